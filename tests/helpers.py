"""Shared helpers for the parity tests (oracle = checker, never the thing under test)."""
import numpy as np

import cilqr_b200 as cb
from oracle import oracle_py as op


def rollout(params, N, x0, u, dtype="f64"):
    """x[k+1] = f(x[k], u[k]) with the oracle's propagate (forward pass with K = 0, d = 0)."""
    x = np.zeros((N + 1, 4))
    x[0] = x0
    nu, nx = op.forward(params, N, u, x, np.zeros((N, 2)), np.zeros((N, 2, 4)), 0.0, dtype)
    return nx


def perturbed_trajectories(pb, seed=0, scale=(0.8, 0.03)):
    """A trajectory per instance: smooth random controls rolled out from x0."""
    rng = np.random.default_rng(seed)
    B, N = pb.B, pb.N
    u = np.zeros((B, N, 2))
    x = np.zeros((B, N + 1, 4))
    for b in range(B):
        a = np.cumsum(rng.normal(0, 0.25, N)) * scale[0] / 2
        s = np.cumsum(rng.normal(0, 0.25, N)) * scale[1] / 2
        u[b, :, 0] = np.clip(a, -2.5, 2.5)
        u[b, :, 1] = np.clip(s, -0.1, 0.1)
        x[b] = rollout(pb.templates[pb.tmpl[b]].params, N, pb.x0[b], u[b])
    return u, x


def oracle_stage(pb, b, u, x, dtype="f64"):
    td = pb.templates[pb.tmpl[b]]
    J, sc = op.total_cost(td, pb.N, pb.ref_velo[b], pb.n_obs[b], pb.obs[b], pb.borders[b], u, x, dtype)
    dv = op.cost_derivs(td, pb.N, pb.ref_velo[b], pb.n_obs[b], pb.obs[b], pb.borders[b], u, x, dtype)
    A, Bm = op.dyn_derivs(td.params, pb.N, u, x, dtype)
    idx = op.ref_match(td.wx, td.wy, x, dtype)
    return J, sc, dv, A, Bm, idx


def relerr(a, b, floor=1.0, loose_nonfinite=False):
    """max |a - b| / max(|b|, floor); entries where both sides hold the same inf (an overflowed barrier term) or
    both NaN count as equal, a non-finite value on one side only as an infinite error.  loose_nonfinite: an entry
    that overflowed on both sides counts as equal whether it ended up inf or NaN (fp32 with several overflowed
    terms: inf - inf versus inf + NaN depends on the order the terms meet)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if not a.size:
        return 0.0
    fin = np.isfinite(a) & np.isfinite(b)
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    if loose_nonfinite:
        same |= ~np.isfinite(a) & ~np.isfinite(b)
    if np.any(~fin & ~same):
        return float("inf")
    with np.errstate(invalid="ignore"):
        e = np.where(fin, np.abs(a - b) / np.maximum(np.abs(b), floor), 0.0)
    return float(np.max(e))


def compare_traces(gpu_solver, out, ref, cap):
    """Lockstep comparison of per-iteration decisions (status, accepted alpha) of a GPU solve
    against the oracle's.  Returns (same, prefix_len): `same[b]` is True when the whole decision
    sequence agrees; prefix_len[b] is the number of leading iterations that agree."""
    st, al, co = gpu_solver.get_trace(len(out.iters))
    B = len(out.iters)
    same = np.zeros(B, bool)
    prefix = np.zeros(B, int)
    worst_prefix_cost = 0.0
    for b in range(B):
        n = min(out.iters[b], ref.iters[b], cap)
        eq = (st[b, :n] == ref.tr_status[b, :n]) & (al[b, :n] == ref.tr_alpha[b, :n])
        k = n if eq.all() else int(np.argmin(eq))
        prefix[b] = k
        same[b] = (k == n) and out.iters[b] == ref.iters[b] and out.exit_reason[b] == ref.exit_reason[b]
        if k > 0:
            worst_prefix_cost = max(worst_prefix_cost, relerr(co[b, :k], ref.tr_cost[b, :k]))
    return same, prefix, worst_prefix_cost
