"""Solve-level parity through the C ABI: trajectories, costs, gains, status and iteration counts
against the CPU oracle, plus size-independent properties at the benchmark batch size."""
import numpy as np
import pytest

import cilqr_b200 as cb
from oracle import oracle_py as op
from helpers import compare_traces, relerr, rollout

pytestmark = pytest.mark.gpu


def _compare(solver, out, ref, tol, min_agree, cap=100):
    """The problem is chaotic for a few percent of instances: the oracle itself changes the decision
    trace of ~5 % of C1 instances when only FMA contraction is switched on (parity).  So:
    (1) most instances must reproduce the oracle's whole decision trace, and on those the trajectory,
    costs and gains must match to the north-star tolerance; (2) every instance must follow the oracle
    in lockstep up to its first differing decision, with matching per-iteration costs."""
    same, prefix, prefix_cost = compare_traces(solver, out, ref, cap)
    assert same.mean() >= min_agree, "decision traces agree on %d/%d only" % (same.sum(), len(same))
    assert prefix_cost < 1e-6, prefix_cost
    assert np.median(prefix / np.maximum(np.minimum(out.iters, ref.iters), 1)) == 1.0
    ex = np.abs(out.x[same] - ref.x[same]).max()
    eu = np.abs(out.u[same] - ref.u[same]).max()
    eJ = relerr(out.J[same], ref.J[same])
    assert ex < tol and eu < tol, (ex, eu)
    assert eJ < tol, eJ
    eK = relerr(out.K[same], ref.K[same])
    ed = relerr(out.d[same], ref.d[same])
    assert eK < tol * 100 and ed < tol * 100, (eK, ed)  # gains amplify through Quu^-1
    assert np.array_equal(out.status[same], ref.status[same])
    return same


@pytest.mark.parametrize("name", cb.templates.TEMPLATE_ORDER)
@pytest.mark.parametrize("N", [30, 50])
def test_templates_first_solve_fp64(name, N):
    scn = cb.get_scenario(name)
    pb = cb.single_problem(scn, N)
    o = op.Solver(scn.params, N)
    r = o.solve(pb.templates[0], pb.ref_velo[0], pb.n_obs[0], pb.obs[0], pb.borders[0], pb.x0[0])
    with cb.BatchSolver(pb.templates, 1, N, pb.max_obs, "f64") as s:
        out = s.solve(pb)
    # the nominal two_straight start sits on waypoints to ~1e-14 m: its border-gradient direction is
    # rounding noise (SURVEY hard part 1), so only the decision-stable templates are held to 1e-6
    if out.iters[0] == r.iters and out.exit_reason[0] == r.exit_reason:
        assert np.abs(out.x[0] - r.x).max() < 1e-6
        assert np.abs(out.u[0] - r.u).max() < 1e-6
        assert relerr(out.J[0], r.J) < 1e-6
    else:
        assert name == "two_straight" or name == "three_bend", (name, out.iters, r.iters)


@pytest.mark.parametrize("cfg", ["C1", "C3"])
def test_batch_fp64(cfg):
    pb = cb.synthetic_batch(cfg, 256, N=50)
    ref = op.solve_batch(pb, "f64", trace_cap=100)
    with cb.BatchSolver(pb.templates, pb.B, pb.N, pb.max_obs, "f64") as s:
        s.enable_trace(100)
        out = s.solve(pb)
        cnt = s.counters()
        _compare(s, out, ref, 1e-6, 0.90)
        assert cnt["total_iters"] == int(out.iters.sum())
        assert cnt["total_trials"] >= cnt["total_iters"] - int((out.status == 2).sum()) * 100
        # the wide line search is an execution strategy only: one alpha per round gives the same bits
        s.set_option(s.OPT_WIDE_SEARCH, 0)
        narrow = s.solve(pb)
        cnt2 = s.counters()
    for f in ("u", "x", "J", "K", "d", "iters", "status", "exit_reason", "step_cost"):
        assert np.array_equal(getattr(out, f), getattr(narrow, f)), f
    assert cnt2["rounds"] >= cnt["rounds"] and cnt2["total_trials"] <= cnt["total_trials"]


def test_batch_fp32_vs_fp32_oracle():
    pb = cb.synthetic_batch("C1", 256, N=50)
    ref = op.solve_batch(pb, "f32")
    with cb.BatchSolver(pb.templates, pb.B, pb.N, pb.max_obs, "f32") as s:
        out = s.solve(pb)
    # fp32 flips discrete decisions far more often (SURVEY hard part 3): parity is asserted on the
    # instances whose decision traces agree, and those must be the majority
    same = (out.iters == ref.iters) & (out.exit_reason == ref.exit_reason)
    assert same.mean() > 0.5
    assert np.abs(out.x[same] - ref.x[same]).max() < 5e-2


def test_fp32_final_trajectories_vs_fp64():
    pb = cb.synthetic_batch("C1", 256, N=50)
    ref = op.solve_batch(pb, "f64")
    with cb.BatchSolver(pb.templates, pb.B, pb.N, pb.max_obs, "f32") as s:
        out = s.solve(pb)
    conv = (out.exit_reason == 1) & (ref.exit_reason == 1)
    assert conv.mean() > 0.5
    # both converged to the 0.01 cost threshold: same local optimum, trajectories close
    err = np.abs(out.x[conv] - ref.x[conv]).max(axis=(1, 2))
    assert np.median(err) < 5e-2


def test_warm_start_sequence():
    """use_last_solution across receding-horizon ticks (template three_straight, cpp:97-102, :144)."""
    scn = cb.get_scenario("three_straight")
    N = 30
    o = op.Solver(scn.params, N)
    x0 = scn.x0.copy()
    with cb.BatchSolver([cb.scenario.template_data(scn)], 1, N, len(scn.ic) - 1, "f64") as s:
        for tick in range(6):
            pb = cb.single_problem(scn, N, tick=tick, x0=x0)
            r = o.solve(pb.templates[0], pb.ref_velo[0], pb.n_obs[0], pb.obs[0], pb.borders[0], pb.x0[0])
            out = s.solve(pb)
            assert out.iters[0] == r.iters, tick
            assert np.abs(out.x[0] - r.x).max() < 1e-6, tick
            assert np.abs(out.u[0] - r.u).max() < 1e-6, tick
            x0 = r.x[1].copy()  # ego_state = new_x.row(1) (motion_planning.cpp:197)


def test_compat_class_and_short_track():
    scn = cb.get_scenario("two_borrow")
    cfg = dict(scn.cfg)
    cfg["lqr/N"] = 50
    sol = cb.CILQRSolver(cfg)
    tracks = [scn.tracks[j] for j in range(scn.tracks.shape[0])]
    u, x = sol.solve(scn.x0, scn.ref, scn.target_velocity, tracks, scn.borders)
    o = op.Solver(scn.params, 50)
    pb = cb.single_problem(scn, 50)
    r = o.solve(pb.templates[0], pb.ref_velo[0], pb.n_obs[0], pb.obs[0], pb.borders[0], pb.x0[0])
    assert u.shape == (50, 2) and x.shape == (51, 4)
    assert np.abs(x - r.x).max() < 1e-6
    with pytest.raises(IndexError):
        sol.solve(scn.x0, scn.ref, scn.target_velocity, [t[:20] for t in tracks], scn.borders)
    sol.close()


def test_properties_at_benchmark_size():
    """B = 4096 (config C1): properties that need no oracle run."""
    pb = cb.synthetic_batch("C1", 4096, N=50)
    with cb.BatchSolver(pb.templates, pb.B, pb.N, pb.max_obs, "f64") as s:
        a = s.solve(pb)
        b = s.solve(pb)  # idempotence: barrier mode without warm start keeps no state
        sub = s.solve(pb.slice(1000, 1500))  # instances are independent: a slice solves identically
    for f in ("u", "x", "J", "K", "d", "iters", "status"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
        assert np.array_equal(getattr(a, f)[1000:1500], getattr(sub, f)), f
    assert np.all(a.J[:, 1] <= a.J[:, 0] * (1 + 1e-12))  # accepted steps only ever decrease the cost
    assert np.allclose(a.step_cost.sum(axis=1), a.J[:, 1], rtol=1e-12)
    assert np.all(a.iters >= 1) and np.all(a.iters <= 100)
    # x is the rollout of u from x0 (spot-check against the oracle's model)
    for i in range(0, 4096, 512):
        assert np.abs(rollout(pb.templates[0].params, pb.N, pb.x0[i], a.u[i]) - a.x[i]).max() < 1e-9


def test_no_obstacles_and_errors():
    pb = cb.synthetic_batch("C1", 32, N=50)
    pb.n_obs[:] = 0
    ref = op.solve_batch(pb, "f64")
    with cb.BatchSolver(pb.templates, pb.B, pb.N, pb.max_obs, "f64") as s:
        out = s.solve(pb)
        same = out.iters == ref.iters
        assert same.mean() > 0.9 and np.abs(out.x[same] - ref.x[same]).max() < 1e-6
        with pytest.raises(cb.CilqrError) as e:
            s.solve(cb.synthetic_batch("C1", 64, N=50))  # more than max_batch
        assert e.value.code == -1
        short = cb.synthetic_batch("C1", 8, N=50)
        short.obs = np.ascontiguousarray(short.obs[:, :, :40])
        with pytest.raises(cb.CilqrError) as e:
            s.solve(short)
        assert e.value.code == -2  # RoutingLine index out of range
