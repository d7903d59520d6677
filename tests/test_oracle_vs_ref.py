"""Restatement vs the reference's own sources (oracle/_ref) on fresh random instances.
Runs only where oracle/_ref was built (this container); the golden-vector tests cover the rest."""
import numpy as np
import pytest

import cilqr_b200 as cb
from oracle import oracle_py as op, ref_py as rp
from helpers import perturbed_trajectories

pytestmark = pytest.mark.skipif(not rp.available(), reason="oracle/_ref not built (no reference tree)")


@pytest.mark.parametrize("cfg,seed", [("C1", 11), ("C3", 12), ("C3", 13)])
def test_full_solves_bit_identical(cfg, seed):
    pb = cb.synthetic_batch(cfg, 24, N=50, seed=seed)
    for b in range(pb.B):
        td = pb.templates[pb.tmpl[b]]
        args = (td, pb.ref_velo[b], pb.n_obs[b], pb.obs[b], pb.borders[b])
        ru, rx, rst = rp.RefSolver(td.params, pb.N).solve(*args, pb.x0[b])
        o = op.Solver(td.params, pb.N).solve(*args, pb.x0[b])
        assert np.array_equal(ru, o.u) and np.array_equal(rx, o.x) and rst == o.status, b


def test_stage_outputs_bit_identical():
    pb = cb.synthetic_batch("C3", 16, N=40, seed=3)
    u, x = perturbed_trajectories(pb, seed=8)
    for b in range(pb.B):
        td = pb.templates[pb.tmpl[b]]
        args = (pb.ref_velo[b], pb.n_obs[b], pb.obs[b], pb.borders[b])
        rs = rp.RefSolver(td.params, pb.N)
        lamb = [0.0, 1.0, 64.0][b % 3]
        r = rs.backward_pass(td, *args, u[b], x[b], lamb)
        dv = op.cost_derivs(td, pb.N, *args, u[b], x[b])
        A, Bm = op.dyn_derivs(td.params, pb.N, u[b], x[b])
        d, K, dV, st = op.riccati(pb.N, dv["lx"], dv["lu"], dv["lxx"], dv["luu"], A, Bm, lamb)
        assert rs.total_cost(td, *args, u[b], x[b]) == op.total_cost(td, pb.N, *args, u[b], x[b])[0]
        for k in ("lx", "lu", "lxx", "luu"):
            assert np.array_equal(r[k], dv[k]), k
        assert np.array_equal(r["A"].reshape(pb.N, 4, 4), A) and np.array_equal(r["B"].reshape(pb.N, 4, 2), Bm)
        assert np.array_equal(r["d"], d) and np.array_equal(r["K"], K) and np.array_equal(r["dV"], dV)
        assert r["status"] == st


def test_alm_and_short_track():
    scn = cb.get_scenario("two_borrow")
    pb = cb.single_problem(scn, 30)
    td = pb.templates[0]
    p = dict(td.params, solve_type=1)
    args = (td, pb.ref_velo[0], pb.n_obs[0], pb.obs[0], pb.borders[0])
    ru, rx, rst = rp.RefSolver(p, 30).solve(*args, pb.x0[0])
    o = op.Solver(p, 30).solve(*args, pb.x0[0])
    assert np.array_equal(ru, o.u) and np.array_equal(rx, o.x) and rst == o.status
    short = (td, pb.ref_velo[0], pb.n_obs[0], pb.obs[0][:, :20], pb.borders[0])
    with pytest.raises(IndexError):
        rp.RefSolver(td.params, 30).solve(*short, pb.x0[0])
    with pytest.raises(IndexError):
        op.Solver(td.params, 30).solve(*short, pb.x0[0])
