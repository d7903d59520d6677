"""Augmented-Lagrangian solve type (lqr/slove_type: "alm", cpp:47-52, :88-93, :253-261, :581-643,
:665-680, :701-713, :377-378) on the GPU against the oracle (which is pinned to the reference's own
ALM code by tests/golden/alm_*.npz)."""
import numpy as np
import pytest

import cilqr_b200 as cb
from oracle import oracle_py as op
from helpers import perturbed_trajectories, relerr

pytestmark = pytest.mark.gpu


def _alm(pb, **kw):
    for td in pb.templates:
        td.params = dict(td.params, solve_type=1, alm_rho_init=20.0, alm_gamma=0.0, max_rho=20.0, max_mu=120.0, **kw)
    return pb


def test_alm_stage_cost_and_derivs():
    pb = _alm(cb.synthetic_batch("C1", 32, N=30))
    u, x = perturbed_trajectories(pb, seed=11)
    rng = np.random.default_rng(2)
    cols = 8 + 2 * pb.max_obs
    mu = np.abs(rng.normal(0, 3.0, (pb.B, pb.N, cols))) * (rng.random((pb.B, pb.N, cols)) < 0.5)
    rho = np.full(pb.B, 20.0)
    with cb.BatchSolver(pb.templates, pb.B, pb.N, pb.max_obs, "f64") as s:
        J, sc = s.stage_cost(pb, u, x, alm_mu=mu, alm_rho=rho)
        dv = s.stage_derivs(pb, u, x, alm_mu=mu, alm_rho=rho)
    for b in range(pb.B):
        td = pb.templates[0]
        args = (pb.ref_velo[b], pb.n_obs[b], pb.obs[b], pb.borders[b])
        eJ, esc = op.total_cost(td, pb.N, *args, u[b], x[b], alm_mu=mu[b], alm_rho=20.0)
        edv = op.cost_derivs(td, pb.N, *args, u[b], x[b], alm_mu=mu[b], alm_rho=20.0)
        assert relerr(J[b], eJ) < 1e-9 and relerr(sc[b], esc) < 1e-9
        for k in ("lx", "lu", "lxx", "luu"):
            assert relerr(dv[k][b], edv[k]) < 1e-8, k
        assert relerr(dv["mu_next"][b], edv["mu_next"]) < 1e-10


@pytest.mark.parametrize("max_iter", [1, 3])
def test_alm_first_iterations_lockstep(max_iter):
    pb = _alm(cb.synthetic_batch("C1", 128, N=30), max_iter=max_iter)
    ref = op.solve_batch(pb, "f64", trace_cap=4)
    with cb.BatchSolver(pb.templates, pb.B, pb.N, pb.max_obs, "f64") as s:
        s.enable_trace(4)
        out = s.solve(pb)
        st, al, co = s.get_trace(pb.B)
    ok = ((st[:, :max_iter] == ref.tr_status[:, :max_iter]) & (al[:, :max_iter] == ref.tr_alpha[:, :max_iter])).all(axis=1)
    assert ok.mean() >= 0.97
    assert np.array_equal(out.iters[ok], ref.iters[ok])
    assert np.abs(out.x[ok] - ref.x[ok]).max() < 1e-6 and np.abs(out.u[ok] - ref.u[ok]).max() < 1e-6
    assert relerr(out.J[ok], ref.J[ok]) < 1e-6


@pytest.mark.parametrize("name", ["two_straight", "two_borrow"])
def test_alm_template_solve(name):
    scn = cb.get_scenario(name)
    pb = _alm(cb.single_problem(scn, 30))
    td = pb.templates[0]
    r = op.Solver(td.params, 30).solve(td, pb.ref_velo[0], pb.n_obs[0], pb.obs[0], pb.borders[0], pb.x0[0])
    with cb.BatchSolver(pb.templates, 1, 30, pb.max_obs, "f64") as s:
        out = s.solve(pb)
    if out.iters[0] == r.iters:
        assert np.abs(out.x[0] - r.x).max() < 1e-5 and out.status[0] == r.status
    assert out.J[0, 1] <= out.J[0, 0]
