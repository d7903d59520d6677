"""The oracle restatement against the committed golden vectors, which were produced by the
reference's own sources compiled in place (tests/golden/gen_golden.py, oracle/_ref).
Everything is compared bit-for-bit: the restatement keeps the reference's operation order and
both builds use glibc libm without FMA contraction."""
import os

import numpy as np
import pytest

import cilqr_b200 as cb
from oracle import oracle_py as op

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name))


@pytest.mark.parametrize("name", cb.templates.TEMPLATE_ORDER)
@pytest.mark.parametrize("N", [30, 50])
def test_first_solve_matches_reference(name, N):
    g = _load("solve_%s_N%d.npz" % (name, N))
    scn = cb.get_scenario(name)
    pb = cb.single_problem(scn, N)
    td = pb.templates[0]
    # the scenario arrays handed to the reference are the ones we hand to every solver
    assert np.array_equal(g["wx"], td.wx) and np.array_equal(g["wyaw"], td.wyaw)
    assert np.array_equal(g["obs"], pb.obs[0]) and np.array_equal(g["x0"], pb.x0[0])
    s = op.Solver(dict(td.params, use_last_solution=0), N)
    r = s.solve(td, pb.ref_velo[0], pb.n_obs[0], pb.obs[0], pb.borders[0], pb.x0[0])
    assert np.array_equal(r.u, g["u"])
    assert np.array_equal(r.x, g["x"])
    assert r.status == int(g["status"])


def test_warm_start_sequence_matches_reference():
    g = _load("warm_three_straight_N30.npz")
    scn = cb.get_scenario("three_straight")
    s = op.Solver(scn.params, 30)
    x0 = scn.x0.copy()
    for tick in range(6):
        pb = cb.single_problem(scn, 30, tick=tick, x0=x0)
        r = s.solve(pb.templates[0], pb.ref_velo[0], pb.n_obs[0], pb.obs[0], pb.borders[0], pb.x0[0])
        assert np.array_equal(r.u, g["u"][tick]), tick
        assert np.array_equal(r.x, g["x"][tick]), tick
        assert r.status == g["status"][tick]
        x0 = r.x[1].copy()


@pytest.mark.parametrize("name", ["two_straight", "two_borrow"])
def test_alm_solve_matches_reference(name):
    g = _load("alm_%s_N30.npz" % name)
    scn = cb.get_scenario(name)
    pb = cb.single_problem(scn, 30)
    td = pb.templates[0]
    s = op.Solver(dict(td.params, solve_type=1, use_last_solution=0), 30)
    r = s.solve(td, pb.ref_velo[0], pb.n_obs[0], pb.obs[0], pb.borders[0], pb.x0[0])
    assert np.array_equal(r.u, g["u"]) and np.array_equal(r.x, g["x"])
    assert r.status == int(g["status"])


def test_stages_match_reference():
    g = _load("stages_C3_B8_N50.npz")
    pb = cb.synthetic_batch("C3", 8, N=50)
    for b in range(pb.B):
        td = pb.templates[pb.tmpl[b]]
        args = (pb.ref_velo[b], pb.n_obs[b], pb.obs[b], pb.borders[b])
        u, x = g["u"][b], g["x"][b]
        J, _ = op.total_cost(td, pb.N, *args, u, x)
        assert J == g["J"][b]
        dv = op.cost_derivs(td, pb.N, *args, u, x)
        for k in ("lx", "lu", "lxx", "luu"):
            assert np.array_equal(dv[k], g[k][b]), k
        A, Bm = op.dyn_derivs(td.params, pb.N, u, x)
        assert np.array_equal(A, g["A"][b]) and np.array_equal(Bm, g["B"][b])
        d, K, dV, st = op.riccati(pb.N, dv["lx"], dv["lu"], dv["lxx"], dv["luu"], A, Bm, 0.5)
        assert np.array_equal(d, g["d"][b]) and np.array_equal(K, g["K"][b]) and np.array_equal(dV, g["dV"][b])
        assert st == g["status"][b]
        idx = op.ref_match(td.wx, td.wy, x)
        assert np.array_equal(np.stack([td.wx[idx], td.wy[idx], td.wyaw[idx]], 1), g["ref_pts"][b])
        nu, nx = op.forward(td.params, pb.N, u, x, d, K, 0.25)
        assert np.array_equal(nu, g["fwd_u"][b]) and np.array_equal(nx, g["fwd_x"][b])


def test_reference_lines_match_reference_spline():
    """Host-side scenario prep (numpy spline) against the reference's CubicSpline2D/ReferenceLine."""
    g = _load("reference_lines.npz")
    for name in cb.templates.TEMPLATE_ORDER:
        ref = cb.get_scenario(name).ref
        assert len(ref.x) == len(g[name + "_x"])
        assert np.abs(ref.x - g[name + "_x"]).max() < 1e-12
        assert np.abs(ref.y - g[name + "_y"]).max() < 1e-12
        assert np.abs(ref.yaw - g[name + "_yaw"]).max() < 1e-12
