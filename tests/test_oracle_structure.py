"""Analytic self-checks of the oracle: the derivative stage against finite differences
of the cost stage, the model Jacobians against the model, and the Riccati recursion on an LQ problem."""
import numpy as np
import pytest

import cilqr_b200 as cb
from oracle import oracle_py as op
from helpers import perturbed_trajectories


def _instance(cfg="C3", b=1, N=20):
    pb = cb.synthetic_batch(cfg, 8, N=N)
    u, x = perturbed_trajectories(pb, seed=2)
    td = pb.templates[pb.tmpl[b]]
    return pb, td, (pb.ref_velo[b], pb.n_obs[b], pb.obs[b], pb.borders[b]), u[b], x[b]


@pytest.mark.parametrize("b", [0, 1, 2, 3])
def test_cost_gradient_matches_finite_differences(b):
    """l_x, l_u are the exact gradient of get_total_cost w.r.t. independent (x, u) entries wherever the
    matched waypoint does not change; l_xx omits the constraints' own curvature (Gauss-Newton), so only
    gradients are checked."""
    pb, td, args, u, x = _instance(b=b)
    dv = op.cost_derivs(td, pb.N, *args, u, x)
    idx0 = op.ref_match(td.wx, td.wy, x)
    h = 1e-6
    for (k, c) in [(3, 0), (7, 1), (11, 2), (15, 3), (pb.N, 1), (pb.N, 3)]:
        xp, xm = x.copy(), x.copy()
        xp[k, c] += h
        xm[k, c] -= h
        if not (np.array_equal(op.ref_match(td.wx, td.wy, xp), idx0) and np.array_equal(op.ref_match(td.wx, td.wy, xm), idx0)):
            continue
        fd = (op.total_cost(td, pb.N, *args, u, xp)[0] - op.total_cost(td, pb.N, *args, u, xm)[0]) / (2 * h)
        assert abs(fd - dv["lx"][k, c]) < 1e-4 * max(1.0, abs(fd)), (k, c, fd, dv["lx"][k, c])
    for (k, c) in [(0, 0), (4, 1), (pb.N - 1, 0), (pb.N - 1, 1)]:
        up, um = u.copy(), u.copy()
        up[k, c] += h
        um[k, c] -= h
        fd = (op.total_cost(td, pb.N, *args, up, x)[0] - op.total_cost(td, pb.N, *args, um, x)[0]) / (2 * h)
        assert abs(fd - dv["lu"][k, c]) < 1e-4 * max(1.0, abs(fd)), (k, c)
    # structure: first state row carries the tracking term only; l_xx symmetric; l_uu diagonal
    Q = np.diag([td.params["w_pos"], td.params["w_pos"], td.params["w_vel"], td.params["w_yaw"]])
    assert np.array_equal(dv["lxx"][0], 2 * Q)
    assert np.abs(dv["lxx"] - dv["lxx"].transpose(0, 2, 1)).max() == 0.0
    assert np.all(dv["luu"][:, 0, 1] == 0) and np.all(dv["luu"][:, 1, 0] == 0)


def test_model_jacobians():
    """Rear-centre mode: A, B are the exact Jacobians.  Gravity-centre mode reproduces the reference's
    beta mismatch: B's steering column is NOT the true derivative; assert the formula."""
    h = 1e-6
    for name in ("two_straight", "two_borrow"):
        p = cb.get_scenario(name).params
        x = np.array([1.0, 2.0, 6.0, 0.3])
        u = np.array([0.7, 0.08])
        A, Bm = op.dyn_derivs(p, 1, u[None], np.stack([x, x]))
        fdA = np.zeros((4, 4))
        for c in range(4):
            e = np.zeros(4); e[c] = h
            fdA[:, c] = (op.propagate(p, x + e, u) - op.propagate(p, x - e, u)) / (2 * h)
        fdB = np.zeros((4, 2))
        for c in range(2):
            e = np.zeros(2); e[c] = h
            fdB[:, c] = (op.propagate(p, x, u + e) - op.propagate(p, x, u - e)) / (2 * h)
        if p["reference_point"] == 0:
            assert np.abs(A[0] - fdA).max() < 1e-8 and np.abs(Bm[0] - fdB).max() < 1e-8
        else:
            assert np.abs(A[0][:, :2] - fdA[:, :2]).max() < 1e-8
            beta_j = np.arctan(np.tan(u[1] / 2))  # the Jacobian's beta
            assert A[0][0, 2] == np.cos(beta_j + x[3]) * p["dt"]
            assert abs(A[0][0, 2] - fdA[0, 2]) > 1e-9  # differs from the true derivative, by design
            assert np.abs(Bm[0][2] - fdB[2]).max() < 1e-9


def test_riccati_on_lq_problem():
    """On a linear-quadratic problem the backward pass gives the optimal gains: one full step reaches
    the analytic optimum, and dV predicts the cost decrease exactly."""
    rng = np.random.default_rng(0)
    N = 12
    A = np.tile(np.eye(4), (N, 1, 1)); A[:, 0, 2] = 0.1; A[:, 1, 3] = 0.1
    Bm = np.zeros((N, 4, 2)); Bm[:, 2, 0] = 0.1; Bm[:, 3, 1] = 0.1
    Qm, Rm = np.diag([1.0, 2.0, 0.5, 0.3]), np.diag([0.4, 0.7])
    x = np.zeros((N + 1, 4)); x[0] = rng.normal(size=4)
    u = rng.normal(size=(N, 2)) * 0.1
    for k in range(N):
        x[k + 1] = A[k] @ x[k] + Bm[k] @ u[k]
    cost = lambda xx, uu: sum(xx[k] @ Qm @ xx[k] for k in range(N + 1)) + sum(uu[k] @ Rm @ uu[k] for k in range(N))
    lx, lu = 2 * x @ Qm, 2 * u @ Rm
    lxx, luu = np.tile(2 * Qm, (N + 1, 1, 1)), np.tile(2 * Rm, (N, 1, 1))
    d, K, dV, st = op.riccati(N, lx, lu, lxx, luu, A, Bm, 0.0)
    assert st == 0
    nx, nu = x.copy(), u.copy()
    for k in range(N):
        nu[k] = u[k] + K[k] @ (nx[k] - x[k]) + d[k]
        nx[k + 1] = A[k] @ nx[k] + Bm[k] @ nu[k]
    assert abs((cost(x, u) - cost(nx, nu)) - (-(dV[0] + dV[1]))) < 1e-9
    d2, K2, dV2, _ = op.riccati(N, 2 * nx @ Qm, 2 * nu @ Rm, lxx, luu, A, Bm, 0.0)
    assert np.abs(d2).max() < 1e-9  # already optimal
    # non-PD control Hessian -> BACKWARD_PASS_FAIL, rows below the failing step stay zero
    luu_bad = luu.copy(); luu_bad[5] = -np.eye(2) * 100
    d3, K3, _, st3 = op.riccati(N, lx, lu, lxx, luu_bad, A, Bm, 0.0)
    assert st3 == 2 and np.all(d3[:6] == 0) and np.all(K3[:6] == 0) and np.any(d3[6:] != 0)
    # NaN passes the LLT test exactly like Eigen's (comparison with NaN is false)
    luu_nan = luu.copy(); luu_nan[5, 0, 0] = np.nan
    assert op.riccati(N, lx, lu, lxx, luu_nan, A, Bm, 0.0)[3] == 0


def test_ref_match_first_local_minimum():
    wx = np.arange(0, 10, 0.1); wy = np.zeros_like(wx)
    x = np.zeros((4, 4)); x[:, 0] = [2.04, 2.06, 1.0, 9.95]
    idx = op.ref_match(wx, wy, x)
    # never moves backwards (third point is behind the second match), clamps at the last waypoint
    assert list(idx) == [20, 21, 21, 99]
    x[:, 0] = [np.nan, 3.0, 3.0, 3.0]
    assert op.ref_match(wx, wy, x)[0] == 0  # NaN distance stops the scan at the start index


def test_status_machine_quirks():
    scn = cb.get_scenario("two_borrow")
    pb = cb.single_problem(scn, 30)
    td = pb.templates[0]
    r = op.Solver(td.params, 30).solve(td, pb.ref_velo[0], pb.n_obs[0], pb.obs[0], pb.borders[0], pb.x0[0])
    tr = r.trace
    assert r.status == 1 and r.exit_reason == 1
    # the converging step is discarded: effective == 0 on the last row (cpp:358-361), so the returned
    # trajectory's cost is the cost *before* that step
    assert tr[-1, 2] == 0 and tr[-1, 0] == 1
    assert r.J[1] == tr[-1, 3]
    # lambda: init 0 stays 0 through successes (0 * decay)
    assert np.all(tr[:, 5] == 0)
