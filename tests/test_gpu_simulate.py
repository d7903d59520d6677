"""Receding-horizon closed loop on the device (cilqr_b200_simulate) against the same loop driven
through the oracle: solve at tick t on the tracks from t on, apply x.row(1), warm start carried
(src/motion_planning.cpp:180-197, src/utils.cpp:88-103)."""
import numpy as np
import pytest

import cilqr_b200 as cb
from oracle import oracle_py as op

pytestmark = pytest.mark.gpu


def _oracle_loop(scn, N, ticks, x0, dtype="f64"):
    """The reference's loop (motion_planning.cpp:180-197): t accumulates in floating point and the obstacle
    window starts at size_t(t / delta_t) — 0,1,2,3,4,5,5,6,... for delta_t = 0.1."""
    o = op.Solver(scn.params, N, dtype)
    ego, iters, t = [np.array(x0, dtype=np.float64)], [], 0.0
    for _ in range(ticks):
        index = int(t / scn.dt)
        pb = cb.single_problem(scn, N, tick=index, x0=ego[-1])
        r = o.solve(pb.templates[0], pb.ref_velo[0], pb.n_obs[0], pb.obs[0], pb.borders[0], pb.x0[0])
        ego.append(r.x[1].copy())
        iters.append(r.iters)
        t += scn.dt
    return np.array(ego), np.array(iters)


def _setup(scn, B):
    rng = np.random.default_rng(5)
    x0 = np.tile(scn.x0, (B, 1))
    x0[1:, 1] += rng.uniform(-0.3, 0.3, B - 1)   # instance 0 is the YAML scenario itself
    x0[1:, 2] += rng.uniform(-1.0, 1.0, B - 1)
    tracks = np.tile(scn.tracks[None], (B, 1, 1, 1))
    return x0, tracks, scn.tracks.shape[0]


@pytest.mark.parametrize("name", ["two_straight", "two_borrow", "three_straight", "three_bend"])
def test_closed_loop_parity_build_is_bit_exact(name):
    """16 ticks (past the repeated index 5 of the reference's t / delta_t sequence) of the receding-horizon loop on the
    device, warm start carried where the scenario asks for it: the parity build returns the oracle's ego states and
    iteration counts bit for bit, for every instance and tick."""
    scn = cb.get_scenario(name)
    N, ticks, B = 30, 16, 6
    x0, tracks, nobs = _setup(scn, B)
    with cb.BatchSolver([cb.scenario.template_data(scn)], B, N, nobs, "f64", flavour="parity") as s:
        ego, iters, status = s.simulate(x0, np.full(B, scn.target_velocity), np.tile(scn.borders, (B, 1)),
                                        np.zeros(B, np.int32), np.full(B, nobs, np.int32), tracks, ticks)
    for b in range(B):
        rego, riters = _oracle_loop(scn, N, ticks, x0[b], "f64pm")
        assert np.array_equal(iters[b], riters), (b, iters[b], riters)
        assert np.array_equal(ego[b], rego), (b, np.abs(ego[b] - rego).max())


@pytest.mark.parametrize("name", ["two_borrow", "three_straight", "three_bend"])
def test_closed_loop_matches_oracle(name):
    """The default build: every tick up to an instance's first decision that differs from the oracle's is within 1e-6
    (ticks before it start from states that agree to rounding), and the first tick — a plain first solve — always."""
    scn = cb.get_scenario(name)
    N, ticks, B = 30, 16, 6
    x0, tracks, nobs = _setup(scn, B)
    with cb.BatchSolver([cb.scenario.template_data(scn)], B, N, nobs, "f64") as s:
        ego, iters, status = s.simulate(x0, np.full(B, scn.target_velocity), np.tile(scn.borders, (B, 1)),
                                        np.zeros(B, np.int32), np.full(B, nobs, np.int32), tracks, ticks)
        with pytest.raises(cb.CilqrError) as e:   # tracks too short for the last tick
            s.simulate(x0, np.full(B, scn.target_velocity), np.tile(scn.borders, (B, 1)), np.zeros(B, np.int32),
                       np.full(B, nobs, np.int32), tracks[:, :, : N + 5], ticks)
        assert e.value.code == -2
    assert np.array_equal(ego[:, 0], x0)
    agree = []
    for b in range(B):
        rego, riters = _oracle_loop(scn, N, ticks, x0[b])
        k = 0
        while k < ticks and iters[b, k] == riters[k]:
            k += 1
        agree.append(k)
        assert k >= 1, (b, iters[b], riters)
        assert np.abs(ego[b, : k + 1] - rego[: k + 1]).max() < 1e-6, (b, k, np.abs(ego[b, : k + 1] - rego[: k + 1]).max())
    print("%s: ticks in agreement with the oracle per instance: %s of %d" % (name, agree, ticks))
