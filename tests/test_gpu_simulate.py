"""Receding-horizon closed loop on the device (cilqr_b200_simulate) against the same loop driven
through the oracle: solve at tick t on the tracks from t on, apply x.row(1), warm start carried
(src/motion_planning.cpp:180-197, src/utils.cpp:88-103)."""
import numpy as np
import pytest

import cilqr_b200 as cb
from oracle import oracle_py as op

pytestmark = pytest.mark.gpu


def _oracle_loop(scn, N, ticks, x0):
    o = op.Solver(scn.params, N)
    ego, iters = [np.array(x0, dtype=np.float64)], []
    for t in range(ticks):
        pb = cb.single_problem(scn, N, tick=t, x0=ego[-1])
        r = o.solve(pb.templates[0], pb.ref_velo[0], pb.n_obs[0], pb.obs[0], pb.borders[0], pb.x0[0])
        ego.append(r.x[1].copy())
        iters.append(r.iters)
    return np.array(ego), np.array(iters)


@pytest.mark.parametrize("name", ["two_borrow", "three_straight", "three_bend"])
def test_closed_loop_matches_oracle(name):
    scn = cb.get_scenario(name)
    N, ticks = 30, 8
    B = 6
    rng = np.random.default_rng(5)
    x0 = np.tile(scn.x0, (B, 1))
    x0[1:, 1] += rng.uniform(-0.3, 0.3, B - 1)   # instance 0 is the YAML scenario itself
    x0[1:, 2] += rng.uniform(-1.0, 1.0, B - 1)
    tracks = np.tile(scn.tracks[None], (B, 1, 1, 1))
    nobs = scn.tracks.shape[0]
    with cb.BatchSolver([cb.scenario.template_data(scn)], B, N, nobs, "f64") as s:
        ego, iters, status = s.simulate(x0, np.full(B, scn.target_velocity), np.tile(scn.borders, (B, 1)),
                                        np.zeros(B, np.int32), np.full(B, nobs, np.int32), tracks, ticks)
        with pytest.raises(cb.CilqrError) as e:   # tracks too short for the last tick
            s.simulate(x0, np.full(B, scn.target_velocity), np.tile(scn.borders, (B, 1)), np.zeros(B, np.int32),
                       np.full(B, nobs, np.int32), tracks[:, :, : ticks + N - 1], ticks)
        assert e.value.code == -2
    assert np.array_equal(ego[:, 0], x0)
    n_same = 0
    for b in range(B):
        rego, riters = _oracle_loop(scn, N, ticks, x0[b])
        same = np.array_equal(iters[b], riters)
        n_same += same
        if same:
            assert np.abs(ego[b] - rego).max() < 1e-5, (b, np.abs(ego[b] - rego).max())
        # the first tick is a plain first solve: always comparable
        if iters[b, 0] == riters[0]:
            assert np.abs(ego[b, 1] - rego[1]).max() < 1e-6
    assert n_same >= B // 2
