"""Edge cases the reference's semantics define: empty and ragged batches, no obstacles at all,
non-finite inputs (NaN comparisons are false, so every line search fails and lambda climbs to its
ceiling, cpp:118-133), maximum waypoint count."""
import numpy as np
import pytest

import cilqr_b200 as cb
from oracle import oracle_py as op

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B", [1, 3, 33, 129])
def test_ragged_batch_sizes(B):
    full = cb.synthetic_batch("C1", 160, N=30)
    for td in full.templates:
        td.params = dict(td.params, max_iter=3)
    pb = full.slice(7, 7 + B)
    ref = op.solve_batch(pb, "f64")
    with cb.BatchSolver(pb.templates, 160, pb.N, pb.max_obs, "f64") as s:
        out = s.solve(pb)
        whole = s.solve(full)
    assert np.array_equal(out.iters, ref.iters)
    assert np.abs(out.x - ref.x).max() < 1e-6
    assert np.array_equal(out.x, whole.x[7:7 + B])  # position in the batch does not matter


def test_empty_batch_and_zero_obstacle_capacity():
    pb = cb.synthetic_batch("C1", 16, N=30)
    pb.n_obs[:] = 0
    nob = cb.BatchProblem(pb.templates, pb.N, pb.x0, pb.ref_velo, pb.borders, pb.tmpl, pb.n_obs,
                          np.zeros((pb.B, 0, pb.N + 1, 3)))
    ref = op.solve_batch(pb, "f64")
    with cb.BatchSolver(pb.templates, 16, pb.N, 0, "f64") as s:   # max_obs = 0
        out = s.solve(nob)
        empty = s.solve(pb.slice(0, 0))
        assert empty.x.shape[0] == 0
    same = out.iters == ref.iters
    assert same.mean() > 0.8 and np.abs(out.x[same] - ref.x[same]).max() < 1e-6


def test_non_finite_inputs_terminate_like_the_reference():
    pb = cb.synthetic_batch("C1", 8, N=30)
    pb.x0[2, 1] = np.nan          # NaN lateral position
    pb.x0[5, 2] = np.inf          # infinite speed
    ref = op.solve_batch(pb, "f64")
    with cb.BatchSolver(pb.templates, 8, pb.N, pb.max_obs, "f64") as s:
        out = s.solve(pb)
    for b in (2, 5):
        assert out.iters[b] == ref.iters[b] and out.exit_reason[b] == ref.exit_reason[b] == 2  # MAX_LAMB
        assert out.status[b] == ref.status[b]
    ok = [b for b in range(8) if b not in (2, 5)]
    assert np.array_equal(out.iters[ok], ref.iters[ok])


def test_long_reference_line_and_template_errors():
    scn = cb.get_scenario("two_straight")
    td = cb.scenario.template_data(scn)
    M = 65535
    td.wx = np.linspace(-10.0, -10.0 + 0.1 * (M - 1), M)
    td.wy = np.zeros(M)
    td.wyaw = np.zeros(M)
    pb = cb.single_problem(scn, 30)
    pb.templates = [td]
    r = op.Solver(td.params, 30).solve(td, pb.ref_velo[0], pb.n_obs[0], pb.obs[0], pb.borders[0], pb.x0[0])
    with cb.BatchSolver([td], 1, 30, pb.max_obs, "f64") as s:
        out = s.solve(pb)
        assert out.iters[0] == r.iters and np.abs(out.x[0] - r.x).max() < 1e-6
        with pytest.raises(cb.CilqrError) as e:   # the reference indexes waypoints with uint16_t (cpp:291)
            s.set_template(0, None, np.zeros(65536), np.zeros(65536), np.zeros(65536))
        assert e.value.code == -1
        bad = cb.single_problem(scn, 30)
        bad.tmpl[:] = 3                             # template never set
        with pytest.raises(cb.CilqrError):
            s.solve(bad)


@pytest.mark.parametrize("max_iter", [0, 1, 2, 7])
def test_lookahead_rounds_at_the_limits(max_iter):
    """Look-ahead rounds (the default for batches this small) where the job machinery has little or nothing to
    adopt: no iteration at all, a single one, a few; one instance and a handful; NaN / inf instances whose every
    line search fails (the all-rejected job chain up to max_lamb) next to ordinary ones.  Same results as the
    sequential rounds and as the oracle."""
    pb = cb.synthetic_batch("C3", 21, N=30)
    for td in pb.templates:
        td.params = dict(td.params, max_iter=max_iter)
    pb.x0[4, 1] = np.nan
    pb.x0[9, 2] = np.inf
    ref = op.solve_batch(pb, "f64")
    with cb.BatchSolver(pb.templates, 21, pb.N, pb.max_obs, "f64") as s:
        for bound in (0, 1):
            s.set_option(s.OPT_LOOKAHEAD, bound)
            outs = [s.solve(pb), s.solve(pb.slice(3, 4))]
            if bound == 0:
                seq = outs
            assert np.array_equal(outs[0].iters, ref.iters)
            assert np.array_equal(outs[0].exit_reason, ref.exit_reason)
            ok = np.isfinite(ref.x).all(axis=(1, 2))
            assert np.abs(outs[0].x[ok] - ref.x[ok]).max() < 1e-6
            assert np.array_equal(outs[1].x, outs[0].x[3:4], equal_nan=True)
        for a, b in zip(seq, outs):
            for f in ("u", "x", "J", "K", "d", "iters", "status", "exit_reason", "step_cost"):
                assert np.array_equal(getattr(a, f), getattr(b, f), equal_nan=True), f
