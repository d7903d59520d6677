import numpy as np

import cilqr_b200 as cb


def test_counter_rng_is_sliceable_and_uniform():
    a = cb.synthetic_batch("C3", 64, N=20)
    b = cb.synthetic_batch("C3", 16, N=20, first_id=32)
    for f in ("x0", "ref_velo", "borders", "tmpl", "n_obs", "obs"):
        assert np.array_equal(getattr(a, f)[32:48], getattr(b, f)), f
    u = cb.scenario.u01(1, np.arange(200000), 3)
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 5e-3 and abs(u.var() - 1 / 12) < 2e-3
    assert not np.array_equal(cb.scenario.u01(1, np.arange(8), 3), cb.scenario.u01(2, np.arange(8), 3))


def test_synthetic_configs():
    for cfg, N, nobs in (("C1", 50, 3), ("C2", 100, 3), ("C4", 200, 5)):
        pb = cb.synthetic_batch(cfg, 32)
        assert pb.N == N and pb.obs.shape == (32, nobs, N + 1, 3)
        assert np.all(pb.n_obs == nobs) and np.all(pb.tmpl == 0)
        assert np.all(np.isfinite(pb.obs)) and np.all(np.isfinite(pb.x0))
    pb = cb.synthetic_batch("C3", 32)
    assert list(pb.n_obs[:4]) == [3, 4, 8, 3] and len(pb.templates) == 4
    assert all(td.params["use_last_solution"] == 0 for td in pb.templates)
    s = pb.slice(4, 12)
    assert s.B == 8 and np.array_equal(s.x0, pb.x0[4:12])
