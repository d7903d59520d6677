"""Roofline leg (K5 alone on resident records, L2 flushed) for the three backward kernels:
python tests/dev/backward_variants.py  -> GB/s per (dtype, batch, variant 0 streaming / 1 prefetch / 2 staged)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cilqr_b200 as cb
seed = cb.synthetic_batch("C1", 4096, N=50)
for dtype in ("f64", "f32"):
    for Br in (65536, 262144, 1048576):
        rs = cb.BatchSolver(seed.templates, Br, 50, seed.max_obs, dtype)
        u0, x0 = rs.stage_init(seed.x0, seed.tmpl)
        rs.stage_derivs(seed, u0, x0)
        rs.bench_tile_records(4096, Br)
        for variant in (0, 1, 2):
            rs.set_option(rs.OPT_BENCH_PREFETCH, variant)
            rs.bench_backward(Br, 0.0, 3, True)
            ms, nbytes = rs.bench_backward(Br, 0.0, 12, True)
            print("%s B=%d variant %d: %.4f ms -> %.0f GB/s" % (dtype, Br, variant, np.mean(ms), nbytes / np.mean(ms) / 1e6), flush=True)
        rs.close()
