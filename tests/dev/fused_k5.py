"""The fused flavour of the backward pass (control half of the records computed in the kernel) against the plain
register-prefetch flavour: same gains?  time per launch on resident data (L2 flushed).  Development experiment."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cilqr_b200 as cb
Br = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
for dtype in ("f64", "f32"):
    seed = cb.synthetic_batch("C1", 4096, N=50)
    with cb.BatchSolver(seed.templates, Br, 50, seed.max_obs, dtype) as rs:
        u0, x0 = rs.stage_init(seed.x0, seed.tmpl)
        rs.stage_derivs(seed, u0, x0)
        rs.bench_tile_records(4096, Br)
        res = {}
        for variant in (1, 3, 1, 3):
            rs.set_option(rs.OPT_BENCH_PREFETCH, variant)
            ms, nbytes = rs.bench_backward(Br, 0.0, 8, True)
            out = rs.download(4096)
            res[variant] = out
            print("B=%d %s variant %d: %.3f ms/launch (median of 8), %.0f GB/s on the plain kernel's byte count" % (
                Br, dtype, variant, np.median(ms[2:]), nbytes / np.median(ms[2:]) / 1e6), flush=True)
        same = np.array_equal(res[1].K, res[3].K, equal_nan=True) and np.array_equal(res[1].d, res[3].d, equal_nan=True)
        print("   same K, d bits:", same, "" if same else "max |dK| %.3e" % np.nanmax(np.abs(res[1].K - res[3].K)))
