"""development: K5 (backward-pass kernel alone) GB/s over dtype x shape x batch x kernel variant"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, cilqr_b200 as cb
lamb = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
for cfg, N in (("C1", 50), ("C4", 200)):
    seed = cb.synthetic_batch(cfg, 4096, N=N)
    for dtype in ("f64", "f32"):
        for Br in (65536, 262144):
            rs = cb.BatchSolver(seed.templates, Br, N, seed.max_obs, dtype)
            u0, x0 = rs.stage_init(seed.x0, seed.tmpl)
            rs.stage_derivs(seed, u0, x0)
            rs.bench_tile_records(4096, Br)
            row = []
            for var in (0, 1):
                rs.set_option(rs.OPT_BENCH_PREFETCH, var)
                ms, nbytes = rs.bench_backward(Br, lamb, 8, True)
                row.append("v%d %.3f ms %5.0f GB/s" % (var, np.median(ms), nbytes / np.median(ms) / 1e6))
            st = rs.download_counts(Br)
            print("%s N=%d %s B=%d lamb=%g: %s" % (cfg, N, dtype, Br, lamb, " | ".join(row)), flush=True)
            rs.close()
