"""The roofline leg alone (K5 on resident records) for `ncu --set full` captures (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cilqr_b200 as cb
Br = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
dtype = sys.argv[2] if len(sys.argv) > 2 else "f64"
seed = cb.synthetic_batch("C1", 4096, N=50)
rs = cb.BatchSolver(seed.templates, Br, 50, seed.max_obs, dtype)
u0, x0 = rs.stage_init(seed.x0, seed.tmpl)
rs.stage_derivs(seed, u0, x0)
rs.bench_tile_records(4096, Br)
ms, nbytes = rs.bench_backward(Br, 0.0, 6, True)
print("B=%d %s: %.3f ms/launch, %.0f GB/s algorithmic" % (Br, dtype, np.mean(ms[2:]), nbytes / np.mean(ms[2:]) / 1e6))
