import sys, os, time, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cilqr_b200 as cb
from importlib import reload
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
pb = cb.synthetic_batch("C1", B, N=50)
for lib in sorted(glob.glob(os.path.join(os.path.dirname(cb.LIB_PATH), "libvariant_*.so"))):
    import cilqr_b200.binding as bd
    bd._lib = None; bd.LIB_PATH = lib
    with cb.BatchSolver(pb.templates, B, 50, pb.max_obs, "f64") as s:
        s.upload(pb)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); s.solve_resident(B); ts.append(time.perf_counter() - t0)
        out = s.download(B, want_gains=False)
    print("%-28s B=%d: %.1f ms  %.2f M iter/s" % (os.path.basename(lib), B, min(ts) * 1e3, out.iters.sum() / min(ts) / 1e6), flush=True)
