"""Small solve for compute-sanitizer (memcheck / synccheck): every regime's kernels on a few instances."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cilqr_b200 as cb
for cfg, B, dt, pipe, thr, repack in (("C1", 96, "f64", 16, 12288, 1), ("C3", 40, "f64", 8, 12288, 1), ("C2", 33, "f32", 1, 12288, 1),
                                     ("C1", 64, "f64", 0, 12288, 1), ("C1", 64, "f64", 1, 0, 1), ("C3", 150, "f64", 1, 12288, 8),
                                     ("C1", 150, "f64", 1, 0, 8)):
    pb = cb.synthetic_batch(cfg, B, N=50)
    for t in pb.templates:
        t.params = dict(t.params, max_iter=6 if repack == 1 else 40)
    with cb.BatchSolver(pb.templates, B, 50, pb.max_obs, dt) as s:
        s.set_option(s.OPT_PIPELINE, pipe)
        s.set_option(s.OPT_PREFETCH_BELOW, thr)
        s.set_option(s.OPT_REPACK, repack)  # 8: repack forced down to 8 slots
        out = s.solve(pb)
        print(cfg, B, dt, "pipeline", pipe, "threshold", thr, "iters", int(out.iters.sum()), "finite", bool(np.isfinite(out.x).all()), flush=True)
