"""B = 1 Mi on one GPU (8192 chunks in the verdict kernel's look-back, one CTA each): checks that every
instance exits and that a slice solved on its own returns the same bits.  Run it under `timeout`."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cilqr_b200 as cb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
pb = cb.synthetic_batch("C1", B, N=50)
with cb.BatchSolver(pb.templates, B, 50, pb.max_obs, "f64") as s:
    s.upload(pb)
    t0 = time.perf_counter(); s.solve_resident(B); dt = time.perf_counter() - t0
    out = s.download(B, want_gains=False)
    c = s.counters()
    lo = B // 2 + 12345
    sub = s.solve(pb.slice(lo, lo + 256), want_gains=False)
same = all(np.array_equal(getattr(out, f)[lo:lo + 256], getattr(sub, f)) for f in ("u", "x", "J", "iters", "status"))
print("B=%d: %.1f ms, %.2f M iter/s, rounds %d, exits %s (sum %d), slice identical: %s"
      % (B, dt * 1e3, out.iters.sum() / dt / 1e6, c["rounds"], c["exits"], sum(c["exits"].values()), same))
