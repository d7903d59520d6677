"""One C1 solve for ncu launch lists (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cilqr_b200 as cb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dtype = sys.argv[2] if len(sys.argv) > 2 else "f64"
pb = cb.synthetic_batch("C1", B, N=50)
with cb.BatchSolver(pb.templates, pb.B, pb.N, pb.max_obs, dtype) as s:
    s.upload(pb)
    s.solve_resident(B)
    print(s.counters())
