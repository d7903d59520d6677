"""Solve time against the regime threshold (CILQR_OPT_PREFETCH_BELOW, applied per round to the number of
instances still running): python tests/dev/crossover.py [dtype]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cilqr_b200 as cb
dtype = sys.argv[1] if len(sys.argv) > 1 else "f64"
for B in (8192, 16384, 32768, 65536, 262144):
    pb = cb.synthetic_batch("C1", B, N=50)
    with cb.BatchSolver(pb.templates, B, 50, pb.max_obs, dtype) as s:
        s.upload(pb)
        line = []
        for thr in (0, 1024, 4096, 8192, 16384, 32768, 65536, 1 << 30):
            if thr != (1 << 30) and thr > 4 * B:
                continue
            s.set_option(s.OPT_PREFETCH_BELOW, thr)
            ts = []
            for _ in range(3):
                t0 = time.perf_counter(); s.solve_resident(B); ts.append(time.perf_counter() - t0)
            line.append("%s: %.1f" % ("inf" if thr == (1 << 30) else thr, min(ts) * 1e3))
        print("%s B=%6d  ms by threshold  %s" % (dtype, B, "  ".join(line)), flush=True)
