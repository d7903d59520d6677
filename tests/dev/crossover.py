import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cilqr_b200 as cb
for B in (8192, 12288, 16384, 24576, 32768):
    pb = cb.synthetic_batch("C1", B, N=50)
    with cb.BatchSolver(pb.templates, B, 50, pb.max_obs, "f64") as s:
        s.upload(pb)
        res = {}
        for name, thr in (("latency", 1 << 30), ("throughput", 0)):
            s.set_option(s.OPT_PREFETCH_BELOW, thr)
            ts = []
            for _ in range(3):
                t0 = time.perf_counter(); s.solve_resident(B); ts.append(time.perf_counter() - t0)
            res[name] = min(ts) * 1e3
        print("B=%6d  latency variants %.1f ms   throughput variants %.1f ms" % (B, res["latency"], res["throughput"]), flush=True)
