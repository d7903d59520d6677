"""A/B of CILQR_OPT_PIPELINE (rollout + waypoint match as one two-stage kernel) on the benchmark batch:
solve time, per-stage device time, and that the two settings return identical bits."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cilqr_b200 as cb

for cfg, B, dt in (("C1", 4096, "f64"), ("C1", 1024, "f64"), ("C3", 4096, "f64"), ("C1", 4096, "f32"), ("C1", 12288, "f64")):
    pb = cb.synthetic_batch(cfg, B, N=50)
    with cb.BatchSolver(pb.templates, B, 50, pb.max_obs, dt) as s:
        s.upload(pb)
        outs = {}
        for pipe in (0, 16, 8, 1):
            s.set_option(s.OPT_PIPELINE, pipe)
            ts = []
            for _ in range(4):
                t0 = time.perf_counter(); s.solve_resident(B); ts.append(time.perf_counter() - t0)
            outs[pipe] = s.download(B)
            s.set_option(s.OPT_PROFILE_STAGES, 1)
            s.solve_resident(B)
            st = s.stage_times()
            s.set_option(s.OPT_PROFILE_STAGES, 0)
            per = " ".join("%s %.1f" % (k, 1e3 * m / max(1, n)) for k, (m, n) in st.items())
            print("%s B=%d %s pipeline=%d  solve %.2f ms  us/launch: %s" % (cfg, B, dt, pipe, min(ts) * 1e3, per), flush=True)
        same = all(all(np.array_equal(np.asarray(getattr(outs[0], f)), np.asarray(getattr(outs[q], f)), equal_nan=True)
                   for f in ("u", "x", "iters", "status")) for q in (16, 8, 1))
        print("   identical results:", same, flush=True)
