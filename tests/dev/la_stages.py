import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cilqr_b200 as cb
spec = sys.argv[1] if len(sys.argv) > 1 else "C1:4096:f64"
cfg, B, dt = spec.split(":"); B = int(B)
N = {"C2": 100, "C4": 200}.get(cfg, 50)
pb = cb.synthetic_batch(cfg, B, N=N)
with cb.BatchSolver(pb.templates, B, N, pb.max_obs, dt) as s:
    s.upload(pb)
    for la in (0, 1):
        s.set_option(s.OPT_LOOKAHEAD, la)
        s.set_option(s.OPT_PROFILE_STAGES, 0)
        ts = []
        for r in range(5):
            t0 = time.perf_counter(); s.solve_resident(B); ts.append(time.perf_counter() - t0)
        c = s.counters()
        s.set_option(s.OPT_PROFILE_STAGES, 1)
        s.solve_resident(B)
        st = s.stage_times()
        print("%s lookahead=%d: %.2f ms best of 5, rounds %d; per-launch us: %s" % (
            spec, la, min(ts) * 1e3, c["rounds"], {k: round(v[0] / max(v[1], 1) * 1e3, 1) for k, v in st.items()}), flush=True)
