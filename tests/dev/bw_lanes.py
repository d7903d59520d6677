"""Backward pass: one lane vs four lanes per trajectory — same bits (stage operator), whole-solve time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cilqr_b200 as cb
for spec in sys.argv[1:] or ["C1:4096:f64", "C1:1024:f64", "C3:4096:f64", "C1:4096:f32", "C2:2048:f64", "C1:512:f64"]:
    cfg, B, dt = spec.split(":"); B = int(B)
    N = {"C2": 100, "C4": 200}.get(cfg, 50)
    pb = cb.synthetic_batch(cfg, B, N=N)
    outs = {}
    with cb.BatchSolver(pb.templates, B, N, pb.max_obs, dt) as s:
        s.upload(pb)
        for staged in (1, 2, 1, 2):
            s.set_option(s.OPT_STAGED_BACKWARD, staged)
            ts = []
            for r in range(4):
                s.reset()
                t0 = time.perf_counter(); s.solve_resident(B); ts.append(time.perf_counter() - t0)
            outs[staged] = s.download(B)
            c = s.counters()
            print("%s staged=%d: %.2f ms best of 4, %.2f M iter/s, rounds %d" % (spec, staged, min(ts) * 1e3, outs[staged].iters.sum() / min(ts) / 1e6, c["rounds"]), flush=True)
    bad = [f for f in ("u", "x", "J", "K", "d", "iters", "status", "exit_reason", "step_cost")
           if not np.array_equal(getattr(outs[1], f), getattr(outs[2], f), equal_nan=True)]
    print("   same bits:", "YES" if not bad else "NO: %s" % bad, flush=True)
