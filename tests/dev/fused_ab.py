"""Bandwidth-bound rounds with the fused backward pass on / off: same bits, solve time.  Development aid."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cilqr_b200 as cb
for spec in sys.argv[1:] or ["C1:262144:f64", "C2:131072:f64", "C3:131072:f64", "C1:65536:f64", "C4:131072:f32", "C1:65536:f32"]:
    cfg, B, dt = spec.split(":"); B = int(B)
    sp = cb.synth_spec(cfg)
    outs = {}
    with cb.BatchSolver(sp.templates, B, sp.N, sp.max_obs, dt) as s:
        s.generate(sp, B)
        for fused in (0, 1, 0, 1):
            s.set_option(s.OPT_FUSED_BACKWARD, fused)
            ts = []
            for r in range(3):
                s.reset()
                t0 = time.perf_counter(); s.solve_resident(B); ts.append(time.perf_counter() - t0)
            outs[fused] = s.download(min(B, 32768))
            st = s.download_counts(B)
            print("%s fused=%d: %.2f ms (best of 3), %d iter_steps, %.2f M iter/s, rounds %d" % (
                spec, fused, min(ts) * 1e3, st["iters"], st["iters"] / min(ts) / 1e6, s.counters()["rounds"]), flush=True)
    bad = [f for f in ("u", "x", "J", "K", "d", "iters", "status", "exit_reason", "step_cost")
           if not np.array_equal(getattr(outs[0], f), getattr(outs[1], f), equal_nan=True)]
    print("   same bits (first 32768 instances):", "YES" if not bad else "NO: %s" % bad, flush=True)
