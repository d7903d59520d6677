"""Look-ahead rounds on / off: same bits, solve time.  `python tests/dev/lookahead_ab.py [C1:4096:f64 ...]`.  Development aid."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cilqr_b200 as cb

specs = sys.argv[1:] or ["C1:4096:f64", "C1:1024:f64", "C3:4096:f64", "C1:4096:f32", "C1:16384:f64", "C2:2048:f64"]
for spec in specs:
    cfg, B, dt = spec.split(":")
    B = int(B)
    N = {"C2": 100, "C4": 200}.get(cfg, 50)
    pb = cb.synthetic_batch(cfg, B, N=N)
    outs = {}
    with cb.BatchSolver(pb.templates, B, N, pb.max_obs, dt) as s:
        s.upload(pb)
        for la in (0, 1, 0, 1):
            s.set_option(s.OPT_LOOKAHEAD, 16384 if la else 0)
            ts = []
            for r in range(4):
                s.reset()
                t0 = time.perf_counter(); s.solve_resident(B); ts.append(time.perf_counter() - t0)
            out = s.download(B)
            c = s.counters()
            outs[la] = out
            print("%s lookahead=%d: %.2f ms (best of 4), %d iter_steps, %.2f M iter/s, rounds %d, trials %d, launches %d"
                  % (spec, la, min(ts) * 1e3, out.iters.sum(), out.iters.sum() / min(ts) / 1e6, c["rounds"], c["total_trials"], c["launches"]), flush=True)
    bad = [f for f in ("u", "x", "J", "K", "d", "iters", "status", "exit_reason", "step_cost")
           if not np.array_equal(getattr(outs[0], f), getattr(outs[1], f), equal_nan=True)]
    print("   same bits:", "YES" if not bad else "NO: %s" % bad, flush=True)
