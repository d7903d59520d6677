"""Look-ahead rounds off / default (hybrid: below 512 running instances) / everywhere: same bits, solve time.  Development aid."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cilqr_b200 as cb

specs = sys.argv[1:] or ["C1:4096:f64", "C1:1024:f64", "C3:4096:f64", "C1:4096:f32", "C1:16384:f64", "C2:2048:f64"]
modes = [int(m) for m in os.environ.get("LA_MODES", "0,1,16384").split(",")]
for spec in specs:
    cfg, B, dt = spec.split(":")
    B = int(B)
    N = {"C2": 100, "C4": 200}.get(cfg, 50)
    pb = cb.synthetic_batch(cfg, B, N=N)
    outs = {}
    with cb.BatchSolver(pb.templates, B, N, pb.max_obs, dt) as s:
        s.upload(pb)
        for la in modes + modes:
            s.set_option(s.OPT_LOOKAHEAD, la)
            ts = []
            for r in range(4):
                s.reset()
                t0 = time.perf_counter(); s.solve_resident(B); ts.append(time.perf_counter() - t0)
            out = s.download(B)
            c = s.counters()
            outs[la] = out
            print("%s lookahead=%d: %.2f ms (best of 4), %d iter_steps, %.2f M iter/s, rounds %d, trials %d, launches %d"
                  % (spec, la, min(ts) * 1e3, out.iters.sum(), out.iters.sum() / min(ts) / 1e6, c["rounds"], c["total_trials"], c["launches"]), flush=True)
    for la in modes[1:]:
        bad = [f for f in ("u", "x", "J", "K", "d", "iters", "status", "exit_reason", "step_cost")
               if not np.array_equal(getattr(outs[modes[0]], f), getattr(outs[la], f), equal_nan=True)]
        print("   lookahead=%d same bits as %d:" % (la, modes[0]), "YES" if not bad else "NO: %s" % bad, flush=True)
