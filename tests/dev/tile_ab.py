"""development: tile kernel on/off on a list of configs: time, rounds, same bits"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, cilqr_b200 as cb
for arg in sys.argv[1:] or ["C1:4096:f64"]:
    cfg, B, dt = arg.split(":"); B = int(B)
    spec = cb.synth_spec(cfg); N = spec.N
    res = {}
    with cb.BatchSolver(spec.templates, B, N, spec.max_obs, dt) as s:
        s.enable_trace(100)
        for tiles in (0, 1):
            s.set_option(s.OPT_TILE_KERNEL, tiles)
            s.generate(spec, B)
            ts = []
            for _ in range(4):
                t0 = time.perf_counter(); s.solve_resident(B); ts.append(time.perf_counter() - t0)
            out = s.download(B); c = s.counters(); tr = s.get_trace(B)
            res[tiles] = (out, tr)
            print("%s B=%d %s tiles=%d: %.2f ms (median %.2f), %.2f M iter/s, rounds %d trials %d launches %d" % (
                cfg, B, dt, tiles, min(ts) * 1e3, np.median(ts) * 1e3, out.iters.sum() / min(ts) / 1e6, c["rounds"], c["total_trials"], c["launches"]), flush=True)
    same = all(np.array_equal(getattr(res[0][0], f), getattr(res[1][0], f), equal_nan=True) for f in ("u", "x", "J", "K", "d", "iters", "status", "exit_reason", "step_cost"))
    same_tr = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(res[0][1], res[1][1]))
    print("   same bits:", same, " same traces:", same_tr, flush=True)
