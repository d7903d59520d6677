"""One (or a few) resident solves of a config, for profiler runs: `python tests/dev/one_solve.py C1:262144:f64 [reps]`.
Prints iterations, solve time, rounds.  Development aid, not a test."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cilqr_b200 as cb

cfg, B, dt = sys.argv[1].split(":")
B = int(B)
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
opts = sys.argv[3] if len(sys.argv) > 3 else ""
N = {"C2": 100, "C4": 200}.get(cfg, 50)
pb = cb.synthetic_batch(cfg, B, N=N)
with cb.BatchSolver(pb.templates, B, N, pb.max_obs, dt) as s:
    for kv in filter(None, opts.split(",")):
        s.set_option(int(kv.split("=")[0]), int(kv.split("=")[1]))
    s.upload(pb)
    for r in range(reps):
        t0 = time.perf_counter(); s.solve_resident(B); t = time.perf_counter() - t0
        out = s.download(B, want_gains=False)
        c = s.counters()
        print("%s B=%d N=%d %s rep %d: %.2f ms, %d iter_steps, %.2f M iter/s, rounds %d, trials %d, launches %d, exits %s"
              % (cfg, B, N, dt, r, t * 1e3, out.iters.sum(), out.iters.sum() / t / 1e6, c["rounds"], c["total_trials"],
                 c["launches"], c["exits"]), flush=True)
