"""Repeat look-ahead solves and compare every one with the sequential result (races between the two streams would show
up as a mismatch in some repetition).  Development aid."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cilqr_b200 as cb
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 25
for cfg, B, dt, bounds in (("C3", 700, "f64", (16384, 1, 100)), ("C1", 2500, "f64", (16384, 512)), ("C2", 300, "f32", (16384, 1)),
                           ("C1", 9000, "f64", (16384, 512)), ("C3", 64, "f64", (1,))):
    pb = cb.synthetic_batch(cfg, B, N={"C2": 100}.get(cfg, 50))
    bad = 0
    with cb.BatchSolver(pb.templates, B, pb.N, pb.max_obs, dt) as s:
        s.set_option(s.OPT_LOOKAHEAD, 0)
        ref = s.solve(pb)
        for r in range(reps):
            s.set_option(s.OPT_LOOKAHEAD, bounds[r % len(bounds)])
            s.reset()
            out = s.solve(pb)
            ok = all(np.array_equal(getattr(ref, f), getattr(out, f), equal_nan=True) for f in ("u", "x", "J", "K", "d", "iters", "status", "exit_reason", "step_cost"))
            bad += not ok
    print("%s x %d %s: %d repetitions, %d mismatches" % (cfg, B, dt, reps, bad), flush=True)
