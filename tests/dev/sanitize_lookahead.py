"""Small look-ahead solves for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): whole-solve look-ahead
rounds, the mid-solve switch, wide line searches (every trial a job) on a few instances."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cilqr_b200 as cb
for cfg, B, dt, bound, iters in (("C1", 96, "f64", 1, 12), ("C3", 70, "f64", 16384, 25), ("C2", 40, "f32", 16384, 10),
                                 ("C3", 150, "f64", 100, 40), ("C1", 33, "f32", 1, 100)):
    pb = cb.synthetic_batch(cfg, B, N={"C2": 100}.get(cfg, 50))
    for t in pb.templates:
        t.params = dict(t.params, max_iter=iters)
    with cb.BatchSolver(pb.templates, B, pb.N, pb.max_obs, dt) as s:
        s.set_option(s.OPT_LOOKAHEAD, bound)
        out = s.solve(pb)
        out2 = s.solve(pb)
        print(cfg, B, dt, "lookahead", bound, "iters", int(out.iters.sum()), "rounds", s.counters()["rounds"],
              "finite", bool(np.isfinite(out.x).all()), flush=True)
