import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cilqr_b200 as cb
spec = sys.argv[1] if len(sys.argv) > 1 else "C1:16384:f64"
cfg, B, dt = spec.split(":"); B = int(B)
N = {"C2": 100, "C4": 200}.get(cfg, 50)
for max_iter in (1, 2):
  for run_ahead in (3, 0):
    pb = cb.synthetic_batch(cfg, B, N=N)
    for td in pb.templates:
        td.params = dict(td.params, max_iter=max_iter)
    with cb.BatchSolver(pb.templates, B, N, pb.max_obs, dt) as s:
        s.set_option(s.OPT_RUN_AHEAD, run_ahead)
        def run(la):
            s.reset(); s.set_option(s.OPT_LOOKAHEAD, la); return s.solve(pb)
        ref = run(0)
        for rep in range(8):
            out = run(1)
            bad = {f: int((~np.isclose(getattr(out, f), getattr(ref, f), rtol=0, atol=0, equal_nan=True)).reshape(B, -1).any(axis=1).sum())
                   for f in ("u", "x", "J", "K", "d", "iters", "status", "step_cost")}
            first = {f: int(np.where((~np.isclose(getattr(out, f), getattr(ref, f), rtol=0, atol=0, equal_nan=True)).reshape(B, -1).any(axis=1))[0][:1].sum()) for f in bad if bad[f]}
            print("max_iter %d run_ahead %d rep %d: differing instances per field %s first %s" % (max_iter, run_ahead, rep, bad, first), flush=True)
