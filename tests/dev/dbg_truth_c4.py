"""development: gpu/yardstick error ratio of a 1-iteration lockstep on C4 for the main library and libvariant_*.so"""
import sys, os, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, cilqr_b200 as cb
import cilqr_b200.binding as bd
from oracle import oracle_py as op
cfg, B, N = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else ("C4", 32, 200)
pb = cb.synthetic_batch(cfg, B, N=N)
for td in pb.templates: td.params = dict(td.params, max_iter=1)
truth = op.solve_batch(pb, "f80", trace_cap=1)
runs = [op.solve_batch(pb, "f64", trace_cap=1), op.solve_batch(pb, "f64pm", trace_cap=1)]
rng = np.random.default_rng(6); x0 = pb.x0.copy()
for _ in range(16):
    pb.x0 = x0 * (1 + 2.0**-52 * rng.uniform(-1, 1, x0.shape)); runs.append(op.solve_batch(pb, "f64", trace_cap=1))
pb.x0 = x0
main = bd.LIB_PATH
for lib in [main] + sorted(glob.glob(os.path.join(os.path.dirname(main), "libvariant_*.so"))):
    bd._lib = None; bd.LIB_PATH = lib
    with cb.BatchSolver(pb.templates, B, N, pb.max_obs, "f64") as s:
        s.enable_trace(1); out = s.solve(pb); gst, gal, gco = s.get_trace(B)
    rx, rJ = [], []
    for b in range(B):
        mine = (gst[b, 0], gal[b, 0])
        peers = [r for r in runs if (r.tr_status[b, 0], r.tr_alpha[b, 0]) == mine]
        if (truth.tr_status[b, 0], truth.tr_alpha[b, 0]) != mine or not peers: continue
        e = lambda a, t: np.abs(a - t).max() / max(np.abs(t).max(), 1.0)
        rx.append(e(out.x[b], truth.x[b]) / max(max(e(r.x[b], truth.x[b]) for r in peers), 1e-12))
        rJ.append(abs(out.J[b, 1] - truth.J[b, 1]) / max(max(abs(r.J[b, 1] - truth.J[b, 1]) for r in peers), 1e-12 * abs(truth.J[b, 1])))
    print(os.path.basename(lib), "x ratio max %.2f median %.2f | J ratio max %.2f median %.2f (n=%d)" % (max(rx), np.median(rx), max(rJ), np.median(rJ), len(rx)))
