import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, cilqr_b200 as cb
from oracle import oracle_py as op
from helpers import oracle_stage, perturbed_trajectories
cfg, B, N, dtype = "C2", 48, 100, "f32"
pb = cb.synthetic_batch(cfg, B, N=N)
u, x = perturbed_trajectories(pb, seed=9)
lx, lu = np.zeros((B, N + 1, 4)), np.zeros((B, N, 2))
lxx, luu = np.zeros((B, N + 1, 4, 4)), np.zeros((B, N, 2, 2))
A, Bm = np.zeros((B, N, 4, 4)), np.zeros((B, N, 4, 2))
for b in range(B):
    _, _, dv, A[b], Bm[b], _ = oracle_stage(pb, b, u[b], x[b], "f64")
    lx[b], lu[b], lxx[b], luu[b] = dv["lx"], dv["lu"], dv["lxx"], dv["luu"]
lx, lu, lxx, luu, A, Bm = [v.astype(np.float32).astype(np.float64) for v in (lx, lu, lxx, luu, A, Bm)]
lamb = np.where(np.arange(B) % 3 == 0, 0.0, 2.0 ** (np.arange(B) % 5))
res = {}
for fl in ("fast", "parity"):
    with cb.BatchSolver(pb.templates, B, N, pb.max_obs, dtype, flavour=fl) as s:
        for var in (0, 1, 2):
            s.set_option(s.OPT_BENCH_PREFETCH, var)
            res[fl, var] = s.stage_backward(lx, lu, lxx, luu, A, Bm, lamb)
def stop(K):
    nz = np.nonzero(np.abs(K).sum(axis=(1, 2)) == 0)[0]
    return int(nz.max()) if len(nz) else -1
for b in range(8):
    o = op.riccati(N, lx[b], lu[b], lxx[b], luu[b], A[b], Bm[b], lamb[b], "f32")
    o64 = op.riccati(N, lx[b], lu[b], lxx[b], luu[b], A[b], Bm[b], lamb[b], "f64")
    print(b, "lamb", lamb[b], "oracle32 st", o[3], "stop", stop(o[1]), "| oracle64 st", o64[3], "stop", stop(o64[1]),
          "| fast", [(int(res["fast", v][3][b]), stop(res["fast", v][1][b])) for v in (0, 1, 2)],
          "| parity", [(int(res["parity", v][3][b]), stop(res["parity", v][1][b])) for v in (0, 1, 2)],
          "Kmax fast", np.abs(res["fast", 1][1][b]).max(), "nan?", np.isnan(res["fast", 1][1][b]).any(), np.isnan(o[1]).any())
