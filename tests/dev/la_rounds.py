import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cilqr_b200 as cb
spec = sys.argv[1] if len(sys.argv) > 1 else "C1:1024:f64"
la = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg, B, dt = spec.split(":"); B = int(B)
N = {"C2": 100, "C4": 200}.get(cfg, 50)
pb = cb.synthetic_batch(cfg, B, N=N)
path = "/tmp/prof_dump.txt"
os.environ["CILQR_PROFILE_DUMP"] = path
with cb.BatchSolver(pb.templates, B, N, pb.max_obs, dt) as s:
    s.upload(pb)
    s.set_option(s.OPT_LOOKAHEAD, la)
    if len(sys.argv) > 3: s.set_option(s.OPT_STAGED_BACKWARD, int(sys.argv[3]))
    s.solve_resident(B)
    s.set_option(s.OPT_PROFILE_STAGES, 1)
    s.solve_resident(B)
rows = np.loadtxt(path)
names = {0: "derivs", 1: "backward", 2: "forward", 3: "ref_match", 4: "cost", 5: "decide", -1: "gap"}
S = rows[rows[:, 0] == 0]; Bm = rows[rows[:, 0] == 1]
ends = np.where(S[:, 1] == -1)[0]
print("%s la=%d rounds=%d total %.2f ms" % (spec, la, len(ends), (S[-1, 2] + S[-1, 3]) / 1e3))
i0 = 0
bj = 0
for r, e in enumerate(ends):
    seg = S[i0:e]
    d = {}
    for x in seg: d[names[int(x[1])]] = d.get(names[int(x[1])], 0) + x[3]
    t0 = seg[0, 2] if len(seg) else 0
    t1 = S[e, 2] + S[e, 3]
    is_la = len(seg) and int(seg[0, 1]) == 2
    if is_la:
        bb = Bm[(Bm[:, 2] >= t0) & (Bm[:, 2] < t1) & (Bm[:, 1] >= 0)]
        d.update({"B:" + names[int(x[1])]: x[3] for x in bb})
        if len(bb): d["B:start"] = bb[0, 2] - t0
    if r < 4 or r % 10 == 0 or r > len(ends) - 4:
        print("  round %3d %s: %.1f us  " % (r, "LA" if is_la else "  ", t1 - t0), {k: round(float(v), 1) for k, v in d.items()})
    i0 = e + 1
