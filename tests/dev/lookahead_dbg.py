"""Look-ahead vs classic from identical starting state (reset before every solve): first differing instance + traces."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cilqr_b200 as cb

specs = sys.argv[1:] or ["C3:4096:f64", "C1:4096:f32", "C2:2048:f64", "C1:16384:f64"]
for spec in specs:
    cfg, B, dt = spec.split(":")
    B = int(B)
    N = {"C2": 100, "C4": 200}.get(cfg, 50)
    pb = cb.synthetic_batch(cfg, B, N=N)
    with cb.BatchSolver(pb.templates, B, N, pb.max_obs, dt) as s:
        s.enable_trace(100)
        def run(la):
            s.reset()
            s.set_option(s.OPT_LOOKAHEAD, la)
            out = s.solve(pb)
            return out, s.get_trace(B), s.counters()
        ref, rtr, rc = run(0)
        ref2, rtr2, _ = run(0)
        print(spec, "classic repeatable:", np.array_equal(ref.iters, ref2.iters) and np.array_equal(ref.x, ref2.x, equal_nan=True), "rounds", rc["rounds"])
        for rep in range(6):
            out, tr, c = run(1)
            d = np.where(out.iters != ref.iters)[0]
            same = all(np.array_equal(getattr(out, f), getattr(ref, f), equal_nan=True) for f in ("u", "x", "J", "K", "d", "iters", "status", "exit_reason", "step_cost"))
            print("  la rep %d: rounds %d same=%s  n_diff_iters=%d" % (rep, c["rounds"], same, len(d)), d[:8])
            if len(d):
                b = d[0]
                n = max(ref.iters[b], out.iters[b])
                k = next((i for i in range(n) if rtr[0][b, i] != tr[0][b, i] or rtr[1][b, i] != tr[1][b, i] or rtr[2][b, i] != tr[2][b, i]), -1)
                print("    inst %d: iters %d vs %d; first differing iteration %d" % (b, ref.iters[b], out.iters[b], k))
                lo = max(0, k - 2)
                print("    classic st", rtr[0][b, lo:k + 3], "al", rtr[1][b, lo:k + 3], "cost", rtr[2][b, lo:k + 3])
                print("    lookahd st", tr[0][b, lo:k + 3], "al", tr[1][b, lo:k + 3], "cost", tr[2][b, lo:k + 3])
                ks = [next((i for i in range(max(ref.iters[q], out.iters[q])) if rtr[0][q, i] != tr[0][q, i] or rtr[1][q, i] != tr[1][q, i] or rtr[2][q, i] != tr[2][q, i]), -1) for q in d[:200]]
                print("    first differing iteration over the first 200 differing instances: hist", np.bincount(np.array(ks) + 1)[:20])
                print("    differing instance ids mod 32 hist", np.bincount(d % 32, minlength=32))
