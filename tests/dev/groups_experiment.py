"""How much would independent sub-batches solved concurrently buy?  G handles of B / G instances each, one Python thread per
handle (the C call releases the GIL), wall clock around all of them, against one handle of B.  Development experiment."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cilqr_b200 as cb

spec = sys.argv[1] if len(sys.argv) > 1 else "C1:4096:f64"
cfg, B, dt = spec.split(":"); B = int(B)
N = {"C2": 100, "C4": 200}.get(cfg, 50)
sp = cb.synth_spec(cfg)
for G in [int(g) for g in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4,8").split(",")]:
    per = B // G
    solvers = []
    for g in range(G):
        s = cb.BatchSolver(sp.templates, per, N, sp.max_obs, dt)
        s.generate(sp, per, first_id=g * per)
        solvers.append(s)
    best = None
    for rep in range(5):
        bar = threading.Barrier(G + 1)
        def work(s):
            bar.wait()
            s.solve_resident(per)
        th = [threading.Thread(target=work, args=(s,)) for s in solvers]
        for t in th: t.start()
        bar.wait()
        t0 = time.perf_counter()
        for t in th: t.join()
        dtm = time.perf_counter() - t0
        best = dtm if best is None else min(best, dtm)
    iters = sum(int(s.download_counts(per)["iters"]) for s in solvers)
    print("%s G=%d x %d: %.2f ms  %d iter_steps  %.2f M iter/s" % (spec, G, per, best * 1e3, iters, iters / best / 1e6), flush=True)
    for s in solvers: s.close()
