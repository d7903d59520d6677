"""Forward-stage time of the first round (all B trials at alpha = 1) for every libvariant_*.so."""
import sys, os, glob, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cilqr_b200 as cb
import cilqr_b200.binding as bd
for B in (1024, 4096):
    pb = cb.synthetic_batch("C1", B, N=50)
    for t in pb.templates:
        t.params["max_iter"] = 1
    for lib in sorted(glob.glob(os.path.join(os.path.dirname(cb.LIB_PATH), "libvariant_*.so"))):
        bd._lib = None; bd.LIB_PATH = lib
        for dt in ("f64", "f32"):
            with cb.BatchSolver(pb.templates, B, 50, pb.max_obs, dt) as s:
                s.upload(pb)
                s.set_option(s.OPT_PROFILE_STAGES, 1)
                best = 1e9
                for _ in range(5):
                    s.solve_resident(B)
                    st = s.stage_times()
                    best = min(best, st["forward"][0] * 1e3 + st["ref_match"][0] * 1e3)
                print("%-26s B=%d %s forward+match first round: %.1f us" % (os.path.basename(lib), B, dt, best), flush=True)
