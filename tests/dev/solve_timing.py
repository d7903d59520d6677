"""Solve time and per-stage device time of every library variant (libcilqr_b200.so plus any
libvariant_*.so next to it) on a list of configs: `python tests/dev/solve_timing.py C1:4096:f64 C1:65536:f64`.
Also checks that all variants return the same bits as the first one (development A/B, not a test)."""
import sys, os, time, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cilqr_b200 as cb
import cilqr_b200.binding as bd

# --sets "6=1;6=0": option sets (id=value,...) to run the main library with, one line each
sets = [""]
args = sys.argv[1:]
if "--sets" in args:
    i = args.index("--sets")
    sets = args[i + 1].split(";")
    del args[i:i + 2]
cfgs = [a.split(":") for a in args] or [["C1", "4096", "f64"]]
main_lib = bd.LIB_PATH
libs = [(l, "") for l in sorted(glob.glob(os.path.join(os.path.dirname(main_lib), "libvariant_*.so")))] + [(main_lib, o) for o in sets]
for cfg, B, dt in cfgs:
    B = int(B)
    N = 100 if cfg == "C2" else 50
    pb = cb.synthetic_batch(cfg, B, N=N)
    first = None
    for lib, opts in libs:
        bd._lib = None
        bd.LIB_PATH = lib
        with cb.BatchSolver(pb.templates, B, N, pb.max_obs, dt) as s:
            for kv in filter(None, opts.split(",")):
                s.set_option(int(kv.split("=")[0]), int(kv.split("=")[1]))
            s.upload(pb)
            ts = []
            for _ in range(5):
                t0 = time.perf_counter(); s.solve_resident(B); ts.append(time.perf_counter() - t0)
            out = s.download(B, want_gains=False)
            c = s.counters()
            s.set_option(s.OPT_PROFILE_STAGES, 1)
            s.solve_resident(B)
            st = s.stage_times()
            per = " ".join("%s %.1f" % (k, 1e3 * m / max(1, n)) for k, (m, n) in st.items())
        if first is None:
            first = out
        same = all(np.array_equal(np.asarray(getattr(first, f)), np.asarray(getattr(out, f)), equal_nan=True)
                   for f in ("u", "x", "iters", "status"))
        print("%-20s %-8s %s B=%d N=%d %s: solve %.2f ms (median %.2f) %.2f M iter/s rounds %d  us/launch: %s  same_bits=%s"
              % (os.path.basename(lib), opts, cfg, B, N, dt, min(ts) * 1e3, np.median(ts) * 1e3,
                 out.iters.sum() / min(ts) / 1e6, c["rounds"], per, same), flush=True)
