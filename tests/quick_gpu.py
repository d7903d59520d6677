"""Ad-hoc GPU timing used during development (not a test, not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cilqr_b200 as cb

def run(cfg, B, N, dtype):
    pb = cb.synthetic_batch(cfg, B, N=N)
    with cb.BatchSolver(pb.templates, pb.B, pb.N, pb.max_obs, dtype) as s:
        s.set_option(s.OPT_FUSED_BACKWARD, 0)  # the backward bench below runs on the records the solves leave
        s.upload(pb)
        for rep in range(3):
            t = time.time(); s.solve_resident(B); out = s.download(B); dt = time.time() - t
            c = s.counters()
            print("%s B=%d N=%d %s: %.1f ms  iters %d (mean %.1f) rounds %d launches %d exits %s -> %.2f M iter/s"
                  % (cfg, B, N, dtype, dt * 1e3, c["total_iters"], out.iters.mean(), c["rounds"], c["launches"], c["exits"], c["total_iters"] / dt / 1e6))
        for pf in (0, 1):
            s.set_option(s.OPT_BENCH_PREFETCH, pf)
            ms, nbytes = s.bench_backward(B, 0.0, 10, True)
            print("  backward B=%d prefetch=%d: median %.3f ms -> %.0f GB/s" % (B, pf, np.median(ms), nbytes / np.median(ms) / 1e6))

if __name__ == "__main__":
    run("C1", 4096, 50, "f64")
    run("C1", 4096, 50, "f32")
    run("C1", 65536, 50, "f64")
    run("C1", 65536, 50, "f32")
    run("C1", 262144, 50, "f64")
