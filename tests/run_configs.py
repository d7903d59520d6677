"""Runs the larger BASELINE.json configs on one GPU and prints one result line each (development /
reporting aid; the contract benchmark is bench.py).  Usage: python tests/run_configs.py C2 [B] [dtype]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cilqr_b200 as cb


def run(cfg, B, dtype, N=None, reps=2):
    t0 = time.perf_counter()
    pb = cb.synthetic_batch(cfg, B, N=N)
    t_gen = time.perf_counter() - t0
    N = pb.N
    with cb.BatchSolver(pb.templates, B, N, pb.max_obs, dtype) as s:
        s.set_option(s.OPT_FUSED_BACKWARD, 0)  # K5 below runs on the records the solve leaves: whole records wanted
        t0 = time.perf_counter()
        s.upload(pb)
        t_up = time.perf_counter() - t0
        best = None
        for _ in range(reps):
            t0 = time.perf_counter()
            s.solve_resident(B)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        out = s.download(B, want_gains=False)
        c = s.counters()
        ms, nbytes = s.bench_backward(B, 0.0, 8, True)   # K5 on the records of the final iterates
    sz = 8 if dtype == "f64" else 4
    line = {
        "config": cfg, "B": B, "N": N, "dtype": dtype, "n_obs_max": int(pb.max_obs),
        "solve_ms": round(best * 1e3, 2), "iter_steps": int(out.iters.sum()),
        "iterations_per_s": round(out.iters.sum() / best), "solves_per_s": round(B / best),
        "mean_iters": round(float(out.iters.mean()), 2), "rounds": c["rounds"], "trials": c["total_trials"],
        "exits": c["exits"], "converged_frac": round(float((out.exit_reason == 1).mean()), 4),
        "k5_ms": round(float(np.median(ms)), 4),
        "k5_GBps_compact": round(nbytes / np.median(ms) / 1e6), "k5_bytes_per_traj": (38 * N + 18) * sz,
        "gen_s": round(t_gen, 1), "upload_s": round(t_up, 2),
        "J_final_median": float(np.median(out.J[:, 1])),
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    cfg = sys.argv[1]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else {"C1": 4096, "C2": 262144, "C3": 131072, "C4": 65536}[cfg]
    dtype = sys.argv[3] if len(sys.argv) > 3 else ("f32" if cfg == "C4" else "f64")
    run(cfg, B, dtype)
