#!/usr/bin/env python
"""Benchmark of the batched CILQR hot path (BASELINE.json metric: iLQR iterations/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One step = one batched solve (forward rollout + backward Riccati loop of every instance, to its
own exit) of the C1 workload on each GPU: batch = 4096 randomized scenario_two_straight instances,
N = 50, nx = 4, nu = 2, fp64 (BASELINE.json configs[1]).  Ranks solve disjoint instance-id ranges
(weak scaling, no collective on the solve path).

  value     iterations/s (one iteration = one iter_step of one instance, counted whether or not the
            step was accepted), problems resident in HBM, CUDA events around each solve, L2 flushed
            between steps (untimed), max over ranks.
  e2e       same metric through cilqr_b200_solve_batch with pinned HOST buffers: host->device copy
            of the problems and device->host copy of (u, x, J, status, iters, exit) inside the
            timed region.
  roofline  the backward-pass kernel (K5) alone on HBM-resident derivative records at batch
            262144 (larger than L2), CUDA events per launch inside the library, against the
            measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the CPU oracle port (oracle/, the only thing here that executes it) on all host
            threads, on a bounded sample of the same workload.  Rank 0, N = 1 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ilqr_iterations_per_sec"
UNIT = "iterations/s"
WORKLOAD = "C1: batch=4096 randomized scenario_two_straight instances per GPU, N=50, nx=4, nu=2"
ROOFLINE_BATCH = 262144
# dram__bytes_read.sum + dram__bytes_write.sum per launch of k_backward<T, true> (the variant the solver uses at this
# batch) at B = 262144, N = 50, from the committed `ncu --set full` captures of this round's kernel
# (profiles/r02_k5_ncu_summary.txt): fp64 2.9675 GB read + 1.0297 GB written = 0.994 x the algorithmic 4.0223 GB;
# fp32 1.4941 + 0.5156 GB = 0.999 x 2.0112 GB.  Only valid for that batch.
ROOFLINE_TRAFFIC_BYTES = {"f64": 3.9972e9, "f32": 2.0098e9}
ROOFLINE_TRAFFIC_SOURCE = ("constant from the committed ncu --set full capture of this kernel at this batch "
                           "(profiles/r02_k5_ncu_summary.txt), not measured in this run")
# the other BASELINE configs, one GPU's share each, device-generated; (config, instances, dtype)
EXTRA_CONFIGS = [("C1", 65536, "f64"), ("C1", 262144, "f64"), ("C2", 262144, "f64"), ("C3", 131072, "f64"),
                 ("C4", 131072, "f32")]
MULTI_GPU_C3_PER_RANK = 131072  # x 8 ranks = BASELINE config C3 (1 048 576 mixed instances)
MULTI_GPU_C4_PER_RANK = 524288  # x 8 ranks = BASELINE config C4 (4 194 304 instances, N = 200, 5 obstacles, fp32)
CHECK_SLICE = 4096


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(cb, seed_batch, budget_s=15.0):
    """Oracle port on all host threads over a bounded prefix of the workload."""
    from oracle import oracle_py as op
    cores = os.cpu_count() or 1
    probe = seed_batch.slice(0, min(256, seed_batch.B))
    t0 = time.perf_counter()
    r = op.solve_batch(probe, "f64", nthreads=cores, want_traj=False)
    dt = time.perf_counter() - t0
    rate = r.iters.sum() / dt
    per_inst = max(r.iters.mean(), 1.0)
    n = int(min(seed_batch.B, max(256, rate * budget_s / per_inst)))
    sample = seed_batch.slice(0, n)
    t0 = time.perf_counter()
    r = op.solve_batch(sample, "f64", nthreads=cores, want_traj=False)
    dt = time.perf_counter() - t0
    return {"value": float(r.iters.sum() / dt), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "first %d instances of the C1 batch, %d iter_steps, %.1f s, one solve per thread" %
                      (n, int(r.iters.sum()), dt)}


def reference_sources_rate(full, cores, per_thread=6):
    """The reference's own, unmodified sources (oracle/_ref: compiled in place against the stand-in for the
    Eigen API, oracle/shim) on a small sample, for the record: informational, not the baseline — the
    stand-in's dynamic matrices are far slower than real Eigen, so the port above is the stronger arm."""
    try:
        from oracle import ref_py, oracle_py as op
        if not ref_py.available():
            return None
        from concurrent.futures import ThreadPoolExecutor
        n = min(full.B, cores * per_thread)
        sample = full.slice(0, n)
        iters = int(op.solve_batch(sample, "f64", nthreads=cores, want_traj=False).iters.sum())  # same bits, same counts
        td = sample.templates[0]

        def work(ids):
            s = ref_py.RefSolver(td.params, sample.N)
            for b in ids:
                s.solve(td, sample.ref_velo[b], int(sample.n_obs[b]), sample.obs[b], sample.borders[b], sample.x0[b])
                s.close()
                s = ref_py.RefSolver(td.params, sample.N)  # a fresh solver per problem: first solve, no warm start
            s.close()
        chunks = [list(range(i, n, cores)) for i in range(cores)]
        t0 = time.perf_counter()
        with ThreadPoolExecutor(cores) as ex:
            list(ex.map(work, chunks))
        dt = time.perf_counter() - t0
        return {"value": iters / dt, "unit": UNIT, "cores": cores, "kind": "reference sources + Eigen stand-in",
                "sample": "first %d instances, %d iter_steps, %.1f s" % (n, iters, dt)}
    except Exception as e:  # informational only
        return {"unavailable": str(e)[:200]}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; Eigen is absent)
    on all host threads.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import cilqr_b200 as cb
    from oracle import oracle_py as op
    cores = os.cpu_count() or 1
    full = cb.synthetic_batch("C1", 4096, N=50)
    # bounded sample per step, sized from a probe so the whole run ends within a few minutes
    probe = full.slice(0, 128)
    t0 = time.perf_counter()
    r = op.solve_batch(probe, "f64", nthreads=cores, want_traj=False)
    rate = r.iters.sum() / (time.perf_counter() - t0)
    total_steps = args.steps + args.warmup
    n = int(min(4096, max(128, rate * (150.0 / total_steps) / max(r.iters.mean(), 1.0))))
    sample = full.slice(0, n)
    times, iters = [], 0
    for i in range(total_steps):
        t0 = time.perf_counter()
        r = op.solve_batch(sample, "f64", nthreads=cores, want_traj=False)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
            iters += int(r.iters.sum())
    T = sum(times)
    value = iters / T
    desc = "first %d instances of the C1 batch per step, one solve per thread" % n
    shim = reference_sources_rate(full, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * T / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": desc,
                   "note": "reference CPU path = Eigen-free restatement (oracle/), bit-identical to the "
                           "reference sources compiled against oracle/shim; the real Eigen build is not possible here"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "reference_sources": shim,
    }))


def run_config(cb, cfg, B, dtype, device, first_id=0, reps=2, k5=True):
    """One GPU's share of a BASELINE config, generated on the device: best-of-`reps` resident solve (wall clock
    around cilqr_b200_solve_resident, which synchronises) and — k5 — the backward-pass kernel alone (L2 flushed, CUDA
    events) on the derivative records of the FIRST iteration of that workload at lambda = 0, with the fraction of
    instances whose recursion runs all N steps (the reference's recursion stops at a non-PD Q_uu, cpp:415-420: an
    instance that stops early moves fewer bytes than the algorithmic count, so the GB/s of a config where many
    stop is a lower bound on what the kernel streams)."""
    spec = cb.synth_spec(cfg)
    N = spec.N
    line = {"config": cfg, "instances": B, "N": N, "dtype": dtype}
    with cb.BatchSolver(spec.templates, B, N, spec.max_obs, dtype, device=device) as s:
        if k5:
            for t, td in enumerate(spec.templates):
                s.set_template(t, dict(td.params, max_iter=1))
            s.generate(spec, B, first_id=first_id)
            s.set_option(s.OPT_FUSED_BACKWARD, 0)  # (whole records wanted: the fused rounds do not store the control half)
            s.solve_resident(B)  # one iter_step: leaves the records of the initial trajectories
            s.set_option(s.OPT_FUSED_BACKWARD, 1)
            ms, nbytes = s.bench_backward(B, 0.0, 8, True)
            done = s.download(B, want_gains=False).status
            for t, td in enumerate(spec.templates):
                s.set_template(t, td.params)
            line["k5_GBps"] = round(nbytes / float(np.median(ms)) / 1e6)
            line["k5_bytes_per_trajectory"] = (38 * N + 18) * (8 if dtype == "f64" else 4)
            line["k5_full_recursions"] = round(float((done == 0).mean()), 4)
        s.generate(spec, B, first_id=first_id)
        best = None
        for i in range(reps + 1):  # first solve = warm-up
            t0 = time.perf_counter()
            s.solve_resident(B)
            dt = time.perf_counter() - t0
            if i > 0:
                best = dt if best is None else min(best, dt)
        c = s.counters()
        head = s.download(min(B, CHECK_SLICE), want_gains=False)
        st = s.download_counts(B)
        line.update({"solve_ms": round(best * 1e3, 2), "iter_steps": int(st["iters"]),
                     "iterations_per_s": round(st["iters"] / best), "rounds": c["rounds"], "trials": c["total_trials"],
                     "launches": c["launches"], "exits": st["exits"],
                     "data": "generated on the device (cilqr_b200_synth_generate)"})
    return line, head


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="instances per GPU (default: C1)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the extra BASELINE configs / multi-GPU C3 leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import cilqr_b200 as cb

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(device=dev)
    B, N = args.batch, 50

    # this rank's slice of the global instance-id range (no cross-rank data on the solve path)
    lo, hi = cb.shard.weak_range(B, rank)
    pb = cb.synthetic_batch("C1", hi - lo, N=N, first_id=lo)
    solver = cb.BatchSolver(pb.templates, B, N, pb.max_obs, args.dtype, device=local)
    solver.set_stream(stream.cuda_stream)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)  # 256 MiB > L2

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- value: problems resident in HBM ------------------------------------------------------
    solver.upload(pb)
    sampler = ClockSampler(local)
    step_ms, iters_total, launches = [], 0, 0
    with torch.cuda.stream(stream):
        for i in range(args.warmup + args.steps):
            if i == args.warmup:
                barrier()
                if rank == 0:
                    sampler.start()
            flush.add_(1.0)  # untimed L2 flush
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            solver.solve_resident(B)
            e1.record(stream)
            e1.synchronize()
            if i >= args.warmup:
                step_ms.append(e0.elapsed_time(e1))
                out = solver.download(B, want_gains=False)
                iters_total += int(out.iters.sum())
                launches += solver.counters()["launches"]
        barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_local = sum(step_ms) / 1e3

    # one extra, untimed solve with a CUDA event in front of every stage launch: where the step's
    # device time goes, and the backward-pass kernel's live duration inside the step
    in_step = None
    if rank == 0:
        # (sequential rounds for this one: in look-ahead rounds the stages overlap on two streams and their
        # per-stream intervals include each other's waits)
        solver.set_option(solver.OPT_PROFILE_STAGES, 1)
        solver.set_option(solver.OPT_LOOKAHEAD, 0)
        with torch.cuda.stream(stream):
            flush.add_(1.0)
            solver.solve_resident(B)
        st = solver.stage_times()
        solver.set_option(solver.OPT_PROFILE_STAGES, 0)
        solver.set_option(solver.OPT_LOOKAHEAD, 1)
        total = sum(v[0] for v in st.values())
        bw_ms, bw_n = st["backward"]
        sz = 8 if args.dtype == "f64" else 4
        in_step = {
            "stage_ms": {k: round(v[0], 3) for k, v in st.items()}, "rounds": bw_n,
            "backward_share": bw_ms / total if total else None,
            "backward_us_per_launch": 1e3 * bw_ms / max(bw_n, 1),
            "backward_GBps_if_all_instances_ran": (38 * N + 18) * sz * B / (1e6 * bw_ms / max(bw_n, 1)) if bw_ms else None,
            "note": "B=4096 records (63 MB) are L2-resident and only the running instances take part in a "
                    "round: the step is latency-bound, the HBM roofline of this kernel is measured at B=262144; "
                    "stage times of a solve in sequential rounds (look-ahead rounds are used "
                    "only for batches of up to 512 instances by default)",
        }

    # ---- e2e: pinned host buffers through cilqr_b200_solve_batch ---------------------------------
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()
    keep = []
    hp = {}
    for f in ("x0", "ref_velo", "borders", "tmpl", "n_obs", "obs"):
        t, v = pinned(getattr(pb, f))
        keep.append(t)
        hp[f] = v
    pbp = cb.BatchProblem(pb.templates, N, hp["x0"], hp["ref_velo"], hp["borders"], hp["tmpl"], hp["n_obs"], hp["obs"])
    res = solver._alloc_out(B, want_gains=False)
    for f in ("u", "x", "J", "step_cost", "status", "iters", "exit_reason"):
        t, v = pinned(getattr(res, f))
        keep.append(t)
        setattr(res, f, v)
    res.step_cost = None
    h2d = sum(hp[f].nbytes for f in hp)
    d2h = sum(getattr(res, f).nbytes for f in ("u", "x", "J", "status", "iters", "exit_reason"))
    e2e_ms, e2e_iters = [], 0
    with torch.cuda.stream(stream):
        for i in range(args.warmup + args.steps):
            if i == args.warmup:
                barrier()
            flush.add_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            solver.solve(pbp, out=res, want_gains=False)
            e1.record(stream)
            e1.synchronize()
            if i >= args.warmup:
                e2e_ms.append(e0.elapsed_time(e1))
                e2e_iters += int(res.iters.sum())
        barrier()
    t_e2e_local = sum(e2e_ms) / 1e3

    # ---- reduce over ranks: max time, summed work --------------------------------------------------
    per_rank_us = cb.shard.gather_ints(dist, dev, [int(1e6 * t_local / max(args.steps, 1)), int(1e6 * t_e2e_local / max(args.steps, 1))])
    (t_max, t_e2e), (iters_total, e2e_iters, launches) = cb.shard.reduce_max_sum(
        dist, dev, [t_local, t_e2e_local], [iters_total, e2e_iters, launches])
    solver.close()

    # ---- roofline: K5 alone at a batch larger than L2 (rank 0) ---------------------------------------
    def roofline_leg(dtype):
        peak, peak_src = measured_peak()
        Br = ROOFLINE_BATCH
        rs = cb.BatchSolver(pb.templates, Br, N, pb.max_obs, dtype, device=local)
        seed = cb.synthetic_batch("C1", 4096, N=N)
        u0, x0 = rs.stage_init(seed.x0, seed.tmpl)
        rs.stage_derivs(seed, u0, x0)          # real l_*, A, B of iteration 0 (K3 + K4)
        rs.bench_tile_records(4096, Br)        # replicated on the device up to the roofline batch
        rs.bench_backward(Br, 0.0, 3, True)    # warm-up launches
        ms, nbytes = rs.bench_backward(Br, 0.0, 20, True)
        # the flavour the solver runs in bandwidth-bound rounds: control half of the records (l_u, l_uu, A, B) computed
        # in the kernel from (v, yaw, u) instead of read back — fewer bytes, more arithmetic, same K / d bits
        rs.set_option(rs.OPT_BENCH_PREFETCH, 3)
        rs.bench_backward(Br, 0.0, 3, True)
        ms_f, _ = rs.bench_backward(Br, 0.0, 20, True)
        sz = 8 if dtype == "f64" else 4
        bytes_f = float(28 * N + 18) * sz * Br
        rs.close()
        achieved = nbytes / (float(np.mean(ms)) * 1e-3) / 1e9
        fused = {"kernel": "k_backward<T, prefetch, fused control half> (what the solver launches in bandwidth-bound rounds; "
                           "replaces this kernel + the control half of the derivative stage)",
                 "ms_per_launch": float(np.mean(ms_f)), "bytes_per_launch": bytes_f,
                 "achieved": bytes_f / (float(np.mean(ms_f)) * 1e-3) / 1e9, "unit": "GB/s",
                 "layout": "14 record fields + v, yaw, u read, K + d written: (28*N+18)*sizeof(T) bytes per trajectory",
                 "note": "no longer bound by HBM: 8 more transcendentals per step in the recursion's thread"}
        return {"bound": "hbm", "kernel": "k_backward<T, prefetch> (backward_pass Riccati recursion, cpp:383-440)",
                "fused_flavour": fused,
                "dtype": dtype, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ROOFLINE_TRAFFIC_BYTES.get(dtype),
                "traffic_source": ROOFLINE_TRAFFIC_SOURCE if dtype in ROOFLINE_TRAFFIC_BYTES else None,
                "peak_source": peak_src,
                "bytes_per_launch": nbytes, "ms_per_launch": float(np.mean(ms)),
                "batch": Br, "layout": "compact record, (38*N+18)*sizeof(T) bytes per trajectory",
                "frac_of_nominal_8000": achieved / 8000.0}

    roofline = roofline_f32 = None
    if rank == 0 and not args.no_roofline:
        roofline = roofline_leg(args.dtype)
        if args.dtype == "f64" and world == 1:
            roofline_f32 = roofline_leg("f32")  # BASELINE's "roofline run" (C4) is fp32: the same kernel in that type

    # ---- the other BASELINE configs (rank 0, one GPU's share each) and, under N > 1, config C3 sharded over
    # the ranks with a cross-rank bit-equality check -------------------------------------------------------
    configs, multi = None, None
    if not args.no_configs:
        if rank == 0 and world == 1:
            configs = []
            for cfg, Bc, dt in EXTRA_CONFIGS:
                try:
                    configs.append(run_config(cb, cfg, Bc, dt, local)[0])
                except Exception as e:  # e.g. not enough free HBM on a shared box
                    configs.append({"config": cfg, "instances": Bc, "dtype": dt, "error": str(e)[:200]})
        if world > 1:
            per = MULTI_GPU_C3_PER_RANK
            lo, hi = cb.shard.weak_range(per, rank)
            barrier()
            t0 = time.perf_counter()
            line, head = run_config(cb, "C3", per, "f64", local, first_id=lo, reps=1, k5=False)
            barrier()
            mine = cb.shard.result_checksum(head)
            # the neighbour's first CHECK_SLICE instances solved here as a small batch of their own
            nlo = cb.shard.weak_range(per, (rank + 1) % world)[0]
            _, nhead = run_config(cb, "C3", CHECK_SLICE, "f64", local, first_id=nlo, reps=1, k5=False)
            theirs = cb.shard.result_checksum(nhead)
            rows = cb.shard.gather_ints(dist, dev, [mine, theirs, line["iter_steps"], int(line["solve_ms"] * 1000)])
            # config C4 at its full per-GPU share (the "roofline run" of BASELINE.json: memory capacity, the
            # > 2^19-instance paths, device-side generation of 524288 x 201 x 5 obstacle samples per rank)
            c4 = [0, 0, 0]
            try:
                lo4 = cb.shard.weak_range(MULTI_GPU_C4_PER_RANK, rank)[0]
                l4, _ = run_config(cb, "C4", MULTI_GPU_C4_PER_RANK, "f32", local, first_id=lo4, reps=1, k5=False)
                c4 = [1, l4["iter_steps"], int(l4["solve_ms"] * 1000)]
            except Exception as e:  # e.g. not enough free HBM on a shared box
                if rank == 0:
                    print("C4 leg skipped: %s" % str(e)[:200], file=sys.stderr)
            rows4 = cb.shard.gather_ints(dist, dev, c4)
            if rank == 0:
                ok = all(rows[(r + 1) % world][0] == rows[r][1] for r in range(world))
                t_solve = max(r[3] for r in rows) / 1e6
                total_iters = sum(r[2] for r in rows)
                multi = {"config": "C3", "instances": per * world, "instances_per_gpu": per, "N": 50, "dtype": "f64",
                         "solve_ms": round(t_solve * 1e3, 2), "iter_steps": total_iters,
                         "iterations_per_s": round(total_iters / t_solve),
                         "slice_check": {"ok": bool(ok), "slice": CHECK_SLICE,
                                         "what": "rank r re-solves the first %d instances of rank r+1's range as a "
                                                 "batch of their own; 64-bit checksums of (x, u, J, iters) bits "
                                                 "all-gathered and compared" % CHECK_SLICE},
                         "data": "generated on the device, ids [rank*%d, (rank+1)*%d)" % (per, per)}
                if all(r[0] for r in rows4):
                    t4 = max(r[2] for r in rows4) / 1e6
                    multi["C4"] = {"instances": MULTI_GPU_C4_PER_RANK * world, "instances_per_gpu": MULTI_GPU_C4_PER_RANK,
                                   "N": 200, "dtype": "f32", "solve_ms": round(t4 * 1e3, 2),
                                   "iter_steps": sum(r[1] for r in rows4),
                                   "iterations_per_s": round(sum(r[1] for r in rows4) / t4)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(cb, cb.synthetic_batch("C1", 4096, N=N))

    if rank == 0:
        value = iters_total / t_max
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_max / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "N": N, "max_iter": 100,
                       "iterations_per_step": iters_total // max(args.steps, 1),
                       "l2": "flushed between steps (256 MiB write, untimed)",
                       "parallelism": "independent instance-id ranges per GPU, no collective on the solve path"},
            "clocks": clocks,
            "e2e": {"value": e2e_iters / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d) * world,
                    "d2h_bytes_per_step": int(d2h) * world, "ms_per_step": 1e3 * t_e2e / max(args.steps, 1)},
            "gpu_launches": launches,
            # (value uses the max over ranks.  On the 8-GPU box of this round seven ranks run 15.2-15.7 ms per step and
            # one — always the same GPU, with or without the ranks pinned to cores of their own — 16.3)
            "per_rank_ms_per_step": [round(r[0] / 1e3, 3) for r in per_rank_us],
            "roofline": roofline,
            "roofline_f32": roofline_f32,
            "in_step": in_step,
            "cpu_baseline": cpu,
            "configs": configs,
            "multi_gpu": multi,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
