"""N > 1 host logic on CPU: two gloo ranks own disjoint instance-id ranges, solve them independently
(with the CPU oracle — there is no GPU here), and the concatenation equals the single-process result
bit for bit; timings reduce with MAX, work counters with SUM, exactly as bench.py does under NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cilqr_b200 as cb
from oracle import oracle_py as op

PER_RANK, WORLD = 6, 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        lo, hi = cb.shard.weak_range(PER_RANK, rank)
        assert (lo, hi) == cb.shard.shard_range(PER_RANK * WORLD, rank, WORLD)
        pb = cb.synthetic_batch("C3", hi - lo, N=20, first_id=lo)
        for td in pb.templates:
            td.params = dict(td.params, max_iter=3)
        r = op.solve_batch(pb, "f64", nthreads=1)
        t_local = 1.0 + rank  # stand-in for the CUDA-event time of this rank
        (t_max,), (iters,) = cb.shard.reduce_max_sum(dist, torch.device("cpu"), [t_local], [int(r.iters.sum())])
        x_all = cb.shard.gather_rows(dist, torch.device("cpu"), r.x, PER_RANK * WORLD)
        it_all = cb.shard.gather_rows(dist, torch.device("cpu"), r.iters.astype(np.int64), PER_RANK * WORLD)
        if rank == 0:
            q.put((t_max, iters, x_all, it_all))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_shards_equal_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    t_max, iters, x_all, it_all = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = cb.synthetic_batch("C3", PER_RANK * WORLD, N=20)
    for td in full.templates:
        td.params = dict(td.params, max_iter=3)
    ref = op.solve_batch(full, "f64", nthreads=1)
    assert t_max == 2.0                      # max over ranks
    assert iters == int(ref.iters.sum())     # sum over ranks
    assert np.array_equal(it_all, ref.iters)
    assert np.array_equal(x_all, ref.x)      # no cross-instance arithmetic: shards are bit-identical


def test_shard_ranges_cover_without_overlap():
    for total, world in ((1048576, 8), (10, 3), (7, 8)):
        edges = [cb.shard.shard_range(total, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
