"""Host-side sharding of a batch of independent planning problems over ranks.

Instances never exchange data, so the only multi-process logic is: which instance ids a rank owns,
and how per-rank timings / counters are combined (time = max over ranks, work = sum over ranks).
Works with any initialised torch.distributed backend (NCCL on GPUs, gloo in the CPU tests).
"""


def shard_range(total, rank, world):
    """Contiguous slice [lo, hi) of the instance-id range for `rank` (GPU g <- [g*B/G, (g+1)*B/G))."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    lo = (total * rank) // world
    hi = (total * (rank + 1)) // world
    return lo, hi


def weak_range(per_rank, rank):
    """Weak scaling: every rank owns `per_rank` consecutive instance ids."""
    return rank * per_rank, (rank + 1) * per_rank


def reduce_max_sum(dist, device, maxima, sums):
    """all_reduce: element-wise MAX of `maxima` (floats), SUM of `sums` (ints).  dist=None: identity."""
    if dist is None:
        return list(maxima), [int(v) for v in sums]
    import torch
    t = torch.tensor(list(maxima), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    s = torch.tensor([int(v) for v in sums], dtype=torch.int64, device=device)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return t.tolist(), [int(v) for v in s.tolist()]


def gather_rows(dist, device, local, total_rows):
    """all_gather of equally sized per-rank row blocks (numpy in, numpy out) — used by tests to check
    that the concatenated shards equal the single-process result."""
    import numpy as np
    import torch
    if dist is None:
        return np.asarray(local)
    t = torch.from_numpy(np.ascontiguousarray(local)).to(device)
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t)
    return torch.cat(parts, 0).cpu().numpy()[:total_rows]


def result_checksum(out, lo=0, hi=None):
    """Order-sensitive 64-bit checksum of the bits of a solve's results over instances [lo, hi): trajectories,
    controls, final cost and iteration counts.  Two solves agree bit for bit iff (up to hash collisions) their
    checksums agree; used to compare a slice solved inside a big multi-GPU batch with the same slice solved alone."""
    import numpy as np
    acc = np.uint64(1469598103934665603)
    with np.errstate(over="ignore"):
        for a in (out.x[lo:hi], out.u[lo:hi], out.J[lo:hi]):
            w = np.ascontiguousarray(a, dtype=np.float64).view(np.uint64).ravel()
            k = np.arange(1, w.size + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
            acc = (acc ^ np.bitwise_xor.reduce(w * k + (w >> np.uint64(29)))) * np.uint64(1099511628211)
        it = np.ascontiguousarray(out.iters[lo:hi], dtype=np.int64).view(np.uint64)
        acc = (acc ^ np.uint64(int((it * np.arange(1, it.size + 1, dtype=np.uint64)).sum()))) * np.uint64(1099511628211)
    return int(acc)


def gather_ints(dist, device, values):
    """all_gather of a short list of 64-bit integers per rank -> list of lists (rank order).  dist=None: [values]."""
    if dist is None:
        return [[int(v) for v in values]]
    import torch
    t = torch.tensor([int(v) - (1 << 64) if int(v) >= (1 << 63) else int(v) for v in values], dtype=torch.int64, device=device)
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t)
    return [[int(v) & ((1 << 64) - 1) for v in p.tolist()] for p in parts]
