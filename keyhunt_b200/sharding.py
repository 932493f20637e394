"""Range sharding across GPUs (one process per GPU) and the tiny end-of-run hit gather.

The reference's only parallelism is a mutex-guarded cursor handing out chunks of N_SEQUENTIAL_MAX keys
(keyhunt.cpp:3314-3329; BSGS windows of 2N keys, :4600-4617).  Chunks are independent, so across GPUs the
range is split into contiguous runs of whole chunks with NO collective on the data path; only the
few-byte hit records and per-rank counters are gathered at the end (NCCL on GPUs, gloo in CPU tests).
"""

GRP = 1024  # CPU_GRP_SIZE (keyhunt.cpp:299)


def scanned_points(start, end, chunk):
    """keys the reference really scans for `-r start:end -n chunk`: every claimed chunk is scanned whole,
    so the range is rounded UP to a multiple of the chunk (keyhunt.cpp:3314-3324,:3856)"""
    if chunk <= 0 or chunk % GRP:
        raise ValueError("chunk must be a positive multiple of 1024")
    if end <= start:
        return 0
    return -(-(end - start) // chunk) * chunk


def shard_range(start, n_points, world, rank, chunk=GRP):
    """contiguous shard [s, s+n) of rank `rank`: whole chunks, earlier ranks get the remainder chunks"""
    if n_points % chunk:
        raise ValueError("n_points must be a multiple of the chunk")
    if not (0 <= rank < world):
        raise ValueError("bad rank")
    chunks = n_points // chunk
    base, rem = divmod(chunks, world)
    mine = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return start + first * chunk, mine * chunk


def shard_windows(n_windows, world, rank):
    """BSGS: contiguous block of 2N-key windows for this rank -> (first_window, count)"""
    base, rem = divmod(n_windows, world)
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def gather_hits(dist, hits):
    """all ranks -> every rank gets the merged, de-duplicated, sorted hit list.  `hits` = list of tuples."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return sorted(set(hits))
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, list(hits))
    return sorted({h for p in parts for h in p})


def max_over_ranks(dist, value, device=None):
    """device-timed durations are reported as the max over ranks"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
