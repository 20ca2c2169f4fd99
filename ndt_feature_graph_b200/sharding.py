"""Multi-GPU plumbing of the edge path: graph edges / scan pairs are independent units (ndt_feature_graph.cpp:349-352
loops over them serially), so a batch shards across ranks with NO data-path collective; the only exchange is the final
gather of the fixed-size result records (pose, covariance, status) — SURVEY.md §8e.

One process per GPU (torchrun); backend NCCL on GPUs, gloo in the CPU tests.  Nothing here computes: the records come
from the C ABI (ndtb_d2d_match_batch / ndtb_register_scans)."""
import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous block of `n_items` owned by `rank`: sizes differ by at most one, order preserved."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_items, world):
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def balance_by_cost(costs, world):
    """Greedy longest-processing-time assignment of edges to ranks by an a-priori cost (e.g. source cells x expected
    hits); returns a list of index arrays, each ascending.  Used when edge sizes are very uneven; the default is
    shard_range (equal-sized scans)."""
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")
    load = np.zeros(world)
    owner = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        owner[r].append(int(i))
        load[r] += float(costs[i])
    return [np.array(sorted(o), dtype=np.int64) for o in owner]


def deal_by_cost(costs, world):
    """Equal-COUNT shards of equal total cost: items in descending cost, dealt boustrophedon over the ranks (0..w-1, w-1..0,
    ...).  len(costs) must be a multiple of world.  Returns owner[i] = rank of item i.  bench.py deals the world x B scan
    pairs of a multi-GPU run this way from the derivative passes of an untimed run: the step time is the MAX over ranks,
    and a rank that drew more registrations that run into ITR_MAX would set it."""
    costs = np.asarray(costs)
    n = costs.shape[0]
    if n % world:
        raise ValueError("deal_by_cost: item count must be a multiple of the world size")
    order = np.argsort(-costs, kind="stable")
    pos = np.arange(n)
    q, r = pos // world, pos % world
    owner = np.empty(n, np.int64)
    owner[order] = np.where(q % 2 == 0, r, world - 1 - r)
    return owner


def gather_results(local_records, n_total, rank, world, device=None):
    """All-gather of per-edge result records.  `local_records`: structured numpy array (api.RESULT_DTYPE) or a uint8
    torch tensor already on `device`, holding this rank's shard_range block.  Returns the n_total records in edge order
    (numpy structured array when given numpy, else a uint8 tensor).  Ranks may own blocks that differ by one record:
    blocks are padded to the largest and trimmed after the collective."""
    import torch
    import torch.distributed as dist

    is_np = isinstance(local_records, np.ndarray)
    if is_np:
        dtype = local_records.dtype
        rec = dtype.itemsize
        t = torch.from_numpy(np.frombuffer(local_records.tobytes(), dtype=np.uint8).copy())
        if device is not None:
            t = t.to(device)
    else:
        t = local_records
        rec = t.numel() // max(1, shard_sizes(n_total, world)[rank]) if t.numel() else 0
    if world == 1:
        return local_records
    sizes = shard_sizes(n_total, world)
    if is_np is False and rec == 0:
        raise ValueError("empty tensor shard: record size unknown")
    if min(sizes) == max(sizes):  # equal blocks (bench.py, weak scaling): one collective straight into the output
        out = torch.empty(world * t.numel(), dtype=torch.uint8, device=t.device)
        dist.all_gather_into_tensor(out, t)
        return np.frombuffer(out.cpu().numpy().tobytes(), dtype=dtype).copy() if is_np else out
    cap = max(sizes) * rec
    pad = torch.zeros(cap, dtype=torch.uint8, device=t.device)
    pad[: t.numel()] = t
    out = torch.empty(world * cap, dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, pad)
    parts = [out[r * cap: r * cap + sizes[r] * rec] for r in range(world)]
    full = torch.cat(parts)
    if is_np:
        return np.frombuffer(full.cpu().numpy().tobytes(), dtype=dtype).copy()
    return full
