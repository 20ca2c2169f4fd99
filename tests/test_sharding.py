"""Host-side multi-GPU logic on CPU: shard bookkeeping and the result gather over gloo with world_size 2 (the N>1
path of bench.py uses the same functions over NCCL)."""
import os
import socket

import numpy as np
import pytest

from ndt_feature_graph_b200 import api, sharding


def test_shard_range_partitions():
    for n in (0, 1, 7, 256, 1999):
        for w in (1, 2, 3, 8):
            blocks = [sharding.shard_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1 and sizes == sharding.shard_sizes(n, w)


def test_balance_by_cost():
    rng = np.random.default_rng(0)
    costs = rng.integers(1, 1000, size=257).astype(float)
    owners = sharding.balance_by_cost(costs, 8)
    allidx = np.sort(np.concatenate(owners))
    assert np.array_equal(allidx, np.arange(257))
    loads = np.array([costs[o].sum() for o in owners])
    assert loads.max() - loads.min() <= costs.max()


def test_deal_by_cost_equal_counts_and_balanced_sums():
    """the multi-GPU bench deals world x B scan pairs by measured passes: equal counts, sums within one item's spread"""
    rng = np.random.default_rng(1)
    for world in (2, 4, 8):
        B = 592
        # the C2 distribution: most registrations take 20-60 passes, ~2 % run into ITR_MAX with 200-340
        costs = rng.integers(20, 60, size=world * B)
        slow = rng.choice(world * B, size=int(0.02 * world * B), replace=False)
        costs[slow] = rng.integers(200, 340, size=slow.size)
        owner = sharding.deal_by_cost(costs, world)
        assert np.array_equal(np.bincount(owner, minlength=world), np.full(world, B))
        loads = np.array([costs[owner == r].sum() for r in range(world)])
        natural = np.array([costs[r * B:(r + 1) * B].sum() for r in range(world)])
        assert loads.max() - loads.min() <= 340 and loads.max() - loads.min() <= natural.max() - natural.min()
        heavy = np.array([(costs[owner == r] >= 200).sum() for r in range(world)])
        assert heavy.max() - heavy.min() <= 1
        assert np.array_equal(owner, sharding.deal_by_cost(costs, world))  # deterministic: every rank computes the same deal
    with pytest.raises(ValueError):
        sharding.deal_by_cost(np.ones(7), 2)


def test_workload_pairs_are_a_function_of_their_index():
    """a rank regenerates the pairs it is dealt: velodyne_batch(indices=...) returns the pairs of the contiguous call"""
    from ndt_feature_graph_b200 import synth

    a = synth.velodyne_batch(3, n_base=2, seed=0, start=4, n_rings=8, n_az=100)
    b = synth.velodyne_batch(0, n_base=2, seed=0, indices=[6, 4], n_rings=8, n_az=100)
    assert np.array_equal(a[0][2], b[0][0]) and np.array_equal(a[1][2], b[1][0]) and np.array_equal(a[2][2], b[2][0])
    assert np.array_equal(a[0][0], b[0][1]) and np.array_equal(a[3][0], b[3][1])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(n_total, rank, world)
    rec = np.zeros(hi - lo, api.RESULT_DTYPE)
    for i in range(lo, hi):  # a recognisable record per edge
        rec["T"][i - lo] = np.arange(16) + 100.0 * i
        rec["score"][i - lo] = -float(i)
        rec["iterations"][i - lo] = i
        rec["status"][i - lo] = 1 + (i % 3)
    full = sharding.gather_results(rec, n_total, rank, world)
    ok = full.shape[0] == n_total and all(
        full["iterations"][i] == i and full["score"][i] == -float(i) and full["T"][i][5] == 5 + 100.0 * i
        and full["status"][i] == 1 + (i % 3) for i in range(n_total))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 256])
def test_gather_results_gloo_world2(n_total):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == {0: True, 1: True}
