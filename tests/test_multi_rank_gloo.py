"""CPU, world_size 2 over gloo: the N>1 plumbing of bench.py (gaps sharded by rank, no data-path
collective; units summed and time maxed over ranks) and the C ABI's gap partitioner."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    gaps = 3
    first = bench.rank_first_seed(100, rank, gaps)
    seqs, pairs, cells, per_gap, _ = bench.build_workload(gaps, first, config="tiny")
    seeds = [None] * world
    dist.all_gather_object(seeds, list(range(first, first + gaps)))
    my_ms, my_e2e = 10.0 * (rank + 1), 30.0 - 5.0 * rank
    tot_cells, tot_gaps, max_ms, max_e2e = bench.reduce_over_ranks(dist, torch.device("cpu"), cells, gaps, my_ms, my_e2e)
    allc = [None] * world
    dist.all_gather_object(allc, cells)
    if rank == 0:
        q.put(dict(seeds=seeds, tot_cells=tot_cells, tot_gaps=tot_gaps, max_ms=max_ms, max_e2e=max_e2e, cells=allc, n_pairs=len(pairs)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_and_reduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    flat = [s for part in res["seeds"] for s in part]
    assert len(set(flat)) == len(flat) == 6                 # disjoint shards
    assert res["tot_cells"] == sum(res["cells"]) and res["tot_gaps"] == 6
    assert res["max_ms"] == 20.0 and res["max_e2e"] == 30.0  # max over ranks
    assert res["n_pairs"] > 0


def test_partition_is_balanced_and_deterministic():
    import gappadder_b200 as g
    rng = np.random.default_rng(5)
    costs = rng.integers(1, 10_000_000, size=500).astype(np.uint64)
    for n in (1, 2, 4, 8):
        part = g.partition_gaps(costs, n)
        assert part.min() >= 0 and part.max() < n
        loads = np.array([costs[part == k].sum() for k in range(n)], dtype=np.float64)
        assert loads.max() - loads.min() <= costs.max()     # LPT bound
        assert (g.partition_gaps(costs, n) == part).all()
    assert g.estimate_gap_cells([100, 200]) == 2 * 300 * 300 + 100 * 100 + 200 * 200
    with pytest.raises(g.GpError):
        g.partition_gaps(costs, 0)
