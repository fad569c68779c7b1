"""Multi-process (gloo, world_size 2 and 3) test of the N>1 host logic on CPU: sequence-chunk
sharding with k-1 overlap + one integer all-reduce reproduces the unsharded currents.  The CPU
oracle stands in for the per-rank count kernel (the GPU kernel itself is covered by -m gpu)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import random_dna

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, k, pool, seed, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from neurokmer_b200.counter import flatten
    from neurokmer_b200.shard import shard_batch
    from oracle.oracle_py import COracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    seqs = [random_dna(rng, n, 0.01, 0.02) for n in (5000, 0, 17, k - 1, k, 12001, 333, 40000, 29, 31)]
    bases, offsets = flatten(seqs)
    c = COracle()
    b_r, o_r = shard_batch(bases, offsets, k, world, rank)
    cur, tot = c.accumulate(b_r, o_r, k, pool, True)
    t = torch.from_numpy(cur.astype(np.int64))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    n = torch.tensor([tot]); dist.all_reduce(n)
    if rank == 0:
        full, ftot = c.accumulate(bases, offsets, k, pool, True)
        q.put((bool((t.numpy().astype(np.uint64) == full).all()), int(n.item()), int(ftot)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,k", [(2, 31), (3, 21), (2, 1)])
def test_sharded_currents_allreduce(world, k):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, k, 5003, 99, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    same, tot, ftot = q.get(timeout=5)
    assert same and tot == ftot


def test_shard_partition_properties():
    """Every window start is owned by exactly one rank, for ragged batches and any world size."""
    from neurokmer_b200.counter import flatten
    from neurokmer_b200.shard import shard_batch
    rng = np.random.default_rng(3)
    for k in (1, 5, 31, 32):
        seqs = [random_dna(rng, int(n)) for n in rng.integers(0, 300, size=40)]
        bases, offsets = flatten(seqs)
        want = sum(max(0, len(s) - k + 1) for s in seqs)
        for world in (1, 2, 3, 7, 8, 64):
            got = 0
            for r in range(world):
                b, o = shard_batch(bases, offsets, k, world, r)
                lens = np.diff(o.astype(np.int64))
                got += int(np.maximum(lens - k + 1, 0).sum())
                assert b.size == int(o[-1])
            assert got == want, (k, world)
