"""world_size-2 gloo tests of the multi-GPU gather's host logic (shard partition, size exchange,
offset rebasing order, bitmap seams, lazily-absent validity).  The shard-local batches come from the
oracle and the two device fix-up kernels are replaced by numpy stand-ins from tests/dist_helpers.py;
the real NCCL + CUDA path is exercised by tests/dist_gpu_check.py on a multi-GPU box."""
import os
import random

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyoracle as po
from pyruhvro_b200 import distributed as D
from tests.parity import expected_schema, gen_case


def test_shard_bounds_cover_and_align():
    for n, w in [(10_000_000, 8), (1000, 2), (257, 2), (5, 4), (0, 2), (100_000_003, 8)]:
        b = [D.shard_bounds(n, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        for (a0, a1), (b0, b1) in zip(b, b[1:]):
            assert a1 == b0 and a0 <= a1
        assert all(x[0] % 256 == 0 for x in b if x[0] < n or n == 0)


def _worker(rank, world, port, seeds, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.dist_helpers import NumpyOps, flat_from_canon
    co = po.COracle()
    try:
        for seed in seeds:
            sj, recs, data, off = gen_case(seed, n=random.Random(seed).choice([600, 1000, 1301]))
            n = len(recs)
            schema = po.to_arrow_schema(po.parse_schema(sj))
            r0, r1 = D.shard_bounds(n, world, rank)
            local_cols = co.decode(sj, recs[r0:r1])
            local = flat_from_canon(local_cols, schema)
            batch = D.gather_batch(local, schema, ops=NumpyOps(), device="cpu")
            want = co.decode(sj, recs)
            batch.validate(full=True)
            got = po.canon_from_batch(batch)
            # a shard-local lazily-absent bitmap may become present after the gather; compare logically there
            diff = po.canon_diff(got, want)
            if diff is not None and "validity presence" not in diff:
                raise AssertionError(f"seed {seed} rank {rank}: {diff}")
            assert batch.equals(po.canon_to_batch(want, schema)), f"seed {seed}: logical mismatch"
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, f"FAIL {type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


def test_gather_two_ranks_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + random.randint(0, 300)
    procs = [ctx.Process(target=_worker, args=(r, world, port, list(range(12)), q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in results), results
