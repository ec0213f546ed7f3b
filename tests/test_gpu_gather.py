"""GPU tests of the multi-GPU gather (pyruhvro_b200.distributed over the C ABI's rv_gather_* entry points): the size
exchange, the plan, the push kernel (offset rebase + bitmap shift fused into the copy) and the result hand-over.  With one
GPU the gather runs as a world of one (the push kernel copies into the rank's own arena); with >= 2 visible GPUs two
NCCL ranks are spawned and the non-leader pushes into the leader's arena through CUDA IPC peer memory over NVLink."""
import os
import random
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard(name, n, seed, world, rank, dev):
    import torch
    import workloads
    from pyruhvro_b200 import distributed as D
    r0, r1 = D.shard_bounds(n, world, rank)
    sj, data, off = workloads.generate(name, r1 - r0, seed=seed, r0=r0)
    d_data = torch.zeros(len(data) + 64, dtype=torch.uint8, device=dev)
    d_data[: len(data)].copy_(torch.from_numpy(np.ascontiguousarray(data)))
    return sj, d_data, torch.from_numpy(off).to(dev), r1 - r0


def _check_against_oracle(name, n, seed, batches):
    import workloads
    from oracle import pyoracle as po
    from tests.parity import expected_schema
    sj, data, off = workloads.generate(name, n, seed=seed)
    assert len(batches) == 1 and batches[0].num_rows == n
    assert batches[0].schema.equals(expected_schema(sj), check_metadata=True)
    diff = po.canon_diff(po.canon_from_batch(batches[0]), po.COracle().decode_packed(sj, data, off, n))
    assert diff is None, diff


def test_gather_world_of_one():
    import torch
    from pyruhvro_b200 import distributed as D
    dev = torch.device("cuda", 0)
    for name, n in [("kafka", 100_003), ("wide", 20_001), ("flat", 70_000), ("array_map", 33_333)]:
        sj, d_data, d_off, n_local = _shard(name, n, 5, 1, 0, dev)
        out = D.decode_and_gather(sj, d_data, d_off, n_local, to_host=True)
        assert out["n_batches"] == 1 and out["remote_bytes"] == 0
        _check_against_oracle(name, n, 5, out["batches"])


def _rank_main(rank, world, port, q):
    try:
        os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank)})
        sys.path.insert(0, ROOT)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", device_id=dev)
        from pyruhvro_b200 import distributed as D
        log = []
        for name, n in [("kafka", 100_003), ("wide", 50_001), ("flat", 70_000), ("array_map", 33_333), ("kafka", 517)]:
            sj, d_data, d_off, n_local = _shard(name, n, 5, world, rank, dev)
            out = D.decode_and_gather(sj, d_data, d_off, n_local, to_host=True)
            if rank == 0:
                _check_against_oracle(name, n, 5, out["batches"])
                log.append(f"{name} n={n} world={world}: OK ({out['remote_bytes']} bytes pushed over NVLink)")
            else:
                assert out["batches"] == []
        # timing at a larger size (C5 shape, scaled): device-resident gather
        n = 4_000_000 * world
        sj, d_data, d_off, n_local = _shard("kafka", n, 42, world, rank, dev)
        best = None
        for _ in range(4):
            out = D.decode_and_gather(sj, d_data, d_off, n_local)
            best = out if best is None or out["gather_ms"] < best["gather_ms"] else best
        if rank == 0:
            log.append(f"kafka {n} rows over {world} GPUs: shard decode {best['decode_ms']:.2f} ms, gather {best['gather_ms']:.2f} ms, "
                       f"{best['remote_bytes'] / 1e9:.2f} GB over NVLink = {best['remote_bytes'] / best['gather_ms'] / 1e6:.0f} GB/s")
        dist.barrier(device_ids=[rank])
        dist.destroy_process_group()
        q.put((rank, "ok", log))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL " + traceback.format_exc()[-1500:], []))


def test_gather_two_gpus_nccl_ipc():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + random.randint(0, 200)
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in results), results
    lines = [l for r in results for l in r[2]]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gather_2gpu.log"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))
