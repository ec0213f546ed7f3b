#!/usr/bin/env python
"""Multi-GPU check of pyruhvro_b200.distributed (run under torchrun on a box with >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_gpu_check.py

Every rank decodes its shard on its GPU, the ranks all-gather the Arrow buffers over NCCL and fix them up on
the device; the single gathered RecordBatch is compared buffer-for-buffer with the oracle's decode of the whole
input.  Also prints the gather's timing for a larger shard (config C5 shape, scaled)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import workloads
    from oracle import pyoracle as po
    from pyruhvro_b200 import distributed as D
    from tests.parity import expected_schema

    def shard(name, n, seed):
        r0, r1 = D.shard_bounds(n, world, rank)
        sj, data, off = workloads.generate(name, r1 - r0, seed=seed, r0=r0)
        d_data = torch.from_numpy(np.ascontiguousarray(data)).to(dev)
        d_data = torch.cat([d_data, torch.zeros(64, dtype=torch.uint8, device=dev)])
        d_off = torch.from_numpy(off).to(dev)
        return sj, d_data, d_off, r1 - r0

    ok = True
    for name, n in [("kafka", 100_003), ("wide", 50_001), ("flat", 70_000), ("array_map", 33_333)]:
        sj, d_data, d_off, n_local = shard(name, n, 5)
        batch = D.decode_sharded_gather(sj, d_data, d_off, n_local)
        assert batch.num_rows == n
        if rank == 0:
            _, data, off = workloads.generate(name, n, seed=5)
            want = po.COracle().decode_packed(sj, data, off, n)
            diff = po.canon_diff(po.canon_from_batch(batch), want)
            assert batch.schema.equals(expected_schema(sj), check_metadata=True)
            print(f"[gather] {name} n={n} world={world}: {'OK' if diff is None else 'DIFF ' + diff}", flush=True)
            ok &= diff is None
    # timing at a larger size
    n = 4_000_000 * world
    sj, d_data, d_off, n_local = shard("kafka", n, 42)
    for it in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s, h = D.decode_sharded(sj, d_data, d_off, n_local)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        from pyruhvro_b200 import lib
        lib.rv_result_free(h)
        batch = D.decode_sharded_gather(sj, d_data, d_off, n_local)
        t2 = time.perf_counter()
        if rank == 0 and it == 2:
            print(f"[gather] kafka {n} rows over {world} GPUs: shard decode {1e3 * (t1 - t0):.1f} ms, decode+all-gather+fix-up+D2H+assemble "
                  f"{1e3 * (t2 - t1):.1f} ms -> one RecordBatch of {batch.num_rows} rows, {batch.nbytes / 1e9:.2f} GB", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("[gather] ALL OK" if ok else "[gather] FAILED", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
