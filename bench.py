#!/usr/bin/env python
"""bench.py — Avro->Arrow direct-decode throughput on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload kafka|flat|wide] [--records R]
    python bench.py --impl reference ...        # the CPU arm: the oracle port on all host threads

A "step" is one pass of the hot path over one batch of synthetic Avro records:
  value   records/s with the packed input already resident in HBM and the Arrow buffers left in
          HBM (rv_decode_device); CUDA events on the launch stream; max over ranks.
  e2e     the same batch through the reference-facing C-ABI call rv_decode_host with HOST buffers:
          H2D of the packed input (pinned), the kernels, D2H of every Arrow buffer (pinned).
  roofline  algorithmic bytes of the dominant kernel (emit_kernel: it reads every input byte and
          offset and writes every Arrow buffer byte) / its CUDA-event duration / measured HBM peak.
  cpu_baseline  the C oracle (row-at-a-time port of fast_decode.rs) on the host cores, bounded sample.
Multi-GPU: records shard by message with no data-path collective (weak scaling: R records per GPU).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "avro_to_arrow_records_per_sec"
UNIT = "records/s"
WORKLOAD_DESC = {
    "kafka": "C3: scripts/generate_avro.py Kafka schema (nullable unions, nested structs, array, map, 4-variant union, enum)",
    "flat": "C2: flat primitives (benches/common/mod.rs FLAT_PRIMITIVES)",
    "wide": "C4: four 8-variant sparse unions + three maps (divergence stress)",
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="kafka", choices=list(WORKLOAD_DESC))
    ap.add_argument("--records", type=int, default=10_000_000, help="records per GPU")
    ap.add_argument("--num-chunks", type=int, default=8, help="output batches per call (README bench: 8)")
    ap.add_argument("--cpu-sample", type=int, default=4_000_000)
    ap.add_argument("--seed", type=int, default=42)
    return ap.parse_args()


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed regions run."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.active = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10}
        while not self._stop_evt.is_set():
            if self.active.is_set():
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for k, bit in names.items():
                        if mask & bit:
                            self.reasons.add(k)
                except Exception:
                    pass
            time.sleep(0.004)

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def cpu_arm(args, workload_name, steps, warmup, emit_line):
    """The oracle port of the reference's per_datum_deserialize_threaded on all host threads."""
    import workloads
    from oracle import pyoracle as po
    co = po.COracle()
    cores = os.cpu_count() or 1
    n = min(args.cpu_sample, args.records)
    schema, data, offsets = workloads.generate(workload_name, n, seed=args.seed)
    # num_chunks is the reference's own tuning knob (its README uses 8).  Four chunks per hardware thread keep every
    # worker busy to the end and each chunk's builders cache-resident: the fastest setting for this implementation
    # (measured 13.5 M rec/s at 1 chunk/thread -> 20 M rec/s at 16-32 chunks/thread on an 8-core host)
    k = 4 * cores
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        co.decode_threaded_packed(schema, data, offsets, n, k, cores, materialize=False)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    value = n * len(times) / total
    base = {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} records of the same workload, {k} chunks on {cores} threads, {len(times)} timed passes "
                      f"(oracle/avro_oracle.c; the Rust reference cannot be built here)"}
    if emit_line:
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
                "warmup": warmup, "ms_per_step": 1000.0 * total / len(times), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": WORKLOAD_DESC[workload_name], "records_per_step": n, "num_chunks": k, "seed": args.seed},
                "cpu_baseline": base,
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
    return base


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1 and "RANK" not in os.environ:
        # convenience: re-launch under torchrun the way the driver does
        import subprocess
        port = os.environ.get("MASTER_PORT", "29517")
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", port, os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))

    steps, warmup = max(1, args.steps), max(3, args.warmup)

    if args.impl == "reference":
        if rank == 0:
            cpu_arm(args, args.workload, steps, warmup, emit_line=True)
        return

    import torch
    import torch.distributed as dist
    import pyruhvro_b200 as pr
    import workloads

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the decode path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = pr.lib

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])

    # ---- workload: this rank's shard, generated straight into pinned host memory ----------
    n = args.records
    pinned_blocks = []

    def alloc_pinned(nbytes):
        p = L.rv_host_alloc(nbytes)
        if not p:
            raise SystemExit("rv_host_alloc failed: " + pr._last_error())
        pinned_blocks.append(p)
        return np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(p))

    schema_json, h_data, h_off = workloads.generate(args.workload, n, seed=args.seed, r0=rank * n, alloc=alloc_pinned)
    total_in = int(h_off[n])
    schema = pr._get_or_parse_schema(schema_json)

    # device-resident copy of the input for the `value` measurement
    d_data = torch.empty(total_in + 64, dtype=torch.uint8, device=dev)
    d_off = torch.empty(n + 1, dtype=torch.int64, device=dev)
    d_data[:total_in].copy_(torch.from_numpy(h_data))
    d_off.copy_(torch.from_numpy(h_off))
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    vp = ctypes.c_void_p

    def step_device():
        h = vp()
        rc = L.rv_decode_device(schema.handle, d_data.data_ptr(), d_off.data_ptr(), n, args.num_chunks, stream.cuda_stream, ctypes.byref(h))
        if rc:
            raise SystemExit("rv_decode_device: " + pr._last_error())
        return h

    def step_host():
        h = vp()
        rc = L.rv_decode_host(schema.handle, h_data.ctypes.data, h_off.ctypes.data, n, args.num_chunks, ctypes.byref(h))
        if rc:
            raise SystemExit("rv_decode_host: " + pr._last_error())
        return h

    sampler = ClockSampler(local_rank if "CUDA_VISIBLE_DEVICES" not in os.environ else int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    sampler.start()

    # ---- value: device-resident ---------------------------------------------------------------
    arrow_bytes = buffer_bytes = 0
    for _ in range(warmup):
        h = step_device()
        arrow_bytes, buffer_bytes = L.rv_result_arrow_bytes(h), L.rv_result_buffer_bytes(h)
        L.rv_result_free(h)
    overflow_tiles = 0
    kt = np.zeros(6, dtype=np.float64)
    tbuf = (ctypes.c_float * 6)()
    launches = 0
    barrier()
    torch.cuda.synchronize()
    sampler.active.set()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        h = step_device()
        L.rv_last_timings(tbuf, 6)
        kt += np.frombuffer(tbuf, dtype=np.float32)
        launches += L.rv_last_launch_count()
        overflow_tiles = L.rv_last_slow_tiles()
        L.rv_result_free(h)
    e1.record(stream)
    torch.cuda.synchronize()
    sampler.active.clear()
    barrier()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    value = world * n * steps / (ms_total / 1000.0)
    kt /= steps

    # ---- e2e: host buffers in, host Arrow buffers out, through the C ABI --------------------------
    for _ in range(warmup):
        L.rv_result_free(step_host())
    barrier()
    torch.cuda.synchronize()
    sampler.active.set()
    t0 = time.perf_counter()
    h2d_ms = d2h_ms = 0.0
    for _ in range(steps):
        h = step_host()
        L.rv_last_timings(tbuf, 6)
        h2d_ms += tbuf[4]
        d2h_ms += tbuf[5]
        L.rv_result_free(h)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    sampler.active.clear()
    barrier()
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * n * steps / e2e_s
    sampler.stop()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    peak, peak_src = hbm_peak()
    idx_bytes = 8 * (n + 1)
    emit_bytes = total_in + idx_bytes + arrow_bytes      # reads every input byte + offset, writes every Arrow byte
    count_bytes = total_in + idx_bytes                   # reads every input byte + offset
    emit_gbs = emit_bytes / (kt[0] * 1e-3) / 1e9 if kt[0] > 0 else 0.0
    count_gbs = 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("workload") == args.workload and tj.get("records") == n:
                traffic = tj.get("emit_kernel_dram_bytes")
        except Exception:
            traffic = None

    cpu = cpu_arm(args, args.workload, steps=3, warmup=1, emit_line=False) if world == 1 else None

    # ---- the reverse direction, informational: Arrow -> Avro through serialize_record_batch ----
    encode = None
    if world == 1:
        try:
            batch = pr.decode_packed(h_data, h_off, n, schema_json, 1)[0]
            # steady state of a caller that keeps the latest result while asking for the next one (so two sets of
            # output slabs circulate through the pinned cache): three warm-up calls held the same way, then timed
            out = None
            for _ in range(3):
                out = pr.serialize_record_batch(batch, schema_json, args.num_chunks)
            t0 = time.perf_counter()
            reps = 5
            for _ in range(reps):
                out = pr.serialize_record_batch(batch, schema_json, args.num_chunks)
            dt = (time.perf_counter() - t0) / reps
            encode = {"value": n / dt, "unit": UNIT, "ms_per_step": 1000.0 * dt, "path": "pyruhvro.serialize_record_batch (host Arrow in, host Avro out)",
                      "avro_bytes": int(sum(a.nbytes for a in out))}
            del out, batch
        except Exception as e:  # pragma: no cover
            encode = {"error": str(e)[:200]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload], "records_per_gpu": n, "num_chunks": args.num_chunks, "seed": args.seed,
                   "input_bytes_per_gpu": total_in, "arrow_bytes_per_gpu": arrow_bytes,
                   "l2": "no flush: each step streams %.2f GB in + %.2f GB out, far larger than the 126 MB L2" % (total_in / 1e9, arrow_bytes / 1e9),
                   "sharding": "records by message, contiguous ranges per rank, no collective"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": total_in + idx_bytes, "d2h_bytes_per_step": buffer_bytes,
                "ms_per_step": 1000.0 * e2e_s / steps, "h2d_ms": h2d_ms / steps, "d2h_ms": d2h_ms / steps},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "rvj_fused" if pr.last_walker() == "jit" else "fused_kernel", "achieved": emit_gbs, "peak": peak, "unit": "GB/s",
                     "frac": emit_gbs / peak, "traffic": traffic, "algorithmic_bytes": emit_bytes, "kernel_ms": kt[0],
                     "peak_source": peak_src, "path_frac": emit_bytes / (ms_total / steps * 1e-3) / 1e9 / peak,
                     "extra_pass_ms": kt[1], "null_count_kernel_ms": kt[3], "walker": pr.last_walker(),
                     "slow_tiles_per_step": overflow_tiles},
        "clocks": sampler.summary(),
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if encode is not None:
        line["encode"] = encode
    print(json.dumps(line), flush=True)
    for p in pinned_blocks:
        L.rv_host_free(p)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
