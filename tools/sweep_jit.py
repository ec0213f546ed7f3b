"""Device-resident decode timing under different JIT code-generation knobs, in ONE process.

    python tools/sweep_jit.py [--workload kafka] [--records 10000000] VAR=a,b,c [VAR2=x,y] ...

Each combination of the listed environment variables gets a fresh Schema handle (the knobs are read at kernel
generation time and are part of the cubin cache key), 3 warm-up + 20 timed rv_decode_device calls, and one line
with the fused kernel's time from rv_last_timings.  Development tool; needs a GPU.
"""
import argparse
import ctypes
import itertools
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="kafka")
    ap.add_argument("--records", type=int, default=10_000_000)
    ap.add_argument("--num-chunks", type=int, default=8)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("knobs", nargs="*")
    args = ap.parse_args()
    import torch
    import pyruhvro_b200 as pr
    import workloads
    L = pr.lib
    n = args.records
    sj, h_data, h_off = workloads.generate(args.workload, n, seed=42)
    total = int(h_off[n])
    d_data = torch.empty(total + 64, dtype=torch.uint8, device="cuda")
    d_off = torch.from_numpy(h_off).cuda()
    d_data[:total].copy_(torch.from_numpy(h_data))
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    names = [k.split("=", 1)[0] for k in args.knobs]
    values = [k.split("=", 1)[1].split(",") for k in args.knobs]
    for combo in itertools.product(*values) if names else [()]:
        for k, v in zip(names, combo):
            os.environ[k] = v
        schema = pr.Schema(sj)
        tb = (ctypes.c_float * 6)()
        kt = np.zeros(6)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(3 + args.steps):
            if i == 3:
                torch.cuda.synchronize()
                e0.record(stream)
            h = ctypes.c_void_p()
            rc = L.rv_decode_device(schema.handle, d_data.data_ptr(), d_off.data_ptr(), n, args.num_chunks,
                                    stream.cuda_stream, ctypes.byref(h))
            if rc:
                raise SystemExit(pr._last_error())
            if i >= 3:
                L.rv_last_timings(tb, 6)
                kt += np.frombuffer(tb, dtype=np.float32)
            L.rv_result_free(h)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        kt /= args.steps
        print(dict(zip(names, combo)), f"step {ms:.3f} ms  {n / ms / 1e6:.2f} G rec/s  fused {kt[0]:.3f}  extra-pass {kt[1]:.3f}  "
              f"nullcount {kt[3]:.3f}  walker {pr.last_walker()}  passes {L.rv_last_passes()}  slow tiles {L.rv_last_slow_tiles()}", flush=True)


if __name__ == "__main__":
    main()
