"""Times serialize_record_batch on a decoded batch (pinned buffers) and on a pageable copy of it.
    RV_TRACE=1 python tools/encode_probe.py [records]      (development tool; needs a GPU)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyarrow as pa

import pyruhvro_b200 as pr
import workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
sj, data, off = workloads.generate("kafka", n, seed=42)
batch = pr.decode_packed(data, off, n, sj, 1)[0]
sink = pa.BufferOutputStream()
with pa.ipc.new_stream(sink, batch.schema) as w:
    w.write_batch(batch)
pageable = pa.ipc.open_stream(sink.getvalue()).read_next_batch()   # same batch in ordinary heap memory
for name, b in (("pinned", batch), ("pageable", pageable)):
    for i in range(4):
        t0 = time.perf_counter()
        out = pr.serialize_record_batch(b, sj, 8)
        t1 = time.perf_counter()
        del out
        t2 = time.perf_counter()
        print(f"{name} call {i}: encode {1e3 * (t1 - t0):.1f} ms, free {1e3 * (t2 - t1):.1f} ms", flush=True)
