"""Helpers for the multi-process tests of pyruhvro_b200.distributed (test infrastructure)."""
import numpy as np
import pyarrow as pa
import torch

from oracle import pyoracle as po


class NumpyOps:
    """CPU stand-ins for the two device fix-up kernels, used ONLY by the gloo tests of the gather's host
    logic (the product's ops are CudaOps and refuse CPU tensors)."""

    def rebase_i32(self, dst, src, count, add):
        d = dst.numpy().view(np.int32)
        s = src.numpy().view(np.int32)
        d[:count] = s[:count] + np.int32(add)

    def concat_bits(self, dst_words, dst_bit, src_words, nbits):
        d = dst_words.numpy()
        bits = np.unpackbits(src_words.numpy(), bitorder="little")[:nbits]
        full = np.unpackbits(d, bitorder="little")
        full[dst_bit:dst_bit + nbits] |= bits
        d[:] = np.packbits(full, bitorder="little")[: d.size]


def _pad4(b: bytes) -> torch.Tensor:
    n = (len(b) + 3) // 4 * 4
    a = np.zeros(n, dtype=np.uint8)
    a[: len(b)] = np.frombuffer(b, dtype=np.uint8)
    return torch.from_numpy(a)


def flat_from_canon(cols, schema: pa.Schema):
    """Oracle canonical columns -> the flat description gather_batch consumes (CPU tensors)."""
    out = []

    def rec(c, t):
        d = {"type": t, "rows": c["length"], "null_count": c["null_count"], "validity": None, "bufs": []}
        out.append(d)
        if c["validity"] is not None and c["length"] > 0:
            d["validity"] = _pad4(c["validity"])
        if c["kind"] == "bool":
            d["bufs"] = [_pad4(c["buffers"][0])]
        else:
            d["bufs"] = [torch.from_numpy(np.frombuffer(b, dtype=np.uint8).copy()) for b in c["buffers"]]
        from pyruhvro_b200.distributed import _children_types
        for ch, ct in zip(c["children"], _children_types(t)):
            rec(ch, ct)

    for i, c in enumerate(cols):
        rec(c, schema.field(i).type)
    return out
