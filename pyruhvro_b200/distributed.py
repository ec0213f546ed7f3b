"""Multi-GPU decode: one process per GPU (torch.distributed), records sharded by message.

The reference's own contract needs no collective: `per_datum_deserialize_threaded` returns one
RecordBatch per contiguous chunk (ruhvro/src/deserialize.rs:57-68,115-119), and rank g decoding rows
`shard_bounds(n, world, g)` yields exactly those batches.  `decode_sharded` is that path.

When ONE RecordBatch over all ranks' rows is wanted, `gather_batch` all-gathers every Arrow buffer
(NCCL over NVLink/NVSwitch; every peer is one hop away, so a flat all-gather is the right shape) and
fixes the pieces up ON THE DEVICE with two kernels of the C ABI: `rv_dev_rebase_i32` (a shard's i32
offsets += totals of the shards before it) and `rv_dev_concat_bits` (bit-level concatenation of
validity / boolean bitmaps whose row counts are not byte multiples).  Arrow's i32 offsets cap a single
batch at 2^31-1 bytes per Utf8 column; beyond that the gather raises, as arrow-rs would panic.

torch is used for what it is here for: device tensors, streams and the process group.
"""
from __future__ import annotations

import ctypes
from typing import Callable, List, Optional

import numpy as np
import pyarrow as pa

INT32_MAX = 2**31 - 1


def shard_bounds(n: int, world: int, rank: int, align: int = 256):
    """Contiguous row range of `rank`: build_slices semantics (floor division, remainder to the last
    rank) with shard starts aligned to `align` rows so bitmaps concatenate on word boundaries."""
    base = (n // world) // align * align if world > 1 else n
    if base == 0:  # tiny inputs: everything on the last rank
        return (0, 0) if rank < world - 1 else (0, n)
    r0 = rank * base
    return (r0, n) if rank == world - 1 else (r0, r0 + base)


# --------------------------------------------------------------------------------------------------
# flat description of a batch: pre-order list of arrays, each {type, rows, null_count, validity, bufs}
# --------------------------------------------------------------------------------------------------
_FIXED_W = {pa.int32(): 4, pa.int64(): 8, pa.float32(): 4, pa.float64(): 8, pa.date32(): 4}


def _width(t: pa.DataType) -> int:
    if pa.types.is_timestamp(t):
        return 8
    return _FIXED_W[t]


def _children_types(t: pa.DataType) -> List[pa.DataType]:
    if pa.types.is_struct(t) or pa.types.is_union(t):
        return [t.field(i).type for i in range(t.num_fields)]
    if pa.types.is_map(t):
        return [pa.struct([t.key_field, t.item_field])]
    if pa.types.is_list(t):
        return [t.value_type]
    return []


class _DevMem:
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (max(nbytes, 0),), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _dev_tensor(ptr: int, nbytes: int):
    import torch
    if nbytes <= 0 or not ptr:
        return torch.empty(0, dtype=torch.uint8, device="cuda")
    return torch.as_tensor(_DevMem(ptr, nbytes), device="cuda")


def describe_device_batch(arr_addr: int, schema: pa.Schema) -> List[dict]:
    """Walks an exported ArrowDeviceArray (struct array of the columns) into the flat description with
    zero-copy uint8 CUDA tensors over its buffers."""
    from . import _ArrowArray
    out: List[dict] = []

    def rec(addr: int, t: pa.DataType):
        a = _ArrowArray.from_address(addr)
        rows = a.length
        bufs = (ctypes.c_void_p * a.n_buffers).from_address(a.buffers) if a.n_buffers else []
        kids = (ctypes.c_void_p * a.n_children).from_address(a.children) if a.n_children else []
        d = {"type": t, "rows": rows, "null_count": a.null_count, "validity": None, "bufs": []}
        out.append(d)
        has_validity_slot = not (pa.types.is_null(t) or pa.types.is_union(t))
        if has_validity_slot and bufs[0]:
            d["validity"] = _dev_tensor(bufs[0], ((rows + 31) // 32) * 4)
        if pa.types.is_boolean(t):
            d["bufs"] = [_dev_tensor(bufs[1], ((rows + 31) // 32) * 4)]
        elif pa.types.is_string(t):
            off = _dev_tensor(bufs[1], 4 * (rows + 1))
            last = int(off[4 * rows:4 * rows + 4].view(dtype=__import__("torch").int32).item()) if rows >= 0 else 0
            d["bufs"] = [off, _dev_tensor(bufs[2], last)]
        elif pa.types.is_list(t) or pa.types.is_map(t):
            d["bufs"] = [_dev_tensor(bufs[1], 4 * (rows + 1))]
        elif pa.types.is_union(t):
            d["bufs"] = [_dev_tensor(bufs[0], rows)]
        elif pa.types.is_struct(t) or pa.types.is_null(t):
            pass
        else:
            d["bufs"] = [_dev_tensor(bufs[1], rows * _width(t))]
        for k, ct in zip(kids, _children_types(t)):
            rec(k, ct)

    top = _ArrowArray.from_address(arr_addr)
    cols = (ctypes.c_void_p * top.n_children).from_address(top.children)
    for i in range(top.n_children):
        rec(cols[i], schema.field(i).type)
    return out


class CudaOps:
    """Device fix-ups through the C ABI (rv_dev_rebase_i32 / rv_dev_concat_bits) on torch's current stream."""

    def __init__(self):
        from . import lib
        self.lib = lib
        lib.rv_dev_rebase_i32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p]
        lib.rv_dev_concat_bits.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]

    @staticmethod
    def _stream():
        import torch
        return torch.cuda.current_stream().cuda_stream

    def rebase_i32(self, dst, src, count: int, add: int):
        from . import _check
        if not (dst.is_cuda and src.is_cuda):
            raise ValueError("gather fix-ups run on the GPU only")
        _check(self.lib.rv_dev_rebase_i32(dst.data_ptr(), src.data_ptr(), count, add, self._stream()))

    def concat_bits(self, dst_words, dst_bit: int, src_words, nbits: int):
        from . import _check
        if not (dst_words.is_cuda and src_words.is_cuda):
            raise ValueError("gather fix-ups run on the GPU only")
        _check(self.lib.rv_dev_concat_bits(dst_words.data_ptr(), dst_bit, src_words.data_ptr(), nbits, self._stream()))


def _all_gather_var(t, sizes: List[int], group, device):
    """All-gather of uint8 tensors of different lengths: pad to the maximum (shards are balanced, so the
    padding is small) and slice.  Returns the list of per-rank tensors."""
    import torch
    import torch.distributed as dist
    world = len(sizes)
    mx = (max(sizes) + 15) // 16 * 16
    if mx == 0:
        return [torch.empty(0, dtype=torch.uint8, device=device) for _ in sizes]
    send = torch.zeros(mx, dtype=torch.uint8, device=device)
    send[: t.numel()] = t
    recv = torch.empty(world * mx, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(recv, send, group=group)
    return [recv[g * mx: g * mx + sizes[g]] for g in range(world)]


def gather_batch(local: List[dict], schema: pa.Schema, group=None, ops=None, device="cuda") -> pa.RecordBatch:
    """All ranks call this with the flat description of their shard's batch; every rank gets the single
    RecordBatch over all shards' rows (in rank order)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    ops = ops or CudaOps()
    meta_local = [{"rows": d["rows"], "nulls": d["null_count"], "validity": d["validity"] is not None,
                   "sizes": [int(b.numel()) for b in d["bufs"]]} for d in local]
    metas: List[Optional[list]] = [None] * world
    dist.all_gather_object(metas, meta_local, group=group)

    final: List[dict] = []
    for i, d in enumerate(local):
        t = d["type"]
        rows = [metas[g][i]["rows"] for g in range(world)]
        row_base = np.concatenate([[0], np.cumsum(rows)]).tolist()
        total_rows = row_base[-1]
        f = {"type": t, "rows": total_rows, "null_count": sum(metas[g][i]["nulls"] for g in range(world)), "validity": None, "bufs": []}
        # ---- validity: present if any shard has one (an absent shard bitmap means all valid) ----
        if any(metas[g][i]["validity"] for g in range(world)) and not (pa.types.is_null(t) or pa.types.is_union(t)):
            nwords = (total_rows + 31) // 32
            dst = torch.zeros(nwords * 4, dtype=torch.uint8, device=device)
            vsizes = [((rows[g] + 31) // 32) * 4 if metas[g][i]["validity"] else 0 for g in range(world)]
            mine = d["validity"] if d["validity"] is not None else torch.empty(0, dtype=torch.uint8, device=device)
            parts = _all_gather_var(mine, vsizes, group, device)
            ones = None
            for g in range(world):
                if rows[g] == 0:
                    continue
                src = parts[g]
                if not metas[g][i]["validity"]:
                    if ones is None or ones.numel() < ((rows[g] + 31) // 32) * 4:
                        ones = torch.full((((max(rows) + 31) // 32) * 4,), 0xFF, dtype=torch.uint8, device=device)
                    src = ones
                ops.concat_bits(dst, row_base[g], src, rows[g])
            f["validity"] = dst
        # ---- data buffers ----
        if pa.types.is_null(t):
            f["null_count"] = total_rows
        for bi in range(len(d["bufs"])):
            sizes = [metas[g][i]["sizes"][bi] for g in range(world)]
            parts = _all_gather_var(d["bufs"][bi], sizes, group, device)
            is_offsets = (pa.types.is_string(t) or pa.types.is_list(t) or pa.types.is_map(t)) and bi == 0
            if pa.types.is_boolean(t):
                dst = torch.zeros(((total_rows + 31) // 32) * 4, dtype=torch.uint8, device=device)
                for g in range(world):
                    if rows[g]:
                        ops.concat_bits(dst, row_base[g], parts[g], rows[g])
            elif is_offsets:
                dst = torch.zeros(4 * (total_rows + 1), dtype=torch.uint8, device=device)
                lasts = [int(parts[g][4 * rows[g]:4 * rows[g] + 4].view(dtype=torch.int32).item()) for g in range(world)]
                add = 0
                for g in range(world):
                    if add + lasts[g] > INT32_MAX:
                        raise ValueError("Arrow i32 offset overflow: the gathered column does not fit a single RecordBatch")
                    if rows[g]:
                        ops.rebase_i32(dst[4 * (row_base[g] + 1):], parts[g][4:], rows[g], add)
                    add += lasts[g]
            else:
                dst = torch.cat(parts) if sum(sizes) else torch.empty(0, dtype=torch.uint8, device=device)
            f["bufs"].append(dst)
        final.append(f)
    if device != "cpu":
        _stage_to_pinned(final)
    return _to_arrow(final, schema)


class _PinnedBlock:
    """One pinned host slab (library cache) that backs every buffer of a gathered batch; freed when the last
    pyarrow buffer referencing it dies."""

    def __init__(self, nbytes: int):
        from . import lib
        self._lib = lib
        self.ptr = lib.rv_host_alloc(max(nbytes, 64))
        if not self.ptr:
            raise MemoryError("rv_host_alloc failed")

    def __del__(self):
        if getattr(self, "ptr", None):
            self._lib.rv_host_free(self.ptr)
            self.ptr = None


def _stage_to_pinned(flat: List[dict]) -> None:
    """Device tensors -> one pinned slab with async copies on the current stream and a single sync; the
    entries of `flat` are replaced by zero-copy pyarrow buffers over the slab."""
    import torch
    tensors = []
    for d in flat:
        if d["validity"] is not None:
            tensors.append((d, "validity", None))
        for i in range(len(d["bufs"])):
            tensors.append((d, "bufs", i))
    sizes = [((d[k] if i is None else d[k][i]).numel() + 63) // 64 * 64 for d, k, i in tensors]
    total = sum(sizes)
    block = _PinnedBlock(total)
    host = torch.from_numpy(np.ctypeslib.as_array((ctypes.c_uint8 * max(total, 1)).from_address(block.ptr)))
    off = 0
    for (d, k, i), sz in zip(tensors, sizes):
        t = d[k] if i is None else d[k][i]
        n = t.numel()
        if n:
            host[off:off + n].copy_(t, non_blocking=True)  # pinned destination: a true async D2H on the current stream
        buf = pa.foreign_buffer(block.ptr + off, n, base=block)
        if i is None:
            d[k] = buf
        else:
            d[k][i] = buf
        off += sz
    torch.cuda.current_stream().synchronize()


def _np(t) -> pa.Buffer:
    if isinstance(t, pa.Buffer):
        return t
    return pa.py_buffer(t.cpu().numpy().tobytes()) if t.numel() else pa.py_buffer(b"")


def _head(t, nbytes: int):
    return t.slice(0, nbytes) if isinstance(t, pa.Buffer) else t[:nbytes]


def _to_arrow(flat: List[dict], schema: pa.Schema) -> pa.RecordBatch:
    pos = 0

    def rec(t: pa.DataType) -> pa.Array:
        nonlocal pos
        d = flat[pos]
        pos += 1
        n, nc = d["rows"], d["null_count"]
        v = _np(_head(d["validity"], (n + 7) // 8)) if d["validity"] is not None else None
        kids = [rec(ct) for ct in _children_types(t)]
        if pa.types.is_null(t):
            return pa.nulls(n)
        if pa.types.is_union(t):
            return pa.Array.from_buffers(t, n, [None, _np(d["bufs"][0])], children=kids)
        if pa.types.is_struct(t):
            return pa.Array.from_buffers(t, n, [v], null_count=nc, children=kids)
        if pa.types.is_boolean(t):
            return pa.Array.from_buffers(t, n, [v, _np(_head(d["bufs"][0], (n + 7) // 8))], null_count=nc)
        if pa.types.is_string(t):
            return pa.Array.from_buffers(t, n, [v, _np(d["bufs"][0]), _np(d["bufs"][1])], null_count=nc)
        if pa.types.is_list(t) or pa.types.is_map(t):
            return pa.Array.from_buffers(t, n, [v, _np(d["bufs"][0])], null_count=nc, children=kids)
        return pa.Array.from_buffers(t, n, [v, _np(d["bufs"][0])], null_count=nc)

    arrays = [rec(schema.field(i).type) for i in range(len(schema))]
    return pa.RecordBatch.from_arrays(arrays, schema=schema)


# --------------------------------------------------------------------------------------------------
# public entry points
# --------------------------------------------------------------------------------------------------
def decode_sharded(schema_json: str, d_data, d_offsets, n_local: int, num_chunks: int = 1):
    """This rank's shard -> device-resident result handle (the reference's per-chunk batches; no collective)."""
    import torch
    from . import _check, _get_or_parse_schema, lib
    s = _get_or_parse_schema(schema_json)
    h = ctypes.c_void_p()
    _check(lib.rv_decode_device(s.handle, d_data.data_ptr(), d_offsets.data_ptr(), n_local, num_chunks,
                                torch.cuda.current_stream().cuda_stream, ctypes.byref(h)))
    return s, h


def decode_sharded_gather(schema_json: str, d_data, d_offsets, n_local: int, group=None) -> pa.RecordBatch:
    """Decode this rank's shard on its GPU, then all-gather into ONE RecordBatch (returned on every rank)."""
    from . import _ArrowArray, _check, lib

    class _DeviceArray(ctypes.Structure):
        _fields_ = [("array", _ArrowArray), ("device_id", ctypes.c_int64), ("device_type", ctypes.c_int32),
                    ("sync_event", ctypes.c_void_p), ("reserved", ctypes.c_int64 * 3)]

    s, h = decode_sharded(schema_json, d_data, d_offsets, n_local, 1)
    try:
        da = _DeviceArray()
        _check(lib.rv_result_export_device(h, 0, ctypes.addressof(da), None))
        try:
            local = describe_device_batch(ctypes.addressof(da.array), s.arrow_schema)
            return gather_batch(local, s.arrow_schema, group=group)
        finally:
            rel = ctypes.CFUNCTYPE(None, ctypes.c_void_p)(da.array.release)
            rel(ctypes.addressof(da.array))
    finally:
        lib.rv_result_free(h)
