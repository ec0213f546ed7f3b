"""Host emulation of the GPU pipeline (test infrastructure; see emu.cpp)."""
import ctypes
import os
import subprocess

import numpy as np
import pyarrow as pa

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "..", "pyruhvro_b200", "csrc")
SO = os.path.join(HERE, "libemu.so")
SRCS = [os.path.join(HERE, "emu.cpp")] + [os.path.join(CSRC, f) for f in ("schema.cpp", "plan.cpp", "result.cpp")]
DEPS = SRCS + [os.path.join(CSRC, f) for f in ("dev_core.cuh", "dev_types.h", "interp.cuh", "plan.hpp", "result.hpp", "schema.hpp", "json.hpp")]


class EmuError(ValueError):
    def __init__(self, code, record, msg=""):
        super().__init__(f"emu error {code} at record {record} {msg}")
        self.code, self.record = code, record


def build(gen_source: str = None):
    """gen_source=None: the interpreter walker.  Otherwise the generated, schema-specialised walker
    (what NVRTC compiles for the GPU) is compiled for the host and driven by the same emulation."""
    if gen_source is None:
        so, extra = SO, []
    else:
        import hashlib
        h = hashlib.sha1(gen_source.encode()).hexdigest()[:16]
        gdir = os.path.join(HERE, "_gen")
        os.makedirs(gdir, exist_ok=True)
        hdr = os.path.join(gdir, f"walker_{h}.cuh")
        if not os.path.exists(hdr):
            with open(hdr, "w") as f:
                f.write(gen_source)
        so = os.path.join(gdir, f"libemu_{h}.so")
        extra = ["-I", CSRC, f'-DEMU_GEN_WALKER="{hdr}"']
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in DEPS):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-Wall", "-x", "c++"] + extra + ["-o", so] + SRCS)
    return so


_libs = {}


def walker_source(schema_json: str) -> str:
    import pyruhvro_b200 as pr
    return pr.Schema(schema_json).walker_source


def decode(schema_json: str, data, offsets, n: int, num_chunks: int = 1, walker: str = "interp"):
    from pyruhvro_b200 import _ArrowArray, _ArrowSchema
    so = build(walker_source(schema_json) if walker == "gen" else None)
    _lib = _libs.get(so)
    if _lib is None:
        _lib = _libs[so] = ctypes.CDLL(so)
        _lib.emu_decode.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64),
                                    ctypes.c_char_p, ctypes.c_size_t]
    data = np.ascontiguousarray(data, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    k_max = min(max(num_chunks, 1), max(n, 1))
    arrs = (_ArrowArray * k_max)()
    sch = _ArrowSchema()
    k = ctypes.c_int64(0)
    rec = ctypes.c_int64(-1)
    msg = ctypes.create_string_buffer(512)
    raw = schema_json.encode()
    rc = _lib.emu_decode(raw, len(raw), data.ctypes.data if data.size else None, offsets.ctypes.data, n, num_chunks,
                         ctypes.addressof(arrs), ctypes.addressof(sch), ctypes.byref(k), ctypes.byref(rec), msg, 512)
    if rc != 0:
        raise EmuError(rc, rec.value, msg.value.decode())
    schema = pa.Schema._import_from_c(ctypes.addressof(sch))
    return [pa.RecordBatch._import_from_c(ctypes.addressof(arrs[i]), schema) for i in range(k.value)]
