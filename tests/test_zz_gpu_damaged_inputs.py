"""Damaged inputs through the CUDA path (C ABI: rv_decode_host) against the C oracle: the same buffers, or the same first
failing record with the same error category (the rv_status).  The host-emulation twin is tests/test_mutation_fuzz.py."""
import ctypes
import re

import numpy as np
import pytest

import pyruhvro_b200 as pr
from oracle import pyoracle as po
from tests import mutation as M
from tests.test_gpu_parity import JIT_SEEDS

pytestmark = pytest.mark.gpu


class GpuDecodeError(Exception):
    def __init__(self, status, message):
        super().__init__(message)
        m = re.search(r"\(record (-?\d+)\)", message)
        self.category, self.record = po.ERR_NAMES.get(status, str(status)), int(m.group(1)) if m else -1


def _gpu(sj, data, off, n, k):
    s = pr._get_or_parse_schema(sj)
    data = np.ascontiguousarray(data, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    h = ctypes.c_void_p()
    rc = pr.lib.rv_decode_host(s.handle, data.ctypes.data if data.size else None, off.ctypes.data, n, k, ctypes.byref(h))
    if rc:
        raise GpuDecodeError(rc, pr._last_error())
    return pr._export_batches(h.value, s)


def _gpu_error(e):
    return (e.category, e.record) if isinstance(e, GpuDecodeError) else None


@pytest.fixture(params=["jit", "interp"])
def walker(request):
    pr.set_jit_enabled(1 if request.param == "jit" else 0)
    yield request.param
    pr.set_jit_enabled(-1)


@pytest.mark.timeout(180, method="thread")
def test_forged_block_count_ends_with_the_records_bytes(coracle, walker):
    """The record the fuzz found (tests/mutation.py): a forged map count of i64::MAX whose item ends in a union.  Before the
    fix of item_stop (dev_core.cuh) the fast COUNT walk of that lane went on for 2^63 items — on a GPU, a hung kernel."""
    recs = [M.HANG_RECORD[:40]] * 300 + [M.HANG_RECORD] + [M.HANG_RECORD[:40]] * 200
    want = M.expected(coracle, M.HANG_SCHEMA, recs)
    data, off = po.pack_records(recs)
    with pytest.raises(GpuDecodeError) as e:
        _gpu(M.HANG_SCHEMA, data, off, len(recs), 3)
    assert _gpu_error(e.value) == want
    assert pr.last_walker() == walker


def test_block_count_of_i64_min_is_an_empty_block(coracle, walker):
    sj, r = M.MIN_BLOCK_SCHEMA, M.MIN_BLOCK_RECORD
    recs = [r] * 600
    data, off = po.pack_records(recs)
    from tests.parity import assert_matches_oracle
    assert_matches_oracle(coracle, _gpu(sj, data, off, len(recs), 2), sj, data, off, len(recs), 2)


@pytest.mark.timeout(600, method="thread")
def test_damaged_batches(coracle):
    """120 damaged batches over random schemas through the interpreter kernel, 40 through generated walkers (the schemas
    tools/warm_jit_cache.py precompiles)."""
    seen = {"decoded": 0, "error": 0}
    try:
        pr.set_jit_enabled(0)
        for seed in range(930000, 930120):
            sj, recs, k = M.damaged_case(seed)
            if pr.Schema(sj).is_supported:
                seen[M.check(coracle, _gpu, _gpu_error, sj, recs, k)] += 1
        pr.set_jit_enabled(1)
        for seed in range(940000, 940040):
            sj, recs, k = M.damaged_case(seed, schema_seed=JIT_SEEDS[seed % len(JIT_SEEDS)])
            seen[M.check(coracle, _gpu, _gpu_error, sj, recs, k)] += 1
            assert pr.last_walker() == "jit"
    finally:
        pr.set_jit_enabled(-1)
    assert seen["decoded"] > 10 and seen["error"] > 10


def test_bytearray_elements_decode_like_bytes(coracle):
    """PyBackedBytes (src/lib.rs:29-33) takes `bytes` or `bytearray` elements; both gather into the same packed buffer."""
    from tests.parity import assert_matches_oracle, gen_case
    sj, recs, data, off = gen_case(7, n=300)
    mixed = [bytearray(r) if i % 2 else r for i, r in enumerate(recs)]
    assert_matches_oracle(coracle, pr.deserialize_array_threaded(mixed, sj, 3), sj, data, off, len(recs), 3)
