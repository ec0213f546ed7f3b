"""GPU parity: the CUDA path (through the C ABI, rv_decode_host / the pyruhvro Python surface)
against the C oracle on the same bytes, buffer-for-buffer.  Needs a B200: `pytest -m gpu`."""
import random

import numpy as np
import pyarrow as pa
import pyarrow.compute  # noqa: F401
import pytest

import pyruhvro_b200 as pr
from oracle import pyoracle as po
from tests.golden import reference_datums as G
from tests.parity import assert_matches_oracle, gen_case

pytestmark = pytest.mark.gpu

JIT_SEEDS = list(range(40))       # schemas whose specialised kernels tools/warm_jit_cache.py precompiles
INTERP_SEEDS = list(range(80))


@pytest.fixture(params=["jit", "interp"])
def walker(request):
    """Runs the test once per GPU walker: NVRTC-specialised kernels and the generic interpreter kernels."""
    pr.set_jit_enabled(1 if request.param == "jit" else 0)
    yield request.param
    pr.set_jit_enabled(-1)


def test_goldens_values_and_buffers(coracle, walker):
    recs = [bytes.fromhex(h) for h in (G.G3_HEX, G.G4_HEX, G.G5_HEX)]
    b = pr.deserialize_array(recs, G.G345_SCHEMA)
    rows = b.to_pylist()
    for got, want in zip(rows, G.G345_ROWS):
        for key in ("name", "age", "emails", "address", "phone_numbers", "preferences", "status"):
            assert got[key] == want[key], key
    for sj, recs in [(G.G1_SCHEMA, [bytes.fromhex(G.G1_HEX)] * 4), (G.G2_SCHEMA, [bytes.fromhex(G.G2_HEX)]),
                     (G.G345_SCHEMA, recs)]:
        data, off = po.pack_records(recs)
        for k in (1, 2, 8):
            assert_matches_oracle(coracle, pr.deserialize_array_threaded(recs, sj, k), sj, data, off, len(recs), k)
    g1 = pr.deserialize_array([bytes.fromhex(G.G1_HEX)] * 4, G.G1_SCHEMA)
    assert g1.num_columns == 8 and g1.num_rows == 4       # deserialize.rs:248-249
    assert g1.column("age").to_pylist() == [28] * 4
    g2 = pr.deserialize_array([bytes.fromhex(G.G2_HEX)], G.G2_SCHEMA)
    assert g2.to_pylist() == [G.G2_ROW]                    # deserialize.rs:307-308 (+ values)
    assert pr.last_walker() == walker


def _random_case(coracle, seed):
    rng = random.Random(1000 + seed)
    n = rng.choice([1, 31, 32, 33, 255, 256, 257, 600, 2500])
    sj, recs, data, off = gen_case(seed, n=n)
    k = rng.choice([1, 1, 2, 3, 8, 1000])
    assert_matches_oracle(coracle, pr.deserialize_array_threaded(recs, sj, k), sj, data, off, n, k)


@pytest.mark.parametrize("seed", JIT_SEEDS)
def test_random_schemas_jit(coracle, seed):
    pr.set_jit_enabled(1)
    try:
        _random_case(coracle, seed)
        assert pr.last_walker() == "jit"
    finally:
        pr.set_jit_enabled(-1)


@pytest.mark.parametrize("seed", INTERP_SEEDS)
def test_random_schemas_interp(coracle, seed):
    pr.set_jit_enabled(0)
    try:
        _random_case(coracle, seed)
        assert pr.last_walker() == "interp"
    finally:
        pr.set_jit_enabled(-1)


def test_packed_c_abi_path(coracle, walker):
    sj, recs, data, off = gen_case(7, n=5000)
    assert_matches_oracle(coracle, pr.decode_packed(data, off, len(recs), sj, 8), sj, data, off, len(recs), 8)
    # a window into a larger packed buffer: offsets[0] != 0 (absolute offsets are kept, the base is biased)
    sub_off = off[777:]
    got = pr.decode_packed(data, sub_off, len(recs) - 777, sj, 3)
    data2, off2 = po.pack_records(recs[777:])
    assert_matches_oracle(coracle, got, sj, data2, off2, len(recs) - 777, 3)
    # more chunks than the pipeline wants to run separately: one launch set handles all of them
    import workloads
    sjk, dk, ok = workloads.generate("kafka", 70_000, seed=4)
    assert_matches_oracle(coracle, pr.decode_packed(dk, ok, 70_000, sjk, 64), sjk, dk, ok, 70_000, 64, full_validate=False)


def test_empty_and_partition(coracle):
    data, off = po.pack_records([])
    b = pr.deserialize_array_threaded([], G.G345_SCHEMA, 8)
    assert len(b) == 1 and b[0].num_rows == 0
    assert_matches_oracle(coracle, b, G.G345_SCHEMA, data, off, 0, 8)
    for n, k, sizes in [(10, 4, [2, 2, 2, 4]), (10, 0, [10]), (3, 8, [1, 1, 1]), (7, 7, [1] * 7)]:
        sj, recs, data, off = gen_case(5, n=n)
        b = pr.deserialize_array_threaded(recs, sj, k)
        assert [x.num_rows for x in b] == sizes
        assert_matches_oracle(coracle, b, sj, data, off, n, k)


def test_malformed_inputs(coracle, walker):
    from tests import malformed as M
    good = [M.good_record(i) for i in range(700)]
    for name, code, bad in M.cases():
        for pos in (0, 300, 699):
            recs = list(good)
            recs[pos] = bad
            with pytest.raises(ValueError) as e:
                pr.deserialize_array_threaded(recs, M.FLAT, 3)
            assert f"(record {pos})" in str(e.value), (name, str(e.value))
            with pytest.raises(po.DecodeError) as oe:
                coracle.decode(M.FLAT, recs)
            assert oe.value.record == pos
    recs = [M.good_record(i) + b"junk" * (i % 3) for i in range(400)]  # trailing bytes ignored (:825-828)
    data, off = po.pack_records(recs)
    assert_matches_oracle(coracle, pr.deserialize_array_threaded(recs, M.FLAT, 2), M.FLAT, data, off, len(recs), 2)


def test_benchmark_workloads_small(coracle, walker):
    import workloads
    for name in ("flat", "kafka", "wide", "array_map", "nested", "nullable"):
        sj, data, off = workloads.generate(name, 30000, seed=7)
        assert_matches_oracle(coracle, pr.decode_packed(data, off, 30000, sj, 8), sj, data, off, 30000, 8, full_validate=False)
        assert pr.last_walker() == walker


def test_pipelined_chunks_large_batch(coracle, walker):
    """n >= 65536 with num_chunks > 1 takes the pipelined rv_decode_host path (H2D / decode / D2H of
    different chunks overlapped on worker streams); results and error reporting must not change."""
    import workloads
    sj, data, off = workloads.generate("kafka", 300_000, seed=11)
    for k in (2, 8, 13):
        assert_matches_oracle(coracle, pr.decode_packed(data, off, 300_000, sj, k), sj, data, off, 300_000, k, full_validate=False)
    from tests import malformed as M
    good = [M.good_record(i) for i in range(200_000)]
    for pos in (5, 70_123, 199_999):
        recs = list(good)
        recs[pos] = b"\x80"
        with pytest.raises(ValueError) as e:
            pr.deserialize_array_threaded(recs, M.FLAT, 8)
        assert f"(record {pos})" in str(e.value), str(e.value)
    recs = list(good)
    recs[150_000] = b"\x80"
    recs[20_000] = b"\xff" * 11
    with pytest.raises(ValueError) as e:
        pr.deserialize_array_threaded(recs, M.FLAT, 8)
    assert "(record 20000)" in str(e.value) and "varint" in str(e.value)   # first failing chunk wins


def test_huge_zero_width_lists_take_the_exact_tile_scan(coracle, walker):
    """Arrays of `null` items cost no bytes per item, so a 256-record tile can hold > 2^31 rows: the count kernel's
    tile scan switches to 64-bit arithmetic for such tiles (valid below the i32 ceiling, RV_ERR_OVERFLOW above)."""
    sj = '{"type":"record","name":"Z","fields":[{"name":"a","type":{"type":"array","items":"null"}},{"name":"b","type":"long"}]}'
    recs = [po.zigzag_bytes(1 << 23) + b"\x00" + po.zigzag_bytes(i) for i in range(64)]
    data, off = po.pack_records(recs)
    got = pr.deserialize_array_threaded(recs, sj, 1)
    assert_matches_oracle(coracle, got, sj, data, off, len(recs), 1)
    assert got[0].column("a").offsets[-1].as_py() == 64 << 23
    bad = [po.zigzag_bytes(1 << 30) + b"\x00" + po.zigzag_bytes(i) for i in range(3)]
    with pytest.raises(ValueError) as e:
        pr.deserialize_array(bad, sj)
    assert "overflow" in str(e.value)
    assert pr.deserialize_array_threaded(bad, sj, 3)[2].column("a").offsets[-1].as_py() == 1 << 30


def test_arrow_array_ingest_matches_list_ingest(coracle):
    """deserialize_arrow_array (Binary / LargeBinary / sliced / chunked) == deserialize_array_threaded on the list."""
    sj, recs, _, _ = gen_case(7, 300)
    want = pr.deserialize_array_threaded(recs, sj, 3)
    large = pa.array(recs, type=pa.large_binary())
    for arr in (pa.array(recs, type=pa.binary()), large,
                pa.chunked_array([pa.array(recs[:100], type=pa.binary()), pa.array(recs[100:], type=pa.binary())])):
        got = pr.deserialize_arrow_array(arr, sj, 3)
        assert len(got) == len(want) and all(g.equals(w) for g, w in zip(got, want))
    data, off = po.pack_records(recs[50:250])
    assert_matches_oracle(coracle, pr.deserialize_arrow_array(large.slice(50, 200), sj, 2), sj, data, off, 200, 2)
    assert pr.deserialize_arrow_array(pa.array([], type=pa.binary()), sj, 4)[0].num_rows == 0


def test_i32_offset_ceiling_and_huge_records(coracle):
    """A Utf8 column of one batch may not exceed 2^31-1 bytes (Arrow i32 offsets; arrow-rs panics, we return
    RV_ERR_OVERFLOW); split over two chunks the same input decodes.  36 KB records also force every tile through
    the non-staged (global-memory) interpreter pass."""
    sj = '{"type":"record","name":"Big","fields":[{"name":"id","type":"long"},{"name":"s","type":"string"}]}'
    n, slen = 60_000, 36_000
    body = np.frombuffer((b"0123456789abcdef" * (slen // 16 + 1))[:slen], dtype=np.uint8)
    head = po.zigzag_bytes(slen)
    rec_len = 3 + len(head) + slen          # ids below are 3-byte varints
    data = np.empty(n * rec_len, dtype=np.uint8)
    view = data.reshape(n, rec_len)
    ids = np.arange(n, dtype=np.int64) + 70_000
    zz = (ids << 1).astype(np.uint64)
    view[:, 0] = (zz & 0x7F) | 0x80
    view[:, 1] = ((zz >> 7) & 0x7F) | 0x80
    view[:, 2] = (zz >> 14) & 0x7F
    view[:, 3:3 + len(head)] = np.frombuffer(head, dtype=np.uint8)
    view[:, 3 + len(head):] = body
    off = np.arange(n + 1, dtype=np.int64) * rec_len
    assert n * slen > 2**31
    with pytest.raises(ValueError) as e:
        pr.decode_packed(data, off, n, sj, 1)
    assert "overflow" in str(e.value)
    got = pr.decode_packed(data, off, n, sj, 2)
    assert [b.num_rows for b in got] == [30_000, 30_000]
    for j, b in enumerate(got):
        assert b.column("id").to_pylist()[:3] == [70_000 + 30_000 * j + i for i in range(3)]
        s_col = b.column("s")
        assert s_col.buffers()[1].size >= 4 * 30_001 and s_col[29_999].as_py() == body.tobytes().decode()
        assert pa.compute.utf8_length(s_col).to_pylist()[:2] == [slen, slen]
    assert pr.lib.rv_last_slow_tiles() >= 0


def test_concurrent_calls_from_python_threads(coracle):
    """The reference releases the GIL and is re-entrant (src/lib.rs:82-87); so is this library."""
    import threading
    import workloads
    cases = []
    for i, name in enumerate(("kafka", "flat", "wide", "array_map")):
        sj, data, off = workloads.generate(name, 20_000 + 1000 * i, seed=20 + i)
        cases.append((sj, data, off, 20_000 + 1000 * i))
    errors = []

    def work(case, reps):
        sj, data, off, n = case
        try:
            for r in range(reps):
                assert_matches_oracle(coracle, pr.decode_packed(data, off, n, sj, 1 + r % 3), sj, data, off, n, 1 + r % 3, full_validate=False)
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(c, 4)) for c in cases for _ in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:2]


def test_error_surface():
    with pytest.raises(TypeError):
        pr.deserialize_array([b"ok", "not-bytes"], G.G1_SCHEMA)
    with pytest.raises(ValueError):
        pr.deserialize_array([b""], '{"type":"record","name":"B","fields":[{"name":"x","type":{"type":"long","logicalType":"local-timestamp-millis"}}]}')  # no CPU fallback
    with pytest.raises(ValueError):
        pr.deserialize_array([b""], "{not json")


def test_tiles_that_do_not_fit_shared_memory_take_the_global_walk(coracle, monkeypatch, walker):
    """The fused kernel walks a tile in shared memory when its bytes fit the input window and its strings the staging
    area; with either window clamped below the largest tile (RV_OUT_CLAMP / RV_IN_CLAMP) the rest must come out right
    through the in-kernel global-memory walk."""
    import workloads
    sj, data, off = workloads.generate("kafka", 60_000, seed=5)
    pr.decode_packed(data, off, 60_000, sj, 3)       # the schema handle now knows its sizes
    monkeypatch.setenv("RV_OUT_CLAMP", "1.0")
    got = pr.decode_packed(data, off, 60_000, sj, 3)
    assert pr.last_walker() == walker and pr.lib.rv_last_slow_tiles() > 0
    assert_matches_oracle(coracle, got, sj, data, off, 60_000, 3)
    monkeypatch.setenv("RV_IN_CLAMP", "1.0")     # and with the input window clamped too
    got = pr.decode_packed(data, off, 60_000, sj, 3)
    assert pr.lib.rv_last_slow_tiles() > 0
    assert_matches_oracle(coracle, got, sj, data, off, 60_000, 3)


def test_capacity_plan_measures_first_then_runs_one_pass_and_repeats_when_outgrown(coracle, walker):
    """Output buffers are sized from what earlier calls on the schema handle needed: the first call measures (2 passes),
    the same data again takes 1 pass, and data whose strings are far longer than planned repeats once with exact sizes."""
    sj = '{"type":"record","name":"C","fields":[{"name":"id","type":"long"},{"name":"s","type":["null","string"]},' \
         '{"name":"xs","type":{"type":"array","items":"int"}}]}'
    s = po.parse_schema(sj)
    rng = random.Random(11)

    def batch(n, slen, items):
        recs = [po.encode_datum(s, {"id": i, "s": (0, None) if rng.random() < 0.3 else (1, "y" * rng.randrange(slen)), "xs": list(range(rng.randrange(items)))})
                for i in range(n)]
        return recs, *po.pack_records(recs)

    schema = pr._get_or_parse_schema(sj)
    pr.lib.rv_schema_forget_stats(schema.handle)
    recs, data, off = batch(5000, 10, 3)
    assert_matches_oracle(coracle, pr.decode_packed(data, off, len(recs), sj, 3), sj, data, off, len(recs), 3)
    assert pr.lib.rv_last_passes() == 2
    assert_matches_oracle(coracle, pr.decode_packed(data, off, len(recs), sj, 3), sj, data, off, len(recs), 3)
    assert pr.lib.rv_last_passes() == 1
    recs, data, off = batch(5000, 400, 40)            # 40x the bytes, 13x the child rows per record
    assert_matches_oracle(coracle, pr.decode_packed(data, off, len(recs), sj, 3), sj, data, off, len(recs), 3)
    assert pr.lib.rv_last_passes() == 2
    assert_matches_oracle(coracle, pr.decode_packed(data, off, len(recs), sj, 3), sj, data, off, len(recs), 3)
    assert pr.lib.rv_last_passes() == 1
    recs, data, off = batch(700, 10, 3)               # smaller again: fits the larger plan
    assert_matches_oracle(coracle, pr.decode_packed(data, off, len(recs), sj, 1), sj, data, off, len(recs), 1)
    assert pr.lib.rv_last_passes() == 1


def test_capacity_plan_is_bounded_by_what_the_input_can_hold(coracle):
    """History from a batch of a few huge records must not size the buffers of a batch of many small ones (the plan is
    capped by the input's bytes), and the other way round the data simply outgrows the plan and the pass repeats."""
    sj = '{"type":"record","name":"H","fields":[{"name":"s","type":"string"},{"name":"xs","type":{"type":"array","items":"string"}}]}'
    s = po.parse_schema(sj)
    schema = pr._get_or_parse_schema(sj)
    pr.lib.rv_schema_forget_stats(schema.handle)
    big = [po.encode_datum(s, {"s": "x" * 3_000_000, "xs": ["y" * 1_000_000] * 3}) for _ in range(3)]
    data, off = po.pack_records(big)
    assert_matches_oracle(coracle, pr.decode_packed(data, off, 3, sj, 1), sj, data, off, 3, 1)
    small = [po.encode_datum(s, {"s": "ab", "xs": ["c"]}) for _ in range(400_000)]
    data, off = po.pack_records(small)
    got = pr.decode_packed(data, off, len(small), sj, 2)          # 400 k rows x 6 MB/row of history would be terabytes
    assert_matches_oracle(coracle, got, sj, data, off, len(small), 2, full_validate=False)
    data, off = po.pack_records(big)
    assert_matches_oracle(coracle, pr.decode_packed(data, off, 3, sj, 1), sj, data, off, 3, 1)


@pytest.mark.parametrize("name,n,k", [("kafka", 10_000_000, 8), ("flat", 2_000_000, 8), ("wide", 2_000_000, 8)])
def test_bench_scale_parity(coracle, name, n, k):
    """The configurations bench.py times (C3 at 10 M records / 8 chunks; C2 and C4 at 2 M), every exported buffer of
    every batch compared with the C oracle — through the host C ABI and through the device-resident path."""
    import ctypes
    import torch
    import workloads
    sj, data, off = workloads.generate(name, n, seed=42)
    got = pr.decode_packed(data, off, n, sj, k)
    assert_matches_oracle(coracle, got, sj, data, off, n, k, full_validate=False)
    del got
    # device-resident decode (what bench.py's `value` times), brought to the host afterwards
    schema = pr._get_or_parse_schema(sj)
    total = int(off[n])
    d_data = torch.empty(total + 64, dtype=torch.uint8, device="cuda")
    d_data[:total].copy_(torch.from_numpy(data))
    d_off = torch.from_numpy(off).cuda()
    h = ctypes.c_void_p()
    pr._check(pr.lib.rv_decode_device(schema.handle, d_data.data_ptr(), d_off.data_ptr(), n, k, torch.cuda.current_stream().cuda_stream, ctypes.byref(h)))
    assert pr.lib.rv_last_passes() == 1 and pr.lib.rv_last_slow_tiles() == 0
    pr._check(pr.lib.rv_result_to_host(h))
    assert_matches_oracle(coracle, pr._export_batches(h.value, schema), sj, data, off, n, k, full_validate=False)


def test_large_records_use_global_path(coracle, walker):
    """Records far larger than the shared-memory tile exercise the direct-from-global walk inside the fused kernel."""
    sj = '{"type":"record","name":"L","fields":[{"name":"s","type":"string"},{"name":"a","type":{"type":"array","items":"long"}}]}'
    s = po.parse_schema(sj)
    rng = random.Random(9)
    recs = [po.encode_datum(s, {"s": "x" * rng.choice([10, 5000, 70000]), "a": list(range(rng.choice([0, 3, 4000])))}) for _ in range(300)]
    data, off = po.pack_records(recs)
    assert_matches_oracle(coracle, pr.deserialize_array_threaded(recs, sj, 2), sj, data, off, len(recs), 2)


def test_zero_width_items_and_deep_nesting(coracle, walker):
    sj = '{"type":"record","name":"Z","fields":[{"name":"z","type":{"type":"array","items":"null"}},' \
         '{"name":"m","type":{"type":"map","values":{"type":"array","items":{"type":"array","items":["null","string"]}}}}]}'
    s = po.parse_schema(sj)
    rng = random.Random(3)
    recs = [po.encode_datum(s, po.random_value(s, rng)) for _ in range(3000)]
    data, off = po.pack_records(recs)
    assert_matches_oracle(coracle, pr.deserialize_array_threaded(recs, sj, 2), sj, data, off, len(recs), 2)
