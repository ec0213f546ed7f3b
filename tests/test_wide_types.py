"""The wider subset: bytes, fixed, uuid, decimal, time-millis / time-micros and references to
named types — schemas the reference sends to its Value-tree fallback (fast_decode.rs:16-17,59), which cannot build
these columns (complex.rs:431 `unimplemented!`).  Arrow types follow schema_translate.rs:58,133-143; values follow the
Avro specification.  CPU tests drive the product's schema front-end, plan, walkers and export through the host
emulation (tests/emu); `-m gpu` tests drive the CUDA path through the Python surface."""
import datetime
import decimal
import json
import random
import uuid

import pyarrow as pa
import pytest

import pyruhvro_b200 as pr
from oracle import pyoracle as po
from tests import emu
from tests.parity import assert_matches_pyoracle_wide, expected_schema_wide, gen_case_wide

ALL_WIDE = json.dumps({"type": "record", "name": "W", "namespace": "ns", "fields": [
    {"name": "b", "type": "bytes"},
    {"name": "fx", "type": {"type": "fixed", "name": "Md5", "size": 16, "doc": "digest", "aliases": ["Hash"]}},
    {"name": "fx2", "type": "Md5"},                                   # a reference by name
    {"name": "u", "type": ["null", {"type": "string", "logicalType": "uuid"}]},
    {"name": "d", "type": {"type": "bytes", "logicalType": "decimal", "precision": 20, "scale": 4}},
    {"name": "df", "type": ["null", {"type": "fixed", "name": "Money", "size": 9, "logicalType": "decimal", "precision": 18, "scale": 2}]},
    {"name": "tm", "type": {"type": "int", "logicalType": "time-millis"}},
    {"name": "tu", "type": ["null", {"type": "long", "logicalType": "time-micros"}]},
    {"name": "xs", "type": {"type": "array", "items": ["null", "ns.Money"]}},
    {"name": "m", "type": {"type": "map", "values": "bytes"}},
    {"name": "un", "type": ["bytes", "int", {"type": "fixed", "name": "Four", "size": 4}]},
]})


def _dec_bytes(v: decimal.Decimal, scale: int, n=None) -> bytes:
    unscaled = int(v.scaleb(scale))
    n = n or max(1, (unscaled.bit_length() + 8) // 8)
    return unscaled.to_bytes(n, "big", signed=True)


def _all_wide_rows():
    s = po.parse_schema(ALL_WIDE, wide=True)
    rows, recs = [], []
    rng = random.Random(4)
    for i in range(300):
        u = uuid.UUID(int=rng.getrandbits(128))
        d = decimal.Decimal(rng.randint(-10**15, 10**15)).scaleb(-4)
        df = decimal.Decimal(rng.randint(-10**12, 10**12)).scaleb(-2)
        xs = [None if rng.random() < 0.3 else decimal.Decimal(rng.randint(-10**6, 10**6)).scaleb(-2) for _ in range(rng.randrange(4))]
        fx = bytes(rng.randrange(256) for _ in range(16))
        row = {"b": bytes(rng.randrange(256) for _ in range(rng.randrange(40))), "fx": fx, "fx2": fx[::-1],
               "u": None if i % 5 == 0 else u.bytes, "d": d, "df": None if i % 7 == 0 else df,
               "tm": datetime.time(rng.randrange(24), rng.randrange(60), rng.randrange(60), 1000 * rng.randrange(1000)),
               "tu": None if i % 3 == 0 else datetime.time(rng.randrange(24), rng.randrange(60), rng.randrange(60), rng.randrange(10**6)),
               "xs": xs, "m": [("k%d" % j, bytes([j] * j)) for j in range(rng.randrange(3))],
               "un": [b"raw", 7, b"\x01\x02\x03\x04"][i % 3]}
        t = row["tm"]
        tu = row["tu"]
        val = {"b": row["b"], "fx": fx, "fx2": fx[::-1],
               "u": (0, None) if row["u"] is None else (1, str(u) if i % 2 else u.hex),
               "d": _dec_bytes(d, 4), "df": (0, None) if row["df"] is None else (1, _dec_bytes(df, 2, 9)),
               "tm": ((t.hour * 60 + t.minute) * 60 + t.second) * 1000 + t.microsecond // 1000,
               "tu": (0, None) if tu is None else (1, ((tu.hour * 60 + tu.minute) * 60 + tu.second) * 10**6 + tu.microsecond),
               "xs": [(0, None) if x is None else (1, _dec_bytes(x, 2, 9)) for x in xs],
               "m": row["m"], "un": (i % 3, row["un"])}
        rows.append(row)
        recs.append(po.encode_datum(s, val))
    return rows, recs


def _check_values(batch: pa.RecordBatch, rows):
    """pyarrow's own reading of the buffers against independently computed Python values."""
    got = batch.to_pylist()
    assert len(got) == len(rows)
    for g, w in zip(got, rows):
        for key in ("b", "fx", "fx2", "u", "d", "df", "tm", "tu", "xs", "un"):
            assert g[key] == w[key], (key, g[key], w[key])
        assert [tuple(kv) for kv in g["m"]] == w["m"]


def test_schema_translation_and_gate():
    s = pr.Schema(ALL_WIDE)
    assert s.is_supported
    assert s.arrow_schema.equals(expected_schema_wide(ALL_WIDE), check_metadata=True)
    t = s.arrow_schema
    assert t.field("b").type == pa.binary() and t.field("fx").type == pa.binary(16) and t.field("fx2").type == pa.binary(16)
    assert t.field("u").type == pa.binary(16) and t.field("u").nullable
    assert t.field("d").type == pa.decimal128(20, 4) and t.field("df").type == pa.decimal128(18, 2)
    assert t.field("tm").type == pa.time32("ms") and t.field("tu").type == pa.time64("us")
    assert t.field("fx").metadata == {b"avro::doc": b"digest", b"avro::aliases": b"[ns.Hash]"}
    assert [f.name for f in t.field("un").type] == ["varbinary", "int", "fixedsizebinary"]
    for seed in range(120):
        sj = po.random_schema_json(random.Random(seed), wide=True)
        assert pr.Schema(sj).arrow_schema.equals(expected_schema_wide(sj), check_metadata=True), sj
    # still outside: recursion, 256-bit decimals, duration, local timestamps, and a fixed of size 0 (no wire bytes per value:
    # a list header could announce 2^31 of them)
    for bad in ['{"type":"record","name":"R","fields":[{"name":"next","type":["null","R"]}]}',
                '{"type":"record","name":"R","fields":[{"name":"a","type":{"type":"array","items":{"type":"fixed","name":"F","size":0}}}]}',
                '{"type":"record","name":"R","fields":[{"name":"d","type":{"type":"fixed","name":"D","size":0,"logicalType":"decimal","precision":1}}]}',
                '{"type":"record","name":"R","fields":[{"name":"d","type":{"type":"bytes","logicalType":"decimal","precision":39,"scale":0}}]}',
                '{"type":"record","name":"R","fields":[{"name":"d","type":{"type":"fixed","name":"D","size":12,"logicalType":"duration"}}]}',
                '{"type":"record","name":"R","fields":[{"name":"t","type":{"type":"long","logicalType":"local-timestamp-millis"}}]}']:
        assert not pr.Schema(bad).is_supported, bad
        assert not po.is_supported(po.parse_schema(bad, wide=True)), bad


@pytest.mark.parametrize("walker", ["interp", "gen"])
def test_emulated_walkers_all_wide_types(walker):
    rows, recs = _all_wide_rows()
    data, off = po.pack_records(recs)
    for k in (1, 3):
        got = emu.decode(ALL_WIDE, data, off, len(recs), k, walker=walker)
        assert_matches_pyoracle_wide(got, ALL_WIDE, recs, k)
    _check_values(emu.decode(ALL_WIDE, data, off, len(recs), 1, walker=walker)[0], rows)


@pytest.mark.parametrize("seed", range(40))
def test_emulated_random_wide_schemas(seed):
    sj, recs, data, off = gen_case_wide(seed)
    k = random.Random(seed).choice([1, 2, 5])
    assert_matches_pyoracle_wide(emu.decode(sj, data, off, len(recs), k, walker="interp"), sj, recs, k)
    if seed < 12:
        assert_matches_pyoracle_wide(emu.decode(sj, data, off, len(recs), k, walker="gen"), sj, recs, k)


def _one(sj, datum):
    data, off = po.pack_records([datum])
    return emu.decode(sj, data, off, 1, 1)


def test_value_errors_and_null_slots():
    uu = '{"type":"record","name":"R","fields":[{"name":"u","type":{"type":"string","logicalType":"uuid"}}]}'
    ok = str(uuid.UUID(int=5)).encode()
    assert _one(uu, po.zigzag_bytes(len(ok)) + ok)[0].column(0)[0].as_py() == uuid.UUID(int=5).bytes
    for bad in (b"not-a-uuid", ok[:-1] + b"g", ok.replace(b"-", b"+"), ok + b"0"):
        with pytest.raises(emu.EmuError) as e:
            _one(uu, po.zigzag_bytes(len(bad)) + bad)
        assert e.value.code == 11 and e.value.record == 0
    dd = '{"type":"record","name":"R","fields":[{"name":"d","type":{"type":"bytes","logicalType":"decimal","precision":38,"scale":0}}]}'
    for raw, want in ((b"", 0), (b"\x7f", 127), (b"\xff", -1), (b"\x80" + bytes(15), -(1 << 127)), (b"\x00\x01\x00", 256)):
        assert _one(dd, po.zigzag_bytes(len(raw)) + raw)[0].column(0)[0].as_py() == decimal.Decimal(want)
    with pytest.raises(emu.EmuError) as e:
        _one(dd, po.zigzag_bytes(17) + bytes(17))
    assert e.value.code == 11
    with pytest.raises(emu.EmuError) as e:      # fixed(4) cut short: end of buffer
        _one('{"type":"record","name":"R","fields":[{"name":"f","type":{"type":"fixed","name":"F","size":4}}]}', b"\x01\x02")
    assert e.value.code == 1
    # null slots of the wide columns hold zero bytes
    sj = '{"type":"record","name":"R","fields":[{"name":"u","type":["null",{"type":"string","logicalType":"uuid"}]},' \
         '{"name":"f","type":["null",{"type":"fixed","name":"F","size":3}]}]}'
    b = _one(sj, b"\x00\x00")[0]
    assert b.column(0).buffers()[1].to_pybytes()[:16] == bytes(16) and b.column(1).buffers()[1].to_pybytes()[:3] == bytes(3)
    assert b.column(0).null_count == 1 and b.column(1).null_count == 1


def test_encode_rejects_the_wider_subset():
    """serialize_record_batch keeps the reference's subset (fast_encode.rs has no arm for these types either)."""
    s = pr.Schema(ALL_WIDE)
    assert s.is_supported  # the decode gate; the encode gate is checked on a GPU box in test_gpu_encode.py


# ---- GPU ------------------------------------------------------------------------------------------------------------
@pytest.fixture(params=["jit", "interp"])
def walker_gpu(request):
    pr.set_jit_enabled(1 if request.param == "jit" else 0)
    yield request.param
    pr.set_jit_enabled(-1)


@pytest.mark.gpu
def test_gpu_all_wide_types(walker_gpu):
    rows, recs = _all_wide_rows()
    for k in (1, 3):
        assert_matches_pyoracle_wide(pr.deserialize_array_threaded(recs, ALL_WIDE, k), ALL_WIDE, recs, k)
    _check_values(pr.deserialize_array(recs, ALL_WIDE), rows)
    assert pr.last_walker() == walker_gpu


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(30))
def test_gpu_random_wide_schemas(seed):
    rng = random.Random(2000 + seed)
    sj, recs, data, off = gen_case_wide(seed, n=rng.choice([1, 33, 256, 257, 900, 2500]))
    k = rng.choice([1, 2, 8])
    pr.set_jit_enabled(1 if seed % 3 else 0)
    try:
        assert_matches_pyoracle_wide(pr.deserialize_array_threaded(recs, sj, k), sj, recs, k)
    finally:
        pr.set_jit_enabled(-1)


@pytest.mark.gpu
def test_gpu_value_errors():
    uu = '{"type":"record","name":"R","fields":[{"name":"id","type":"long"},{"name":"u","type":{"type":"string","logicalType":"uuid"}}]}'
    s = po.parse_schema(uu, wide=True)
    recs = [po.encode_datum(s, {"id": i, "u": str(uuid.UUID(int=i))}) for i in range(700)]
    recs[413] = po.encode_datum(s, {"id": 413, "u": "zz" + str(uuid.UUID(int=1))[2:]})
    with pytest.raises(ValueError, match=r"logical type.*record 413"):
        pr.deserialize_array(recs, uu)
    with pytest.raises(ValueError):
        b = pr.deserialize_array(recs[:100], uu)
        pr.serialize_record_batch(b, uu, 1)   # the encode direction keeps the reference's subset
