"""The reference's own unit tests, restated one by one (same names, same schemas, same rows).

ruhvro/src/fast_decode.rs:955-1010 (`supports_*`: the gate), :1012-1226 (`decodes_*`: the fast path equals the
Value-tree baseline on rows built by a closure) and ruhvro/src/fast_encode.rs:638-830 (`encodes_*`: avro -> fast decode
-> fast encode -> baseline decode gives the batch back).  The reference builds its datums with `to_avro_datum` at run
time; here the same rows are encoded by the oracle's wire writer and the expected values are the rows themselves
(which is what "equals the baseline" means).

CPU (`-m "not gpu"`): both oracles decode every case to the expected rows — this pins the oracle on the reference's
test corpus beyond the five literal datums of tests/golden.  GPU (`-m gpu`): the CUDA path through the reference-facing
Python API gives the same rows, matches the oracle buffer for buffer, and `serialize_record_batch` returns the datums.
"""
import datetime

import pyarrow as pa
import pytest

from oracle import pyoracle as po

P_FLAT = """{"type":"record","name":"P","fields":[
    {"name":"i","type":"int"},{"name":"l","type":"long"},{"name":"f","type":"float"},
    {"name":"d","type":"double"},{"name":"b","type":"boolean"},{"name":"s","type":"string"}]}"""
P_NULLABLE = """{"type":"record","name":"P","fields":[
    {"name":"i","type":["null","int"],"default":null},
    {"name":"s","type":["string","null"],"default":""}]}"""
L_LOGICAL = """{"type":"record","name":"L","fields":[
    {"name":"d","type":{"type":"int","logicalType":"date"}},
    {"name":"tm","type":{"type":"long","logicalType":"timestamp-millis"}},
    {"name":"tu","type":{"type":"long","logicalType":"timestamp-micros"}}]}"""
R_ENUM = """{"type":"record","name":"R","fields":[
    {"name":"e","type":{"type":"enum","name":"E","symbols":["A","B","C"]}}]}"""
O_NESTED = """{"type":"record","name":"O","fields":[
    {"name":"outer_id","type":"long"},
    {"name":"inner","type":{"type":"record","name":"I","fields":[
        {"name":"x","type":"int"},{"name":"label","type":"string"}]}}]}"""
O_NULLABLE_NESTED = """{"type":"record","name":"O","fields":[
    {"name":"inner","type":["null",{"type":"record","name":"I","fields":[{"name":"x","type":"int"}]}],"default":null}]}"""
M_UNION = """{"type":"record","name":"M","fields":[{"name":"u","type":["null","string","int","boolean"]}]}"""
C_ARRAY_STR = """{"type":"record","name":"C","fields":[{"name":"tags","type":{"type":"array","items":"string"}}]}"""
C_ARRAY_INT = """{"type":"record","name":"C","fields":[{"name":"tags","type":{"type":"array","items":"int"}}]}"""
C_MAP_STR = """{"type":"record","name":"C","fields":[{"name":"props","type":{"type":"map","values":"string"}}]}"""

F32 = lambda x: float(pa.scalar(x, pa.float32()).as_py())  # noqa: E731
EPOCH = datetime.date(1970, 1, 1)


def _union_row(i):
    return [(0, None), (1, f"s-{i}"), (2, i * 11), (3, i % 8 == 3)][i % 4]


# name -> (schema, n, wire value of row i, expected Arrow row i)
CASES = {
    "decodes_flat_primitives": (P_FLAT, 5,
        lambda i: {"i": i, "l": i * 100, "f": i * 1.5, "d": i * 2.25, "b": i % 2 == 0, "s": f"row-{i}"},
        lambda i: {"i": i, "l": i * 100, "f": F32(i * 1.5), "d": i * 2.25, "b": i % 2 == 0, "s": f"row-{i}"}),
    "decodes_nullable_primitives": (P_NULLABLE, 6,
        lambda i: {"i": (1, i) if i % 2 == 0 else (0, None), "s": (1, None) if i % 3 == 0 else (0, f"v-{i}")},
        lambda i: {"i": i if i % 2 == 0 else None, "s": None if i % 3 == 0 else f"v-{i}"}),
    "decodes_logical_types": (L_LOGICAL, 4,
        lambda i: {"d": i * 7, "tm": 1_700_000_000_000 + i, "tu": 1_700_000_000_000_000 + i},
        lambda i: {"d": EPOCH + datetime.timedelta(days=7 * i),
                   "tm": datetime.datetime(1970, 1, 1) + datetime.timedelta(milliseconds=1_700_000_000_000 + i),
                   "tu": datetime.datetime(1970, 1, 1) + datetime.timedelta(microseconds=1_700_000_000_000_000 + i)}),
    "decodes_enum": (R_ENUM, 6, lambda i: {"e": i % 3}, lambda i: {"e": "ABC"[i % 3]}),
    "decodes_nested_record": (O_NESTED, 5,
        lambda i: {"outer_id": i, "inner": {"x": i, "label": f"lbl-{i}"}},
        lambda i: {"outer_id": i, "inner": {"x": i, "label": f"lbl-{i}"}}),
    "decodes_nullable_nested_record": (O_NULLABLE_NESTED, 6,
        lambda i: {"inner": (1, {"x": i}) if i % 2 == 0 else (0, None)},
        lambda i: {"inner": {"x": i} if i % 2 == 0 else None}),
    "decodes_multi_variant_union": (M_UNION, 8, lambda i: {"u": _union_row(i)}, lambda i: {"u": _union_row(i)[1]}),
    "decodes_array_of_string": (C_ARRAY_STR, 6,
        lambda i: {"tags": [f"t-{i}-a", f"t-{i}-b"]}, lambda i: {"tags": [f"t-{i}-a", f"t-{i}-b"]}),
    "decodes_empty_array": (C_ARRAY_INT, 3, lambda i: {"tags": []}, lambda i: {"tags": []}),
    "decodes_map_of_string": (C_MAP_STR, 4,
        lambda i: {"props": [(f"k{i}-1", f"v{i}-1"), (f"k{i}-2", f"v{i}-2")]},
        lambda i: {"props": [(f"k{i}-1", f"v{i}-1"), (f"k{i}-2", f"v{i}-2")]}),
}

SUPPORTS = {
    "supports_flat_primitives": """{"type":"record","name":"P","fields":[{"name":"i","type":"int"},{"name":"s","type":"string"}]}""",
    "supports_logical_types_and_enum": """{"type":"record","name":"L","fields":[
        {"name":"d","type":{"type":"int","logicalType":"date"}},
        {"name":"tm","type":{"type":"long","logicalType":"timestamp-millis"}},
        {"name":"tu","type":{"type":"long","logicalType":"timestamp-micros"}},
        {"name":"e","type":{"type":"enum","name":"E","symbols":["A","B","C"]}}]}""",
    "supports_nested_record": """{"type":"record","name":"O","fields":[
        {"name":"inner","type":{"type":"record","name":"I","fields":[{"name":"x","type":"int"}]}}]}""",
    "supports_multi_variant_union": M_UNION,
}


def _datums(name):
    sj, n, wire, want = CASES[name]
    s = po.parse_schema(sj)
    return sj, s, [po.encode_datum(s, wire(i)) for i in range(n)], [want(i) for i in range(n)]


def _rows(batch):
    rows = batch.to_pylist()
    for r in rows:  # pyarrow renders map rows as lists of (key, value) tuples; normalise lists of lists
        for k, v in r.items():
            if isinstance(v, list) and v and isinstance(v[0], (list, tuple)):
                r[k] = [tuple(x) for x in v]
    return rows


@pytest.mark.parametrize("name", sorted(SUPPORTS))
def test_gate(name):
    """fast_decode.rs:955-1010: every `supports_*` schema is inside the direct-decode subset."""
    import pyruhvro_b200 as pr
    assert po.is_supported(po.parse_schema(SUPPORTS[name]))
    assert pr.Schema(SUPPORTS[name]).is_supported


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracles_reproduce_the_reference_case(name, coracle):
    sj, s, recs, want = _datums(name)
    assert po.is_supported(s)
    arrow_schema = po.to_arrow_schema(s)
    c_batch = po.canon_to_batch(coracle.decode(sj, recs), arrow_schema)
    py_batch = po.canon_to_batch(po.py_decode(s, recs), arrow_schema)
    c_batch.validate(full=True)
    assert _rows(c_batch) == want
    assert _rows(py_batch) == want
    assert po.canon_diff(coracle.decode(sj, recs), po.py_decode(s, recs)) is None


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_decodes_the_reference_case(name, coracle):
    import pyruhvro_b200 as pr
    from tests.parity import assert_matches_oracle
    sj, s, recs, want = _datums(name)
    batch = pr.deserialize_array(recs, sj)
    batch.validate(full=True)
    assert _rows(batch) == want
    data, off = po.pack_records(recs)
    for k in (1, 2):
        assert_matches_oracle(coracle, pr.deserialize_array_threaded(recs, sj, k), sj, data, off, len(recs), k)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_encodes_the_reference_case(name):
    """fast_encode.rs:638-830 (`encodes_*`, same schemas and rows): decode -> serialize_record_batch(…, 1) returns one
    chunk whose datums decode back to the batch — here they are byte-identical to the input datums."""
    import pyruhvro_b200 as pr
    sj, s, recs, _ = _datums(name)
    batch = pr.deserialize_array(recs, sj)
    chunks = pr.serialize_record_batch(batch, sj, 1)
    assert len(chunks) == 1 and chunks[0].type == pa.binary()
    out = [bytes(x.as_py()) for x in chunks[0]]
    assert out == recs
    assert pr.deserialize_array(out, sj).equals(batch)
