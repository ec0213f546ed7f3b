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
    if name == "decodes_multi_variant_union":   # the null variant keeps its real type_id (fast_decode.rs:649-656)
        assert c_batch.column("u").type_codes.to_pylist() == [0, 1, 2, 3] * 2
        assert pa.types.is_null(c_batch.column("u").type.field(0).type)


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


# ---- ruhvro/src/schema_translate.rs:304-341 -------------------------------------------------------------------
def test_field_names_use_avro_field_name():
    import pyruhvro_b200 as pr
    sj = """{"type": "record", "name": "User", "namespace": "com.example", "fields": [
        {"name": "userid", "type": "string"},
        {"name": "address", "type": ["null", {"type": "record", "name": "Address", "fields": [
            {"name": "street", "type": "string"}, {"name": "city", "type": "string"}]}], "default": null},
        {"name": "class", "type": {"type": "enum", "name": "ClassEnum", "symbols": ["A", "B", "C"]}}]}"""
    for schema in (po.to_arrow_schema(po.parse_schema(sj)), pr.Schema(sj).arrow_schema):
        assert schema.names == ["userid", "address", "class"]
        address = schema.field("address").type
        assert pa.types.is_struct(address) and [address.field(i).name for i in range(address.num_fields)] == ["street", "city"]


# ---- ruhvro/src/serialize.rs:114-207, 263-310, 361-422 (Arrow built by hand -> serialize -> deserialize) ---------
CONVERT_SCHEMA = """{"type": "record", "name": "test", "fields": [
    {"name": "int_arr_1", "type": "int"},
    {"name": "int_arr_2", "type": ["null","int"]},
    {"name": "str_arr_1", "type": "string"},
    {"name": "str_arr_2", "type": ["null", "string"]},
    {"name": "list_arr", "type": "array", "items": ["null", "int"]},
    {"name": "list_arr2", "type": ["null", {"type": "array", "items": ["null", "int"]}]},
    {"name": "timestamp_arr", "type": {"type": "long", "logicalType": "timestamp-millis"}}]}"""
MAP_SCHEMA = """{"type": "record", "name": "test", "fields": [{"name": "map_field", "type": {"type": "map", "values": "int"}}]}"""
AB_SCHEMA = """{"type": "record", "name": "test", "fields": [{"name": "a", "type": "int"}, {"name": "b", "type": "string"}]}"""


def _convert_batch():
    """test_convert_to_avro's StructArray (serialize.rs:116-170).  Note the schema's field-level
    {"type": "array", "items": ...} shorthand (:185): apache-avro reads a field's type attributes from the field."""
    item = pa.field("item", pa.int32(), True)
    return pa.RecordBatch.from_arrays([
        pa.array([1, 2, 3, 4], pa.int32()),
        pa.array([1, None, None, 2], pa.int32()),
        pa.array(["one", "two", "three", "four"]),
        pa.array(["one", None, "three", None]),
        pa.array([[1, 2, 3], [2, 3], [], [None, 2, 3]], pa.list_(item)),
        pa.array([[1, 2, 3], None, [], [None, 2, 3]], pa.list_(item)),
        pa.array([1, 2, 3, 4], pa.timestamp("ms")),
    ], schema=pa.schema([
        pa.field("int_arr_1", pa.int32(), False), pa.field("int_arr_2", pa.int32(), True),
        pa.field("str_arr_1", pa.string(), False), pa.field("str_arr_2", pa.string(), True),
        pa.field("list_arr", pa.list_(item), False), pa.field("list_arr2", pa.list_(item), True),
        pa.field("timestamp_arr", pa.timestamp("ms"), False)]))


def _map_batch():
    """test_map_record_round_trip's MapArray::new_from_strings(["a","b"], [13, 2], offsets [0,1,2]) (serialize.rs:279-296)."""
    typ = pa.map_(pa.string(), pa.field("values", pa.int32(), False))
    return pa.RecordBatch.from_arrays([pa.array([[("a", 13)], [("b", 2)]], typ)],
                                      schema=pa.schema([pa.field("map_field", typ, False)]))


def _same_rows(a: pa.RecordBatch, b: pa.RecordBatch):
    assert a.num_rows == b.num_rows and a.schema.names == b.schema.names
    assert _rows(a) == _rows(b)


def test_convert_to_avro_and_map_round_trip_through_the_oracles(coracle):
    """serialize (fast_encode.rs restatement) -> deserialize (fast_decode.rs restatement) returns the batch."""
    for sj, batch in ((CONVERT_SCHEMA, _convert_batch()), (MAP_SCHEMA, _map_batch())):
        s = po.parse_schema(sj)
        assert po.is_supported(s)
        chunks = po.py_encode(s, batch, 1)
        assert len(chunks) == 1 and len(chunks[0]) == batch.num_rows
        back = po.canon_to_batch(coracle.decode(sj, chunks[0]), po.to_arrow_schema(s))
        _same_rows(back, batch)
        assert po.canon_diff(coracle.decode(sj, chunks[0]), po.py_decode(s, chunks[0])) is None


def test_serialize_matches_columns_by_name_and_errors_on_missing_column_oracle():
    s = po.parse_schema(AB_SCHEMA)
    a, b = pa.array([1, 2, 3], pa.int32()), pa.array(["x", "y", "z"])
    in_order = pa.RecordBatch.from_arrays([a, b], names=["a", "b"])
    reversed_ = pa.RecordBatch.from_arrays([b, a], names=["b", "a"])
    assert po.py_encode(s, in_order, 1) == po.py_encode(s, reversed_, 1)
    with pytest.raises(po.EncodeError, match="missing column 'b'"):
        po.py_encode(s, pa.RecordBatch.from_arrays([a], names=["a"]), 1)


@pytest.mark.gpu
def test_gpu_convert_to_avro_and_map_round_trip(coracle):
    import pyruhvro_b200 as pr
    for sj, batch in ((CONVERT_SCHEMA, _convert_batch()), (MAP_SCHEMA, _map_batch())):
        chunks = pr.serialize_record_batch(batch, sj, 1)
        assert len(chunks) == 1 and len(chunks[0]) == batch.num_rows
        datums = [bytes(x.as_py()) for x in chunks[0]]
        assert [datums] == po.py_encode(po.parse_schema(sj), batch, 1)      # same bytes as the fast_encode.rs restatement
        _same_rows(pr.deserialize_array(datums, sj), batch)                 # and the reference's assertion


@pytest.mark.gpu
def test_gpu_serialize_matches_columns_by_name_and_errors_on_missing_column():
    import pyruhvro_b200 as pr
    a, b = pa.array([1, 2, 3], pa.int32()), pa.array(["x", "y", "z"])
    in_order = pr.serialize_record_batch(pa.RecordBatch.from_arrays([a, b], names=["a", "b"]), AB_SCHEMA, 1)
    reversed_ = pr.serialize_record_batch(pa.RecordBatch.from_arrays([b, a], names=["b", "a"]), AB_SCHEMA, 1)
    assert len(in_order) == len(reversed_) == 1 and in_order[0].equals(reversed_[0])
    with pytest.raises(ValueError, match="missing column 'b'"):
        pr.serialize_record_batch(pa.RecordBatch.from_arrays([a], names=["a"]), AB_SCHEMA, 1)


# ---- SURVEY.md A.3 edge cases that need no GPU -------------------------------------------------------------------
def test_records_without_fields_fail_like_the_reference(coracle):
    """Nested: "RecordDecoder produced a record with 0 fields" (fast_decode.rs:633-635); top level:
    RecordBatch::try_new with no columns (:834).  The reference's gate (:38-61) still says "supported" — the failure is
    at finish; rv_schema_is_supported answers "can this library decode it" and may say 0 (nothing can decode it)."""
    import numpy as np
    import pyruhvro_b200 as pr
    top = '{"type":"record","name":"E","fields":[]}'
    nested = ('{"type":"record","name":"O","fields":[{"name":"a","type":"int"},'
              '{"name":"e","type":["null",{"type":"record","name":"E","fields":[]}]}]}')
    for sj, text in ((top, "at least one column"), (nested, "record with 0 fields")):
        s = po.parse_schema(sj)
        assert po.is_supported(s)
        with pytest.raises(po.DecodeError):
            po.py_decode(s, [b"\x02\x00"])
        with pytest.raises(po.DecodeError):
            coracle.decode(sj, [b"\x02\x00"])
        with pytest.raises(ValueError, match=text):   # raised from the plan, before any CUDA call
            pr.decode_packed(np.frombuffer(b"\x02\x00", dtype=np.uint8), np.array([0, 2]), 1, sj, 1)


def test_two_variant_union_without_null_is_a_sparse_union(coracle):
    """["string","int"]: no null branch, so it is NOT folded into a nullable column — a non-nullable sparse union of two
    (fast_decode.rs:376-384, schema_translate.rs:78-104)."""
    import pyruhvro_b200 as pr
    sj = '{"type":"record","name":"U","fields":[{"name":"u","type":["string","int"]}]}'
    s = po.parse_schema(sj)
    for schema in (po.to_arrow_schema(s), pr.Schema(sj).arrow_schema):
        f = schema.field("u")
        assert pa.types.is_union(f.type) and f.type.mode == "sparse" and f.type.num_fields == 2 and not f.nullable
        assert [f.type.field(i).name for i in range(2)] == ["varchar", "int"]
    recs = [po.encode_datum(s, {"u": (0, "abc")}), po.encode_datum(s, {"u": (1, -7)}), po.encode_datum(s, {"u": (0, "")})]
    c = coracle.decode(sj, recs)
    assert po.canon_diff(c, po.py_decode(s, recs)) is None
    batch = po.canon_to_batch(c, po.to_arrow_schema(s))
    assert batch.column("u").to_pylist() == ["abc", -7, ""]
    assert batch.column("u").type_codes.to_pylist() == [0, 1, 0]
    with pytest.raises(po.DecodeError):                    # branch index 2 is out of range (:643-658)
        coracle.decode(sj, [b"\x04"])


def test_chunk_partition_rule():
    """deserialize.rs:53-68: k = clamp(num_chunks, 1, max(n, 1)); equal floors, the last chunk takes the remainder."""
    import pyruhvro_b200 as pr
    assert [b - a for a, b in po.chunk_bounds(10, po.clamp_chunks(4, 10))] == [2, 2, 2, 4]
    assert po.clamp_chunks(0, 10) == 1 and po.clamp_chunks(99, 3) == 3 and po.clamp_chunks(5, 0) == 1
    shards = pr.distributed.shard_bounds if hasattr(pr, "distributed") else None
    assert shards is None or callable(shards)
