"""Shared parity helpers: compare decoded batches (product or emulation) with the C oracle,
buffer-for-buffer (SURVEY.md A.2 level L2) plus schema and pyarrow validation."""
import json

import numpy as np
import pyarrow as pa

from oracle import pyoracle as po


def normalize_map_names(t: pa.DataType) -> pa.DataType:
    """pyarrow's C-Data import renames a map's key/value fields to "key"/"value"; apply the same
    renaming to the expected type (the C-level names are checked separately in test_schema.py)."""
    if pa.types.is_map(t):
        return pa.map_(pa.field("key", normalize_map_names(t.key_type), False),
                       pa.field("value", normalize_map_names(t.item_type), t.item_field.nullable))
    if pa.types.is_list(t):
        f = t.value_field
        return pa.list_(pa.field(f.name, normalize_map_names(f.type), f.nullable, f.metadata))
    if pa.types.is_struct(t):
        return pa.struct([pa.field(f.name, normalize_map_names(f.type), f.nullable, f.metadata) for f in t])
    if pa.types.is_union(t):
        return pa.union([pa.field(f.name, normalize_map_names(f.type), f.nullable, f.metadata) for f in t],
                        mode="sparse", type_codes=t.type_codes)
    return t


def expected_schema(schema_json: str) -> pa.Schema:
    s = po.to_arrow_schema(po.parse_schema(schema_json))
    return pa.schema([pa.field(f.name, normalize_map_names(f.type), f.nullable, f.metadata) for f in s])


def assert_matches_oracle(coracle, batches, schema_json, data, offsets, n, num_chunks, full_validate=True):
    """`batches`: list[pa.RecordBatch] from the implementation under test."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    want = coracle.decode_threaded_packed(schema_json, data, offsets, n, num_chunks, threads=4)
    assert len(batches) == len(want), (len(batches), len(want))
    exp_schema = expected_schema(schema_json)
    bounds = po.chunk_bounds(n, po.clamp_chunks(num_chunks, n))
    for i, (b, w) in enumerate(zip(batches, want)):
        assert b.schema.equals(exp_schema, check_metadata=True), f"batch {i} schema\n{b.schema}\n!=\n{exp_schema}"
        assert b.num_rows == bounds[i][1] - bounds[i][0]
        if full_validate:
            b.validate(full=True)
        d = po.canon_diff(po.canon_from_batch(b), w, f"batch[{i}]")
        assert d is None, d


def gen_case(seed: int, n=None, max_depth: int = 3):
    import random
    rng = random.Random(seed)
    sj = po.random_schema_json(rng, max_depth=max_depth)
    s = po.parse_schema(sj)
    if n is None:
        n = rng.choice([1, 2, 31, 32, 33, 255, 256, 257, 600])
    neg = rng.random() < 0.3
    recs = [po.encode_datum(s, po.random_value(s, rng), neg_blocks=neg) for _ in range(n)]
    data, off = po.pack_records(recs)
    return sj, recs, data, off


# ---- the wider subset (bytes / fixed / uuid / decimal / time-* / named references) -------------------------------
# The reference has no implementation of these, so the checker is the pure-Python restatement of the Avro
# specification + schema_translate.rs types in oracle/pyoracle.py (wide=True), itself cross-checked against pyarrow's
# own reading of the buffers (RecordBatch.validate(full) + to_pylist values in tests/test_wide_types.py).
def expected_schema_wide(schema_json: str) -> pa.Schema:
    s = po.to_arrow_schema(po.parse_schema(schema_json, wide=True))
    return pa.schema([pa.field(f.name, normalize_map_names(f.type), f.nullable, f.metadata) for f in s])


def assert_matches_pyoracle_wide(batches, schema_json, recs, num_chunks, full_validate=True):
    s = po.parse_schema(schema_json, wide=True)
    n = len(recs)
    bounds = po.chunk_bounds(n, po.clamp_chunks(num_chunks, n))
    assert len(batches) == len(bounds), (len(batches), len(bounds))
    exp_schema = expected_schema_wide(schema_json)
    for i, (b, (r0, r1)) in enumerate(zip(batches, bounds)):
        assert b.schema.equals(exp_schema, check_metadata=True), f"batch {i} schema\n{b.schema}\n!=\n{exp_schema}"
        assert b.num_rows == r1 - r0
        if full_validate:
            b.validate(full=True)
        d = po.canon_diff(po.canon_from_batch(b), po.py_decode(s, recs[r0:r1]), f"batch[{i}]")
        assert d is None, d


def gen_case_wide(seed: int, n=None, max_depth: int = 3):
    import random
    rng = random.Random(seed)
    sj = po.random_schema_json(rng, max_depth=max_depth, wide=True)
    s = po.parse_schema(sj, wide=True)
    if n is None:
        n = rng.choice([1, 2, 31, 32, 33, 255, 256, 257, 600])
    neg = rng.random() < 0.3
    recs = [po.encode_datum(s, po.random_value(s, rng), neg_blocks=neg) for _ in range(n)]
    data, off = po.pack_records(recs)
    return sj, recs, data, off
