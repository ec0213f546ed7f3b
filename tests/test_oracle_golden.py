"""Pins both oracles (C restatement + independent pure-Python restatement) against the
literal golden datums the reference's tests hold, and against each other on random schemas."""
import datetime
import random

import pyarrow as pa
import pytest

from oracle import pyoracle as po
from tests.golden import reference_datums as G


def _decoders(coracle):
    return [("py", lambda sj, recs: po.py_decode(po.parse_schema(sj), recs)),
            ("c", lambda sj, recs: coracle.decode(sj, recs))]


def _batch(sj, cols):
    schema = po.to_arrow_schema(po.parse_schema(sj))
    b = po.canon_to_batch(cols, schema)
    b.validate(full=True)
    return b


@pytest.mark.parametrize("which", ["py", "c"])
def test_g1_values(coracle, which):
    dec = dict(_decoders(coracle))[which]
    rec = bytes.fromhex(G.G1_HEX)
    assert len(rec) == 167
    b = _batch(G.G1_SCHEMA, dec(G.G1_SCHEMA, [rec] * 4))
    assert b.num_columns == 8 and b.num_rows == 4  # what the reference asserts (deserialize.rs:248-249)
    for row in b.to_pylist():
        assert row["userId"] == G.G1_ROW["userId"]
        assert row["age"] == 28
        assert row["fullName"] == G.G1_ROW["fullName"]
        assert row["email"] == G.G1_ROW["email"]
        assert row["phoneNumbers"] == G.G1_ROW["phoneNumbers"]
        assert row["isPremiumMember"] is False
        assert row["favoriteItems"] == G.G1_ROW["favoriteItems"]
        assert row["registrationDate"] == datetime.datetime(2022, 1, 2, 20, 19, 16)
    assert b.column("registrationDate").cast(pa.int64()).to_pylist() == [G.G1_ROW["registrationDate_ms"]] * 4


@pytest.mark.parametrize("which", ["py", "c"])
def test_g2_values(coracle, which):
    dec = dict(_decoders(coracle))[which]
    rec = bytes.fromhex(G.G2_HEX)
    assert len(rec) == 87
    b = _batch(G.G2_SCHEMA, dec(G.G2_SCHEMA, [rec]))
    assert b.num_columns == 5 and b.num_rows == 1  # deserialize.rs:307-308
    assert b.to_pylist() == [G.G2_ROW]


@pytest.mark.parametrize("which", ["py", "c"])
def test_g345_values(coracle, which):
    dec = dict(_decoders(coracle))[which]
    recs = [bytes.fromhex(h) for h in (G.G3_HEX, G.G4_HEX, G.G5_HEX)]
    assert [len(r) for r in recs] == [162, 113, 81]
    cols = dec(G.G345_SCHEMA, recs)  # G3 carries 5 trailing bytes ("staff") that must be ignored
    b = _batch(G.G345_SCHEMA, cols)
    assert b.num_rows == 3 and b.num_columns == 7
    rows = b.to_pylist()
    for got, want in zip(rows, G.G345_ROWS):
        for key in ("name", "age", "emails", "address", "phone_numbers", "preferences", "status"):
            assert got[key] == want[key], key
    status = cols[6]
    assert status["kind"] == "union" and list(status["buffers"][0]) == [r["status_type_id"] for r in G.G345_ROWS]
    # 3-variant non-null union: field is non-nullable, children all nullable, names varchar/int/bit
    f = b.schema.field("status")
    assert not f.nullable and [f.type.field(i).name for i in range(3)] == ["varchar", "int", "bit"]
    # address is a nullable struct whose validity is always materialised; children carry propagated nulls
    addr = cols[3]
    assert addr["validity"] == bytes([0b010]) and addr["null_count"] == 2
    assert addr["children"][0]["validity"] == bytes([0b010])
    # name never null in rows 1,2 but null in row 0 -> lazy bitmap present
    assert cols[0]["validity"] == bytes([0b110])


def test_oracles_agree_on_goldens(coracle):
    for sj, recs in [(G.G1_SCHEMA, [bytes.fromhex(G.G1_HEX)] * 4), (G.G2_SCHEMA, [bytes.fromhex(G.G2_HEX)]),
                     (G.G345_SCHEMA, [bytes.fromhex(h) for h in (G.G3_HEX, G.G4_HEX, G.G5_HEX)])]:
        a = po.py_decode(po.parse_schema(sj), recs)
        b = coracle.decode(sj, recs)
        assert po.canon_diff(a, b) is None


@pytest.mark.parametrize("seed", range(40))
def test_oracles_agree_on_random_schemas(coracle, seed):
    rng = random.Random(seed)
    sj = po.random_schema_json(rng)
    s = po.parse_schema(sj)
    assert po.is_supported(s) and coracle.is_supported(sj)
    n = rng.choice([0, 1, 7, 33, 100])
    recs = [po.encode_datum(s, po.random_value(s, rng), neg_blocks=rng.random() < 0.3) for _ in range(n)]
    a = po.py_decode(s, recs)
    b = coracle.decode(sj, recs)
    assert po.canon_diff(a, b) is None, sj
    batch = po.canon_to_batch(b, po.to_arrow_schema(s))
    batch.validate(full=True)
    assert po.canon_diff(po.canon_from_batch(batch), b) is None  # canon <-> arrow round trip is lossless
