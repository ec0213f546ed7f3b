"""GPU parity of the Arrow -> Avro direct encode (serialize_record_batch) against the
pure-Python restatement of fast_encode.rs (oracle.pyoracle.py_encode) and against the round-trip property the
reference's own tests use (fast_encode.rs:616-637: avro -> decode -> encode gives the datums back)."""
import random

import pyarrow as pa
import pytest

import pyruhvro_b200 as pr
from oracle import pyoracle as po
from tests.golden import reference_datums as G

pytestmark = pytest.mark.gpu


def _datums(arrays):
    return [bytes(x.as_py()) for a in arrays for x in a]


@pytest.mark.parametrize("seed", range(40))
def test_round_trip_and_oracle(coracle, seed):
    rng = random.Random(seed)
    sj = po.random_schema_json(rng)
    s = po.parse_schema(sj)
    n = rng.choice([1, 33, 257, 700])
    recs = [po.encode_datum(s, po.random_value(s, rng)) for _ in range(n)]
    batch = pr.deserialize_array(recs, sj)
    k = rng.choice([1, 2, 5])
    out = pr.serialize_record_batch(batch, sj, k)
    assert len(out) == po.clamp_chunks(k, n) and all(a.type == pa.binary() for a in out)
    assert [len(a) for a in out] == [b - a for a, b in po.chunk_bounds(n, po.clamp_chunks(k, n))]
    assert _datums(out) == recs                                     # avro -> arrow -> avro is the identity
    want = po.py_encode(s, po.canon_to_batch(coracle.decode(sj, recs), po.to_arrow_schema(s)), k)
    assert [[bytes(x.as_py()) for x in a] for a in out] == want     # and equals the fast_encode.rs restatement


def test_reference_golden_datums():
    recs = [bytes.fromhex(h) for h in (G.G3_HEX, G.G4_HEX, G.G5_HEX)]
    batch = pr.deserialize_array(recs, G.G345_SCHEMA)
    out = _datums(pr.serialize_record_batch(batch, G.G345_SCHEMA, 1))   # what lib.rs:174-178 does (asserting only "no error")
    assert out[0] == recs[0][:157] and out[1] == recs[1] and out[2] == recs[2]
    for sj, hx in ((G.G1_SCHEMA, G.G1_HEX), (G.G2_SCHEMA, G.G2_HEX)):
        r = bytes.fromhex(hx)
        assert _datums(pr.serialize_record_batch(pr.deserialize_array([r, r], sj), sj, 2)) == [r, r]


def test_columns_matched_by_name_and_slices():
    import workloads
    sj, data, off = workloads.generate("kafka", 3000, seed=3)
    recs = [bytes(data[off[i]:off[i + 1]]) for i in range(3000)]
    batch = pr.deserialize_array(recs, sj)
    names = list(batch.schema.names)
    shuffled = pa.RecordBatch.from_arrays([batch.column(n) for n in reversed(names)], names=list(reversed(names)))
    assert _datums(pr.serialize_record_batch(shuffled, sj, 3)) == recs       # fast_encode.rs:157-181
    sl = batch.slice(100, 1234)                                              # offsets / slices are honoured
    assert _datums(pr.serialize_record_batch(sl, sj, 4)) == recs[100:1334]
    extra = shuffled.append_column("unused", pa.array(range(3000)))
    assert _datums(pr.serialize_record_batch(extra, sj, 1)) == recs


def test_large_batch_takes_the_staged_upload_and_round_trips():
    """Above 4 MiB of Arrow buffers the upload goes through the pinned gather pipeline (encode.cu); the bytes
    that come back must still be the input datums (decode -> encode identity), for whole and sliced batches."""
    import numpy as np
    import workloads
    n = 200_000
    sj, data, off = workloads.generate("kafka", n, seed=11)
    batch = pr.decode_packed(data, off, n, sj, 1)[0]
    assert batch.nbytes > (16 << 20)
    for b, lo, k in ((batch, 0, 3), (batch.slice(12_345, 150_000), 12_345, 2)):
        out = pr.serialize_record_batch(b, sj, k)
        offs = [np.frombuffer(a.buffers()[1], dtype=np.int32)[:len(a) + 1] for a in out]
        got = np.concatenate([np.frombuffer(a.buffers()[2], dtype=np.uint8)[:o[-1]] for a, o in zip(out, offs)])
        assert got.tobytes() == data[off[lo]:off[lo + b.num_rows]].tobytes()
        lens = np.concatenate([np.diff(o) for o in offs])
        assert np.array_equal(lens, np.diff(off[lo:lo + b.num_rows + 1]))


def test_row_groups_keep_each_chunk_alive_on_its_own():
    """From 2^18 rows the chunks are encoded in concurrent row groups (encode.cu: rv_encode_host).  Every exported array
    must own ITS chunk's memory: drop all but one array, churn the pinned cache with further calls, and the survivor
    still holds its datums."""
    import gc
    import numpy as np
    import workloads
    n, k = 600_000, 8
    sj, data, off = workloads.generate("kafka", n, seed=5)
    batch = pr.decode_packed(data, off, n, sj, 1)[0]
    rows = n // k
    for keep in (0, 3, 7):
        out = pr.serialize_record_batch(batch, sj, k)
        assert [len(a) for a in out] == [rows] * (k - 1) + [n - rows * (k - 1)]
        survivor = out[keep]
        del out
        gc.collect()
        for _ in range(2):                                          # re-uses whatever the dropped arrays gave back
            pr.serialize_record_batch(batch.slice(17, 300_000), sj, 5)
        lo = keep * rows
        o = np.frombuffer(survivor.buffers()[1], dtype=np.int32)[:len(survivor) + 1]
        got = np.frombuffer(survivor.buffers()[2], dtype=np.uint8)[:o[-1]]
        assert got.tobytes() == data[off[lo]:off[lo + len(survivor)]].tobytes()
        assert np.array_equal(np.diff(o), np.diff(off[lo:lo + len(survivor) + 1]))


def test_error_surface():
    recs = [bytes.fromhex(G.G2_HEX)]
    batch = pr.deserialize_array(recs, G.G2_SCHEMA)
    with pytest.raises(ValueError) as e:
        pr.serialize_record_batch(batch.drop_columns(["age"]), G.G2_SCHEMA, 1)
    assert "Arrow struct missing column 'age' required by Avro schema. Available columns:" in str(e.value)   # :171-179
    with pytest.raises(ValueError):
        pr.serialize_record_batch(batch.set_column(2, "age", pa.array(["x"])), G.G2_SCHEMA, 1)               # downcast failure
    sj = '{"type":"record","name":"E","fields":[{"name":"e","type":{"type":"enum","name":"S","symbols":["A","B"]}}]}'
    with pytest.raises(ValueError) as e:
        pr.serialize_record_batch(pa.record_batch({"e": ["A", "Z", "B"]}), sj, 1)
    assert "enum symbol" in str(e.value) and "(row 1)" in str(e.value)                                        # :573-576
    assert _datums(pr.serialize_record_batch(pa.record_batch({"e": ["A", "B", "B"]}), sj, 1)) == [b"\x00", b"\x02", b"\x02"]
    out = pr.serialize_record_batch(batch.slice(0, 0), G.G2_SCHEMA, 8)
    assert len(out) == 1 and len(out[0]) == 0                                                                  # n = 0 -> one empty chunk


def test_null_slots_of_non_nullable_fields_encode_their_raw_value():
    """`array.value(row)` ignores validity for non-nullable Avro fields (fast_encode.rs:401-409)."""
    sj = '{"type":"record","name":"N","fields":[{"name":"a","type":"long"},{"name":"b","type":["null","long"]}]}'
    b = pa.record_batch({"a": pa.array([1, None, 3], pa.int64()), "b": pa.array([None, 5, None], pa.int64())})
    got = _datums(pr.serialize_record_batch(b, sj, 1))
    want = [d for ch in po.py_encode(po.parse_schema(sj), b, 1) for d in ch]
    assert got == want == [b"\x02\x00", b"\x00\x02\x0a", b"\x06\x00"]
