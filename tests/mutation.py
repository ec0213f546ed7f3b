"""Random byte-level damage to valid Avro datums, and the comparison every decoder under test must survive: the C oracle
(the reference's semantics) decides whether the damaged batch still decodes — then the buffers must match — or which record
fails first with which error category — then the decoder must report exactly that.  Shared by tests/test_mutation_fuzz.py
(host emulation of the product's readers), tests/test_zz_gpu_damaged_inputs.py (the CUDA path; named to run last) and tools/mutation_fuzz.py."""
import random

from oracle import pyoracle as po
from tests.parity import assert_matches_oracle, gen_case

# found by the fuzz (seed 300386): a map whose forged block count is i64::MAX and whose item record ENDS in a union — the
# fast reader's end-of-buffer check parks the cursor AT the end, the item loop's own check then never sees it past the end
HANG_SCHEMA = ('{"type": "record", "name": "Top", "fields": [{"name": "c1", "type": {"type": "array", "items": "float"}}, '
               '{"name": "c2", "type": {"type": "array", "items": {"type": "int", "logicalType": "date"}}}, '
               '{"name": "c3", "type": {"type": "record", "name": "R4", "fields": [{"name": "f5", "type": "boolean"}, '
               '{"name": "f6", "type": "double"}, {"name": "f7", "type": "long"}]}}, '
               '{"name": "c8", "type": {"type": "map", "values": {"type": "record", "name": "R9", "fields": ['
               '{"name": "f10", "type": {"type": "enum", "name": "E11", "symbols": ["S0", "S1S1", "S2S2S2", "S3"]}}, '
               '{"name": "f12", "type": "string"}, {"name": "f13", "type": {"type": "enum", "name": "E14", "symbols": ["S0S0S0", "S1S1"]}}, '
               '{"name": "f15", "type": ["long", "null", "string", {"type": "enum", "name": "E16", "symbols": ["S0", "S1S1S1"]}]}]}}}, '
               '{"name": "c17", "type": ["null", {"type": "int", "logicalType": "date"}]}, {"name": "c18", "type": {"type": "int", "logicalType": "date"}}]}')
HANG_RECORD = bytes.fromhex(
    "0004b7cc277f00017cbf9b10ffffffffffffff42feffffffffffffffff01040c6b33363333340400000422776f6a6be29c937a666d657861716c"
    "63780a6b3934373000486ae29c936d6fe29c93647a7463626871686f686a6c61656b676e7478646675c3a96179680004407a626e786a6c73716e"
    "626a6475766f6477766864706a76c3a9627466737463680000b6b010")

# a block count of i64::MIN: `-n` wraps in the reference's release build, `0..n` is an empty range, the next block header follows
MIN_BLOCK_SCHEMA = '{"type":"record","name":"R","fields":[{"name":"m","type":{"type":"map","values":"long"}},{"name":"x","type":"int"}]}'
MIN_BLOCK_RECORD = bytes.fromhex("ffffffffffffffffff01" "00" "02" "026b" "54" "00" "0e")   # i64::MIN, size 0 | 1 entry k -> 42 | end | x = 7


def damage(rng: random.Random, recs):
    """A few byte-level mutations (overwrite, 0xFF runs, truncate, extend, insert) in a few of the records."""
    recs = [bytearray(r) for r in recs]
    for _ in range(rng.choice([1, 1, 2, 4])):
        r = recs[rng.randrange(len(recs))]
        op = rng.choice(["set", "set", "trunc", "ext", "ins", "ff"])
        if op == "set" and r:
            r[rng.randrange(len(r))] = rng.randrange(256)
        elif op == "ff" and r:
            j = rng.randrange(len(r))
            r[j:j + rng.randrange(1, 12)] = b"\xff" * rng.randrange(1, 12)
        elif op == "trunc" and r:
            del r[rng.randrange(len(r)):]
        elif op == "ext":
            r.extend(bytes(rng.randrange(256) for _ in range(rng.randrange(1, 9))))
        elif op == "ins":
            j = rng.randrange(len(r) + 1)
            r[j:j] = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 5)))
    return [bytes(r) for r in recs]


def damaged_case(seed: int, schema_seed=None):
    rng = random.Random(seed * 977 + 5)
    sj, recs, _, _ = gen_case(seed if schema_seed is None else schema_seed, n=rng.choice([3, 40, 257, 300]))
    recs = damage(rng, recs)
    return sj, recs, rng.choice([1, 2, 3, 8])


def expected(coracle, sj, recs):
    """None when the batch still decodes, else (error category, index of the first failing record)."""
    try:
        coracle.decode(sj, recs)
        return None
    except po.DecodeError as e:
        return (e.code, e.record)


def check(coracle, decode, error_of, sj, recs, k):
    """`decode(sj, data, off, n, k)` -> batches or raises; `error_of(exc)` -> (category, record) or None if `exc` is not a
    decode error of the implementation under test."""
    data, off = po.pack_records(recs)
    want = expected(coracle, sj, recs)
    try:
        got = decode(sj, data, off, len(recs), k)
    except Exception as e:  # noqa: BLE001 - classified right below
        g = error_of(e)
        if g is None:
            raise
        assert want is not None, f"decoder reports {g}, the reference decodes this batch"
        assert g == want, f"decoder reports {g}, the reference {want}"
        return "error"
    assert want is None, f"decoder accepted a batch the reference rejects with {want}"
    assert_matches_oracle(coracle, got, sj, data, off, len(recs), k, full_validate=False)   # (damaged strings need not be UTF-8)
    return "decoded"


# ---- forged varints: structured damage (a second generator; `damage` above and its seeds stay as they are) --------------
# Values a length / count / index / branch reader has an edge at, canonical and not.
FORGED_INTS = [0, 1, -1, 2, 63, 64, -64, -65, 127, 128, 255, 256, 257, 300, 8191, 8192, 2**20, 2**30, -2**30, 2**31 - 2, 2**31 - 1, 2**31,
               -2**31, -2**31 - 1, 2**32 - 1, 2**32, 2**42, 2**56, 2**62, -2**62, 2**63 - 1, -2**63 + 1, -2**63]


def _leb128(u: int) -> bytes:
    out = bytearray()
    while True:
        b = u & 0x7F
        u >>= 7
        if u:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def forged_varint(rng: random.Random) -> bytes:
    r = rng.random()
    if r < 0.6:
        return po.zigzag_bytes(rng.choice(FORGED_INTS))
    if r < 0.75:   # not canonical: the same value padded with continuation bytes
        b = bytearray(po.zigzag_bytes(rng.choice(FORGED_INTS[:12])))
        b[-1] |= 0x80
        return bytes(b) + b"\x80" * rng.randrange(0, 8) + b"\x00"
    if r < 0.85:   # raw 64-bit patterns (the top bit of the tenth byte, all ones)
        return _leb128(rng.choice([2**64 - 1, 2**63, 2**64 - 2, 2**35, 2**63 - 1]))
    if r < 0.93:   # too long
        return b"\xff" * rng.randrange(9, 12) + bytes([rng.randrange(256)])
    return po.zigzag_bytes(rng.randrange(-2**63, 2**63))


def forge_varints(rng: random.Random, recs):
    """One to three forged varints per batch, each either spliced in or written over the varint that starts at a random byte."""
    recs = [bytearray(r) for r in recs]
    for _ in range(rng.choice([1, 1, 2, 3])):
        r = recs[rng.randrange(len(recs))]
        v = forged_varint(rng)
        j = rng.randrange(len(r) + 1)
        if rng.random() < 0.5 and j < len(r):
            e = j
            while e < len(r) and r[e] & 0x80 and e - j < 10:
                e += 1
            r[j:e + 1] = v
        else:
            r[j:j] = v
    return [bytes(r) for r in recs]


def forged_case(seed: int, schema_seed=None):
    rng = random.Random(seed * 7919 + 11)
    sj, recs, _, _ = gen_case(seed if schema_seed is None else schema_seed, n=rng.choice([3, 40, 257, 300]))
    return sj, forge_varints(rng, recs), rng.choice([1, 2, 3, 8])


# ---- the wider subset: the checker is the pure-Python oracle (the reference cannot decode these schemas at all) ----------
def damaged_case_wide(seed: int):
    from tests.parity import gen_case_wide
    rng = random.Random(seed * 977 + 5)
    sj, recs, _, _ = gen_case_wide(seed, n=rng.choice([3, 40, 257]))
    return sj, damage(rng, recs), rng.choice([1, 2, 3])


def expected_wide(sj, recs):
    s = po.parse_schema(sj, wide=True)
    for i, r in enumerate(recs):      # py_decode raises without an index: first failing record, one at a time
        try:
            po.py_decode(s, [r])
        except po.DecodeError as e:
            return (e.code, i)
    return None


def check_wide(decode, error_of, sj, recs, k):
    from tests.parity import assert_matches_pyoracle_wide
    data, off = po.pack_records(recs)
    want = expected_wide(sj, recs)
    try:
        got = decode(sj, data, off, len(recs), k)
    except Exception as e:  # noqa: BLE001
        g = error_of(e)
        if g is None:
            raise
        assert g == want, f"decoder reports {g}, the oracle {want}"
        return "error"
    assert want is None, f"decoder accepted a batch the oracle rejects with {want}"
    assert_matches_pyoracle_wide(got, sj, recs, k, full_validate=False)
    return "decoded"
