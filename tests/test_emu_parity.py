"""Logic parity of the product's walker / plan / layout / Arrow export (run through the host
emulation in tests/emu) against the C oracle.  The CUDA kernels themselves are covered by the
`-m gpu` tests in test_gpu_parity.py with the same helpers."""
import random

import numpy as np
import pytest

from oracle import pyoracle as po
from tests import emu
from tests.golden import reference_datums as G
from tests.parity import assert_matches_oracle, gen_case


def test_goldens(coracle):
    for sj, recs in [(G.G1_SCHEMA, [bytes.fromhex(G.G1_HEX)] * 4), (G.G2_SCHEMA, [bytes.fromhex(G.G2_HEX)]),
                     (G.G345_SCHEMA, [bytes.fromhex(h) for h in (G.G3_HEX, G.G4_HEX, G.G5_HEX)])]:
        data, off = po.pack_records(recs)
        for k in (1, 2, 8):
            assert_matches_oracle(coracle, emu.decode(sj, data, off, len(recs), k), sj, data, off, len(recs), k)


@pytest.mark.parametrize("seed", range(60))
def test_random_schemas(coracle, seed):
    sj, recs, data, off = gen_case(seed)
    k = random.Random(seed).choice([1, 1, 2, 3, 8, 1000])
    assert_matches_oracle(coracle, emu.decode(sj, data, off, len(recs), k), sj, data, off, len(recs), k)


@pytest.mark.parametrize("seed", range(100, 112))
def test_generated_walker_random_schemas(coracle, seed):
    """The schema-specialised walker source that NVRTC compiles for the GPU, compiled for the host."""
    sj, recs, data, off = gen_case(seed)
    k = random.Random(seed).choice([1, 2, 5])
    assert_matches_oracle(coracle, emu.decode(sj, data, off, len(recs), k, walker="gen"), sj, data, off, len(recs), k)


def test_generated_walker_goldens_and_errors(coracle):
    from tests import malformed as M
    for sj, recs in [(G.G1_SCHEMA, [bytes.fromhex(G.G1_HEX)] * 4), (G.G2_SCHEMA, [bytes.fromhex(G.G2_HEX)]),
                     (G.G345_SCHEMA, [bytes.fromhex(h) for h in (G.G3_HEX, G.G4_HEX, G.G5_HEX)])]:
        data, off = po.pack_records(recs)
        assert_matches_oracle(coracle, emu.decode(sj, data, off, len(recs), 2, walker="gen"), sj, data, off, len(recs), 2)
    good = [M.good_record(i) for i in range(300)]
    for name, code, bad in M.cases():
        recs = list(good)
        recs[37] = bad
        data, off = po.pack_records(recs)
        with pytest.raises(emu.EmuError) as ee:
            emu.decode(M.FLAT, data, off, len(recs), 3, walker="gen")
        assert (po.ERR_NAMES[ee.value.code], ee.value.record) == (code, 37), name


def test_empty_input(coracle):
    sj = G.G345_SCHEMA
    data, off = po.pack_records([])
    b = emu.decode(sj, data, off, 0, 8)
    assert len(b) == 1 and b[0].num_rows == 0  # deserialize.rs:53-55,81: n = 0 -> exactly one empty batch
    assert_matches_oracle(coracle, b, sj, data, off, 0, 8)


@pytest.mark.parametrize("n,k,sizes", [(10, 4, [2, 2, 2, 4]), (10, 0, [10]), (3, 8, [1, 1, 1]), (7, 7, [1] * 7)])
def test_chunk_partition(coracle, n, k, sizes):
    """build_slices semantics (deserialize.rs:57-68): floor division, remainder to the last chunk."""
    sj, recs, data, off = gen_case(5, n=n)
    b = emu.decode(sj, data, off, n, k)
    assert [x.num_rows for x in b] == sizes
    assert_matches_oracle(coracle, b, sj, data, off, n, k)


def test_malformed_inputs_match_oracle(coracle):
    from tests import malformed as M
    good = [M.good_record(i) for i in range(300)]
    for name, code, bad in M.cases():
        for pos in (0, 37, 299):
            recs = list(good)
            recs[pos] = bad
            data, off = po.pack_records(recs)
            with pytest.raises(po.DecodeError) as oe:
                coracle.decode(M.FLAT, recs)
            assert (oe.value.code, oe.value.record) == (code, pos), name
            with pytest.raises(emu.EmuError) as ee:
                emu.decode(M.FLAT, data, off, len(recs), 3)
            assert (po.ERR_NAMES[ee.value.code], ee.value.record) == (code, pos), name


def test_trailing_bytes_ignored(coracle):
    from tests import malformed as M
    recs = [M.good_record(i) + b"junk" * (i % 3) for i in range(40)]
    data, off = po.pack_records(recs)
    assert_matches_oracle(coracle, emu.decode(M.FLAT, data, off, len(recs), 2), M.FLAT, data, off, len(recs), 2)


def test_zero_width_items_and_deep_nesting(coracle):
    sj = '{"type":"record","name":"Z","fields":[{"name":"z","type":{"type":"array","items":"null"}},' \
         '{"name":"m","type":{"type":"map","values":{"type":"array","items":{"type":"array","items":["null","string"]}}}}]}'
    s = po.parse_schema(sj)
    rng = random.Random(3)
    recs = [po.encode_datum(s, po.random_value(s, rng)) for _ in range(300)]
    data, off = po.pack_records(recs)
    assert_matches_oracle(coracle, emu.decode(sj, data, off, len(recs), 2), sj, data, off, len(recs), 2)
