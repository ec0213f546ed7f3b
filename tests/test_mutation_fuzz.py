"""Damaged inputs through the host emulation of the product's readers (tests/emu: fast count -> precise repeat -> emit,
dev_core.cuh / interp.cuh / the generated walkers compiled for the host) against the C oracle: same buffers, or the same
first failing record with the same error category.  tools/mutation_fuzz.py runs the same comparison at any size."""
import pytest

from oracle import pyoracle as po
from tests import emu, mutation as M


def _emu(walker):
    return lambda sj, data, off, n, k: emu.decode(sj, data, off, n, k, walker=walker)


def _emu_error(e):
    return (po.ERR_NAMES.get(e.code, str(e.code)), e.record) if isinstance(e, emu.EmuError) else None


@pytest.mark.timeout(120)
@pytest.mark.parametrize("walker", ["interp", "gen"])
def test_forged_block_count_ends_with_the_records_bytes(coracle, walker):
    """A forged list / map count must cost no more than the record's bytes, whatever the item's last node is (the case the
    fuzz found walked 2^63 items: the cursor parked AT the end of the record never tripped the item loop's own check)."""
    recs = [M.HANG_RECORD[:40]] * 4 + [M.HANG_RECORD] + [M.HANG_RECORD[:40]] * 3
    assert M.expected(coracle, M.HANG_SCHEMA, [M.HANG_RECORD]) == ("enum", 0)
    data, off = po.pack_records(recs)
    with pytest.raises(emu.EmuError) as ee:
        emu.decode(M.HANG_SCHEMA, data, off, len(recs), 2, walker=walker)
    want = M.expected(coracle, M.HANG_SCHEMA, recs)
    assert _emu_error(ee.value) == want and want[1] <= 4


def test_block_count_of_i64_min_is_an_empty_block_in_every_implementation(coracle):
    """`Ok(-n)` (fast_decode.rs:695) wraps for i64::MIN in the release build: the count stays negative, `0..n` is empty and
    the next block header is read.  Both oracles and the product's readers agree (the C oracle used to negate with signed
    overflow, the Python one with unbounded integers)."""
    sj, r = M.MIN_BLOCK_SCHEMA, M.MIN_BLOCK_RECORD
    want = po.py_decode(po.parse_schema(sj), [r])
    assert po.canon_diff(coracle.decode(sj, [r]), want) is None
    data, off = po.pack_records([r])
    for walker in ("interp", "gen"):
        b = emu.decode(sj, data, off, 1, 1, walker=walker)[0]
        assert b.column("x").to_pylist() == [7] and b.column("m").to_pylist() == [[("k", 42)]]
        assert po.canon_diff(po.canon_from_batch(b), want) is None


@pytest.mark.parametrize("seed", range(910000, 910300))
def test_damaged_batches_interpreter(coracle, seed):
    sj, recs, k = M.damaged_case(seed)
    M.check(coracle, _emu("interp"), _emu_error, sj, recs, k)


@pytest.mark.parametrize("seed", range(920000, 920060))
def test_damaged_batches_generated_walkers(coracle, seed):
    sj, recs, k = M.damaged_case(seed, schema_seed=100 + seed % 12)   # the schemas test_emu_parity.py already compiles
    M.check(coracle, _emu("gen"), _emu_error, sj, recs, k)


@pytest.mark.parametrize("seed", range(960000, 960200))
def test_forged_varints_interpreter(coracle, seed):
    """Structured damage: edge-value, padded, over-long and raw 64-bit varints spliced in or written over existing ones."""
    sj, recs, k = M.forged_case(seed)
    M.check(coracle, _emu("interp"), _emu_error, sj, recs, k)


@pytest.mark.parametrize("seed", range(970000, 970060))
def test_forged_varints_generated_walkers(coracle, seed):
    sj, recs, k = M.forged_case(seed, schema_seed=100 + seed % 12)
    M.check(coracle, _emu("gen"), _emu_error, sj, recs, k)


@pytest.mark.parametrize("seed", range(950000, 950100))
def test_damaged_batches_wider_subset(seed):
    """bytes / fixed / uuid / decimal / time-* / named references: value errors (RV_ERR_VALUE) included."""
    sj, recs, k = M.damaged_case_wide(seed)
    M.check_wide(_emu("interp"), _emu_error, sj, recs, k)
