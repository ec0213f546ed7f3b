#!/usr/bin/env python
"""CPU-only campaigns through the host emulation (tests/emu) at any size — the same comparisons the CPU suite makes on a
few dozen seeds:

    python tools/emu_fuzz.py valid  FIRST_SEED N    random schemas x valid records, sizes around the tile / warp boundaries,
                                                    1..5000 output batches, every buffer against the C oracle
    python tools/emu_fuzz.py wide   FIRST_SEED N    the same over the wider type subset, against the pure-Python oracle
    python tools/emu_fuzz.py gather FIRST_SEED N    2..8 emulated ranks: shard bounds, the product's gather plan and every push
                                                    job applied to one arena, the gathered RecordBatch against the C oracle
"""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    mode, first, count = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    import numpy as np
    from oracle import pyoracle as po
    from pyruhvro_b200 import distributed as D
    from tests import emu
    from tests.parity import assert_matches_oracle, assert_matches_pyoracle_wide, gen_case, gen_case_wide
    co = po.COracle()
    bad, t0 = 0, time.time()
    for seed in range(first, first + count):
        rng = random.Random(seed * 3 + 1)
        try:
            if mode in ("valid", "wide"):
                n = rng.choice([0, 1, 31, 32, 33, 255, 256, 257, 511, 513, 700, 1500])
                k = rng.choice([1, 2, 3, 7, 8, 64, 5000])
                if mode == "wide":
                    sj, recs, data, off = gen_case_wide(seed, n=n)
                    assert_matches_pyoracle_wide(emu.decode(sj, data, off, len(recs), k), sj, recs, k)
                else:
                    sj, recs, data, off = gen_case(seed, n=n)
                    assert_matches_oracle(co, emu.decode(sj, data, off, len(recs), k), sj, data, off, len(recs), k)
            else:
                world = rng.choice([2, 2, 3, 4, 5, 8])
                sj, recs, _, _ = gen_case(seed, n=rng.choice([0, 1, 255, 256, 257, 600, 1000, 1301, 2100, 4097]))
                n = len(recs)
                shards = []
                for r in range(world):
                    r0, r1 = D.shard_bounds(n, world, r)
                    d, o = po.pack_records(recs[r0:r1])
                    shards.append(emu.Shard(sj, d, o, r1 - r0))
                metas = np.stack([s.meta() for s in shards])
                groups = shards[0].groups(metas)
                assert len(groups) == 1 and groups[0][1] == world and groups[0][3] == n, groups
                arena = np.zeros(max(groups[0][2], 64), dtype=np.uint8)
                for r in range(world):       # (the ranks' pushes touch disjoint bytes except for OR-merged bitmap seams)
                    mine = np.zeros_like(arena)
                    shards[r].apply(metas, r, mine)
                    arena |= mine
                batch = shards[0].export(metas, 0, arena)
                batch.validate(full=True)
                diff = po.canon_diff(po.canon_from_batch(batch), co.decode(sj, recs))
                assert diff is None, diff
        except AssertionError as e:
            bad += 1
            print(f"FAIL seed={seed}: {str(e)[:300]}", flush=True)
    print(f"emu fuzz ({mode}): {count} cases, failures {bad}, {int(time.time() - t0)} s")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
