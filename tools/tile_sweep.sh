#!/bin/bash
# Tile-size variants of the library (built with RV_KBLOCK=N into build/variants/; development aid).
cd "$(dirname "$0")/.."
echo "== default (256)"; timeout 150 python tools/sweep_jit.py --records 10000000 --steps 10 2>&1 | tail -1
echo "== default C4"; timeout 150 python tools/sweep_jit.py --workload wide --records 10000000 --steps 5 2>&1 | tail -1
for spec in "224 4" "224 3" "320 2" "384 2"; do
  set -- $spec
  echo "== kblock $1 minb $2"; RV_LIB_PATH=build/variants/lib_k$1.so RV_JIT_MINB=$2 timeout 200 python tools/sweep_jit.py --records 10000000 --steps 10 2>&1 | tail -1
done
