#!/bin/bash
# One GPU visit, several measurements (development aid).  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== C4 ablations"
for d in "" "RV_ABL_NOCOPY=1" "RV_ABL_NOWALK=1" "RV_ABL_NOWALK=1;RV_ABL_NOCOUNTWALK=1" "RV_ABL_NOBITS=1" "RV_ABL_NOSTORE=1"; do
  echo "-- $d"; RV_JIT_DEFS="$d" timeout 150 python tools/sweep_jit.py --workload wide --records 10000000 --steps 5 2>&1 | tail -1
done
echo "== C4 source-level capture"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rvj_ -s 4 -c 1 -o gpurun_out/r02_c4 python tools/sweep_jit.py --workload wide --records 2000000 --steps 2 > gpurun_out/r02_c4.log 2>&1; tail -1 gpurun_out/r02_c4.log
echo "== C3 full capture (10 M records) + launch list of the bench command"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rvj_ -s 4 -c 1 -o gpurun_out/r02_c3 python tools/sweep_jit.py --records 10000000 --steps 2 > gpurun_out/r02_c3.log 2>&1; tail -1 gpurun_out/r02_c3.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/r02_launches_bench.log 2>&1; tail -c 300 gpurun_out/r02_launches_bench.log
