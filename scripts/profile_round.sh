#!/bin/bash
# Round profile capture (run under gpurun): launch lists + one full ncu capture per hot kernel.
# Every ncu invocation is preceded by the same command run plainly (exit 0 required).
set -u
R=${1:-r1}
OUT=gpurun_out
for wl in deskew_c2 register_c3 stabilize_c4; do
  CMD="python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --units 2"
  $CMD > $OUT/plain_${wl}_${R}.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
      --log-file $OUT/launches_${wl}_${R}.csv $CMD > $OUT/ncu_launches_${wl}_${R}.log 2>&1
done
CMD="python bench.py --workload deskew_c2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --units 2"
$CMD > $OUT/plain_full_deskew_${R}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:deskew_tma -s 2 -c 1 \
    -o $OUT/prof_deskew_c2_${R} $CMD > $OUT/ncu_full_deskew_${R}.log 2>&1
CMD="python bench.py --workload register_c3 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --units 2"
$CMD > $OUT/plain_full_register_${R}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:affine_zsep -s 2 -c 1 \
    -o $OUT/prof_register_c3_${R} $CMD > $OUT/ncu_full_register_${R}.log 2>&1
CMD="python bench.py --workload stabilize_c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --units 2"
$CMD > $OUT/plain_full_stabilize_${R}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:affine_zsep -s 2 -c 1 \
    -o $OUT/prof_stabilize_c4_${R} $CMD > $OUT/ncu_full_stabilize_${R}.log 2>&1
ls -la $OUT | tail -20
