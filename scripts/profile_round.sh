#!/bin/bash
# Round profile capture (run under gpurun): launch lists + one full ncu capture per hot kernel.
# Every ncu invocation is preceded by the same command run plainly (exit 0 required).
set -u
R=${1:-r1}
OUT=gpurun_out
for wl in deskew_c2 deskew_c1 register_c3 stabilize_c4 register_generic; do
  CMD="python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --units 2"
  $CMD > $OUT/plain_${wl}_${R}.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
      --log-file $OUT/launches_${wl}_${R}.csv $CMD > $OUT/ncu_launches_${wl}_${R}.log 2>&1
done
for pair in "deskew_c2:deskew_" "deskew_c1:deskew_" "register_c3:affine_zsep" "stabilize_c4:affine_zsep" "register_generic:affine_brick"; do
  wl=${pair%%:*}; pat=${pair##*:}
  CMD="python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --units 2"
  $CMD > $OUT/plain_full_${wl}_${R}.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 2 -c 1 \
      -o $OUT/prof_${wl}_${R} $CMD > $OUT/ncu_full_${wl}_${R}.log 2>&1
done
ls $OUT | grep "_${R}" | wc -l
