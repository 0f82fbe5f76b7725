#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r2_t7.log; tail -2 $O/r2_t7.log
python bench.py --workload register_generic --no-extra --no-cpu-baseline --no-e2e --steps 10 2>/dev/null | grep '^{' > $O/r2_gen_h.json; cut -c1-220 $O/r2_gen_h.json
python scripts/spline_small.py > $O/plain_spline.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_spline_r2.csv python scripts/spline_small.py > $O/ncu_spline.log 2>&1
grep -E "spline3|convert" $O/launches_spline_r2.csv | awk -F'","' '{print $5, $NF}' | tail -12
