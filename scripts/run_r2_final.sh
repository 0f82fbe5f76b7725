#!/bin/bash
# final verification of the round: full GPU suite, the driver's two bench commands (timed), the
# spline probe, then the ncu profile round of the final kernels
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r2_t8.log; tail -2 $O/r2_t8.log
/usr/bin/time -f "reference arm wall %e s" python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_final_ref.json 2> $O/r2_final_ref.err; tail -1 $O/r2_final_ref.err
/usr/bin/time -f "b200 arm wall %e s" python bench.py --steps 20 --warmup 5 > $O/r2_final_bench.json 2> $O/r2_final_bench.err; tail -1 $O/r2_final_bench.err
python scripts/spline_probe.py > $O/r2_spline_probe2.txt 2>&1; head -3 $O/r2_spline_probe2.txt
bash scripts/profile_round.sh r2f
