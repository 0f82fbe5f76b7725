#!/bin/bash
# GPU bundle: full GPU suite, the three generic-kernel variants, then compute-sanitizer memcheck
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r2_t5.log; tail -3 $O/r2_t5.log
B="python bench.py --workload register_generic --no-extra --no-cpu-baseline --no-e2e --steps 10"
$B > $O/r2_gen_pers.json 2> $O/r2_gen_pers.err; cut -c1-260 $O/r2_gen_pers.json
B2_BRICK_PERSISTENT=0 $B > $O/r2_gen_tz16.json 2> $O/r2_gen_tz16.err; cut -c1-260 $O/r2_gen_tz16.json
B2_BRICK_PERSISTENT=0 B2_BRICK_TZ=8 $B > $O/r2_gen_tz8.json 2> $O/r2_gen_tz8.err; cut -c1-260 $O/r2_gen_tz8.json
python scripts/sanitize_cases.py > $O/r2_sanitize_plain.log 2>&1 && \
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_cases.py > $O/r2_sanitize_memcheck.log 2>&1
tail -5 $O/r2_sanitize_plain.log | head -3; tail -6 $O/r2_sanitize_memcheck.log
