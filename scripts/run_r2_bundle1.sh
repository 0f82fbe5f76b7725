#!/bin/bash
# GPU bundle: full GPU suite, generic-kernel bench, then compute-sanitizer memcheck
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r2_t5.log; tail -3 $O/r2_t5.log
B="python bench.py --workload register_generic --no-extra --no-cpu-baseline --no-e2e --steps 10"
$B > $O/r2_gen_e.json 2> $O/r2_gen_e.err; cut -c1-260 $O/r2_gen_e.json; tail -2 $O/r2_gen_e.err
python scripts/sanitize_cases.py > $O/r2_sanitize_plain.log 2>&1 && \
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_cases.py > $O/r2_sanitize_memcheck.log 2>&1
tail -5 $O/r2_sanitize_plain.log | head -3; tail -6 $O/r2_sanitize_memcheck.log
