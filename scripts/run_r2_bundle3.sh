#!/bin/bash
O=gpurun_out
python -m pytest tests/test_affine_gpu.py tests/test_fuzz_gpu.py -m gpu -x -q 2>&1 | tail -3
B="python bench.py --workload register_generic --no-extra --no-cpu-baseline --no-e2e --steps 10"
$B > $O/r2_gen_g.json 2> $O/r2_gen_g.err; cut -c1-260 $O/r2_gen_g.json; tail -2 $O/r2_gen_g.err
B2_BRICK_PREFETCH=0 $B > $O/r2_gen_g0.json 2> $O/r2_gen_g0.err; cut -c1-260 $O/r2_gen_g0.json
B2_BRICK_PREFETCH=1184 $B > $O/r2_gen_g2.json 2> $O/r2_gen_g2.err; cut -c1-260 $O/r2_gen_g2.json
bash scripts/run_r2_boundscheck.sh
