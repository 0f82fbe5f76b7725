#!/bin/bash
O=gpurun_out
python -m pytest tests/test_affine_gpu.py tests/test_fuzz_gpu.py tests/test_edge_cases_gpu.py tests/test_large_index_gpu.py tests/test_reference_pins_gpu.py -m gpu -x -q 2>&1 | tail -4 > $O/r2_t6.log; tail -2 $O/r2_t6.log
python bench.py --workload register_generic --no-extra --no-cpu-baseline --no-e2e --steps 10 > $O/r2_gen_f.json 2> $O/r2_gen_f.err; cut -c1-260 $O/r2_gen_f.json; tail -2 $O/r2_gen_f.err
bash scripts/profile_round.sh r2
