#!/bin/bash
# flat-field median rewrite: parity tests, then timing of the mantis-sized volume
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_flatfield_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/ff_tests.log
cat gpurun_out/ff_tests.log
timeout 300 python scripts/ff_time.py 2>&1 | tail -2 | tee gpurun_out/ff_bench.log
