#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_affine_gpu.py tests/test_fuzz_gpu.py tests/test_full_size_gpu.py -x -q 2>&1 | tail -3
timeout 300 python scripts/rot90_probe.py 2>&1 | tee gpurun_out/ly_variants.log
