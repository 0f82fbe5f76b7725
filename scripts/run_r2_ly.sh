#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_affine_gpu.py tests/test_fuzz_gpu.py -x -q 2>&1 | tail -2
{
echo "new"; timeout 300 python scripts/rot90_probe.py 2>&1 | sed -n 2,4p
for v in biahub_b200/_lib/variants/*.so; do echo $v; BIAHUB_B200_LIB=/root/repo/$v timeout 300 python scripts/rot90_probe.py 2>&1 | sed -n 2,4p; done
} | tee gpurun_out/ly_variants.log
