#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_affine_gpu.py tests/test_fuzz_gpu.py -x -q 2>&1 | tail -2
{
echo "new (+ mid/half from thread 0)"; timeout 300 python scripts/rot90_probe.py 2>&1 | tail -2
echo "vd (zero-fill vec + k*plane addressing)"; BIAHUB_B200_LIB=/root/repo/biahub_b200/_lib/variants/libb2_vd.so timeout 300 python scripts/rot90_probe.py 2>&1 | tail -2
echo "vb (HEAD)"; BIAHUB_B200_LIB=/root/repo/biahub_b200/_lib/variants/libb2_vb.so timeout 300 python scripts/rot90_probe.py 2>&1 | tail -2
} | tee gpurun_out/gen_variants.log
