#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
{
echo "base"; timeout 300 python scripts/ff_time.py 2>&1 | tail -1
for v in biahub_b200/_lib/variants/*.so; do
  echo "$v"; BIAHUB_B200_LIB=/root/repo/$v timeout 300 python scripts/ff_time.py 2>&1 | tail -1
done
} | tee gpurun_out/ff_variants.log
