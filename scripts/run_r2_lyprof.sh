#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/ly_probe.py scaled > gpurun_out/ly_probe_plain.log 2>&1 || { echo plain failed; tail gpurun_out/ly_probe_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:affine_zsep -s 2 -c 1 -f -o gpurun_out/prof_zsep_ly_scaled_final python scripts/ly_probe.py scaled > gpurun_out/ly_probe_ncu.log 2>&1
tail -2 gpurun_out/ly_probe_ncu.log
