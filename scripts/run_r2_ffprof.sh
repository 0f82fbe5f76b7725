#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/ff_probe.py > gpurun_out/ff_probe_plain.log 2>&1 || { echo plain failed; tail gpurun_out/ff_probe_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flatfield -c 2 -f -o gpurun_out/prof_ffmedian python scripts/ff_probe.py > gpurun_out/ff_probe_ncu.log 2>&1
tail -3 gpurun_out/ff_probe_ncu.log
ls -la gpurun_out/prof_ffmedian.ncu-rep
