#!/bin/bash
# after the flat-field rewrite: full GPU suite, the driver's two bench commands, flat-field
# timing (device + host API) and the ncu capture of both flat-field kernels
cd /root/repo
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r2_t9.log; tail -2 $O/r2_t9.log
python bench.py --impl reference --steps 20 --warmup 5 2> $O/r2_final3_ref.err | grep '^{' > $O/r2_final3_ref.json; tail -c 300 $O/r2_final3_ref.json
python bench.py --steps 20 --warmup 5 2> $O/r2_final3_bench.err | grep '^{' > $O/r2_final3_bench.json; tail -c 600 $O/r2_final3_bench.json
python bench.py --workload flatfield --no-extra --steps 10 --warmup 3 2> $O/r2_final3_ff.err | grep '^{' > $O/r2_final3_ff.json; tail -c 900 $O/r2_final3_ff.json
python scripts/flatfield_bench.py > $O/r2_ff_bench.txt 2>&1; cat $O/r2_ff_bench.txt
python scripts/ff_probe.py > $O/ff_probe_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:flatfield -c 2 -f -o $O/prof_flatfield_r2g python scripts/ff_probe.py > $O/ff_probe_ncu.log 2>&1
ls -la $O/prof_flatfield_r2g.ncu-rep
