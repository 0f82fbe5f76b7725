#!/bin/bash
# full GPU suite + generic/flat-field bench lines + profile of the final generic kernel
cd /root/repo
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r2_t10.log; tail -2 $O/r2_t10.log
python bench.py --workload register_generic --no-extra --no-cpu-baseline --steps 10 --warmup 3 2> $O/r2_final4_gen.err | grep '^{' > $O/r2_final4_gen.json; tail -c 700 $O/r2_final4_gen.json
CMD="python bench.py --workload register_generic --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --units 2"
$CMD > $O/plain_full_register_generic_r2g.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:affine_brick -s 2 -c 1 -f -o $O/prof_register_generic_r2g $CMD > $O/ncu_full_register_generic_r2g.log 2>&1
ls -la $O/prof_register_generic_r2g.ncu-rep
