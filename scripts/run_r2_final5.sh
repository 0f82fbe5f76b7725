#!/bin/bash
# end-of-round verification: full GPU suite, smoke(), the driver's two bench commands, the
# rot90 probe and the launch-list + full captures of the round's final kernels
cd /root/repo
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r2_t11.log; tail -2 $O/r2_t11.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
s=$(date +%s); python bench.py --impl reference --steps 20 --warmup 5 2> $O/r2_final5_ref.err | grep '^{' > $O/r2_final5_ref.json; echo "reference arm wall $(( $(date +%s) - s )) s"
s=$(date +%s); python bench.py --steps 20 --warmup 5 2> $O/r2_final5_bench.err | grep '^{' > $O/r2_final5_bench.json; echo "b200 arm wall $(( $(date +%s) - s )) s"
python scripts/rot90_probe.py > $O/r2_rot90_final.txt 2>&1; cat $O/r2_rot90_final.txt
bash scripts/profile_round.sh r2h
