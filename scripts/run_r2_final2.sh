#!/bin/bash
O=gpurun_out
T0=$(date +%s); python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_final_ref.json 2> $O/r2_final_ref.err; echo "reference arm wall $(( $(date +%s) - T0 )) s" | tee $O/r2_final_wall.txt
T0=$(date +%s); python bench.py --steps 20 --warmup 5 > $O/r2_final_bench.json 2> $O/r2_final_bench.err; echo "b200 arm wall $(( $(date +%s) - T0 )) s" | tee -a $O/r2_final_wall.txt
cut -c1-300 $O/r2_final_bench.json
