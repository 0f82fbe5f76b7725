#!/bin/bash
cd /root/repo
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r2_t12.log; tail -2 $O/r2_t12.log
python bench.py --workload register_generic --no-extra --no-cpu-baseline --steps 10 --warmup 3 2> $O/r2_final6_gen.err | grep '^{' > $O/r2_final6_gen.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_final6_gen.json")); print(d["value"], d["roofline"]["frac"], d["roofline"]["launch_ms"], d["e2e"]["value"], d["clocks"])
PY
python scripts/rot90_probe.py > $O/r2_rot90_final.txt 2>&1; cat $O/r2_rot90_final.txt
