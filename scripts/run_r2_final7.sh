#!/bin/bash
# last verification of the round on a fresh box: what the driver runs
cd /root/repo
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r2_t13.log; tail -2 $O/r2_t13.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
s=$(date +%s); python bench.py --impl reference --steps 20 --warmup 5 2> $O/r2_final7_ref.err | grep '^{' > $O/r2_final7_ref.json; echo "reference arm wall $(( $(date +%s) - s )) s"
s=$(date +%s); python bench.py --steps 20 --warmup 5 2> $O/r2_final7_bench.err | grep '^{' > $O/r2_final7_bench.json; echo "b200 arm wall $(( $(date +%s) - s )) s"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_final7_bench.json"))
print(d["value"], d["roofline"]["frac"], d["roofline"]["launch_ms"], d["e2e"]["value"], d["e2e"]["ceiling"]["value"], d["clocks"])
for w in d["workloads"]: print("  ", w.get("value"), w.get("roofline", {}).get("frac"), w.get("e2e", {}).get("value"), w.get("clocks", {}).get("sm_mhz"))
r = json.load(open("gpurun_out/r2_final7_ref.json")); print(r["value"], r["cpu_baseline"]["kind"], r["config"] == d["config"])
PY
