#!/bin/bash
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1; tail -2 $O/r2_smoke.log
python bench.py --workload plate_c5 --no-extra --no-cpu-baseline --steps 5 2> $O/r2_plate_n1.err | grep '^{' > $O/r2_plate_n1.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_plate_n1.json"))
print("plate n1", d["plate"]["value"], d["plate"]["seconds_per_plate"], d["plate"]["ceiling"])
PY
B="python bench.py --workload register_generic --no-extra --no-cpu-baseline --no-e2e --steps 10"
for m in 5 3; do
  B2_NVCC_EXTRA=-DB2_BRICK_MINB=$m python -m biahub_b200._build --force > /dev/null 2>&1
  $B 2> /dev/null | grep '^{' > $O/r2_gen_minb$m.json; cut -c1-200 $O/r2_gen_minb$m.json
done
python -m biahub_b200._build --force > /dev/null 2>&1
$B 2> /dev/null | grep '^{' > $O/r2_gen_minb4.json; cut -c1-200 $O/r2_gen_minb4.json
