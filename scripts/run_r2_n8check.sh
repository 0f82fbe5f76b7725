#!/bin/bash
cd /root/repo
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29621 bench.py --gpus 8 --steps 10 --warmup 3 2> $O/r2_n8check.err | grep '^{' > $O/r2_n8check.json
timeout 400 $TR --nproc-per-node 8 --master-port 29622 bench.py --gpus 8 --workload plate_c5 --no-extra --steps 5 2> $O/r2_n8plate.err | grep '^{' > $O/r2_n8plate.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_n8check.json")); print(d["n_gpus"], d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"].get("ceiling"), [ (w.get("value"), w.get("roofline",{}).get("frac")) for w in d.get("workloads", [])])
p = json.load(open("gpurun_out/r2_n8plate.json")); print(p["n_gpus"], p["value"], p.get("ms_per_step"), p["config"].get("workload","")[:80])
PY
