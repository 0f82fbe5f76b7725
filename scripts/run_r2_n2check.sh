#!/bin/bash
cd /root/repo
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 2> $O/r2_n2check.err | grep '^{' > $O/r2_n2check.json
timeout 300 $TR --nproc-per-node 2 --master-port 29612 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 2> $O/r2_n2check_ref.err | grep '^{' > $O/r2_n2check_ref.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_n2check.json")); print(d["n_gpus"], d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"].get("ceiling"), len(d.get("workloads", [])))
r = json.load(open("gpurun_out/r2_n2check_ref.json")); print(r["impl"], r["value"], r["n_gpus"])
PY
timeout 300 python -m pytest tests/test_reference_pins_gpu.py::test_host_calls_leave_the_current_device_alone tests/test_workers_gpu.py -q 2>&1 | tail -2
