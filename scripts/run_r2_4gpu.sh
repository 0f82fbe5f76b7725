#!/bin/bash
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L > $O/r2_4gpu_devices.txt
timeout 400 $TR --nproc-per-node 4 --master-port 29611 bench.py --gpus 4 --workload plate_c5 --no-extra --steps 5 2> $O/r2_plate_n4.err | grep '^{' > $O/r2_plate_n4.json
timeout 400 $TR --nproc-per-node 2 --master-port 29612 bench.py --gpus 2 --workload plate_c5 --no-extra --steps 5 2> $O/r2_plate_n2.err | grep '^{' > $O/r2_plate_n2.json
timeout 300 $TR --nproc-per-node 4 --master-port 29613 bench.py --gpus 4 --no-extra --steps 5 2> $O/r2_bench_n4.err | grep '^{' > $O/r2_bench_n4.json
python - <<'PY'
import json
for f in ("r2_plate_n4", "r2_plate_n2", "r2_bench_n4"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["value"], d["e2e"]["value"], d["e2e"].get("ceiling"), d.get("plate", {}).get("seconds_per_plate"))
    except Exception as e:
        print(f, "ERR", e)
PY
