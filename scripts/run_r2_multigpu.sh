#!/bin/bash
# 8-GPU bundle (one call): host<->device ceilings lever by lever, the 2-GPU-only tests, the plate
# run and the default bench at N = 2 / 8.  Everything under `timeout`.
O=gpurun_out
nvidia-smi topo -m > $O/r2_topo.txt 2>&1
free -g | head -2 > $O/r2_host.txt; nproc >> $O/r2_host.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 8 --master-port 29601 scripts/pcie_probe2.py > $O/r2_pcie_n8.txt 2>&1
timeout 200 $TR --nproc-per-node 2 --master-port 29602 scripts/pcie_probe2.py > $O/r2_pcie_n2.txt 2>&1
timeout 200 $TR --nproc-per-node 4 --master-port 29603 scripts/pcie_probe2.py > $O/r2_pcie_n4.txt 2>&1
grep -h "alloc=" $O/r2_pcie_n8.txt $O/r2_pcie_n2.txt $O/r2_pcie_n4.txt | cut -c1-230
timeout 300 python -m pytest tests/test_reference_pins_gpu.py::test_host_calls_leave_the_current_device_alone tests/test_workers_gpu.py -q 2>&1 | tail -3
timeout 400 $TR --nproc-per-node 8 --master-port 29604 bench.py --gpus 8 --workload plate_c5 --no-extra --steps 5 > $O/r2_plate_n8.json 2> $O/r2_plate_n8.err
cut -c1-200 $O/r2_plate_n8.json; tail -2 $O/r2_plate_n8.err
timeout 300 $TR --nproc-per-node 2 --master-port 29605 bench.py --gpus 2 --no-extra --steps 5 > $O/r2_bench_n2.json 2> $O/r2_bench_n2.err
timeout 300 $TR --nproc-per-node 8 --master-port 29606 bench.py --gpus 8 --no-extra --steps 5 > $O/r2_bench_n8.json 2> $O/r2_bench_n8.err
python - <<'PY'
import json
for n in (2, 8):
    try:
        d = json.load(open(f"gpurun_out/r2_bench_n{n}.json"))
        print(n, d["value"], d["e2e"]["value"], d["e2e"]["ceiling"], d.get("device"))
    except Exception as e:
        print(n, "ERR", e)
PY
