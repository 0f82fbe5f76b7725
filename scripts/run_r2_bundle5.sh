#!/bin/bash
O=gpurun_out
python scripts/rot90_probe.py > $O/r2_rot90_probe.txt 2>&1; cat $O/r2_rot90_probe.txt
python scripts/spline_probe.py > $O/r2_spline_probe.txt 2>&1; cat $O/r2_spline_probe.txt
