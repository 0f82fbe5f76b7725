#!/bin/bash
# Rebuild the library with -DB2_BOUNDS_CHECK on the GPU box and run the per-family cases: the
# substitute for compute-sanitizer, which is closed on this pool.
O=gpurun_out
cp biahub_b200/_lib/libbiahub_b200.so /tmp/release.so
B2_NVCC_EXTRA=-DB2_BOUNDS_CHECK python -m biahub_b200._build --force > $O/r2_bc_build.log 2>&1
python scripts/sanitize_cases.py > $O/r2_bounds_check.log 2>&1; echo "rc=$?" >> $O/r2_bounds_check.log
cp /tmp/release.so biahub_b200/_lib/libbiahub_b200.so
tail -4 $O/r2_bc_build.log; grep -E "sanitize_cases|rc=|Error|assert" $O/r2_bounds_check.log | head
