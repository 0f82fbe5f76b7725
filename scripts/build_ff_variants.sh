#!/bin/bash
# builds variant libraries of the flat-field kernels for A/B timing: _lib/variants/libb2_<name>.so
cd /root/repo
mkdir -p biahub_b200/_lib/variants
python -c "from biahub_b200 import _build; _build.build()"
objs=$(ls biahub_b200/_lib/obj/*.o | grep -v b2_flatfield.o)
build() { # name flags...
  name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -I include -c biahub_b200/csrc/b2_flatfield.cu -o /tmp/ffv_$name.o || exit 1
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o biahub_b200/_lib/variants/libb2_$name.so $objs /tmp/ffv_$name.o -cudart static || exit 1
}
build t32 -DB2_FM_THREADS=32 &
build g4 -DB2_FM_GROUP=4 &
build r16g4 -DB2_FM_RING=16 -DB2_FM_GROUP=4 &
build r32g16 -DB2_FM_RING=32 -DB2_FM_GROUP=16 &
wait
ls -la biahub_b200/_lib/variants/
