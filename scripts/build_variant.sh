#!/bin/bash
# usage: scripts/build_variant.sh <name> <source.cu> [nvcc flags...]
# builds biahub_b200/_lib/variants/libb2_<name>.so = the current library with one source file
# recompiled with extra flags (A/B timing with BIAHUB_B200_LIB=...)
cd /root/repo
name=$1; src=$2; shift 2
mkdir -p biahub_b200/_lib/variants
python -c "from biahub_b200 import _build; _build.build()"
base=$(basename "$src" .cu)
objs=$(ls biahub_b200/_lib/obj/*.o | grep -v "/$base.o")
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -I include -c "$src" -o /tmp/var_$name.o || exit 1
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o biahub_b200/_lib/variants/libb2_$name.so $objs /tmp/var_$name.o -cudart static || exit 1
ls -la biahub_b200/_lib/variants/libb2_$name.so
