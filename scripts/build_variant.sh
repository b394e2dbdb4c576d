#!/bin/bash
# usage: scripts/build_variant.sh NAME "<extra nvcc flags>"
# Rebuilds only the fp32 engine with the extra flags and links a development copy of the
# library at simpleds_b200/variants/libsds_NAME.so (select it with SDS_LIB=<path>).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p simpleds_b200/variants
python -c "from simpleds_b200.build import build_lib; build_lib()"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 $@ -Xcompiler -fPIC \
  -c simpleds_b200/csrc/sds_engine_f32.cu -o simpleds_b200/variants/engine_f32_$name.o
objs=$(ls simpleds_b200/csrc/*.o | grep -v sds_engine_f32.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o simpleds_b200/variants/libsds_$name.so $objs simpleds_b200/variants/engine_f32_$name.o
echo simpleds_b200/variants/libsds_$name.so
