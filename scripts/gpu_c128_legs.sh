#!/bin/bash
# usage: c128legs.sh variant...
for v in "$@"; do
  if [ "$v" = default ]; then unset SDS_LIB; else export SDS_LIB=$PWD/simpleds_b200/variants/libsds_$v.so; fi
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:engine_ -c 4 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-breakdown --no-modes --no-fixed-job --precision complex128 > gpurun_out/ncu_c128_$v.log 2>&1
  echo "== $v"; grep -A4 "void engine" gpurun_out/ncu_c128_$v.log | grep "void\|gpu__time" | cut -c1-60 | paste - - | awk '{print $2,$3,$4,$NF}'
done
