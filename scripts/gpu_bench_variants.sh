#!/bin/bash
# usage (on the GPU box): scripts/gpu_bench_variants.sh [bench args] -- name1 name2 ...
# runs bench.py once per development library simpleds_b200/variants/libsds_<name>.so ("default" = libsds.so)
args=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do args+=("$1"); shift; done
shift
for v in "$@"; do
  if [ "$v" = default ]; then unset SDS_LIB; else export SDS_LIB=$PWD/simpleds_b200/variants/libsds_$v.so; fi
  python bench.py --no-e2e --no-cpu "${args[@]}" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$v: ms/step %.3f  frac %.4f  GB/s %.1f  value %.3e' % (d['ms_per_step'], d['roofline']['frac'], d['roofline']['achieved'], d['value']))
    elif 'rror' in l: print(l.strip())
"
done
