#!/bin/bash
# usage: try_variants.sh "<flags1>" "<flags2>" ...   (run on the GPU box; rebuilds the engine per variant)
for fl in "$@"; do
  echo "=== variant: $fl"
  SDS_NVCC_FLAGS="$fl" python simpleds_b200/build.py --force > /dev/null 2>&1 || { echo build failed; continue; }
  python bench.py --steps 5 --warmup 3 --times 4 --no-e2e --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.3f  frac %.4f  GB/s %.1f  value %.3e' % (d['ms_per_step'], d['roofline']['frac'], d['roofline']['achieved'], d['value']))
"
done
