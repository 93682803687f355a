#!/bin/bash
# A/B: for each variant run c2, c2s, c4 kernel-only
for v in "$@"; do
  for w in c2 c2s c4; do
    VHP_LIB_VARIANT=$v python bench.py --workload $w --no-e2e --no-cpu --no-planner --steps 10 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$v','$w','%.1f Gcells/s  %.3f ms  frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['frac']))"
  done
done
