#!/bin/bash
# A/B over the warps-per-CTA choice of the batched tile sweep (VHP_TILE_WARPS): c2, c2s kernel-only
for w in "$@"; do
  for wl in c2 c2s; do
    VHP_TILE_WARPS=$w python bench.py --workload $wl --no-e2e --no-cpu --no-planner --steps 10 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('warps=$w','$wl','%.1f Gcells/s  %.3f ms  frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['frac']))"
  done
done
