#!/bin/bash
# planner batch (bench.py's 8192 problems) with 8 / 12 / 16 warps per CTA
for w in 8 6; do
  VHP_PLANNER_WARPS=$w python bench.py --steps 5 --warmup 3 --no-giant --no-penumbra --no-legs --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['planner']; print('warps $w', 'host', round(p['value']), 'dev ms', round(p['device']['ms_per_batch'],3), 'frac', round(p['roofline']['frac'],4))"
done
