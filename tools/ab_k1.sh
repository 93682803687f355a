#!/bin/bash
# A/B of library variants (tools/build_variant.py): K1 on c2 / c2s / c2d / c4 and the planner batch
for v in "$@"; do
  [ "$v" = base ] && v=""
  VHP_LIB_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-legs --no-e2e --no-cpu --no-giant 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['planner']
print('[$v]', 'c2', round(d['value']), {k:(round(v['value']),round(v['frac_of_hbm_peak'],3)) for k,v in d['penumbra'].items()}, 'planner ms', round(p['device']['ms_per_batch'],2), round(p['roofline']['frac'],3))"
done
