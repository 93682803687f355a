#!/bin/bash
# ncu captures of the round's kernels (run on the GPU box through gpurun; outputs under gpurun_out/):
#   k1_<workload>.ncu-rep   sweep_tile_kernel, --set full, one launch of a subset of the batch
#   bits_<workload>.ncu-rep sweep_tile_kernel writing bits (kFmtBits), --set full
#   planner.ncu-rep         planner_kernel, --set full, 592 problems
#   launches_default.csv    gpu__time_duration of every launch of the default bench command (short)
set -x
O=gpurun_out
for w in c2 c2s c2d; do
  ncu --set full --clock-control none --import-source on -k regex:sweep_tile_kernel -c 1 -f -o $O/k1_$w \
    python bench.py --workload $w --pairs 1184 --steps 1 --warmup 3 --no-giant --no-planner --no-legs --no-e2e --no-cpu > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:sweep_tile_kernel -c 1 -f -o $O/k1_c4 \
  python bench.py --workload c4 --steps 1 --warmup 3 --no-giant --no-planner --no-legs --no-e2e --no-cpu > /dev/null 2>&1
# the bit-writing sweep (vhp_visibility_batch_bin_dev) on the all-lit and the dense-obstacle batch
for w in c2 c2d; do
  ncu --set full --clock-control none --import-source on -k regex:sweep_tile_kernel -c 1 -f -o $O/bits_$w \
    python tools/profile_bits.py $w 1184 > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:planner_kernel -c 1 -f -o $O/planner python tools/profile_planner.py 592 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_default.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu > $O/launches_default_bench.log 2>&1
ls -la $O/*.ncu-rep $O/launches_default.csv
