#!/bin/bash
# One frame's kernels under ncu: (1) launch list with durations, (2) --set full with source for the 11 kernels of
# one warm frame. Run on the GPU box AFTER `python bench.py` exited 0 without ncu. Usage: tools/profile_frame.sh TAG
set -u
TAG=${1:-rX}
OUT=gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-modes"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_ncu_launch.log 2>&1
# a warm frame: skip the first 3 frames (9 kernels each since round 2: project, compaction, depth bucket scatter + local sort,
# expansion, one tile-sort pass, tile chunk count + place, blend)
KPF=${KPF:-10}
SKIP=$((3 * KPF))
N=$(grep -c "project_cull_mono_kernel" $OUT/${TAG}_launches.csv)
ncu --set full --import-source on --clock-control none -k regex:"project_cull|compact_visible|onesweep_pass|onesweep_pair|bucket_rank|bucket_scatter|bucket_local_sort|create_instances|tile_chunk|tile_lower_bounds|blend_mono" \
    --launch-skip $SKIP -c $KPF -o $OUT/${TAG}_frame $BENCH > $OUT/${TAG}_ncu_full.log 2>&1
echo "launch list rows: $N, skipped $SKIP"
ls -la $OUT/${TAG}_frame.ncu-rep
