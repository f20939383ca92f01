#!/bin/bash
# Launch list of the DEFAULT bench command (4096 views) + full captures of the hot kernels at 1184 views (same launch shapes as the
# default command: the cull kernel switches to its small-batch shape at <= 640 views).  Run on the GPU box under gpurun.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > $OUT/plain_full_$TAG.json 2> $OUT/plain_full_$TAG.err || { echo "plain run failed"; tail -5 $OUT/plain_full_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_full_$TAG.csv $CMD > $OUT/ncu_launches_full_$TAG.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --views 1184 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD2 > $OUT/plain_1184_$TAG.json 2> $OUT/plain_1184_$TAG.err || { echo "plain run failed"; exit 1; }
for K in ${KERNELS:-k_cull_views k_fill_tiles k_setup_triangles k_clip_triangles}; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $OUT/prof_${TAG}_$K $CMD2 > $OUT/ncu_${TAG}_$K.log 2>&1
  echo "$K rc=$?"
done
