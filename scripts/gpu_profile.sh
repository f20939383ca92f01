#!/bin/bash
# Run on the GPU box (under gpurun): plain bench first, then the ncu launch list and one full capture per hot kernel.
# usage: scripts/gpu_profile.sh <tag> [views]
set -u
TAG=${1:-r01}
VIEWS=${2:-592}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --views $VIEWS --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > $OUT/plain_$TAG.json 2> $OUT/plain_$TAG.err || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
for K in ${KERNELS:-k_cull_views k_fill_tiles k_setup_triangles}; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $OUT/prof_${TAG}_$K $CMD > $OUT/ncu_${TAG}_$K.log 2>&1
  echo "$K rc=$?"
done
ls -la $OUT | tail -12
