#!/bin/bash
# Round-2 profile set (run on the GPU box under gpurun, after the plain command has exited 0): the ncu launch list of the default bench
# command and one --set full capture per hot kernel, from a single-slot run so that the captured launch is one whole 4096-view batch.
set -u
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > $OUT/plain_r02.json 2> $OUT/plain_r02.err || { echo "plain run failed"; tail -5 $OUT/plain_r02.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches_r02.csv $CMD > $OUT/ncu_launches_r02.log 2>&1
echo "launch list rc=$?"
CMD1="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --inflight 1"
for K in k_dense_occlusion_walk k_fill_tiles k_dense_frustum; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 1 -f -o $OUT/prof_r02_$K $CMD1 > $OUT/ncu_r02_$K.log 2>&1
  echo "$K rc=$?"
done
# the octree walk itself (its seed kernel shares the name's prefix)
ncu --set full --clock-control none --import-source on -k 'regex:^k_dense_tree$' -s 6 -c 1 -f -o $OUT/prof_r02_k_dense_tree_main $CMD1 > $OUT/ncu_r02_k_dense_tree.log 2>&1
echo "k_dense_tree rc=$?"
# the first setup launch of the run is pass A of the two-phase drawing (a later index may land on the deferred count-only pass of the statistics read)
ncu --set full --clock-control none --import-source on -k regex:k_setup_triangles -s 0 -c 1 -f -o $OUT/prof_r02_k_setup_triangles $CMD1 > $OUT/ncu_r02_k_setup_triangles.log 2>&1
echo "k_setup_triangles rc=$?"
