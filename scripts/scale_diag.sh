#!/bin/bash
# Diagnostic sweep for the multi-GPU step overhead (run on an N-GPU box): bash scripts/scale_diag.sh N
N=${1:-2}
out=gpurun_out
run() { # name, env..., -- args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  if [ "$N" -gt 1 ]; then
    env "${envs[@]}" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 "$@" > $out/diag_$name.json 2> $out/diag_$name.err
  else
    env "${envs[@]}" timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" > $out/diag_$name.json 2> $out/diag_$name.err
  fi
  echo "$name rc=$?"
}
if [ "$N" -gt 1 ]; then
  run n${N}_if2 A=1 --
  run n${N}_if1 A=1 -- --inflight 1
  run n${N}_if3 A=1 -- --inflight 3
  run n${N}_if2_nogather U3D_BENCH_NO_GATHER=1 --
  run n${N}_if2_ctas2 NCCL_MAX_CTAS=2 --
  run n${N}_if2_sb0 U3D_CULL_SMALL_BATCH=0 --
else
  run v4096_if1 A=1 -- --inflight 1
  run v4096_if2 A=1 -- --inflight 2
  run v512_if1 A=1 -- --views 512 --inflight 1
  run v512_if2 A=1 -- --views 512 --inflight 2
  run v512_if3 A=1 -- --views 512 --inflight 3
  run v512_if2_sb0 U3D_CULL_SMALL_BATCH=0 -- --views 512 --inflight 2
  run v512_if3_sb0 U3D_CULL_SMALL_BATCH=0 -- --views 512 --inflight 3
  run v1024_if2 A=1 -- --views 1024 --inflight 2
  run v1024_if2_sb2000 U3D_CULL_SMALL_BATCH=2000 -- --views 1024 --inflight 2
fi
