#!/bin/bash
# Diagnostic sweep for the multi-GPU step overhead (run on an N-GPU box): bash scripts/scale_diag.sh N
# N > 1: slots in flight, peer-memory exchange vs the NCCL baseline vs no exchange at all; N = 1: what a rank's share of the batch costs alone.
N=${1:-2}
out=gpurun_out
mkdir -p $out
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
  run n${N}_default A=1 --
  run n${N}_interleaved A=1 -- --shard interleaved
  run n${N}_if12 A=1 -- --inflight 12
  run n${N}_if4 A=1 -- --inflight 4
  run n${N}_if1 A=1 -- --inflight 1
  run n${N}_nccl A=1 -- --gather nccl
  run n${N}_nogather U3D_BENCH_NO_GATHER=1 --
  run n${N}_nccl_ctas2 NCCL_MAX_CTAS=2 -- --gather nccl
else
  run v4096_if1 A=1 -- --inflight 1
  run v4096_if8 A=1 -- --inflight 8
  for v in 512 1024 2048; do
    run v${v}_if1 A=1 -- --views $v --inflight 1
    run v${v}_if4 A=1 -- --views $v --inflight 4
    run v${v}_if12 A=1 -- --views $v --inflight 12
  done
fi
