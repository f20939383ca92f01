#!/bin/bash
# bench.py at N GPUs of this box -> gpurun_out/r02_scale_n$N.json     usage: bash scripts/scale_run.sh N [extra bench args]
N=$1; shift
if [ "$N" -gt 1 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 "$@" > gpurun_out/r02_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err
else
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r02_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err
fi
echo "N=$N rc=$?"
