#!/bin/bash
# First GPU call of the next round, one B200 (about 3 minutes): runs everything that was written after round 1's GPU budget was spent.
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash scripts/round2_first_call.sh'
# Outputs land in gpurun_out/r02_first_*.
out=gpurun_out
mkdir -p $out
# 1. device tests that have never run (wild inputs: under a timeout, in their own process)
U3D_RUN_PENDING=1 timeout 300 python -m pytest tests/pending_gpu -m gpu -q -x > $out/r02_first_pending.log 2>&1; echo "pending rc=$?"
# 2. GPUCompute.h's in-engine classes against the reference's own classes, in one process
U3D_EXPECT_DEVICE=1 timeout 120 tests/cpp/test_engine_types.bin > $out/r02_first_engine_types.log 2>&1; echo "engine types rc=$?"
# 3. the default bench line (new fields: parity_check, cpu_baseline.engine_threads)
timeout 400 python bench.py > $out/r02_first_bench_n1.json 2> $out/r02_first_bench_n1.err; echo "bench rc=$?"
# 4. fill phase with the bulk asynchronous tile store (TMA) against the default build: parity first, then the phase times
make -s -j8 -C urho3d_b200/csrc OUT=$PWD/urho3d_b200/lib_vbulk EXTRA=-DU3D_FILL_BULK_STORE=1 > $out/r02_first_bulk_build.log 2>&1 && {
  U3D_LIB=$PWD/urho3d_b200/lib_vbulk/libu3dgpu.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "batched_views or golden or million or full_size" > $out/r02_first_bulk_tests.log 2>&1; echo "bulk tests rc=$?"
  for lib in lib lib_vbulk; do
    U3D_LIB=$PWD/urho3d_b200/$lib/libu3dgpu.so timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); p = d['phases_rank0']
print('$lib', round(d['value']), 'views/s', ' '.join('%s %.3f' % (k, v['ms']) for k, v in p.items() if v['ms'] > 0.05))" | tee -a $out/r02_first_bulk_ab.txt
  done
}
tail -3 $out/r02_first_pending.log $out/r02_first_engine_types.log
