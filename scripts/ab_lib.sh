#!/bin/bash
# A/B of library builds on the default bench.  usage: scripts/ab_lib.sh lib lib_vX ...   (directories under urho3d_b200/)
for lib in "$@"; do
  U3D_LIB=$PWD/urho3d_b200/$lib/libu3dgpu.so timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); p = d['phases_rank0']
print('%-16s' % '$lib', round(d['value']), 'views/s e2e', round(d['e2e']['value']), ' '.join('%s %.3f' % (k, v['ms']) for k, v in p.items() if v['ms'] > 0.05 and 'total' not in k))"
done
