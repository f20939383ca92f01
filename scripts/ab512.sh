#!/bin/bash
# like ab.sh, at 512 views per step with 12 slots in flight (one rank's share of the 8-GPU run)
for o in "$@"; do
  opts=$o; [ "$o" = "-" ] && opts=""
  U3D_DEBUG_OPTS=$opts timeout 300 python bench.py --views 512 --inflight 12 --steps 40 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][0]); p = d['phases_rank0']
print('%-40s' % '$o', round(d['value']), 'views/s e2e', round(d['e2e']['value']), ' '.join('%s %.3f' % (k, v['ms']) for k, v in p.items() if v['ms'] > 0.02 and 'total' not in k))"
done
