#!/bin/bash
# A/B of tuning options on the default bench (one line per option set).  usage: scripts/ab.sh "opts1" "opts2" ...   (opts: name=val,name=val; "-" = defaults)
for o in "$@"; do
  opts=$o; [ "$o" = "-" ] && opts=""
  U3D_DEBUG_OPTS=$opts timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); p = d['phases_rank0']
print('%-40s' % '$o', round(d['value']), 'views/s e2e', round(d['e2e']['value']), ' '.join('%s %.3f' % (k, v['ms']) for k, v in p.items() if v['ms'] > 0.05 and 'total' not in k))"
done
