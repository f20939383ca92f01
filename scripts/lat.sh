#!/bin/bash
# tests/config_latency.py for the given configs under each option set.   usage: scripts/lat.sh "C2 C5" - tile_target=444 ...
cfgs=$1; shift
for o in "$@"; do
  opts=$o; [ "$o" = "-" ] && opts=""
  U3D_DEBUG_OPTS=$opts timeout 300 python tests/config_latency.py $cfgs 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); g = d['gpu_ms']
    print('%-22s' % '$o', d['config'], 'total %.3f call %.3f ' % (g['total'], d['call_ms']), ' '.join('%s %.3f' % (k, v) for k, v in g.items() if v > 0.02 and k != 'total'))"
done
