#!/bin/bash
# Tuning sweep on one GPU: bash scripts/sweep.sh "opts1" "opts2" ...  (each a U3D_DEBUG_OPTS string; "-" = defaults).  Prints views/s and the phases.
for o in "$@"; do
  [ "$o" = "-" ] && oo="" || oo="$o"
  U3D_DEBUG_OPTS="$oo" python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${SWEEP_ARGS:-} 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); p=d['phases_rank0']
print('%-28s %8.0f views/s  %.3f ms/step  e2e %8.0f | ' % ('$o', d['value'], d['ms_per_step'], d['e2e']['value']) + ' '.join('%s %.3f' % (k, v['ms']) for k, v in p.items() if v['ms'] > 0.05))"
done
