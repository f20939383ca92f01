# slots in flight at the default 4096-view step
for f in ${@:-4 6 8 12}; do
  timeout 300 python bench.py --inflight $f --steps 24 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][0])
print('inflight $f', round(d['value']), 'views/s e2e', round(d['e2e']['value']))"
done
