for f in 4 8 12 16 24; do
  timeout 300 python bench.py --views 512 --inflight $f --steps 48 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][0])
print('inflight $f', round(d['value']), 'views/s e2e', round(d['e2e']['value']))"
done
