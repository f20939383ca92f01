for inf in 12 4 1; do
timeout 300 python bench.py --views 512 --inflight $inf --steps 40 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][0]); p = d['phases_rank0']
print('inflight $inf', round(d['value']), 'views/s e2e', round(d['e2e']['value']), ' '.join('%s %.3f' % (k, v['ms']) for k, v in p.items() if 'total' not in k))"
done
