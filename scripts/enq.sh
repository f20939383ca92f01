# host enqueue time per step against the device time per step, at 4096 and at 512 views
for a in "--views 4096 --inflight 8 --steps 20" "--views 512 --inflight 12 --steps 40"; do
  timeout 300 python bench.py $a --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][0])
print('$a', round(d['value']), 'views/s  device ms/step', round(d['ms_per_step'], 3), ' host enqueue ms/step', d['host_enqueue_ms_per_step'], ' launches/step', d['gpu_launches'] / d['steps'])"
done
