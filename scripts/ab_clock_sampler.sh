mkdir -p gpurun_out/s7
for rep in 1 2; do
for per in 0.002 5.0; do
  LITE_BENCH_CLOCK_PERIOD=$per python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-secondary > gpurun_out/s7/ab_${per}_$rep.json 2>/dev/null
  python - <<P
import json
d=json.loads(open('gpurun_out/s7/ab_${per}_$rep.json').read().strip().splitlines()[-1])
print('period', '$per', 'rep', $rep, 'value ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'samples', d['clocks']['samples'], d['clocks'].get('e2e_region',{}).get('samples'))
P
done
done
