# A/B of the prefetch modes of the chain step (LITE_SMF_PREFETCH bit mask) on config 5
mkdir -p gpurun_out/s8
for chains in 8192 65536; do
for mode in ${MODES:-0 1 2 4 3 5 7 0}; do
  LITE_SMF_PREFETCH=$mode python bench.py --workload c5 --chains $chains --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/s8/c5_${chains}_$mode.json 2>/dev/null
  python - <<P
import json
d=json.loads(open('gpurun_out/s8/c5_${chains}_$mode.json').read().strip().splitlines()[-1])
print('chains', $chains, 'prefetch', $mode, 'chain-steps/s %.4g' % d['value'], 'ms/step %.3f' % d['ms_per_step'])
P
done
done
