# running block sums (LITE_SMF_PREFETCH bit 3) against the exact re-sums, 8192 chains; then the chain tests in both modes
mkdir -p gpurun_out/s8
for mode in ${MODES:-7 15 7 15}; do
  LITE_SMF_PREFETCH=$mode python bench.py --workload c5 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/s8/c5_sums_$mode.json 2>/dev/null
  python - <<P
import json
d=json.loads(open('gpurun_out/s8/c5_sums_$mode.json').read().strip().splitlines()[-1])
print('mode', $mode, 'chain-steps/s %.4g' % d['value'], 'ms/step %.3f' % d['ms_per_step'])
P
done
LITE_SMF_PREFETCH=15 python -m pytest tests/test_gpu_smf.py -m gpu -x -q 2>&1 | tail -2
