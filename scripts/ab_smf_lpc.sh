# chains per warp of the chain kernel (LITE_SMF_LPC = lanes per chain: 1 -> 32 chains per warp, 2 -> 16, 4 -> 8) at 8192 chains
mkdir -p gpurun_out/s8
for lpc in ${LPCS:-2 1 4 2}; do
  LITE_SMF_LPC=$lpc python bench.py --workload c5 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/s8/c5_lpc_$lpc.json 2>/dev/null
  python - <<P
import json
d=json.loads(open('gpurun_out/s8/c5_lpc_$lpc.json').read().strip().splitlines()[-1])
print('lanes per chain', $lpc, 'chain-steps/s %.4g' % d['value'], 'ms/step %.3f' % d['ms_per_step'])
P
done
