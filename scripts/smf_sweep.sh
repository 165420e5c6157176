python -m pytest tests/test_gpu_smf.py -x -q -m gpu 2>&1 | tail -3
for minb in 4 8 16; do for lpc in 1 2 4 8 32; do
echo "== n=8192 MINB=$minb LPC=$lpc"; LITE_SMF_MINB=$minb LITE_SMF_LPC=$lpc python scripts/smf_probe.py 8192 1020 1500 300 2>&1 | tail -1
done; done
for minb in 8 16; do for lpc in 1 2; do
echo "== n=65536 MINB=$minb LPC=$lpc"; LITE_SMF_MINB=$minb LITE_SMF_LPC=$lpc python scripts/smf_probe.py 65536 1020 1500 300 2>&1 | tail -1
done; done
