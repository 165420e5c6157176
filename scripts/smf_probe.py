"""Throughput probe of the SMF chain ensemble (chain-steps/s) on one GPU.
usage: python scripts/smf_probe.py [n_chains] [max_segments] [burn_in] [steps] [tau]"""
import math
import sys
import time

sys.path.insert(0, ".")
import lite_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 1020
burn = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
tau = float(sys.argv[5]) if len(sys.argv) > 5 else 6.0
ctx = lite_b200.Context(0)
t0 = time.time()
ctx.smf_create(n, cap, seed=5)
print("create %.3fs  bytes/chain %d" % (time.time() - t0, ctx.smf_count()[3]))
ctx.smf_set_model(["is_segment_internal", "edge_length"], [-math.log(tau), (tau - 1) / math.pi])
done = 0
for chunk in ([burn] if burn else []) + [steps] * 3:
    rows = ctx.smf_run(1, chunk)
    ms = ctx.smf_last_ms()
    done += chunk
    print("after %6d steps: mean n(T) %.1f  max %d  | %d steps x %d chains in %.1f ms = %.3e chain-steps/s" % (
        done, rows[0, 1].mean(), rows[0, 1].max(), chunk, n, ms, chunk * n / (ms * 1e-3)), flush=True)
ctx.close()
