"""The whole config 5 on one GPU: 65536 chains x 100 batches x 250 proposals (checkACS,
tau = 6), then the consistency check of every chain and the detailed-balance identity."""
import math, sys, time
sys.path.insert(0, ".")
import numpy as np
import lite_b200
from lite_b200.capi import SMF_COLS
COL = {n: i for i, n in enumerate(SMF_COLS)}
n, tau = 65536, 6.0
ctx = lite_b200.Context(0)
ctx.smf_create(n, 1020, seed=5)
ctx.smf_set_model(["is_segment_internal", "edge_length"], [-math.log(tau), (tau - 1) / math.pi])
t0 = time.time()
rows = ctx.smf_run(100, 250)
dt = time.time() - t0
print("100 x 250 proposals on %d chains: %.2f s wall, kernel %.2f s, %.3e chain-steps/s" % (
    n, dt, ctx.smf_last_ms() / 1e3, n * 25000 / (ctx.smf_last_ms() / 1e3)))
ctx.smf_stats(validate=True)
tail = rows[80:]
est = tail[:, COL["n_nbl"]].mean() * math.pi / (2 * tail[:, COL["l"]] + 4).mean()
print("mean n(T) %.1f  max %d  n_nb*pi/int_length = %.4f (tau = %g)" % (
    rows[-1, COL["n"]].mean(), rows[:, COL["n"]].max(), est, tau))
acc = rows[:, [COL["acc_S"], COL["acc_M"], COL["acc_F"]]].sum(axis=(0, 2)) / rows[:, [COL["prop_S"], COL["prop_M"], COL["prop_F"]]].sum(axis=(0, 2))
print("acceptance S/M/F", acc)
ctx.close()
