"""Create / use / destroy contexts repeatedly (thread pool, streams, pinned buffers, ensemble)."""
import sys, math
sys.path.insert(0, ".")
import numpy as np
import lite_b200
t, _ = lite_b200.generate_crtt(2000, seed=5)
soa = t.export()
for k in range(40):
    ctx = lite_b200.Context(0)
    ctx.tess_upload(soa); ctx.energy_set(["minus_is_segment_internal"])
    ctx.merges_flips(); ctx.add_splits_philox(50000, k, 0)
    lpl, g, H = ctx.pl_eval([0.3])
    ctx.smf_create(128, 100, seed=k)
    ctx.smf_set_model(["is_segment_internal"], [-math.log(3.0)])
    ctx.smf_run(1, 200, rows=False)
    ctx.close()
print("churn ok", lpl)
