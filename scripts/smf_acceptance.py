import sys; sys.path.insert(0,'.')
import numpy as np, math, lite_b200
ctx=lite_b200.Context(0)
ctx.smf_create(8192,1020,seed=5)
ctx.smf_set_model(["is_segment_internal","edge_length"],[-math.log(6.0),5/math.pi],0.33,0.33)
ctx.smf_run(1,20000,rows=False)
r=ctx.smf_run(4,250)
col={n:i for i,n in enumerate(lite_b200.SMF_COLS)}
m=r.mean(axis=(0,2))
print({k:float(m[col[k]]) for k in lite_b200.SMF_COLS})
print("acc S %.3f M %.3f F %.3f overall %.3f"%(m[col['acc_S']]/m[col['prop_S']], m[col['acc_M']]/m[col['prop_M']], m[col['acc_F']]/m[col['prop_F']], (m[col['acc_S']]+m[col['acc_M']]+m[col['acc_F']])/250))
