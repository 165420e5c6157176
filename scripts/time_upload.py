import sys, time, os
sys.path.insert(0,'/root/repo')
import numpy as np, lite_b200
t,tau=lite_b200.generate_crtt(10000,seed=5); soa=t.export()
ctx=lite_b200.Context(0)
ACS=["edge_length","face_area_2","face_sum_of_angles","is_segment_internal"]
for k in range(4):
    t0=time.perf_counter(); ctx.tess_upload(soa); t1=time.perf_counter(); ctx.energy_set(ACS); t2=time.perf_counter(); ctx.merges_flips(); t3=time.perf_counter()
    ctx.clear_splits(); ctx.add_splits_philox(10**7,1,0); t4=time.perf_counter(); ctx.pl_eval([1,1,1,1]); t5=time.perf_counter()
    print("upload %.3f energy %.3f mf %.3f add %.3f eval %.3f ms"%tuple(1e3*x for x in (t1-t0,t2-t1,t3-t2,t4-t3,t5-t4)))
