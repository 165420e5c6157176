"""Does a host->device copy slow down when the pinned source was just written by the CPU?"""
import numpy as np, torch, time
dev = torch.device("cuda", 0)
n = 3_805_184
h = torch.empty(n, dtype=torch.uint8).pin_memory()
src = np.random.randint(0, 255, n, dtype=np.uint8)
d = torch.empty(n, dtype=torch.uint8, device=dev)
s = torch.cuda.Stream()
hn = h.numpy()
def run(write, label, gap=0.0):
    ts = []
    with torch.cuda.stream(s):
        for k in range(8):
            if write:
                np.copyto(hn, src)
            if gap:
                time.sleep(gap)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); d.copy_(h, non_blocking=True); e1.record(s); s.synchronize()
            ts.append(e0.elapsed_time(e1))
    print(label, ["%.3f" % t for t in ts])
run(False, "clean source   ")
run(True, "just written   ")
run(True, "written, 2ms gap", 0.002)
run(False, "idle 2 ms before", 0.002)
run(False, "idle 20 ms before", 0.02)
