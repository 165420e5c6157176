"""Copy time of a pinned block that several threads have just written."""
import numpy as np, torch, threading
dev = torch.device("cuda", 0)
n = 3_805_184
h = torch.empty(n, dtype=torch.uint8).pin_memory()
src = np.random.randint(0, 255, n, dtype=np.uint8)
d = torch.empty(n, dtype=torch.uint8, device=dev)
s = torch.cuda.Stream()
hn = h.numpy()
def write(nt):
    if nt == 1:
        np.copyto(hn, src); return
    step = n // nt
    th = [threading.Thread(target=lambda a=a: np.copyto(hn[a:a + step], src[a:a + step])) for a in range(0, step * nt, step)]
    [t.start() for t in th]; [t.join() for t in th]
for nt in (1, 2, 4, 8, 1, 4):
    ts = []
    with torch.cuda.stream(s):
        for k in range(8):
            write(nt)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); d.copy_(h, non_blocking=True); e1.record(s); s.synchronize()
            ts.append(e0.elapsed_time(e1))
    print("written by %d thread(s):" % nt, ["%.3f" % t for t in ts])
