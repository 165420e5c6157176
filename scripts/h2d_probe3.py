"""Copy time of a 3.8 MB host->device transfer from differently allocated pinned buffers."""
import ctypes, numpy as np, torch
dev = torch.device("cuda", 0)
torch.cuda.init()
rt = ctypes.CDLL("libcudart.so.12")
n = 3_805_184
d = torch.empty(8 * n, dtype=torch.uint8, device=dev)
s = torch.cuda.Stream()
def timeit(ptr, label, nbytes=n):
    ts = []
    with torch.cuda.stream(s):
        for k in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            rt.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(nbytes), 1, ctypes.c_void_p(s.cuda_stream))
            e1.record(s); s.synchronize()
            ts.append(e0.elapsed_time(e1))
    print(label, ["%.3f" % t for t in ts])
for cap in (n, 5_900_000, 16 << 20, 64 << 20):
    p = ctypes.c_void_p()
    assert rt.cudaMallocHost(ctypes.byref(p), ctypes.c_size_t(cap)) == 0
    ctypes.memset(p, 1, cap)
    timeit(p.value, "cudaMallocHost cap %9d, copy 3.8 MB" % cap)
    if cap >= 16 << 20:
        timeit(p.value, "   same, copy whole %9d" % cap, cap)
    p2 = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p2), ctypes.c_size_t(cap), 4) == 0   # cudaHostAllocWriteCombined
    ctypes.memset(p2, 1, cap)
    timeit(p2.value, "write-combined cap %9d, copy 3.8 MB" % cap)
