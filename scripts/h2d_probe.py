"""Host->device copy time against size from pinned memory (what one lite_tess_upload pays):
python scripts/h2d_probe.py"""
import torch
dev = torch.device("cuda", 0)
for mb in (0.25, 0.5, 1, 2, 4, 8, 16, 64, 256):
    n = int(mb * 1024 * 1024)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            d.copy_(h, non_blocking=True)
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record(s)
        for _ in range(reps):
            d.copy_(h, non_blocking=True)
        e1.record(s)
        s.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%8.2f MB  %8.3f ms  %7.2f GB/s" % (mb, ms, n / ms / 1e6))
