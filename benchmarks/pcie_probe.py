"""D2H copy rate of this box (pinned host memory), one stream and two concurrent streams:
the ceiling of the end-to-end range query, whose rows are 12 B per hit."""
import time
import torch

n = 2_790_000_000
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for label, parts in (("one stream", 1), ("two streams", 2), ("one stream, 8 pieces", -8)):
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if parts == 1:
            with torch.cuda.stream(s1):
                h.copy_(d, non_blocking=True)
        elif parts == 2:
            m = n // 2
            with torch.cuda.stream(s1):
                h[:m].copy_(d[:m], non_blocking=True)
            with torch.cuda.stream(s2):
                h[m:].copy_(d[m:], non_blocking=True)
        else:
            m = n // 8
            with torch.cuda.stream(s1):
                for k in range(8):
                    h[k * m:(k + 1) * m].copy_(d[k * m:(k + 1) * m], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print("%-22s %.2f ms  %.1f GB/s" % (label, dt * 1e3, n / dt / 1e9), flush=True)
