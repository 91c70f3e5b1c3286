"""Development probe: D2H bandwidth of a 110 MB pinned copy alone, and while a 34 MB H2D copy runs on another stream."""
import time
import torch
d = torch.empty(110 << 20, dtype=torch.uint8, device="cuda"); h = torch.empty(110 << 20, dtype=torch.uint8).pin_memory()
d2 = torch.empty(34 << 20, dtype=torch.uint8, device="cuda"); h2 = torch.empty(34 << 20, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for both in (False, True):
    for _ in range(3):
        with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10):
        with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                for _ in range(3): d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
    print("D2H 110 MiB", "with concurrent 3x34 MiB H2D" if both else "alone", f"{dt*1e3:.3f} ms", (110 << 20) / dt / 1e9, "GB/s")
# chunked D2H: 9 copies of 12 MiB back to back on one stream
parts = [(i * (12 << 20), 12 << 20) for i in range(9)]
for _ in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    with torch.cuda.stream(s1):
        for o, n in parts: h[o:o + n].copy_(d[o:o + n], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
print("9 x 12 MiB D2H back to back", f"{dt*1e3:.3f} ms", 9 * (12 << 20) / dt / 1e9, "GB/s")
