"""PCIe probe for the e2e analysis of bench.py: pinned H2D of 133 MB, D2H of 113 MB, alone and concurrently (two streams)."""
import time, torch
dev = torch.device("cuda", 0)
hin = torch.empty(133_000_000 // 8, dtype=torch.float64).pin_memory()
hout = torch.empty(113_000_000 // 8, dtype=torch.float64).pin_memory()
din = torch.empty_like(hin, device=dev); dout = torch.empty_like(hout, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def run(h2d, d2h, reps=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
for _ in range(2): run(True, True, 2)
a, b, c = run(True, False), run(False, True), run(True, True)
print(f"H2D 133 MB alone {a:.2f} ms ({133/a:.1f} GB/s)   D2H 113 MB alone {b:.2f} ms ({113/b:.1f} GB/s)   both concurrently {c:.2f} ms")
