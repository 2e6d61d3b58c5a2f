"""Developer probe: pinned host<->device copy bandwidth, one direction and both at once (the bound of the e2e leg)."""
import time
import torch
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5, chunk=None):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if chunk is None:
            if h2d:
                with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
        else:
            for o in range(0, n, chunk):
                if h2d:
                    with torch.cuda.stream(s1): d_a[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2): h_out[o:o + chunk].copy_(d_b[o:o + chunk], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return reps * n * (int(h2d) + int(d2h)) / dt / 1e9
for _ in range(2): run(True, True, 1)
print("h2d GB/s", round(run(True, False), 1))
print("d2h GB/s", round(run(False, True), 1))
print("both GB/s (sum)", round(run(True, True), 1))
print("both, 32 MB chunks", round(run(True, True, chunk=32 << 20), 1))
print("both, 4 MB chunks", round(run(True, True, chunk=4 << 20), 1))
