"""Developer probe: ONE BlockGf (2 blocks, the literal configs[1]) per call, device resident: event time of the round trip against the sum of
its kernels (what is left is launch gaps and the status read-back that makes every call synchronous)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import triqs_b200
F = triqs_b200.FERMION
ctx = triqs_b200.Context(0, print_warnings=False)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
g = torch.Generator(device="cuda").manual_seed(0)
gt = torch.randn(2, 100001, 25, dtype=torch.float64, device="cuda", generator=g).to(torch.complex128) * 1e-3
gw = torch.empty(2, 20000, 25, dtype=torch.complex128, device="cuda")
def step():
    ctx.fourier_tau_to_iw(gt, 100.0, F, 10000, out=gw)
    ctx.fourier_iw_to_tau(gw, 100.0, F, 100001, out=gt)
for _ in range(5): step()
ctx.synchronize()
t0 = time.perf_counter()
for _ in range(50): step()
ctx.synchronize()
wall = (time.perf_counter() - t0) / 50 * 1e3
ctx.profile_enable(True)
for _ in range(20): step()
ctx.synchronize()
prof = ctx.profile_report()
ks = {k: round(v["ms"] / 20, 4) for k, v in prof.items()}
print({"wall_ms_per_round_trip": round(wall, 4), "kernel_sum_ms": round(sum(ks.values()), 4), "kernels": ks})
