"""Developer probe: per-kernel times of the C2 round trip (32 blocks, known moments), via the library's event profiling."""
import sys
import torch
sys.path.insert(0, ".")
import triqs_b200
F = triqs_b200.FERMION
ctx = triqs_b200.Context(0, print_warnings=False)
g = torch.Generator(device="cuda").manual_seed(0)
R = 16
gt = torch.randn(2 * R, 100001, 25, dtype=torch.float64, device="cuda", generator=g).to(torch.complex128) * 1e-3
gw = torch.empty(2 * R, 20000, 25, dtype=torch.complex128, device="cuda")
mom = torch.zeros(2 * R, 4, 25, dtype=torch.complex128, device="cuda"); mom[:, 1] = 1.0
def step():
    ctx.fourier_tau_to_iw(gt, 100.0, F, 10000, moments=mom, out=gw)
    ctx.fourier_iw_to_tau(gw, 100.0, F, 100001, moments=mom, out=gt)
for _ in range(3): step()
ctx.synchronize()
ctx.profile_enable(True)
for _ in range(5): step()
ctx.synchronize()
prof = ctx.profile_report()
print({k: round(v["ms"] / 5, 4) for k, v in prof.items()})
