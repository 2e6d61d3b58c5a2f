"""Developer probe: C1 (scalar gfs, L = 10^4, known moments) transform time, batch 8192."""
import sys
import torch
sys.path.insert(0, ".")
import triqs_b200
F = triqs_b200.FERMION
ctx = triqs_b200.Context(0, print_warnings=False)
g = torch.Generator(device="cuda").manual_seed(0)
B = 8192
gt = torch.randn(B, 10001, 1, dtype=torch.float64, device="cuda", generator=g).to(torch.complex128) * 1e-3
gw = torch.empty(B, 2050, 1, dtype=torch.complex128, device="cuda")
mom = torch.zeros(B, 4, 1, dtype=torch.complex128, device="cuda"); mom[:, 1] = 1.0
def step():
    ctx.fourier_tau_to_iw(gt, 10.0, F, 1025, moments=mom, out=gw)
for _ in range(3): step()
ctx.synchronize()
ctx.profile_enable(True)
for _ in range(5): step()
ctx.synchronize()
prof = ctx.profile_report()
print({k: round(v["ms"] / 5, 4) for k, v in prof.items()})
