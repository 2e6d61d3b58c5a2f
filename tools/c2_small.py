"""Developer probe: one small C2-shaped forward + inverse transform (2 blocks) for ncu captures."""
import sys
import torch
sys.path.insert(0, ".")
import triqs_b200
F = triqs_b200.FERMION
ctx = triqs_b200.Context(0, print_warnings=False)
g = torch.Generator(device="cuda").manual_seed(0)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1
gt = torch.randn(2 * R, 100001, 25, dtype=torch.float64, device="cuda", generator=g).to(torch.complex128) * 1e-3
gw = torch.empty(2 * R, 20000, 25, dtype=torch.complex128, device="cuda")
mom = torch.zeros(2 * R, 4, 25, dtype=torch.complex128, device="cuda"); mom[:, 1] = 1.0
for _ in range(2):
    ctx.fourier_tau_to_iw(gt, 100.0, F, 10000, moments=mom, out=gw)
    ctx.fourier_iw_to_tau(gw, 100.0, F, 100001, moments=mom, out=gt)
ctx.synchronize()
print("ok")
