"""Planned pipelined kernels at L = 10^4 x 25 columns (and forced at L = 10^5) for an ncu capture."""
import sys

import torch

sys.path.insert(0, ".")
import triqs_b200  # noqa: E402

F = triqs_b200.FERMION
ctx = triqs_b200.Context(0, print_warnings=False)
dev = torch.device("cuda", 0)
L, n_iw, nc = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (10000, 1025, 25)
batch = max(1, int(0.4e9 / ((L + 1) * nc * 16)))
gt = torch.view_as_complex(torch.randn(batch, L + 1, nc, 2, dtype=torch.float64, device=dev)).mul_(1e-3)
gw = torch.empty(batch, 2 * n_iw, nc, dtype=torch.complex128, device=dev)
gt2 = torch.empty_like(gt)
mom = torch.zeros(batch, 4, nc, dtype=torch.complex128, device=dev)
mom[:, 1] = 1.0
if L == 100000:
    ctx.set_option("force_planned", 1)
for _ in range(2):
    ctx.fourier_tau_to_iw(gt, 20.0, F, n_iw, moments=mom, out=gw)
    ctx.fourier_iw_to_tau(gw, 20.0, F, L + 1, moments=mom, out=gt2)
ctx.synchronize()
ctx.close()
print("done")
