"""Kernels added in the second half of round 2, for an ncu capture: the Bluestein convolution (L = 19999 and 9999 x 25: pre / post kernels, the
planned lattice kernel with and without the fused product), the real-valued boundary kernels (promote / real part), the heavy F1 kernel
with 7 workers (L = 61500) and a "gfs as columns" batch of scalar gfs (L = 20000)."""
import sys

import torch

sys.path.insert(0, ".")
import triqs_b200  # noqa: E402

F = triqs_b200.FERMION
ctx = triqs_b200.Context(0, print_warnings=False)
dev = torch.device("cuda", 0)


def matsubara(L, n_iw, nc, gbytes=0.4e9, real=False):
    batch = max(1, int(gbytes / ((L + 1) * nc * 16)))
    gt = torch.view_as_complex(torch.randn(batch, L + 1, nc, 2, dtype=torch.float64, device=dev)).mul_(1e-3)
    gw = torch.empty(batch, 2 * n_iw, nc, dtype=torch.complex128, device=dev)
    gt2 = torch.empty_like(gt)
    mom = torch.zeros(batch, 4, nc, dtype=torch.complex128, device=dev)
    mom[:, 1] = 1.0
    gr = gt.real.contiguous() if real else None
    gr2 = torch.empty_like(gr) if real else None
    for _ in range(2):
        ctx.fourier_tau_to_iw(gr if real else gt, 20.0, F, n_iw, moments=mom, out=gw)
        ctx.fourier_iw_to_tau(gw, 20.0, F, L + 1, moments=mom, out=gr2 if real else gt2)
    ctx.synchronize()


matsubara(19999, 3333, 25)
matsubara(9999, 1000, 25)
matsubara(100000, 10000, 25, real=True)
matsubara(61500, 10250, 25)
matsubara(20000, 3333, 1)
ctx.close()
print("done")
