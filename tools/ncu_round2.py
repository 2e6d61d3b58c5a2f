"""Round-2 kernels for an ncu capture: planned Matsubara kernels (L = 10^4 x 25), heavy (Rader) ones (L = 6150 x 25), the planned lattice
kernel (96 x 96), the new interpolation kernel, the warp-per-matrix inversion (n = 16) and the C2 compile-time kernels."""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import triqs_b200  # noqa: E402

F = triqs_b200.FERMION
ctx = triqs_b200.Context(0, print_warnings=False)
dev = torch.device("cuda", 0)


def matsubara(L, n_iw, nc, gbytes=0.4e9):
    batch = max(1, int(gbytes / ((L + 1) * nc * 16)))
    gt = torch.view_as_complex(torch.randn(batch, L + 1, nc, 2, dtype=torch.float64, device=dev)).mul_(1e-3)
    gw = torch.empty(batch, 2 * n_iw, nc, dtype=torch.complex128, device=dev)
    gt2 = torch.empty_like(gt)
    mom = torch.zeros(batch, 4, nc, dtype=torch.complex128, device=dev)
    mom[:, 1] = 1.0
    for _ in range(2):
        ctx.fourier_tau_to_iw(gt, 20.0, F, n_iw, moments=mom, out=gw)
        ctx.fourier_iw_to_tau(gw, 20.0, F, L + 1, moments=mom, out=gt2)
    ctx.synchronize()


matsubara(10000, 1025, 25)
matsubara(6150, 1025, 25)
matsubara(100000, 10000, 25)
x = torch.view_as_complex(torch.randn(96 * 96, 9216, 2, dtype=torch.float64, device=dev))
y = torch.empty_like(x)
for _ in range(2):
    ctx.fourier_lattice(x, (96, 96, 1), True, out=y)
del x, y
npts = 10_000_000
table = torch.from_numpy(bench.block_gtau(0, 8192, 10.0)).to(dev)
taus = torch.rand(npts, dtype=torch.float64, device=dev) * 10.0
o = torch.empty(npts, 25, dtype=torch.complex128, device=dev)
for _ in range(2):
    ctx.interp_tau(table, 10.0, taus, out=o)
del o
a = torch.randn(200_000, 16, 16, dtype=torch.complex128, device=dev)
ctx.invert_in_place(a)
ctx.synchronize()
ctx.close()
print("done")
