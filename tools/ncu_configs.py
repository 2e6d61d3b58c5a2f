"""One launch set of every config's kernels at profiling-friendly sizes.  Run plain first, then under ncu:
    python tools/ncu_configs.py [c1 c2 c3 c4 c5 fit density]
Sizes: large enough to fill the machine for several waves, small enough that ~40 ncu replays per launch stay short."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import triqs_b200  # noqa: E402

F = triqs_b200.FERMION
which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5", "fit", "density"]
ctx = triqs_b200.Context(0, print_warnings=False)
dev = torch.device("cuda", 0)
reps = 2

if "c1" in which:
    B = 2048
    gt = torch.from_numpy(bench.c1_inputs(B)).to(dev)
    gw = torch.empty(B, 2050, 1, dtype=torch.complex128, device=dev)
    for _ in range(reps):
        ctx.fourier_tau_to_iw(gt, 10.0, F, 1025, out=gw)
        ctx.fit_tail(gw, 10.0, F, hermitian=True, inner_dim=1)
    del gt, gw
if "c2" in which:
    R = 4
    blocks = [torch.from_numpy(bench.block_gtau(b)) for b in range(2)]
    gt = torch.stack([blocks[i % 2] * (1 + 0.01 * (i // 2)) for i in range(2 * R)]).to(dev)
    gw = torch.empty(2 * R, 20000, 25, dtype=torch.complex128, device=dev)
    gt2 = torch.empty_like(gt)
    for _ in range(reps):
        ctx.fourier_tau_to_iw(gt, 100.0, F, 10000, out=gw)
        ctx.fourier_iw_to_tau(gw, 100.0, F, 100001, out=gt2)
    del gt, gw, gt2
if "c3" in which:
    nk, no = 65536, 2304
    x = torch.randn(nk, no, dtype=torch.float64, device=dev).to(torch.complex128)
    y = torch.empty_like(x)
    for _ in range(reps):
        ctx.fourier_lattice(x, (256, 256, 1), True, out=y)
    del x, y
if "c4" in which:
    displ, hop, sig_h, dims = bench.c4_inputs()
    eps = ctx.tight_binding_fourier(displ, torch.from_numpy(hop).to(dev), dims)
    sig = torch.from_numpy(sig_h).to(dev)
    for _ in range(reps):
        ctx.dyson_ksum(eps[:65536].contiguous(), sig, 10.0, F, 0.1)
    a = torch.randn(1_000_000, 5, 5, dtype=torch.complex128, device=dev)
    ctx.invert_in_place(a)
    del eps, a
if "c5" in which:
    npts = 10_000_000
    table = torch.from_numpy(bench.block_gtau(0, 8192, 10.0)).to(dev)
    taus = torch.rand(npts, dtype=torch.float64, device=dev) * 10.0
    o = torch.empty(npts, 25, dtype=torch.complex128, device=dev)
    for _ in range(reps):
        ctx.interp_tau(table, 10.0, taus, out=o)
    del o
if "fit" in which:
    gw = torch.from_numpy(np.stack([bench.block_gtau(b)[:20000] for b in range(2)])).to(dev)  # any smooth [2][20000][25] data
    gw = gw.repeat(16, 1, 1)
    for _ in range(reps):
        ctx.fit_tail(gw, 100.0, F)
        ctx.fit_tail(gw, 100.0, F, hermitian=True, inner_dim=5)
if "density" in which:
    nb, nw, n = 32, 20000, 5
    gw = torch.randn(nb, nw, n, n, dtype=torch.complex128, device=dev) * 1e-3
    mom = torch.zeros(nb, 4, n, n, dtype=torch.complex128, device=dev)
    mom[:, 1] = torch.eye(n, dtype=torch.complex128, device=dev)
    for _ in range(reps):
        ctx.density(gw, 100.0, F, moments=mom)
ctx.synchronize()
ctx.close()
print("ncu_configs done")
