"""Developer probe: times the main kernels at (near) full config sizes with device-resident inputs.
Not part of the product or the tests; run under gpurun."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import triqs_b200

F = triqs_b200.FERMION
ctx = triqs_b200.Context(0, print_warnings=False)
stream = torch.cuda.ExternalStream(ctx.stream)
res = {}


def timeit(name, fn, reps=5, warm=2, bytes_=None, pts=None):
    for _ in range(warm):
        fn()
    ctx.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    r = {"ms": ms}
    if bytes_:
        r["GBps"] = bytes_ / ms / 1e6
    if pts:
        r["Mpts_s"] = pts / ms / 1e3
    res[name] = r
    print(name, json.dumps(r), flush=True)


which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
g = torch.Generator(device="cuda").manual_seed(0)

if "c1" in which:
    B = 8192
    gt = torch.randn(B, 10001, 1, dtype=torch.float64, device="cuda", generator=g).to(torch.complex128)
    gt = torch.cumsum(gt, dim=1) * 1e-3  # smooth-ish
    out = torch.empty(B, 2050, 1, dtype=torch.complex128, device="cuda")
    mom = torch.zeros(B, 4, 1, dtype=torch.complex128, device="cuda")
    mom[:, 1] = 1.0
    timeit("c1_tau2iw_B8192_known", lambda: ctx.fourier_tau_to_iw(gt, 10.0, F, 1025, moments=mom, out=out), bytes_=B * 192816.0, pts=B * 2050)
    timeit("c1_tau2iw_B8192_fit", lambda: ctx.fourier_tau_to_iw(gt, 10.0, F, 1025, out=out), bytes_=B * 192816.0, pts=B * 2050)
    del gt, out

if "c2" in which:
    R = 16
    gt = torch.randn(2 * R, 100001, 25, dtype=torch.float64, device="cuda", generator=g).to(torch.complex128) * 1e-3
    gw = torch.empty(2 * R, 20000, 25, dtype=torch.complex128, device="cuda")
    mom = torch.zeros(2 * R, 4, 25, dtype=torch.complex128, device="cuda")
    mom[:, 1] = 1.0
    fb = 2 * R * (100001 + 20000) * 25 * 16.0
    timeit("c2_tau2iw_x16_known", lambda: ctx.fourier_tau_to_iw(gt, 100.0, F, 10000, moments=mom, out=gw), bytes_=fb, pts=2 * R * 20000)
    timeit("c2_tau2iw_x16_fit", lambda: ctx.fourier_tau_to_iw(gt, 100.0, F, 10000, out=gw), bytes_=fb, pts=2 * R * 20000)
    gw.normal_(generator=g) if False else None
    timeit("c2_iw2tau_x16_known", lambda: ctx.fourier_iw_to_tau(gw, 100.0, F, 100001, moments=mom, out=gt), bytes_=fb, pts=2 * R * 100001)
    # single BlockGf (2 blocks)
    timeit("c2_tau2iw_x1_known", lambda: ctx.fourier_tau_to_iw(gt[:2], 100.0, F, 10000, moments=mom[:2], out=gw[:2]), bytes_=fb / R,
           pts=2 * 20000)
    del gt, gw

if "c3" in which:
    nk, no = 65536, 9216
    x = torch.randn(nk, no, dtype=torch.float64, device="cuda", generator=g).to(torch.complex128)
    y = torch.empty_like(x)
    timeit("c3_lattice_256x256x9216", lambda: ctx.fourier_lattice(x, (256, 256, 1), True, out=y), reps=10, warm=3, bytes_=2.0 * nk * no * 16,
           pts=nk * 1024)
    del x, y

if "c4" in which:
    nk, nw, n = 32768, 2048, 5
    eps = torch.randn(nk, n, n, dtype=torch.complex128, device="cuda", generator=g) * 0.3
    eps = eps + eps.conj().transpose(1, 2)
    sig = torch.randn(nw, n, n, dtype=torch.complex128, device="cuda", generator=g) * 0.1
    out = torch.empty_like(sig)
    timeit("c4_dyson_32k_k_x2048w", lambda: ctx.dyson_ksum(eps, sig, 10.0, F, 0.1, out=out), reps=3, warm=1, pts=nk * nw)
    a = torch.randn(4_000_000, n, n, dtype=torch.complex128, device="cuda", generator=g)
    timeit("invert_4M_5x5", lambda: ctx.invert_in_place(a), reps=3, warm=1, bytes_=2.0 * a.numel() * 16, pts=a.shape[0])
    del eps, a

if "c5" in which:
    npts = 20_000_000
    table = torch.randn(8192, 25, dtype=torch.complex128, device="cuda", generator=g)
    taus = torch.rand(npts, dtype=torch.float64, device="cuda", generator=g) * 10.0
    out = torch.empty(npts, 25, dtype=torch.complex128, device="cuda")
    timeit("c5_interp_2e7", lambda: ctx.interp_tau(table, 10.0, taus, out=out), bytes_=npts * 408.0, pts=npts)

json.dump(res, open("gpurun_out/probe.json", "w"), indent=1)
