"""Developer probe: latency of SMALL calls (what an interactive TRIQS session issues): one BlockGf of 2 blocks, host NumPy buffers and device tensors,
wall time per call, next to the CPU oracle (NumPy / pocketfft) on the same input."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import triqs_b200
from oracle import triqs_oracle as O
F = triqs_b200.FERMION
ctx = triqs_b200.Context(0, print_warnings=False)
def wall(fn, n=50, warm=5):
    for _ in range(warm): fn()
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    ctx.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
for L, n_iw, nc in [(6150, 1025, 9), (10000, 1025, 1), (10000, 1025, 25), (1000, 100, 4), (100000, 10000, 25)]:
    n_tau = L + 1
    rng = np.random.default_rng(0)
    gt = (rng.normal(size=(2, n_tau, nc)) * 1e-3).astype(complex)
    mom = np.zeros((2, 4, nc), complex); mom[:, 1] = 1
    gw = np.empty((2, 2 * n_iw, nc), complex)
    gtd = torch.from_numpy(gt).cuda(); gwd = torch.empty(2, 2 * n_iw, nc, dtype=torch.complex128, device="cuda"); momd = torch.from_numpy(mom).cuda()
    res = {"L": L, "n_cols": nc}
    res["host_fwd_ms"] = round(wall(lambda: ctx.fourier_tau_to_iw(gt, 10.0, F, n_iw, moments=mom, out=gw)), 4)
    res["host_inv_ms"] = round(wall(lambda: ctx.fourier_iw_to_tau(gw, 10.0, F, n_tau, moments=mom, out=gt)), 4)
    res["host_fwd_unknown_moments_ms"] = round(wall(lambda: ctx.fourier_tau_to_iw(gt, 10.0, F, n_iw, out=gw)), 4)
    res["host_inv_fitted_tail_ms"] = round(wall(lambda: ctx.fourier_iw_to_tau(gw, 10.0, F, n_tau, out=gt)), 4)
    res["device_fwd_ms"] = round(wall(lambda: ctx.fourier_tau_to_iw(gtd, 10.0, F, n_iw, moments=momd, out=gwd)), 4)
    res["device_inv_ms"] = round(wall(lambda: ctx.fourier_iw_to_tau(gwd, 10.0, F, n_tau, moments=momd, out=gtd)), 4)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        for b in range(2):
            O.fourier_tau_to_iw(gt[b], 10.0, F, n_iw, mom[b])
    res["cpu_oracle_fwd_ms"] = round((time.perf_counter() - t0) / reps * 1e3, 3)
    print(res, flush=True)
