"""Developer probe: round-trip throughput for a few mesh lengths (25 columns, 1.2 GB per call, known moments)."""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
import triqs_b200
F = triqs_b200.FERMION
ctx = triqs_b200.Context(0, print_warnings=False)
import os
if os.environ.get("BLUE_CHUNK_MB"):
    ctx.set_option("blue_chunk_mb", int(os.environ["BLUE_CHUNK_MB"]))
if os.environ.get("FORCE_BLUESTEIN"):
    ctx.set_option("force_bluestein", 1)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
def timed(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    ctx.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
cases = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(6150, 1025, 25), (61500, 10250, 25), (6000, 1000, 25)]
for L, n_iw, nc in cases:
    n_tau, n_w = L + 1, 2 * n_iw
    batch = max(1, int(1.2e9 / (n_tau * nc * 16)))
    gt = torch.view_as_complex(torch.randn(batch, n_tau, nc, 2, dtype=torch.float64, device=dev)).mul_(1e-3)
    gw = torch.empty(batch, n_w, nc, dtype=torch.complex128, device=dev)
    gt2 = torch.empty_like(gt)
    mom = torch.zeros(batch, 4, nc, dtype=torch.complex128, device=dev); mom[:, 1] = 1.0
    ms_f = timed(lambda: ctx.fourier_tau_to_iw(gt, 20.0, F, n_iw, moments=mom, out=gw))
    ms_i = timed(lambda: ctx.fourier_iw_to_tau(gw, 20.0, F, n_tau, moments=mom, out=gt2))
    ctx.profile_enable(True)
    ctx.fourier_tau_to_iw(gt, 20.0, F, n_iw, moments=mom, out=gw)
    ctx.fourier_iw_to_tau(gw, 20.0, F, n_tau, moments=mom, out=gt2)
    k = {a: round(b["ms"], 3) for a, b in ctx.profile_report().items()}
    ctx.profile_enable(False)
    by = batch * (n_tau + n_w) * nc * 16.0
    print(json.dumps({"L": L, "n_cols": nc, "batch": batch, "ms_fwd": round(ms_f, 3), "ms_inv": round(ms_i, 3), "GBps_roundtrip": round(2 * by / (ms_f + ms_i) / 1e6), "kernels": k}), flush=True)
    del gt, gw, gt2, mom
    torch.cuda.empty_cache()
