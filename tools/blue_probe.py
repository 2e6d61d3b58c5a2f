"""Developer probe: lengths without a plan -- Bluestein against the direct O(N^2) window DFT."""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
import triqs_b200
F = triqs_b200.FERMION
ctx = triqs_b200.Context(0, print_warnings=False)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
def timed(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    ctx.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
for L, n_iw, nc, batch in [(1999, 333, 25, 64), (4999, 833, 25, 32), (8191, 1365, 25, 32), (19999, 3333, 25, 16), (49999, 8333, 25, 8), (199999, 10000, 25, 2), (20000, 3333, 25, 16)]:
    n_tau, n_w = L + 1, 2 * n_iw
    gt = torch.view_as_complex(torch.randn(batch, n_tau, nc, 2, dtype=torch.float64, device=dev)).mul_(1e-3)
    gw = torch.empty(batch, n_w, nc, dtype=torch.complex128, device=dev)
    gt2 = torch.empty_like(gt)
    mom = torch.zeros(batch, 4, nc, dtype=torch.complex128, device=dev); mom[:, 1] = 1.0
    res = {"L": L, "n_cols": nc, "batch": batch}
    for name, nb in (("bluestein", 0), ("direct", 1)):
        ctx.set_option("no_bluestein", nb)
        res[name + "_ms_fwd"] = timed(lambda: ctx.fourier_tau_to_iw(gt, 20.0, F, n_iw, moments=mom, out=gw))
        res[name + "_ms_inv"] = timed(lambda: ctx.fourier_iw_to_tau(gw, 20.0, F, n_tau, moments=mom, out=gt2))
    ctx.set_option("no_bluestein", 0)
    by = batch * (n_tau + n_w) * nc * 16.0
    res["bluestein_GBps_roundtrip"] = 2 * by / (res["bluestein_ms_fwd"] + res["bluestein_ms_inv"]) / 1e6
    res["direct_GBps_roundtrip"] = 2 * by / (res["direct_ms_fwd"] + res["direct_ms_inv"]) / 1e6
    print(json.dumps(res), flush=True)
