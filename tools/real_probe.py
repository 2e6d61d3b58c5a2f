"""Developer probe: device-resident real-valued G(tau), forward leg: two-for-one against promotion (C2 batch)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import triqs_b200
F = triqs_b200.FERMION
ctx = triqs_b200.Context(0, print_warnings=False)
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
for nc in (25, 16):
    B, n_tau, n_iw = 32, 100001, 10000
    gr = torch.randn(B, n_tau, nc, dtype=torch.float64, device=dev) * 1e-3
    gc = gr.to(torch.complex128)
    gw = torch.empty(B, 2 * n_iw, nc, dtype=torch.complex128, device=dev)
    mom = torch.zeros(B, 4, nc, dtype=torch.complex128); mom[:, 1] = 1.0
    momd = mom.to(dev)
    res = {"n_cols": nc}
    res["complex_ms"] = timed(lambda: ctx.fourier_tau_to_iw(gc, 100.0, F, n_iw, moments=momd, out=gw))
    res["real_two_for_one_ms"] = timed(lambda: ctx.fourier_tau_to_iw(gr, 100.0, F, n_iw, moments=mom.numpy(), out=gw))
    ctx.set_option("no_two_for_one", 1)
    res["real_promoted_ms"] = timed(lambda: ctx.fourier_tau_to_iw(gr, 100.0, F, n_iw, moments=mom.numpy(), out=gw))
    ctx.set_option("no_two_for_one", 0)
    res["real_two_for_one_unknown_moments_ms"] = timed(lambda: ctx.fourier_tau_to_iw(gr, 100.0, F, n_iw, out=gw))
    ctx.profile_enable(True)
    ctx.fourier_tau_to_iw(gr, 100.0, F, n_iw, moments=mom.numpy(), out=gw)
    res["kernels"] = {a: round(b["ms"], 3) for a, b in ctx.profile_report().items()}
    ctx.profile_enable(False)
    print(res, flush=True)
