"""Round-trip throughput of the Matsubara transforms over mesh lengths, and of the lattice transform over axis lengths
("a bench table over L in {6000, 6150, 9999, 1e4 x 25 cols, 61500, 1e5, 2e5} and lattice axes {32, 96, 512}").
Device-resident inputs, CUDA events, known moments (transforms only).  Prints one JSON object per case."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import triqs_b200  # noqa: E402

F = triqs_b200.FERMION
ctx = triqs_b200.Context(0, print_warnings=False)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
which = sys.argv[1:] or ["matsubara", "lattice"]
import os
if "TQ_LAG" in os.environ:
    ctx.set_option("pipe_lag0", int(os.environ["TQ_LAG"]))


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ctx.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


out = {"matsubara": [], "lattice": []}
target_bytes = 1.2e9   # G(tau) bytes per call: larger than L2
cases = [(6000, 1000, 25), (6144, 1024, 25), (6150, 1025, 25), (9999, 1000, 25), (10000, 1025, 25), (12000, 2000, 25), (61500, 10250, 25),
         (100000, 10000, 25), (200000, 20000, 25), (65536, 8192, 25), (10000, 1025, 9), (6150, 1025, 4), (100000, 10000, 9)]
if "matsubara" in which:
    for L, n_iw, nc in cases:
        n_tau, n_w = L + 1, 2 * n_iw
        batch = max(1, int(target_bytes / (n_tau * nc * 16)))
        g = torch.Generator(device=dev).manual_seed(L)
        gt = torch.randn(batch, n_tau, nc, 2, dtype=torch.float64, device=dev, generator=g)
        gt = torch.view_as_complex(gt).mul_(1e-3)
        gw = torch.empty(batch, n_w, nc, dtype=torch.complex128, device=dev)
        gt2 = torch.empty_like(gt)
        mom = torch.zeros(batch, 4, nc, dtype=torch.complex128, device=dev)
        mom[:, 1] = 1.0
        rec = {"L": L, "n_iw": n_iw, "n_cols": nc, "batch": batch}
        variants = [("planned", {}), ("generic", {"no_pipe": 1})]
        if L == 100000:
            variants.append(("planned_forced", {"force_planned": 1}))
        for label, opts in variants:
            for k, v in opts.items():
                ctx.set_option(k, v)
            try:
                ms_f = timed(lambda: ctx.fourier_tau_to_iw(gt, 20.0, F, n_iw, moments=mom, out=gw))
                ms_i = timed(lambda: ctx.fourier_iw_to_tau(gw, 20.0, F, n_tau, moments=mom, out=gt2))
                bytes_rt = 2.0 * batch * (n_tau + n_w) * nc * 16
                rec[label] = {"ms_fwd": ms_f, "ms_inv": ms_i, "GBps_round_trip": bytes_rt / ((ms_f + ms_i) * 1e-3) / 1e9}
                ctx.profile_enable(True)
                ctx.fourier_tau_to_iw(gt, 20.0, F, n_iw, moments=mom, out=gw)
                ctx.fourier_iw_to_tau(gw, 20.0, F, n_tau, moments=mom, out=gt2)
                rec[label]["kernels_ms"] = {k: round(v["ms"], 4) for k, v in ctx.profile_report().items()}
                ctx.profile_enable(False)
            except Exception as e:
                rec[label] = {"error": repr(e)[:200]}
            for k in opts:
                ctx.set_option(k, 0)
        out["matsubara"].append(rec)
        print(json.dumps(rec), flush=True)
        del gt, gw, gt2, mom
        torch.cuda.empty_cache()

if "lattice" in which:
    for dims, no in [((256, 256, 1), 2304), ((32, 32, 32), 4096), ((96, 96, 1), 9216), ((512, 512, 1), 1024), ((64, 64, 64), 1024),
                     ((100, 100, 1), 9216), ((48, 48, 48), 1024)]:
        nk = int(np.prod(dims))
        x = torch.view_as_complex(torch.randn(nk, no, 2, dtype=torch.float64, device=dev))
        y = torch.empty_like(x)
        rec = {"dims": dims, "n_others": no}
        for label, opts in (("pipe", {}), ("generic", {"no_pipe": 1})):
            for k, v in opts.items():
                ctx.set_option(k, v)
            try:
                ms = timed(lambda: ctx.fourier_lattice(x, dims, True, out=y), reps=5, warm=2)
                rec[label] = {"ms": ms, "GBps_algorithmic": 2.0 * nk * no * 16 / (ms * 1e-3) / 1e9}
            except Exception as e:
                rec[label] = {"error": repr(e)[:200]}
            for k in opts:
                ctx.set_option(k, 0)
        out["lattice"].append(rec)
        print(json.dumps(rec), flush=True)
        del x, y
        torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/bench_lengths.json", "w"), indent=1)
