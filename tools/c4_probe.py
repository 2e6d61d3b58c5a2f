"""Dyson k-sum timing at the full C4 size (64^3 x 2048 x 5x5) with BASELINE inputs."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import triqs_b200  # noqa: E402

F = triqs_b200.FERMION
ctx = triqs_b200.Context(0, print_warnings=False)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
displ, hop, sig_h, dims = bench.c4_inputs()
eps = ctx.tight_binding_fourier(displ, torch.from_numpy(hop).to(dev), dims)
sig = torch.from_numpy(sig_h).to(dev)
gl = torch.empty_like(sig)
run = lambda: ctx.dyson_ksum(eps, sig, 10.0, F, 0.1, out=gl)
run()
ctx.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); run(); e1.record(stream); e1.synchronize(); ts.append(e0.elapsed_time(e1))
print(json.dumps({"lib": triqs_b200.LIB_PATH.split("/")[-1], "ms": float(np.median(ts)), "G_inv_s": 64 ** 3 * 2048 / float(np.median(ts)) / 1e6,
                  "checksum": complex(gl.sum().cpu())}.__repr__()))
