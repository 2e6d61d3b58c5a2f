"""Does the second axis pass of a 2-D lattice transform run faster when the slab it reads was just written and fits L2?
(premise of fusing both axes per column slab).  Rotates over several arrays so that inputs are cold."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import triqs_b200  # noqa: E402

ctx = triqs_b200.Context(0, print_warnings=False)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
dims = (256, 256, 1)
nk = 65536
for no in (16, 32, 48, 64, 96, 128, 256):
    nrot = max(2, int(1.2e9 / (nk * no * 16)))
    xs = [torch.view_as_complex(torch.randn(nk, no, 2, dtype=torch.float64, device=dev)) for _ in range(nrot)]
    ys = [torch.empty_like(x) for x in xs]
    for i in range(nrot):
        ctx.fourier_lattice(xs[i], dims, True, out=ys[i])
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record(stream)
    for _ in range(reps):
        for i in range(nrot):
            ctx.fourier_lattice(xs[i], dims, True, out=ys[i])
    e1.record(stream)
    e1.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * nrot)
    mb = nk * no * 16 / 1e6
    print(json.dumps({"n_others": no, "slab_MB": mb, "us_per_2d_transform": us, "GBps_algorithmic": 2 * mb / us * 1e3 / 1e3,
                      "us_per_64col_equiv": us * 64 / no}), flush=True)
    del xs, ys
    torch.cuda.empty_cache()
