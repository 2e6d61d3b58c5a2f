"""End-to-end (host buffers through the C ABI) sweep over the internal lane count and chunk size: one host thread, one call per
direction for the whole batch of 16 BlockGfs -- pinned and pageable buffers."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import triqs_b200  # noqa: E402

F = triqs_b200.FERMION
numa = bench.bind_to_gpu_numa_node(0) if "--bind" in sys.argv else {"bound": False}
ctx = triqs_b200.Context(0, print_warnings=False)
R, NB = 16, 2
blocks = [torch.from_numpy(bench.block_gtau(b)) for b in range(NB)]
host_gt = torch.empty((NB * R, bench.N_TAU, bench.NCOL), dtype=torch.complex128).pin_memory()
for i in range(NB * R):
    host_gt[i] = blocks[i % NB] * (1.0 + 0.01 * (i // NB))
host_gw = torch.empty((NB * R, bench.N_W, bench.NCOL), dtype=torch.complex128).pin_memory()
host_gt2 = torch.empty_like(host_gt).pin_memory()
pin = (host_gt.numpy(), host_gw.numpy(), host_gt2.numpy())
page = (np.array(pin[0]), np.empty_like(pin[1]), np.empty_like(pin[2]))
pts = R * bench.POINTS_PER_BLOCKGF
print(json.dumps({"numa": numa}), flush=True)
for kind, (a, b, c) in (("pageable", page), ("pinned", pin)):
    for lanes in ((1, 2, 3, 4, 6, 8) if kind == "pageable" else (3,)):
        for chunk_mb in ((48, 96) if kind == "pageable" else (96,)):
            ctx.set_option("host_lanes", lanes)
            ctx.set_option("host_chunk_kb", chunk_mb << 10)

            def step():
                ctx.fourier_tau_to_iw(a, bench.BETA, F, bench.N_IW, out=b)
                ctx.fourier_iw_to_tau(b, bench.BETA, F, bench.N_TAU, out=c)
            step()
            t0 = time.perf_counter()
            n = 3
            for _ in range(n):
                step()
            dt = (time.perf_counter() - t0) / n
            print(json.dumps({"memory": kind, "lanes": lanes, "chunk_mb": chunk_mb, "ms": dt * 1e3, "Mpts_s": pts / dt / 1e6,
                              "GBps_each_way": 1.536 / dt}), flush=True)
        if kind == "pageable" and lanes == 1:
            pass
