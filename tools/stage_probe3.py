import sys, time
import numpy as np
sys.path.insert(0, ".")
import triqs_b200
F = triqs_b200.FERMION
ctx = triqs_b200.Context(0, print_warnings=False)
def wall(fn, n=15, warm=3):
    for _ in range(warm): fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e3
rng = np.random.default_rng(0)
for nb in (2, 8, 32):
    gt = (rng.normal(size=(nb, 100001, 25)) * 1e-3).astype(complex)
    mom = np.zeros((nb, 4, 25), complex); mom[:, 1] = 1
    gw = np.empty((nb, 20000, 25), complex)
    for nt in (1, 0, 1, 0):
        ctx.set_option("stage_nt", nt)
        for T in (0, 16):
            ctx.set_option("stage_threads", T)
            f = wall(lambda: ctx.fourier_tau_to_iw(gt, 100.0, F, 10000, moments=mom, out=gw))
            i = wall(lambda: ctx.fourier_iw_to_tau(gw, 100.0, F, 100001, moments=mom, out=gt))
            print({"blocks": nb, "nt": nt, "threads": T, "fwd_ms": round(f, 3), "inv_ms": round(i, 3)}, flush=True)
