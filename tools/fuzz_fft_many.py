"""Developer tool: random ranks / axis lengths / column counts through tq_fft_many and tq_fourier_lattice, against NumPy's FFT (1e-12).
usage: fuzz_fft_many.py [n_cases] [seed]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import triqs_b200

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ctx = triqs_b200.Context(0, print_warnings=False)
rel = lambda x, r: float(np.max(np.abs(np.asarray(x) - r)) / max(np.max(np.abs(r)), 1e-300))
pool = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 15, 16, 20, 24, 25, 27, 30, 32, 36, 41, 48, 50, 64, 72, 96, 100, 101, 120, 125, 128, 131, 149, 200, 243, 256, 300,
        384, 400, 500, 512, 600, 625, 640, 720, 1000, 1009, 1024, 1500, 2003, 2048, 4096, 5000]
worst, bad = 0.0, 0
tags = {}
for case in range(n_cases):
    rank = int(rng.integers(1, 4))
    while True:
        dims = [int(rng.choice(pool)) for _ in range(rank)]
        hm = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 9, 16, 24, 25, 26, 33, 64, 100, 130, 257]))
        n = int(np.prod(dims))
        if n * hm <= 6_000_000:
            break
    x = rng.normal(size=(n, hm)) + 1j * rng.normal(size=(n, hm))
    for sign in (+1, -1):
        ctx.profile_enable(True)
        y = ctx.fft_many(x, dims, sign)
        for k in ctx.profile_report():
            tags[k] = tags.get(k, 0) + 1
        ctx.profile_enable(False)
        xr = x.reshape(tuple(dims) + (hm,))
        axes = tuple(range(rank))
        ref = (np.fft.ifftn(xr, axes=axes) * n if sign > 0 else np.fft.fftn(xr, axes=axes)).reshape(n, hm)
        e = rel(y, ref)
        worst = max(worst, e)
        if not e < 1e-12:
            bad += 1
            print("FAIL", dims, hm, sign, e, flush=True)
    if rank == 3 or (rank == 2 and rng.integers(0, 2)):
        d3 = tuple(dims) + (1,) * (3 - rank)
        gk = ctx.fourier_lattice(x, d3, to_k=True)
        back = ctx.fourier_lattice(gk, d3, to_k=False)
        e = rel(back, x)
        worst = max(worst, e)
        if not e < 1e-12:
            bad += 1
            print("FAIL lattice round trip", d3, hm, e, flush=True)
print(f"{n_cases} cases, worst rel err {worst:.2e}, failures {bad}")
print("kernel tags seen:", dict(sorted(tags.items())))
