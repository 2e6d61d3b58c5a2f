"""Developer check: large PAGEABLE host buffers through the other entry points (lattice, fft_many, interp_tau, dyson, density, evaluate): the copies
spread over helper threads (default) give the same bits as the plain single-stream copies ("host_lanes" = 1)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import triqs_b200
from oracle import triqs_oracle as O
F = triqs_b200.FERMION
ctx = triqs_b200.Context(0, print_warnings=False)
rng = np.random.default_rng(2)
def both(fn):
    ctx.set_option("host_lanes", 1)
    a = fn()
    ctx.set_option("host_lanes", 3)
    b = fn()
    return a, b
ok = True
x = rng.normal(size=(128 * 128, 96)) + 1j * rng.normal(size=(128 * 128, 96))          # 25 MB
a, b = both(lambda: ctx.fourier_lattice(x, (128, 128, 1), True))
print("lattice", np.array_equal(a, b), float(np.max(np.abs(a - O.fourier_lattice_r_to_k(x, (128, 128, 1)))) / np.max(np.abs(a)))); ok &= np.array_equal(a, b)
a, b = both(lambda: ctx.fft_many(x, (128, 128), -1))
print("fft_many", np.array_equal(a, b)); ok &= np.array_equal(a, b)
table = rng.normal(size=(4096, 25)) + 1j * rng.normal(size=(4096, 25))
taus = rng.uniform(0, 10.0, size=2_000_000)                                              # 16 MB in, 800 MB out
a, b = both(lambda: ctx.interp_tau(table, 10.0, taus))
print("interp_tau", np.array_equal(a, b)); ok &= np.array_equal(a, b)
eps = rng.normal(size=(40000, 5, 5)) + 1j * rng.normal(size=(40000, 5, 5)); eps = eps + eps.conj().transpose(0, 2, 1)   # 16 MB
sig = np.stack([np.eye(5) / (w - 0.5) for w in O.imfreq_values(10.0, F, 16)])
a, b = both(lambda: ctx.dyson_ksum(eps, sig, 10.0, F, 0.1))
print("dyson_ksum", np.array_equal(a, b)); ok &= np.array_equal(a, b)
gw = rng.normal(size=(8, 20000, 25)) + 1j * rng.normal(size=(8, 20000, 25))               # 64 MB
a, b = both(lambda: ctx.make_hermitian(gw.reshape(8, 20000, 5, 5)) if hasattr(ctx, "make_hermitian") else gw)
print("make_hermitian", np.array_equal(np.asarray(a), np.asarray(b))); ok &= np.array_equal(np.asarray(a), np.asarray(b))
print("ALL OK" if ok else "MISMATCH")
