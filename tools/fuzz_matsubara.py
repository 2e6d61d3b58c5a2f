"""Developer tool: random mesh lengths / column counts / batch sizes / statistics through both Matsubara transforms, every result against
the oracle (1e-12 with known moments).  Exercises whatever engine the dispatcher picks: planned, heavy, generic, single-column,
"gfs as columns", Bluestein, direct.  usage: fuzz_matsubara.py [n_cases] [seed]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import triqs_b200
from oracle import triqs_oracle as O

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
ctx = triqs_b200.Context(0, print_warnings=False)
rel = lambda x, r: float(np.max(np.abs(np.asarray(x) - r)) / max(np.max(np.abs(r)), 1e-300))
worst, bad = 0.0, 0
engines = {}
for case in range(n_cases):
    kind = rng.integers(0, 5)
    if kind == 0:
        L = int(rng.integers(64, 3000))
    elif kind == 1:
        L = int(rng.choice([1999, 2003, 4999, 8191, 10007, 2 * 5003, 3 * 1009, 7 * 2857]))
    elif kind == 2:
        L = int(rng.choice([6000, 6144, 6150, 9999, 10000, 12000, 6600, 7800, 10200, 4096, 5000, 7500]))
    elif kind == 3:
        L = int(rng.choice([20000, 30000, 61500, 65536, 100000]))
    else:
        L = int(rng.integers(3000, 20000))
    stat = int(rng.choice([O.FERMION, O.BOSON]))
    n_iw = int(rng.integers(8, max(9, L // 6)))
    C = int(rng.choice([1, 1, 2, 3, 4, 5, 9, 16, 25]))
    B = int(rng.choice([1, 2, 3, 9, 17])) if L < 30000 else int(rng.choice([1, 2]))
    beta = float(rng.choice([1.0, 10.0, 40.0]))
    iw = O.imfreq_values(beta, stat, n_iw)
    e = rng.uniform(0.2, 2.0, size=(B, 3, C)) * rng.choice([-1.0, 1.0], size=(B, 3, C))
    r = rng.uniform(0.2, 1.0, size=(B, 3, C))
    gw = np.sum(r[:, None] / (iw[None, :, None, None] - e[:, None]), axis=2)
    mom = np.zeros((B, 4, C), complex)
    for p in range(1, 4):
        mom[:, p] = np.sum(r * e ** (p - 1), axis=1)
    ctx.profile_enable(True)
    gt = ctx.fourier_iw_to_tau(gw, beta, stat, L + 1, moments=mom)
    back = ctx.fourier_tau_to_iw(gt, beta, stat, n_iw, moments=mom)
    rep = ctx.profile_report()
    ctx.profile_enable(False)
    for k in rep:
        engines[k] = engines.get(k, 0) + 1
    for b in sorted({0, B - 1}):
        rt = O.fourier_iw_to_tau(gw[b], beta, stat, L + 1, mom[b])
        e1 = rel(gt[b], rt)
        e2 = rel(back[b], O.fourier_tau_to_iw(rt, beta, stat, n_iw, mom[b]))
        worst = max(worst, e1, e2)
        if not (e1 < 1e-12 and e2 < 1e-12):
            bad += 1
            print("FAIL", dict(L=L, stat=stat, n_iw=n_iw, C=C, B=B, beta=beta, b=b, e1=e1, e2=e2, kernels=sorted(rep)), flush=True)
print(f"{n_cases} cases, worst rel err {worst:.2e}, failures {bad}")
print("kernel tags seen:", dict(sorted(engines.items())))
