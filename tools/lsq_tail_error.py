"""Prints the parity of the return leg (imfreq -> imtime) when the library fits the tail itself, against the oracle, next to
the oracle's own LAPACK-driver spread (gesvd vs gesdd) -- VERDICT r1 "weak" #2.  Run on a GPU box."""
import json
import os
import sys

import numpy as np
import scipy.linalg as sl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import triqs_oracle as O  # noqa: E402


def mk(driver):
    def f(A):
        U, s, Vh = sl.svd(A, full_matrices=True, lapack_driver=driver)
        M, N = A.shape
        return (Vh.conj().T / s[None, :]) @ U[:, :N].conj().T, U.conj().T[N:M, :], s
    return f


def matrix_gf(beta, n_iw, n, seed):
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
    h = a + a.conj().T
    w, v = np.linalg.eigh(h)
    w = -2 + 4 * (w - w.min()) / (w.max() - w.min())
    H = (v * w) @ v.conj().T
    iw = O.imfreq_values(beta, 1, n_iw)
    return np.linalg.inv(iw[:, None, None] * np.eye(n)[None] - H[None]).reshape(len(iw), n * n)


def main():
    import triqs_b200
    ctx = triqs_b200.Context(0, print_warnings=False)
    out = {}
    for name, beta, n_iw, n_tau in [("c2", 100.0, 10000, 100001), ("ref_test_mesh", 10.0, 1000, 10000), ("default_6150", 10.0, 1025, 6151)]:
        gw = matrix_gf(beta, n_iw, 5, 2001)
        res = {}
        for d in ("gesvd", "gesdd"):
            O._svd_pinv = mk(d)
            res[d] = O.fourier_iw_to_tau(gw, beta, 1, n_tau, return_info=True)
        ref, info = res["gesvd"]
        gt = ctx.fourier_iw_to_tau(gw[None], beta, 1, n_tau)[0]
        mom, err = ctx.fit_tail(gw[None], beta, 1, known_moments=np.zeros((1, 1, 25), complex))
        nm = min(mom.shape[1], info["tail"].shape[0])
        out[name] = {
            "gpu_vs_oracle_rel": float(np.max(np.abs(gt - ref)) / np.max(np.abs(ref))),
            "oracle_gesvd_vs_gesdd_rel": float(np.max(np.abs(res["gesdd"][0] - ref)) / np.max(np.abs(ref))),
            "moment_abs_diff_gpu_vs_oracle": [float(np.max(np.abs(mom[0][p] - info["tail"][p]))) for p in range(nm)],
            "moment_abs_diff_gesvd_vs_gesdd": [float(np.max(np.abs(res["gesdd"][1]["tail"][p] - info["tail"][p]))) for p in range(info["tail"].shape[0])],
            "n_moments_gpu": int(mom.shape[1]), "n_moments_oracle": int(info["tail"].shape[0]),
        }
    print(json.dumps(out, indent=1))
    ctx.close()


if __name__ == "__main__":
    main()
