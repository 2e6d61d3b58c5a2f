"""Generates tests/golden/hotpath_v1.npz: seeded inputs and the oracle's outputs for every entry point of the hot path.

The reference keeps no binary fixtures for this path (its tests build analytic inputs), and it cannot be
built or imported in this environment, so the vectors come from the pinned oracle (oracle/triqs_oracle.py, checked against the
reference's known answers in tests/test_oracle_golden.py).  They freeze that oracle: tests/test_golden_fixtures.py checks on
the CPU that the oracle still reproduces them and on the GPU that the CUDA path matches them.

    python tests/golden/make_golden.py        # rewrites the .npz (deterministic)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import triqs_oracle as O  # noqa: E402


def cases():
    """name -> (inputs dict, outputs dict).  Inputs are regenerated from the seeds by the tests; only small ones are stored."""
    out = {}
    # --- C1-like: scalar G(tau), 4 poles, direct transform + hermitian tail fit (BASELINE configs[0] shape, shorter mesh)
    rng = np.random.default_rng(1001)
    beta, n_tau, n_iw = 10.0, 2001, 200
    E, r = rng.uniform(-2, 2, 4), rng.dirichlet(np.ones(4))
    tau = O.tau_values(beta, n_tau)
    gt = np.zeros((n_tau, 1), complex)
    for e, w in zip(E, r):
        gt[:, 0] += w * (-np.exp(-e * tau) / (1 + np.exp(-beta * e)) if e > 0 else -np.exp(e * (beta - tau)) / (1 + np.exp(beta * e)))
    gw = O.fourier_tau_to_iw(gt, beta, O.FERMION, n_iw)
    tail, err = O.fit_hermitian_tail(gw, beta, O.FERMION, inner_dim=1)
    out["c1"] = dict(beta=beta, n_iw=n_iw, gt=gt, gw=gw, tail=tail, err=err)
    # --- C2-like: 5x5 block, G = (iw - H)^-1, round trip with fitted tail
    rng = np.random.default_rng(2001)
    beta, n_iw, n_tau = 20.0, 150, 1201
    a = rng.normal(size=(5, 5)) + 1j * rng.normal(size=(5, 5))
    H = 0.5 * (a + a.conj().T)
    lam, U = np.linalg.eigh(H)
    H = (U * (4 * (lam - lam.min()) / (lam.max() - lam.min()) - 2)) @ U.conj().T
    iw = O.imfreq_values(beta, O.FERMION, n_iw)
    gw = np.linalg.inv(iw[:, None, None] * np.eye(5)[None] - H[None]).reshape(-1, 25)
    gt = O.fourier_iw_to_tau(gw, beta, O.FERMION, n_tau)
    gw_back = O.fourier_tau_to_iw(gt, beta, O.FERMION, n_iw)
    fit, ferr = O.fit_tail(gw, beta, O.FERMION)
    out["c2"] = dict(beta=beta, n_tau=n_tau, gw=gw, gt=gt, gw_back=gw_back, fit=fit, fit_err=ferr, density=O.density(gw.reshape(-1, 5, 5), beta, O.FERMION))
    # --- boson leg with known moments
    rng = np.random.default_rng(2002)
    beta, n_iw, n_tau = 5.0, 60, 481
    iw = O.imfreq_values(beta, O.BOSON, n_iw)
    e = rng.uniform(0.3, 1.5, 3)
    gw = (1 / (iw[:, None] - e[None, :])).astype(complex)
    mom = np.zeros((4, 3), complex)
    mom[1], mom[2], mom[3] = 1, e, e ** 2
    out["boson"] = dict(beta=beta, n_tau=n_tau, gw=gw, mom=mom, gt=O.fourier_iw_to_tau(gw, beta, O.BOSON, n_tau, mom))
    # --- lattice, plain DFT
    rng = np.random.default_rng(3001)
    dims = (6, 5, 4)
    gr = rng.normal(size=(120, 3)) + 1j * rng.normal(size=(120, 3))
    out["lattice"] = dict(dims=np.array(dims), gr=gr, gk=O.fourier_lattice_r_to_k(gr, dims), back=O.fourier_lattice_k_to_r(O.fourier_lattice_r_to_k(gr, dims), dims),
                          fft=O.fft_many(gr, dims, -1))
    # --- Dyson k-sum, inversion
    rng = np.random.default_rng(4001)
    n, nk, n_iw, beta, mu = 5, 48, 20, 10.0, 0.1
    t = rng.normal(size=(nk, n, n)) + 1j * rng.normal(size=(nk, n, n))
    eps = 0.25 * (t + t.conj().transpose(0, 2, 1))
    iw = O.imfreq_values(beta, O.FERMION, n_iw)
    V = rng.normal(size=(2, n)) + 1j * rng.normal(size=(2, n))
    sig = sum(np.outer(V[p], V[p].conj())[None] / (iw[:, None, None] - (0.5 - p)) for p in range(2))
    out["dyson"] = dict(beta=beta, mu=mu, eps=eps, sigma=sig, gloc=O.dyson_ksum(eps, sig, beta, O.FERMION, mu), inv=O.invert_batched(eps + 2 * np.eye(n)[None]))
    # --- interpolation
    rng = np.random.default_rng(5001)
    table = rng.normal(size=(64, 4)) + 1j * rng.normal(size=(64, 4))
    taus = np.concatenate([rng.uniform(0, 10.0, 200), [0.0, 10.0]])
    out["interp"] = dict(beta=10.0, table=table, taus=taus, out=O.interp_tau(table, 10.0, taus))
    # --- companions: real time, Legendre, tight binding, hermiticity helpers
    rng = np.random.default_rng(6001)
    L = 240
    t_lo, t_hi, _ = O.adjoint_real_mesh(-8.0, 8.0, L)
    g = rng.normal(size=(L, 2)) + 1j * rng.normal(size=(L, 2))
    m = rng.normal(size=(3, 2)) + 1j * rng.normal(size=(3, 2))
    m[0] = 0
    out["real"] = dict(meshes=np.array([t_lo, t_hi, -8.0, 8.0]), g=g, mom=m, t2w=O.fourier_t_to_w(g, t_lo, t_hi, -8.0, 8.0, m), w2t=O.fourier_w_to_t(g, t_lo, t_hi, -8.0, 8.0, m))
    gl = (rng.normal(size=(12, 2)) + 1j * rng.normal(size=(12, 2))) / (1 + np.arange(12))[:, None] ** 2
    gtl = O.legendre_to_imtime(gl, 3.0, 301)
    out["legendre"] = dict(beta=3.0, gl=gl, gw=O.legendre_to_imfreq(gl, O.FERMION, 40), gt=gtl, back=O.imtime_to_legendre(gtl, 3.0, 12))
    R = np.array([(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 0), (1, -1, 2)])
    hop = rng.normal(size=(6, 2, 2)) + 1j * rng.normal(size=(6, 2, 2))
    out["tb"] = dict(displ=R, hop=hop, dims=np.array([5, 4, 3]), eps=O.tight_binding_fourier(R, hop, (5, 4, 3)))
    gh = rng.normal(size=(40, 2, 2)) + 1j * rng.normal(size=(40, 2, 2))
    tail = rng.normal(size=(4, 2, 2)) + 1j * rng.normal(size=(4, 2, 2))
    out["herm"] = dict(g=gh, herm=O.make_hermitian(gh), tail=tail, replaced=O.replace_by_tail(gh, 7.0, O.FERMION, tail, 12))
    return out


def main():
    flat = {}
    for name, d in cases().items():
        for k, v in d.items():
            flat[f"{name}.{k}"] = np.asarray(v)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hotpath_v1.npz")
    np.savez_compressed(path, **flat)
    print("wrote", path, os.path.getsize(path), "bytes,", len(flat), "arrays")


if __name__ == "__main__":
    main()
