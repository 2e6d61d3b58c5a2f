"""Committed golden vectors (tests/golden/hotpath_v1.npz, written by tests/golden/make_golden.py):
CPU: the oracle still reproduces them;  GPU: every entry point of the C ABI matches them."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import triqs_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "hotpath_v1.npz"))
TOL = 1e-12


def rel(x, ref):
    x, ref = np.asarray(x), np.asarray(ref)
    return float(np.max(np.abs(x - ref)) / max(np.max(np.abs(ref)), 1e-300))


def test_oracle_reproduces_the_golden_vectors():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    n = 0
    for name, d in mod.cases().items():
        for k, v in d.items():
            v, ref = np.asarray(v), G[f"{name}.{k}"]
            if k.endswith("err"):            # residuals of an exact fit are round-off: BLAS / CPU dependent
                assert abs(float(v) - float(ref)) < 1e-10, (name, k)
            elif k in ("tail", "fit"):       # raw high-order moments are conditioning-limited
                assert v.shape == ref.shape and rel(v[:4], ref[:4]) < 1e-8, (name, k)
            elif name == "c2":               # quantities that use the tail fitted inside the transform
                assert rel(v, ref) < 1e-10, (name, k)
            else:
                assert rel(v, ref) < 1e-11, (name, k)
            n += 1
    assert n == len(G.files)


@pytest.fixture(scope="module")
def ctx():
    import triqs_b200
    c = triqs_b200.Context(0, print_warnings=False)
    yield c
    c.close()


@pytest.mark.gpu
def test_cuda_path_matches_the_golden_vectors(ctx):
    F, B = O.FERMION, O.BOSON
    # C1-like
    beta, n_iw = float(G["c1.beta"]), int(G["c1.n_iw"])
    gw = ctx.fourier_tau_to_iw(G["c1.gt"][None], beta, F, n_iw)
    assert rel(gw[0], G["c1.gw"]) < TOL
    tail, err = ctx.fit_tail(G["c1.gw"][None], beta, F, hermitian=True, inner_dim=1)
    assert tail.shape[1] == G["c1.tail"].shape[0]
    assert rel(tail[0][:4], G["c1.tail"][:4]) < 1e-9 and abs(err[0] - float(G["c1.err"])) < 1e-9
    # C2-like
    beta, n_tau = float(G["c2.beta"]), int(G["c2.n_tau"])
    n_iw = G["c2.gw"].shape[0] // 2
    gt = ctx.fourier_iw_to_tau(G["c2.gw"][None], beta, F, n_tau)
    assert rel(gt[0], G["c2.gt"]) < 1e-10  # tail fitted inside (conditioning-limited)
    assert rel(ctx.fourier_tau_to_iw(G["c2.gt"][None], beta, F, n_iw)[0], G["c2.gw_back"]) < TOL
    fit, ferr = ctx.fit_tail(G["c2.gw"][None], beta, F)
    assert rel(fit[0][:4], G["c2.fit"][:4]) < 1e-8
    assert rel(ctx.density(G["c2.gw"].reshape(1, -1, 5, 5), beta, F)[0], G["c2.density"]) < 1e-10
    # boson, known moments
    gt = ctx.fourier_iw_to_tau(G["boson.gw"][None], float(G["boson.beta"]), B, int(G["boson.n_tau"]), moments=G["boson.mom"][None])
    assert rel(gt[0], G["boson.gt"]) < TOL
    # lattice / plain DFT
    dims = tuple(int(x) for x in G["lattice.dims"])
    gk = ctx.fourier_lattice(G["lattice.gr"], dims, to_k=True)
    assert rel(gk, G["lattice.gk"]) < TOL
    assert rel(ctx.fourier_lattice(G["lattice.gk"], dims, to_k=False), G["lattice.back"]) < TOL
    assert rel(ctx.fft_many(G["lattice.gr"], dims, -1), G["lattice.fft"]) < TOL
    # Dyson, inversion
    gl = ctx.dyson_ksum(G["dyson.eps"], G["dyson.sigma"], float(G["dyson.beta"]), F, float(G["dyson.mu"]))
    assert rel(gl, G["dyson.gloc"]) < TOL
    a = np.ascontiguousarray(G["dyson.eps"] + 2 * np.eye(5)[None])
    assert rel(ctx.invert_in_place(a), G["dyson.inv"]) < TOL
    # interpolation: bit exact
    assert np.array_equal(ctx.interp_tau(G["interp.table"], float(G["interp.beta"]), G["interp.taus"]), G["interp.out"])
    # companions
    t_lo, t_hi, w_lo, w_hi = (float(x) for x in G["real.meshes"])
    assert rel(ctx.fourier_t_to_w(G["real.g"][None], (t_lo, t_hi), (w_lo, w_hi), moments=G["real.mom"][None])[0], G["real.t2w"]) < TOL
    assert rel(ctx.fourier_w_to_t(G["real.g"][None], (t_lo, t_hi), (w_lo, w_hi), moments=G["real.mom"][None])[0], G["real.w2t"]) < TOL
    beta = float(G["legendre.beta"])
    assert rel(ctx.legendre_to_imfreq(G["legendre.gl"][None], F, 40)[0], G["legendre.gw"]) < TOL
    assert rel(ctx.legendre_to_imtime(G["legendre.gl"][None], beta, 301)[0], G["legendre.gt"]) < TOL
    assert rel(ctx.imtime_to_legendre(G["legendre.gt"][None], beta, 12)[0], G["legendre.back"]) < TOL
    assert rel(ctx.tight_binding_fourier(G["tb.displ"], G["tb.hop"], tuple(int(x) for x in G["tb.dims"])), G["tb.eps"]) < TOL
    assert np.array_equal(ctx.make_hermitian(G["herm.g"][None])[0], G["herm.herm"])
    rep = np.ascontiguousarray(G["herm.g"].reshape(1, 40, 4).copy())
    ctx.replace_by_tail(rep, 7.0, F, G["herm.tail"].reshape(1, 4, 4), 12)
    assert rel(rep[0].reshape(40, 2, 2), G["herm.replaced"]) < TOL
