"""Parity of the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Tolerance (BASELINE.json north_star): 1e-12 relative in the max norm,  max|x - x_ref| / max|x_ref|,
the convention of c++/triqs/gfs/gf_tests.hpp:26-32.  Every test needs a B200 (-m gpu).
"""
import math

import numpy as np
import pytest

from oracle import triqs_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-12
F, B = O.FERMION, O.BOSON


def rel(x, ref):
    return float(np.max(np.abs(np.asarray(x) - ref)) / np.max(np.abs(ref)))


@pytest.fixture(scope="module")
def ctx():
    import triqs_b200
    c = triqs_b200.Context(0, print_warnings=False)
    yield c
    c.close()


def poles_gtau(beta, n_tau, E, r, stat=F):
    """G(tau) = sum_p r_p * pole(E_p) in the overflow-safe form of test/c++/gfs/fourier_matsubara.cpp:59-65."""
    tau = O.tau_values(beta, n_tau)
    s = -1.0 if stat == F else 1.0
    g = np.zeros(n_tau)
    for e, w in zip(E, r):
        if e > 0:
            g += w * (-np.exp(-e * tau) / (1 - s * math.exp(-e * beta)))
        else:
            g += w * (s * np.exp(e * (beta - tau)) / (1 - s * math.exp(e * beta)))
    return g.astype(np.complex128)


def c1_input(b):
    rng = np.random.default_rng(1001 + b)
    E = rng.uniform(-2, 2, 4)
    r = rng.dirichlet(np.ones(4))
    return poles_gtau(10.0, 10001, E, r)[:, None]


def matrix_gf(beta, n_iw, n, seed, stat=F):
    """G(iw) = (iw - H)^-1 with H random Hermitian, spectrum in [-2, 2]; returns (gw [n_w, n*n], H)."""
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
    h = a + a.conj().T
    w, v = np.linalg.eigh(h)
    w = -2 + 4 * (w - w.min()) / (w.max() - w.min())
    H = (v * w) @ v.conj().T
    iw = O.imfreq_values(beta, stat, n_iw)
    g = np.linalg.inv(iw[:, None, None] * np.eye(n)[None] - H[None])
    return g.reshape(len(iw), n * n), H


def matrix_gtau(beta, n_tau, H):
    w, v = np.linalg.eigh(H)
    tau = O.tau_values(beta, n_tau)
    f = np.where(w[None, :] > 0, -np.exp(-w[None, :] * tau[:, None]) / (1 + np.exp(-beta * np.abs(w[None, :]))),
                 -np.exp(-w[None, :] * (tau[:, None] - beta)) / (1 + np.exp(-beta * np.abs(w[None, :]))))
    g = np.einsum("ik,tk,jk->tij", v, f, v.conj())
    return g.reshape(n_tau, -1)


# ---------------------------------------------------------------------------------------------
# direct transform
# ---------------------------------------------------------------------------------------------
def test_c1_tau_to_iw_single_and_batch(ctx):
    """configs[0]: beta=10, n_tau=10001, n_iw=1025, scalar."""
    gts = np.stack([c1_input(b) for b in range(5)])
    gw = ctx.fourier_tau_to_iw(gts, 10.0, F, 1025)
    for b in range(5):
        ref = O.fourier_tau_to_iw(gts[b], 10.0, F, 1025)
        assert rel(gw[b], ref) < TOL
    gw0 = ctx.fourier_tau_to_iw(gts[:1], 10.0, F, 1025)
    assert np.array_equal(gw0[0], gw[0])  # batching does not change a bit


@pytest.mark.parametrize("stat", [F, B])
@pytest.mark.parametrize("ncols", [1, 4])
def test_reference_test_mesh_round_trip(ctx, stat, ncols):
    """test/c++/gfs/fourier_matsubara.cpp: beta=10, N_iw=1000, N_tau=10000 (L = 9999 = 3^2 * 11 * 101)."""
    beta, n_iw, n_tau, E = 10.0, 1000, 10000, -1.0
    iw = O.imfreq_values(beta, stat, n_iw)
    gw1 = 1 / (iw - E) + 1 / (iw + 2 * E) - 4.5 / (iw - 1.25 * E)
    gw = np.repeat(gw1[:, None], ncols, axis=1) * (1 + 0.25 * np.arange(ncols))[None, :]
    ref_t, info = O.fourier_iw_to_tau(gw, beta, stat, n_tau, return_info=True)
    # inverse with the oracle's fitted tail supplied: isolates the transform from the SVD
    gt = ctx.fourier_iw_to_tau(gw[None], beta, stat, n_tau, moments=info["tail"][None])[0]
    assert rel(gt, ref_t) < TOL
    # inverse with the library's own tail fit
    gt2 = ctx.fourier_iw_to_tau(gw[None], beta, stat, n_tau)[0]
    assert rel(gt2, ref_t) < TOL
    # direct, with the tau-derivative tail estimate
    ref_w = O.fourier_tau_to_iw(ref_t, beta, stat, n_iw)
    gwb = ctx.fourier_tau_to_iw(ref_t[None], beta, stat, n_iw)[0]
    assert rel(gwb, ref_w) < TOL
    assert np.max(np.abs(gwb - gw)) < 1e-8  # the reference test's own criterion
    # known moments (:74-83)
    km = np.zeros((2, ncols), complex)
    km[1] = -2.5 * (1 + 0.25 * np.arange(ncols))
    ref_km = O.fourier_iw_to_tau(gw, beta, stat, n_tau, km)
    gt_km = ctx.fourier_iw_to_tau(gw[None], beta, stat, n_tau, moments=km[None])[0]
    assert rel(gt_km, ref_km) < TOL


def test_default_python_mesh_6150(ctx):
    """n_iw = 1025 (Python default) -> n_tau = 6151, L = 6150 = 2 * 3 * 5^2 * 41 (hermiticity.cpp:34-48)."""
    beta, n_iw = 10.0, 1025
    gw, H = matrix_gf(beta, n_iw, 2, 11)
    n_tau = O.adjoint_n_tau(n_iw)
    ref_t, info = O.fourier_iw_to_tau(gw, beta, F, n_tau, return_info=True)
    gt = ctx.fourier_iw_to_tau(gw[None], beta, F, n_tau, moments=info["tail"][None])[0]
    assert rel(gt, ref_t) < TOL
    g3 = gt.reshape(-1, 2, 2)
    assert np.max(np.abs(g3 - g3.conj().transpose(0, 2, 1))) < 1e-12
    ref_w = O.fourier_tau_to_iw(ref_t, beta, F, n_iw)
    assert rel(ctx.fourier_tau_to_iw(ref_t[None], beta, F, n_iw)[0], ref_w) < TOL


def test_c2_blockgf_round_trip(ctx):
    """configs[1]: 2 blocks x 5x5, beta=100, n_tau=100001 (L=10^5, four-step path), n_iw=10000."""
    beta, n_tau, n_iw = 100.0, 100001, 10000
    gts, gws = [], []
    for b in range(2):
        gw, H = matrix_gf(beta, n_iw, 5, 2001 + b)
        gts.append(matrix_gtau(beta, n_tau, H))
        gws.append(gw)
    gts = np.stack(gts)
    out_w = ctx.fourier_tau_to_iw(gts, beta, F, n_iw)
    for b in range(2):
        ref_w = O.fourier_tau_to_iw(gts[b], beta, F, n_iw)
        assert rel(out_w[b], ref_w) < TOL
        assert np.max(np.abs(out_w[b] - gws[b])) < 1e-6
    # return leg on the transformed data, tail supplied by the oracle (transform parity) ...
    refs, tails = [], []
    for b in range(2):
        r, info = O.fourier_iw_to_tau(np.asarray(out_w[b]), beta, F, n_tau, return_info=True)
        refs.append(r)
        tails.append(info["tail"])
    out_t = ctx.fourier_iw_to_tau(out_w, beta, F, n_tau, moments=np.stack(tails))
    for b in range(2):
        assert rel(out_t[b], refs[b]) < TOL
    # ... and with the library's own least-squares tail (end-to-end)
    out_t2, err = ctx.fourier_iw_to_tau(out_w, beta, F, n_tau, return_err=True)
    for b in range(2):
        assert rel(out_t2[b], refs[b]) < TOL
        assert np.max(np.abs(out_t2[b] - gts[b])) < 1e-7
    assert np.all(err < 1e-4)


@pytest.mark.parametrize("stat,ncols,n_iw", [(F, 1, 10000), (B, 4, 10000), (F, 30, 7000), (B, 3, 1025), (F, 7, 50000), (F, 16, 10000)])
def test_long_mesh_pipelined_variants(ctx, stat, ncols, n_iw):
    """L = 10^5 goes through the TMA-pipelined four-step kernels (fourier_pipe.cu): both statistics, ragged column
    tiles, windows that do not align with the four-step factors, batch > 1."""
    beta, n_tau = 40.0, 100001
    rng = np.random.default_rng(ncols * 1000 + n_iw)
    iw = O.imfreq_values(beta, stat, n_iw)
    gws = []
    for b in range(2):
        E = rng.uniform(0.3, 2.0, (3, ncols)) * rng.choice([-1, 1], (3, ncols))
        r = rng.normal(size=(3, ncols)) + 1j * rng.normal(size=(3, ncols))
        gws.append(sum(r[p][None, :] / (iw[:, None] - E[p][None, :]) for p in range(3)))
    gws = np.stack(gws)
    refs, tails = [], []
    for b in range(2):
        rt, info = O.fourier_iw_to_tau(gws[b], beta, stat, n_tau, return_info=True)
        refs.append(rt)
        tails.append(info["tail"])
    gt = ctx.fourier_iw_to_tau(gws, beta, stat, n_tau, moments=np.stack(tails))
    for b in range(2):
        assert rel(gt[b], refs[b]) < TOL
    gw2 = ctx.fourier_tau_to_iw(np.stack(refs), beta, stat, n_iw)
    for b in range(2):
        assert rel(gw2[b], O.fourier_tau_to_iw(refs[b], beta, stat, n_iw)) < TOL


def test_long_mesh_chunked_launches_and_wide_targets(ctx):
    """Pipelined path with the batch split over several launch pairs (scratch budget below one batch) and with more
    target columns than one column tile holds (70 columns -> 5 ragged tiles in pass 1, more in pass 2)."""
    ctx.set_option("scratch_chunk_mb", 120)  # 70 columns x 1e5 x 16 B = 112 MB per gf -> one gf per launch pair
    beta, n_tau, n_iw, ncols, stat = 25.0, 100001, 10000, 70, F
    rng = np.random.default_rng(77)
    iw = O.imfreq_values(beta, stat, n_iw)
    gws = []
    for b in range(3):
        E = rng.uniform(0.3, 2.0, (2, ncols)) * rng.choice([-1, 1], (2, ncols))
        r = rng.normal(size=(2, ncols)) + 1j * rng.normal(size=(2, ncols))
        gws.append(sum(r[p][None, :] / (iw[:, None] - E[p][None, :]) for p in range(2)))
    gws = np.stack(gws)
    refs, tails = [], []
    for b in range(3):
        rt, info = O.fourier_iw_to_tau(gws[b], beta, stat, n_tau, return_info=True)
        refs.append(rt)
        tails.append(info["tail"])
    gt = ctx.fourier_iw_to_tau(gws, beta, stat, n_tau, moments=np.stack(tails))
    for b in range(3):
        assert rel(gt[b], refs[b]) < TOL
    gw2 = ctx.fourier_tau_to_iw(np.stack(refs), beta, stat, n_iw)
    ctx.set_option("scratch_chunk_mb", 0)
    for b in range(3):
        assert rel(gw2[b], O.fourier_tau_to_iw(refs[b], beta, stat, n_iw)) < TOL


def test_tau_to_iw_known_moments_and_noise_warning(ctx):
    import triqs_b200
    beta, n_tau, n_iw = 10.0, 1201, 200
    rng = np.random.default_rng(5)
    gt = poles_gtau(beta, n_tau, [0.7, -1.1], [0.4, 0.6])[:, None] + 1e-2 * rng.normal(size=(n_tau, 1))
    ref, info = O.fourier_tau_to_iw(gt, beta, F, n_iw, return_info=True)
    assert info["warn"] & O.WARN_NOISY_TAU
    out = ctx.fourier_tau_to_iw(gt[None], beta, F, n_iw)
    assert ctx.last_warn[0] & triqs_b200.TQ_WARN_NOISY_TAU
    assert rel(out[0], ref) < TOL
    km = np.array([[0.0], [1.0], [0.1], [0.3]], complex)
    ref = O.fourier_tau_to_iw(gt, beta, F, n_iw, km)
    assert rel(ctx.fourier_tau_to_iw(gt[None], beta, F, n_iw, moments=km[None])[0], ref) < TOL
    # short tau mesh warning (L < 6 n_iw)
    ctx.fourier_tau_to_iw(gt[None], beta, F, 300)
    assert ctx.last_warn[0] & triqs_b200.TQ_WARN_SHORT_TAU_MESH


def test_errors_match_reference(ctx):
    import triqs_b200
    gt = np.zeros((1, 101, 1), complex)
    with pytest.raises(triqs_b200.TriqsRuntimeError, match="at least twice as long"):
        ctx.fourier_tau_to_iw(gt, 1.0, F, 60)
    km = np.ones((1, 2, 1), complex)
    with pytest.raises(triqs_b200.TriqsRuntimeError, match="vanishing 0th moment"):
        ctx.fourier_tau_to_iw(np.zeros((1, 1001, 1), complex), 1.0, F, 50, moments=km)
    with pytest.raises(triqs_b200.TriqsRuntimeError, match="full mesh"):
        ctx.fourier_iw_to_tau(np.zeros((1, 101, 1), complex), 1.0, F, 1001)
    # G = iw cannot be fitted: error > 1e-2 (tail_issues.py:30-37 + fourier_matsubara.cpp:205-207)
    iw = O.imfreq_values(10.0, F, 1000)
    with pytest.raises(triqs_b200.TriqsRuntimeError, match="greater than 1e-2"):
        ctx.fourier_iw_to_tau(iw[None, :, None].copy(), 10.0, F, 6001)


# ---------------------------------------------------------------------------------------------
# tail fit
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("herm", [False, True])
@pytest.mark.parametrize("n_fixed", [0, 2])
def test_fit_tail_kernel_against_same_pinv(ctx, herm, n_fixed):
    """The batched kernel against NumPy using the SAME pseudo-inverse (the library's, exported by
    tq_tail_fitter_setup_host): pins gather, known-moment shift, GEMM, residual, rescale, symmetrisation."""
    import triqs_b200
    beta, n_iw, n = 10.0, 1025, 3
    gws = np.stack([matrix_gf(beta, n_iw, n, 40 + b)[0] for b in range(3)])
    known = None
    if n_fixed:
        known = np.zeros((3, 2, n * n), complex)
        known[:, 1] = np.eye(n).reshape(-1)
    s = triqs_b200.tail_fitter_setup(beta, F, n_iw, n_fixed=n_fixed, hermitian=herm)
    mom, err = ctx.fit_tail(gws, beta, F, known_moments=known, hermitian=herm, inner_dim=n)
    f = O.TailFitter(beta, F, n_iw)
    lss = f.setup_lss(n_fixed, herm)
    assert mom.shape[1] == s["n_var"] + n_fixed
    eps64 = np.finfo(float).eps
    for b in range(3):
        # oracle fit with the library's pinv swapped in
        lss2 = dict(lss)
        lss2["pinv"] = s["pinv"]
        f._lss[(n_fixed, herm)] = lss2
        ref, ref_err = f.fit(gws[b], None if known is None else known[b], hermitian=herm, inner_dim=n)
        # x = P B is a cancellation-heavy dot product (|P| ~ 1/sigma_min): two correct evaluations differ by
        # ~ eps * (|P| |B|), which is the bound used here (a wrong index or sign would be O(1) off)
        rows = np.asarray(f.fit_idx) - f.first
        Babs = np.abs(gws[b][rows, :]) + (np.abs(f.vander[:, :n_fixed]) @ np.abs(known[b]) if n_fixed else 0)
        if herm:
            Babs = np.vstack([Babs, Babs.reshape(len(rows), -1, n, n).transpose(0, 1, 3, 2).reshape(len(rows), -1)])
        bound = np.abs(s["pinv"]) @ Babs
        for v in range(s["n_var"]):
            p = n_fixed + v
            tol = 64 * eps64 * np.max(bound[v]) * f.w_max ** p + 1e-13 * np.max(np.abs(ref[p]))
            assert np.max(np.abs(mom[b, p] - ref[p])) < tol, (p,)
        for p in range(n_fixed):
            assert np.array_equal(mom[b, p], known[b, p])
        assert abs(err[b] - ref_err) < 1e-12 * np.max(np.abs(gws[b]))


@pytest.mark.parametrize("eps", [0.001, 0.1, 1, 10, 100])
def test_fit_hermitian_tail_exact_moments(ctx, eps):
    """test/python/base/tail_issues.py:74-102 through the GPU path."""
    Nw = int(max(25, eps * 60))
    H = eps * np.array([[1.0, 0.1j], [-0.1j, 2.0]])
    iw = O.imfreq_values(1.0, F, Nw)
    g = np.linalg.inv(iw[:, None, None] * np.eye(2)[None] - H[None]).reshape(1, -1, 4)
    mx = lambda x: np.max(np.abs(x))
    for km in (None, np.vstack([np.zeros(4), np.eye(2).reshape(4)]).astype(complex)[None]):
        tail, err = ctx.fit_tail(g, 1.0, F, known_moments=km, hermitian=True, inner_dim=2)
        for n in range(3):
            exact = np.linalg.matrix_power(H, n)
            assert mx(exact - tail[0, 1 + n].reshape(2, 2)) / mx(exact) < 1e-4
        ref, ref_err = O.fit_hermitian_tail(g[0], 1.0, F, None if km is None else km[0], inner_dim=2)
        # low moments agree with the LAPACK-based oracle far below the test's own 1e-4
        for n in range(1, 4):
            assert mx(tail[0, n] - ref[n]) / max(mx(ref[n]), 1e-300) < 2e-4


def test_fit_tail_reference_known_answers(ctx):
    """test/c++/gfs/fit_tail_matsubara.cpp:27-106 through the GPU path."""
    beta, N = 10.0, 100
    iw = O.imfreq_values(beta, F, N)
    c = np.array([0, 1, 3, 5], dtype=complex)
    gw = (c[0] + c[1] / iw + c[2] / iw / iw + c[3] / iw / iw / iw)[None, :, None]
    for km in (np.zeros((1, 1, 1), complex), np.array([[[0.0], [1.0]]], complex)):
        tail, err = ctx.fit_tail(gw, beta, F, known_moments=km, tail_fraction=0.3)
        assert np.max(np.abs(tail[0, :4, 0] - c)) < 1e-8
    km = np.array([[[0.0], [1.0]]], complex)
    for stat, n_iw in ((F, N), (B, N - 1)):
        iw = O.imfreq_values(beta, stat, n_iw)
        tail, err = ctx.fit_tail((1 / (iw - 1))[None, :, None], beta, stat, known_moments=km, tail_fraction=0.3)
        assert np.max(np.abs(tail[0, :5, 0] - np.array([0, 1, 1, 1, 1]))) < 1e-9


# ---------------------------------------------------------------------------------------------
# lattice
# ---------------------------------------------------------------------------------------------
def test_g_r_k_exact(ctx):
    """test/c++/gfs/g_r_k.cpp:52-58."""
    N = 10
    k = 2 * math.pi * np.arange(N) / N
    gk = (-2 * (np.cos(k)[:, None] + np.cos(k)[None, :])).reshape(N * N, 1).astype(complex)
    gr = ctx.fourier_lattice(gk, (N, N, 1), to_k=False).reshape(N, N)
    expect = np.zeros((N, N), complex)
    expect[1, 0] = expect[N - 1, 0] = expect[0, 1] = expect[0, N - 1] = -1
    assert np.max(np.abs(gr - expect)) < 1e-13
    assert np.max(np.abs(ctx.fourier_lattice(gr.reshape(-1, 1), (N, N, 1), to_k=True) - gk)) < 1e-12


@pytest.mark.parametrize("dims,n_others", [((64, 64, 1), 36), ((4, 3, 5), 7), ((16, 16, 16), 9), ((2, 2, 1), 1), ((7, 1, 11), 3),
                                           ((256, 256, 1), 18), ((128, 64, 1), 5), ((64, 64, 64), 3), ((256, 1, 128), 130)])
def test_lattice_vs_oracle(ctx, dims, n_others):
    """configs[2] shape family: data = complex normal * exp(-|r|_inf / 8), seed 3001."""
    rng = np.random.default_rng(3001)
    nk = int(np.prod(dims))
    x = rng.normal(size=(nk, n_others)) + 1j * rng.normal(size=(nk, n_others))
    idx = np.stack(np.unravel_index(np.arange(nk), dims), axis=1)
    dist = np.max(np.minimum(idx, np.array(dims)[None, :] - idx), axis=1)
    x *= np.exp(-dist / 8.0)[:, None]
    ref = O.fourier_lattice_r_to_k(x, dims)
    out = ctx.fourier_lattice(x, dims, to_k=True)
    assert rel(out, ref) < TOL
    back = ctx.fourier_lattice(out, dims, to_k=False)
    assert rel(back, x) < TOL
    assert rel(back, O.fourier_lattice_k_to_r(ref, dims)) < TOL


def test_fft_many_semantics(ctx):
    """_fourier_base: rank-1 over L rows of an (L+1)-row buffer is the caller's business; here plain rank 1..3."""
    rng = np.random.default_rng(9)
    for dims in [(12,), (6150,), (5, 8), (3, 4, 5)]:
        n = int(np.prod(dims))
        x = rng.normal(size=(n, 3)) + 1j * rng.normal(size=(n, 3))
        for sign in (1, -1):
            assert rel(ctx.fft_many(x, dims, sign), O.fft_many(x, dims, sign)) < TOL


# ---------------------------------------------------------------------------------------------
# inversion and the Dyson k-sum
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 8, 9, 12, 17, 32, 40, 64])
def test_invert_batched(ctx, n):
    """n <= 8: one thread per matrix in registers; 9..64: one warp per matrix in shared memory (nda::inverse_in_place has no limit)."""
    rng = np.random.default_rng(100 + n)
    a = rng.normal(size=(1000, n, n)) + 1j * rng.normal(size=(1000, n, n))
    ref = O.invert_batched(a)
    out = ctx.invert_in_place(a.copy())
    err = np.max(np.abs(out - ref), axis=(1, 2)) / np.max(np.abs(ref), axis=(1, 2))
    cond = np.linalg.cond(a)
    assert np.all(err < 1e-15 * cond * 50)
    assert np.median(err) < 1e-13 * max(1, n // 4)


def c4_inputs(nk1, n_iw, n=5, beta=10.0):
    """configs[3] family: NN tight-binding eps_k (seed 4001), 2-pole Sigma (seed 4002)."""
    rng = np.random.default_rng(4001)
    t = []
    for _ in range(3):
        m = 0.25 * (rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))) / math.sqrt(n)
        t.append(m)
    e0 = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
    e0 = 0.1 * (e0 + e0.conj().T)
    ks = 2 * math.pi * np.arange(nk1) / nk1
    kx, ky, kz = np.meshgrid(ks, ks, ks, indexing="ij")
    eps = np.broadcast_to(e0, (nk1, nk1, nk1, n, n)).copy()
    for d, kk in enumerate((kx, ky, kz)):
        ph = np.exp(1j * kk)[..., None, None]
        eps += ph * t[d] + ph.conj() * t[d].conj().T
    eps = eps.reshape(-1, n, n)
    rng = np.random.default_rng(4002)
    iw = O.imfreq_values(beta, F, n_iw)
    sig = np.zeros((len(iw), n, n), complex)
    for p in range(2):
        v = rng.normal(size=(n, 1)) + 1j * rng.normal(size=(n, 1))
        e = rng.uniform(-1, 1)
        sig += (v @ v.conj().T)[None] / (iw[:, None, None] - e)
    return eps, sig


def test_dyson_ksum_vs_oracle(ctx):
    eps, sig = c4_inputs(8, 64)
    ref = O.dyson_ksum(eps, sig, 10.0, F, 0.1)
    out = ctx.dyson_ksum(eps, sig, 10.0, F, 0.1)
    assert rel(out, ref) < TOL
    # explicit weights and a ragged k count
    w = np.random.default_rng(1).uniform(0.5, 1.5, 301)
    w /= w.sum()
    ref = O.dyson_ksum(eps[:301], sig, 10.0, F, 0.1, weights=w)
    assert rel(ctx.dyson_ksum(eps[:301], sig, 10.0, F, 0.1, weights=w), ref) < TOL


def test_dyson_other_dims(ctx):
    for n in (1, 2, 3, 9, 14):
        eps, sig = c4_inputs(4, 16, n=n)
        assert rel(ctx.dyson_ksum(eps, sig, 10.0, F, -0.3), O.dyson_ksum(eps, sig, 10.0, F, -0.3)) < TOL


def test_dyson_sharded_sum_equals_whole(ctx):
    """k-slices like mpi.slice_array (mpi_mpi4py.py:96-109), partial sums added = the all-reduce result."""
    eps, sig = c4_inputs(6, 32)
    nk = eps.shape[0]
    whole = ctx.dyson_ksum(eps, sig, 10.0, F, 0.1)
    acc = np.zeros_like(whole)
    for r in range(3):
        lo, hi = O.slice_bounds(nk, 3, r)
        acc += ctx.dyson_ksum(eps[lo:hi], sig, 10.0, F, 0.1, uniform_weight=1.0 / nk)
    assert rel(acc, whole) < 1e-14


# ---------------------------------------------------------------------------------------------
# interpolation
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,axes", [
    ((64, 48, 6), [("linear", 0.0, 10.0), ("linear", 0.0, 10.0)]),                       # g(tau, tau')
    ((12, 10, 8, 16, 4), ["periodic", "periodic", "periodic", "index"]),                # g(k, iw), gf<prod<brzone, imfreq>>
    ((40, 40, 1, 10, 1), ["periodic", "periodic", "periodic", "index"]),                # test/c++/gfs/multivar/g_k_om.cpp:76-114
    ((30, 9), [("linear", -2.0, 3.0)]),                                                  # one component
    ((7, 9, 5, 1), [("linear", 0.0, 1.0), "periodic", ("linear", -1.0, 1.0)]),
    ((5, 6, 7, 8, 2), [("linear", 0.0, 1.0), ("linear", 0.0, 2.0), ("linear", 0.0, 3.0), ("linear", 0.0, 4.0)]),
])
def test_evaluate_prod_bit_exact(ctx, shape, axes):
    """product-mesh evaluation (mesh/evaluate.hpp:97-114): same roundings as the restated recursion, so bit-for-bit"""
    rng = np.random.default_rng(len(shape) * 100 + shape[0])
    table = rng.normal(size=shape) + 1j * rng.normal(size=shape)
    n_pts = 20011
    cols = []
    for d, a in enumerate(axes):
        if a == "periodic":
            cols.append(rng.uniform(-2.5 * shape[d], 2.5 * shape[d], n_pts))   # any real index, wrapped (brzone.hpp:355-359)
        elif a == "index":
            cols.append(rng.integers(0, shape[d], n_pts).astype(float))
        else:
            x = rng.uniform(a[1], a[2], n_pts)
            x[:3] = [a[1], a[2], 0.5 * (a[1] + a[2])]                              # end points (linear.hpp:224-227 clamps)
            cols.append(x)
    coords = np.stack(cols, axis=1)
    ref = O.evaluate_prod(table, axes, coords)
    out = ctx.evaluate_prod(table, axes, coords)
    assert np.array_equal(np.asarray(out), ref)
    import torch
    td = torch.from_numpy(table).cuda()
    out_d = ctx.evaluate_prod(td, axes, torch.from_numpy(coords).cuda())
    assert np.array_equal(out_d.cpu().numpy(), ref)


def test_interp_tau(ctx):
    beta, n_tau = 10.0, 8192
    _, H = matrix_gf(beta, 8, 5, 77)
    table = matrix_gtau(beta, n_tau, H)
    rng = np.random.default_rng(5001)
    taus = np.concatenate([rng.uniform(0, beta, 100000), [0.0, beta, beta / 2, 1e-300, beta - 1e-13]])
    ref = O.interp_tau(table, beta, taus)
    out = ctx.interp_tau(table, beta, taus)
    assert np.array_equal(out, ref)  # same operations in the same order: bit exact
    # test/c++/meshes/eval.cpp:62-89
    tab = np.array([[1.0], [10.0], [100.0], [1000.0], [1e4]], complex)
    o = ctx.interp_tau(tab, 10.0, np.array([2.0, 3.5]))
    assert np.allclose(o[:, 0], [0.2 * 1 + 0.8 * 10, 0.6 * 10 + 0.4 * 100], rtol=1e-14)


# ---------------------------------------------------------------------------------------------
# device-resident (torch CUDA tensor) path gives the same bits as the host-staged path
# ---------------------------------------------------------------------------------------------
def test_device_pointers_same_result(ctx):
    import torch
    gts = np.stack([c1_input(b) for b in range(3)])
    host = ctx.fourier_tau_to_iw(gts, 10.0, F, 1025)
    dev = ctx.fourier_tau_to_iw(torch.from_numpy(gts).cuda(), 10.0, F, 1025)
    assert dev.is_cuda
    assert np.array_equal(dev.cpu().numpy(), host)


# ---------------------------------------------------------------------------------------------
# density
# ---------------------------------------------------------------------------------------------
def test_density_reference_known_answers(ctx):
    """test/c++/gfs/gf_density.cpp:22-72 and gf_base.cpp:69 through the C ABI."""
    beta, n_iw = 1.0, 400
    iw = O.imfreq_values(beta, F, n_iw)
    g = (1 / (iw + 2.3)).reshape(1, -1, 1, 1)
    assert abs(ctx.density(g, beta, F)[0, 0, 0] - 1 / (1 + math.exp(-beta * 2.3))) < 1e-9
    iwb = O.imfreq_values(beta, B, n_iw)
    gb = (1 / (iwb - 2.3)).reshape(1, -1, 1, 1)
    assert abs(ctx.density(gb, beta, B)[0, 0, 0] - 1 / (-1 + math.exp(beta * 2.3))) < 1e-9
    iw3 = O.imfreq_values(1.0, F, 100)
    g3 = np.zeros((1, 200, 2, 2), complex)
    g3[0, :, 0, 0] = g3[0, :, 1, 1] = 2 / (iw3 + 2.3)
    assert np.max(np.abs(ctx.density(g3, 1.0, F)[0] - np.diag([1.8177540779781256] * 2))) < 1e-7
    import triqs_b200
    with pytest.raises(triqs_b200.TriqsRuntimeError):  # gf_density.cpp:66-72
        ctx.density(np.ones((1, 800, 1, 1), complex), beta, F)
    with pytest.raises(triqs_b200.TriqsRuntimeError):
        ctx.density(iw.reshape(1, -1, 1, 1).copy(), beta, F)


@pytest.mark.parametrize("stat,n,n_iw", [(F, 5, 1025), (B, 3, 700), (F, 1, 10000), (F, 8, 300)])
def test_density_vs_oracle(ctx, stat, n, n_iw):
    """Batch of matrix-valued gfs (not Hermitian on purpose: only the upper triangle may be used), fitted and known tails."""
    rng = np.random.default_rng(100 * n + stat)
    beta = 12.0
    iw = O.imfreq_values(beta, stat, n_iw)
    gs, kms = [], []
    for b in range(3):
        E = rng.uniform(0.4, 2.0, (2, n, n)) * rng.choice([-1, 1], (2, n, n)) if stat == F else rng.uniform(0.4, 2.0, (2, n, n))
        r = rng.normal(size=(2, n, n)) + 1j * rng.normal(size=(2, n, n))
        gs.append(sum(r[p][None] / (iw[:, None, None] - E[p][None]) for p in range(2)))
        kms.append(np.stack([np.zeros((n, n)), r.sum(0), (r * E).sum(0), (r * E * E).sum(0)]))
    gs, kms = np.stack(gs), np.stack(kms)
    out = ctx.density(gs, beta, stat)
    out_km = ctx.density(gs, beta, stat, moments=kms)
    out_k2 = ctx.density(gs, beta, stat, moments=kms[:, :2])
    for b in range(3):
        ref_km = O.density(gs[b], beta, stat, kms[b])
        assert rel(out_km[b], ref_km) < TOL
        # fitted tails: the transform is pinned to 1e-12 above, the LSQ moments to cond * eps
        assert rel(out[b], O.density(gs[b], beta, stat)) < 1e-9
        assert rel(out_k2[b], O.density(gs[b], beta, stat, kms[b][:2])) < 1e-9
        iu = np.triu_indices(n, 1)
        assert np.array_equal(out_km[b].T[iu], out_km[b][iu].conj())  # lower triangle = conjugate of the upper one (:115)


# ---------------------------------------------------------------------------------------------
# hermiticity helpers, replace_by_tail
# ---------------------------------------------------------------------------------------------
def test_hermiticity_helpers(ctx):
    """hermiticity.cpp:25-80 through the C ABI + parity with the oracle on non-symmetric data."""
    beta, n_iw = 1.0, 1025
    iw = O.imfreq_values(beta, F, n_iw)
    h = np.array([[1, 1j], [-1j, 2]], complex)
    g = np.linalg.inv(iw[:, None, None] * np.eye(2) - h)[None]
    assert ctx.is_gf_hermitian(g)
    assert np.max(np.abs(ctx.make_hermitian(g) - g)) < 1e-15
    hb = np.array([[1, 1j], [1j, 2]], complex)
    gb = np.linalg.inv(iw[:, None, None] * np.eye(2) - hb)[None]
    assert not ctx.is_gf_hermitian(gb)
    assert ctx.is_gf_hermitian(ctx.make_hermitian(gb))
    rng = np.random.default_rng(3)
    for npts, d, imtime in [(2050, 3, False), (1999, 4, False), (1001, 2, True), (64, 1, False)]:
        x = rng.normal(size=(3, npts, d, d)) + 1j * rng.normal(size=(3, npts, d, d))
        out = ctx.make_hermitian(x, imtime=imtime)
        for b in range(3):
            assert np.array_equal(out[b], O.make_hermitian(x[b], imtime=imtime))  # one add and one multiply by 0.5: bit exact
        assert ctx.is_gf_hermitian(out, imtime=imtime, tolerance=1e-15)
        assert not ctx.is_gf_hermitian(x, imtime=imtime)


@pytest.mark.parametrize("stat", [F, B])
def test_replace_by_tail(ctx, stat):
    beta, n_iw, ncols = 10.0, 300, 4
    rng = np.random.default_rng(12 + stat)
    n_w = 2 * n_iw - (0 if stat == F else 1)
    g = rng.normal(size=(2, n_w, ncols)) + 1j * rng.normal(size=(2, n_w, ncols))
    tail = rng.normal(size=(2, 6, ncols)) + 1j * rng.normal(size=(2, 6, ncols))
    ref = np.stack([O.replace_by_tail(g[b], beta, stat, tail[b], 120) for b in range(2)])
    out = ctx.replace_by_tail(g.copy(), beta, stat, tail, 120)
    assert rel(out, ref) < TOL
    first, _ = O.imfreq_first_last(stat, n_iw)
    n = first + np.arange(n_w)
    inside = (n < 120) & (n >= -120)
    assert np.array_equal(out[:, inside], g[:, inside])


def test_tight_binding_fourier(ctx):
    """tight_binding::fourier(k_mesh): known square-lattice dispersion and a random 3-D multi-orbital model vs the oracle."""
    displ = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0)]
    e = ctx.tight_binding_fourier(displ, -np.ones((4, 1, 1), complex), (10, 10, 1)).reshape(10, 10)
    k = 2 * np.pi * np.arange(10) / 10
    assert np.max(np.abs(e - (-2 * (np.cos(k)[:, None] + np.cos(k)[None, :])))) < 1e-14
    rng = np.random.default_rng(4001)
    R = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1), (0, 0, 0), (2, -1, 3), (-2, 1, -3)]
    t = rng.normal(size=(9, 5, 5)) + 1j * rng.normal(size=(9, 5, 5))
    for dims in [(16, 12, 10), (64, 64, 1), (7, 1, 5)]:
        out = ctx.tight_binding_fourier(R, t, dims)
        assert rel(out, O.tight_binding_fourier(R, t, dims)) < TOL


def test_fourier_real_time_frequency(ctx):
    """retime <-> refreq (fourier_real.cpp): the reference test's meshes and known answers (fourier_real.cpp:27-87), random data
    and moments against the oracle, both directions, batch > 1, lengths with large prime factors."""
    w_max, L, E = 40.0, 1001, 1.0
    t_lo, t_hi, _ = O.adjoint_real_mesh(-w_max, w_max, L)
    w, t = O.linear_values(-w_max, w_max, L), O.linear_values(t_lo, t_hi, L)
    gw = np.repeat((2 * E / (w * w + E * E)).astype(complex)[:, None], 4, axis=1)[None]
    tail = np.zeros((1, 3, 4), complex)
    tail[0, 2] = 2 * E
    gt = ctx.fourier_w_to_t(gw, (t_lo, t_hi), (-w_max, w_max), moments=tail)
    assert np.max(np.abs(gt[0] - np.exp(-E * np.abs(t))[:, None])) < 1e-4
    assert rel(gt[0], O.fourier_w_to_t(gw[0], t_lo, t_hi, -w_max, w_max, tail[0])) < TOL
    back = ctx.fourier_t_to_w(gt, (t_lo, t_hi), (-w_max, w_max), moments=tail)
    assert np.max(np.abs(back - gw)) < 1e-4
    assert rel(back[0], O.fourier_t_to_w(gt[0], t_lo, t_hi, -w_max, w_max, tail[0])) < TOL
    rng = np.random.default_rng(77)
    for (L, C, B, half) in [(1001, 8, 3, False), (600, 1, 2, True), (2048, 25, 1, False), (97, 5, 2, False)]:
        w_lo, w_hi = -12.5, 12.5
        if half:
            t_lo, t_hi, _ = O.adjoint_real_mesh(w_lo, w_hi, L, shift_half_bin=True)
        else:
            t_lo, t_hi, _ = O.adjoint_real_mesh(w_lo, w_hi, L)
        g = rng.normal(size=(B, L, C)) + 1j * rng.normal(size=(B, L, C))
        mom = rng.normal(size=(B, 3, C)) + 1j * rng.normal(size=(B, 3, C))
        mom[:, 0] = 0
        for m in (mom, None):
            f = ctx.fourier_t_to_w(g, (t_lo, t_hi), (w_lo, w_hi), moments=m)
            i = ctx.fourier_w_to_t(g, (t_lo, t_hi), (w_lo, w_hi), moments=m)
            for b in range(B):
                mb = None if m is None else m[b]
                assert rel(f[b], O.fourier_t_to_w(g[b], t_lo, t_hi, w_lo, w_hi, mb)) < TOL
                assert rel(i[b], O.fourier_w_to_t(g[b], t_lo, t_hi, w_lo, w_hi, mb)) < TOL
    # error behaviour of the reference: incompatible meshes, too few moments, non-vanishing 0th moment
    import triqs_b200
    g = np.zeros((1, 64, 2), complex)
    t_lo, t_hi, _ = O.adjoint_real_mesh(-5.0, 5.0, 64)
    with pytest.raises(triqs_b200.TriqsRuntimeError, match="not compatible"):
        ctx.fourier_t_to_w(g, (t_lo, 1.01 * t_hi), (-5.0, 5.0))
    with pytest.raises(triqs_b200.TriqsRuntimeError, match="order 2"):
        ctx.fourier_t_to_w(g, (t_lo, t_hi), (-5.0, 5.0), moments=np.zeros((1, 2, 2), complex))
    with pytest.raises(triqs_b200.TriqsRuntimeError, match="vanishing 0th moment"):
        ctx.fourier_w_to_t(g, (t_lo, t_hi), (-5.0, 5.0), moments=np.ones((1, 3, 2), complex))


def test_legendre_matsubara(ctx):
    """Legendre <-> Matsubara (legendre_matsubara.hpp): known answers of gf_legendre.cpp:22-72 and parity with the oracle for
    all four transforms; imfreq -> legendre goes through the prime-length (L = 49999) direct transform."""
    beta, n_l, n_tau = 2.0, 10, 5
    g = np.zeros((1, n_l, 1), complex)
    g[0, 0] = 1
    assert np.max(np.abs(ctx.legendre_to_imtime(g, beta, n_tau) - 1 / beta)) < 1e-16
    g[:] = 0
    g[0, 2] = 1
    x = 2 * O.tau_values(beta, n_tau) / beta - 1
    assert np.max(np.abs(ctx.legendre_to_imtime(g, beta, n_tau)[0, :, 0] - math.sqrt(5) / beta * (3 * x * x - 1) / 2)) < 1e-15
    rng = np.random.default_rng(31)
    B, C, n_l = 2, 6, 40
    gl = (rng.normal(size=(B, n_l, C)) + 1j * rng.normal(size=(B, n_l, C))) / (1 + np.arange(n_l))[None, :, None] ** 2
    for stat, n_iw in ((O.FERMION, 300), (O.BOSON, 129)):
        gw = ctx.legendre_to_imfreq(gl, stat, n_iw)
        for b in range(B):
            assert rel(gw[b], O.legendre_to_imfreq(gl[b], stat, n_iw)) < TOL
    for n_tau in (1001, 4097, 64):
        gt = ctx.legendre_to_imtime(gl, 7.0, n_tau)
        back = ctx.imtime_to_legendre(gt, 7.0, n_l)
        for b in range(B):
            ref = O.legendre_to_imtime(gl[b], 7.0, n_tau)
            assert rel(gt[b], ref) < TOL
            assert rel(back[b], O.imtime_to_legendre(ref, 7.0, n_l)) < TOL
    # imfreq -> legendre of a physical G (tail fitted inside, 50000 tau points)
    beta, n_iw = 10.0, 400
    iw = O.imfreq_values(beta, O.FERMION, n_iw)
    H = np.array([[0.7, 0.2], [0.2, -0.4]])
    gw = np.linalg.inv(iw[:, None, None] * np.eye(2)[None] - H[None]).reshape(1, -1, 4)
    got = ctx.imfreq_to_legendre(gw, beta, O.FERMION, 30)
    ref = O.imfreq_to_legendre(gw[0], beta, O.FERMION, 30)
    assert rel(got[0], ref) < 1e-10  # the LSQ tail fit inside is conditioning-limited
    with pytest.raises(Exception, match="128 Legendre"):
        ctx.legendre_to_imtime(np.zeros((1, 200, 1), complex), 1.0, 10)


def test_matsubara_meshes_with_large_prime_factors(ctx):
    """time meshes whose length the tile FFT cannot factor (FFTW takes any length, fourier_common.cpp:30): direct window DFT"""
    rng = np.random.default_rng(12)
    for (L, n_iw, stat, C, B) in [(6 * 1031, 1031, O.FERMION, 4, 2), (4999, 500, O.BOSON, 3, 1), (49999, 64, O.FERMION, 9, 1)]:
        beta = 5.0
        iw = O.imfreq_values(beta, stat, n_iw)
        e = rng.uniform(0.3, 2.0, size=(B, C))
        gw = 1 / (iw[None, :, None] - e[:, None, :])
        mom = np.zeros((B, 4, C), complex)
        mom[:, 1], mom[:, 2], mom[:, 3] = 1, e, e ** 2
        gt = ctx.fourier_iw_to_tau(gw, beta, stat, L + 1, moments=mom)
        back = ctx.fourier_tau_to_iw(gt, beta, stat, n_iw, moments=mom)
        for b in range(B):
            rt = O.fourier_iw_to_tau(gw[b], beta, stat, L + 1, mom[b])
            assert rel(gt[b], rt) < TOL
            assert rel(back[b], O.fourier_tau_to_iw(rt, beta, stat, n_iw, mom[b])) < TOL


@pytest.mark.parametrize("L,n_iw,stat,C,B", [(1999, 300, O.FERMION, 5, 3), (4999, 833, O.BOSON, 3, 2), (8191, 1024, O.FERMION, 25, 2),
                                            (19999, 2000, O.FERMION, 4, 1), (2 * 10007, 1500, O.BOSON, 2, 2)])
def test_bluestein_for_lengths_without_a_plan(ctx, L, n_iw, stat, C, B):
    """Round n_tau values (2000, 5000, 8192, 20000 ...) are one off a smooth length and have a prime factor > 127: chirp-z (Bluestein)
    convolution through the plain axis engine instead of the O(N^2) window DFT.  Checked against the oracle (1e-12) and against the
    direct evaluation ("no_bluestein"), both legs, known moments and the library's own tail."""
    rng = np.random.default_rng(L)
    beta = 7.0
    iw = O.imfreq_values(beta, stat, n_iw)
    e = rng.uniform(0.3, 2.0, size=(B, C)) * rng.choice([-1.0, 1.0], size=(B, C))
    gw = 1 / (iw[None, :, None] - e[:, None, :])
    mom = np.zeros((B, 4, C), complex)
    mom[:, 1], mom[:, 2], mom[:, 3] = 1, e, e ** 2
    ctx.profile_enable(True)
    gt = ctx.fourier_iw_to_tau(gw, beta, stat, L + 1, moments=mom)
    back = ctx.fourier_tau_to_iw(gt, beta, stat, n_iw, moments=mom)
    rep = ctx.profile_report()
    ctx.profile_enable(False)
    assert "iw2tau_bluestein_pre" in rep and "tau2iw_bluestein_post" in rep, rep
    for b in range(B):
        rt = O.fourier_iw_to_tau(gw[b], beta, stat, L + 1, mom[b])
        assert rel(gt[b], rt) < TOL
        assert rel(back[b], O.fourier_tau_to_iw(rt, beta, stat, n_iw, mom[b])) < TOL
    ctx.set_option("no_bluestein", 1)
    try:
        gt_d = ctx.fourier_iw_to_tau(gw, beta, stat, L + 1, moments=mom)
        back_d = ctx.fourier_tau_to_iw(gt, beta, stat, n_iw, moments=mom)
    finally:
        ctx.set_option("no_bluestein", 0)
    assert rel(gt, gt_d) < TOL and rel(back, back_d) < TOL
    ctx.set_option("no_pipe", 1)   # the axis transforms through the generic tile kernel, the pointwise products in passes of their own
    try:
        gt_g = ctx.fourier_iw_to_tau(gw, beta, stat, L + 1, moments=mom)
        back_g = ctx.fourier_tau_to_iw(gt, beta, stat, n_iw, moments=mom)
    finally:
        ctx.set_option("no_pipe", 0)
    assert rel(gt, gt_g) < TOL and rel(back, back_g) < TOL
    # unknown moments: tau-derivative estimate forward, least-squares tail on the way back
    w2 = ctx.fourier_tau_to_iw(gt, beta, stat, n_iw)
    assert rel(w2[0], O.fourier_tau_to_iw(np.asarray(gt[0]), beta, stat, n_iw)) < TOL


def test_fft_many_bluestein_axes(ctx):
    """_fourier_base takes any length (fourier_common.cpp:30-44): plain axes with a prime factor > 127 run as a Bluestein convolution
    (N >= 512; shorter ones by direct summation), also beyond the 65536 points the direct sum was limited to.  Against numpy's FFT."""
    rng = np.random.default_rng(5)
    for dims, hm in [((1009,), 7), ((2 * 1013, 3), 5), ((4, 4999), 25), ((70001,), 3), ((8191,), 1)]:
        n = int(np.prod(dims))
        x = rng.normal(size=(n, hm)) + 1j * rng.normal(size=(n, hm))
        for sign in (+1, -1):
            y = ctx.fft_many(x, dims, sign)
            xr = x.reshape(tuple(dims) + (hm,))
            axes = tuple(range(len(dims)))
            ref = (np.fft.ifftn(xr, axes=axes) * n if sign > 0 else np.fft.fftn(xr, axes=axes)).reshape(n, hm)
            assert rel(y, ref) < TOL, (dims, hm, sign)
    ctx.profile_enable(True)
    ctx.fft_many(rng.normal(size=(1009, 4)) + 0j, (1009,), 1)
    assert "fft_axis_bluestein" in ctx.profile_report()
    ctx.profile_enable(False)


def test_tau_to_iw_along_inner_mesh_is_one_gf(ctx):
    """tq_fourier_tau_to_iw_axis = _fourier<N>, N > 0 (fourier.hpp:78-84): [outer][n_tau][inner] is ONE gf with outer*inner
    columns; the noise check (fourier_matsubara.cpp:100-114) is joint.  Column (0, 0) is G = 1/(i w) whose G(tau) = -1/2 is
    flat: on its own its differences are pure round-off and the per-gf heuristic would drop the tail."""
    beta, n_iw, n_tau, outer, inner = 1.0, 100, 201, 5, 3
    iw = O.imfreq_values(beta, O.FERMION, n_iw)
    e = np.linspace(0.0, 2.0, outer * inner).reshape(outer, inner)
    gw = 1 / (iw[None, :, None] - e[:, None, :])
    gt = np.stack([O.fourier_iw_to_tau(gw[o], beta, O.FERMION, n_tau) for o in range(outer)])
    flat = np.ascontiguousarray(gt.transpose(1, 0, 2).reshape(n_tau, outer * inner))   # flatten_2d<N>
    ref = O.fourier_tau_to_iw(flat, beta, O.FERMION, n_iw).reshape(2 * n_iw, outer, inner).transpose(1, 0, 2)
    got = ctx.fourier_tau_to_iw(gt, beta, O.FERMION, n_iw, one_gf=True)
    assert rel(got, ref) < TOL
    assert np.max(np.abs(got - gw)) < 1e-4
    assert not (ctx.last_warn[0] & 1)
    # per-gf semantics (independent gfs): the flat one is declared noisy, like the reference called on that gf alone
    sep = ctx.fourier_tau_to_iw(gt, beta, O.FERMION, n_iw)
    assert rel(sep[0], O.fourier_tau_to_iw(gt[0], beta, O.FERMION, n_iw)) < TOL


def test_fit_tail_real_frequency(ctx):
    """fit_tail on a refreq mesh: test/c++/gfs/fit_tail_real.cpp:24-91 (Basic, Complex, the k x w multivar case as a batch)
    and parity with the oracle's fitter."""
    L, om = 201, 10.0
    w = O.linear_values(-om, om, L)
    c = np.array([0, 1, 3, 5], complex)
    g = (c[0] + c[1] / (w + 1e-14) + c[2] / (w * w + 1e-14) + c[3] / (w * w * w + 1e-14)).astype(complex)[None, :, None]
    for known in (np.zeros((1, 1, 1), complex), np.array([[[0.0], [1.0]]], complex)):
        tail, err = ctx.fit_tail_refreq(g, (-om, om), known_moments=known, tail_fraction=0.3)
        assert np.max(np.abs(tail[0, :4, 0] - c)) < 1e-9
    om = 100.0
    w = O.linear_values(-om, om, L)
    a = 1.0 + 0.4j
    known = np.array([[[0.0], [1.0]]], complex)
    tail, err = ctx.fit_tail_refreq((1 / (w - a))[None, :, None], (-om, om), known_moments=known)
    assert np.max(np.abs(tail[0, :5, 0] - np.array([0, 1, a, a ** 2, a ** 3]))) < 1e-7
    # Multivar (:95-130): g(k, w) = 1 / (w - cos kx cos ky - i eta), one fit per k-point = batch
    kk = 2 * np.pi * np.arange(4) / 4
    eps = (np.cos(kk)[:, None] * np.cos(kk)[None, :]).reshape(-1)
    gk = 1 / (w[None, :, None] - eps[:, None, None] - 0.001j)
    known = np.zeros((16, 2, 1), complex)
    known[:, 1] = 1.0
    tail, err = ctx.fit_tail_refreq(gk, (-om, om), known_moments=known)
    for k in range(16):
        ex = np.array([0, 1, eps[k] + 0.001j, (eps[k] + 0.001j) ** 2])
        assert np.max(np.abs(tail[k, :4, 0] - ex)) < 1e-6
        rt, re = O.fit_tail_refreq(gk[k], -om, om, known[k])
        assert tail.shape[1] == rt.shape[0]
        assert np.max(np.abs(tail[k, :4] - rt[:4])) < 1e-9
        assert abs(err[k] - re) < 1e-9 + 1e-6 * re


def test_plain_and_lattice_lengths_the_tile_fft_cannot_factor(ctx):
    """FFTW takes any length (fourier_common.cpp:30): axis lengths with a prime factor > 128, or longer than one SM's shared
    memory, go through the direct evaluation -- plain DFT (both signs, in place), lattice meshes, real-time meshes."""
    rng = np.random.default_rng(91)
    for (dims, hm) in [((131,), 5), ((2 * 139, 3), 9), ((7, 149, 2), 4), ((20011,), 2)]:
        n = int(np.prod(dims))
        x = rng.normal(size=(n, hm)) + 1j * rng.normal(size=(n, hm))
        for sign in (1, -1):
            assert rel(ctx.fft_many(x, dims, sign), O.fft_many(x, dims, sign)) < TOL
        y = x.copy()
        ctx.fft_many(y, dims, 1, out=y)  # in place
        assert rel(y, O.fft_many(x, dims, 1)) < TOL
    dims = (131, 4, 1)
    g = rng.normal(size=(131 * 4, 6)) + 1j * rng.normal(size=(131 * 4, 6))
    gk = ctx.fourier_lattice(g, dims, to_k=True)
    assert rel(gk, O.fourier_lattice_r_to_k(g, dims)) < TOL
    assert rel(ctx.fourier_lattice(gk, dims, to_k=False), g) < TOL
    # retime <-> refreq on a prime-length mesh
    L, C = 1009, 3
    w_lo, w_hi = -9.0, 9.0
    t_lo, t_hi, _ = O.adjoint_real_mesh(w_lo, w_hi, L)
    g = rng.normal(size=(2, L, C)) + 1j * rng.normal(size=(2, L, C))
    mom = rng.normal(size=(2, 3, C)) + 1j * rng.normal(size=(2, 3, C))
    mom[:, 0] = 0
    f = ctx.fourier_t_to_w(g, (t_lo, t_hi), (w_lo, w_hi), moments=mom)
    i = ctx.fourier_w_to_t(g, (t_lo, t_hi), (w_lo, w_hi), moments=mom)
    for b in range(2):
        assert rel(f[b], O.fourier_t_to_w(g[b], t_lo, t_hi, w_lo, w_hi, mom[b])) < TOL
        assert rel(i[b], O.fourier_w_to_t(g[b], t_lo, t_hi, w_lo, w_hi, mom[b])) < TOL


def test_matsubara_transforms_randomised_shapes(ctx):
    """Seeded sweep over mesh lengths (single pass, generic four-step, direct), statistics, column counts, batch sizes and
    known / fitted moments: both directions against the oracle."""
    rng = np.random.default_rng(20261019)
    lengths = [96, 210, 1000, 1024, 1500, 3125, 4097, 6144, 12000, 14406, 30000, 2 * 3 * 5 * 7 * 11 * 13]
    for case in range(18):
        L = int(rng.choice(lengths))
        stat = int(rng.integers(0, 2))
        n_iw = int(rng.integers(3, max(4, min(L // 2, 600))))
        C = int(rng.choice([1, 2, 3, 5, 9, 16, 25, 36]))
        B = int(rng.choice([1, 2, 5]))
        beta = float(rng.choice([1.0, 10.0, 50.0]))
        iw = O.imfreq_values(beta, stat, n_iw)
        e = rng.uniform(0.2, 2.0, size=(B, 3, C)) * rng.choice([-1.0, 1.0], size=(B, 3, C))
        r = rng.uniform(0.2, 1.0, size=(B, 3, C))
        gw = np.sum(r[:, None] / (iw[None, :, None, None] - e[:, None]), axis=2)
        known = case % 2 == 0
        mom = None
        if known:
            mom = np.zeros((B, 4, C), complex)
            for p in range(1, 4):
                mom[:, p] = np.sum(r * e ** (p - 1), axis=1)
        try:
            refs = [O.fourier_iw_to_tau(gw[b], beta, stat, L + 1, None if mom is None else mom[b]) for b in range(B)]
        except Exception:  # e.g. a window too short for the tail fit: the reference throws, and so must the library
            with pytest.raises(Exception):
                ctx.fourier_iw_to_tau(gw, beta, stat, L + 1, moments=mom)
            continue
        gt = ctx.fourier_iw_to_tau(gw, beta, stat, L + 1, moments=mom)
        back = ctx.fourier_tau_to_iw(gt, beta, stat, n_iw, moments=mom)
        for b in range(B):
            rt = refs[b]
            tol = TOL if known else 1e-10  # the LSQ tail fit inside is conditioning-limited
            assert rel(gt[b], rt) < tol, (case, L, stat, n_iw, C, B, known)
            rb = O.fourier_tau_to_iw(rt, beta, stat, n_iw, None if mom is None else mom[b])
            assert rel(back[b], rb) < tol, (case, L, stat, n_iw, C, B, known)


def test_host_buffer_calls_overlap_lanes_do_not_change_a_bit(ctx):
    """Host-pointer calls are cut into chunks over internal lanes (H2D || kernels || D2H); every gf is independent, so the
    result must be bit-identical to the single-stream call, warnings and tail errors land in the right slots."""
    beta, n_tau, n_iw, ncols = 10.0, 1201, 200, 3
    rng = np.random.default_rng(9)
    gts = []
    for b in range(7):
        E = rng.uniform(-2, 2, 3)
        gts.append(np.stack([poles_gtau(beta, n_tau, E * (1 + 0.1 * c), [0.2, 0.3, 0.5])[:] for c in range(ncols)], axis=1))
    gts = np.stack(gts)
    gts[4] += 1e-2 * rng.normal(size=gts[4].shape)   # one noisy gf: its warning bit must stay at index 4
    ctx.set_option("host_lanes", 1)
    w1 = ctx.fourier_tau_to_iw(gts, beta, F, n_iw)
    warn1 = np.array(ctx.last_warn)
    t1, e1 = ctx.fourier_iw_to_tau(w1, beta, F, n_tau, return_err=True)
    ctx.set_option("host_lanes", 3)
    ctx.set_option("host_chunk_kb", 2 * n_tau * ncols * 16 // 1024 + 1)   # two gfs per chunk -> 4 chunks over 3 lanes
    w3 = ctx.fourier_tau_to_iw(gts, beta, F, n_iw)
    warn3 = np.array(ctx.last_warn)
    t3, e3 = ctx.fourier_iw_to_tau(w3, beta, F, n_tau, return_err=True)
    ctx.set_option("host_chunk_kb", 48 << 10)
    assert np.array_equal(w1, w3) and np.array_equal(t1, t3)
    assert np.array_equal(warn1, warn3) and warn1[4] & 1 and not warn1[0] & 1
    assert np.array_equal(e1, e3)
    for b in (0, 6):
        assert rel(w3[b], O.fourier_tau_to_iw(gts[b], beta, F, n_iw)) < TOL


def test_pageable_buffers_through_the_lanes_pinned_slots(ctx):
    """NumPy (pageable) buffers above 12 MiB are staged piecewise through pinned slots by a few helper threads (tq_stage_in /
    tq_stage_out_finish, stage_parallel): same bits as the plain single-stream copy ("host_lanes" = 1), whatever the helper count."""
    beta, n_tau, n_iw, ncols = 10.0, 120001, 3000, 5     # 9.6 MB per gf: two pieces each way
    rng = np.random.default_rng(21)
    iw = O.imfreq_values(beta, F, n_iw)
    gws = np.stack([sum((rng.normal(size=ncols) + 1j * rng.normal(size=ncols))[None, :] / (iw[:, None] - e) for e in rng.uniform(-2, 2, 3))
                    for _ in range(5)])
    ctx.set_option("host_lanes", 1)
    t1 = ctx.fourier_iw_to_tau(gws, beta, F, n_tau)
    w1 = ctx.fourier_tau_to_iw(t1, beta, F, n_iw)
    ctx.set_option("host_lanes", 3)
    ctx.set_option("host_chunk_kb", 10 << 10)            # one gf per chunk -> 5 chunks over 3 lanes
    t3 = ctx.fourier_iw_to_tau(gws, beta, F, n_tau)
    w3 = ctx.fourier_tau_to_iw(t3, beta, F, n_iw)
    ctx.set_option("host_chunk_kb", 48 << 10)
    assert np.array_equal(t1, t3) and np.array_equal(w1, w3)
    for helpers in (1, 3, 16):
        ctx.set_option("stage_threads", helpers)
        assert np.array_equal(ctx.fourier_iw_to_tau(gws, beta, F, n_tau), t1)
        assert np.array_equal(ctx.fourier_tau_to_iw(t1, beta, F, n_iw), w1)
    ctx.set_option("stage_threads", 0)
    assert rel(t3[4], O.fourier_iw_to_tau(gws[4], beta, F, n_tau)) < TOL


def test_real_valued_gtau_moves_reals_only_and_matches_the_promoted_call(ctx):
    """Real-valued G(tau) (gf<imtime, T::real_t>): the reference promotes it to complex before the transform (fourier.hpp:78-81).
    tq_fourier_tau_to_iw_real takes the float64 buffer: with 8+ columns two real columns travel as ONE complex column ("two for one",
    separated through G(i w_-n) = conj G(i w_n)): same result as the promoted call to rounding; otherwise (or with "no_two_for_one") the columns are
    promoted on the device: same BITS as promoting on the host.  tq_fourier_iw_to_tau_real hands back real(g) with max |Im| per gf
    (functions2.hpp:223-246).  Host buffers (chunked over the lanes), device tensors, odd and even column counts, both statistics, known moments."""
    import torch
    beta, n_iw = 10.0, 200
    rng = np.random.default_rng(33)
    for n_tau, ncols, B, stat in [(1201, 5, 7, F), (6151, 4, 3, F), (10001, 1, 9, F), (6001, 25, 3, F), (4097, 16, 2, O.BOSON), (2000, 9, 2, F)]:
        gts = []
        for b in range(B):
            E = rng.uniform(-2, 2, 3)
            gts.append(np.stack([poles_gtau(beta, n_tau, E * (1 + 0.1 * c), [0.2, 0.3, 0.5], stat)[:] for c in range(ncols)], axis=1))
        gc = np.stack(gts)
        assert np.max(np.abs(gc.imag)) == 0.0
        gr = np.ascontiguousarray(gc.real)
        w_c = ctx.fourier_tau_to_iw(gc, beta, stat, n_iw)
        warn_c = np.array(ctx.last_warn)
        ctx.set_option("host_chunk_kb", 2 * n_tau * ncols * 8 // 1024 + 1)  # two gfs per chunk
        w_r = ctx.fourier_tau_to_iw(gr, beta, stat, n_iw)
        ctx.set_option("host_chunk_kb", 48 << 10)
        assert np.array_equal(warn_c, np.array(ctx.last_warn))
        if ncols >= 8:   # two for one
            assert rel(w_r, w_c) < 1e-13
        else:
            assert np.array_equal(w_c, w_r)
        ctx.set_option("no_two_for_one", 1)
        assert np.array_equal(ctx.fourier_tau_to_iw(gr, beta, stat, n_iw), w_c)
        ctx.set_option("no_two_for_one", 0)
        assert rel(w_r[0], O.fourier_tau_to_iw(gc[0], beta, stat, n_iw)) < TOL
        # known moments: real ones keep the packed path, complex ones (not the moments of a real G) fall back to the promoted one
        mom = np.zeros((B, 4, ncols), complex)
        mom[:, 1] = 1.0
        mom[:, 2] = rng.normal(size=(B, ncols))
        mom[:, 3] = rng.normal(size=(B, ncols))
        assert rel(ctx.fourier_tau_to_iw(gr, beta, stat, n_iw, moments=mom), ctx.fourier_tau_to_iw(gc, beta, stat, n_iw, moments=mom)) < 1e-13
        momc = mom + 1e-3j * rng.normal(size=mom.shape) * (np.arange(4)[None, :, None] > 1)
        assert np.array_equal(ctx.fourier_tau_to_iw(gr, beta, stat, n_iw, moments=momc), ctx.fourier_tau_to_iw(gc, beta, stat, n_iw, moments=momc))
        # device-resident real input
        w_d = ctx.fourier_tau_to_iw(torch.from_numpy(gr).cuda(), beta, stat, n_iw)
        assert np.array_equal(w_d.cpu().numpy(), w_r)
        # and back: real part + max |Im|
        t_c = ctx.fourier_iw_to_tau(w_c, beta, stat, n_tau)
        t_r = ctx.fourier_iw_to_tau(w_c, beta, stat, n_tau, real_out=True)
        assert t_r.dtype == np.float64 and np.array_equal(t_r, t_c.real)
        assert np.array_equal(ctx.last_max_imag, np.max(np.abs(t_c.imag), axis=(1, 2)))
        assert np.max(np.abs(t_r - gr)) < 5e-6  # (round trip: limited by the tail fit, like the complex path)
        out = torch.empty(B, n_tau, ncols, dtype=torch.float64, device="cuda")
        ctx.fourier_iw_to_tau(torch.from_numpy(w_c).cuda(), beta, stat, n_tau, out=out)
        assert np.array_equal(out.cpu().numpy(), t_r)
    # the zeroth-moment assertion (:116-118) survives the packing
    bad = np.zeros((1, 4, 9), complex)
    bad[:, 0, 3] = 1e-3
    bad[:, 1] = 1.0
    with pytest.raises(Exception, match="0th moment"):
        ctx.fourier_tau_to_iw(np.zeros((1, 1201, 9)), beta, F, n_iw, moments=bad)


def test_gf_mirror_real_valued_gf(ctx):
    """Gf(..., is_real=True) / real(g) / is_gf_real(g): python/triqs/gf/gf.py `is_real`, functions2.hpp:223-248."""
    from triqs_b200.gf import Gf, MeshImTime, make_gf_from_fourier, is_gf_real, real
    beta, n_tau = 10.0, 1201
    m = MeshImTime(beta, F, n_tau)
    g = Gf(m, (2, 2), ctx=ctx, is_real=True)
    assert g.data.dtype == np.float64
    for i in range(2):
        for j in range(2):
            g.data[:, i, j] = poles_gtau(beta, n_tau, [0.3 + i, -0.7 * (j + 1)], [0.4, 0.6]).real * (1.0 if i == j else 0.1)
    gw = make_gf_from_fourier(g)
    gc = Gf(m, (2, 2), g.data.astype(complex), ctx)
    assert np.array_equal(gw.data, make_gf_from_fourier(gc).data)
    back = make_gf_from_fourier(gw)
    assert is_gf_real(back, 1e-9) and not is_gf_real(gw)
    rb = real(back)
    assert rb.data.dtype == np.float64 and np.max(np.abs(rb.data - g.data)) < 1e-6


def test_large_pageable_buffers_of_the_other_entry_points(ctx):
    """tq_stage_in / tq_stage_out_finish spread every large copy from / to pageable memory over helper threads (tq_ctx.cu, stage_parallel): lattice,
    plain DFT, interpolation and k-sum calls with NumPy buffers above 12 MiB give the bits of the plain single-stream copies ("host_lanes" = 1)."""
    rng = np.random.default_rng(2)

    def both(fn):
        ctx.set_option("host_lanes", 1)
        a = fn()
        ctx.set_option("host_lanes", 3)
        return a, fn()
    x = rng.normal(size=(128 * 128, 96)) + 1j * rng.normal(size=(128 * 128, 96))          # 25 MB each way
    a, b = both(lambda: ctx.fourier_lattice(x, (128, 128, 1), True))
    assert np.array_equal(a, b) and rel(a, O.fourier_lattice_r_to_k(x, (128, 128, 1))) < TOL
    a, b = both(lambda: ctx.fft_many(x, (128, 128), -1))
    assert np.array_equal(a, b)
    table = rng.normal(size=(4096, 4)) + 1j * rng.normal(size=(4096, 4))
    taus = rng.uniform(0, 10.0, size=2_000_000)                                            # 16 MB in, 128 MB out
    a, b = both(lambda: ctx.interp_tau(table, 10.0, taus))
    assert np.array_equal(a, b)
    eps = rng.normal(size=(40000, 5, 5)) + 1j * rng.normal(size=(40000, 5, 5))            # 16 MB
    eps = eps + eps.conj().transpose(0, 2, 1)
    sig = np.stack([np.eye(5) / (w - 0.5) for w in O.imfreq_values(10.0, F, 16)])
    a, b = both(lambda: ctx.dyson_ksum(eps, sig, 10.0, F, 0.1))
    assert np.array_equal(a, b)
