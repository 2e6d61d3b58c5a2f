"""Pins the CPU oracle against the known answers the reference's own tests assert
(SURVEY.md section 4 / section 8c).  No GPU, no reference tree needed at run time."""
import math

import numpy as np
import pytest

from oracle import triqs_oracle as O

F, B = O.FERMION, O.BOSON


def one_pole_tau(En, tau, beta, stat):
    """closed form used by test/c++/gfs/fourier_matsubara.cpp:56-66."""
    s = -1.0 if stat == F else 1.0
    if En > 0:
        return -np.exp(-En * tau) / (1 - s * math.exp(-En * beta))
    return s * np.exp(En * (beta - tau)) / (1 - s * math.exp(En * beta))


@pytest.mark.parametrize("stat", [F, B])
@pytest.mark.parametrize("ncols", [1, 4])
def test_fourier_matsubara_cpp(stat, ncols):
    """test/c++/gfs/fourier_matsubara.cpp:23-114: beta=10, N_iw=1000, N_tau=10000, E=-1."""
    beta, n_iw, n_tau, E = 10.0, 1000, 10000, -1.0
    iw = O.imfreq_values(beta, stat, n_iw)
    gw1 = 1 / (iw - E) + 1 / (iw + 2 * E) - 4.5 / (iw - 1.25 * E)
    gw = np.repeat(gw1[:, None], ncols, axis=1)
    gt = O.fourier_iw_to_tau(gw, beta, stat, n_tau)
    gwb = O.fourier_tau_to_iw(gt, beta, stat, n_iw)
    assert np.max(np.abs(gw - gwb)) < 1e-8
    tau = O.tau_values(beta, n_tau)
    exact = one_pole_tau(E, tau, beta, stat) + one_pole_tau(-2 * E, tau, beta, stat) - 4.5 * one_pole_tau(1.25 * E, tau, beta, stat)
    assert np.max(np.abs(gt - exact[:, None])) < 1e-8
    # tail from fit_tail passed explicitly (:68-72)
    tail, err = O.fit_tail(gw, beta, stat)
    gt_tail = O.fourier_iw_to_tau(gw, beta, stat, n_tau, tail)
    assert np.max(np.abs(gt_tail - exact[:, None])) < 1e-8
    # only 0th and 1st moment known (:74-83)
    km = np.zeros((2, ncols), dtype=complex)
    km[1] = -2.5
    gt_km = O.fourier_iw_to_tau(gw, beta, stat, n_tau, km)
    assert np.max(np.abs(gt_km - exact[:, None])) < 1e-8
    # real-valued G(tau) input (:91-96)
    gwc = O.fourier_tau_to_iw(np.repeat(exact[:, None], ncols, axis=1), beta, stat, n_iw)
    assert np.max(np.abs(gw - gwc)) < 1e-8
    # with a self energy, then 10 repeated round trips (:98-113)
    sig = 1 / (iw - 2) + 3 / (iw + 3)
    gw2 = np.repeat((1 / (iw - E - sig))[:, None], ncols, axis=1)
    g = gw2
    for _ in range(11):
        g = O.fourier_tau_to_iw(O.fourier_iw_to_tau(g, beta, stat, n_tau), beta, stat, n_iw)
    assert np.max(np.abs(g - gw2)) < 1e-8


def test_fourier_positive_only_throws():
    """test/c++/gfs/fourier_matsubara.cpp:127-136 (a positive-only mesh has odd/half size)."""
    gw = np.zeros((1000 + 1, 1), dtype=complex)
    with pytest.raises(O.TriqsRuntimeError):
        O.fourier_iw_to_tau(gw, 10.0, F, 6001)


def test_fit_tail_basic():
    """test/c++/gfs/fit_tail_matsubara.cpp:27-72: c={0,1,3,5}, tail_fraction 0.3, 1e-8."""
    beta, N = 10.0, 100
    iw = O.imfreq_values(beta, F, N)
    c = np.array([0, 1, 3, 5], dtype=complex)
    gw = (c[0] + c[1] / iw + c[2] / iw / iw + c[3] / iw / iw / iw)[:, None]
    for km in (np.zeros((1, 1), complex), np.array([[0.0], [1.0]], complex)):
        tail, err = O.fit_tail(gw, beta, F, km, tail_fraction=0.3)
        assert np.max(np.abs(tail[:4, 0] - c)) < 1e-8


def test_fit_tail_real_f_and_b():
    """fit_tail_matsubara.cpp:76-106: 1/(iw-1) moments {0,1,1,1,1}, 1e-9, Fermion and Boson."""
    beta, N = 10.0, 100
    km = np.array([[0.0], [1.0]], complex)
    for stat, n_iw in ((F, N), (B, N - 1)):
        iw = O.imfreq_values(beta, stat, n_iw)
        tail, err = O.fit_tail((1 / (iw - 1))[:, None], beta, stat, km, tail_fraction=0.3)
        assert np.max(np.abs(tail[:5, 0] - np.array([0, 1, 1, 1, 1]))) < 1e-9


def test_fit_tail_complex():
    """fit_tail_matsubara.cpp:110-130: complex pole a=1+0.4i, N=200, 1e-7."""
    beta, N = 10.0, 200
    a = 1.0 + 0.4j
    iw = O.imfreq_values(beta, F, N)
    tail, err = O.fit_tail((1 / (iw - a))[:, None], beta, F, np.array([[0.0], [1.0]], complex))
    assert np.max(np.abs(tail[:5, 0] - np.array([0, 1, a, a ** 2, a ** 3]))) < 1e-7


@pytest.mark.parametrize("eps", [0.001, 0.01, 0.1, 1, 10, 100])
def test_fit_hermitian_tail_exact_moments(eps):
    """test/python/base/tail_issues.py:74-102."""
    Nw = int(max(25, eps * 60))
    beta = 1.0
    H = eps * np.array([[1.0, 0.1j], [-0.1j, 2.0]])
    iw = O.imfreq_values(beta, F, Nw)
    g = np.linalg.inv(iw[:, None, None] * np.eye(2)[None] - H[None]).reshape(-1, 4)
    mx = lambda x: np.max(np.abs(x))
    for km in (None, np.vstack([np.zeros(4), np.eye(2).reshape(4)]).astype(complex)):
        tail, err = O.fit_hermitian_tail(g, beta, F, km, inner_dim=2)
        for n in range(3):
            exact = np.linalg.matrix_power(H, n)
            assert mx(exact - tail[1 + n].reshape(2, 2)) / mx(exact) < 1e-4
        # result is hermitian moment by moment
        for m in tail:
            mm = m.reshape(2, 2)
            assert mx(mm - mm.conj().T) < 1e-12 * max(1.0, mx(mm))


def test_fit_iomega_n_error_large():
    """tail_issues.py:30-37: fitting G = iw gives err > 1e-2."""
    iw = O.imfreq_values(10.0, F, 1000)
    g = np.stack([iw, 0 * iw, 0 * iw, iw], axis=1)
    tail, err = O.fit_tail(g, 10.0, F)
    assert err > 1e-2


@pytest.mark.parametrize("eps", [0.001, 1, 100])
def test_multi_fft(eps):
    """tail_issues.py:41-72 (10 iterations instead of 100 to keep the CPU suite short)."""
    Nw = int(max(25, eps * 60))
    Ntau = 6 * Nw + 1
    beta = 1.0
    iw = O.imfreq_values(beta, F, Nw)
    g_ref = np.linalg.inv((iw[:, None, None] + eps) * np.eye(2)[None]).reshape(-1, 4)
    g = g_ref.copy()
    fitter = O.TailFitter(beta, F, Nw)
    for _ in range(10):
        gt = O.fourier_iw_to_tau(g, beta, F, Ntau, fitter=fitter)
        g = O.fourier_tau_to_iw(gt, beta, F, Nw)
        assert np.max(np.abs(g - g_ref)) < 1e-9
        assert np.linalg.norm(g[0].reshape(2, 2) - g[-1].reshape(2, 2).conj().T) < 1e-9


def test_imag_gt_semicircular():
    """tail_issues.py:104-112: Im G(tau) < 1e-11 for a semicircular DOS, beta=50, n=2000."""
    beta, n = 50.0, 2000
    iw = O.imfreq_values(beta, F, n)
    D = 1.0
    om = iw.imag
    # SemiCircular(1): G(iw) = 2 (iw - i sign(w) sqrt(w^2 + D^2)) / D^2
    g = (2 * (iw - 1j * np.sign(om) * np.sqrt(om ** 2 + D ** 2)) / D ** 2)[:, None]
    gt = O.fourier_iw_to_tau(g, beta, F, O.adjoint_n_tau(n))
    assert np.max(np.abs(gt.imag)) < 1e-11


def test_hermiticity_default_meshes():
    """test/c++/gfs/functions/hermiticity.cpp:34-48: n_iw=1025 -> n_tau=6151; G(iw)=G(-iw)^dagger kept to 1e-12."""
    beta, n_iw = 10.0, 1025
    assert O.adjoint_n_tau(n_iw) == 6151
    iw = O.imfreq_values(beta, F, n_iw)
    H = np.array([[1.0, 1j], [-1j, 2.0]])
    g = np.linalg.inv(iw[:, None, None] * np.eye(2)[None] - H[None]).reshape(-1, 4)
    gt = O.fourier_iw_to_tau(g, beta, F, 6151)
    gt2 = gt.reshape(-1, 2, 2)
    assert np.max(np.abs(gt2 - gt2.conj().transpose(0, 2, 1))) < 1e-12


def test_g_r_k_exact():
    """test/c++/gfs/g_r_k.cpp:28-62: FT of -2(cos kx + cos ky) on 10x10 is -1 at r=(+-1,0),(0,+-1)."""
    N = 10
    k = 2 * math.pi * np.arange(N) / N
    gk = (-2 * (np.cos(k)[:, None] + np.cos(k)[None, :])).reshape(N * N, 1).astype(complex)
    gr = O.fourier_lattice_k_to_r(gk, (N, N, 1)).reshape(N, N)
    expect = np.zeros((N, N), complex)
    expect[1, 0] = expect[N - 1, 0] = expect[0, 1] = expect[0, N - 1] = -1
    assert np.max(np.abs(gr - expect)) < 1e-13
    back = O.fourier_lattice_r_to_k(gr.reshape(N * N, 1), (N, N, 1))
    assert np.max(np.abs(back - gk)) < 1e-13


def test_lattice_round_trip_3d():
    """test/c++/gfs/fourier_lattice.cpp:27-62 style round trip, plus sign convention on a plane wave."""
    rng = np.random.default_rng(7)
    dims = (4, 3, 5)
    x = rng.normal(size=(60, 3)) + 1j * rng.normal(size=(60, 3))
    assert np.max(np.abs(O.fourier_lattice_k_to_r(O.fourier_lattice_r_to_k(x, dims), dims) - x)) < 1e-13
    # delta at r=(1,0,0) -> e^{+2 pi i k0/d0}
    d = np.zeros((60, 1), complex)
    d[1 * 15] = 1
    gk = O.fourier_lattice_r_to_k(d, dims).reshape(dims)
    assert np.allclose(gk[:, 0, 0], np.exp(2j * math.pi * np.arange(4) / 4))


def test_interp_weights():
    """test/c++/meshes/eval.cpp:62-89: imtime(beta=10, 5 pts): tau=2 -> .2 g0 + .8 g1; 3.5 -> .6 g1 + .4 g2."""
    tab = np.array([[1.0], [10.0], [100.0], [1000.0], [1e4]], complex)
    out = O.interp_tau(tab, 10.0, np.array([2.0, 3.5, 10.0, 0.0]))
    assert np.allclose(out[:, 0], [0.2 * 1 + 0.8 * 10, 0.6 * 10 + 0.4 * 100, 1e4, 1.0], rtol=1e-14)


def test_tail_fit_indices_c1():
    """n_iw=1025 fermion -> 60 rows, first -1025 step 6.8333, second start 819."""
    f = O.TailFitter(10.0, F, 1025)
    assert len(f.fit_idx) == 60
    assert f.fit_idx[0] == -1025 and f.fit_idx[1] == 1024 - 205
    assert f.fit_idx[2] == int(-1025 + 205 / 30)


def test_dyson_small_against_direct():
    rng = np.random.default_rng(3)
    n, nk, n_iw = 3, 17, 8
    eps = rng.normal(size=(nk, n, n)) + 1j * rng.normal(size=(nk, n, n))
    eps = eps + eps.conj().transpose(0, 2, 1)
    iw = O.imfreq_values(10.0, F, n_iw)
    sig = np.stack([np.eye(n) / (w - 0.5) for w in iw])
    G = O.dyson_ksum(eps, sig, 10.0, F, 0.1, k_chunk=5)
    for wi, w in enumerate(iw):
        ref = sum(np.linalg.inv((w + 0.1) * np.eye(n) - e - sig[wi]) for e in eps) / nk
        assert np.max(np.abs(G[wi] - ref)) < 1e-13


def test_slice_bounds_matches_mpi_rule():
    """python/triqs/utility/mpi_mpi4py.py:96-109."""
    n, size = 11, 4
    pieces = [O.slice_bounds(n, size, r) for r in range(size)]
    assert pieces == [(0, 3), (3, 6), (6, 9), (9, 11)]


# ---------------------------------------------------------------------------------------------
# density  (test/c++/gfs/gf_density.cpp:22-72, gf_base.cpp:55-69)
# ---------------------------------------------------------------------------------------------
def test_density_reference_known_answers():
    beta, n_iw = 1.0, 400
    iw = O.imfreq_values(beta, O.FERMION, n_iw)
    g = (1 / (iw + 2.3)).reshape(-1, 1, 1)
    assert abs(O.density(g, beta, O.FERMION)[0, 0] - 1 / (1 + math.exp(-beta * 2.3))) < 1e-9  # gf_density.cpp:35
    iwb = O.imfreq_values(beta, O.BOSON, n_iw)
    gb = (1 / (iwb - 2.3)).reshape(-1, 1, 1)
    assert abs(O.density(gb, beta, O.BOSON)[0, 0] - 1 / (-1 + math.exp(beta * 2.3))) < 1e-9  # gf_density.cpp:53
    # gf_base.cpp:40-69: G3 = 2 / (iw + 2.3) on a 2x2 diagonal target, beta = 1, 100 frequencies
    iw3 = O.imfreq_values(1.0, O.FERMION, 100)
    g3 = np.zeros((200, 2, 2), complex)
    g3[:, 0, 0] = g3[:, 1, 1] = 2 / (iw3 + 2.3)
    n = O.density(g3, 1.0, O.FERMION)
    assert np.max(np.abs(n - np.diag([1.8177540779781256] * 2))) < 1e-7
    # gf_density.cpp:66-72: a constant / a linear function have no usable tail -> runtime_error
    with pytest.raises(O.TriqsRuntimeError):
        O.density(np.ones((800, 1, 1), complex), beta, O.FERMION)
    with pytest.raises(O.TriqsRuntimeError):
        O.density(iw.reshape(-1, 1, 1).copy(), beta, O.FERMION)


def test_density_matrix_hermitian_and_known_moments():
    """(iw - H)^-1 with Hermitian H: n = f(H) (Fermi function of the matrix); supplying exact moments changes nothing."""
    rng = np.random.default_rng(5)
    beta, n_iw, n = 8.0, 600, 3
    a = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
    H = (a + a.conj().T) / 2
    iw = O.imfreq_values(beta, O.FERMION, n_iw)
    g = np.linalg.inv(iw[:, None, None] * np.eye(n) - H)
    w, v = np.linalg.eigh(H)
    exact = (v * (1 / (1 + np.exp(beta * w)))) @ v.conj().T
    d = O.density(g, beta, O.FERMION)
    # the reference returns G(tau = 0^-) = <c^dag_j c_i>, the transpose of f(H)_ij ... for Hermitian H: conj
    assert np.max(np.abs(d - exact)) < 2e-6 or np.max(np.abs(d - exact.conj())) < 2e-6
    km = np.stack([np.zeros((n, n)), np.eye(n), H, H @ H]).astype(complex)
    d2 = O.density(g, beta, O.FERMION, km)
    assert np.max(np.abs(d2 - d)) < 2e-6


# ---------------------------------------------------------------------------------------------
# hermiticity helpers, replace_by_tail  (test/c++/gfs/functions/hermiticity.cpp:25-80, fit_and_replace.cpp:30-72)
# ---------------------------------------------------------------------------------------------
def test_hermiticity_reference_known_answers():
    beta, n_iw = 1.0, 1025  # gf<imfreq>{{beta, Fermion}, {2, 2}}: default n_iw
    iw = O.imfreq_values(beta, O.FERMION, n_iw)
    h = np.array([[1, 1j], [-1j, 2]], complex)
    g = np.linalg.inv(iw[:, None, None] * np.eye(2) - h)
    assert O.is_gf_hermitian(g)                                            # hermiticity.cpp:38
    assert np.max(np.abs(O.make_hermitian(g) - g)) < 1e-15                 # :41
    hb = np.array([[1, 1j], [1j, 2]], complex)                             # :49 broken hermiticity
    gb = np.linalg.inv(iw[:, None, None] * np.eye(2) - hb)
    assert not O.is_gf_hermitian(gb)                                       # :52
    assert O.is_gf_hermitian(O.make_hermitian(gb))                         # :55
    # tensor_valued<4> case (:63-69) viewed as a 4 x 4 (ij),(kl) matrix
    idx = np.arange(2)
    i, j, k, l = np.meshgrid(idx, idx, idx, idx, indexing="ij")
    chi = 1.0 / (iw[:, None, None, None, None] - (1 + 1j) * (i * j)[None] - (1 - 1j) * (k * l)[None])
    chi4 = chi.reshape(-1, 4, 4)
    assert O.is_gf_hermitian(chi4)
    assert np.max(np.abs(O.make_hermitian(chi4) - chi4)) < 1e-15
    # imtime: G(tau) of the hermitian h is hermitian point by point
    w, v = np.linalg.eigh(h)
    tau = O.tau_values(beta, 201)
    f = -np.exp(-w[None, :] * tau[:, None]) / (1 + np.exp(-beta * w[None, :]))
    gt = np.einsum("ik,tk,jk->tij", v, f, v.conj())
    assert O.is_gf_hermitian(gt, imtime=True)
    assert np.max(np.abs(O.make_hermitian(gt, imtime=True) - gt)) < 1e-12


def test_replace_by_tail_reference_behaviour():
    """fit_and_replace.cpp:42-72 pattern: outside +-n_min the gf is replaced by its tail expansion; inside it is untouched."""
    beta, n_iw = 10.0, 100
    iw = O.imfreq_values(beta, O.FERMION, n_iw)
    g = (1 / (iw - 1.0) + 1 / (iw + 2.0)).reshape(-1, 1)
    tail = np.array([[0.0], [2.0], [-1.0], [5.0], [-7.0], [17.0]], complex)  # moments of the two poles: sum E^(n-1)
    out = O.replace_by_tail(g, beta, O.FERMION, tail, 50)
    first, _ = O.imfreq_first_last(O.FERMION, n_iw)
    n = first + np.arange(2 * n_iw)
    inside = (n < 50) & (n >= -50)
    assert np.array_equal(out[inside], g[inside])
    assert np.max(np.abs(out[~inside] - g[~inside])) < 1e-6 and np.max(np.abs(out[~inside] - g[~inside])) > 0


def test_tight_binding_fourier_square_lattice():
    """Nearest-neighbour hopping t = -1 on the square lattice: eps_k = -2 (cos kx + cos ky)  (the dispersion of g_r_k.cpp:40-58)."""
    displ = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0)]
    hop = -np.ones((4, 1, 1), complex)
    N = 10
    e = O.tight_binding_fourier(displ, hop, (N, N, 1)).reshape(N, N)
    k = 2 * np.pi * np.arange(N) / N
    assert np.max(np.abs(e - (-2 * (np.cos(k)[:, None] + np.cos(k)[None, :])))) < 1e-14
    # and it is the lattice Fourier transform (cyclat -> brzone) of the hopping placed on the real-space mesh
    gr = np.zeros((N, N), complex)
    gr[1, 0] = gr[N - 1, 0] = gr[0, 1] = gr[0, N - 1] = -1
    assert np.max(np.abs(O.fourier_lattice_r_to_k(gr.reshape(-1, 1), (N, N, 1)).reshape(N, N) - e)) < 1e-13


def _real_test_meshes():
    w_max, Nw = 40.0, 1001  # fourier_real.cpp:29-31
    t_lo, t_hi, _ = O.adjoint_real_mesh(-w_max, w_max, Nw)
    return -w_max, w_max, t_lo, t_hi, Nw


def test_fourier_real_lorentzian_known_answers():
    """fourier_real.cpp:49-62 (Lorentzian <-> exp(-E|t|)) and :66-87 (theta x exponential <-> two poles), precision 1e-4."""
    w_lo, w_hi, t_lo, t_hi, L = _real_test_meshes()
    E = 1.0
    w, t = O.linear_values(w_lo, w_hi, L), O.linear_values(t_lo, t_hi, L)
    gw = np.repeat((2 * E / (w * w + E * E)).astype(complex)[:, None], 4, axis=1)
    tail = np.zeros((3, 4), complex)
    tail[2] = 2 * E  # the exact high-frequency expansion 2E / w^2 (the reference fits it, :52)
    gt = O.fourier_w_to_t(gw, t_lo, t_hi, w_lo, w_hi, tail)
    assert np.max(np.abs(gt - np.exp(-E * np.abs(t))[:, None])) < 1e-4
    assert np.max(np.abs(O.fourier_t_to_w(gt, t_lo, t_hi, w_lo, w_hi, tail) - gw)) < 1e-4
    theta = lambda x: np.where(x > 0, 1.0, np.where(x < 0, 0.0, 0.5))
    gt2 = np.repeat((0.5j * (np.exp(-E * np.abs(t)) * theta(-t) - np.exp(-E * np.abs(t)) * theta(t)))[:, None], 4, axis=1)
    km = np.zeros((3, 4), complex)
    km[1] = 1.0
    gw2 = O.fourier_t_to_w(gt2, t_lo, t_hi, w_lo, w_hi, km)
    assert np.max(np.abs(gw2 - (0.5 / (w + 1j * E) + 0.5 / (w - 1j * E))[:, None])) < 1e-4
    assert np.max(np.abs(O.fourier_w_to_t(gw2, t_lo, t_hi, w_lo, w_hi, km) - gt2)) < 1e-4
    # without moments the raw data is transformed (:41-43); incompatible meshes throw (:50-51)
    assert np.max(np.abs(O.fourier_w_to_t(O.fourier_t_to_w(gt2, t_lo, t_hi, w_lo, w_hi), t_lo, t_hi, w_lo, w_hi) - gt2)) < 1e-12
    with pytest.raises(RuntimeError):
        O.fourier_t_to_w(gt2, t_lo, 1.01 * t_hi, w_lo, w_hi)


def test_legendre_known_answers():
    """gf_legendre.cpp:22-72: g_l = delta_{l0} -> g(tau) = 1/beta;  g_l = delta_{l2} -> sqrt(5)/beta (3x^2-1)/2; and the
    properties of T_{nl} (legendre.cpp:24-42): T_{-n-1,l} = conj T_{nl}, unitarity of the l <-> tau pair."""
    beta, n_l, n_tau = 2.0, 10, 5
    g = np.zeros((n_l, 1), complex)
    g[0] = 1
    assert np.array_equal(O.legendre_to_imtime(g, beta, n_tau)[:, 0], np.full(n_tau, 1 / beta, complex))
    g[:] = 0
    g[2] = 1
    x = 2 * O.tau_values(beta, n_tau) / beta - 1
    assert np.max(np.abs(O.legendre_to_imtime(g, beta, n_tau)[:, 0] - math.sqrt(5) / beta * (3 * x * x - 1) / 2)) < 1e-15
    for n in range(0, 5):
        for l in range(0, 6):
            assert O.legendre_T(-n - 1, l) == np.conj(O.legendre_T(n, l))
    # T_{n0} = i (-1)^n j_0((n+1/2) pi) = i / ((n+1/2) pi):  G(iw) of g_l = delta_{l0} is 1/(i w_n) x (-2/beta)... check the value itself
    assert abs(O.legendre_T(3, 0) - 1j * (-1) ** 3 * math.sin(3.5 * math.pi) / (3.5 * math.pi)) < 1e-15
    # tau -> l -> tau is the identity on the span of the first n_l polynomials (up to the trapezoid error)
    rng = np.random.default_rng(5)
    gl = rng.normal(size=(8, 3)) + 1j * rng.normal(size=(8, 3))
    back = O.imtime_to_legendre(O.legendre_to_imtime(gl, 3.0, 20001), 3.0, 8)
    assert np.max(np.abs(back - gl)) < 1e-5
    # l -> iw -> l (through 50000 tau points, legendre_matsubara.hpp:85) for a smooth physical G
    beta = 10.0
    tau = O.tau_values(beta, 40001)
    gtau = (-np.exp(-0.7 * tau) / (1 + np.exp(-beta * 0.7)))[:, None]
    gl = O.imtime_to_legendre(gtau, beta, 30)
    gw = O.legendre_to_imfreq(gl, O.FERMION, 200)
    iw = O.imfreq_values(beta, O.FERMION, 200)
    assert np.max(np.abs(gw[:, 0] - 1 / (iw - 0.7))) < 2e-6
    assert np.max(np.abs(O.imfreq_to_legendre(gw, beta, O.FERMION, 30) - gl)) < 1e-5


def test_fit_tail_real_known_answers():
    """test/c++/gfs/fit_tail_real.cpp:24-91 (Basic, Complex)."""
    L, om = 201, 10.0
    w = O.linear_values(-om, om, L)
    c = np.array([0, 1, 3, 5], complex)
    g = (c[0] + c[1] / (w + 1e-14) + c[2] / (w * w + 1e-14) + c[3] / (w * w * w + 1e-14)).astype(complex)[:, None]
    for known in (np.zeros((1, 1), complex), np.array([[0.0], [1.0]], complex)):
        tail, err = O.fit_tail_refreq(g, -om, om, known, tail_fraction=0.3)
        assert np.max(np.abs(tail[:4, 0] - c)) < 1e-9
    om = 100.0
    w = O.linear_values(-om, om, L)
    a = 1.0 + 0.4j
    g = (1 / (w - a))[:, None]
    tail, err = O.fit_tail_refreq(g, -om, om, np.array([[0.0], [1.0]], complex))
    assert np.max(np.abs(tail[:5, 0] - np.array([0, 1, a, a ** 2, a ** 3]))) < 1e-7


def test_evaluate_prod_reference_known_answers():
    """test/c++/gfs/multivar/g_k_om.cpp:76-114 (Gkom.Eval): g(k, iw) on a 40 x 40 brzone x imfreq mesh re-evaluated on the mesh itself
    (1e-14) and on a 3 x finer k grid (within 5 %, the reference's own bound); the interpolation rules are
    mesh/evaluate.hpp:97-114, brzone.hpp:355-373, linear.hpp:221-228."""
    beta, n_k, n_w = 5.0, 40, 5
    iw = O.imfreq_values(beta, O.FERMION, n_w)
    k = np.arange(n_k) / n_k * 2 * np.pi
    g = 1 / (iw[None, None, None, :] + 3 + 2 * (np.cos(k)[:, None, None, None] + np.cos(k)[None, :, None, None]))
    g = np.ascontiguousarray(np.broadcast_to(g, (n_k, n_k, 1, len(iw))))[..., None]
    axes = ["periodic", "periodic", "periodic", "index"]
    pts = np.array([[ix, iy, 0, n] for ix in range(n_k) for iy in range(n_k) for n in range(len(iw))], float)
    r = O.evaluate_prod(g, axes, pts)
    assert np.max(np.abs(r[:, 0] - g[..., 0].reshape(-1))) < 1e-14
    n2 = 3 * n_k
    pts, ref = [], []
    for a in range(n2):
        for b in range(n2):
            kx, ky = a / n2 * 2 * np.pi, b / n2 * 2 * np.pi
            for n in range(len(iw)):
                pts.append([kx / (2 * np.pi) * n_k, ky / (2 * np.pi) * n_k, 0.0, n])  # transpose(units_inv) * k, brzone.hpp:368
                ref.append(1 / (iw[n] + 3 + 2 * (np.cos(kx) + np.cos(ky))))
    r = O.evaluate_prod(g, axes, np.array(pts))[:, 0]
    ref = np.array(ref)
    assert np.all(np.abs(r - ref) <= 0.05 * np.abs(ref))
    # one linear component == gf_evaluator<imtime> (evaluator.hpp:35-43)
    rng = np.random.default_rng(0)
    tab = rng.normal(size=(101, 3)) + 1j * rng.normal(size=(101, 3))
    taus = rng.uniform(0, 7.0, 500)
    assert np.array_equal(O.evaluate_prod(tab, [("linear", 0.0, 7.0)], taus[:, None]), O.interp_tau(tab, 7.0, taus))
    # a bilinear function is reproduced exactly by two linear components (evaluate.hpp:104-110: the first component is the inner sum)
    t1, t2 = O.linear_values(0.0, 2.0, 11), O.linear_values(-1.0, 1.0, 21)
    f = (1 + 2 * t1[:, None] - 3 * t2[None, :] + 0.5 * t1[:, None] * t2[None, :]).astype(complex)[..., None]
    xy = np.stack([rng.uniform(0, 2, 300), rng.uniform(-1, 1, 300)], axis=1)
    r = O.evaluate_prod(f, [("linear", 0.0, 2.0), ("linear", -1.0, 1.0)], xy)[:, 0]
    assert np.max(np.abs(r - (1 + 2 * xy[:, 0] - 3 * xy[:, 1] + 0.5 * xy[:, 0] * xy[:, 1]))) < 1e-13
