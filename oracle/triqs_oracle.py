"""CPU restatement (NumPy, FP64) of the TRIQS Green's-function hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  The product (``triqs_b200``) never does and
fails loudly when the CUDA library is missing.

Every function cites the reference file:line it restates (paths relative to the
TRIQS 3.3.1 tree).  The reference itself cannot be built here (it needs nda, h5,
mpi, FFTW, Boost, HDF5 -- none present, no network), so the oracle is pinned
against the known answers the reference's own tests assert
(``tests/test_oracle_golden.py``).

Third-party arithmetic that is not in the reference tree and is restated from
its published algorithm:
  * FFTW3 (unpinned system lib)  -> ``numpy.fft`` (pocketfft); FFTW_BACKWARD is
    the unnormalised e^{+2 pi i jk/L} sum == ``ifft * L``; FFTW_FORWARD == ``fft``.
  * TRIQS/nda 1.3.x ``nda::lapack::gelss_worker{,_hermitian}`` -> SVD pseudo
    inverse (``numpy.linalg.svd`` = LAPACK zgesdd; the reference uses zgesvd --
    both backward stable; results agree to ~cond*eps).
  * ``nda::inverse_in_place`` -> ``numpy.linalg.inv`` (LAPACK zgetrf/zgetri).
"""
from __future__ import annotations

import math
import numpy as np

FERMION = 1  # c++/triqs/mesh/domains/matsubara.hpp: statistic_enum {Boson = 0, Fermion = 1}
BOSON = 0
REFREQ = 2  # (oracle-internal tag: the tail fitter on a real-frequency mesh)

# ----------------------------------------------------------------------------------------------
# meshes
# ----------------------------------------------------------------------------------------------


def tau_values(beta: float, n_tau: int) -> np.ndarray:
    """imtime mesh values.  c++/triqs/mesh/bases/linear.hpp:154-160 (to_value):
    wr = index/(L-1); xmin*(1-wr) + xmax*wr with xmin=0, xmax=beta, L=n_tau."""
    idx = np.arange(n_tau, dtype=np.float64)
    wr = idx / (n_tau - 1)
    return 0.0 * (1 - wr) + beta * wr


def tau_delta(beta: float, n_tau: int) -> float:
    """c++/triqs/mesh/bases/linear.hpp:51  delta_x = (b-a)/(L-1)."""
    return (beta - 0.0) / (n_tau - 1)


def imfreq_first_last(stat: int, n_iw: int, positive_only: bool = False):
    """c++/triqs/mesh/imfreq.hpp:58-60,77-80."""
    last = n_iw - 1
    first = 0 if positive_only else -(last + (1 if stat == FERMION else 0))
    return first, last


def imfreq_values(beta: float, stat: int, n_iw: int, positive_only: bool = False) -> np.ndarray:
    """i*omega_n for the whole mesh.  c++/triqs/mesh/domains/matsubara.hpp:55:
    complex(0, pi*(2n+statistic)/beta)."""
    first, last = imfreq_first_last(stat, n_iw, positive_only)
    n = np.arange(first, last + 1, dtype=np.int64)
    return 1j * (math.pi * (2 * n + stat).astype(np.float64) / beta)


def adjoint_n_iw(n_tau: int) -> int:
    """c++/triqs/mesh/adjoint.hpp:31-34."""
    return (n_tau - 1) // 6


def adjoint_n_tau(n_iw: int) -> int:
    """c++/triqs/mesh/adjoint.hpp:35-38."""
    return 6 * n_iw + 1


# ----------------------------------------------------------------------------------------------
# 3-pole tail model
# ----------------------------------------------------------------------------------------------


def one_fermion(a, b: float, tau, beta: float):
    """c++/triqs/gfs/transform/fourier_matsubara.cpp:28-30."""
    if b >= 0:
        f = np.exp(-b * tau) / (1 + math.exp(-beta * b))
    else:
        f = np.exp(b * (beta - tau)) / (1 + math.exp(beta * b))
    return -np.multiply.outer(f, a)


def one_boson(a, b: float, tau, beta: float):
    """c++/triqs/gfs/transform/fourier_matsubara.cpp:32-34."""
    if b >= 0:
        f = np.exp(-b * tau) / (math.exp(-beta * b) - 1)
    else:
        f = np.exp(b * (beta - tau)) / (1 - math.exp(b * beta))
    return np.multiply.outer(f, a)


def _poles(stat: int, m1, m2, m3):
    """c++/triqs/gfs/transform/fourier_matsubara.cpp:151-169 and :238-252."""
    if stat == FERMION:
        b = (0.0, 1.0, -1.0)
        a = (m1 - m3, (m2 + m3) / 2, (m3 - m2) / 2)
    else:
        b = (-0.5, -1.0, 1.0)
        a = (4 * (m1 - m3) / 3, m3 - (m1 + m2) / 2, m1 / 6 + m2 / 2 + m3 / 3)
    return a, b


# literal digits of the inverse Vandermonde, c++/triqs/gfs/transform/fourier_matsubara.cpp:52-67
V_INV = np.array([
    [8.000000000008853, -28.000000000055913, 56.000000000151815, -70.00000000021936, 56.00000000020795, -28.000000000115904,
     8.00000000003683, -1.0000000000049232],
    [-13.742857142879107, 62.10000000013776, -133.5333333337073, 172.75000000055635, -141.00000000052023, 71.43333333362274,
     -20.600000000091182, 2.592857142869304],
    [9.694444444465196, -50.98055555568518, 119.55000000035201, -161.8888888894241, 135.86111111160685, -70.12500000027573,
     20.494444444530682, -2.6055555555670558],
    [-3.6555555555654573, 21.23472222228416, -53.60000000016842, 76.34027777803497, -66.27777777801586, 35.03750000013273,
     -10.422222222263619, 1.3430555555610917],
    [0.7986111111137331, -4.97222222223865, 13.312500000044732, -19.888888888957325, 17.923611111174495, -9.750000000035401,
     2.9652777777888164, -0.38888888889036743],
    [-0.1013888888892789, 0.6638888888913357, -1.862500000006671, 2.902777777787997, -2.71527777778725, 1.5250000000052975,
     -0.47638888889054154, 0.06388888888911043],
    [0.0069444444444749, -0.047222222222413485, 0.13750000000052204, -0.22222222222302282, 0.2152777777785205,
     -0.12500000000041575, 0.04027777777790756, -0.00555555555557296],
    [-0.0001984126984136688, 0.0013888888888949882, -0.00416666666668333, 0.006944444444470023, -0.006944444444468193,
     0.004166666666679969, -0.0013888888888930434, 0.00019841269841325577]])


def fit_tail_tau(gt: np.ndarray, beta: float, stat: int) -> np.ndarray:
    """tau-derivative estimate of moments 0..3 of G(tau) given as [n_tau, n_others].
    c++/triqs/gfs/transform/fourier_matsubara.cpp:39-92."""
    n_tau = gt.shape[0]
    tau = tau_values(beta, n_tau)
    fit_order = 8
    dl = np.empty((fit_order, gt.shape[1]), dtype=np.complex128)
    dr = np.empty_like(dl)
    for m in range(1, fit_order + 1):
        dl[m - 1] = (gt[m] - gt[0]) / tau[m]
        dr[m - 1] = (gt[n_tau - 1] - gt[n_tau - 1 - m]) / tau[m]
    gl = V_INV @ dl
    gr = V_INV @ dr
    sign = -1.0 if stat == FERMION else 1.0
    tail = np.zeros((4, gt.shape[1]), dtype=np.complex128)
    tail[1] = -(gt[0] - sign * gt[n_tau - 1])
    tail[2] = gl[0] - sign * gr[0]
    tail[3] = -(gl[1] + sign * gr[1]) * 2 / tau_delta(beta, n_tau)
    return tail


class TriqsRuntimeError(RuntimeError):
    """stands in for triqs::runtime_error (c++/triqs/utility/exceptions.hpp:31-37)."""


WARN_NOISY_TAU = 1      # fourier_matsubara.cpp:109-110
WARN_SHORT_TAU_MESH = 2  # fourier_matsubara.cpp:131-133
WARN_TAIL_ERR = 4       # fourier_matsubara.cpp:208-210


def fourier_tau_to_iw(gt: np.ndarray, beta: float, stat: int, n_iw: int, known_moments: np.ndarray | None = None,
                      return_info: bool = False):
    """Direct transform G(tau)[n_tau, n_others] -> G(iw)[2*n_iw(-1 boson), n_others].
    c++/triqs/gfs/transform/fourier_matsubara.cpp:96-186."""
    gt = np.asarray(gt, dtype=np.complex128)
    n_tau, n_others = gt.shape
    warn = 0
    delta = tau_delta(beta, n_tau)
    if known_moments is None or known_moments.size == 0:
        der_1 = np.max(np.abs(gt[1] - gt[0]) + np.abs(gt[n_tau - 2] - gt[n_tau - 1])) / delta
        der_2 = 0.5 * np.max(np.abs(gt[2] - gt[0]) + np.abs(gt[n_tau - 3] - gt[n_tau - 1])) / delta
        if der_1 < 0.95 * der_2 or der_1 > 1.05 * der_2:
            warn |= WARN_NOISY_TAU
            tail = np.zeros((4, n_others), dtype=np.complex128)
        else:
            tail = fit_tail_tau(gt, beta, stat)
    else:
        known_moments = np.asarray(known_moments, dtype=np.complex128)
        abs0 = np.max(np.abs(known_moments[0]))
        if not abs0 < 1e-8:
            raise TriqsRuntimeError("ERROR: Direct Fourier implementation requires vanishing 0th moment\n  error is :%f" % abs0)
        nk = min(known_moments.shape[0], 4)
        tail = np.zeros((4, n_others), dtype=np.complex128)
        tail[:nk] = known_moments[:nk]

    first, last = imfreq_first_last(stat, n_iw)
    L = n_tau - 1
    if L < 2 * (last + 1):
        raise TriqsRuntimeError("Fourier: The time mesh mush be at least twice as long as the number of positive frequencies")
    if L < 6 * (last + 1):
        warn |= WARN_SHORT_TAU_MESH

    is_fermion = stat == FERMION
    fact = beta / L
    m1, m2, m3 = tail[1], tail[2], tail[3]
    a, b = _poles(stat, m1, m2, m3)
    tau = tau_values(beta, n_tau)
    if is_fermion:
        model = one_fermion(a[0], b[0], tau, beta) + one_fermion(a[1], b[1], tau, beta) + one_fermion(a[2], b[2], tau, beta)
        phase = np.exp(1j * (math.pi / beta) * tau)
        gin = fact * phase[:, None] * (gt - model)
    else:
        model = one_boson(a[0], b[0], tau, beta) + one_boson(a[1], b[1], tau, beta) + one_boson(a[2], b[2], tau, beta)
        gin = fact * (gt - model)
    # FFTW_BACKWARD over rows 0..L-1 (row L is computed but not transformed), fourier_common.cpp:30-44
    gout = np.fft.ifft(gin[:L], axis=0) * L
    corr = -0.5 * fact * (gt[0] + m1 + (1 if is_fermion else -1) * gt[L])
    iw = imfreq_values(beta, stat, n_iw)
    n = np.arange(first, last + 1)
    gw = (gout[(n + L) % L] + corr[None, :] + a[0][None, :] / (iw[:, None] - b[0]) + a[1][None, :] / (iw[:, None] - b[1])
          + a[2][None, :] / (iw[:, None] - b[2]))
    if return_info:
        return gw, {"warn": warn, "tail": tail}
    return gw


def fourier_iw_to_tau(gw: np.ndarray, beta: float, stat: int, n_tau: int, known_moments: np.ndarray | None = None,
                      return_info: bool = False, fitter: "TailFitter | None" = None):
    """Inverse transform G(iw)[n_w, n_others] -> G(tau)[n_tau, n_others].
    c++/triqs/gfs/transform/fourier_matsubara.cpp:190-271."""
    gw = np.asarray(gw, dtype=np.complex128)
    n_w, n_others = gw.shape
    # full mesh only (:192): size = 2*n_iw (F) or 2*n_iw-1 (B)
    if stat == FERMION:
        if n_w % 2 != 0:
            raise TriqsRuntimeError("Fourier is only implemented for g(i omega_n) with full mesh (positive and negative frequencies)")
        n_iw = n_w // 2
    else:
        if n_w % 2 != 1:
            raise TriqsRuntimeError("Fourier is only implemented for g(i omega_n) with full mesh (positive and negative frequencies)")
        n_iw = (n_w + 1) // 2
    warn = 0
    tail_err = 0.0
    if known_moments is None or known_moments.size == 0:
        known_moments = np.zeros((1, n_others), dtype=np.complex128)  # make_zero_tail(gw, 1), :197
    known_moments = np.asarray(known_moments, dtype=np.complex128)
    abs0 = np.max(np.abs(known_moments[0]))
    if not abs0 < 1e-8:
        raise TriqsRuntimeError("ERROR: Inverse Fourier implementation requires vanishing 0th moment\n  error is :%f\n" % abs0)
    if known_moments.shape[0] < 4:
        if fitter is None:
            fitter = TailFitter(beta, stat, n_iw)
        t, err = fitter.fit(gw, known_moments)
        tail_err = err
        if not err < 1e-2:
            raise TriqsRuntimeError("ERROR: High frequency moments have an error greater than 1e-2.\n  Error = %f" % err)
        if err > 1e-4:
            warn |= WARN_TAIL_ERR
        if not t.shape[0] > 3:
            raise TriqsRuntimeError("ERROR: Inverse Fourier implementation requires at least a proper 3rd high-frequency moment\n")
        tail = t
    else:
        tail = known_moments

    first, last = imfreq_first_last(stat, n_iw)
    L = n_tau - 1
    if L < 2 * (last + 1):
        raise TriqsRuntimeError("Inverse Fourier: The time mesh mush be at least twice as long as the freq mesh")
    is_fermion = stat == FERMION
    fact = 1.0 / beta
    m1, m2, m3 = tail[1], tail[2], tail[3]
    a, b = _poles(stat, m1, m2, m3)
    iw = imfreq_values(beta, stat, n_iw)
    n = np.arange(first, last + 1)
    gin = np.zeros((L, n_others), dtype=np.complex128)
    gin[(n + L) % L] = fact * (gw - (a[0][None, :] / (iw[:, None] - b[0]) + a[1][None, :] / (iw[:, None] - b[1])
                                     + a[2][None, :] / (iw[:, None] - b[2])))
    gout = np.fft.fft(gin, axis=0)  # FFTW_FORWARD
    tau = tau_values(beta, n_tau)
    gt = np.empty((n_tau, n_others), dtype=np.complex128)
    if is_fermion:
        model = one_fermion(a[0], b[0], tau, beta) + one_fermion(a[1], b[1], tau, beta) + one_fermion(a[2], b[2], tau, beta)
        gt[:L] = gout * np.exp(-1j * (math.pi / beta) * tau[:L])[:, None] + model[:L]
    else:
        model = one_boson(a[0], b[0], tau, beta) + one_boson(a[1], b[1], tau, beta) + one_boson(a[2], b[2], tau, beta)
        gt[:L] = gout + model[:L]
    pm = -1.0 if is_fermion else 1.0
    gt[L] = pm * (gt[0] + m1)
    if return_info:
        return gt, {"warn": warn, "tail": tail, "tail_err": tail_err}
    return gt


# ----------------------------------------------------------------------------------------------
# least-squares tail fit
# ----------------------------------------------------------------------------------------------


def _svd_pinv(A: np.ndarray):
    """gelss_worker ctor (nda 1.3.x nda/lapack/gelss_worker.hpp, not vendored in the reference
    tree): SVD A = U S V^H; pinv = V S^-1 U^H; U_null^H = U^H[N:M, :]."""
    U, s, Vh = np.linalg.svd(A, full_matrices=True)
    M, N = A.shape
    pinv = (Vh.conj().T / s[None, :]) @ U[:, :N].conj().T
    uh_null = U.conj().T[N:M, :]
    return pinv, uh_null, s


class TailFitter:
    """c++/triqs/mesh/tail_fitter.hpp:67-298 for an imfreq mesh (beta, stat, n_iw)."""

    max_order = 9
    rcond = 1e-8

    def __init__(self, beta: float, stat: int, n_iw: int, tail_fraction: float = 0.2, n_tail_max: int = 30,
                 expansion_order: int | None = None):
        self.beta, self.stat, self.n_iw = beta, stat, n_iw
        self.tail_fraction, self.n_tail_max = tail_fraction, n_tail_max
        self.adjust_order = expansion_order is None
        self.expansion_order = self.max_order if expansion_order is None else expansion_order
        if stat == REFREQ:
            # the fitter is generic in the mesh (tail_fitter.hpp:105-181): refreq(w_min, w_max, n_w) is passed as
            # (beta = (w_min, w_max), stat = REFREQ, n_iw = n_w); first_index 0, w_max() = xmax (refreq.hpp:52)
            w_lo, w_hi = beta
            self.first, self.last = 0, n_iw - 1
            self.size = n_iw
            self.w_max = abs(w_hi)
            self.fit_idx = self.get_tail_fit_indices()
            wr = np.asarray(self.fit_idx, dtype=np.float64) / (n_iw - 1)
            iw = (w_lo * (1 - wr) + w_hi * wr).astype(np.complex128)
        else:
            self.first, self.last = imfreq_first_last(stat, n_iw)
            self.size = self.last - self.first + 1
            self.w_max = math.pi * (2 * self.last + stat) / beta  # |imfreq::w_max()|, imfreq.hpp:128
            self.fit_idx = self.get_tail_fit_indices()
            iw = 1j * (math.pi * (2 * np.asarray(self.fit_idx, dtype=np.int64) + stat).astype(np.float64) / beta)
        # vander, tail_fitter.hpp:34-44 (z *= p accumulation), :149-155 (C = om_max / iw)
        C = self.w_max / iw
        V = np.empty((len(self.fit_idx), self.expansion_order + 1), dtype=np.complex128)
        z = np.ones(len(self.fit_idx), dtype=np.complex128)
        for p in range(self.expansion_order + 1):
            V[:, p] = z
            z = z * C
        self.vander = V
        self._lss = {}

    def get_tail_fit_indices(self):
        """tail_fitter.hpp:105-128 (double accumulated by += step, truncated by long())."""
        n_rng = int(round_half_away(self.tail_fraction * self.size / 2))
        n_tail = min(n_rng, self.n_tail_max)
        if n_tail <= 0:
            return []
        step = float(n_rng) / n_tail
        idx1 = float(self.first)
        idx2 = float(self.last - n_rng)
        out = []
        for _ in range(n_tail):
            out.append(int(idx1))  # long(): truncation toward zero
            out.append(int(idx2))
            idx1 += step
            idx2 += step
        return out

    def setup_lss(self, n_fixed: int, hermitian: bool):
        """tail_fitter.hpp:140-181."""
        V = self.vander
        if n_fixed + 1 > V.shape[0] // 2:
            raise TriqsRuntimeError("Insufficient data points for least square procedure")

        def worker(n):
            A = V[:, n_fixed:n + 1]
            pinv, uh_null, s = _svd_pinv(A)
            if hermitian:
                As = np.vstack([A, A.conj()])
                pinv_h, uh_null_h, _ = _svd_pinv(As)
                return {"A": A, "pinv": pinv_h, "uh_null": uh_null_h, "s": s, "n_var": A.shape[1]}
            return {"A": A, "pinv": pinv, "uh_null": uh_null, "s": s, "n_var": A.shape[1]}

        if not self.adjust_order:
            return worker(self.expansion_order)
        n_max = min(self.max_order, int(1.0 + 16.0 / math.log10(1 + abs(self.w_max))))
        n_max = min(n_max, V.shape[0] // 2)
        for n in range(n_max, n_fixed - 1, -1):
            w = worker(n)
            if w["s"][-1] > self.rcond:
                return w
        raise TriqsRuntimeError("Conditioning of tail-fit violates boundary")

    def _get(self, n_fixed, hermitian):
        key = (n_fixed, hermitian)
        if key not in self._lss:
            self._lss[key] = self.setup_lss(n_fixed, hermitian)
        return self._lss[key]

    def fit(self, g_data: np.ndarray, known_moments: np.ndarray | None = None, hermitian: bool = False,
            inner_dim: int | None = None, normalize: bool = True):
        """tail_fitter.hpp:192-288.  g_data [n_w, ncols] (mesh first), known [n_fixed, ncols].
        Returns (moments [n_moments, ncols], err)."""
        g_data = np.asarray(g_data, dtype=np.complex128)
        ncols = g_data.shape[1]
        if known_moments is None:
            known_moments = np.zeros((0, ncols), dtype=np.complex128)
        known_moments = np.asarray(known_moments, dtype=np.complex128)
        if hermitian and inner_dim is None:
            raise TriqsRuntimeError("Enforcing the hermiticity in tail_fit requires inner matrix dimension")
        n_fixed = known_moments.shape[0]
        if n_fixed > self.expansion_order:
            return known_moments.copy(), 0.0
        lss = self._get(n_fixed, hermitian)
        n_moments = lss["n_var"] + n_fixed
        rows = np.asarray(self.fit_idx, dtype=np.int64) - self.first
        g_mat = g_data[rows, :].copy()
        if n_fixed > 0:
            if known_moments.shape[1] != ncols:
                raise TriqsRuntimeError("known_moments shape incompatible with shape of data")
            z = 1.0
            km = np.empty_like(known_moments)
            for order in range(n_fixed):
                km[order] = z * known_moments[order]
                z /= self.w_max
            g_mat -= self.vander[:, :n_fixed] @ km
        M = g_mat.shape[0]
        if hermitian:
            d = inner_dim

            def inner_adjoint(C):
                sh = C.shape
                return C.reshape(sh[0], -1, d, d).transpose(0, 1, 3, 2).conj().reshape(sh)

            b_stack = np.vstack([g_mat, inner_adjoint(g_mat)])
            x = lss["pinv"] @ b_stack
            err = float(np.max(np.linalg.norm(lss["uh_null"] @ b_stack, axis=0)) / math.sqrt(2 * M))
            a_mat = 0.5 * (x + inner_adjoint(x))
        else:
            a_mat = lss["pinv"] @ g_mat
            err = float(np.max(np.linalg.norm(lss["uh_null"] @ g_mat, axis=0)) / math.sqrt(M)) if M != lss["n_var"] else 0.0
        if normalize:
            z = 1.0
            for _ in range(n_fixed):
                z *= self.w_max
            for i in range(a_mat.shape[0]):
                a_mat[i] *= z
                z *= self.w_max
        res = np.empty((n_moments, ncols), dtype=np.complex128)
        res[:n_fixed] = known_moments
        res[n_fixed:] = a_mat
        return res, err


def round_half_away(x: float) -> float:
    """std::round semantics (half away from zero), used at tail_fitter.hpp:89,108."""
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


def fit_tail(gw, beta, stat, known_moments=None, **fitter_kw):
    """c++/triqs/gfs/functions/functions2.hpp:50-56.  gw [n_w, ncols]."""
    n_w = gw.shape[0]
    n_iw = n_w // 2 if stat == FERMION else (n_w + 1) // 2
    return TailFitter(beta, stat, n_iw, **fitter_kw).fit(gw, known_moments)


def fit_tail_refreq(gw, w_min: float, w_max: float, known_moments=None, **fitter_kw):
    """fit_tail of a real-frequency gf (functions2.hpp:50-56 on a refreq mesh; test/c++/gfs/fit_tail_real.cpp).  gw [n_w, ncols]."""
    return TailFitter((w_min, w_max), REFREQ, gw.shape[0], **fitter_kw).fit(gw, known_moments)


def fit_hermitian_tail(gw, beta, stat, known_moments=None, inner_dim: int = 1, **fitter_kw):
    """c++/triqs/gfs/functions/functions2.hpp:94-110 (inner_matrix_dim = 1 scalar / n matrix)."""
    n_w = gw.shape[0]
    n_iw = n_w // 2 if stat == FERMION else (n_w + 1) // 2
    return TailFitter(beta, stat, n_iw, **fitter_kw).fit(gw, known_moments, hermitian=True, inner_dim=inner_dim)


def tail_eval(moments: np.ndarray, om: complex) -> np.ndarray:
    """sum_n A_n / om^n.  c++/triqs/mesh/tail_fitter.hpp:49-64."""
    z = 1.0 + 0j
    res = np.zeros(moments.shape[1:], dtype=np.complex128)
    for n in range(moments.shape[0]):
        res = res + moments[n] * z
        z = z / om
    return res


# ----------------------------------------------------------------------------------------------
# retime <-> refreq transforms     c++/triqs/gfs/transform/fourier_real.cpp:35-131
# ----------------------------------------------------------------------------------------------


def linear_values(xmin: float, xmax: float, n: int) -> np.ndarray:
    """mesh/bases/linear.hpp:154-160:  x_i = xmin (1 - i/(n-1)) + xmax i/(n-1)."""
    if n == 1:
        return np.array([xmin])
    wr = np.arange(n, dtype=np.float64) / (n - 1)
    return xmin * (1 - wr) + xmax * wr


def adjoint_real_mesh(xmin: float, xmax: float, n: int, shift_half_bin: bool = False):
    """mesh/adjoint.hpp:43-54 (the same formula in both directions)."""
    delta = (xmax - xmin) / (n - 1)
    m = math.pi * (n - 1) / (n * delta)
    if shift_half_bin:
        return -m + math.pi / n / delta, m + math.pi / n / delta, n
    return -m, m, n


def _real_meshes(t_min, t_max, w_min, w_max, L):
    t, w = linear_values(t_min, t_max, L), linear_values(w_min, w_max, L)
    dt, dw = (t_max - t_min) / (L - 1), (w_max - w_min) / (L - 1)
    if abs(dt * dw * L / (2 * math.pi) - 1) > 1e-10:  # fourier_real.cpp:50-51, :100-101
        raise RuntimeError("Meshes are not compatible")
    return t, w, dt, dw


def _real_tail(known_moments, n_cols, dw, L):
    if known_moments is None or np.size(known_moments) == 0:
        known_moments = np.zeros((3, n_cols), complex)  # :41-43 make_zero_tail(g, 3)
    known_moments = np.asarray(known_moments, dtype=np.complex128)
    if known_moments.shape[0] < 3:
        raise RuntimeError("Direct RealTime Fourier transform requires known moments up to order 2.")
    if np.max(np.abs(known_moments[0])) >= 1e-8:
        raise RuntimeError("Fourier implementation requires vanishing 0th moment")
    a = dw * math.sqrt(float(L))  # :60
    m1, m2 = known_moments[1], known_moments[2]
    return a, (m1 + 1j * m2 / a) / 2.0, (m1 - 1j * m2 / a) / 2.0  # :66


def _th_expo(t, a):  # :27
    return np.where(t > 0, -1j * np.exp(-a * t), np.where(t < 0, 0.0, -0.5j * np.exp(-a * t)))


def _th_expo_neg(t, a):  # :28
    return np.where(t < 0, 1j * np.exp(a * t), np.where(t > 0, 0.0, 0.5j * np.exp(a * t)))


def fourier_t_to_w(gt: np.ndarray, t_min: float, t_max: float, w_min: float, w_max: float, known_moments=None) -> np.ndarray:
    """fourier_real.cpp:35-78.  gt [L, n_cols] -> gw [L, n_cols]."""
    gt = np.asarray(gt, dtype=np.complex128)
    L, C = gt.shape
    t, w, dt, dw = _real_meshes(t_min, t_max, w_min, w_max, L)
    a, a1, a2 = _real_tail(known_moments, C, dw, L)
    gin = (gt - (a1[None] * _th_expo(t, a)[:, None] + a2[None] * _th_expo_neg(t, a)[:, None])) * np.exp(1j * t * w_min)[:, None]
    gout = np.fft.ifft(gin, axis=0) * L  # FFTW_BACKWARD, unnormalised
    return dt * np.exp(1j * (w - w_min) * t_min)[:, None] * gout + a1[None] / (w + 1j * a)[:, None] + a2[None] / (w - 1j * a)[:, None]


def fourier_w_to_t(gw: np.ndarray, t_min: float, t_max: float, w_min: float, w_max: float, known_moments=None) -> np.ndarray:
    """fourier_real.cpp:82-131.  gw [L, n_cols] -> gt [L, n_cols]."""
    gw = np.asarray(gw, dtype=np.complex128)
    L, C = gw.shape
    t, w, dt, dw = _real_meshes(t_min, t_max, w_min, w_max, L)
    a, a1, a2 = _real_tail(known_moments, C, dw, L)
    gin = (gw - a1[None] / (w + 1j * a)[:, None] - a2[None] / (w - 1j * a)[:, None]) * np.exp(-1j * w * t_min)[:, None]
    gout = np.fft.fft(gin, axis=0)  # FFTW_FORWARD
    corr = 1.0 / (dt * L)
    return corr * np.exp(1j * w_min * (t_min - t))[:, None] * gout + a1[None] * _th_expo(t, a)[:, None] + a2[None] * _th_expo_neg(t, a)[:, None]


# ----------------------------------------------------------------------------------------------
# Legendre <-> Matsubara     c++/triqs/gfs/transform/legendre_matsubara.hpp:37-129, utility/legendre.{hpp,cpp}
# ----------------------------------------------------------------------------------------------


def legendre_T(n: int, l: int) -> complex:
    """legendre.cpp:24-42:  sqrt(2l+1)/sqrt(2n+1) e^{i(n+1/2)pi} i^l J_{l+1/2}((n+1/2)pi), T_{-n-1,l} = conj(T_{n,l}).
    J_{l+1/2}(x) = sqrt(2x/pi) j_l(x) and sqrt(2x/pi) = sqrt(2n+1) at x = (n+1/2)pi."""
    from scipy.special import spherical_jn
    neg = n < 0
    if neg:
        n = abs(n + 1)
    x = (n + 0.5) * math.pi
    res = math.sqrt(2 * l + 1) * np.exp(1j * x) * (1j ** l) * spherical_jn(l, x)
    return np.conj(res) if neg else res


def legendre_polynomials(x: np.ndarray, n_l: int) -> np.ndarray:
    """legendre_generator (legendre.hpp:39-58): P_0 = 1, P_1 = x, n P_n = (2n-1) x P_{n-1} - (n-1) P_{n-2}.  Returns [n_l, len(x)]."""
    x = np.asarray(x, dtype=np.float64)
    P = np.empty((n_l, x.size))
    P[0] = 1.0
    if n_l > 1:
        P[1] = x
    for n in range(2, n_l):
        P[n] = ((2 * n - 1) * x * P[n - 1] - (n - 1) * P[n - 2]) / n
    return P


def legendre_to_imfreq(gl: np.ndarray, stat: int, n_iw: int) -> np.ndarray:
    """legendre_matsubara.hpp:37-50.  gl [n_l, n_cols] -> gw [n_w, n_cols] (the Matsubara *index* enters T whatever the statistic)."""
    gl = np.asarray(gl, dtype=np.complex128)
    first, last = imfreq_first_last(stat, n_iw)
    T = np.array([[legendre_T(n, l) for l in range(gl.shape[0])] for n in range(first, last + 1)])
    return T @ gl


def legendre_to_imtime(gl: np.ndarray, beta: float, n_tau: int) -> np.ndarray:
    """legendre_matsubara.hpp:56-70."""
    gl = np.asarray(gl, dtype=np.complex128)
    n_l = gl.shape[0]
    P = legendre_polynomials(2 * tau_values(beta, n_tau) / beta - 1, n_l)
    f = np.sqrt(2 * np.arange(n_l) + 1.0) / beta
    return (P * f[:, None]).T @ gl


def imtime_to_legendre(gt: np.ndarray, beta: float, n_l: int) -> np.ndarray:
    """legendre_matsubara.hpp:97-118: trapezoid rule over tau."""
    gt = np.asarray(gt, dtype=np.complex128)
    n_tau = gt.shape[0]
    P = legendre_polynomials(2 * tau_values(beta, n_tau) / beta - 1, n_l)
    coef = np.ones(n_tau)
    coef[0] = coef[-1] = 0.5
    f = np.sqrt(2 * np.arange(n_l) + 1.0)
    return ((P * f[:, None]) * coef[None, :]) @ gt * tau_delta(beta, n_tau)


def imfreq_to_legendre(gw: np.ndarray, beta: float, stat: int, n_l: int) -> np.ndarray:
    """legendre_matsubara.hpp:76-93: through G(tau) on 50000 points, tail fitted by the transform."""
    return imtime_to_legendre(fourier_iw_to_tau(gw, beta, stat, 50000), beta, n_l)


# ----------------------------------------------------------------------------------------------
# tight-binding dispersion on a brzone mesh  
# ----------------------------------------------------------------------------------------------


def tight_binding_fourier(displ, hop: np.ndarray, dims) -> np.ndarray:
    """h_k = sum_j hop[j] exp(2 pi i k . displ[j]) with k = (i0/d0, i1/d1, i2/d2), C-order k index.
    c++/triqs/lattice/tight_binding.hpp:88-123; mesh/brzone.hpp:187-190,285-288."""
    displ = np.asarray(displ, dtype=np.int64).reshape(-1, 3)
    hop = np.asarray(hop, dtype=np.complex128)
    d = np.asarray(dims, dtype=np.int64)
    idx = np.stack(np.unravel_index(np.arange(int(np.prod(d))), tuple(d)), axis=1)  # [nk, 3]
    kfrac = idx / d[None, :]
    phase = np.exp(2j * np.pi * (kfrac @ displ.T))  # [nk, n_r]
    return np.einsum("kj,jab->kab", phase, hop)


# ----------------------------------------------------------------------------------------------
# hermiticity helpers and replace_by_tail  
# ----------------------------------------------------------------------------------------------


def make_hermitian(g: np.ndarray, imtime: bool = False) -> np.ndarray:
    """0.5 (g[w] + dagger(g[-w])) on a full imfreq mesh (-w is the mirrored data index), 0.5 (g[t] + dagger(g[t])) on imtime.
    g [n_pts, d, d].  c++/triqs/gfs/functions/imfreq.hpp:175-215."""
    g = np.asarray(g, dtype=np.complex128)
    other = g if imtime else g[::-1]
    return 0.5 * (g + other.conj().transpose(0, 2, 1))


def is_gf_hermitian(g: np.ndarray, imtime: bool = False, tolerance: float = 1e-12) -> bool:
    """imfreq.hpp:81-124."""
    g = np.asarray(g, dtype=np.complex128)
    other = g if imtime else g[::-1]
    return bool(np.max(np.abs(other.conj().transpose(0, 2, 1) - g)) <= tolerance)


def replace_by_tail(gw: np.ndarray, beta: float, stat: int, tail: np.ndarray, n_min: int) -> np.ndarray:
    """g[iw] = tail_eval(tail, iw) for iw.n >= n_min or iw.n < -n_min.  gw [n_w, ncols], tail [n_mom, ncols].  imfreq.hpp:258-261."""
    gw = np.array(gw, dtype=np.complex128)
    n_w = gw.shape[0]
    n_iw = n_w // 2 if stat == FERMION else (n_w + 1) // 2
    first, _ = imfreq_first_last(stat, n_iw)
    iw = imfreq_values(beta, stat, n_iw)
    for i in range(n_w):
        n = first + i
        if n >= n_min or n < -n_min:
            gw[i] = tail_eval(tail, iw[i])
    return gw


# ----------------------------------------------------------------------------------------------
# density  (the immediate consumer of G_loc in a DMFT loop)
# ----------------------------------------------------------------------------------------------


def density(gw: np.ndarray, beta: float, stat: int, known_moments: np.ndarray | None = None, return_info: bool = False):
    """n = -xi [ (1/beta) sum_n (G(i w_n) - tail model) + m1 + sum_i F(a_i, b_i) ] for a matrix-valued gf on the FULL
    Matsubara mesh.  gw [n_w, n, n]; known_moments [n_mom, n, n] or None.  c++/triqs/gfs/functions/density.cpp:32-120:
    only the upper triangle is summed, the lower one is its conjugate (:115); moments are fitted with fit_tail when
    fewer than four are known (:47-58)."""
    gw = np.asarray(gw, dtype=np.complex128)
    n_w, n1, n2 = gw.shape
    info = {"tail_err": 0.0, "warn": 0}
    if known_moments is None or np.size(known_moments) == 0:
        known_moments = np.zeros((1, n1, n2), dtype=np.complex128)  # make_zero_tail(g, 1)   :40
    km = np.asarray(known_moments, dtype=np.complex128)
    a0 = float(np.max(np.abs(km[0])))
    if not a0 < 1e-8:
        raise TriqsRuntimeError("ERROR: Density implementation requires vanishing 0th moment\n  error is :%f\n" % a0)  # :42-44
    if km.shape[0] < 4:
        tail, err = fit_tail(gw.reshape(n_w, n1 * n2), beta, stat, km.reshape(km.shape[0], n1 * n2))  # :47
        info["tail_err"] = err
        if not err < 1e-2:
            raise TriqsRuntimeError("ERROR: High frequency moments have an error greater than 1e-2.\n  Error = %f\n" % err)
        if err > 1e-4:
            info["warn"] |= WARN_TAIL_ERR
        if not tail.shape[0] > 3:
            raise TriqsRuntimeError("ERROR: Density implementation requires at least a proper 3rd high-frequency moment\n")
        km = tail.reshape(tail.shape[0], n1, n2)
    m1, m2, m3 = km[1], km[2], km[3]
    if stat == FERMION:  # :70-79
        xi, b = -1.0, (0.0, 1.0, -1.0)
        a = (m1 - m3, (m2 + m3) / 2, (m3 - m2) / 2)  # :98-101
    else:
        xi, b = 1.0, (-1.0, 1.0, 0.5)
        a = (m1 / 6.0 - m2 / 2.0 + m3 / 3.0, -m1 / 2.0 + m2 / 2.0 + m3, 4.0 * m1 / 3.0 - 4.0 * m3 / 3.0)  # :102-105
    n_iw = n_w // 2 if stat == FERMION else (n_w + 1) // 2
    iw = imfreq_values(beta, stat, n_iw)
    res = np.zeros((n1, n2), dtype=np.complex128)
    F = lambda amp, bb: xi * amp / (-xi + math.exp(-beta * bb))  # :84
    model = sum(a[i][None, :, :] / (iw[:, None, None] - b[i]) for i in range(3))
    r = np.sum(gw - model, axis=0)  # :110
    full = (r / beta + m1 + F(a[0], b[0]) + F(a[1], b[1]) + F(a[2], b[2])) * (-xi)  # :111-112
    for i in range(n1):
        for j in range(i, n2):
            res[i, j] = full[i, j]
            if j > i:
                res[j, i] = np.conj(full[i, j])  # :115
    return (res, info) if return_info else res


# ----------------------------------------------------------------------------------------------
# lattice transforms
# ----------------------------------------------------------------------------------------------


def fourier_lattice_r_to_k(gr: np.ndarray, dims) -> np.ndarray:
    """cyclat -> brzone: FFTW_BACKWARD rank 3, no scaling.  gr [N_r, n_others], C-order r index
    i0*d1*d2 + i1*d2 + i2.  c++/triqs/gfs/transform/fourier_lattice.cpp:30-47,59."""
    d0, d1, d2 = dims
    n_others = gr.shape[1]
    x = np.asarray(gr, dtype=np.complex128).reshape(d0, d1, d2, n_others)
    y = np.fft.ifftn(x, axes=(0, 1, 2)) * (d0 * d1 * d2)
    return y.reshape(d0 * d1 * d2, n_others)


def fourier_lattice_k_to_r(gk: np.ndarray, dims) -> np.ndarray:
    """brzone -> cyclat: FFTW_FORWARD then /= N_k.  fourier_lattice.cpp:51-55."""
    d0, d1, d2 = dims
    n_others = gk.shape[1]
    x = np.asarray(gk, dtype=np.complex128).reshape(d0, d1, d2, n_others)
    y = np.fft.fftn(x, axes=(0, 1, 2)) / (d0 * d1 * d2)
    return y.reshape(d0 * d1 * d2, n_others)


def fft_many(x: np.ndarray, dims, sign: int) -> np.ndarray:
    """_fourier_base: unnormalised rank-len(dims) DFT over the leading (flattened) axis of
    [prod(dims), n_others]; sign=+1 is FFTW_BACKWARD.  fourier_common.cpp:25-45."""
    n_others = x.shape[1]
    shp = tuple(dims) + (n_others,)
    axes = tuple(range(len(dims)))
    xx = np.asarray(x, dtype=np.complex128).reshape(shp)
    if sign > 0:
        y = np.fft.ifftn(xx, axes=axes) * int(np.prod(dims))
    else:
        y = np.fft.fftn(xx, axes=axes)
    return y.reshape(x.shape)


# ----------------------------------------------------------------------------------------------
# Dyson k-sum and inversion
# ----------------------------------------------------------------------------------------------


def invert_batched(a: np.ndarray) -> np.ndarray:
    """invert_in_place on [n_pts, n, n].  c++/triqs/gfs/functions/functions2.hpp:197-201."""
    return np.linalg.inv(np.asarray(a, dtype=np.complex128))


def dyson_ksum(eps_k: np.ndarray, sigma: np.ndarray, beta: float, stat: int, mu: float, weights: np.ndarray | None = None,
               k_chunk: int = 4096) -> np.ndarray:
    """G_loc(iw_n) = sum_k w_k [(iw_n + mu) 1 - eps_k - Sigma(iw_n)]^-1.
    eps_k [n_k, n, n]; sigma [n_w, n, n] on the full imfreq mesh (n_w = 2 n_iw fermion).
    python/triqs/sumk/sumk_discrete.py:140-166 (weights default 1/N_k, :58)."""
    n_k, n, _ = eps_k.shape
    n_w = sigma.shape[0]
    n_iw = n_w // 2 if stat == FERMION else (n_w + 1) // 2
    iw = imfreq_values(beta, stat, n_iw)
    if weights is None:
        weights = np.full(n_k, 1.0 / n_k)
    eye = np.eye(n)
    tmp = (iw[:, None, None] + mu) * eye[None] - sigma  # [n_w, n, n]
    G = np.zeros((n_w, n, n), dtype=np.complex128)
    for k0 in range(0, n_k, k_chunk):
        e = eps_k[k0:k0 + k_chunk]
        m = tmp[None, :, :, :] - e[:, None, :, :]
        inv = np.linalg.inv(m)
        G += np.einsum("k,kwij->wij", weights[k0:k0 + k_chunk], inv)
    return G


def slice_bounds(n: int, size: int, rank: int):
    """k-point slice of rank `rank` out of `size`: python/triqs/utility/mpi_mpi4py.py:96-109
    (first n % size ranks get one extra)."""
    q, r = divmod(n, size)
    lo = rank * q + min(rank, r)
    hi = lo + q + (1 if rank < r else 0)
    return lo, hi


# ----------------------------------------------------------------------------------------------
# off-mesh evaluation
# ----------------------------------------------------------------------------------------------


def interp_tau(table: np.ndarray, beta: float, taus: np.ndarray) -> np.ndarray:
    """Linear interpolation of G(tau) stored as [n_tau, n_elem] at arbitrary tau in [0, beta].
    c++/triqs/mesh/bases/linear.hpp:221-228 via c++/triqs/gfs/evaluator.hpp:35-43."""
    n_tau = table.shape[0]
    delta = (beta - 0.0) / (n_tau - 1)
    delta_inv = 1.0 / delta
    x = np.maximum(np.asarray(taus, dtype=np.float64), 0.0)
    a = (x - 0.0) * delta_inv
    i = np.minimum(a.astype(np.int64), n_tau - 2)
    w = np.minimum(a - i, 1.0)
    return (1 - w)[:, None] * table[i] + w[:, None] * table[i + 1]


def evaluate_prod(table: np.ndarray, axes, coords: np.ndarray) -> np.ndarray:
    """Off-mesh evaluation on a product mesh: c++/triqs/mesh/evaluate.hpp:97-114 (recursion over the component meshes, the first
    component innermost) with linear::evaluate (c++/triqs/mesh/bases/linear.hpp:221-228), brzone1d (c++/triqs/mesh/brzone.hpp:355-359)
    and index pass-through (evaluate.hpp:73).  table [s_0, ..., s_{d-1}, n_elem]; axes[d] = ("linear", xmin, xmax) | "periodic" | "index";
    coords [n_pts, d] (periodic: in units of the mesh index, i.e. after transpose(units_inv) * k of brzone.hpp:368)."""
    table = np.asarray(table, dtype=np.complex128)
    coords = np.asarray(coords, dtype=np.float64)
    d = len(axes)
    n_pts = coords.shape[0]
    lo, hi, w = [], [], []
    for j, a in enumerate(axes):
        n = table.shape[j]
        x = coords[:, j]
        if a == "periodic":
            i = np.floor(x).astype(np.int64)
            w.append(x - i)
            lo.append(np.mod(i, n))
            hi.append(np.zeros_like(i) if n == 1 else np.mod(i + 1, n))
        elif a == "index":
            i = np.rint(x).astype(np.int64)
            lo.append(i)
            hi.append(i)
            w.append(None)
        else:
            xmin, xmax = float(a[1]), float(a[2])
            delta_inv = 1.0 / ((xmax - xmin) / (n - 1))
            xc = np.maximum(x, xmin)
            aa = (xc - xmin) * delta_inv
            i = np.minimum(aa.astype(np.int64), n - 2)
            lo.append(i)
            hi.append(i + 1)
            w.append(np.minimum(aa - i, 1.0))

    def rec(j, idx):
        # value with components j.. (j = d - 1 is the outermost sum) still to be combined; idx: chosen indices of components > j
        if j < 0:
            return table[tuple(idx[k] for k in range(d))]
        if w[j] is None:
            return rec(j - 1, {**idx, j: lo[j]})
        A, B = rec(j - 1, {**idx, j: lo[j]}), rec(j - 1, {**idx, j: hi[j]})
        return (1 - w[j])[:, None] * A + w[j][:, None] * B

    return rec(d - 1, {}).reshape(n_pts, table.shape[-1])
