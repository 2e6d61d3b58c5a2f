"""Thin, array-level Python face of the C ABI (include/tqgpu.h).

Arrays are either NumPy arrays (host memory: the library stages them through the GPU inside the call)
or torch CUDA tensors of dtype complex128/float64 (device memory: used in place, results are torch
tensors on the same device).  Shapes follow the reference's flattened views ``[batch][mesh][n_others]``.
PyTorch is used only as the device-memory owner; all compute happens in libtqgpu.so.
"""
from __future__ import annotations

import ctypes
import sys
import threading
import warnings

import numpy as np

from . import _lib as L

FERMION, BOSON = L.TQ_FERMION, L.TQ_BOSON


class TriqsRuntimeError(RuntimeError):
    """Raised where the reference throws triqs::runtime_error (c++/triqs/utility/exceptions.hpp:31-37)."""


class TqError(RuntimeError):
    pass


# messages the reference prints to std::cerr (c++/triqs/gfs/transform/fourier_matsubara.cpp:109-110, 131-133, 208-210)
WARNING_TEXT = {
    L.TQ_WARN_NOISY_TAU: "WARNING: Direct Fourier cannot deduce the high-frequency moments of G(tau) due to noise or a coarse tau-grid. "
                         "Please specify the high-frequency moments for higher accuracy.\n",
    L.TQ_WARN_SHORT_TAU_MESH: "[Direct Fourier] WARNING: The imaginary time mesh is less than six times as long as the number of positive "
                              "frequencies.\nThis can lead to substantial numerical inaccuracies at the boundary of the frequency mesh.\n",
    L.TQ_WARN_TAIL_ERR: "WARNING: High frequency moments have an error greater than 1e-4.\n"
                        " Please make sure you treat the constant offset analytically!\n",
}


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _as_c128(x):
    if _is_torch(x):
        import torch
        if x.dtype != torch.complex128:
            x = x.to(torch.complex128)
        return x.contiguous()
    return np.ascontiguousarray(x, dtype=np.complex128)


def _is_real_dtype(x):
    """real-valued data (gf<mesh, T::real_t>): float64 / float32 NumPy arrays or torch tensors"""
    if _is_torch(x):
        return not x.dtype.is_complex
    return not np.iscomplexobj(x)


def _as_f64(x):
    if _is_torch(x):
        import torch
        return x.to(torch.float64).contiguous()
    return np.ascontiguousarray(x, dtype=np.float64)


_tls = threading.local()  # set by _ptr when an argument is a torch CUDA tensor; read and cleared by Context._run


def _ptr(x):
    if x is None:
        return None
    if _is_torch(x):
        if x.is_cuda:
            _tls.device_args = True
        return ctypes.c_void_p(x.data_ptr())
    return ctypes.c_void_p(x.ctypes.data)


def _empty_like_kind(ref, shape, dtype=np.complex128):
    if _is_torch(ref):
        import torch
        td = torch.complex128 if dtype == np.complex128 else torch.float64
        return torch.empty(shape, dtype=td, device=ref.device)
    return np.empty(shape, dtype=dtype)


def plan_describe(n_tau, n_iw, statistic, n_cols=1, batch=1) -> dict:
    """tq_plan_describe: which engine an imtime <-> imfreq transform of this shape takes (host only, works without a GPU)."""
    import json
    lib = L.load()
    buf = ctypes.create_string_buffer(1024)
    st = lib.tq_plan_describe(int(n_tau), int(n_iw), int(statistic), int(n_cols), int(batch), buf, len(buf))
    if st != 0:
        raise ValueError("tq_plan_describe: invalid argument")
    return json.loads(buf.value.decode())


class Context:
    """One per process/GPU: owns the stream, the plan caches and (optionally) the NCCL communicator."""

    def __init__(self, device: int = 0, print_warnings: bool = True, torch_ordering: bool = True):
        self._lib = L.load()  # raises if libtqgpu.so is missing -- no CPU fallback
        h = ctypes.c_void_p()
        st = self._lib.tq_create(device, ctypes.byref(h))
        if st == L.TQ_ERR_NO_DEVICE:
            raise TqError("tq_create: no CUDA device -- triqs_b200 never computes on the CPU")
        if st != L.TQ_OK:
            raise TqError(f"tq_create failed with status {st}")
        self._h = h
        self.device = device
        self.print_warnings = print_warnings
        self.torch_ordering = torch_ordering  # False: the caller orders the streams itself (tq_synchronize / events)
        self.last_warn = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.tq_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing -----------------------------------------------------------------------------
    def _check(self, st):
        if st == L.TQ_OK:
            return
        msg = self._lib.tq_last_error(self._h).decode(errors="replace")
        if st == L.TQ_ERR_RUNTIME:
            raise TriqsRuntimeError(msg)
        raise TqError(f"status {st}: {msg}")

    def _emit(self, warn):
        self.last_warn = warn
        if not self.print_warnings:
            return
        bits = int(np.bitwise_or.reduce(warn)) if len(warn) else 0
        for b, text in WARNING_TEXT.items():
            if bits & b:
                sys.stderr.write(text)

    def _run(self, name, *args):
        """One library call.  When an argument is a torch CUDA tensor the call is asynchronous on the context's own
        (non-blocking) stream: it is ordered after the work already enqueued on torch's current stream, and torch's current
        stream is ordered after it (include/tqgpu.h "Stream-ordering contract"), so torch producers / consumers on either side
        need no manual synchronisation."""
        fn = getattr(self._lib, name)
        device_args = getattr(_tls, "device_args", False)
        _tls.device_args = False
        if not (device_args and self.torch_ordering):
            return fn(self._h, *args)
        import torch
        other = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        self._check(self._lib.tq_stream_wait_external(self._h, other))
        st = fn(self._h, *args)
        self._check(self._lib.tq_stream_signal_external(self._h, other))
        return st

    def synchronize(self):
        self._check(self._lib.tq_synchronize(self._h))

    def set_option(self, name: str, value: int):
        """tq_set_option: "no_pipe" (generic engine only), "scratch_chunk_mb" -- neither changes results."""
        self._check(self._lib.tq_set_option(self._h, name.encode(), int(value)))

    @property
    def stream(self):
        return self._lib.tq_stream(self._h)

    def launch_count(self) -> int:
        return int(self._lib.tq_launch_count(self._h))

    def profile_enable(self, on: bool = True):
        self._check(self._lib.tq_profile_enable(self._h, int(on)))

    def profile_report(self) -> dict:
        """per-kernel {"tag": {"launches": n, "ms": total}} since the last report (CUDA events on the ctx stream)."""
        import json
        buf = ctypes.create_string_buffer(1 << 16)
        self._check(self._lib.tq_profile_report(self._h, buf, len(buf)))
        return json.loads(buf.value.decode())

    # -- Matsubara transforms -------------------------------------------------------------------
    def fourier_tau_to_iw(self, gt, beta, statistic, n_iw, moments=None, out=None, one_gf=False):
        """gt [batch, n_tau, n_others] -> gw [batch, n_w, n_others]  (tq_fourier_tau_to_iw).
        one_gf: the leading index is an outer mesh of ONE gf (tq_fourier_tau_to_iw_axis): joint noise check, one warning.
        A real-valued gt (float64) goes through tq_fourier_tau_to_iw_real: only the reals cross PCIe, the promotion to complex
        (fourier.hpp:78-81) happens on the device."""
        real_in = _is_real_dtype(gt) and not one_gf
        gt = _as_f64(gt) if real_in else _as_c128(gt)
        batch, n_tau, n_others = gt.shape
        n_w = 2 * n_iw if statistic == FERMION else 2 * n_iw - 1
        n_mom = 0
        if moments is not None:
            moments = _as_c128(moments)
            n_mom = moments.shape[1]
        gw = out if out is not None else _empty_like_kind(gt, (batch, n_w, n_others))
        warn = (ctypes.c_int * (1 if one_gf else batch))()
        if real_in:
            self._check(self._run("tq_fourier_tau_to_iw_real", beta, statistic, n_tau, n_iw, n_others, batch, _ptr(gt), _ptr(moments),
                                  n_mom, _ptr(gw), warn))
        elif one_gf:
            self._check(self._run("tq_fourier_tau_to_iw_axis", beta, statistic, n_tau, n_iw, batch, n_others, _ptr(gt), _ptr(moments),
                                                            n_mom, _ptr(gw), warn))
        else:
            self._check(self._run("tq_fourier_tau_to_iw", beta, statistic, n_tau, n_iw, n_others, batch, _ptr(gt), _ptr(moments),
                                                       n_mom, _ptr(gw), warn))
        self._emit(np.frombuffer(warn, dtype=np.int32))
        return gw

    def fourier_iw_to_tau(self, gw, beta, statistic, n_tau, moments=None, out=None, return_err=False, real_out=False):
        """gw [batch, n_w, n_others] (full mesh) -> gt [batch, n_tau, n_others]  (tq_fourier_iw_to_tau).
        real_out (or a float64 `out`): tq_fourier_iw_to_tau_real -- gt receives Re G(tau) as float64 (real(g), functions2.hpp:242)
        and `last_max_imag` [batch] max |Im G(tau)|, what is_gf_real (functions2.hpp:223) compares with its tolerance."""
        gw = _as_c128(gw)
        batch, n_w, n_others = gw.shape
        if statistic == FERMION:
            if n_w % 2:
                raise TriqsRuntimeError("Fourier is only implemented for g(i omega_n) with full mesh (positive and negative frequencies)")
            n_iw = n_w // 2
        else:
            if n_w % 2 == 0:
                raise TriqsRuntimeError("Fourier is only implemented for g(i omega_n) with full mesh (positive and negative frequencies)")
            n_iw = (n_w + 1) // 2
        n_mom = 0
        if moments is not None:
            moments = _as_c128(moments)
            n_mom = moments.shape[1]
        real_out = real_out or (out is not None and _is_real_dtype(out))
        if real_out and out is not None:
            if not _is_real_dtype(out) or tuple(out.shape) != (batch, n_tau, n_others) or not (
                    out.is_contiguous() if _is_torch(out) else (out.flags.c_contiguous and out.dtype == np.float64)):
                raise ValueError("real_out: `out` must be a contiguous float64 [batch, n_tau, n_others] array")
        gt = out if out is not None else _empty_like_kind(gw, (batch, n_tau, n_others), np.float64 if real_out else np.complex128)
        warn = (ctypes.c_int * batch)()
        err = (ctypes.c_double * batch)()
        if real_out:
            mi = (ctypes.c_double * batch)()
            self._check(self._run("tq_fourier_iw_to_tau_real", beta, statistic, n_iw, n_tau, n_others, batch, _ptr(gw), _ptr(moments), n_mom,
                                  _ptr(gt), warn, err, mi))
            self.last_max_imag = np.frombuffer(mi, dtype=np.float64).copy()
        else:
            self._check(self._run("tq_fourier_iw_to_tau", beta, statistic, n_iw, n_tau, n_others, batch, _ptr(gw), _ptr(moments), n_mom,
                                  _ptr(gt), warn, err))
        self._emit(np.frombuffer(warn, dtype=np.int32))
        if return_err:
            return gt, np.frombuffer(err, dtype=np.float64).copy()
        return gt

    # -- retime <-> refreq (fourier_real.cpp:35-131) ---------------------------------------------
    def _fourier_real(self, fn, g, t_mesh, w_mesh, moments, out):
        g = _as_c128(g)
        batch, n_pts, n_others = g.shape
        n_mom = 0
        if moments is not None:
            moments = _as_c128(moments)
            n_mom = moments.shape[1]
        res = out if out is not None else _empty_like_kind(g, (batch, n_pts, n_others))
        self._check(self._run(fn, float(t_mesh[0]), float(t_mesh[1]), float(w_mesh[0]), float(w_mesh[1]), n_pts, n_others, batch, _ptr(g),
                       _ptr(moments), n_mom, _ptr(res)))
        return res

    def fourier_t_to_w(self, gt, t_mesh, w_mesh, moments=None, out=None):
        """gt [batch, n, n_others] on retime(t_min, t_max, n) -> gw on refreq(w_min, w_max, n)  (tq_fourier_t_to_w)."""
        return self._fourier_real("tq_fourier_t_to_w", gt, t_mesh, w_mesh, moments, out)

    def fourier_w_to_t(self, gw, t_mesh, w_mesh, moments=None, out=None):
        """gw [batch, n, n_others] on refreq(w_min, w_max, n) -> gt on retime(t_min, t_max, n)  (tq_fourier_w_to_t)."""
        return self._fourier_real("tq_fourier_w_to_t", gw, t_mesh, w_mesh, moments, out)

    # -- Legendre <-> Matsubara (legendre_matsubara.hpp:37-129) ------------------------------------
    def legendre_to_imfreq(self, gl, statistic, n_iw, out=None):
        """gl [batch, n_l, n_others] -> gw [batch, n_w, n_others] on the full imfreq mesh."""
        gl = _as_c128(gl)
        batch, n_l, n_others = gl.shape
        n_w = 2 * n_iw if statistic == FERMION else 2 * n_iw - 1
        res = out if out is not None else _empty_like_kind(gl, (batch, n_w, n_others))
        self._check(self._run("tq_legendre_to_imfreq", statistic, n_l, n_iw, n_others, batch, _ptr(gl), _ptr(res)))
        return res

    def legendre_to_imtime(self, gl, beta, n_tau, out=None):
        gl = _as_c128(gl)
        batch, n_l, n_others = gl.shape
        res = out if out is not None else _empty_like_kind(gl, (batch, n_tau, n_others))
        self._check(self._run("tq_legendre_to_imtime", beta, n_l, n_tau, n_others, batch, _ptr(gl), _ptr(res)))
        return res

    def imtime_to_legendre(self, gt, beta, n_l, out=None):
        gt = _as_c128(gt)
        batch, n_tau, n_others = gt.shape
        res = out if out is not None else _empty_like_kind(gt, (batch, n_l, n_others))
        self._check(self._run("tq_imtime_to_legendre", beta, n_tau, n_l, n_others, batch, _ptr(gt), _ptr(res)))
        return res

    def imfreq_to_legendre(self, gw, beta, statistic, n_l, out=None):
        gw = _as_c128(gw)
        batch, n_w, n_others = gw.shape
        n_iw = n_w // 2 if statistic == FERMION else (n_w + 1) // 2
        res = out if out is not None else _empty_like_kind(gw, (batch, n_l, n_others))
        self._check(self._run("tq_imfreq_to_legendre", beta, statistic, n_iw, n_l, n_others, batch, _ptr(gw), _ptr(res)))
        return res

    # -- density ---------------------------------------------------------------------------------
    def density(self, gw, beta, statistic, moments=None, out=None, return_err=False):
        """gw [batch, n_w, n, n] (full mesh) -> n [batch, n, n]  (tq_density; density.cpp:32-120)."""
        gw = _as_c128(gw)
        batch, n_w, n, n2 = gw.shape
        if n != n2:
            raise TriqsRuntimeError("density: the target must be square")
        if (statistic == FERMION) == bool(n_w % 2):
            raise TriqsRuntimeError("density is only implemented for g(i omega_n) with full mesh (positive and negative frequencies)")
        n_iw = n_w // 2 if statistic == FERMION else (n_w + 1) // 2
        n_mom = 0
        if moments is not None:
            moments = _as_c128(moments)
            n_mom = moments.shape[1]
        res = out if out is not None else _empty_like_kind(gw, (batch, n, n))
        warn = (ctypes.c_int * batch)()
        err = (ctypes.c_double * batch)()
        self._check(self._run("tq_density", beta, statistic, n_iw, n, batch, _ptr(gw), _ptr(moments), n_mom, _ptr(res), warn, err))
        self._emit(np.frombuffer(warn, dtype=np.int32))
        if return_err:
            return res, np.frombuffer(err, dtype=np.float64).copy()
        return res

    # -- tight-binding dispersion (lattice/tight_binding.hpp:88-123) ------------------------------
    def tight_binding_fourier(self, displ, hop, dims, out=None):
        """eps_k [d0*d1*d2, n, n] = sum_j hop[j] exp(2 pi i k.displ[j]);  displ [n_r, 3] ints (host), hop [n_r, n, n]."""
        displ = np.ascontiguousarray(np.asarray(displ, dtype=np.int32).reshape(-1, 3))
        hop = _as_c128(hop)
        n_r, n, _ = hop.shape
        d = (ctypes.c_long * 3)(*[int(x) for x in dims])
        nk = int(dims[0]) * int(dims[1]) * int(dims[2])
        res = out if out is not None else _empty_like_kind(hop, (nk, n, n))
        self._check(self._run("tq_tight_binding_fourier", n, n_r, displ.ctypes.data_as(ctypes.c_void_p), _ptr(hop), d, _ptr(res)))
        return res

    # -- hermiticity helpers, replace_by_tail (imfreq.hpp:81-215, 258-261) ----------------------
    def make_hermitian(self, g, imtime=False, out=None):
        """g [batch, n_pts, d, d] -> 0.5 (g[w] + dagger(g[-w]))  (full imfreq mesh, or imtime: the point itself)."""
        g = _as_c128(g)
        batch, n_pts, d, d2 = g.shape
        res = out if out is not None else _empty_like_kind(g, g.shape)
        self._check(self._run("tq_make_hermitian", 1 if imtime else 0, n_pts, d, batch, _ptr(g), _ptr(res)))
        return res

    def is_gf_hermitian(self, g, imtime=False, tolerance=1e-12):
        g = _as_c128(g)
        batch, n_pts, d, d2 = g.shape
        r = ctypes.c_int(0)
        self._check(self._run("tq_is_gf_hermitian", 1 if imtime else 0, n_pts, d, batch, _ptr(g), float(tolerance), ctypes.byref(r)))
        return bool(r.value)

    def replace_by_tail(self, g, beta, statistic, tail, n_min):
        """in place: g [batch, n_w, n_cols], tail [batch, n_moments, n_cols]."""
        batch, n_w, n_cols = g.shape
        n_iw = n_w // 2 if statistic == FERMION else (n_w + 1) // 2
        tail = _as_c128(tail)
        self._check(self._run("tq_replace_by_tail", beta, statistic, n_iw, n_cols, batch, _ptr(g), _ptr(tail), tail.shape[1], int(n_min)))
        return g

    # -- tail fit ----------------------------------------------------------------------------
    def fit_tail(self, g, beta, statistic, known_moments=None, hermitian=False, inner_dim=1, tail_fraction=0.2, n_tail_max=30,
                 expansion_order=None):
        """g [batch, n_w, n_cols] -> (moments [batch, n_moments, n_cols], err [batch])  (tq_fit_tail)."""
        g = _as_c128(g)
        batch, n_w, n_cols = g.shape
        n_iw = n_w // 2 if statistic == FERMION else (n_w + 1) // 2
        n_known = 0
        if known_moments is not None:
            known_moments = _as_c128(known_moments)
            n_known = known_moments.shape[1]
            if known_moments.shape[2] != n_cols:
                raise TriqsRuntimeError("known_moments shape incompatible with shape of data")
        out = _empty_like_kind(g, (batch, max(L.TQ_MAX_MOMENTS, n_known), n_cols))
        nm = ctypes.c_int(0)
        err = (ctypes.c_double * batch)()
        self._check(self._run("tq_fit_tail", beta, statistic, n_iw, tail_fraction, n_tail_max,
                                          -1 if expansion_order is None else int(expansion_order), int(bool(hermitian)), int(inner_dim), n_cols,
                                          batch, _ptr(g), _ptr(known_moments), n_known, _ptr(out), ctypes.byref(nm), err))
        n_moments = nm.value
        flat = out.reshape(-1)[: batch * n_moments * n_cols].reshape(batch, n_moments, n_cols)
        return flat, np.frombuffer(err, dtype=np.float64).copy()

    def fit_tail_refreq(self, g, w_mesh, known_moments=None, tail_fraction=0.2, n_tail_max=30, expansion_order=None):
        """fit_tail of real-frequency gfs g [batch, n_w, n_cols] on refreq(w_min, w_max, n_w)  (tq_fit_tail_refreq)."""
        g = _as_c128(g)
        batch, n_w, n_cols = g.shape
        n_known = 0
        if known_moments is not None:
            known_moments = _as_c128(known_moments)
            n_known = known_moments.shape[1]
            if known_moments.shape[2] != n_cols:
                raise TriqsRuntimeError("known_moments shape incompatible with shape of data")
        out = _empty_like_kind(g, (batch, max(L.TQ_MAX_MOMENTS, n_known), n_cols))
        nm = ctypes.c_int(0)
        err = (ctypes.c_double * batch)()
        self._check(self._run("tq_fit_tail_refreq", float(w_mesh[0]), float(w_mesh[1]), n_w, tail_fraction, n_tail_max,
                                                 -1 if expansion_order is None else int(expansion_order), n_cols, batch, _ptr(g),
                                                 _ptr(known_moments), n_known, _ptr(out), ctypes.byref(nm), err))
        n_moments = nm.value
        flat = out.reshape(-1)[: batch * n_moments * n_cols].reshape(batch, n_moments, n_cols)
        return flat, np.frombuffer(err, dtype=np.float64).copy()

    def tail_fitter_setup(self, beta, statistic, n_iw, n_fixed=0, hermitian=False, tail_fraction=0.2, n_tail_max=30, expansion_order=None):
        return tail_fitter_setup(beta, statistic, n_iw, n_fixed, hermitian, tail_fraction, n_tail_max, expansion_order)

    # -- plain DFT / lattice -------------------------------------------------------------------
    def fft_many(self, x, dims, sign, out=None):
        """x [prod(dims), howmany]; unnormalised; sign=+1 is FFTW_BACKWARD  (tq_fft_many)."""
        x = _as_c128(x)
        total, howmany = x.shape
        assert total == int(np.prod(dims))
        y = out if out is not None else _empty_like_kind(x, x.shape)
        d = (ctypes.c_int * len(dims))(*[int(v) for v in dims])
        self._check(self._run("tq_fft_many", len(dims), d, howmany, _ptr(x), _ptr(y), int(sign)))
        return y

    def fourier_lattice(self, x, dims, to_k: bool, out=None):
        """x [d0*d1*d2, n_others]; to_k: cyclat -> brzone, else brzone -> cyclat  (tq_fourier_lattice)."""
        x = _as_c128(x)
        nk, n_others = x.shape
        assert nk == int(np.prod(dims)) and len(dims) == 3
        y = out if out is not None else _empty_like_kind(x, x.shape)
        d = (ctypes.c_long * 3)(*[int(v) for v in dims])
        self._check(self._run("tq_fourier_lattice", d, int(bool(to_k)), n_others, _ptr(x), _ptr(y)))
        return y

    # -- inversion / Dyson ----------------------------------------------------------------------
    def invert_in_place(self, a):
        """a [n_pts, n, n] complex128 (C-contiguous), inverted in place  (tq_invert_batched)."""
        if not _is_torch(a):
            assert a.dtype == np.complex128 and a.flags.c_contiguous
        n_pts, n, n2 = a.shape
        assert n == n2
        self._check(self._run("tq_invert_batched", n, n_pts, _ptr(a)))
        return a

    def dyson_ksum(self, eps_k, sigma, beta, statistic, mu, weights=None, uniform_weight=None, all_reduce=False, out=None):
        """eps_k [n_k, n, n], sigma [n_w, n, n] -> G_loc [n_w, n, n]  (tq_dyson_ksum)."""
        eps_k = _as_c128(eps_k)
        sigma = _as_c128(sigma)
        n_k, n, _ = eps_k.shape
        n_w = sigma.shape[0]
        n_iw = n_w // 2 if statistic == FERMION else (n_w + 1) // 2
        if weights is not None:
            weights = _as_f64(weights)
        if uniform_weight is None:
            uniform_weight = 1.0 / max(n_k, 1)
        g = out if out is not None else _empty_like_kind(sigma, sigma.shape)
        self._check(self._run("tq_dyson_ksum", n, n_k, n_iw, beta, statistic, mu, _ptr(eps_k), _ptr(weights), float(uniform_weight),
                                            _ptr(sigma), _ptr(g), int(bool(all_reduce))))
        return g

    # -- interpolation ---------------------------------------------------------------------------
    def interp_tau(self, table, beta, taus, out=None):
        """table [n_tau, n_elem], taus [n_pts] -> [n_pts, n_elem]  (tq_interp_tau)."""
        table = _as_c128(table)
        taus = _as_f64(taus)
        n_tau, n_elem = table.shape
        n_pts = taus.shape[0]
        y = out if out is not None else _empty_like_kind(table, (n_pts, n_elem))
        self._check(self._run("tq_interp_tau", beta, n_tau, n_elem, _ptr(table), n_pts, _ptr(taus), _ptr(y)))
        return y

    def evaluate_imfreq(self, g, beta, statistic, ns, out=None, return_err=False):
        """g [batch, n_w, n_cols] (full mesh), ns: Matsubara indices (ints, any range) -> [batch, len(ns), n_cols]  (tq_evaluate_imfreq)."""
        g = _as_c128(g)
        batch, n_w, n_cols = g.shape
        n_iw = n_w // 2 if statistic == FERMION else (n_w + 1) // 2
        ns = np.ascontiguousarray(np.asarray(ns, dtype=np.int64).reshape(-1))
        res = out if out is not None else _empty_like_kind(g, (batch, len(ns), n_cols))
        err = (ctypes.c_double * batch)()
        self._check(self._run("tq_evaluate_imfreq", beta, statistic, n_iw, n_cols, batch, _ptr(g), len(ns), _ptr(ns), _ptr(res), err))
        if return_err:
            return res, np.frombuffer(err, dtype=np.float64).copy()
        return res

    def evaluate_prod(self, table, axes, coords, out=None):
        """Multi-linear interpolation on a product mesh (tq_evaluate_prod; c++/triqs/mesh/evaluate.hpp:97-114).
        table [size_0, ..., size_{d-1}, n_elem]; axes: one entry per component mesh --
          ("linear", xmin, xmax)   imtime / retime / refreq (linear.hpp:221-228), coordinate = the value,
          "periodic"               one brzone direction (brzone.hpp:355-359), coordinate in units of the mesh index,
          "index"                  pass-through (imfreq: n - first), coordinate = the data index;
        coords [n_pts, d] -> [n_pts, n_elem]."""
        table = _as_c128(table)
        coords = _as_f64(coords)
        n_dims = len(axes)
        if table.ndim != n_dims + 1 or coords.ndim != 2 or coords.shape[1] != n_dims:
            raise TqError("evaluate_prod: table must be [size_0, ..., size_{d-1}, n_elem] and coords [n_pts, d]")
        arr = (L.TqAxis * n_dims)()
        for d, a in enumerate(axes):
            if a == "periodic":
                arr[d] = L.TqAxis(L.TQ_AXIS_PERIODIC, table.shape[d], 0.0, 0.0)
            elif a == "index":
                arr[d] = L.TqAxis(L.TQ_AXIS_INDEX, table.shape[d], 0.0, 0.0)
            else:
                arr[d] = L.TqAxis(L.TQ_AXIS_LINEAR, table.shape[d], float(a[1]), float(a[2]))
        n_pts, n_elem = coords.shape[0], table.shape[-1]
        y = out if out is not None else _empty_like_kind(table, (n_pts, n_elem))
        self._check(self._run("tq_evaluate_prod", n_dims, arr, n_elem, _ptr(table), n_pts, _ptr(coords), _ptr(y)))
        return y

    # -- multi GPU ---------------------------------------------------------------------------------
    def comm_init(self, n_ranks: int, rank: int, unique_id: bytes):
        buf = (ctypes.c_ubyte * L.TQ_COMM_ID_BYTES).from_buffer_copy(unique_id)
        self._check(self._run("tq_comm_init", n_ranks, rank, buf))

    def all_reduce_sum(self, t):
        """in-place sum over ranks of a torch CUDA tensor (float64 or complex128)."""
        import torch
        assert _is_torch(t) and t.is_cuda and t.is_contiguous()
        count = t.numel() * (2 if t.dtype == torch.complex128 else 1)
        self._check(self._run("tq_all_reduce_sum", _ptr(t), count))
        return t


def comm_unique_id() -> bytes:
    lib = L.load()
    buf = (ctypes.c_ubyte * L.TQ_COMM_ID_BYTES)()
    st = lib.tq_comm_unique_id(buf)
    if st != L.TQ_OK:
        raise TqError("tq_comm_unique_id failed (libnccl.so.2 not loadable?)")
    return bytes(buf)


def tail_fitter_setup(beta, statistic, n_iw, n_fixed=0, hermitian=False, tail_fraction=0.2, n_tail_max=30, expansion_order=None):
    """Mesh-only part of the tail fit (host code, no GPU needed): fit indices, pseudo-inverse, singular values."""
    lib = L.load()
    n_idx = ctypes.c_int(0)
    n_var = ctypes.c_int(0)
    idx = (ctypes.c_long * (2 * n_tail_max))()
    pinv = np.zeros(L.TQ_MAX_MOMENTS * 4 * n_tail_max, dtype=np.complex128)
    sv = (ctypes.c_double * L.TQ_MAX_MOMENTS)()
    err = ctypes.create_string_buffer(512)
    st = lib.tq_tail_fitter_setup_host(beta, statistic, n_iw, tail_fraction, n_tail_max, -1 if expansion_order is None else expansion_order,
                                       n_fixed, int(bool(hermitian)), ctypes.byref(n_idx), idx, ctypes.byref(n_var),
                                       ctypes.c_void_p(pinv.ctypes.data), sv, err, 512)
    if st == L.TQ_ERR_RUNTIME:
        raise TriqsRuntimeError(err.value.decode())
    if st != L.TQ_OK:
        raise TqError(f"tq_tail_fitter_setup_host: status {st}")
    rows = n_idx.value * (2 if hermitian else 1)
    return {
        "idx": np.array(idx[: n_idx.value], dtype=np.int64),
        "n_var": n_var.value,
        "pinv": pinv[: n_var.value * rows].reshape(n_var.value, rows).copy(),
        "sv": np.array(sv[: n_var.value]),
    }
