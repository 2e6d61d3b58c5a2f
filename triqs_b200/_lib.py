"""ctypes binding of libtqgpu.so (the C ABI declared in include/tqgpu.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is present, importing
works (so that CPU-only tooling can inspect the package) but creating a Context raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_long, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TQ_LIB_PATH", os.path.join(_HERE, "libtqgpu.so"))  # (override: A/B builds of the kernels)

TQ_OK = 0
TQ_ERR_CUDA, TQ_ERR_INVALID, TQ_ERR_RUNTIME, TQ_ERR_UNSUPPORTED, TQ_ERR_NO_DEVICE, TQ_ERR_COMM = -1, -2, -3, -4, -5, -6
TQ_BOSON, TQ_FERMION = 0, 1
TQ_WARN_NOISY_TAU, TQ_WARN_SHORT_TAU_MESH, TQ_WARN_TAIL_ERR = 1, 2, 4
TQ_MAX_MOMENTS = 10
TQ_MAX_MATRIX_DIM = 64
TQ_COMM_ID_BYTES = 128
TQ_PROD_MAX_DIMS = 4
TQ_AXIS_LINEAR, TQ_AXIS_PERIODIC, TQ_AXIS_INDEX = 0, 1, 2


class TqAxis(ctypes.Structure):
    """tq_axis of include/tqgpu.h"""
    _fields_ = [("kind", c_int), ("size", c_long), ("xmin", c_double), ("xmax", c_double)]


# every symbol include/tqgpu.h declares: (restype, argtypes)
SIGNATURES = {
    "tq_create": (c_int, [c_int, POINTER(c_void_p)]),
    "tq_destroy": (None, [c_void_p]),
    "tq_last_error": (c_char_p, [c_void_p]),
    "tq_stream": (c_void_p, [c_void_p]),
    "tq_synchronize": (c_int, [c_void_p]),
    "tq_stream_wait_external": (c_int, [c_void_p, c_void_p]),
    "tq_stream_signal_external": (c_int, [c_void_p, c_void_p]),
    "tq_set_option": (c_int, [c_void_p, c_char_p, c_long]),
    "tq_plan_describe": (c_int, [c_long, c_long, c_int, c_long, c_long, ctypes.c_char_p, c_size_t]),
    "tq_launch_count": (c_uint64, [c_void_p]),
    "tq_version": (c_char_p, []),
    "tq_profile_enable": (c_int, [c_void_p, c_int]),
    "tq_profile_report": (c_int, [c_void_p, ctypes.c_char_p, c_size_t]),
    "tq_fourier_tau_to_iw": (c_int, [c_void_p, c_double, c_int, c_long, c_long, c_long, c_long, c_void_p, c_void_p, c_int, c_void_p,
                                     POINTER(c_int)]),
    "tq_fourier_tau_to_iw_axis": (c_int, [c_void_p, c_double, c_int, c_long, c_long, c_long, c_long, c_void_p, c_void_p, c_int, c_void_p,
                                          POINTER(c_int)]),
    "tq_fourier_iw_to_tau": (c_int, [c_void_p, c_double, c_int, c_long, c_long, c_long, c_long, c_void_p, c_void_p, c_int, c_void_p,
                                     POINTER(c_int), POINTER(c_double)]),
    "tq_fourier_tau_to_iw_real": (c_int, [c_void_p, c_double, c_int, c_long, c_long, c_long, c_long, c_void_p, c_void_p, c_int, c_void_p,
                                          POINTER(c_int)]),
    "tq_fourier_iw_to_tau_real": (c_int, [c_void_p, c_double, c_int, c_long, c_long, c_long, c_long, c_void_p, c_void_p, c_int, c_void_p,
                                          POINTER(c_int), POINTER(c_double), POINTER(c_double)]),
    "tq_density": (c_int, [c_void_p, c_double, c_int, c_long, c_int, c_long, c_void_p, c_void_p, c_int, c_void_p, POINTER(c_int),
                           POINTER(c_double)]),
    "tq_fourier_t_to_w": (c_int, [c_void_p, c_double, c_double, c_double, c_double, c_long, c_long, c_long, c_void_p, c_void_p, c_int, c_void_p]),
    "tq_fourier_w_to_t": (c_int, [c_void_p, c_double, c_double, c_double, c_double, c_long, c_long, c_long, c_void_p, c_void_p, c_int, c_void_p]),
    "tq_legendre_to_imfreq": (c_int, [c_void_p, c_int, c_long, c_long, c_long, c_long, c_void_p, c_void_p]),
    "tq_legendre_to_imtime": (c_int, [c_void_p, c_double, c_long, c_long, c_long, c_long, c_void_p, c_void_p]),
    "tq_imtime_to_legendre": (c_int, [c_void_p, c_double, c_long, c_long, c_long, c_long, c_void_p, c_void_p]),
    "tq_imfreq_to_legendre": (c_int, [c_void_p, c_double, c_int, c_long, c_long, c_long, c_long, c_void_p, c_void_p]),
    "tq_tight_binding_fourier": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, POINTER(c_long), c_void_p]),
    "tq_make_hermitian": (c_int, [c_void_p, c_int, c_long, c_int, c_long, c_void_p, c_void_p]),
    "tq_is_gf_hermitian": (c_int, [c_void_p, c_int, c_long, c_int, c_long, c_void_p, c_double, POINTER(c_int)]),
    "tq_replace_by_tail": (c_int, [c_void_p, c_double, c_int, c_long, c_long, c_long, c_void_p, c_void_p, c_int, c_int]),
    "tq_fit_tail": (c_int, [c_void_p, c_double, c_int, c_long, c_double, c_int, c_int, c_int, c_int, c_long, c_long, c_void_p, c_void_p,
                            c_int, c_void_p, POINTER(c_int), POINTER(c_double)]),
    "tq_fit_tail_refreq": (c_int, [c_void_p, c_double, c_double, c_long, c_double, c_int, c_int, c_long, c_long, c_void_p, c_void_p, c_int,
                                   c_void_p, POINTER(c_int), POINTER(c_double)]),
    "tq_tail_fitter_setup_host": (c_int, [c_double, c_int, c_long, c_double, c_int, c_int, c_int, c_int, POINTER(c_int), POINTER(c_long),
                                          POINTER(c_int), c_void_p, POINTER(c_double), ctypes.c_char_p, c_size_t]),
    "tq_fft_many": (c_int, [c_void_p, c_int, POINTER(c_int), c_long, c_void_p, c_void_p, c_int]),
    "tq_fourier_lattice": (c_int, [c_void_p, POINTER(c_long), c_int, c_long, c_void_p, c_void_p]),
    "tq_invert_batched": (c_int, [c_void_p, c_int, c_long, c_void_p]),
    "tq_dyson_ksum": (c_int, [c_void_p, c_int, c_long, c_long, c_double, c_int, c_double, c_void_p, c_void_p, c_double, c_void_p, c_void_p,
                              c_int]),
    "tq_interp_tau": (c_int, [c_void_p, c_double, c_long, c_long, c_void_p, c_long, c_void_p, c_void_p]),
    "tq_evaluate_imfreq": (c_int, [c_void_p, c_double, c_int, c_long, c_long, c_long, c_void_p, c_long, c_void_p, c_void_p, POINTER(c_double)]),
    "tq_evaluate_prod": (c_int, [c_void_p, c_int, POINTER(TqAxis), c_long, c_void_p, c_long, c_void_p, c_void_p]),
    "tq_comm_unique_id": (c_int, [c_void_p]),
    "tq_comm_init": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "tq_all_reduce_sum": (c_int, [c_void_p, c_void_p, c_size_t]),
    "tq_device_alloc": (c_int, [c_void_p, c_size_t, POINTER(c_void_p)]),
    "tq_device_free": (c_int, [c_void_p, c_void_p]),
    "tq_memcpy_h2d": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t]),
    "tq_memcpy_d2h": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t]),
}

_lib = None


class TqLibraryMissing(RuntimeError):
    pass


def load():
    """dlopen libtqgpu.so and attach the prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TqLibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(triqs_b200 has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
