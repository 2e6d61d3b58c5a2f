"""triqs_b200 -- B200-native (sm_100a) implementation of the TRIQS Green's-function hot path.

Only the data-parallel path is here: imtime <-> imfreq transforms with tail handling, tail fits,
cyclat <-> brzone transforms, the lattice Dyson k-sum and off-mesh G(tau) evaluation.  Everything
computes in ``libtqgpu.so`` (hand-written CUDA, C ABI in ``include/tqgpu.h``); there is no CPU path.
"""
from ._lib import (TQ_BOSON, TQ_FERMION, TQ_WARN_NOISY_TAU, TQ_WARN_SHORT_TAU_MESH, TQ_WARN_TAIL_ERR, LIB_PATH, TqLibraryMissing)
from .context import BOSON, FERMION, Context, TqError, TriqsRuntimeError, comm_unique_id, plan_describe, tail_fitter_setup

__all__ = ["Context", "TriqsRuntimeError", "TqError", "TqLibraryMissing", "FERMION", "BOSON", "comm_unique_id", "plan_describe", "tail_fitter_setup",
           "LIB_PATH"]
