"""Mesh-only tail-fit setup (host code of libtqgpu: indices, Vandermonde, Jacobi-SVD pseudo-inverse, order
choice) against the oracle's LAPACK-based restatement of tail_fitter.hpp:105-181.  No GPU needed."""
import numpy as np
import pytest

import triqs_b200
from oracle import triqs_oracle as O

CASES = [
    # beta, stat, n_iw, n_fixed, tail_fraction
    (10.0, O.FERMION, 1025, 0, 0.2),   # C1
    (10.0, O.FERMION, 1025, 1, 0.2),
    (100.0, O.FERMION, 10000, 1, 0.2),  # C2 (inverse transform default)
    (10.0, O.FERMION, 100, 1, 0.3),    # fit_tail_matsubara.cpp
    (10.0, O.FERMION, 100, 2, 0.3),
    (10.0, O.BOSON, 99, 2, 0.3),
    (1.0, O.FERMION, 25, 0, 0.2),      # tail_issues.py small meshes
    (1.0, O.FERMION, 6000, 2, 0.2),
    (50.0, O.FERMION, 2000, 1, 0.2),
]


@pytest.mark.parametrize("beta,stat,n_iw,n_fixed,frac", CASES)
@pytest.mark.parametrize("herm", [False, True])
def test_setup_matches_oracle(beta, stat, n_iw, n_fixed, frac, herm):
    s = triqs_b200.tail_fitter_setup(beta, stat, n_iw, n_fixed=n_fixed, hermitian=herm, tail_fraction=frac)
    f = O.TailFitter(beta, stat, n_iw, tail_fraction=frac)
    lss = f.setup_lss(n_fixed, herm)
    # bit-exact integer work: the fit indices and the chosen order
    assert list(s["idx"]) == list(f.fit_idx)
    assert s["n_var"] == lss["n_var"]
    # singular values: Jacobi and LAPACK agree to high relative accuracy
    assert np.max(np.abs(s["sv"] - lss["s"]) / lss["s"]) < 1e-9
    # pseudo-inverse: P A = 1, and agreement with LAPACK's within the conditioning of A
    A = lss["A"] if not herm else np.vstack([lss["A"], lss["A"].conj()])
    P = s["pinv"]
    n = s["n_var"]
    cond = lss["s"][0] / lss["s"][-1]
    assert np.max(np.abs(P @ A - np.eye(n))) < 1e-15 * cond * 50
    assert np.max(np.abs(P - lss["pinv"])) / np.max(np.abs(lss["pinv"])) < 1e-15 * cond * 50


def test_insufficient_points_raises():
    with pytest.raises(triqs_b200.TriqsRuntimeError, match="Insufficient data points"):
        triqs_b200.tail_fitter_setup(1.0, O.FERMION, 5, n_fixed=3)
    with pytest.raises(O.TriqsRuntimeError, match="Insufficient data points"):
        O.TailFitter(1.0, O.FERMION, 5).setup_lss(3, False)


def test_fixed_expansion_order():
    s = triqs_b200.tail_fitter_setup(10.0, O.FERMION, 200, n_fixed=1, expansion_order=5)
    f = O.TailFitter(10.0, O.FERMION, 200, expansion_order=5)
    lss = f.setup_lss(1, False)
    assert s["n_var"] == lss["n_var"] == 5
