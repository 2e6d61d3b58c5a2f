"""Device math (tile-FFT engine, register-resident inversion) compiled for the HOST and emulated
single-threaded: pins butterflies, twiddles, digit reversal and pivoting logic on the CPU-only box.
The emulation library is test infrastructure; the product never links it."""
import ctypes
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "hostsim", "libhostsim.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        pytest.skip("tests/hostsim/libhostsim.so not built (run __graft_entry__.build())")
    return ctypes.CDLL(LIB)


SIZES = [1, 2, 3, 4, 5, 6, 8, 10, 16, 20, 25, 32, 49, 64, 100, 125, 128, 250, 256, 400, 1000, 246, 6150, 9999, 10000, 1001, 404]


@pytest.mark.parametrize("N", SIZES)
def test_tile_fft_engine(lib, N):
    rng = np.random.default_rng(N)
    for TC in (1, 3):
        for sign in (1, -1):
            for direct in (0, 1):
                for skew in (0, 1):
                    x = rng.normal(size=(N, TC)) + 1j * rng.normal(size=(N, TC))
                    y = x.copy()
                    assert lib.hostsim_fft(N, TC, sign, direct, skew, y.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) >= 0
                    ref = np.fft.ifft(x, axis=0) * N if sign > 0 else np.fft.fft(x, axis=0)
                    assert np.max(np.abs(y - ref)) / np.max(np.abs(ref)) < 5e-15


@pytest.mark.parametrize("N,RA", [(250, 10), (400, 20), (256, 16), (100, 10), (500, 20), (625, 25), (128, 8)])
def test_two_pass_tile_fft(lib, N, RA):
    """tq_fft2.cuh: compile-time (RA x RB) passes incl. the radix-20 prime-factor and radix-25 butterflies."""
    rng = np.random.default_rng(N + RA)
    for tcw in (1, 5, 13):
        for sign in (1, -1):
            x = rng.normal(size=(N, tcw)) + 1j * rng.normal(size=(N, tcw))
            y = x.copy()
            assert lib.hostsim_fft2(N, RA, tcw, sign, y.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) == 0
            ref = np.fft.ifft(x, axis=0) * N if sign > 0 else np.fft.fft(x, axis=0)
            assert np.max(np.abs(y - ref)) / np.max(np.abs(ref)) < 5e-15


def test_three_pass_small_radix_tile_fft(lib):
    """tools/tq_fft3.cuh: N = R1 R2 R3 with radices <= 10 (an experiment: correct, but slower than the two-pass engine)."""
    rng = np.random.default_rng(9)
    for (N, R1, R2) in [(250, 10, 5), (250, 5, 10), (400, 10, 8), (400, 8, 10), (400, 5, 8)]:
        for tcw in (1, 5, 13):
            for sign in (1, -1):
                x = rng.normal(size=(N, tcw)) + 1j * rng.normal(size=(N, tcw))
                y = np.ascontiguousarray(x.copy())
                assert lib.hostsim_fft3(N, R1, R2, tcw, sign, y.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) == 0
                ref = np.fft.ifft(x, axis=0) * N if sign > 0 else np.fft.fft(x, axis=0)
                assert np.max(np.abs(y - ref)) / np.max(np.abs(ref)) < 5e-15


def test_radix_choice(lib):
    rad = (ctypes.c_int * 16)()
    n = lib.hostsim_radices(10000, rad)
    assert list(rad[:n]) == [10, 10, 10, 10]
    n = lib.hostsim_radices(6150, rad)
    assert sorted(rad[:n]) == [3, 5, 10, 41]
    assert lib.hostsim_radices(2 * 131, rad) == -1  # prime factor above the generic-radix limit


@pytest.mark.parametrize("n", range(1, 9))
def test_register_inverse(lib, n):
    rng = np.random.default_rng(n)
    a = rng.normal(size=(500, n, n)) + 1j * rng.normal(size=(500, n, n))
    # force pivoting: zero some diagonal entries
    if n > 1:
        a[::3, 0, 0] = 0
    if n > 2:
        a[::5, 1, 1] = 1e-18
    b = a.copy()
    assert lib.hostsim_invert(n, 500, b.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) == 0
    ref = np.linalg.inv(a)
    err = np.max(np.abs(b - ref), axis=(1, 2)) / np.max(np.abs(ref), axis=(1, 2))
    assert np.all(err < 1e-15 * 100 * np.linalg.cond(a))


@pytest.mark.parametrize("n", [2, 3, 5, 8])
def test_register_inverse_threshold_no_interchange(lib, n):
    """tq_invert_inplace_nopivot: accepted matrices are inverted to LAPACK accuracy, matrices whose natural pivot is too
    small are rejected (the kernels then fall back to the partially pivoted routine)."""
    rng = np.random.default_rng(100 + n)
    # Dyson-like matrices: (i w + mu) - H - Sigma with a definite anti-Hermitian part
    h = rng.normal(size=(400, n, n)) + 1j * rng.normal(size=(400, n, n))
    h = h + h.conj().transpose(0, 2, 1)
    w = rng.uniform(0.05, 30.0, 400)
    a = (1j * w + 0.1)[:, None, None] * np.eye(n)[None] - h
    a[::4, 0, 0] = 0.0       # natural first pivot unusable -> must be rejected
    b = a.copy()
    ok = np.zeros(400, dtype=np.int32)
    assert lib.hostsim_invert_nopivot(n, 400, b.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ok.ctypes.data_as(ctypes.POINTER(ctypes.c_int))) == 0
    assert not ok[::4].any()
    acc = ok.astype(bool)
    assert acc.sum() > 200
    ref = np.linalg.inv(a[acc])
    err = np.max(np.abs(b[acc] - ref), axis=(1, 2)) / np.max(np.abs(ref), axis=(1, 2))
    assert np.all(err < 1e-15 * 200 * np.linalg.cond(a[acc]))


REG_RADICES = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 16, 17, 20, 25, 32, 41]


@pytest.mark.parametrize("RA", REG_RADICES)
def test_planned_two_pass_engine_every_radix_pair(lib, RA):
    """tq_fft2.cuh fft2_item_a / fft2_item_b behind TQ_RADIX_SWITCH: every (RA, RB) pair of register radices, plus the direct-sum
    pass B for radices without a register butterfly (11, 13, 41, 101: the prime factors of the reference's own meshes)."""
    rng = np.random.default_rng(100 + RA)
    fn = lib.hostsim_fft2_rt
    for RB, generic in [(r, 0) for r in REG_RADICES] + [(11, 1), (13, 1), (41, 1), (101, 1), (22, 1), (8, 1), (9, 1)]:
        N = RA * RB
        for tcw, sign in ((1, 1), (7, -1)):
            x = rng.normal(size=(N, tcw)) + 1j * rng.normal(size=(N, tcw))
            y = np.ascontiguousarray(x.copy())
            assert fn(RA, RB, generic, tcw, sign, y.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) == 0
            ref = np.fft.ifft(x, axis=0) * N if sign > 0 else np.fft.fft(x, axis=0)
            assert np.max(np.abs(y - ref)) / np.max(np.abs(ref)) < 2e-14, (RA, RB, generic)
