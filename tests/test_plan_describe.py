"""Host-side planning of the imtime <-> imfreq transforms (tq_plan_describe: no device needed): which engine a mesh takes.  FFTW plans any
length per call (c++/triqs/gfs/transform/fourier_common.cpp:30-44); here the choice is a cached plan, and these are its rules."""
import re

import pytest

import triqs_b200
from triqs_b200 import plan_describe

F, B = triqs_b200.FERMION, triqs_b200.BOSON


def smooth(n):
    for p in (2, 3, 5):
        while n % p == 0:
            n //= p
    return n == 1


def tiles(text):
    m = re.match(r"Nb=(\d+)\((\d+)x(\d+)\) Na=(\d+)\((\d+)x(\d+)\) cost=([\d.]+) heavy=(\d)(\d) generic=(\d)(\d)", text)
    assert m, text
    return [int(x) if i != 6 else float(x) for i, x in enumerate(m.groups())]


def test_headline_mesh_keeps_its_compile_time_plan():
    d = plan_describe(100001, 10000, F, 25, 2)
    assert d["L"] == 100000 and d["n_w"] == 20000 and d["engine_fwd"] == d["engine_inv"] == "planned"
    assert d["plan_fwd"] == d["plan_inv"] and tiles(d["plan_fwd"])[:6] == [250, 10, 25, 400, 20, 20]


@pytest.mark.parametrize("L", [6000, 6144, 10000, 12000, 65536, 200000, 20000, 30000])
def test_smooth_lengths_have_a_two_radix_plan_in_registers(L):
    d = plan_describe(L + 1, L // 6, F, 25, 1)
    nb, ra1, rb1, na, ra2, rb2, cost, h1, h2, g1, g2 = tiles(d["plan_fwd"])
    assert nb * na == L and ra1 * rb1 == nb and ra2 * rb2 == na and max(nb, na) <= 640
    assert (h1, h2, g1, g2) == (0, 0, 0, 0) and d["engine_fwd"] == "planned" and d["plan_fwd"] == d["plan_inv"]


@pytest.mark.parametrize("L,prime", [(6150, 41), (61500, 41), (6600, 11), (7800, 13), (10200, 17)])
def test_rader_tile_sits_in_pass_1_forward_and_in_pass_2_on_the_way_back(L, prime):
    """a heavy (Rader) tile costs 1.2 ms in F1 / 1.15 ms in I2 against 2.0 ms in F2 / 2.8 ms in I1: transposed plans"""
    d = plan_describe(L + 1, L // 6, F, 25, 4)
    f, i = tiles(d["plan_fwd"]), tiles(d["plan_inv"])
    assert (f[7], f[8]) == (1, 0) and (i[7], i[8]) == (0, 1)
    assert f[0] % prime == 0 and i[3] % prime == 0 and f[0] * f[3] == L and i[0] * i[3] == L
    assert d["engine_fwd"] == d["engine_inv"] == "planned"


@pytest.mark.parametrize("n_tau,n_iw,stat", [(2000, 300, F), (5000, 833, B), (8192, 1024, F), (20000, 2000, F), (2 * 10007 + 1, 1500, B)])
def test_lengths_without_a_plan_take_the_bluestein_convolution(n_tau, n_iw, stat):
    d = plan_describe(n_tau, n_iw, stat, 5, 3)
    assert d["plan_fwd"] == "none" and d["plan_cost"] < 0 and d["engine_fwd"] == d["engine_inv"] == "bluestein"
    m1, m2 = d["bluestein_M"]
    assert smooth(m1) and smooth(m2) and 16 <= m1 <= m2 <= 640
    assert m1 * m2 >= d["L"] + d["n_w"] - 1            # linear convolution of L inputs onto n_w outputs without wrap-around
    assert m1 * m2 <= 1.12 * (d["L"] + d["n_w"] - 1)   # ... and not much longer than that


def test_costly_direct_sum_plans_are_replaced_by_bluestein():
    """101 | 9999 (the reference's own test mesh, test/c++/gfs/fourier_matsubara.cpp:26-29): a 101-point direct-sum pass costs 404 FP64
    instructions per element; beyond a plan cost of 9.5 the convolution is faster"""
    d = plan_describe(10000, 1000, F, 25, 2)
    assert d["plan_cost"] > 9.5 and tiles(d["plan_fwd"])[9:] == [0, 1] and d["engine_fwd"] == "bluestein"
    assert d["bluestein_M"] == [100, 120]
    assert plan_describe(10000, 1000, F, 2, 8)["engine_fwd"] == "bluestein"     # batches of narrow targets: wide buffer
    assert plan_describe(10000, 1000, F, 2, 1)["engine_fwd"] != "bluestein"     # one narrow gf: planned / single-kernel
    assert plan_describe(1000, 100, F, 25, 2)["engine_fwd"] == "planned"        # 999 = 27 x 37: a small direct-sum radix keeps its plan


def test_short_meshes_and_invalid_arguments():
    d = plan_describe(201, 30, F, 4, 1)
    assert d["engine_fwd"] == "generic" and d["bluestein_M"] is None
    assert plan_describe(10001, 1025, F, 1, 8192)["engine_fwd"] == "single-column"
    for bad in [(1, 10, F, 1, 1), (100, 0, F, 1, 1), (100, 10, 7, 1, 1), (100, 10, F, 0, 1)]:
        with pytest.raises(ValueError):
            plan_describe(*bad)
