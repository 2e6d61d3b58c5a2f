"""The planned four-step engine (fourier_pipe*.cu): any mesh length L = Na * Nb whose tile lengths split into a register radix and
a second radix (register or direct-sum), both directions, both statistics, ragged column tiles -- against the oracle at 1e-12 and
bit-for-bit consistent with nothing else; every case is ALSO run through the generic tile kernels (tq_set_option "no_pipe") so
that the two engines are checked against each other on identical inputs.

Lengths: the default adjoint mesh L = 6 n_iw (mesh/adjoint.hpp:35-38: 6000, 6144, 6150 = 2 3 5^2 41), the reference's own test mesh
L = 9999 = 3^2 11 101 (test/c++/gfs/fourier_matsubara.cpp:26-29), 10^4 with a matrix target, 12000, 61500, 2 10^5, powers of two."""
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


def pole_gw(beta, stat, n_iw, ncols, seed):
    rng = np.random.default_rng(seed)
    iw = O.imfreq_values(beta, stat, n_iw)
    E = rng.uniform(0.3, 2.0, (3, ncols)) * rng.choice([-1, 1], (3, ncols))
    r = rng.normal(size=(3, ncols)) + 1j * rng.normal(size=(3, ncols))
    return sum(r[p][None, :] / (iw[:, None] - E[p][None, :]) for p in range(3))


CASES = [
    # L, n_iw, ncols, stat, batch
    (600, 100, 5, F, 2), (1200, 200, 4, B, 1), (6000, 1000, 25, F, 2), (6144, 1024, 9, F, 1), (6150, 1025, 25, F, 2), (6150, 1025, 4, B, 1),
    (9999, 1000, 4, F, 2), (9999, 1000, 25, B, 1), (10000, 1025, 25, F, 2), (12000, 2000, 7, F, 1), (8192, 1024, 16, F, 1),
    (61500, 10250, 9, F, 1), (65536, 8192, 5, B, 1), (200000, 20000, 6, F, 1), (20000, 10000, 5, F, 1), (30030, 2000, 4, F, 1),
    (100000, 3000, 9, F, 1), (50000, 25000, 3, B, 2),
]


@pytest.mark.parametrize("L,n_iw,ncols,stat,batch", CASES)
def test_planned_engine_both_directions(ctx, L, n_iw, ncols, stat, batch):
    beta, n_tau = 20.0, L + 1
    gws = np.stack([pole_gw(beta, stat, n_iw, ncols, 1000 * L % 7919 + b) for b in range(batch)])
    refs, tails = [], []
    for b in range(batch):
        r, info = O.fourier_iw_to_tau(gws[b], beta, stat, n_tau, return_info=True)
        refs.append(r)
        tails.append(info["tail"])
    tails = np.stack(tails)
    gt = ctx.fourier_iw_to_tau(gws, beta, stat, n_tau, moments=tails)
    for b in range(batch):
        assert rel(gt[b], refs[b]) < TOL
    gt_own = ctx.fourier_iw_to_tau(gws, beta, stat, n_tau)      # the library's own least-squares tail
    for b in range(batch):
        assert rel(gt_own[b], refs[b]) < TOL
    gw2 = ctx.fourier_tau_to_iw(np.stack(refs), beta, stat, n_iw)
    for b in range(batch):
        assert rel(gw2[b], O.fourier_tau_to_iw(refs[b], beta, stat, n_iw)) < TOL
    # the generic engine on the same inputs
    ctx.set_option("no_pipe", 1)
    try:
        gt_g = ctx.fourier_iw_to_tau(gws, beta, stat, n_tau, moments=tails)
        gw_g = ctx.fourier_tau_to_iw(np.stack(refs), beta, stat, n_iw)
    finally:
        ctx.set_option("no_pipe", 0)
    assert rel(gt_g, gt) < TOL and rel(gw_g, gw2) < TOL


def test_planned_engine_launches_the_pipelined_kernels(ctx):
    """the fast path really is the one that runs off the benchmark mesh: kernel tags of the profile report"""
    beta, L, n_iw, ncols = 10.0, 6150, 1025, 25
    gw = pole_gw(beta, F, n_iw, ncols, 3)[None]
    ctx.profile_enable(True)
    gt = ctx.fourier_iw_to_tau(gw, beta, F, L + 1)
    ctx.fourier_tau_to_iw(gt, beta, F, n_iw)
    rep = ctx.profile_report()
    ctx.profile_enable(False)
    for tag in ("iw2tau_pass1", "iw2tau_pass2", "tau2iw_pass1", "tau2iw_pass2"):
        assert tag in rep, rep


@pytest.mark.parametrize("dims,n_others", [((32, 32, 32), 5), ((96, 96, 1), 9), ((512, 512, 1), 3), ((100, 100, 1), 12), ((48, 48, 48), 2), ((7, 11, 13), 6),
                                            ((41, 1, 1), 40), ((1000, 1, 1), 9), ((1, 303, 1), 17), ((12, 1, 320), 8), ((640, 2, 1), 4)])
def test_planned_lattice_axes(ctx, dims, n_others):
    """cyclat <-> brzone on axis lengths outside the compile-time set {64, 128, 256}: planned pipelined kernel vs oracle and vs
    the generic tile kernel (fourier_lattice.cpp:30-59; FFTW takes any length, fourier_common.cpp:30-44)."""
    rng = np.random.default_rng(sum(dims) + n_others)
    nk = int(np.prod(dims))
    x = rng.normal(size=(nk, n_others)) + 1j * rng.normal(size=(nk, n_others))
    yk = ctx.fourier_lattice(x, dims, True)
    assert rel(yk, O.fourier_lattice_r_to_k(x, dims)) < TOL
    yr = ctx.fourier_lattice(yk, dims, False)
    assert rel(yr, x) < TOL
    ctx.set_option("no_pipe", 1)
    try:
        yk_g = ctx.fourier_lattice(x, dims, True)
    finally:
        ctx.set_option("no_pipe", 0)
    assert rel(yk_g, yk) < TOL


NARROW = [
    # L, n_iw, ncols, stat, batch        ("gfs as columns": batch * ncols >= 8, ncols < 4)
    (6150, 1025, 1, F, 8), (6150, 1025, 1, B, 37), (6000, 1000, 2, F, 13), (10000, 1025, 1, F, 9), (10000, 1025, 3, F, 11), (9999, 1000, 1, F, 24),
    (600, 100, 3, B, 40), (20000, 2000, 1, F, 10), (100000, 10000, 1, F, 8), (100000, 10000, 2, B, 5), (1200, 200, 1, F, 300), (6144, 1024, 3, F, 3),
]


def pole_gw_moments(beta, stat, n_iw, ncols, seed):
    """sum_p r_p / (i w - E_p) and its exact moments [0, sum r, sum r E, sum r E^2]"""
    rng = np.random.default_rng(seed)
    iw = O.imfreq_values(beta, stat, n_iw)
    E = rng.uniform(0.3, 2.0, (3, ncols)) * rng.choice([-1, 1], (3, ncols))
    r = rng.normal(size=(3, ncols)) + 1j * rng.normal(size=(3, ncols))
    gw = sum(r[p][None, :] / (iw[:, None] - E[p][None, :]) for p in range(3))
    return gw, np.stack([np.zeros(ncols, complex), r.sum(0), (r * E).sum(0), (r * E * E).sum(0)])


@pytest.mark.parametrize("L,n_iw,ncols,stat,batch", NARROW)
def test_narrow_targets_run_as_columns_of_one_wide_gf(ctx, L, n_iw, ncols, stat, batch):
    """Batches of scalar / 1-3 column gfs (configs[0] is one): the planned engine treats the batch as ONE gf whose columns are all
    (gf, column) pairs, gathered by tensor maps; checked against the oracle on the first, middle and last gf and against the per-gf
    kernels ("no_tr") on all of them."""
    beta, n_tau = 15.0, L + 1
    pick = sorted({0, batch // 2, batch - 1})
    both = [pole_gw_moments(beta, stat, n_iw, ncols, 77 * L % 7919 + b) for b in range(batch)]
    gws, tails = np.stack([x[0] for x in both]), np.stack([x[1] for x in both])
    ctx.profile_enable(True)
    gt = np.asarray(ctx.fourier_iw_to_tau(gws, beta, stat, n_tau, moments=tails))
    rep = ctx.profile_report()
    ctx.profile_enable(False)
    # (single-column gfs on smooth short meshes keep the single-kernel path for the inverse leg: half-sector writes)
    # (... and lengths whose plan needs a large direct-sum radix, 101 | 9999, run as a Bluestein convolution with all gfs side by side)
    assert ("iw2tau_pass1" in rep and "iw2tau_pass2" in rep) or (ncols == 1 and "iw2tau_single" in rep) or "iw2tau_bluestein_pre" in rep, rep
    for b in pick:
        assert rel(gt[b], O.fourier_iw_to_tau(gws[b], beta, stat, n_tau, known_moments=tails[b])) < TOL
    gw2 = np.asarray(ctx.fourier_tau_to_iw(gt, beta, stat, n_iw, moments=tails))
    for b in pick:
        assert rel(gw2[b], O.fourier_tau_to_iw(gt[b], beta, stat, n_iw, known_moments=tails[b])) < TOL
    ctx.set_option("no_tr", 1)
    try:
        gt_g = np.asarray(ctx.fourier_iw_to_tau(gws, beta, stat, n_tau, moments=tails))
        gw_g = np.asarray(ctx.fourier_tau_to_iw(gt, beta, stat, n_iw, moments=tails))
    finally:
        ctx.set_option("no_tr", 0)
    for b in range(batch):
        assert rel(gt_g[b], gt[b]) < TOL and rel(gw_g[b], gw2[b]) < TOL


@pytest.mark.parametrize("n_tau,n_iw,ncols,batch", [(100001, 10000, 25, 2), (6151, 1025, 9, 2), (10000, 1000, 25, 2), (5000, 833, 5, 3), (10001, 1025, 1, 16),
                                                    (10000, 1000, 2, 8), (8192, 1000, 1, 16), (12001, 2000, 4, 1)])
def test_plan_describe_names_the_engine_that_runs(ctx, n_tau, n_iw, ncols, batch):
    """tq_plan_describe (host only) against the kernels that actually run (profile tags of the forward call)"""
    from triqs_b200 import plan_describe
    d = plan_describe(n_tau, n_iw, F, ncols, batch)
    rng = np.random.default_rng(3)
    gt = rng.normal(size=(batch, n_tau, ncols)) * 1e-3 + 0j
    mom = np.zeros((batch, 4, ncols), complex)
    mom[:, 1] = 1.0
    ctx.profile_enable(True)
    ctx.fourier_tau_to_iw(gt, 10.0, F, n_iw, moments=mom)
    rep = ctx.profile_report()
    ctx.profile_enable(False)
    eng = d["engine_fwd"]
    if eng == "bluestein":
        assert "tau2iw_bluestein_pre" in rep, (d, sorted(rep))
    elif eng == "planned":
        assert "tau2iw_pass1" in rep and "tau2iw_pass2" in rep, (d, sorted(rep))
    elif eng == "single-column":
        assert "tau2iw_col" in rep, (d, sorted(rep))
    else:
        assert "tau2iw_bluestein_pre" not in rep, (d, sorted(rep))
