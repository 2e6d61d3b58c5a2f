"""The Python mirror of the reference API (triqs_b200.gf) on the GPU: same call shapes as python/triqs/gf."""
import numpy as np
import pytest

from oracle import triqs_oracle as O

pytestmark = pytest.mark.gpu


def _gw(beta, n_iw, n, seed):
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
    H = 0.5 * (a + a.conj().T)
    iw = O.imfreq_values(beta, O.FERMION, n_iw)
    return np.linalg.inv(iw[:, None, None] * np.eye(n)[None] - H[None]), H


def test_gf_round_trip_blockgf_and_tail():
    from triqs_b200 import gf as G
    beta, n_iw = 10.0, 300
    blocks = []
    for b in range(2):
        d, H = _gw(beta, n_iw, 2, b)
        blocks.append(G.Gf(G.MeshImFreq(beta, O.FERMION, n_iw), (2, 2), d))
    bg = G.BlockGf(["up", "down"], blocks)
    bgt = G.make_gf_from_fourier(bg)              # default n_tau = 6 n_iw + 1
    assert bgt["up"].mesh.n_tau == 6 * n_iw + 1
    bgw = G.make_gf_from_fourier(bgt, n_iw)
    for b in range(2):
        ref_t = O.fourier_iw_to_tau(blocks[b].flat(), beta, O.FERMION, 6 * n_iw + 1)
        assert np.max(np.abs(bgt[b].flat() - ref_t)) / np.max(np.abs(ref_t)) < 1e-12
        assert np.max(np.abs(bgw[b].data - blocks[b].data)) < 1e-8
    tail, err = G.fit_hermitian_tail(blocks[0])
    assert tail.shape[1:] == (2, 2) and err < 1e-8
    assert np.max(np.abs(tail[1] - np.eye(2))) < 1e-6
    tails, max_err = G.fit_hermitian_tail(bg)
    assert len(tails) == 2 and max_err < 1e-8
    # single gf, known moments
    km = G.make_zero_tail(blocks[0], 2)
    km[1] = np.eye(2)
    gt = G.make_gf_from_fourier(blocks[0], 2001, km)
    ref = O.fourier_iw_to_tau(blocks[0].flat(), beta, O.FERMION, 2001, km.reshape(2, 4))
    assert np.max(np.abs(gt.flat() - ref)) / np.max(np.abs(ref)) < 1e-12


def test_gf_call_inverse_and_lattice():
    from triqs_b200 import gf as G
    d, H = _gw(10.0, 50, 3, 3)
    g = G.Gf(G.MeshImFreq(10.0, O.FERMION, 50), (3, 3), d)
    gi = G.inverse(g)
    iw = g.mesh.values()
    assert np.max(np.abs(gi.data - (iw[:, None, None] * np.eye(3)[None] - H[None]))) < 1e-12
    gt = G.make_gf_from_fourier(g, 601)
    v = gt(np.array([0.0, 1.234, 10.0]))
    ref = O.interp_tau(gt.flat(), 10.0, np.array([0.0, 1.234, 10.0])).reshape(3, 3, 3)
    assert np.array_equal(v, ref)
    assert np.array_equal(gt(1.234), ref[1])
    rng = np.random.default_rng(0)
    gr = G.Gf(G.MeshCycLat((6, 5, 4)), (2,), rng.normal(size=(120, 2)) + 1j * rng.normal(size=(120, 2)))
    gk = G.make_gf_from_fourier(gr)
    assert isinstance(gk.mesh, G.MeshBrZone)
    assert np.max(np.abs(gk.flat() - O.fourier_lattice_r_to_k(gr.flat(), (6, 5, 4)))) < 1e-12
    back = G.make_gf_from_fourier(gk)
    assert np.max(np.abs(back.data - gr.data)) < 1e-13


def test_dyson_k_sum_sharded_on_one_gpu():
    """ranks emulated sequentially on one GPU: partial sums with the 1/N_k(total) weight add up to the whole."""
    import triqs_b200
    from triqs_b200 import gf as G
    ctx = triqs_b200.Context(0, print_warnings=False)
    rng = np.random.default_rng(4)
    eps = rng.normal(size=(101, 5, 5)) + 1j * rng.normal(size=(101, 5, 5))
    eps = eps + eps.conj().transpose(0, 2, 1)
    iw = O.imfreq_values(10.0, O.FERMION, 40)
    sig = np.stack([0.2 * np.eye(5) / (w + 0.3) for w in iw])
    sigma = G.Gf(G.MeshImFreq(10.0, O.FERMION, 40), (5, 5), sig)
    ref = O.dyson_ksum(eps, sig, 10.0, O.FERMION, 0.2)
    parts = []
    for r in range(4):
        parts.append(G.dyson_k_sum(ctx, eps, sigma, 0.2, rank=r, n_ranks=4, reduce_fn=lambda p: p).data)
    assert np.max(np.abs(sum(parts) - ref)) / np.max(np.abs(ref)) < 1e-12
    ctx.close()


def test_density_of_local_gf():
    """density(G_loc) after a Dyson k-sum, the DMFT-loop use of the two together (lattice_utils.py:155-166 pattern)."""
    import triqs_b200
    from triqs_b200 import gf as G
    ctx = triqs_b200.Context(0, print_warnings=False)
    rng = np.random.default_rng(8)
    n, beta, n_iw = 3, 10.0, 300
    eps = rng.normal(size=(64, n, n)) + 1j * rng.normal(size=(64, n, n))
    eps = 0.3 * (eps + eps.conj().transpose(0, 2, 1))
    iw = O.imfreq_values(beta, O.FERMION, n_iw)
    sig = np.stack([0.1 * np.eye(n) / (w - 0.4) for w in iw])
    sigma = G.Gf(G.MeshImFreq(beta, O.FERMION, n_iw), (n, n), sig, ctx=ctx)
    gloc = G.dyson_k_sum(ctx, eps, sigma, 0.1, weights=np.full(64, 1 / 64), reduce_fn=lambda p: p)
    dens = G.density(gloc)
    ref = O.density(O.dyson_ksum(eps, sig, beta, O.FERMION, 0.1, weights=np.full(64, 1 / 64)), beta, O.FERMION)
    assert np.max(np.abs(dens - ref)) < 1e-9
    assert 0.0 < dens.trace().real < n
    ctx.close()


def test_hermiticity_and_replace_by_tail_mirror():
    """hermiticity.cpp:25-80 and fit_and_replace.cpp:42-72 with the Gf mirror (fit on the full mesh instead of a window)."""
    import triqs_b200
    from triqs_b200 import gf as G
    ctx = triqs_b200.Context(0, print_warnings=False)
    beta = 1.0
    mesh = G.MeshImFreq(beta, O.FERMION, 1025)
    iw = O.imfreq_values(beta, O.FERMION, 1025)
    h = np.array([[1, 1j], [-1j, 2]], complex)
    g = G.Gf(mesh, (2, 2), np.linalg.inv(iw[:, None, None] * np.eye(2) - h), ctx)
    assert G.is_gf_hermitian(g)
    assert np.max(np.abs(G.make_hermitian(g).data - g.data)) < 1e-15
    gt = G.make_gf_from_fourier(g)
    assert G.is_gf_hermitian(gt, 1e-10)
    hb = np.array([[1, 1j], [1j, 2]], complex)
    gb = G.Gf(mesh, (2, 2), np.linalg.inv(iw[:, None, None] * np.eye(2) - hb), ctx)
    assert not G.is_gf_hermitian(gb)
    assert G.is_gf_hermitian(G.make_hermitian(gb))
    # replace the outer part of a gf by its fitted tail: close to the exact function
    beta2, n_iw = 10.0, 400
    mesh2 = G.MeshImFreq(beta2, O.FERMION, n_iw)
    iw2 = O.imfreq_values(beta2, O.FERMION, n_iw)
    exact = (1 / (iw2 - 1.0 - (1 / (iw2 - 2) + 3 / (iw2 + 3)))).reshape(-1, 1, 1)
    g2 = G.Gf(mesh2, (1, 1), exact.copy(), ctx)
    tail, err = G.fit_tail(g2)
    n = -n_iw + np.arange(2 * n_iw)
    g2.data[np.abs(n + 0.5) > 330] = 123.0  # garbage in the outer part
    G.replace_by_tail(g2, tail, 330)
    assert np.max(np.abs(g2.data - exact)) < 1e-8
    ctx.close()


def test_real_time_transform_reads_like_the_reference_test():
    """fourier_real.cpp:27-87 through the mirror: Gt = make_gf_from_fourier(Gw, t_mesh, tail), matrix-valued target."""
    from triqs_b200 import gf as G
    w_mesh = G.MeshReFreq(-40.0, 40.0, 1001)
    t_mesh = G.make_adjoint_mesh(w_mesh)
    E = 1.0
    w, t = w_mesh.values(), t_mesh.values()
    Gw1 = G.Gf(w_mesh, (2, 2), np.repeat((2 * E / (w * w + E * E)).astype(complex)[:, None], 4, axis=1).reshape(-1, 2, 2))
    tail = np.zeros((3, 2, 2), complex)
    tail[2] = 2 * E
    Gt1 = G.make_gf_from_fourier(Gw1, t_mesh, tail)
    assert isinstance(Gt1.mesh, G.MeshReTime) and Gt1.mesh == t_mesh
    assert np.max(np.abs(Gt1.data - np.exp(-E * np.abs(t))[:, None, None])) < 1e-4
    Gw1b = G.make_gf_from_fourier(Gt1, w_mesh, tail)
    assert np.max(np.abs(Gw1b.data - Gw1.data)) < 1e-4
    # exactly as the reference test: the tail comes from fit_tail(Gw1) on the refreq mesh (fourier_real.cpp:52-53)
    ftail, err = G.fit_tail(Gw1)
    assert ftail.shape[1:] == (2, 2) and abs(ftail[2, 0, 0] - 2 * E) < 1e-3
    Gt1f = G.make_gf_from_fourier(Gw1, t_mesh, ftail)
    assert np.max(np.abs(Gt1f.data - np.exp(-E * np.abs(t))[:, None, None])) < 1e-4
    assert np.max(np.abs(G.make_gf_from_fourier(Gt1f, w_mesh, ftail).data - Gw1.data)) < 1e-4
    theta = np.where(t > 0, 1.0, np.where(t < 0, 0.0, 0.5))
    Gt2 = G.Gf(t_mesh, (2, 2))
    Gt2.data[:] = (0.5j * (np.exp(-E * np.abs(t)) * (1 - theta) - np.exp(-E * np.abs(t)) * theta))[:, None, None]
    km = np.zeros((3, 2, 2), complex)
    km[1] = 1.0
    Gw2 = G.make_gf_from_fourier(Gt2, w_mesh, km)
    assert np.max(np.abs(Gw2.data - (0.5 / (w + 1j * E) + 0.5 / (w - 1j * E))[:, None, None])) < 1e-4
    assert np.max(np.abs(G.make_gf_from_fourier(Gw2, t_mesh, km).data - Gt2.data)) < 1e-4


def test_legendre_descriptors():
    """issue459.py / tail_issues.py:176-183: gl << MatsubaraToLegendre(gw); gw << LegendreToMatsubara(gl)."""
    from triqs_b200 import gf as G
    beta, n_iw = 10.0, 300
    d, H = _gw(beta, n_iw, 2, 3)
    gw = G.Gf(G.MeshImFreq(beta, O.FERMION, n_iw), (2, 2), d)
    gl = G.matsubara_to_legendre(gw, 40)
    assert isinstance(gl.mesh, G.MeshLegendre) and gl.data.shape == (40, 2, 2)
    gw2 = G.legendre_to_matsubara(gl, gw.mesh)
    assert np.max(np.abs(gw2.data - gw.data)) < 1e-4     # 40 coefficients resolve this G to ~1e-5
    gt = G.legendre_to_matsubara(gl, G.MeshImTime(beta, O.FERMION, 1001))
    ref = O.fourier_iw_to_tau(gw.flat(), beta, O.FERMION, 1001).reshape(1001, 2, 2)
    assert np.max(np.abs(gt.data - ref)) < 1e-4


def test_multivar_fourier_reads_like_the_reference_test():
    """test/c++/gfs/multivar/fourier.cpp:24-93: g(k, iW, iw) = 1 / (iw + iW - cos kx cos ky) on brzone x imfreq(Boson) x imfreq;
    transform mesh 0, mesh 2, and both, and come back (precision 1e-4); each leg also against the oracle."""
    from triqs_b200 import gf as G
    beta, N_iw, N_iW, N_k = 1.0, 100, 4, 4
    k_mesh, r_mesh = G.MeshBrZone((N_k, N_k, 1)), G.MeshCycLat((N_k, N_k, 1))
    iw_mesh, iW_mesh = G.MeshImFreq(beta, O.FERMION, N_iw), G.MeshImFreq(beta, O.BOSON, N_iW)
    kk = 2 * np.pi * np.arange(N_k) / N_k
    eps = (np.cos(kk)[:, None] * np.cos(kk)[None, :]).reshape(-1)
    val = 1 / (iw_mesh.values()[None, None, :] + iW_mesh.values()[None, :, None] - eps[:, None, None])
    for target in ((), (2, 2)):
        data = val.reshape(val.shape + (1,) * len(target)) * np.ones(target)
        g = G.Gf(G.MeshProduct((k_mesh, iW_mesh, iw_mesh)), target, np.ascontiguousarray(data))
        # mesh 0 and back
        g_r = G.make_gf_from_fourier_axes(g, (0,), (r_mesh,))
        assert g_r.mesh[0] == r_mesh
        ref = O.fourier_lattice_k_to_r(g.data.reshape(N_k * N_k, -1), (N_k, N_k, 1)).reshape(g.data.shape)
        assert np.max(np.abs(g_r.data - ref)) < 1e-13
        assert np.max(np.abs(G.make_gf_from_fourier_axes(g_r, (0,), (k_mesh,)).data - g.data)) < 1e-12
        # mesh 2 and back
        tau_mesh = G.MeshImTime(beta, O.FERMION, 2 * N_iw + 1)
        g_tau = G.make_gf_from_fourier_axes(g, (2,), (tau_mesh,))
        assert g_tau.data.shape[:3] == (N_k * N_k, 2 * N_iW - 1, 2 * N_iw + 1)
        d3 = g.data.reshape(N_k * N_k * (2 * N_iW - 1), 2 * N_iw, -1)
        for b in (0, 17, d3.shape[0] - 1):
            ref = O.fourier_iw_to_tau(d3[b], beta, O.FERMION, 2 * N_iw + 1)
            got = g_tau.data.reshape(d3.shape[0], 2 * N_iw + 1, -1)[b]
            assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-10
        gb = G.make_gf_from_fourier_axes(g_tau, (2,), (iw_mesh,))
        assert np.max(np.abs(gb.data - g.data)) < 1e-4
        # meshes 0 and 2 together, default adjoint meshes, and back with the meshes given
        g_r_tau = G.make_gf_from_fourier_axes(g, (0, 2))
        assert isinstance(g_r_tau.mesh[0], G.MeshCycLat) and isinstance(g_r_tau.mesh[2], G.MeshImTime)
        gb1 = G.make_gf_from_fourier_axes(g_r_tau, (0, 2), (k_mesh, iw_mesh))
        assert np.max(np.abs(gb1.data - g.data)) < 1e-4


def test_imfreq_evaluator_uses_the_tail_outside_the_mesh():
    """evaluator.hpp:50-76: g(n) = data inside the window, the fitted tail expansion outside."""
    from triqs_b200 import gf as G
    beta, n_iw = 10.0, 400
    d, H = _gw(beta, n_iw, 2, 11)
    g = G.Gf(G.MeshImFreq(beta, O.FERMION, n_iw), (2, 2), d)
    assert np.array_equal(g(5), d[5 + n_iw])
    for n in (n_iw + 10, -3 * n_iw, 5000):
        iw = 1j * np.pi * (2 * n + 1) / beta
        exact = np.linalg.inv(iw * np.eye(2) - H)
        tail, _ = O.fit_tail(g.flat(), beta, O.FERMION)
        ref = O.tail_eval(tail, iw).reshape(2, 2)
        assert np.max(np.abs(g(n) - ref)) < 1e-10
        assert np.max(np.abs(g(n) - exact)) < 1e-6
    # batched: many indices in one call (tq_evaluate_imfreq), inside and outside the window
    ns = np.array([-3 * n_iw, -n_iw - 1, -n_iw, -1, 0, 5, n_iw - 1, n_iw, n_iw + 10, 5000])
    vals = g(ns)
    tail, _ = O.fit_tail(g.flat(), beta, O.FERMION)
    for i, n in enumerate(ns):
        if -n_iw <= n <= n_iw - 1:
            assert np.array_equal(vals[i], d[n + n_iw])
        else:
            ref = O.tail_eval(tail, 1j * np.pi * (2 * n + 1) / beta).reshape(2, 2)
            assert np.max(np.abs(vals[i] - ref)) < 1e-10


def test_torch_producer_and_consumer_need_no_manual_sync():
    """the context's stream is non-blocking; the binding orders it against torch's current stream on both sides
    (tq_stream_wait_external / tq_stream_signal_external), so torch kernel -> library -> torch read chains are race free."""
    import torch
    import triqs_b200
    ctx = triqs_b200.Context(0, print_warnings=False)
    dev = torch.device("cuda", 0)
    n, npts = 4, 200_000
    g = torch.Generator(device=dev).manual_seed(3)
    base = torch.randn(npts, n, n, dtype=torch.complex128, device=dev, generator=g)
    eye = torch.eye(n, dtype=torch.complex128, device=dev)
    for it in range(5):
        big = torch.randn(4096, 4096, device=dev)
        junk = big @ big                      # keeps torch's stream busy so that an unordered library call would start too early
        a = base * (1.0 + it) + (3.0 + it) * eye * (junk[0, 0] * 0 + 1)   # produced on torch's stream, after the matmul
        ref = torch.linalg.inv(a.clone())
        ctx.invert_in_place(a)                # library stream
        err = (a - ref).abs().max() / ref.abs().max()   # consumed on torch's stream
        assert float(err) < 1e-12
    # device outputs of a transform feed a torch op directly
    beta, n_iw, n_tau = 10.0, 100, 601
    iw = torch.from_numpy(O.imfreq_values(beta, O.FERMION, n_iw)).to(dev)
    gw = (1.0 / (iw - 0.7))[None, :, None].contiguous()
    gt = ctx.fourier_iw_to_tau(gw, beta, O.FERMION, n_tau)
    s = gt.sum().cpu()
    ref = O.fourier_iw_to_tau(gw[0].cpu().numpy(), beta, O.FERMION, n_tau)
    assert abs(complex(s) - ref.sum()) < 1e-9 * abs(ref.sum())
    ctx.close()


def test_gf_invert_strided_view_and_device_replace_by_tail():
    """Gf.invert on a non-contiguous view must write the result back; replace_by_tail must work on torch data."""
    import torch
    from triqs_b200 import gf as G
    d, H = _gw(10.0, 40, 2, 11)
    wide = np.zeros((80, 2, 4), complex)
    wide[:, :, :2] = d
    g = G.Gf(G.MeshImFreq(10.0, O.FERMION, 40), (2, 2), wide[:, :, :2])   # strided view
    assert not g.data.flags.c_contiguous
    g.invert()
    iw = g.mesh.values()
    assert np.max(np.abs(wide[:, :, :2] - (iw[:, None, None] * np.eye(2)[None] - H[None]))) < 1e-12
    # device-resident gf
    gd = G.Gf(G.MeshImFreq(10.0, O.FERMION, 40), (2, 2), torch.from_numpy(d.copy()).cuda())
    tail = np.zeros((4, 2, 2), complex)
    tail[1] = np.eye(2)
    tail[2] = H
    tail[3] = H @ H
    G.replace_by_tail(gd, tail, 30)
    ref = O.replace_by_tail(d.reshape(80, 4), 10.0, O.FERMION, tail.reshape(4, 4), 30).reshape(80, 2, 2)
    assert np.max(np.abs(gd.data.cpu().numpy() - ref)) < 1e-14
    gd.invert()
    assert np.max(np.abs(gd.data.cpu().numpy() - np.linalg.inv(ref))) / np.max(np.abs(np.linalg.inv(ref))) < 1e-12


def test_product_mesh_evaluation_reads_like_the_reference_test():
    """test/c++/gfs/multivar/g_k_om.cpp:76-114 (Gkom.Eval): g(k, w) on gf<prod<brzone, imfreq>, scalar_valued>"""
    from triqs_b200 import gf as G
    beta, n_k, n_w = 5.0, 40, 5
    bz, wm = G.MeshBrZone((n_k, n_k)), G.MeshImFreq(beta, O.FERMION, n_w)
    iw = wm.values()
    k = np.arange(n_k) / n_k * 2 * np.pi
    data = 1 / (iw[None, None, :] + 3 + 2 * (np.cos(k)[:, None, None] + np.cos(k)[None, :, None]))
    g = G.Gf(G.MeshProduct((bz, wm)), (), data.reshape(n_k * n_k, len(iw)))
    # re-evaluate on the mesh itself
    ks = np.array([[k[ix], k[iy], 0.0] for ix in range(n_k) for iy in range(n_k)])
    for n in range(wm.first_index, wm.last_index + 1):
        r = g(ks, n)
        assert np.max(np.abs(r - data[:, :, n - wm.first_index].reshape(-1))) < 1e-14
    # a three times finer grid: within 5 % (the reference's bound)
    n2 = 3 * n_k
    kk = np.arange(n2) / n2 * 2 * np.pi
    ks = np.array([[kx, ky, 0.0] for kx in kk for ky in kk])
    for n in (wm.first_index, 0, wm.last_index):
        w = iw[n - wm.first_index]
        res = 1 / (w + 3 + 2 * (np.cos(ks[:, 0]) + np.cos(ks[:, 1])))
        assert np.all(np.abs(g(ks, n) - res) <= 0.05 * np.abs(res))
    assert abs(g(np.array([np.pi, np.pi, 0.0]), 0) - 1 / (iw[n_w] + 3 - 4)) < 1e-14  # a single point (pi, pi) is a mesh point
    # g(tau, tau') on prod<imtime, imtime>: bilinear data are reproduced
    tm = G.MeshImTime(2.0, O.FERMION, 21)
    t = tm.values()
    f = (1 + t[:, None] - 2 * t[None, :] + 0.25 * t[:, None] * t[None, :]).astype(complex)
    g2 = G.Gf(G.MeshProduct((tm, tm)), (), f)
    rng = np.random.default_rng(3)
    x, y = rng.uniform(0, 2, 100), rng.uniform(0, 2, 100)
    assert np.max(np.abs(g2(x, y) - (1 + x - 2 * y + 0.25 * x * y))) < 1e-13


@pytest.mark.gpu
def test_lazy_descriptors_fourier_legendre_lshift():
    """`gw << Fourier(gt)` / `gt << Fourier(gw)` / BlockGf `<<` / Legendre descriptors (python/triqs/gf/descriptors.py:156-196,
    gf.py:486-519): the same numbers as the factory functions, written INTO the existing gf."""
    import numpy as np
    from oracle import triqs_oracle as O
    import triqs_b200
    from triqs_b200.gf import (BlockGf, Fourier, Gf, LegendreToMatsubara, MatsubaraToLegendre, MeshImFreq, MeshImTime, MeshLegendre,
                               make_gf_from_fourier)
    from triqs_b200.context import TriqsRuntimeError
    F = triqs_b200.FERMION
    beta, n_iw, n_tau = 10.0, 200, 1201
    mw, mt = MeshImFreq(beta, F, n_iw), MeshImTime(beta, F, n_tau)
    iw = O.imfreq_values(beta, F, n_iw)
    gw = Gf(mw, (2, 2))
    for i in range(2):
        for j in range(2):
            gw.data[:, i, j] = (1.0 if i == j else 0.2) / (iw - 0.5 * (i + 1) + 0.3 * j)
    gt = Gf(mt, (2, 2))
    assert (gt << Fourier(gw)) is gt
    assert np.array_equal(gt.data, make_gf_from_fourier(gw, n_tau).data)
    gw2 = Gf(mw, (2, 2))
    gw2 << Fourier(gt)
    assert np.max(np.abs(gw2.data - gw.data)) < 1e-6
    # a BlockGf of equally shaped blocks: one batched launch
    bw = BlockGf(["up", "down"], [gw, gw2])
    bt = BlockGf(["up", "down"], [Gf(mt, (2, 2)), Gf(mt, (2, 2))])
    bt << Fourier(bw)
    assert np.array_equal(bt["up"].data, gt.data)
    # scalars, gfs, mesh mismatch
    gw2 << 0.5
    assert gw2.data[3, 0, 0] == 0.5 and gw2.data[3, 0, 1] == 0
    gw2 << gw
    assert np.array_equal(gw2.data, gw.data)
    with pytest.raises(TriqsRuntimeError):
        gt << Fourier(gt)
    with pytest.raises(TriqsRuntimeError):
        Gf(MeshImTime(2 * beta, F, n_tau), (2, 2)) << Fourier(gw)
    # Legendre
    gl = Gf(MeshLegendre(beta, F, 30), (2, 2))
    gl << MatsubaraToLegendre(gt)
    gw3 = Gf(mw, (2, 2))
    gw3 << LegendreToMatsubara(gl)
    from triqs_b200.gf import legendre_to_matsubara
    assert np.array_equal(gw3.data, legendre_to_matsubara(gl, mw).data)
    assert np.max(np.abs(gw3.data - gw.data)) < 5e-3  # (30 Legendre coefficients)
