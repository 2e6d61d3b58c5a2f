"""Python host mirror of the reference's gf surface for the hot path (names follow python/triqs/gf).

    MeshImTime, MeshImFreq, MeshBrZone, MeshCycLat        python/triqs/gf/meshes.py (C++: c++/triqs/mesh/*.hpp)
    Gf, BlockGf                                           python/triqs/gf/gf.py, block_gf.py
    make_gf_from_fourier, fit_tail, fit_hermitian_tail    python/triqs/gf/gf_fnt_desc.py:24-44,129-170, gf_factories_desc.py:46-124
    make_zero_tail, inverse, Gf.invert, Gf.__call__       python/triqs/gf/gf.py:598-611; c++/triqs/gfs/evaluator.hpp:31-44
    SumkDiscrete-style k-sum with MPI-like slicing        python/triqs/sumk/sumk_discrete.py:140-166, utility/mpi_mpi4py.py:96-109

Only containers and argument plumbing live here; every number is computed by libtqgpu.so through
``triqs_b200.context.Context``.  ``Gf.data`` is a NumPy array (host) or a torch CUDA tensor (device resident).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from .context import BOSON, FERMION, Context, TriqsRuntimeError, _is_torch

_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


# ---------------------------------------------------------------------------------------------
# meshes
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class MeshImTime:
    beta: float
    statistic: int  # FERMION / BOSON
    n_tau: int

    def __len__(self):
        return self.n_tau

    @property
    def delta(self):
        return (self.beta - 0.0) / (self.n_tau - 1)

    def values(self):
        """c++/triqs/mesh/bases/linear.hpp:154-160"""
        wr = np.arange(self.n_tau, dtype=np.float64) / (self.n_tau - 1)
        return 0.0 * (1 - wr) + self.beta * wr


@dataclass(frozen=True)
class MeshImFreq:
    beta: float
    statistic: int
    n_iw: int = 1025  # c++/triqs/mesh/imfreq.hpp:77

    @property
    def last_index(self):
        return self.n_iw - 1

    @property
    def first_index(self):
        return -(self.last_index + (1 if self.statistic == FERMION else 0))

    def __len__(self):
        return self.last_index - self.first_index + 1

    def values(self):
        """c++/triqs/mesh/domains/matsubara.hpp:55"""
        n = np.arange(self.first_index, self.last_index + 1, dtype=np.int64)
        return 1j * (math.pi * (2 * n + self.statistic).astype(np.float64) / self.beta)


@dataclass(frozen=True)
class _MeshLinear:
    """c++/triqs/mesh/bases/linear.hpp: n points on [xmin, xmax], both ends included."""
    xmin: float
    xmax: float
    n: int

    def __len__(self):
        return self.n

    @property
    def delta(self):
        return (self.xmax - self.xmin) / (self.n - 1)

    def values(self):
        wr = np.arange(self.n, dtype=np.float64) / (self.n - 1)
        return self.xmin * (1 - wr) + self.xmax * wr


class MeshReTime(_MeshLinear):
    """c++/triqs/mesh/retime.hpp:39"""


class MeshReFreq(_MeshLinear):
    """c++/triqs/mesh/refreq.hpp:38"""


@dataclass(frozen=True)
class MeshLegendre:
    """c++/triqs/mesh/legendre.hpp: n_l Legendre coefficients of a gf on [0, beta]"""
    beta: float
    statistic: int
    n_l: int

    def __len__(self):
        return self.n_l


@dataclass(frozen=True)
class MeshCycLat:
    dims: tuple

    def __len__(self):
        return int(np.prod(self.dims))


@dataclass(frozen=True)
class MeshBrZone:
    """c++/triqs/mesh/brzone.hpp: dims = points per reciprocal direction; `units` (rows = the mesh steps in k space, brzone.hpp:60-75)
    is only needed to evaluate at a k vector -- default: the simple (hyper)cubic lattice with unit lattice constant, 2 pi / dims."""
    dims: tuple
    units: tuple | None = None

    def __len__(self):
        return int(np.prod(self.dims))

    def k_to_index_units(self, k):
        """transpose(units_inv) * k   (brzone.hpp:368): a k vector in units of the mesh index"""
        d = tuple(self.dims) + (1,) * (3 - len(self.dims))
        U = np.diag([2 * np.pi / x for x in d]) if self.units is None else np.asarray(self.units, dtype=float)
        k = np.atleast_2d(np.asarray(k, dtype=float))
        if k.shape[1] < 3:
            k = np.concatenate([k, np.zeros((k.shape[0], 3 - k.shape[1]))], axis=1)
        return k @ np.linalg.inv(U)


@dataclass(frozen=True)
class MeshProduct:
    """c++/triqs/mesh/prod.hpp:72-208: a tuple of meshes; the data rank grows by one per mesh (C order)."""
    meshes: tuple

    def __len__(self):
        return int(np.prod([len(m) for m in self.meshes]))

    @property
    def shape(self):
        return tuple(len(m) for m in self.meshes)

    def __getitem__(self, i):
        return self.meshes[i]


def make_adjoint_mesh(m, n=-1):
    """c++/triqs/mesh/adjoint.hpp:31-57"""
    if isinstance(m, MeshImTime):
        return MeshImFreq(m.beta, m.statistic, (m.n_tau - 1) // 6 if n == -1 else n)
    if isinstance(m, MeshImFreq):
        return MeshImTime(m.beta, m.statistic, 6 * (m.last_index + 1) + 1 if n == -1 else n)
    if isinstance(m, MeshCycLat):
        return MeshBrZone(m.dims)
    if isinstance(m, MeshBrZone):
        return MeshCycLat(m.dims)
    if isinstance(m, (MeshReTime, MeshReFreq)):  # adjoint.hpp:43-54; `n` plays the role of shift_half_bin (True / False)
        xm = math.pi * (m.n - 1) / (m.n * m.delta)
        sh = math.pi / m.n / m.delta if n is True else 0.0
        return (MeshReFreq if isinstance(m, MeshReTime) else MeshReTime)(-xm + sh, xm + sh, m.n)
    raise TypeError(m)


# ---------------------------------------------------------------------------------------------
# containers
# ---------------------------------------------------------------------------------------------
class Gf:
    """mesh + data[mesh, target...] (C order, complex128) -- c++/triqs/gfs/gf/gf.hpp:86-375."""

    def __init__(self, mesh, target_shape=(), data=None, ctx: Context | None = None, is_real=False):
        """is_real: real-valued target (python/triqs/gf/gf.py `is_real`, gf<mesh, T::real_t>): float64 data.  Fourier transforms
        of a real-valued G(tau) move only the reals to the GPU (Context.fourier_tau_to_iw)."""
        self.mesh = mesh
        self.target_shape = tuple(target_shape)
        shape = (mesh.shape if isinstance(mesh, MeshProduct) else (len(mesh),)) + self.target_shape
        if data is None:
            data = np.zeros(shape, dtype=np.float64 if is_real else np.complex128)
        elif tuple(data.shape) != shape:
            raise TriqsRuntimeError(f"Gf: data shape {tuple(data.shape)} does not match mesh/target {shape}")
        self.data = data
        self._ctx = ctx

    @property
    def ctx(self) -> Context:
        return self._ctx or default_context()

    @property
    def n_others(self):
        return int(np.prod(self.target_shape)) if self.target_shape else 1

    def flat(self):
        """flatten_gf_2d<0> (c++/triqs/gfs/gf/flatten.hpp:61-64): a view, no copy -- the layout is already [mesh, n_others]."""
        return self.data.reshape(len(self.mesh), self.n_others)

    def copy(self):
        d = self.data.clone() if _is_torch(self.data) else self.data.copy()
        return Gf(self.mesh, self.target_shape, d, self._ctx)

    def invert(self):
        """python/triqs/gf/gf.py:598-611 -> _gf_invert_data_in_place"""
        n = self.target_shape[0]
        if len(self.target_shape) != 2 or self.target_shape[1] != n:
            raise TriqsRuntimeError("invert: matrix_valued square target required")
        d = self.data
        contiguous = d.is_contiguous() if _is_torch(d) else d.flags.c_contiguous
        if contiguous:
            self.ctx.invert_in_place(d.reshape(-1, n, n))  # a view: inverted in place
        else:  # strided view (e.g. a slice of a larger gf): reshape would copy, so invert a dense copy and write it back
            dense = d.contiguous() if _is_torch(d) else np.ascontiguousarray(d)
            self.ctx.invert_in_place(dense.reshape(-1, n, n))
            self.data[...] = dense
        return self

    # -- the Python call surface of the transforms: `g << Fourier(g_other)` (python/triqs/gf/descriptors.py:156-196) ------------
    def _assign(self, data):
        if tuple(data.shape) != tuple(self.data.shape):
            raise TriqsRuntimeError(f"Gf: cannot assign data of shape {tuple(data.shape)} to a gf of shape {tuple(self.data.shape)}")
        if _is_torch(self.data) and not _is_torch(data):
            import torch
            data = torch.from_numpy(np.ascontiguousarray(data)).to(self.data.device)
        elif _is_torch(data) and not _is_torch(self.data):
            data = data.cpu().numpy()
        if not _is_real_data(self.data) or _is_real_data(data):
            self.data[...] = data
        else:
            raise TriqsRuntimeError("Gf: cannot store complex values in a real-valued gf (use real(g) after is_gf_real(g))")
        return self

    def set_from_fourier(self, g, known_moments=None):
        """G2.set_from_fourier(G[, known_moments]) (python/triqs/gf/gf_fnt.py via cpp2py; fourier.hpp:257-276): the transform of `g`
        ONTO this gf's mesh -- imtime <-> imfreq, retime <-> refreq, cyclat <-> brzone."""
        if tuple(g.target_shape) != tuple(self.target_shape):
            raise TriqsRuntimeError("set_from_fourier: target shapes differ")
        pairs = {MeshImTime: MeshImFreq, MeshImFreq: MeshImTime, MeshReTime: MeshReFreq, MeshReFreq: MeshReTime, MeshCycLat: MeshBrZone,
                 MeshBrZone: MeshCycLat}
        if pairs.get(type(g.mesh)) is not type(self.mesh):  # (the reference: static_assert on the adjoint mesh types, fourier.hpp:101,106)
            raise TriqsRuntimeError(f"set_from_fourier: {type(self.mesh).__name__} is not the adjoint mesh type of {type(g.mesh).__name__}")
        for attr in ("beta", "statistic", "dims"):
            if hasattr(g.mesh, attr) and getattr(g.mesh, attr) != getattr(self.mesh, attr):
                raise TriqsRuntimeError(f"set_from_fourier: the meshes differ in {attr}")
        res = make_gf_from_fourier(g, self.mesh, known_moments)
        return self._assign(res.data)

    def set_from_legendre(self, gl):
        """descriptors.py:170-182"""
        return self._assign(legendre_to_matsubara(gl, self.mesh).data)

    def set_from_imfreq(self, g):
        """descriptors.py:184-196 (MatsubaraToLegendre; the argument may live on an imfreq or an imtime mesh)"""
        return self._assign(matsubara_to_legendre(g, len(self.mesh)).data)

    set_from_imtime = set_from_imfreq

    def __lshift__(self, rhs):
        """`g << rhs` (python/triqs/gf/gf.py:486-519): a lazy descriptor (Fourier, LegendreToMatsubara, MatsubaraToLegendre), another gf
        on the same mesh, an array, or a scalar (times the identity for a matrix-valued target)."""
        if callable(rhs) and not isinstance(rhs, Gf):
            return rhs(self)
        if isinstance(rhs, Gf):
            if rhs.mesh != self.mesh:
                raise TriqsRuntimeError("Gf <<: the meshes differ")
            return self._assign(rhs.data)
        if np.isscalar(rhs):
            self.data[...] = 0
            if len(self.target_shape) == 2 and self.target_shape[0] == self.target_shape[1]:
                for i in range(self.target_shape[0]):
                    self.data[..., i, i] = rhs
            else:
                self.data[...] = rhs
            return self
        return self._assign(rhs)

    def _call_prod(self, args):
        """g(x1, x2, ...) on a product mesh (or a single linear / brzone mesh): c++/triqs/mesh/evaluate.hpp:97-114 via tq_evaluate_prod.
        Per component: imtime / retime / refreq take values, brzone a k vector (or an array [n_pts, 3] of them), imfreq a Matsubara
        index inside the mesh, cyclat a linear index."""
        meshes = self.mesh.meshes if isinstance(self.mesh, MeshProduct) else (self.mesh,)
        if len(args) != len(meshes):
            raise TriqsRuntimeError(f"evaluation of a gf on {len(meshes)} meshes needs {len(meshes)} arguments")
        axes, cols, shape, scalar = [], [], [], True
        for m, x in zip(meshes, args):
            if isinstance(m, MeshImTime):
                axes.append(("linear", 0.0, m.beta))
                shape.append(len(m))
                cols.append(np.atleast_1d(np.asarray(x, dtype=float)))
                scalar = scalar and np.isscalar(x)
            elif isinstance(m, _MeshLinear):
                axes.append(("linear", m.xmin, m.xmax))
                shape.append(len(m))
                cols.append(np.atleast_1d(np.asarray(x, dtype=float)))
                scalar = scalar and np.isscalar(x)
            elif isinstance(m, MeshBrZone):
                v = m.k_to_index_units(x)
                d = tuple(m.dims) + (1,) * (3 - len(m.dims))
                for j in range(3):
                    axes.append("periodic")
                    shape.append(d[j])
                    cols.append(v[:, j])
                scalar = scalar and np.ndim(x) == 1
            elif isinstance(m, MeshImFreq):
                n = np.atleast_1d(np.asarray(x, dtype=np.int64))
                if np.any(n < m.first_index) or np.any(n > m.last_index):
                    raise TriqsRuntimeError("product-mesh evaluation: Matsubara index outside the mesh")
                axes.append("index")
                shape.append(len(m))
                cols.append((n - m.first_index).astype(float))
                scalar = scalar and np.isscalar(x)
            else:
                axes.append("index")
                shape.append(len(m))
                cols.append(np.atleast_1d(np.asarray(x, dtype=float)))
                scalar = scalar and np.isscalar(x)
        n_pts = max(len(c) for c in cols)
        coords = np.stack([np.broadcast_to(c, (n_pts,)) for c in cols], axis=1)
        d = self.data
        table = d.reshape(tuple(shape) + (-1,))
        out = self.ctx.evaluate_prod(table, axes, coords)
        out = out.reshape((n_pts,) + tuple(self.target_shape))
        return out[0] if scalar else out

    def __call__(self, tau, *more):
        """off-mesh evaluation: linear interpolation on an imtime mesh (c++/triqs/gfs/evaluator.hpp:35-43; a scalar or an array
        of tau), or g(n) on an imfreq mesh -- the mesh value inside the window, the fitted tail outside (evaluator.hpp:50-76);
        product meshes, real-time / real-frequency and brzone meshes: multi-linear interpolation (mesh/evaluate.hpp:97-114)."""
        if more or isinstance(self.mesh, (MeshProduct, MeshBrZone, _MeshLinear)):
            return self._call_prod((tau,) + more)
        if isinstance(self.mesh, MeshImFreq):
            m = self.mesh
            if not np.isscalar(tau):  # an array of Matsubara indices: one batched call (tq_evaluate_imfreq)
                ns = np.asarray(tau, dtype=np.int64).reshape(-1)
                out = self.ctx.evaluate_imfreq(self.flat()[None], m.beta, m.statistic, ns)[0]
                return out.reshape((len(ns),) + tuple(self.target_shape))
            n = int(tau)
            if m.first_index <= n <= m.last_index:
                return self.data[n - m.first_index]
            tail, _ = fit_tail(self)  # moments A_p: g = sum_p A_p / (i w)^p  (fit_tail_no_normalize with x = |w_max| / i w is the same sum)
            iw = 1j * math.pi * (2 * n + m.statistic) / m.beta
            z, res = 1.0 + 0j, 0
            tail = tail.cpu().numpy() if _is_torch(tail) else tail
            for p in range(tail.shape[0]):
                res = res + tail[p] * z
                z = z / iw
            return res
        if not isinstance(self.mesh, MeshImTime):
            raise TriqsRuntimeError("Gf.__call__ is implemented for imtime meshes")
        scalar = np.isscalar(tau)
        taus = np.atleast_1d(np.asarray(tau, dtype=np.float64)) if not _is_torch(tau) else tau
        out = self.ctx.interp_tau(self.flat(), self.mesh.beta, taus)
        out = out.reshape((out.shape[0],) + self.target_shape)
        return out[0] if scalar else out


class _LazyTransform:
    """BaseBlock of python/triqs/gf/descriptors.py: a lazy expression `X(G)`; `G2 << X(G)` calls X(G)(G2).  Applied to a BlockGf it
    acts block by block -- equally shaped blocks in ONE batched launch (map_block_gf, block/map.hpp:35-40)."""
    _method = None

    def __init__(self, G, *args, **kw):
        self.G, self.args, self.kw = G, args, kw

    def __call__(self, G2):
        if isinstance(G2, BlockGf):
            src = self.G
            if not isinstance(src, BlockGf) or len(src) != len(G2):
                raise TriqsRuntimeError("lazy transform: BlockGf arguments of equal length are required")
            b0, s0 = G2.blocks[0], src.blocks[0]
            same = all(b.mesh == b0.mesh and b.target_shape == b0.target_shape for b in G2.blocks) and all(
                b.mesh == s0.mesh and b.target_shape == s0.target_shape for b in src.blocks)
            if same and self._method == "set_from_fourier" and not self.args and not self.kw:
                res = make_gf_from_fourier(src, b0.mesh)
                for b, r in zip(G2.blocks, res.blocks):
                    b._assign(r.data)
                return G2
            for b, sb in zip(G2.blocks, src.blocks):
                getattr(b, self._method)(sb, *self.args, **self.kw)
            return G2
        getattr(G2, self._method)(self.G, *self.args, **self.kw)
        return G2


class Fourier(_LazyTransform):
    """descriptors.py:156-168: `gw << Fourier(gt)`, `gt << Fourier(gw)` (any pair of adjoint meshes)"""
    _method = "set_from_fourier"


class LegendreToMatsubara(_LazyTransform):
    """descriptors.py:170-182"""
    _method = "set_from_legendre"


class MatsubaraToLegendre(_LazyTransform):
    """descriptors.py:184-196"""
    _method = "set_from_imfreq"


def _is_real_data(d):
    return (not d.dtype.is_complex) if _is_torch(d) else (not np.iscomplexobj(d))


class BlockGf:
    """named list of Gf -- c++/triqs/gfs/block/block_gf.hpp:104-295, python/triqs/gf/block_gf.py."""

    def __lshift__(self, rhs):
        """`G << Fourier(G_other)` etc. block by block (python/triqs/gf/block_gf.py:306-316)"""
        if callable(rhs) and not isinstance(rhs, BlockGf):
            return rhs(self)
        if isinstance(rhs, BlockGf):
            for b, r in zip(self.blocks, rhs.blocks):
                b << r
            return self
        for b in self.blocks:
            b << rhs
        return self

    def __init__(self, name_list, block_list):
        if len(name_list) != len(block_list):
            raise TriqsRuntimeError("BlockGf: names and blocks differ in length")
        self.names = list(name_list)
        self.blocks = list(block_list)

    def __iter__(self):
        return iter(zip(self.names, self.blocks))

    def __getitem__(self, key):
        return self.blocks[self.names.index(key)] if isinstance(key, str) else self.blocks[key]

    def __len__(self):
        return len(self.blocks)


def is_gf_real(g, tolerance=1e-13) -> bool:
    """c++/triqs/gfs/functions/functions2.hpp:223-232: max |Im g| <= tolerance (a BlockGf: every block)."""
    if isinstance(g, BlockGf):
        return all(is_gf_real(b, tolerance) for b in g.blocks)
    d = g.data
    if _is_torch(d):
        return (not d.dtype.is_complex) or float(d.imag.abs().max()) <= tolerance
    return (not np.iscomplexobj(d)) or float(np.max(np.abs(d.imag))) <= tolerance


def real(g):
    """c++/triqs/gfs/functions/functions2.hpp:242-248: the real-valued gf {g.mesh(), real(g.data())}."""
    if isinstance(g, BlockGf):
        return BlockGf(g.names, [real(b) for b in g.blocks])
    d = g.data
    if _is_torch(d):
        d = d.real.contiguous() if d.dtype.is_complex else d
    else:
        d = np.ascontiguousarray(d.real) if np.iscomplexobj(d) else d
    return Gf(g.mesh, g.target_shape, d, g._ctx)


def make_zero_tail(g: Gf, n_moments=10):
    """c++/triqs/gfs/functions/functions2.hpp:154-162"""
    return np.zeros((n_moments,) + g.target_shape, dtype=np.complex128)


# ---------------------------------------------------------------------------------------------
# Fourier
# ---------------------------------------------------------------------------------------------
def _stack(blocks):
    d0 = blocks[0].flat()
    if _is_torch(d0):
        import torch
        return torch.stack([b.flat() for b in blocks])
    return np.stack([b.flat() for b in blocks])


def make_gf_from_fourier_axes(g: Gf, axes, meshes=None) -> Gf:
    """make_gf_from_fourier<N1, N2, ...>(g[, m1, m2, ...]) on a product mesh (fourier.hpp:60-85, :159-181;
    test/c++/gfs/multivar/fourier.cpp).  The reference copies the data twice per transformed mesh (flatten_2d / unflatten_2d,
    flatten.hpp:33-47); here mesh N of [n_0 .. n_N .. n_k, target] is simply the middle index of [outer][n_N][inner]."""
    if not isinstance(g.mesh, MeshProduct):
        raise TriqsRuntimeError("make_gf_from_fourier<N...>: a product mesh is required")
    axes = tuple(axes)
    meshes = tuple(meshes) if meshes is not None else (None,) * len(axes)
    cur_meshes, data = list(g.mesh.meshes), g.data
    for N, mo in zip(axes, meshes):
        mi = cur_meshes[N]
        mo = mo if mo is not None else make_adjoint_mesh(mi)
        shp = tuple(len(m) for m in cur_meshes)
        outer = int(np.prod(shp[:N])) if N else 1
        inner = int(np.prod(shp[N + 1:])) * g.n_others
        d3 = data.reshape(outer, shp[N], inner)
        out = _transform(g.ctx, mi, d3, mo, None, one_gf=True)
        cur_meshes[N] = mo
        data = out.reshape(tuple(len(m) for m in cur_meshes) + g.target_shape)
    return Gf(MeshProduct(tuple(cur_meshes)), g.target_shape, data, g._ctx)


def make_gf_from_fourier(g, n=-1, known_moments=None):
    """fourier.hpp:93-245: imtime <-> imfreq and cyclat <-> brzone; a BlockGf of equally shaped blocks is ONE batched launch."""
    if isinstance(g, BlockGf):
        same = all(b.mesh == g.blocks[0].mesh and b.target_shape == g.blocks[0].target_shape for b in g.blocks)
        if not same or known_moments is not None:
            kms = known_moments if known_moments is not None else [None] * len(g)
            return BlockGf(g.names, [make_gf_from_fourier(b, n, km) for b, km in zip(g.blocks, kms)])
        b0 = g.blocks[0]
        out = _transform(b0.ctx, b0.mesh, _stack(g.blocks), n, None)
        mo = _out_mesh(b0.mesh, n)
        return BlockGf(g.names, [Gf(mo, b0.target_shape, out[i].reshape((len(mo),) + b0.target_shape), b0._ctx) for i in range(len(g))])
    km = None
    if known_moments is not None:
        km = known_moments.reshape(1, known_moments.shape[0], g.n_others)
    out = _transform(g.ctx, g.mesh, g.flat()[None], n, km)
    mo = _out_mesh(g.mesh, n)
    return Gf(mo, g.target_shape, out[0].reshape((len(mo),) + g.target_shape), g._ctx)


def legendre_to_matsubara(gl: "Gf", mesh) -> "Gf":
    """`g << LegendreToMatsubara(gl)` (python/triqs/gf/descriptors.py:170-182; legendre_matsubara.hpp:37-70): onto an imfreq
    or an imtime mesh."""
    if not isinstance(gl.mesh, MeshLegendre):
        raise TriqsRuntimeError("LegendreToMatsubara needs a Legendre gf")
    if isinstance(mesh, MeshImFreq):
        out = gl.ctx.legendre_to_imfreq(gl.flat()[None], mesh.statistic, mesh.n_iw)
    elif isinstance(mesh, MeshImTime):
        out = gl.ctx.legendre_to_imtime(gl.flat()[None], mesh.beta, mesh.n_tau)
    else:
        raise TypeError(mesh)
    return Gf(mesh, gl.target_shape, out[0].reshape((len(mesh),) + gl.target_shape), gl._ctx)


def matsubara_to_legendre(g: "Gf", n_l: int) -> "Gf":
    """`gl << MatsubaraToLegendre(g)` (descriptors.py:184-196; legendre_matsubara.hpp:76-118) from an imfreq or an imtime gf."""
    m = g.mesh
    if isinstance(m, MeshImFreq):
        out = g.ctx.imfreq_to_legendre(g.flat()[None], m.beta, m.statistic, n_l)
    elif isinstance(m, MeshImTime):
        out = g.ctx.imtime_to_legendre(g.flat()[None], m.beta, n_l)
    else:
        raise TypeError(m)
    ml = MeshLegendre(m.beta, m.statistic, n_l)
    return Gf(ml, g.target_shape, out[0].reshape((n_l,) + g.target_shape), g._ctx)


def _out_mesh(mesh, n):
    """the target mesh: given explicitly (make_gf_from_fourier(g, mesh, ...), fourier.hpp:93-114) or the adjoint one"""
    if hasattr(n, "__len__") and not isinstance(n, (int, bool)):
        return n
    return make_adjoint_mesh(mesh, n)


def _transform(ctx, mesh, data3, n, km, one_gf=False):
    if isinstance(mesh, MeshImTime):
        mo = _out_mesh(mesh, n)
        return ctx.fourier_tau_to_iw(data3, mesh.beta, mesh.statistic, mo.n_iw, moments=km, one_gf=one_gf)
    if isinstance(mesh, MeshImFreq):
        mo = _out_mesh(mesh, n)
        return ctx.fourier_iw_to_tau(data3, mesh.beta, mesh.statistic, mo.n_tau, moments=km)
    if isinstance(mesh, (MeshReTime, MeshReFreq)):  # fourier_real.cpp:35-131
        mo = n if isinstance(n, (MeshReTime, MeshReFreq)) else make_adjoint_mesh(mesh, n)
        if isinstance(mesh, MeshReTime):
            return ctx.fourier_t_to_w(data3, (mesh.xmin, mesh.xmax), (mo.xmin, mo.xmax), moments=km)
        return ctx.fourier_w_to_t(data3, (mo.xmin, mo.xmax), (mesh.xmin, mesh.xmax), moments=km)
    if isinstance(mesh, (MeshCycLat, MeshBrZone)):
        outs = [ctx.fourier_lattice(d, tuple(mesh.dims), to_k=isinstance(mesh, MeshCycLat)) for d in data3]
        if _is_torch(outs[0]):
            import torch
            return torch.stack(outs)
        return np.stack(outs)
    raise TypeError(mesh)


def fit_tail(g: Gf, known_moments=None, hermitian=False, **fit_params):
    """functions2.hpp:50-56 (and :94-110 with hermitian=True): returns (tail[n_moments, target...], err)."""
    if isinstance(g.mesh, MeshReFreq) and not hermitian:  # fit_tail_real.cpp
        km = None if known_moments is None else np.asarray(known_moments).reshape(1, np.shape(known_moments)[0], g.n_others)
        tail, err = g.ctx.fit_tail_refreq(g.flat()[None], (g.mesh.xmin, g.mesh.xmax), known_moments=km, **fit_params)
        return tail[0].reshape((tail.shape[1],) + g.target_shape), float(err[0])
    if not isinstance(g.mesh, MeshImFreq):
        raise TriqsRuntimeError("fit_tail needs an imfreq gf")
    inner = 1
    if hermitian:
        if len(g.target_shape) == 0:
            inner = 1
        elif len(g.target_shape) == 2 and g.target_shape[0] == g.target_shape[1]:
            inner = g.target_shape[0]
        else:
            raise TriqsRuntimeError("Incompatible target_shape for fit_hermitian_tail")  # functions2.hpp:105
    km = None if known_moments is None else known_moments.reshape(1, known_moments.shape[0], g.n_others)
    mom, err = g.ctx.fit_tail(g.flat()[None], g.mesh.beta, g.mesh.statistic, known_moments=km, hermitian=hermitian, inner_dim=inner,
                              **fit_params)
    return mom[0].reshape((mom.shape[1],) + g.target_shape), float(err[0])


def fit_hermitian_tail(g, known_moments=None, **fit_params):
    if isinstance(g, BlockGf):  # functions2.hpp:121-134: list of tails, maximum error
        res = [fit_tail(b, None if known_moments is None else known_moments[i], hermitian=True, **fit_params) for i, b in enumerate(g.blocks)]
        return [r[0] for r in res], max(r[1] for r in res)
    return fit_tail(g, known_moments, hermitian=True, **fit_params)


def _herm_view(g):
    """(data [1, n_pts, d, d], imtime?) of an imfreq / imtime gf with target rank 0, 2 or 4 (imfreq.hpp:81-215)."""
    if not isinstance(g.mesh, (MeshImFreq, MeshImTime)):
        raise TriqsRuntimeError("requires an imfreq or imtime Green function")
    ts = g.target_shape
    if len(ts) == 0:
        d = 1
    elif len(ts) == 2 and ts[0] == ts[1]:
        d = ts[0]
    elif len(ts) == 4 and ts[0] == ts[2] and ts[1] == ts[3]:
        d = ts[0] * ts[1]
    else:
        raise TriqsRuntimeError("requires a Green function with a (square) target of rank 0, 2 or 4")
    return g.flat().reshape(1, -1, d, d), isinstance(g.mesh, MeshImTime)


def make_hermitian(g):
    """imfreq.hpp:175-215 (block overload: map over blocks)."""
    if isinstance(g, BlockGf):
        return BlockGf(g.names, [make_hermitian(b) for b in g.blocks])
    data, imtime = _herm_view(g)
    out = g.ctx.make_hermitian(data, imtime=imtime)
    return Gf(g.mesh, g.target_shape, out.reshape((len(g.mesh),) + tuple(g.target_shape)), g.ctx)


def is_gf_hermitian(g, tolerance: float = 1e-12) -> bool:
    """imfreq.hpp:81-124."""
    if isinstance(g, BlockGf):
        return all(is_gf_hermitian(b, tolerance) for b in g.blocks)
    data, imtime = _herm_view(g)
    return g.ctx.is_gf_hermitian(data, imtime=imtime, tolerance=tolerance)


def replace_by_tail(g: Gf, tail, n_min: int) -> None:
    """imfreq.hpp:258-261: in place."""
    if not isinstance(g.mesh, MeshImFreq):
        raise TriqsRuntimeError("replace_by_tail needs an imfreq gf")
    if _is_torch(g.data):
        import torch
        d = g.data
        dense = d if d.is_contiguous() else d.contiguous()
        t = tail if _is_torch(tail) else torch.as_tensor(np.asarray(tail), device=d.device)
        g.ctx.replace_by_tail(dense.reshape(1, len(g.mesh), g.n_others), g.mesh.beta, g.mesh.statistic,
                              t.reshape(1, t.shape[0], g.n_others), n_min)
        if dense is not d:
            d[...] = dense
        return
    dense = np.ascontiguousarray(g.data, dtype=np.complex128)
    if not dense.flags.writeable or dense is not g.data:
        dense = dense.copy()
    g.ctx.replace_by_tail(dense.reshape(1, len(g.mesh), g.n_others), g.mesh.beta, g.mesh.statistic,
                          np.asarray(tail).reshape(1, np.shape(tail)[0], g.n_others), n_min)
    g.data[...] = dense


def density(g, known_moments=None):
    """density.cpp:32-128: n[target] of an imfreq gf (matrix- or scalar-valued); for a BlockGf the list over blocks."""
    if isinstance(g, BlockGf):
        return [density(b, None if known_moments is None else known_moments[i]) for i, b in enumerate(g.blocks)]
    if not isinstance(g.mesh, MeshImFreq):
        raise TriqsRuntimeError("density needs an imfreq gf")
    ts = g.target_shape
    if len(ts) == 0:
        n = 1
    elif len(ts) == 2 and ts[0] == ts[1]:
        n = ts[0]
    else:
        raise TriqsRuntimeError("density: the target must be scalar or a square matrix")
    data = g.flat().reshape(1, -1, n, n)
    km = None if known_moments is None else known_moments.reshape(1, known_moments.shape[0], n, n)
    res = g.ctx.density(data, g.mesh.beta, g.mesh.statistic, moments=km)[0]
    return res[0, 0] if len(ts) == 0 else res


def inverse(g: Gf) -> Gf:
    """functions2.hpp:203-209"""
    return g.copy().invert()


# ---------------------------------------------------------------------------------------------
# lattice Dyson k-sum, sharded like SumkDiscrete (sumk_discrete.py:140-166)
# ---------------------------------------------------------------------------------------------
def slice_bounds(n: int, size: int, rank: int):
    """mpi.slice_array bounds (python/triqs/utility/mpi_mpi4py.py:96-109): the first n % size ranks get one extra."""
    q, r = divmod(n, size)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def dyson_k_sum(ctx: Context, eps_k, sigma: Gf, mu: float, weights=None, rank: int = 0, n_ranks: int = 1, reduce_fn=None):
    """G_loc = sum_k w_k [(iw + mu) - eps_k - Sigma]^-1 with k-points sliced over ranks.

    ``reduce_fn(partial)`` sums the per-rank partial over ranks: None -> the library's NCCL all-reduce (needs
    ctx.comm_init and a device-resident sigma), or e.g. a torch.distributed all_reduce (gloo on CPU tests)."""
    n_k = eps_k.shape[0]
    lo, hi = slice_bounds(n_k, n_ranks, rank)
    w = None if weights is None else weights[lo:hi]
    n = sigma.target_shape[0]
    use_lib = reduce_fn is None and n_ranks > 1
    part = ctx.dyson_ksum(eps_k[lo:hi], sigma.data.reshape(-1, n, n), sigma.mesh.beta, sigma.mesh.statistic, mu, weights=w,
                          uniform_weight=1.0 / n_k, all_reduce=use_lib)
    if reduce_fn is not None and n_ranks > 1:
        part = reduce_fn(part)
    return Gf(sigma.mesh, sigma.target_shape, part.reshape((len(sigma.mesh),) + sigma.target_shape), ctx)


# ---------------------------------------------------------------------------------------------
# HDF5 archives (c++/triqs/gfs/h5.hpp:72-91; mesh groups imfreq.hpp:261-292, imtime.hpp:83-106, brzone.hpp:311-343, prod.hpp:179-195)
# ---------------------------------------------------------------------------------------------
def h5_write(path, **objects):
    """h5_write("file.h5", G=g, Sigma=bg): one group per object with the reference's layout (Format tags "Gf" / "BlockGf",
    `data` with the __complex__ attribute, `mesh` sub-group).  No libhdf5 needed (triqs_b200/h5lite.py)."""
    from . import h5lite
    h5lite.write_gf(path, objects)


def h5_read(path, name=None, ctx=None):
    """-> the Gf / BlockGf stored under `name`, or a dict of all of them; reads archives written by TRIQS itself."""
    from . import h5lite
    return h5lite.read_gf(path, name, ctx)
