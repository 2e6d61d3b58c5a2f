"""N > 1 host logic on CPU: world_size-2 gloo processes shard the k-points like mpi.slice_array
(python/triqs/utility/mpi_mpi4py.py:96-109) and sum the partial G_loc with an all-reduce
(sumk_discrete.py:163).  The GPU context is replaced by a stand-in that evaluates the per-rank partial
with the oracle, so only the sharding / weighting / reduction plumbing of triqs_b200.gf.dyson_k_sum is tested."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import triqs_oracle as O


class _OracleCtx:
    """test stand-in with the signature of triqs_b200.Context.dyson_ksum"""

    def dyson_ksum(self, eps_k, sigma, beta, statistic, mu, weights=None, uniform_weight=None, all_reduce=False, out=None):
        assert not all_reduce
        w = weights if weights is not None else np.full(eps_k.shape[0], uniform_weight)
        if eps_k.shape[0] == 0:
            return np.zeros_like(sigma)
        return O.dyson_ksum(eps_k, sigma, beta, statistic, mu, weights=w)


def _inputs(nk=37, n=3, n_iw=6):
    rng = np.random.default_rng(11)
    eps = rng.normal(size=(nk, n, n)) + 1j * rng.normal(size=(nk, n, n))
    eps = eps + eps.conj().transpose(0, 2, 1)
    iw = O.imfreq_values(10.0, O.FERMION, n_iw)
    sig = np.stack([0.3 * np.eye(n) / (w - 0.4) for w in iw])
    return eps, sig


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from triqs_b200 import gf as G
    eps, sig = _inputs()
    sigma = G.Gf(G.MeshImFreq(10.0, O.FERMION, 6), (3, 3), sig.copy())

    def reduce_fn(part):
        t = torch.from_numpy(np.ascontiguousarray(part).view(np.float64))
        dist.all_reduce(t)
        return t.numpy().view(np.complex128).reshape(part.shape)

    g = G.dyson_k_sum(_OracleCtx(), eps, sigma, 0.1, rank=rank, n_ranks=world, reduce_fn=reduce_fn)
    ret[rank] = g.data
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_k_sharded_dyson_equals_single_rank(world):
    eps, sig = _inputs()
    ref = O.dyson_ksum(eps, sig, 10.0, O.FERMION, 0.1)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        assert np.max(np.abs(ret[r] - ref)) / np.max(np.abs(ref)) < 1e-13


def test_slice_bounds_rule():
    from triqs_b200.gf import slice_bounds
    for n in (0, 1, 7, 262144):
        for size in (1, 2, 3, 8):
            pieces = [slice_bounds(n, size, r) for r in range(size)]
            assert pieces[0][0] == 0 and pieces[-1][1] == n
            assert all(pieces[i][1] == pieces[i + 1][0] for i in range(size - 1))
            lens = [hi - lo for lo, hi in pieces]
            assert max(lens) - min(lens) <= 1 and lens == sorted(lens, reverse=True)
            assert pieces == [O.slice_bounds(n, size, r) for r in range(size)]
