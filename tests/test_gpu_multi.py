"""Hardware parity of the collective: two processes, one GPU each, NCCL communicator of the library.

Pattern of the reference's own tests: test/c++/gfs/mpi_gf.cpp:58-79 (all_reduce of a gf equals size * g) and
python/triqs/sumk/sumk_discrete.py:149-163 (k-slices per rank, mpi.all_reduce of the partial G_loc).  Skips with one device.
"""
import os
import tempfile
import time

import numpy as np
import pytest

from oracle import triqs_oracle as O

pytestmark = pytest.mark.gpu


def _inputs(nk=1000, n=5, n_iw=64):
    rng = np.random.default_rng(4001)
    eps = rng.normal(size=(nk, n, n)) + 1j * rng.normal(size=(nk, n, n))
    eps = 0.3 * (eps + eps.conj().transpose(0, 2, 1))
    iw = O.imfreq_values(10.0, O.FERMION, n_iw)
    V = rng.normal(size=(2, n, 1)) + 1j * rng.normal(size=(2, n, 1))
    sig = sum((V[p] @ V[p].conj().T)[None] / (iw[:, None, None] - e) for p, e in enumerate((-0.6, 0.9)))
    return eps, sig


def _worker(rank, world, idfile, ret):
    import torch
    import triqs_b200
    torch.cuda.set_device(rank)
    ctx = triqs_b200.Context(rank, print_warnings=False)
    if rank == 0:
        uid = triqs_b200.comm_unique_id()
        with open(idfile + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(idfile + ".tmp", idfile)
    else:
        t0 = time.time()
        while not os.path.exists(idfile):
            if time.time() - t0 > 120:
                raise RuntimeError("no NCCL id from rank 0")
            time.sleep(0.05)
        uid = open(idfile, "rb").read()
    ctx.comm_init(world, rank, uid)
    # (1) all_reduce of a gf's data == sum over ranks           mpi_gf.cpp:58-79
    g = torch.arange(2 * 3 * 3, dtype=torch.float64, device=f"cuda:{rank}").reshape(2, 3, 3).to(torch.complex128) * (1 + 2j) * (rank + 1)
    ctx.all_reduce_sum(g)
    ctx.synchronize()
    # (2) k-sum sharded like mpi.slice_array + all-reduce        sumk_discrete.py:149-163
    eps, sig = _inputs()
    lo, hi = O.slice_bounds(eps.shape[0], world, rank)
    gl = ctx.dyson_ksum(torch.from_numpy(eps[lo:hi].copy()).cuda(rank), torch.from_numpy(sig).cuda(rank), 10.0, O.FERMION, 0.1,
                        uniform_weight=1.0 / eps.shape[0], all_reduce=True)
    ctx.synchronize()
    ret[rank] = (g.cpu().numpy(), gl.cpu().numpy())
    ctx.close()


def test_nccl_all_reduce_and_sharded_dyson_ksum_two_gpus():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, os.path.join(d, "nccl_id"), ret), nprocs=world, join=True)
    base = (np.arange(18, dtype=np.float64).reshape(2, 3, 3) * (1 + 2j))
    eps, sig = _inputs()
    ref = O.dyson_ksum(eps, sig, 10.0, O.FERMION, 0.1)
    for r in range(world):
        g, gl = ret[r]
        assert np.array_equal(g, base * 3.0)          # (1 + 2) * g: exact in floating point
        assert np.max(np.abs(gl - ref)) / np.max(np.abs(ref)) < 1e-12
    assert np.array_equal(ret[0][1], ret[1][1])       # every rank holds the same bits after the all-reduce
