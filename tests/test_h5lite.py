"""HDF5 Gf archives without libhdf5: the minimal reader / writer of triqs_b200/h5lite.py.

The reference's own archives are not on the GPU box, so what was read from their bytes here is pinned as a small golden
fixture (tests/golden/h5_reference_digest.json, written by this file's `_make_digest` when /root/reference exists)."""
import json
import os

import numpy as np
import pytest

from triqs_b200 import h5lite

HERE = os.path.dirname(os.path.abspath(__file__))
DIGEST = os.path.join(HERE, "golden", "h5_reference_digest.json")
REF_FILES = {"block_pyh5_ref": "/root/reference/test/c++/gfs/block_pyh5_ref.h5", "gf_init": "/root/reference/test/python/base/gf_init.ref.h5",
             "pade": "/root/reference/test/python/base/pade.ref.h5", "dos": "/root/reference/test/python/base/dos.ref.h5",
             "U_mat": "/root/reference/test/python/base/U_mat.ref.h5"}


def _flatten(tree, attrs, path=""):
    out = {}
    for k, v in tree.items():
        p = (path + "/" + k) if path else k
        if isinstance(v, dict):
            out[p + "/"] = {a: str(b) for a, b in attrs.get(p, {}).items()}
            out.update(_flatten(v, attrs, p))
        elif isinstance(v, np.ndarray) and v.dtype != object:
            out[p] = {"shape": list(v.shape), "dtype": str(v.dtype), "sum": float(np.sum(v.astype(np.float64))),
                      "abs_sum": float(np.sum(np.abs(v.astype(np.float64)))), "attrs": {a: str(b) for a, b in attrs.get(p, {}).items()}}
        else:
            out[p] = {"value": [str(x) for x in v.reshape(-1)] if isinstance(v, np.ndarray) else str(v)}
    return out


def _make_digest():
    d = {}
    for name, f in REF_FILES.items():
        tree, attrs = h5lite.read(f)
        d[name] = _flatten(tree, attrs)
    json.dump(d, open(DIGEST, "w"), indent=1, sort_keys=True)


def test_reader_against_the_reference_archives():
    """every archive the reference ships for this path parses (old-style groups, compact / contiguous / chunked+deflate datasets,
    fixed and variable-length strings) and matches the digest taken from its bytes"""
    if not os.path.exists(REF_FILES["block_pyh5_ref"]):
        pytest.skip("/root/reference not present (GPU box): the digest fixture was generated where it is")
    if not os.path.exists(DIGEST):
        _make_digest()
    want = json.load(open(DIGEST))
    for name, f in REF_FILES.items():
        tree, attrs = h5lite.read(f)
        got = json.loads(json.dumps(_flatten(tree, attrs)))
        assert got == want[name], name
    # the layout this repository must emit: block_gf of two imfreq gfs (test/c++/gfs/block_pyh5_ref.h5)
    tree, attrs = h5lite.read(REF_FILES["block_pyh5_ref"])
    assert attrs["bgf"]["Format"] == "BlockGf" and attrs["bgf/0"]["Format"] == "Gf" and attrs["bgf/0/mesh"]["Format"] == "MeshImFreq"
    assert tree["bgf"]["0"]["data"].shape == (200, 2) and attrs["bgf/0/data"]["__complex__"] == "1"
    assert int(tree["bgf"]["0"]["mesh"]["size"]) == 200 and tree["bgf"]["0"]["mesh"]["statistic"] == "F"


def test_reference_block_gf_loads_as_blockgf_and_round_trips(tmp_path):
    if not os.path.exists(REF_FILES["block_pyh5_ref"]):
        pytest.skip("/root/reference not present")
    from triqs_b200 import gf as G
    bg = h5lite.read_gf(REF_FILES["block_pyh5_ref"], "bgf")
    assert isinstance(bg, G.BlockGf) and bg.names == ["0", "1"]
    assert isinstance(bg[0].mesh, G.MeshImFreq) and bg[0].mesh.beta == 10.0 and bg[0].mesh.n_iw == 100 and bg[0].data.shape == (200,)
    out = tmp_path / "again.h5"
    h5lite.write_gf(str(out), {"bgf": bg})
    bg2 = h5lite.read_gf(str(out), "bgf")
    for a, b in zip(bg.blocks, bg2.blocks):
        assert a.mesh == b.mesh and np.array_equal(a.data, b.data)
    # same group / attribute structure as the reference file
    t1, a1 = h5lite.read(REF_FILES["block_pyh5_ref"])
    t2, a2 = h5lite.read(str(out))
    assert sorted(_flatten(t1, a1)) == sorted(_flatten(t2, a2))
    for k in ("bgf", "bgf/0", "bgf/0/mesh", "bgf/block_names", "bgf/0/data"):
        assert a1[k] == a2[k], k


def test_write_read_round_trip_every_mesh(tmp_path):
    from triqs_b200 import gf as G
    rng = np.random.default_rng(0)

    def rnd(*shape):
        return rng.normal(size=shape) + 1j * rng.normal(size=shape)
    objs = {
        "g_iw": G.Gf(G.MeshImFreq(10.0, G.FERMION, 50), (3, 3), rnd(100, 3, 3)),
        "g_tau": G.Gf(G.MeshImTime(7.0, G.BOSON, 101), (2,), rnd(101, 2)),
        "g_w": G.Gf(G.MeshReFreq(-5.0, 4.9, 33), (), rnd(33)),
        "g_t": G.Gf(G.MeshReTime(0.0, 9.0, 19), (1, 1), rnd(19, 1, 1)),
        "g_l": G.Gf(G.MeshLegendre(10.0, G.FERMION, 30), (2, 2), rnd(30, 2, 2)),
        "g_k": G.Gf(G.MeshBrZone((4, 3, 2)), (2, 2), rnd(24, 2, 2)),
        "g_kw": G.Gf(G.MeshProduct((G.MeshBrZone((2, 2, 1)), G.MeshImFreq(5.0, G.FERMION, 8))), (2, 2), rnd(4, 16, 2, 2)),
        "bg": G.BlockGf(["up", "down"], [G.Gf(G.MeshImFreq(10.0, G.FERMION, 20), (5, 5), rnd(40, 5, 5)) for _ in range(2)]),
    }
    f = tmp_path / "all.h5"
    h5lite.write_gf(str(f), objs)
    blob = open(f, "rb").read()
    assert blob[:8] == b"\x89HDF\r\n\x1a\n" and len(blob) % 8 == 0
    back = h5lite.read_gf(str(f))
    assert set(back) == set(objs)
    for k, g in objs.items():
        h = back[k]
        if isinstance(g, G.BlockGf):
            assert h.names == g.names
            for a, b in zip(g.blocks, h.blocks):
                assert a.mesh == b.mesh and np.array_equal(a.data, b.data)
        else:
            assert h.mesh == g.mesh and h.target_shape == g.target_shape and np.array_equal(h.data, g.data)
    # many links in one group: more than one symbol node
    tree = {f"d{i:03d}": np.arange(i + 1, dtype=np.int64) for i in range(70)}
    tree["name"] = "a string"
    tree["sub"] = {"x": 1.5, "empty": np.zeros((0, 3))}
    f2 = tmp_path / "many.h5"
    h5lite.write(str(f2), tree, {"": {"note": "root attribute"}, "sub": {"Format": "Dict"}})
    t, a = h5lite.read(str(f2))
    assert a[""]["note"] == "root attribute" and a["sub"]["Format"] == "Dict" and t["name"] == "a string" and t["sub"]["x"] == 1.5
    assert t["sub"]["empty"].shape == (0, 3)
    for i in range(70):
        assert np.array_equal(t[f"d{i:03d}"], np.arange(i + 1))
