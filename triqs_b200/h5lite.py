"""Minimal HDF5 reader / writer for the fixed layout TRIQS emits for Green's functions.

There is no libhdf5 / h5py in the target image, so this module implements exactly the subset of the HDF5 file format
(version 0 superblock, "old style" groups = v1 B-tree + local heap + symbol nodes, version 1 object headers) that
libhdf5 1.8/1.10/1.12 writes by default and that the reference's archives use:

  * groups, nested;
  * datasets with contiguous or compact layout: float64 / float32 / int8..int64 / uint8..uint64, any rank (rank 0 = scalar);
    complex arrays are stored the TRIQS way: a real array with a trailing dimension of 2 and the attribute ``__complex__ = "1"``
    (h5 library of TRIQS; c++/triqs/gfs/h5.hpp:72-91 writes ``data`` and ``mesh``);
  * strings: fixed length (written) and variable length through the global heap (read: h5py writes attributes that way);
  * attributes (version 1 attribute messages) on groups and datasets -- the ``Format`` tag ("Gf", "BlockGf", "MeshImFreq" ...;
    c++/triqs/gfs/gf/_gf_view_common.hpp:152-171, mesh/imfreq.hpp:261-292, mesh/imtime.hpp:83-106).

Chunked datasets with the deflate / shuffle / fletcher32 filters are READ (the Python archive of TRIQS compresses arrays); the
writer always emits contiguous datasets.  Not supported (raises H5LiteError): new-style (v2 B-tree / fractal heap) groups,
compound types, references.  The reader was checked against the bytes of the reference's own archives
(test/c++/gfs/block_pyh5_ref.h5, test/python/base/*.ref.h5); see tests/test_h5lite.py.

The data model is a nested dict: group -> {"name": subgroup-dict | numpy array | str | scalar}, with attributes kept in a
parallel mapping ``attrs`` keyed by the object's path ("" = root, "G/data" ...).
"""
from __future__ import annotations

import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"


class H5LiteError(RuntimeError):
    pass


# ================================================================================================
# reader
# ================================================================================================
class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        if buf[:8] != SIG:
            raise H5LiteError("not an HDF5 file")
        ver = buf[8]
        if ver not in (0, 1):
            raise H5LiteError(f"superblock version {ver} is not supported (only the default 'earliest' format)")
        if buf[13] != 8 or buf[14] != 8:
            raise H5LiteError("only 8-byte offsets / lengths are supported")
        off = 24 + (4 if ver == 1 else 0)
        self.base, _free, self.eof, _drv = struct.unpack_from("<QQQQ", buf, off)
        self.root_entry = off + 32
        self.attrs: dict[str, dict] = {}

    # -- primitives ------------------------------------------------------------------------------
    def u(self, fmt, off):
        return struct.unpack_from("<" + fmt, self.b, off)

    def sym_entry(self, off):
        name_off, ohdr, cache, _ = self.u("QQII", off)
        scratch = self.b[off + 24:off + 40]
        return name_off, ohdr, cache, scratch

    def heap_string(self, heap_addr, name_off):
        if self.b[heap_addr:heap_addr + 4] != b"HEAP":
            raise H5LiteError("bad local heap signature")
        data_addr = self.u("Q", heap_addr + 24)[0]
        s = data_addr + name_off
        e = self.b.index(b"\0", s)
        return self.b[s:e].decode()

    # -- object headers ---------------------------------------------------------------------------
    def messages(self, addr):
        ver = self.b[addr]
        if ver != 1:
            if self.b[addr:addr + 4] == b"OHDR":
                raise H5LiteError("version 2 object headers (libver='latest') are not supported")
            raise H5LiteError(f"object header version {ver}")
        nmsg, = self.u("H", addr + 2)
        size, = self.u("I", addr + 8)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = self.u("HHB", pos)
                body = pos + 8
                if mtype == 0x0010:  # continuation
                    caddr, clen = self.u("QQ", body)
                    blocks.append((caddr, clen))
                out.append((mtype, body, msize))
                pos = body + msize
        return out

    # -- datatypes / dataspaces -------------------------------------------------------------------
    def datatype(self, off):
        """-> (kind, numpy dtype or None, size, consumed bytes).  kind in {"num", "str", "vstr"}"""
        cv = self.b[off]
        cls, ver = cv & 0x0F, cv >> 4
        bits0 = self.b[off + 1]
        size, = self.u("I", off + 4)
        if cls == 0:  # fixed point
            signed = bool(bits0 & 0x08)
            big = bool(bits0 & 0x01)
            dt = np.dtype(("i" if signed else "u") + str(size)).newbyteorder(">" if big else "<")
            return "num", dt, size, 8 + 4
        if cls == 1:  # floating point
            big = bool(bits0 & 0x01)
            dt = np.dtype("f" + str(size)).newbyteorder(">" if big else "<")
            return "num", dt, size, 8 + 12
        if cls == 3:  # fixed-length string
            return "str", None, size, 8
        if cls == 9:  # variable length
            if (bits0 & 0x0F) != 1:
                raise H5LiteError("variable-length sequences are not supported (only strings)")
            return "vstr", None, size, 8 + 16
        raise H5LiteError(f"datatype class {cls} (version {ver}) is not supported")

    def dataspace(self, off):
        ver, rank, flags = self.b[off], self.b[off + 1], self.b[off + 2]
        if ver == 1:
            p = off + 8
        elif ver == 2:
            if self.b[off + 3] == 2:  # null dataspace
                return None
            p = off + 4
        else:
            raise H5LiteError(f"dataspace version {ver}")
        return tuple(self.u("Q", p + 8 * i)[0] for i in range(rank))

    def global_heap_object(self, coll, index):
        if self.b[coll:coll + 4] != b"GCOL":
            raise H5LiteError("bad global heap signature")
        csize, = self.u("Q", coll + 8)
        pos = coll + 16
        while pos < coll + csize:
            idx, _rc, _r, osize = self.u("HHIQ", pos)
            if idx == 0:
                break
            if idx == index:
                return self.b[pos + 16:pos + 16 + osize]
            pos += 16 + ((osize + 7) & ~7)
        raise H5LiteError("global heap object not found")

    def decode(self, kind, dt, size, shape, raw):
        n = 1
        for d in (shape or ()):
            n *= d
        if kind == "num":
            a = np.frombuffer(raw, dtype=dt, count=n).astype(dt.newbyteorder("="))
            return a.reshape(shape) if shape else a.reshape(())[()]
        if kind == "str":
            vals = [raw[i * size:(i + 1) * size].split(b"\0")[0].decode() for i in range(n)]
        else:
            vals = []
            for i in range(n):
                ln, coll, idx = struct.unpack_from("<IQI", raw, 16 * i)
                vals.append(self.global_heap_object(coll, idx)[:ln].decode() if ln else "")
        if not shape:
            return vals[0]
        return np.array(vals, dtype=object).reshape(shape)

    # -- objects ------------------------------------------------------------------------------------
    def read_object(self, ohdr, path):
        msgs = self.messages(ohdr)
        attrs = {}
        space = dtype = layout = symtab = None
        filters = []
        for mtype, body, msize in msgs:
            if mtype == 0x0001:
                space = self.dataspace(body)
            elif mtype == 0x0003:
                dtype = self.datatype(body)
            elif mtype == 0x0008:
                layout = body
            elif mtype == 0x000B:
                filters = self.filter_pipeline(body)
            elif mtype == 0x0011:
                symtab = self.u("QQ", body)
            elif mtype == 0x000C:
                name, val = self.attribute(body)
                attrs[name] = val
            elif mtype == 0x0002:
                raise H5LiteError("new-style groups (link info messages) are not supported")
        if attrs:
            self.attrs[path] = attrs
        if symtab is not None:
            return self.read_group(symtab[0], symtab[1], path)
        if dtype is None or layout is None:
            raise H5LiteError(f"object '{path}' is neither an old-style group nor a dataset")
        kind, dt, size, _ = dtype
        lver, lcls = self.b[layout], self.b[layout + 1]
        if lver != 3:
            raise H5LiteError(f"data layout message version {lver}")
        n = 1
        for d in (space or ()):
            n *= d
        esize = size if kind != "vstr" else 16
        if lcls == 1:
            addr, nbytes = self.u("QQ", layout + 2)
            raw = b"" if addr == UNDEF else self.b[addr:addr + n * esize]
            if addr == UNDEF and n:
                raw = bytes(n * esize)
        elif lcls == 0:
            nbytes, = self.u("H", layout + 2)
            raw = self.b[layout + 4:layout + 4 + nbytes]
        elif lcls == 2:
            raw = self.read_chunked(layout, space, esize, filters, path)
        else:
            raise H5LiteError(f"dataset '{path}': data layout class {lcls}")
        return self.decode(kind, dt, size, space, raw)

    # -- chunked datasets (h5py / the TRIQS Python archive write compressed arrays this way) ------------
    def filter_pipeline(self, body):
        ver, nf = self.b[body], self.b[body + 1]
        if ver != 1:
            raise H5LiteError(f"filter pipeline message version {ver}")
        p = body + 8
        out = []
        for _ in range(nf):
            fid, nlen, _flags, ncd = self.u("HHHH", p)
            p += 8 + ((nlen + 7) & ~7)
            cd = [self.u("I", p + 4 * i)[0] for i in range(ncd)]
            p += 4 * ncd + (4 if ncd % 2 else 0)
            out.append((fid, cd))
        return out

    def chunk_nodes(self, addr, ndim):
        if self.b[addr:addr + 4] != b"TREE" or self.b[addr + 4] != 1:
            raise H5LiteError("bad chunk B-tree node")
        level, used = self.b[addr + 5], self.u("H", addr + 6)[0]
        ksz = 8 + 8 * ndim
        p = addr + 24
        for _ in range(used):
            csize, fmask = self.u("II", p)
            offs = [self.u("Q", p + 8 + 8 * i)[0] for i in range(ndim)]
            child, = self.u("Q", p + ksz)
            p += ksz + 8
            if level > 0:
                yield from self.chunk_nodes(child, ndim)
            else:
                yield csize, fmask, offs, child

    def read_chunked(self, layout, shape, esize, filters, path):
        import zlib
        ndim = self.b[layout + 2]
        btree, = self.u("Q", layout + 3)
        cdims = [self.u("I", layout + 11 + 4 * i)[0] for i in range(ndim)]
        rank = ndim - 1
        if shape is None or len(shape) != rank:
            raise H5LiteError(f"dataset '{path}': chunk rank mismatch")
        out = np.zeros(tuple(shape) + (esize,), dtype=np.uint8)
        if btree == UNDEF:
            return out.tobytes()
        cshape = tuple(cdims[:rank])
        for csize, fmask, offs, addr in self.chunk_nodes(btree, ndim):
            raw = self.b[addr:addr + csize]
            for i, (fid, cd) in reversed(list(enumerate(filters))):
                if fmask & (1 << i):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:  # shuffle: bytes of the elements were transposed
                    n = len(raw) // esize
                    raw = np.frombuffer(raw[:n * esize], dtype=np.uint8).reshape(esize, n).T.tobytes() + raw[n * esize:]
                elif fid == 3:  # fletcher32 checksum appended
                    raw = raw[:-4]
                else:
                    raise H5LiteError(f"dataset '{path}': filter {fid} is not supported")
            chunk = np.frombuffer(raw, dtype=np.uint8, count=int(np.prod(cshape)) * esize).reshape(cshape + (esize,))
            sl_out, sl_in = [], []
            for d in range(rank):
                lo = offs[d]
                hi = min(lo + cshape[d], shape[d])
                sl_out.append(slice(lo, hi))
                sl_in.append(slice(0, hi - lo))
            out[tuple(sl_out)] = chunk[tuple(sl_in)]
        return out.tobytes()

    def attribute(self, body):
        ver = self.b[body]
        if ver not in (1, 2, 3):
            raise H5LiteError(f"attribute message version {ver}")
        nsz, tsz, ssz = self.u("HHH", body + 2)
        pad = (lambda x: (x + 7) & ~7) if ver == 1 else (lambda x: x)
        p = body + 8 + (1 if ver == 3 else 0)
        name = self.b[p:p + nsz].split(b"\0")[0].decode()
        p += pad(nsz)
        kind, dt, size, _ = self.datatype(p)
        p += pad(tsz)
        shape = self.dataspace(p)
        p += pad(ssz)
        n = 1
        for d in (shape or ()):
            n *= d
        raw = self.b[p:p + n * (size if kind != "vstr" else 16)]
        return name, self.decode(kind, dt, size, shape, raw)

    def read_group(self, btree, heap, path):
        out = {}
        for name, ohdr in self.btree_entries(btree, heap):
            out[name] = self.read_object(ohdr, (path + "/" + name) if path else name)
        return out

    def btree_entries(self, addr, heap):
        if self.b[addr:addr + 4] != b"TREE":
            raise H5LiteError("bad B-tree signature")
        ntype, level, used = self.b[addr + 4], self.b[addr + 5], self.u("H", addr + 6)[0]
        if ntype != 0:
            raise H5LiteError("unexpected B-tree node type")
        p = addr + 24
        for i in range(used):
            child, = self.u("Q", p + 8)  # key_i (8), child_i (8)
            p += 16
            if level > 0:
                yield from self.btree_entries(child, heap)
            else:
                if self.b[child:child + 4] != b"SNOD":
                    raise H5LiteError("bad symbol node signature")
                nsym, = self.u("H", child + 6)
                for j in range(nsym):
                    name_off, ohdr, _cache, _sc = self.sym_entry(child + 8 + 40 * j)
                    yield self.heap_string(heap, name_off), ohdr

    def read(self):
        _name_off, ohdr, cache, scratch = self.sym_entry(self.root_entry)
        tree = self.read_object(ohdr, "")
        return tree, self.attrs


def read(path):
    """-> (tree, attrs): nested dict of groups / numpy arrays / strings, and {object path: {attribute: value}}."""
    with open(path, "rb") as f:
        return _Reader(f.read()).read()


# ================================================================================================
# writer
# ================================================================================================
def _pad8(b: bytes) -> bytes:
    return b + bytes((-len(b)) % 8)


def _dtype_msg(kind, dt=None, size=0):
    if kind == "str":
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, size)  # class 3, version 1; null-terminated ASCII
    if dt.kind == "f":
        if dt.itemsize == 8:
            return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 0x3F, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 0x1F, 0, 4, 0, 32, 23, 8, 0, 23, 127)
    if dt.kind in "iu":
        return struct.pack("<BBBBIHH", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    raise H5LiteError(f"cannot store dtype {dt}")


def _space_msg(shape):
    if shape is None:
        shape = ()
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", d) for d in shape)


def _value_parts(v):
    """python / numpy value -> (datatype message, dataspace message, raw bytes)"""
    if isinstance(v, str):
        raw = v.encode() + b"\0"
        return _dtype_msg("str", size=len(raw)), _space_msg(()), raw
    if isinstance(v, (list, tuple)) and v and all(isinstance(x, str) for x in v):
        size = max(len(x.encode()) for x in v) + 1
        raw = b"".join(x.encode().ljust(size, b"\0") for x in v)
        return _dtype_msg("str", size=size), _space_msg((len(v),)), raw
    a = np.asarray(v)
    if a.dtype == np.bool_:
        a = a.astype(np.int8)
    if a.dtype.kind == "c":
        raise H5LiteError("complex values are stored as [..., 2] real arrays with a __complex__ attribute (use write_complex)")
    if a.dtype.kind == "O" or a.dtype.kind == "U":
        return _value_parts([str(x) for x in a.reshape(-1)])
    shape = a.shape  # (ascontiguousarray would turn a 0-d scalar into shape (1,))
    a = np.ascontiguousarray(a.astype(a.dtype.newbyteorder("<")))
    return _dtype_msg("num", a.dtype), _space_msg(shape), a.tobytes()


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)  # superblock (24 + 32 + 40 bytes), filled in at the end

    def alloc(self, data: bytes) -> int:
        addr = len(self.buf)
        self.buf += _pad8(data)
        return addr

    @staticmethod
    def msg(mtype, body, flags=0):
        body = _pad8(body)
        return struct.pack("<HHB3x", mtype, len(body), flags) + body

    def object_header(self, msgs):
        body = b"".join(msgs)
        hdr = struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body))
        return self.alloc(hdr + body)

    def attr_msgs(self, attrs):
        out = []
        for name, val in (attrs or {}).items():
            dtm, spm, raw = _value_parts(val)
            nm = name.encode() + b"\0"
            body = struct.pack("<BBHHH", 1, 0, len(nm), len(dtm), len(spm)) + _pad8(nm) + _pad8(dtm) + _pad8(spm) + raw
            out.append(self.msg(0x000C, body))
        return out

    def dataset(self, value, attrs):
        dtm, spm, raw = _value_parts(value)
        addr = self.alloc(raw) if raw else UNDEF
        layout = struct.pack("<BBQQ", 3, 1, addr, len(raw))
        fill = struct.pack("<BBBB", 2, 2, 0, 0)  # fill value message v2: late allocation, never written, undefined
        msgs = [self.msg(0x0001, spm), self.msg(0x0003, dtm, flags=1), self.msg(0x0005, fill), self.msg(0x0008, layout)] + self.attr_msgs(attrs)
        return self.object_header(msgs)

    def group(self, tree, attrs_all, path):
        entries = []
        for name in sorted(tree):  # B-tree keys must be in increasing (byte-wise) order
            sub = (path + "/" + name) if path else name
            v = tree[name]
            if isinstance(v, dict):
                entries.append((name, self.group(v, attrs_all, sub)))
            else:
                entries.append((name, self.dataset(v, attrs_all.get(sub))))
        # local heap: names, 8-byte aligned; offset 0 is the empty string (key 0 of the B-tree)
        heap_data = bytearray(8)
        offs = []
        for name, _ in entries:
            offs.append(len(heap_data))
            heap_data += _pad8(name.encode() + b"\0")
        free_off = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 16)  # one free block: next = 1 (none), size 16
        data_addr = self.alloc(bytes(heap_data))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, data_addr))
        # symbol nodes: at most 2K = 8 symbols each (group leaf node K = 4), one level-0 B-tree node (<= 2K = 32 children, K = 16)
        per = 8
        nodes = [entries[i:i + per] for i in range(0, len(entries), per)] or [[]]
        if len(nodes) > 32:
            raise H5LiteError("too many links in one group for the single-node B-tree of this writer (max 256)")
        keys, children = [0], []
        pos = 0
        for chunk in nodes:
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk))
            for (name, ohdr) in chunk:
                body += struct.pack("<QQII16x", offs[pos], ohdr, 0, 0)
                pos += 1
            body += bytes(40 * (per - len(chunk)))
            children.append(self.alloc(body))
            keys.append(offs[pos - 1] if chunk else 0)
        tb = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(children) if entries else 0, UNDEF, UNDEF)
        for i, c in enumerate(children):
            tb += struct.pack("<QQ", keys[i], c)
        tb += struct.pack("<Q", keys[-1])
        tb += bytes((2 * 16 + 1) * 8 + 2 * 16 * 8 - (len(tb) - 24))  # full-size node: 2K+1 keys, 2K children
        btree_addr = self.alloc(tb)
        msgs = [self.msg(0x0011, struct.pack("<QQ", btree_addr, heap_addr))] + self.attr_msgs(attrs_all.get(path))
        ohdr = self.object_header(msgs)
        self.last_group = (btree_addr, heap_addr)
        return ohdr

    def finish(self, root_ohdr, root_bt_heap):
        sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, root_ohdr, 1, 0) + struct.pack("<QQ", *root_bt_heap)
        self.buf[:len(sb)] = sb
        return bytes(self.buf)


def write(path, tree: dict, attrs: dict | None = None):
    """tree: nested dict {name: dict | array | str | scalar}; attrs: {object path: {attribute: value}} ("" = root)."""
    w = _Writer()
    root = w.group(tree, attrs or {}, "")
    data = w.finish(root, w.last_group)
    with open(path, "wb") as f:
        f.write(data)


# ================================================================================================
# the TRIQS conventions on top: complex arrays, Gf / BlockGf groups
# ================================================================================================
def put_complex(tree, attrs, path, name, z):
    """complex array -> real [..., 2] dataset + __complex__ attribute (the convention of the TRIQS h5 library)"""
    z = np.asarray(z, dtype=np.complex128)
    tree[name] = np.ascontiguousarray(z).view(np.float64).reshape(z.shape + (2,))
    attrs[(path + "/" + name) if path else name] = {"__complex__": "1"}


def get_array(tree, attrs, path, name):
    a = tree[name]
    at = attrs.get((path + "/" + name) if path else name, {})
    if "__complex__" in at and isinstance(a, np.ndarray) and a.shape[-1] == 2:
        return np.ascontiguousarray(a, dtype=np.float64).view(np.complex128).reshape(a.shape[:-1])
    return a


# ------------------------------------------------------------------------------------------------
# Gf / BlockGf groups: c++/triqs/gfs/h5.hpp:72-91 (data + mesh), _gf_view_common.hpp:152-171 (Format tag "Gf"),
# block_gf.hpp h5 (Format "BlockGf", block_names list, one sub-group per block)
# ------------------------------------------------------------------------------------------------
def _mesh_to_group(m):
    """-> (group dict, Format string): the members each reference mesh writes"""
    from . import gf as G
    stat = lambda s: "F" if s == G.FERMION else "B"
    if isinstance(m, G.MeshImFreq):  # mesh/imfreq.hpp:261-272
        return {"beta": float(m.beta), "statistic": stat(m.statistic), "size": np.int64(len(m)), "positive_freq_only": np.int32(0)}, "MeshImFreq"
    if isinstance(m, G.MeshImTime):  # mesh/imtime.hpp:83-92
        return {"beta": float(m.beta), "statistic": stat(m.statistic), "n_tau": np.int64(m.n_tau)}, "MeshImTime"
    if isinstance(m, (G.MeshReFreq, G.MeshReTime)):  # mesh/bases/linear.hpp:198-204
        return ({"min": float(m.xmin), "max": float(m.xmax), "size": np.int64(m.n)},
                "MeshReFreq" if isinstance(m, G.MeshReFreq) else "MeshReTime")
    if isinstance(m, G.MeshLegendre):  # mesh/legendre.hpp:145-152
        return {"beta": float(m.beta), "statistic": stat(m.statistic), "max_n": np.int64(m.n_l)}, "MeshLegendre"
    if isinstance(m, (G.MeshBrZone, G.MeshCycLat)):  # mesh/brzone.hpp:315-321, cyclat.hpp:215-221 (lattice geometry is not carried here)
        return {"dims": np.asarray(m.dims, dtype=np.int64)}, "MeshBrillouinZone" if isinstance(m, G.MeshBrZone) else "MeshCyclicLattice"
    if isinstance(m, G.MeshProduct):  # mesh/prod.hpp:182-187
        return {f"MeshComponent{i}": _mesh_to_group(c) for i, c in enumerate(m.meshes)}, "MeshProduct"
    raise H5LiteError(f"no HDF5 layout for mesh {type(m).__name__}")


class _ComponentHolder:
    def __init__(self, pair):
        self.pair = pair


def _add_mesh(tree, attrs, path, name, m):
    grp, fmt = m.pair if isinstance(m, _ComponentHolder) else _mesh_to_group(m)
    sub = (path + "/" + name) if path else name
    attrs[sub] = {"Format": fmt}
    out = {}
    for k, v in grp.items():
        if isinstance(v, tuple):
            _add_mesh(out, attrs, sub, k, _ComponentHolder(v))
        else:
            out[k] = v
    tree[name] = out


def _mesh_from_group(grp, fmt):
    from . import gf as G
    stat = lambda s: G.FERMION if s == "F" else G.BOSON
    dom = grp.get("domain", grp)  # backward compatibility of the reference readers (imfreq.hpp:287, imtime.hpp:101)
    if fmt == "MeshImFreq":
        pos = int(grp.get("positive_freq_only", grp.get("start_at_0", 0)))
        size = int(grp["size"])
        if pos:
            raise H5LiteError("positive-only imfreq meshes are not used on this path")
        return G.MeshImFreq(float(dom["beta"]), stat(dom["statistic"]), (size + 1) // 2)
    if fmt == "MeshImTime":
        return G.MeshImTime(float(dom["beta"]), stat(dom["statistic"]), int(grp["n_tau"] if "n_tau" in grp else grp["size"]))
    if fmt in ("MeshReFreq", "MeshReTime"):
        return (G.MeshReFreq if fmt == "MeshReFreq" else G.MeshReTime)(float(grp["min"]), float(grp["max"]), int(grp["size"]))
    if fmt == "MeshLegendre":
        return G.MeshLegendre(float(dom["beta"]), stat(dom["statistic"]), int(grp["max_n"] if "max_n" in grp else grp["n_max"]))
    if fmt in ("MeshBrillouinZone", "MeshCyclicLattice"):
        dims = tuple(int(x) for x in np.asarray(grp["dims"]).reshape(-1))
        return G.MeshBrZone(dims) if fmt == "MeshBrillouinZone" else G.MeshCycLat(dims)
    raise H5LiteError(f"mesh format '{fmt}' is not supported")


def _read_mesh(grp, attrs, path):
    fmt = attrs.get(path, {}).get("Format")
    if fmt == "MeshProduct":
        from . import gf as G
        comps = []
        i = 0
        while f"MeshComponent{i}" in grp:
            comps.append(_read_mesh(grp[f"MeshComponent{i}"], attrs, f"{path}/MeshComponent{i}"))
            i += 1
        return G.MeshProduct(tuple(comps))
    if fmt is None:  # archives written before the Format attribute: guess from the members
        fmt = "MeshImFreq" if "positive_freq_only" in grp or "domain" in grp else ("MeshReFreq" if "min" in grp else None)
    return _mesh_from_group(grp, fmt)


def _gf_into(tree, attrs, path, name, g):
    from . import gf as G
    sub = (path + "/" + name) if path else name
    if isinstance(g, G.BlockGf):
        attrs[sub] = {"Format": "BlockGf"}
        grp = {}
        attrs[sub + "/block_names"] = {"Format": "List"}
        grp["block_names"] = {str(i): str(n) for i, n in enumerate(g.names)}
        for n, b in zip(g.names, g.blocks):
            _gf_into(grp, attrs, sub, str(n), b)
        tree[name] = grp
        return
    attrs[sub] = {"Format": "Gf"}
    grp = {}
    data = g.data.cpu().numpy() if type(g.data).__module__.startswith("torch") else np.asarray(g.data)
    put_complex(grp, attrs, sub, "data", data)
    _add_mesh(grp, attrs, sub, "mesh", g.mesh)
    tree[name] = grp


def _gf_from(grp, attrs, path, ctx=None):
    from . import gf as G
    fmt = attrs.get(path, {}).get("Format", attrs.get(path, {}).get("TRIQS_HDF5_data_scheme"))
    if fmt == "BlockGf":
        names = [grp["block_names"][str(i)] for i in range(len(grp["block_names"]))]
        return G.BlockGf(names, [_gf_from(grp[n], attrs, f"{path}/{n}", ctx) for n in names])
    mesh = _read_mesh(grp["mesh"], attrs, path + "/mesh")
    data = get_array(grp, attrs, path, "data")
    n_mesh = len(mesh.shape) if isinstance(mesh, G.MeshProduct) else 1
    return G.Gf(mesh, tuple(data.shape[n_mesh:]), np.ascontiguousarray(data, dtype=np.complex128), ctx)


def write_gf(path, objects: dict):
    """objects: {name: Gf | BlockGf}  ->  one HDF5 file a TRIQS `h5::file` / `HDFArchive` can open."""
    tree, attrs = {}, {}
    for name, g in objects.items():
        _gf_into(tree, attrs, "", name, g)
    write(path, tree, attrs)


def read_gf(path, name=None, ctx=None):
    """-> the Gf / BlockGf stored under `name` (or a dict of everything that carries a Gf / BlockGf Format tag)."""
    tree, attrs = read(path)
    if name is not None:
        return _gf_from(tree[name], attrs, name, ctx)
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict) and attrs.get(k, {}).get("Format", attrs.get(k, {}).get("TRIQS_HDF5_data_scheme", "")) in ("Gf", "BlockGf", "GfImFreq", "GfReFreq"):
            out[k] = _gf_from(v, attrs, k, ctx)
    return out
