"""Dependency-free readers for the reference's data fixtures (host-side data preparation only).

* ``read_hdf5_datasets``  - minimal HDF5 reader (superblock v0, symbol-table groups, v1 object
  headers, contiguous little-endian datasets), enough for the statismo/Scalismo model files the
  reference loads through ``StatisticalModelIO.readStatisticalMeshModel``
  (reference: src/main/scala/apps/femur/LoadTestData.scala:34-35).
* ``read_binary_stl``      - binary STL with vertex merging in first-appearance order
  (reference: ``MeshIO.readMesh``, src/main/scala/apps/femur/LoadTestData.scala:39-40).
* ``load_gpmm_h5``         - model file -> (ref points, cells, mean deformation, basis U, variance).

This module is not on the hot path; it only produces the arrays that cross the C ABI.
"""
from __future__ import annotations

import struct
import json
import numpy as np

_UNDEF = 0xFFFFFFFFFFFFFFFF


class _H5:
    def __init__(self, buf: bytes):
        self.b = buf
        if buf[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError("not an HDF5 file")
        ver = buf[8]
        if ver != 0:
            raise ValueError(f"only superblock v0 supported (got {ver})")
        so, sl = buf[13], buf[14]
        if so != 8 or sl != 8:
            raise ValueError("only 8-byte offsets/lengths supported")
        # v0: sig8 ver4x1 so sl rsv leafK(2) intK(2) flags(4) base free eof driver | root entry
        self.base = struct.unpack_from("<Q", buf, 24)[0]
        self.root_entry = 24 + 32

    def u(self, fmt, off):
        return struct.unpack_from("<" + fmt, self.b, off)

    # ---- groups -----------------------------------------------------------------------
    def _heap_name(self, heap_addr, off):
        assert self.b[heap_addr:heap_addr + 4] == b"HEAP"
        data_addr = self.u("Q", heap_addr + 24)[0]
        s = data_addr + off
        e = self.b.index(b"\x00", s)
        return self.b[s:e].decode()

    def _iter_btree(self, addr, heap_addr):
        sig = self.b[addr:addr + 4]
        if sig == b"TREE":
            ntype, level, used = self.u("BBH", addr + 4)
            p = addr + 24
            for i in range(used):
                child = self.u("Q", p + 8 + i * 16)[0]
                yield from self._iter_btree(child, heap_addr)
        elif sig == b"SNOD":
            n = self.u("H", addr + 6)[0]
            p = addr + 8
            for i in range(n):
                name_off, ohdr = self.u("QQ", p + i * 40)
                yield self._heap_name(heap_addr, name_off), ohdr
        else:
            raise ValueError(f"unexpected node signature {sig!r} at {addr}")

    def _messages(self, ohdr):
        ver, _, nmsg, _, hsize = self.u("BBHII", ohdr)
        if ver != 1:
            raise ValueError("only v1 object headers supported")
        blocks = [(ohdr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            p, sz = blocks.pop(0)
            end = p + sz
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, mflags = self.u("HHB", p)
                body = p + 8
                if mtype == 0x10:
                    coff, clen = self.u("QQ", body)
                    blocks.append((coff, clen))
                out.append((mtype, body, msize))
                p = body + msize
        return out

    def _walk(self, ohdr, prefix, found):
        msgs = self._messages(ohdr)
        sym = [m for m in msgs if m[0] == 0x11]
        if sym:
            btree, heap = self.u("QQ", sym[0][1])
            for name, child in self._iter_btree(btree, heap):
                self._walk(child, prefix + "/" + name, found)
            return
        shape = dtype = addr = size = None
        for mtype, body, msize in msgs:
            if mtype == 0x1:
                v, rank, flags = self.u("BBB", body)
                doff = body + (8 if v == 1 else 4)
                shape = tuple(self.u("Q" * rank, doff)) if rank else ()
            elif mtype == 0x3:
                cv = self.b[body]
                cls = cv & 0x0F
                bits0 = self.b[body + 1]
                tsize = self.u("I", body + 4)[0]
                if bits0 & 1:
                    raise ValueError("big-endian datasets unsupported")
                if cls == 0:
                    signed = (bits0 >> 3) & 1
                    dtype = np.dtype(("<i" if signed else "<u") + str(tsize))
                elif cls == 1:
                    dtype = np.dtype("<f" + str(tsize))
                else:
                    dtype = None  # strings etc.: ignored
            elif mtype == 0x8:
                v = self.b[body]
                if v == 3:
                    lclass = self.b[body + 1]
                    if lclass == 1:
                        addr, size = self.u("QQ", body + 2)
                    elif lclass == 0:  # compact
                        csize = self.u("H", body + 2)[0]
                        addr, size = body + 4 - self.base, csize
                else:
                    rank = self.b[body + 1]
                    lclass = self.b[body + 2]
                    if lclass == 1:
                        addr = self.u("Q", body + 8)[0]
        if shape is not None and dtype is not None and addr is not None and addr != _UNDEF:
            n = int(np.prod(shape)) if shape else 1
            arr = np.frombuffer(self.b, dtype=dtype, count=n, offset=self.base + addr).reshape(shape)
            found[prefix] = arr

    def datasets(self):
        found = {}
        ohdr = self.u("Q", self.root_entry + 8)[0]
        self._walk(ohdr, "", found)
        return found


def read_hdf5_datasets(path):
    with open(path, "rb") as f:
        return _H5(f.read()).datasets()


def load_gpmm_h5(path):
    """-> dict(ref (N,3) f64, cells (T,3) i32, mean_def (3N,) f64, basis (3N,K) f64, variance (K,) f64).

    Layout follows statismo format 0.9 as written by Scalismo 0.90:
    ``/model/mean`` holds mean *positions* (interleaved xyz); the mean deformation is
    ``mean - representer points``. ``pcaBasis`` is used as stored (no sqrt(lambda) rescale).
    """
    d = read_hdf5_datasets(path)
    pts = np.asarray(d["/representer/points"], dtype=np.float64)  # (3, N)
    cells = np.asarray(d["/representer/cells"], dtype=np.int32)   # (3, T)
    ref = np.ascontiguousarray(pts.T)
    tris = np.ascontiguousarray(cells.T)
    mean = np.asarray(d["/model/mean"], dtype=np.float64).reshape(-1)
    basis = np.asarray(d["/model/pcaBasis"], dtype=np.float64)
    var = np.asarray(d["/model/pcaVariance"], dtype=np.float64).reshape(-1)
    if basis.shape[0] != mean.shape[0]:
        basis = basis.T
    return dict(ref=ref, cells=tris, mean_def=mean - ref.reshape(-1), basis=np.ascontiguousarray(basis),
                variance=var)


def read_binary_stl(path):
    """-> (vertices (N,3) f64 merged in first-appearance order, cells (T,3) i32)."""
    with open(path, "rb") as f:
        raw = f.read()
    ntri = struct.unpack_from("<I", raw, 80)[0]
    rec = np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")])
    facets = np.frombuffer(raw, dtype=rec, count=ntri, offset=84)
    v = facets["v"].reshape(-1, 3)
    key = v.view(np.dtype((np.void, 12))).reshape(-1)
    _, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # unique keys in first-appearance order
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    verts = v[first[order]].astype(np.float64)
    cells = rank[inv].reshape(-1, 3).astype(np.int32)
    return verts, cells


def read_landmarks_json(path):
    with open(path) as f:
        lms = json.load(f)
    return {lm["id"]: np.asarray(lm["coordinates"], dtype=np.float64) for lm in lms}


def rigid_landmark_alignment(src, dst):
    """Least-squares rigid transform (Umeyama/Kabsch, no scale) mapping src landmarks onto dst.

    Mirrors ``LandmarkRegistration.rigid3DLandmarkRegistration`` as called from
    src/main/scala/apps/util/AlignmentTransforms.scala:25-30. Returns (R, t) with y = R x + t.
    """
    names = [n for n in src if n in dst]
    a = np.stack([src[n] for n in names])
    b = np.stack([dst[n] for n in names])
    ca, cb = a.mean(0), b.mean(0)
    h = (a - ca).T @ (b - cb)
    u, _, vt = np.linalg.svd(h)
    d = np.sign(np.linalg.det(vt.T @ u.T))
    r = vt.T @ np.diag([1.0, 1.0, d]) @ u.T
    return r, cb - r @ ca
