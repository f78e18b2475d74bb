"""ORACLE / TEST INFRASTRUCTURE -- minimal reader for MATLAB v7.3 (.mat = HDF5) files, enough for the
reference's `model_approx.mat` (A_s 2883 x 28, b_s 2883 x 1: the first-order PSF model of the estimator,
README.md:294, :478).  h5py is not in the image, so the few HDF5 structures the file uses are parsed by hand:

  512-byte MATLAB user block -> superblock v0 -> root group symbol table (B-tree v1 'TREE' of group nodes 'SNOD',
  names in a local heap 'HEAP') -> v1 object headers -> dataspace (0x01), datatype (0x03, IEEE f64 LE only),
  data layout v3 (0x08: contiguous or chunked), filter pipeline (0x0B: deflate only) -> chunk B-tree v1 -> zlib.

All file addresses are relative to the user block.  HDF5 stores dimensions slowest-first, MATLAB is column-major:
an HDF5 dataset of shape (28, 2883) is the MATLAB array 2883 x 28, i.e. the transpose of the C-order read.
Only product code under tests/, bench.py's baseline leg and __graft_entry__.smoke() may import this module.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class Mat73:
    def __init__(self, path: str):
        self.d = open(path, "rb").read()
        self.base = self.d.find(b"\x89HDF\r\n\x1a\n")
        if self.base < 0:
            raise ValueError("not an HDF5 (MATLAB v7.3) file")
        sb = self.base
        ver = self.d[sb + 8]
        if ver != 0:
            raise ValueError(f"superblock version {ver} not supported")
        so, sl = self.d[sb + 13], self.d[sb + 14]
        if (so, sl) != (8, 8):
            raise ValueError("only 8-byte offsets / lengths are supported")
        # v0: sig8 ver1 fsver1 rgver1 res1 shver1 so1 sl1 res1 leafk2 intk2 flags4 | base8 freesp8 eof8 drv8 | root symbol-table entry
        root = sb + 24 + 32
        _, self.root_ohdr, cache_type = struct.unpack_from("<QQI", self.d, root)
        scratch = root + 24
        if cache_type != 1:
            raise ValueError("root group without cached symbol-table info")
        self.root_btree, self.root_heap = struct.unpack_from("<QQ", self.d, scratch)

    # ---- low level ----
    def at(self, addr: int) -> int:
        return self.base + addr

    def heap_data(self, heap_addr: int) -> int:
        o = self.at(heap_addr)
        assert self.d[o:o + 4] == b"HEAP"
        _, _, data_addr = struct.unpack_from("<QQQ", self.d, o + 8)
        return self.at(data_addr)

    def group_entries(self, btree_addr: int, heap_addr: int) -> dict:
        names = {}
        hd = self.heap_data(heap_addr)

        def walk(addr):
            o = self.at(addr)
            sig = self.d[o:o + 4]
            if sig == b"TREE":
                ntype, level, used = struct.unpack_from("<BBH", self.d, o + 4)
                p = o + 24                                  # sig4 type1 level1 used2 left8 right8
                for i in range(used):
                    child = struct.unpack_from("<Q", self.d, p + 8 + i * 16)[0]     # key(8) child(8) key child ...
                    walk(child)
            elif sig == b"SNOD":
                nsym = struct.unpack_from("<H", self.d, o + 6)[0]
                p = o + 8
                for i in range(nsym):
                    name_off, ohdr = struct.unpack_from("<QQ", self.d, p + i * 40)
                    e = self.d.index(b"\x00", hd + name_off)
                    names[self.d[hd + name_off:e].decode()] = ohdr
            else:
                raise ValueError(f"unexpected node signature {sig!r}")

        walk(btree_addr)
        return names

    def messages(self, ohdr_addr: int):
        """Yields (type, body bytes) of a version-1 object header, following continuation blocks."""
        o = self.at(ohdr_addr)
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", self.d, o)
        if ver != 1:
            raise ValueError(f"object header version {ver} not supported")
        blocks = [(o + 16, hsize)]
        out = []
        while blocks:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", self.d, p)
                body = self.d[p + 8:p + 8 + msize]
                if mtype == 0x10:                           # continuation
                    caddr, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((self.at(caddr), clen))
                out.append((mtype, body))
                p += 8 + msize
        return out

    # ---- datasets ----
    def read(self, name: str) -> np.ndarray:
        """Dataset `name` of the root group as a MATLAB-oriented float64 array."""
        ents = self.group_entries(self.root_btree, self.root_heap)
        if name not in ents:
            raise KeyError(f"{name}: not in {sorted(ents)}")
        dims = layout = None
        deflate = False
        for mtype, body in self.messages(ents[name]):
            if mtype == 0x01:                               # dataspace v1: ver rank flags res(5) dims...
                ver, rank, flags = body[0], body[1], body[2]
                off = 8 if ver == 1 else 4
                dims = struct.unpack_from(f"<{rank}Q", body, off)
            elif mtype == 0x03:                             # datatype: class+version byte, 3 flag bytes, size
                cls = body[0] & 0x0F
                size = struct.unpack_from("<I", body, 4)[0]
                if cls != 1 or size != 8 or (body[1] & 1):
                    raise ValueError("only little-endian IEEE float64 datasets are supported")
            elif mtype == 0x08:                             # data layout v3
                if body[0] != 3:
                    raise ValueError(f"layout version {body[0]} not supported")
                lclass = body[1]
                if lclass == 1:
                    addr, size = struct.unpack_from("<QQ", body, 2)
                    layout = ("contiguous", addr, size)
                elif lclass == 2:
                    rank1 = body[2]
                    addr = struct.unpack_from("<Q", body, 3)[0]
                    cdims = struct.unpack_from(f"<{rank1}I", body, 11)
                    layout = ("chunked", addr, cdims[:-1])
                else:
                    raise ValueError("compact layout not supported")
            elif mtype == 0x0B:                             # filter pipeline v1: ver nfilters res(6) | id namelen flags nvals ...
                nf = body[1]
                p = 8
                for _ in range(nf):
                    fid, nlen, _fl, nvals = struct.unpack_from("<HHHH", body, p)
                    p += 8 + ((nlen + 7) // 8) * 8 + 4 * nvals + (4 if nvals % 2 else 0)
                    if fid == 1:
                        deflate = True
                    else:
                        raise ValueError(f"filter {fid} not supported")
        if dims is None or layout is None:
            raise ValueError("dataset without dataspace / layout message")
        out = np.zeros(dims, dtype="<f8")
        if layout[0] == "contiguous":
            o = self.at(layout[1])
            out[...] = np.frombuffer(self.d, dtype="<f8", count=int(np.prod(dims)), offset=o).reshape(dims)
        else:
            cdims = layout[2]
            rank = len(dims)

            def walk(addr):
                o = self.at(addr)
                assert self.d[o:o + 4] == b"TREE"
                ntype, level, used = struct.unpack_from("<BBH", self.d, o + 4)
                assert ntype == 1
                ksz = 8 + 8 * (rank + 1)                    # chunk size 4, filter mask 4, offsets 8 (rank + 1)
                p = o + 24
                for i in range(used):
                    csize, _mask = struct.unpack_from("<II", self.d, p)
                    offs = struct.unpack_from(f"<{rank}Q", self.d, p + 8)
                    child = struct.unpack_from("<Q", self.d, p + ksz)[0]
                    if level > 0:
                        walk(child)
                    else:
                        raw = self.d[self.at(child):self.at(child) + csize]
                        if deflate:
                            raw = zlib.decompress(raw)
                        blk = np.frombuffer(raw, dtype="<f8").reshape(cdims)
                        sl = tuple(slice(offs[k], min(offs[k] + cdims[k], dims[k])) for k in range(rank))
                        out[sl] = blk[tuple(slice(0, s.stop - s.start) for s in sl)]
                    p += ksz + 8

            walk(layout[1])
        return np.ascontiguousarray(out.T)                  # MATLAB orientation


def load_model_approx(path: str):
    """(A_s (2883, 28), b_s (2883,)) of the reference's model_approx.mat."""
    f = Mat73(path)
    return f.read("A_s"), f.read("b_s").reshape(-1)
