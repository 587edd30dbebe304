"""tests/h5mini.py — an independent reader for the HDF5 subset picsp's output uses, written
straight from the HDF5 File Format Specification (superblock v0, v1 object headers, old-style
groups: symbol-table message -> v1 B-tree -> SNOD + local heap, dataspace v1, datatype v1,
contiguous layout v3, attribute v1).  Test infrastructure: h5py/libhdf5 are not in the image, so
this is what checks picsp_b200's writer; it follows every pointer the way libhdf5 would and
validates signatures, sizes and ordering on the way."""
import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


class File:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        b = self.b
        if b[:8] != SIG:
            raise H5Error("bad signature")
        (sbv, fsv, rgv, _r, shv, so, sl, _r2) = struct.unpack_from("<8B", b, 8)
        if (sbv, fsv, rgv, shv) != (0, 0, 0, 0) or (so, sl) != (8, 8):
            raise H5Error("unsupported superblock")
        self.leaf_k, self.int_k = struct.unpack_from("<HH", b, 16)
        base, fs, eof, drv = struct.unpack_from("<4Q", b, 24)
        if base != 0 or fs != UNDEF or drv != UNDEF or eof != len(b):
            raise H5Error(f"superblock addresses: base {base} eof {eof} len {len(b)}")
        name_off, hdr, cache, _ = struct.unpack_from("<QQII", b, 56)
        self.root_header = hdr
        if cache == 1:
            self.root_btree, self.root_heap = struct.unpack_from("<QQ", b, 80)
        self.root = self._object(hdr)
        if cache == 1 and (self.root["btree"], self.root["heap"]) != (self.root_btree, self.root_heap):
            raise H5Error("root scratch pad disagrees with the symbol table message")

    # -- object header v1 ------------------------------------------------------------------
    def _messages(self, addr):
        b = self.b
        ver, _r, nmsg, refc, size = struct.unpack_from("<BBHII", b, addr)
        if ver != 1 or refc != 1:
            raise H5Error("object header version/refcount")
        p, end, out = addr + 16, addr + 16 + size, []
        for _ in range(nmsg):
            mtype, msize, flags = struct.unpack_from("<HHB", b, p)
            if msize % 8:
                raise H5Error("message not padded to 8")
            out.append((mtype, b[p + 8:p + 8 + msize]))
            p += 8 + msize
        if p != end:
            raise H5Error("object header size does not match its messages")
        return out

    @staticmethod
    def _datatype(d):
        cls, ver = d[0] & 0x0F, d[0] >> 4
        size = struct.unpack_from("<I", d, 4)[0]
        if ver != 1:
            raise H5Error("datatype version")
        if cls == 1:
            off, prec, eloc, esz, mloc, msz, bias = struct.unpack_from("<HHBBBBI", d, 8)
            if (size, d[1], d[2], off, prec, eloc, esz, mloc, msz, bias) != (8, 0x20, 63, 0, 64, 52, 11, 0, 52, 1023):
                raise H5Error("not IEEE f64 little endian")
            return np.dtype("<f8"), 8 + 12
        if cls == 0:
            off, prec = struct.unpack_from("<HH", d, 8)
            if (size, d[1], off, prec) != (4, 0x08, 0, 32):
                raise H5Error("not int32 little endian signed")
            return np.dtype("<i4"), 8 + 4
        raise H5Error("datatype class")

    @staticmethod
    def _dataspace(d):
        ver, rank, flags = d[0], d[1], d[2]
        if ver != 1 or flags != 0:
            raise H5Error("dataspace version/flags")
        return tuple(struct.unpack_from(f"<{rank}Q", d, 8)) if rank else ()

    def _object(self, addr):
        o = {"attrs": {}}
        for mtype, d in self._messages(addr):
            if mtype == 0x0011:
                o["btree"], o["heap"] = struct.unpack_from("<QQ", d, 0)
            elif mtype == 0x0001:
                o["shape"] = self._dataspace(d)
            elif mtype == 0x0003:
                o["dtype"], _ = self._datatype(d)
            elif mtype == 0x0008:
                ver, cls = d[0], d[1]
                if (ver, cls) != (3, 1):
                    raise H5Error("layout must be v3 contiguous")
                o["data_addr"], o["data_size"] = struct.unpack_from("<QQ", d, 2)
            elif mtype == 0x000C:
                ver, _r, nsz, tsz, ssz = struct.unpack_from("<BBHHH", d, 0)
                if ver != 1:
                    raise H5Error("attribute version")
                p = 8
                name = d[p:p + nsz].split(b"\0")[0].decode(); p += (nsz + 7) // 8 * 8
                dt, _ = self._datatype(d[p:p + tsz]); p += (tsz + 7) // 8 * 8
                shape = self._dataspace(d[p:p + ssz]); p += (ssz + 7) // 8 * 8
                if shape != ():
                    raise H5Error("attribute must be scalar")
                o["attrs"][name] = np.frombuffer(d, dtype=dt, count=1, offset=p)[0]
            elif mtype in (0x0000, 0x0005):
                pass
            else:
                raise H5Error(f"unexpected message type {mtype:#x}")
        return o

    # -- old-style group ----------------------------------------------------------------------
    def _heap(self, addr):
        b = self.b
        if b[addr:addr + 4] != b"HEAP" or b[addr + 4] != 0:
            raise H5Error("heap signature")
        size, free, seg = struct.unpack_from("<QQQ", b, addr + 8)
        if free != 1:
            raise H5Error("heap free list must be H5HL_FREE_NULL (1)")
        return b[seg:seg + size]

    def _name(self, heap, off):
        return heap[off:heap.index(b"\0", off)].decode()

    def members(self, obj):
        """name -> object header address, walking B-tree and symbol table nodes like libhdf5."""
        b, heap, out = self.b, self._heap(obj["heap"]), {}

        def node(addr, upper_key=None):
            if b[addr:addr + 4] != b"TREE":
                raise H5Error("btree signature")
            ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
            left, right = struct.unpack_from("<QQ", b, addr + 8)
            if ntype != 0 or left != UNDEF or right != UNDEF or used > 2 * self.int_k:
                raise H5Error("btree node")
            p = addr + 24
            keys = [struct.unpack_from("<Q", b, p + 16 * i)[0] for i in range(used + 1)]
            kids = [struct.unpack_from("<Q", b, p + 16 * i + 8)[0] for i in range(used)]
            for i, child in enumerate(kids):
                lo, hi = self._name(heap, keys[i]), self._name(heap, keys[i + 1])
                if level > 0:
                    node(child)
                    continue
                if b[child:child + 4] != b"SNOD" or b[child + 4] != 1:
                    raise H5Error("SNOD signature")
                nsym = struct.unpack_from("<H", b, child + 6)[0]
                if nsym > 2 * self.leaf_k:
                    raise H5Error("SNOD overfull for the leaf K in the superblock")
                prev = None
                for e in range(nsym):
                    noff, hdr, cache, _r = struct.unpack_from("<QQII", b, child + 8 + 40 * e)
                    nm = self._name(heap, noff)
                    if prev is not None and not prev < nm:
                        raise H5Error("symbol table entries not sorted")
                    if not (lo < nm <= hi):
                        raise H5Error("entry outside its B-tree key range")
                    prev = nm
                    out[nm] = hdr
        node(obj["btree"])
        return out

    # -- public --------------------------------------------------------------------------------
    def attrs(self):
        return dict(self.root["attrs"])

    def groups(self):
        return {n: self._object(a) for n, a in self.members(self.root).items()}

    def datasets(self, group):
        g = self.groups()[group]
        return sorted(self.members(g))

    def read(self, path):
        parts = [p for p in path.split("/") if p]
        obj = self.root
        for p in parts:
            m = self.members(obj)
            if p not in m:
                raise KeyError(path)
            obj = self._object(m[p])
        if "data_addr" not in obj:
            raise H5Error("not a dataset")
        n = int(np.prod(obj["shape"]))
        if obj["data_size"] != n * obj["dtype"].itemsize:
            raise H5Error("layout size mismatch")
        return np.frombuffer(self.b, dtype=obj["dtype"], count=n, offset=obj["data_addr"]).reshape(obj["shape"]).copy()
