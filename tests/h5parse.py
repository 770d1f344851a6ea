"""A small independent reader of the HDF5 structures the drop-in writes (no h5py / libhdf5 in this image): superblock
version 0, version-1 object headers, symbol-table groups (v1 B-tree "TREE" + symbol nodes "SNOD" + local heap "HEAP"),
dataspace v1, datatype v1 (fixed point, IEEE float, compound), contiguous layout v3, attribute v1.  Written from the
HDF5 File Format Specification; it checks signatures, sizes, sort order and B-tree keys as it walks, so a file that a
real libhdf5 would reject for those reasons fails here too.  Test infrastructure."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


class Node:
    def __init__(self, name):
        self.name = name
        self.children = {}      # groups
        self.attrs = {}
        self.dtype = None       # datasets
        self.shape = None
        self.addr = None
        self.nbytes = None
        self.is_group = False


class File:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        b = self.b
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise H5Error("bad signature")
        ver, fsver, rgver, _, shver, so, sl, _ = struct.unpack_from("<8B", b, 8)
        if (ver, fsver, rgver, shver, so, sl) != (0, 0, 0, 0, 8, 8):
            raise H5Error("unexpected superblock versions / sizes")
        self.leaf_k, self.int_k, flags = struct.unpack_from("<HHI", b, 16)
        base, fsaddr, self.eof, drv = struct.unpack_from("<4Q", b, 24)
        if base != 0 or fsaddr != UNDEF or drv != UNDEF or flags != 0:
            raise H5Error("unexpected superblock addresses")
        if self.eof != len(b):
            raise H5Error("end-of-file address %d != file size %d" % (self.eof, len(b)))
        nameoff, ohdr, ctype, _, bt, heap = struct.unpack_from("<QQIIQQ", b, 56)
        self.root = self._object("/", ohdr)
        if ctype == 1 and (bt, heap) != (self.root._stab):
            raise H5Error("root entry scratch pad disagrees with the symbol table message")

    # ------------------------------------------------------------------ object headers
    def _object(self, name, at):
        b = self.b
        ver, _, nmsg, refc, hsize = struct.unpack_from("<BBHII", b, at)
        if ver != 1 or refc < 1 or hsize % 8:
            raise H5Error("bad object header at %d" % at)
        p, end = at + 16, at + 16 + hsize
        node = Node(name)
        layout = None
        seen = 0
        while p < end:
            mtype, msize, mflags = struct.unpack_from("<HHB", b, p)
            if msize % 8:
                raise H5Error("message size not a multiple of 8")
            d = p + 8
            if mtype == 0x0001:
                node.shape = self._dataspace(d)
            elif mtype == 0x0003:
                node.dtype, _ = self._datatype(d)
            elif mtype == 0x0005:
                fv = struct.unpack_from("<4B", b, d)
                if fv[0] != 2:
                    raise H5Error("fill value message version %d" % fv[0])
            elif mtype == 0x0008:
                v, cls = struct.unpack_from("<BB", b, d)
                if (v, cls) != (3, 1):
                    raise H5Error("layout must be version 3 contiguous")
                layout = struct.unpack_from("<QQ", b, d + 2)
            elif mtype == 0x0011:
                node._stab = struct.unpack_from("<QQ", b, d)
                node.is_group = True
            elif mtype == 0x000C:
                k, v = self._attribute(d)
                node.attrs[k] = v
            elif mtype != 0:
                raise H5Error("unexpected message type 0x%04x" % mtype)
            p = d + msize
            seen += 1
        if p != end or seen != nmsg:
            raise H5Error("object header at %d: %d messages / %d bytes declared, %d / %d found" % (at, nmsg, hsize, seen, p - at - 16))
        if node.is_group:
            self._group(node)
        else:
            if node.shape is None or node.dtype is None or layout is None:
                raise H5Error("dataset %s lacks dataspace / datatype / layout" % name)
            node.addr, node.nbytes = layout
            want = int(np.prod(node.shape, dtype=np.int64)) * node.dtype.itemsize
            if node.nbytes != want:
                raise H5Error("dataset %s: layout size %d != %d" % (name, node.nbytes, want))
            if node.addr % 8 or node.addr + node.nbytes > len(b):
                raise H5Error("dataset %s: bad address" % name)
        return node

    def _dataspace(self, d):
        v, rank, flags = struct.unpack_from("<BBB", self.b, d)
        if v != 1 or flags != 0:
            raise H5Error("dataspace version / flags")
        return tuple(struct.unpack_from("<%dQ" % rank, self.b, d + 8))

    def _datatype(self, d):
        b = self.b
        cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", b, d)
        cls, ver = cv & 15, cv >> 4
        if ver != 1:
            raise H5Error("datatype version %d" % ver)
        if cls == 0:
            off, prec = struct.unpack_from("<HH", b, d + 8)
            if (b0 & 1) or off != 0 or prec != 8 * size:
                raise H5Error("fixed-point type is not plain little endian")
            return np.dtype("<%s%d" % ("i" if b0 & 8 else "u", size)), 12
        if cls == 1:
            off, prec, eloc, esize, mloc, msize, bias = struct.unpack_from("<HHBBBBI", b, d + 8)
            if (size, b0, b1, off, prec, eloc, esize, mloc, msize, bias) != (8, 0x20, 63, 0, 64, 52, 11, 0, 52, 1023):
                raise H5Error("float type is not IEEE binary64 little endian")
            return np.dtype("<f8"), 20
        if cls == 6:
            nmem = b0 | (b1 << 8)
            p = d + 8
            names, formats, offsets = [], [], []
            for _ in range(nmem):
                e = b.index(b"\0", p)
                nm = b[p:e].decode()
                p += (e - p + 1 + 7) // 8 * 8
                moff, ndim = struct.unpack_from("<IB", b, p)
                if ndim != 0:
                    raise H5Error("array members are not expected")
                p += 4 + 4 + 4 + 4 + 16
                mt, used = self._datatype(p)
                p += used
                names.append(nm); formats.append(mt); offsets.append(moff)
            return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": size}), p - d
        raise H5Error("datatype class %d" % cls)

    def _attribute(self, d):
        b = self.b
        v, _, nsz, tsz, ssz = struct.unpack_from("<BBHHH", b, d)
        if v != 1:
            raise H5Error("attribute version")
        p = d + 8
        name = b[p:p + nsz - 1].decode()
        if b[p + nsz - 1] != 0:
            raise H5Error("attribute name not terminated")
        p += (nsz + 7) // 8 * 8
        dt, used = self._datatype(p)
        if used != tsz:
            raise H5Error("attribute datatype size %d != %d" % (used, tsz))
        p += (tsz + 7) // 8 * 8
        shape = self._dataspace(p)
        p += (ssz + 7) // 8 * 8
        n = int(np.prod(shape, dtype=np.int64))
        return name, np.frombuffer(b, dtype=dt, count=n, offset=p).reshape(shape).copy()

    # ------------------------------------------------------------------ groups
    def _heap_name(self, heap, off):
        b = self.b
        if b[heap:heap + 4] != b"HEAP" or b[heap + 4] != 0:
            raise H5Error("bad local heap")
        size, free, data = struct.unpack_from("<QQQ", b, heap + 8)
        if free != 1 and free >= size:
            raise H5Error("bad heap free list")
        if off >= size:
            raise H5Error("name offset outside the heap")
        e = b.index(b"\0", data + off)
        return b[data + off:e].decode()

    def _group(self, node):
        bt, heap = node._stab
        names = []
        self._btree(node, bt, heap, names, None)
        if names != sorted(names):
            raise H5Error("group %s: entries are not in name order" % node.name)

    def _btree(self, node, at, heap, names, expect_left):
        b = self.b
        if b[at:at + 4] != b"TREE":
            raise H5Error("bad B-tree node at %d" % at)
        ntype, level, used, left, right = struct.unpack_from("<BBHQQ", b, at + 4)
        if ntype != 0 or used > 2 * self.int_k:
            raise H5Error("bad group B-tree node")
        p = at + 24
        keys = []
        kids = []
        for i in range(used):
            keys.append(struct.unpack_from("<Q", b, p)[0]); kids.append(struct.unpack_from("<Q", b, p + 8)[0]); p += 16
        keys.append(struct.unpack_from("<Q", b, p)[0])
        for i, kid in enumerate(kids):
            before = len(names)
            if level > 0:
                self._btree(node, kid, heap, names, None)
            else:
                self._snod(node, kid, heap, names)
            got = names[before:]
            lo, hi = self._heap_name(heap, keys[i]), self._heap_name(heap, keys[i + 1])
            if got and not (lo < got[0] and got[-1] <= hi and got[-1] == hi):
                raise H5Error("B-tree keys (%r, %r] do not bracket %r..%r" % (lo, hi, got[0], got[-1]))

    def _snod(self, node, at, heap, names):
        b = self.b
        if b[at:at + 4] != b"SNOD" or b[at + 4] != 1:
            raise H5Error("bad symbol node at %d" % at)
        n = struct.unpack_from("<H", b, at + 6)[0]
        if n > 2 * self.leaf_k:
            raise H5Error("symbol node over capacity")
        for i in range(n):
            nameoff, ohdr, ctype, _, s0, s1 = struct.unpack_from("<QQIIQQ", b, at + 8 + 40 * i)
            nm = self._heap_name(heap, nameoff)
            child = self._object(nm, ohdr)
            if ctype == 1 and (s0, s1) != child._stab:
                raise H5Error("cached symbol table of %s is stale" % nm)
            node.children[nm] = child
            names.append(nm)

    # ------------------------------------------------------------------ access
    def __getitem__(self, path):
        node = self.root
        for part in [p for p in path.split("/") if p]:
            node = node.children[part]
        return node

    def read(self, path):
        n = self[path]
        cnt = int(np.prod(n.shape, dtype=np.int64))
        return np.frombuffer(self.b, dtype=n.dtype, count=cnt, offset=n.addr).reshape(n.shape)
