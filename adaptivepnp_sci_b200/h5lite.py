"""Minimal pure-Python reader (and test writer) for the HDF5 subset MATLAB ``-v7.3`` files use.

The reference reads its datasets with h5py (ADMM_TV_Warm_Start_save.py:69-74: ``meas_bayer``, ``mask_bayer``,
``orig_bayer``, ``orig``); h5py / libhdf5 are not installed in this image, so this module restates the file format
(HDF5 File Format Specification, version 1.x structures) for exactly what such files contain:

* an optional user block (MATLAB writes 512 bytes) before the version-0/1 superblock,
* "old-style" groups: symbol-table message -> version-1 B-tree + local heap + symbol-table nodes,
* version-1 object headers with continuation blocks,
* dataspace v1/v2, fixed-point and IEEE floating-point datatypes (little / big endian),
* data layout v3: compact, contiguous, chunked (version-1 chunk B-tree),
* filter pipeline v1/v2 with deflate (1), shuffle (2) and fletcher32 (3).

``File(path)[name]`` returns a numpy array in the HDF5 (C-order) shape, i.e. exactly what ``np.array(h5py.File(path)[name])``
returns, so the transposes of the reference's loader apply unchanged.  Anything outside the subset raises
``H5Unsupported`` naming the structure.  ``write_mat73`` writes such a file (contiguous or chunked + shuffle + deflate)
and exists for the tests: no third-party HDF5 writer is available here, so the reader is validated against files
produced by this spec-following writer (both layouts, both endiannesses of the structures it claims to support)."""
import struct
import zlib

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Unsupported(NotImplementedError):
    pass


class File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.buf = f.read()
        off = 0
        while True:                                      # superblock at 0, 512, 1024, 2048, ...
            if self.buf[off:off + 8] == SIG:
                break
            off = 512 if off == 0 else off * 2
            if off + 8 > len(self.buf):
                raise ValueError("%s: no HDF5 signature found (not a MATLAB v7.3 file?)" % path)
        b = self.buf
        ver = b[off + 8]
        if ver > 1:
            raise H5Unsupported("superblock version %d (only 0/1: MATLAB v7.3)" % ver)
        self.O, self.L = b[off + 13], b[off + 14]
        if self.O != 8 or self.L != 8:
            raise H5Unsupported("offset/length sizes %d/%d" % (self.O, self.L))
        p = off + 24 + (4 if ver == 1 else 0)
        self.base = struct.unpack_from("<Q", b, p)[0]
        p += 32                                           # base, free-space, end-of-file, driver-info addresses
        # root group symbol table entry
        _, hdr_addr, cache_type = struct.unpack_from("<QQI", b, p)
        self.root = self._group_entries(self._messages(hdr_addr))

    # ---- low level ------------------------------------------------------------------------------------
    def _at(self, addr):
        return self.base + addr

    def _messages(self, addr):
        """[(type, bytes)] of a version-1 object header, continuation blocks followed."""
        b = self.buf
        p = self._at(addr)
        if b[p] != 1:
            raise H5Unsupported("object header version %d" % b[p])
        nmsg, _, hsize = struct.unpack_from("<HII", b, p + 2)
        blocks = [(p + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            q, size = blocks.pop(0)
            end = q + size
            while q + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, q)
                data = b[q + 8:q + 8 + msize]
                if mtype == 0x10:
                    caddr, clen = struct.unpack_from("<QQ", data, 0)
                    blocks.append((self._at(caddr), clen))
                out.append((mtype, data))
                q += 8 + msize
        return out

    def _heap_name(self, heap_addr, off):
        b = self.buf
        p = self._at(heap_addr)
        if b[p:p + 4] != b"HEAP":
            raise ValueError("bad local heap")
        seg = struct.unpack_from("<Q", b, p + 24)[0]
        s = self._at(seg) + off
        e = b.index(b"\0", s)
        return b[s:e].decode("ascii")

    def _group_entries(self, msgs):
        for t, d in msgs:
            if t == 0x11:
                btree, heap = struct.unpack_from("<QQ", d, 0)
                ent = {}
                self._walk_group_btree(btree, heap, ent)
                return ent
        raise H5Unsupported("group without a symbol-table message (new-style link messages)")

    def _walk_group_btree(self, addr, heap, ent):
        b = self.buf
        p = self._at(addr)
        if b[p:p + 4] == b"SNOD":
            n = struct.unpack_from("<H", b, p + 6)[0]
            q = p + 8
            for _ in range(n):
                name_off, hdr = struct.unpack_from("<QQ", b, q)
                ent[self._heap_name(heap, name_off)] = hdr
                q += 40
            return
        if b[p:p + 4] != b"TREE" or b[p + 4] != 0:
            raise ValueError("bad group B-tree node")
        n = struct.unpack_from("<H", b, p + 6)[0]
        q = p + 24
        for i in range(n):
            q += 8                                          # key i
            child = struct.unpack_from("<Q", b, q)[0]
            q += 8
            self._walk_group_btree(child, heap, ent)

    # ---- datasets -------------------------------------------------------------------------------------
    def keys(self):
        return list(self.root.keys())

    def __contains__(self, name):
        return name in self.root

    def __getitem__(self, name):
        if name not in self.root:
            raise KeyError(name)
        msgs = self._messages(self.root[name])
        shape = dtype = layout = None
        filters = []
        for t, d in msgs:
            if t == 0x01:
                shape = self._dataspace(d)
            elif t == 0x03:
                dtype = self._datatype(d)
            elif t == 0x08:
                layout = d
            elif t == 0x0B:
                filters = self._filters(d)
        if shape is None or dtype is None or layout is None:
            raise H5Unsupported("%s is not a simple dataset" % name)
        return self._read(shape, dtype, layout, filters)

    @staticmethod
    def _dataspace(d):
        ver, rank = d[0], d[1]
        if ver == 1:
            return tuple(struct.unpack_from("<%dQ" % rank, d, 8))
        if ver == 2:
            return tuple(struct.unpack_from("<%dQ" % rank, d, 4))
        raise H5Unsupported("dataspace version %d" % ver)

    @staticmethod
    def _datatype(d):
        cls, bits0 = d[0] & 0x0F, d[1]
        size = struct.unpack_from("<I", d, 4)[0]
        order = ">" if bits0 & 1 else "<"
        if cls == 1:
            if size not in (2, 4, 8):
                raise H5Unsupported("float size %d" % size)
            return np.dtype(order + "f%d" % size)
        if cls == 0:
            return np.dtype(order + ("i" if bits0 & 8 else "u") + str(size))
        raise H5Unsupported("datatype class %d (only integers and IEEE floats)" % cls)

    @staticmethod
    def _filters(d):
        ver, n = d[0], d[1]
        out = []
        p = 8 if ver == 1 else 2
        for _ in range(n):
            fid = struct.unpack_from("<H", d, p)[0]
            p += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen = struct.unpack_from("<H", d, p)[0]
                p += 2
            _flags, ncd = struct.unpack_from("<HH", d, p)
            p += 4
            if nlen:
                p += (nlen + 7) // 8 * 8 if ver == 1 else nlen
            cd = struct.unpack_from("<%dI" % ncd, d, p)
            p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out

    def _read(self, shape, dtype, layout, filters):
        b = self.buf
        n = int(np.prod(shape)) if shape else 1
        if layout[0] != 3:
            raise H5Unsupported("data layout message version %d" % layout[0])
        cls = layout[1]
        if cls == 0:
            size = struct.unpack_from("<H", layout, 2)[0]
            return np.frombuffer(layout[4:4 + size], dtype=dtype, count=n).reshape(shape).copy()
        if cls == 1:
            addr, _size = struct.unpack_from("<QQ", layout, 2)
            if addr == UNDEF:
                return np.zeros(shape, dtype=dtype)
            return np.frombuffer(b, dtype=dtype, count=n, offset=self._at(addr)).reshape(shape).copy()
        if cls != 2:
            raise H5Unsupported("layout class %d" % cls)
        ndim = layout[2]
        btree = struct.unpack_from("<Q", layout, 3)[0]
        cdims = struct.unpack_from("<%dI" % ndim, layout, 11)
        chunk = tuple(cdims[:-1])
        if len(chunk) != len(shape):
            raise ValueError("chunk rank mismatch")
        out = np.zeros(shape, dtype=dtype)
        if btree != UNDEF:
            self._walk_chunks(btree, ndim, chunk, dtype, filters, out)
        return out

    def _walk_chunks(self, addr, ndim, chunk, dtype, filters, out):
        b = self.buf
        p = self._at(addr)
        if b[p:p + 4] != b"TREE" or b[p + 4] != 1:
            raise ValueError("bad chunk B-tree node")
        level = b[p + 5]
        n = struct.unpack_from("<H", b, p + 6)[0]
        q = p + 24
        ksz = 8 + 8 * ndim
        for _ in range(n):
            size, fmask = struct.unpack_from("<II", b, q)
            offs = struct.unpack_from("<%dQ" % ndim, b, q + 8)
            child = struct.unpack_from("<Q", b, q + ksz)[0]
            q += ksz + 8
            if level > 0:
                self._walk_chunks(child, ndim, chunk, dtype, filters, out)
                continue
            raw = b[self._at(child):self._at(child) + size]
            for i in range(len(filters) - 1, -1, -1):       # undo the pipeline back to front
                if fmask & (1 << i):
                    continue
                fid = filters[i][0]
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = dtype.itemsize
                    a = np.frombuffer(raw, dtype=np.uint8)
                    m = a.size // es
                    raw = a[:m * es].reshape(es, m).T.tobytes() + a[m * es:].tobytes()
                elif fid == 3:
                    raw = raw[:-4]
                else:
                    raise H5Unsupported("filter id %d" % fid)
            blk = np.frombuffer(raw, dtype=dtype, count=int(np.prod(chunk))).reshape(chunk)
            sl_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, out.shape))
            sl_blk = tuple(slice(0, s.stop - s.start) for s in sl_out)
            out[sl_out] = blk[sl_blk]


# ---------------------------------------------------------------------------------------------------------------------
# test writer
# ---------------------------------------------------------------------------------------------------------------------
def write_mat73(path, arrays, chunks=None, userblock=512):
    """Write ``arrays`` (name -> ndarray, stored in the given C-order shape) as a MATLAB-v7.3-style HDF5 file: user block,
    version-0 superblock, one old-style root group, version-1 object headers.  ``chunks`` (name -> chunk shape) selects the
    chunked layout with shuffle + deflate for that dataset, otherwise it is stored contiguously."""
    chunks = chunks or {}
    out = bytearray(b"MATLAB 7.3 MAT-file".ljust(userblock, b" ")) if userblock else bytearray()
    base = len(out)
    out += b"\0" * 96                                        # superblock placeholder (24 + 32 + 40 bytes)

    def align():
        while (len(out) - base) % 8:
            out.append(0)

    def addr():
        return len(out) - base

    def msg(t, data):
        data = bytes(data)
        pad = (-len(data)) % 8
        return struct.pack("<HHB3x", t, len(data) + pad, 0) + data + b"\0" * pad

    def header(msgs):
        align()
        a = addr()
        body = b"".join(msgs)
        out.extend(struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body)
        return a

    names = sorted(arrays)
    hdr_addr = {}
    for name in names:
        a = np.ascontiguousarray(arrays[name])
        dt = a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype
        a = a.astype(dt)
        space = struct.pack("<BBB5x", 1, a.ndim, 0) + struct.pack("<%dQ" % a.ndim, *a.shape)
        if a.dtype.kind == "f":
            exp_bits, mant_bits = {2: (5, 10), 4: (8, 23), 8: (11, 52)}[a.itemsize]
            dtype = struct.pack("<BBBBI", 0x11, 0x20, 8 * a.itemsize - 1, 0, a.itemsize) + \
                struct.pack("<HHBBBBI", 0, 8 * a.itemsize, mant_bits, exp_bits, 0, mant_bits, (1 << (exp_bits - 1)) - 1)
        else:
            dtype = struct.pack("<BBBBI", 0x10, 0x08 if a.dtype.kind == "i" else 0, 0, 0, a.itemsize) + \
                struct.pack("<HH", 0, 8 * a.itemsize)
        if name in chunks:
            ch = tuple(chunks[name])
            recs = []
            grid = [range(0, s, c) for s, c in zip(a.shape, ch)]
            for offs in np.ndindex(*[len(g) for g in grid]):
                o = tuple(g[i] for g, i in zip(grid, offs))
                blk = np.zeros(ch, dtype=a.dtype)
                sl = tuple(slice(oo, min(oo + c, s)) for oo, c, s in zip(o, ch, a.shape))
                blk[tuple(slice(0, s.stop - s.start) for s in sl)] = a[sl]
                raw = np.frombuffer(blk.tobytes(), dtype=np.uint8).reshape(-1, a.itemsize).T.tobytes()      # shuffle
                raw = zlib.compress(raw, 6)                                                                 # deflate
                align()
                recs.append((len(raw), o, addr()))
                out.extend(raw)
            align()
            bt = addr()
            nd = a.ndim + 1
            node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(recs), UNDEF, UNDEF)
            for size, o, ca in recs:
                node += struct.pack("<II", size, 0) + struct.pack("<%dQ" % nd, *(o + (0,))) + struct.pack("<Q", ca)
            node += struct.pack("<II", 0, 0) + struct.pack("<%dQ" % nd, *(a.shape + (0,)))                # final key
            out.extend(node)
            layout = struct.pack("<BBB", 3, 2, nd) + struct.pack("<Q", bt) + struct.pack("<%dI" % nd, *(ch + (a.itemsize,)))
            pipeline = struct.pack("<BB6x", 1, 2) + \
                struct.pack("<HHHH", 2, 0, 1, 1) + struct.pack("<I", a.itemsize) + b"\0" * 4 + \
                struct.pack("<HHHH", 1, 0, 1, 1) + struct.pack("<I", 6) + b"\0" * 4
            msgs = [msg(1, space), msg(3, dtype), msg(0x0B, pipeline), msg(8, layout)]
        else:
            align()
            da = addr()
            out.extend(a.tobytes())
            layout = struct.pack("<BB", 3, 1) + struct.pack("<QQ", da, a.nbytes)
            msgs = [msg(1, space), msg(3, dtype), msg(8, layout)]
        hdr_addr[name] = header(msgs)
    # local heap with the names (offset 0 = empty string for the root)
    heap_data = bytearray(b"\0" * 8)
    name_off = {}
    for name in names:
        name_off[name] = len(heap_data)
        heap_data += name.encode("ascii") + b"\0"
        while len(heap_data) % 8:
            heap_data.append(0)
    align()
    seg = addr()
    out.extend(heap_data)
    align()
    heap = addr()
    out.extend(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, seg))
    align()
    snod = addr()
    node = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for name in names:
        node += struct.pack("<QQII16x", name_off[name], hdr_addr[name], 0, 0)
    out.extend(node)
    align()
    bt = addr()
    out.extend(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod, name_off[names[-1]]))
    root = header([msg(0x11, struct.pack("<QQ", bt, heap))])
    sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0) + \
        struct.pack("<QQQQ", base, UNDEF, len(out) - base, UNDEF) + struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", bt, heap)
    out[base:base + len(sb)] = sb
    with open(path, "wb") as f:
        f.write(bytes(out))
