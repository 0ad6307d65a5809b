"""h5lite - the subset of HDF5 the reference's files use, without h5py.

The reference reads its input spectra with ``h5py.File(name, 'r')['wavelength'|'flux'|'noise'][:]``
(vamp_1.0/vpspectrum.py:58-64) and writes ``*_params.h5`` / ``*_flux_model.h5`` with
``f.create_dataset(key, data=np.array(value))`` (vpspectrum.py:528-538).  Files made that way are plain: a version-0
superblock, a root group stored as a symbol table (``TREE`` B-tree node, ``SNOD`` symbol nodes, ``HEAP`` local
heap), version-1 object headers, contiguous little-endian integer / IEEE datasets, no filters, no attributes.  This
module reads exactly that (plus nested groups of the same kind and compact layouts) and writes files of the same
shape, so that sightlines in the reference's format go in and its output files come out on a box without h5py
(this image has none).  The reader is checked against the reference's own fixtures (written by h5py), the writer
against the reader and, message by message, against the bytes of those fixtures (tests/test_h5lite.py).

Anything outside the subset (chunked or filtered datasets, strings, compound types, newer superblocks) raises
``H5Unsupported`` - never a silent wrong answer.  Format reference: the HDF5 File Format Specification, version 1.1
(superblock 0, B-tree v1, symbol table nodes, local heaps, object header v1 and its messages 0x0001 dataspace,
0x0003 datatype, 0x0005 fill value, 0x0008 layout, 0x0010 continuation, 0x0011 symbol table).
"""
import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Unsupported(ValueError):
    pass


# ------------------------------------------------------------------------------------------------ reading
class _Reader:
    def __init__(self, buf):
        self.b = buf
        if buf[:8] != SIGNATURE:
            raise H5Unsupported("not an HDF5 file (signature at offset 0 expected)")
        ver = buf[8]
        if ver not in (0, 1):
            raise H5Unsupported("superblock version %d (only 0 / 1: files written with h5py's defaults)" % ver)
        if buf[13] != 8 or buf[14] != 8:
            raise H5Unsupported("offset / length sizes %d / %d (8 / 8 expected)" % (buf[13], buf[14]))
        o = 24 + (4 if ver == 1 else 0)
        self.base, _, self.eof, _ = struct.unpack_from("<4Q", buf, o)
        o += 32
        # root group symbol table entry
        _, self.root_header, cache, _ = struct.unpack_from("<QQII", buf, o)

    def u(self, fmt, off):
        return struct.unpack_from("<" + fmt, self.b, off)

    # -- object headers (version 1)
    def messages(self, addr):
        addr += self.base
        ver, _, nmsg, _, size = self.u("BBHII", addr)
        if ver != 1:
            raise H5Unsupported("object header version %d (only 1)" % ver)
        out = []
        blocks = [(addr + 16, size)]
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = self.u("HHB", p)
                data = p + 8
                if mtype == 0x0010:                                   # continuation
                    caddr, clen = self.u("QQ", data)
                    blocks.append((caddr + self.base, clen))
                out.append((mtype, data, msize, flags))
                p = data + msize
        return out

    def heap_name(self, heap_addr, offset):
        h = heap_addr + self.base
        if self.b[h:h + 4] != b"HEAP":
            raise H5Unsupported("local heap signature missing")
        seg = self.u("Q", h + 24)[0] + self.base
        s = seg + offset
        e = self.b.index(b"\0", s)
        return self.b[s:e].decode("utf-8")

    def group_entries(self, btree_addr, heap_addr):
        """(name, object header address) of a symbol-table group, through its v1 B-tree."""
        out = []
        a = btree_addr + self.base
        if self.b[a:a + 4] != b"TREE":
            raise H5Unsupported("B-tree signature missing")
        ntype, level, used = self.u("BBH", a + 4)
        if ntype != 0:
            raise H5Unsupported("B-tree node type %d in a group" % ntype)
        p = a + 24
        for i in range(used):
            child = self.u("Q", p + 8)[0]
            p += 16
            if level > 0:
                out += self.group_entries(child, heap_addr)
            else:
                s = child + self.base
                if self.b[s:s + 4] != b"SNOD":
                    raise H5Unsupported("symbol table node signature missing")
                nsym = self.u("H", s + 6)[0]
                for k in range(nsym):
                    name_off, hdr = self.u("QQ", s + 8 + 40 * k)
                    out.append((self.heap_name(heap_addr, name_off), hdr))
        return out

    def dtype(self, p):
        cv, b0, b1, b2, size = self.u("BBBBI", p)
        cls, ver = cv & 0x0F, cv >> 4
        if b0 & 1:
            raise H5Unsupported("big-endian data")
        if cls == 0:
            return np.dtype("<%s%d" % ("i" if (b0 & 0x08) else "u", size))
        if cls == 1:
            if size not in (2, 4, 8):
                raise H5Unsupported("%d-byte floating point" % size)
            return np.dtype("<f%d" % size)
        raise H5Unsupported("datatype class %d (only integers and IEEE floats)" % cls)

    def dataset(self, msgs):
        shape = dt = None
        layout = None
        for mtype, p, msize, flags in msgs:
            if mtype == 0x0001:
                ver, rank, fl = self.u("BBB", p)
                q = p + (8 if ver == 1 else 4)
                shape = tuple(self.u("%dQ" % rank, q)) if rank else ()
            elif mtype == 0x0003:
                dt = self.dtype(p)
            elif mtype == 0x0008:
                ver, cls = self.u("BB", p)
                if ver != 3:
                    raise H5Unsupported("data layout message version %d (only 3)" % ver)
                if cls == 1:
                    layout = ("contiguous",) + self.u("QQ", p + 2)
                elif cls == 0:
                    layout = ("compact", p + 4, self.u("H", p + 2)[0])
                else:
                    raise H5Unsupported("chunked dataset (only contiguous / compact layouts)")
            elif mtype == 0x000B:
                raise H5Unsupported("filtered dataset")
        if shape is None or dt is None or layout is None:
            return None
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if layout[0] == "contiguous":
            addr = layout[1]
            if addr == UNDEF:                                          # never written: fill value 0
                return np.zeros(shape, dtype=dt)
            start = addr + self.base
        else:
            start = layout[1]
        return np.frombuffer(self.b, dtype=dt, count=n, offset=start).reshape(shape).copy()

    def walk(self, header, prefix, out):
        msgs = self.messages(header)
        for mtype, p, msize, flags in msgs:
            if mtype == 0x0011:                                        # a group: symbol table message
                btree, heap = self.u("QQ", p)
                for name, hdr in self.group_entries(btree, heap):
                    self.walk(hdr, prefix + name + "/", out)
                return
        arr = self.dataset(msgs)
        if arr is not None:
            out[prefix[:-1]] = arr


def read_h5(path):
    """All datasets of the file as ``{name: ndarray}`` (datasets of nested groups as ``group/name``)."""
    with open(path, "rb") as fh:
        buf = fh.read()
    r = _Reader(buf)
    out = {}
    r.walk(r.root_header, "", out)
    return out


# ------------------------------------------------------------------------------------------------ writing
def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _msg(mtype, data, flags=0):
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _dtype_message(dt):
    dt = np.dtype(dt)
    if dt.kind in "iu":
        body = struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize)
        body += struct.pack("<HH", 0, 8 * dt.itemsize)
    elif dt.kind == "f" and dt.itemsize in (4, 8):
        ebits, mbits, bias = (11, 52, 1023) if dt.itemsize == 8 else (8, 23, 127)
        body = struct.pack("<BBBBI", 0x11, 0x20, 8 * dt.itemsize - 1, 0, dt.itemsize)
        body += struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, mbits, ebits, 0, mbits, bias)
    else:
        raise H5Unsupported("dtype %s (integers, float32 and float64 only)" % dt)
    return _msg(0x0003, body, flags=1)


def _dataset_header(arr, data_addr):
    rank = arr.ndim
    space = struct.pack("<BBB5x", 1, rank, 1 if rank else 0)
    if rank:
        space += struct.pack("<%dQ" % rank, *arr.shape) * 2            # dimensions, maximum dimensions
    msgs = [_msg(0x0001, space),
            _dtype_message(arr.dtype),
            _msg(0x0005, bytes([2, 2, 2, 1, 0, 0, 0, 0]), flags=1),    # fill value v2: late allocation, size 0
            _msg(0x0008, struct.pack("<BBQQ", 3, 1, data_addr if arr.nbytes else UNDEF, arr.nbytes))]
    body = b"".join(msgs)
    return struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body


def _normalise(value):
    a = np.asarray(value)
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    if a.dtype.kind == "f" and a.dtype.itemsize not in (4, 8):
        a = a.astype(np.float64)
    if a.dtype.kind not in "iuf":
        raise H5Unsupported("dataset of dtype %s" % a.dtype)
    return np.asarray(a.astype(a.dtype.newbyteorder("<"), copy=False), order="C")     # (0-d arrays stay 0-d)


def write_h5(path, datasets):
    """Write ``{name: array-like}`` as root-level contiguous datasets (what
    ``h5py.File(path, 'w').create_dataset(name, data=...)`` produces for each entry)."""
    items = sorted(((str(k), _normalise(v)) for k, v in datasets.items()), key=lambda kv: kv[0].encode())
    for name, _ in items:
        if "/" in name or not name:
            raise H5Unsupported("dataset name %r (root-level names only)" % name)
    n = len(items)
    leaf_k = max(4, (n + 1) // 2)                                      # one symbol node holds 2 K entries
    if leaf_k > 0x7FFF:
        raise H5Unsupported("too many datasets")
    internal_k = 16
    # ---- local heap data segment: "" at 0, then the names, then one free block
    heap = bytearray(_pad8(b"\0"))
    name_off = []
    for name, _ in items:
        name_off.append(len(heap))
        heap += _pad8(name.encode("utf-8") + b"\0")
    free_off = len(heap)
    heap += struct.pack("<QQ", 1, 16)                                  # free block: next = H5HL_FREE_NULL, size 16 (itself)
    # ---- addresses
    a_root = 96                                                        # superblock (56 bytes) + root symbol table entry (40)
    root_hdr = struct.pack("<BBHII4x", 1, 0, 1, 1, 24)                 # + symbol table message, filled in below
    a_btree = a_root + 16 + 24
    btree_size = 24 + (2 * internal_k + 1) * 8 + 2 * internal_k * 8
    a_heap = a_btree + btree_size
    a_snod = a_heap + 32
    snod_size = 8 + 2 * leaf_k * 40
    a_hdr0 = a_snod + snod_size
    hdr_addrs, p = [], a_hdr0
    hdr_len = [len(_dataset_header(a, 0)) for _, a in items]
    for L in hdr_len:
        hdr_addrs.append(p)
        p += L
    a_heapdata = p
    p += len(heap)
    p += -p % 8
    data_addrs = []
    for _, a in items:
        data_addrs.append(p)
        p += a.nbytes
        p += -p % 8
    eof = p
    # ---- assemble
    out = bytearray(eof)
    out[0:8] = SIGNATURE
    struct.pack_into("<BBBBBBBBHHI", out, 8, 0, 0, 0, 0, 0, 8, 8, 0, leaf_k, internal_k, 0)
    struct.pack_into("<4Q", out, 24, 0, UNDEF, eof, UNDEF)
    struct.pack_into("<QQII", out, 56, 0, a_root, 1, 0)                 # root entry: cached symbol table addresses
    struct.pack_into("<QQ", out, 80, a_btree, a_heap)
    out[a_root: a_root + 16] = root_hdr
    out[a_root + 16: a_root + 40] = _msg(0x0011, struct.pack("<QQ", a_btree, a_heap))
    # B-tree: one leaf-level node with one child (the symbol node); keys: "" and the largest name
    out[a_btree: a_btree + 4] = b"TREE"
    struct.pack_into("<BBHQQ", out, a_btree + 4, 0, 0, 1 if n else 0, UNDEF, UNDEF)
    if n:
        struct.pack_into("<QQQ", out, a_btree + 24, 0, a_snod, name_off[-1])
    out[a_heap: a_heap + 4] = b"HEAP"
    struct.pack_into("<QQQ", out, a_heap + 8, len(heap), free_off, a_heapdata)
    out[a_snod: a_snod + 4] = b"SNOD"
    struct.pack_into("<BBH", out, a_snod + 4, 1, 0, n)
    for k, ((name, a), ha, da) in enumerate(zip(items, hdr_addrs, data_addrs)):
        struct.pack_into("<QQII", out, a_snod + 8 + 40 * k, name_off[k], ha, 0, 0)
        h = _dataset_header(a, da)
        out[ha: ha + len(h)] = h
        out[da: da + a.nbytes] = a.tobytes()
    out[a_heapdata: a_heapdata + len(heap)] = heap
    with open(path, "wb") as fh:
        fh.write(bytes(out))
    return eof
