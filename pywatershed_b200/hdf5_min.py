"""A minimal read-only HDF5 / netCDF-4 reader in pure Python + numpy.

pywatershed's inputs are netCDF-4 files (parameters_<Process>.nc, prcp.nc,
tmax.nc, tmin.nc; utils/netcdf_utils.py:31-353, parameters.py:263-321), i.e.
HDF5 containers.  Where the netCDF4 module is not installed this reader lets
those files load anyway.  It covers what the netCDF-4 library and h5py /
h5netcdf write for such files and nothing more:

  * superblock versions 0-3, object headers v1 and v2 (with continuations)
  * groups: link messages (compact), fractal heap (dense) and old-style symbol
    tables (B-tree v1 + local heap); only the root group's datasets are listed
  * datasets: compact, contiguous and chunked (B-tree v1 index) layouts with the
    deflate, shuffle and fletcher32 filters; integers, floats, fixed strings
  * attributes: compact and dense; numeric, fixed and variable-length strings

Anything else (virtual datasets, layout v4 chunk indexes, compound data,
external links ...) raises NotImplementedError with the feature named.
"""
from __future__ import annotations

import mmap
import os
import struct
import zlib

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


_THREAD_BYTES = 32 << 20  # chunked variables at least this large are decoded by several threads


class H5Error(NotImplementedError):
    pass


class _Reader:
    def __init__(self, buf, off_size=8, len_size=8):
        self.b = buf
        self.O = off_size
        self.L = len_size

    def u(self, pos, n):
        return int.from_bytes(self.b[pos:pos + n], "little")

    def off(self, pos):
        v = self.u(pos, self.O)
        return None if v == (1 << (8 * self.O)) - 1 else v

    def length(self, pos):
        return self.u(pos, self.L)


class Datatype:
    """Decoded datatype message: numpy dtype for fixed-size classes, or
    kind 'vlen_str' / 'vlen' / 'other'."""

    def __init__(self, cls, size, dtype=None, kind="num", base=None, raw_len=0):
        self.cls = cls
        self.size = size
        self.dtype = dtype
        self.kind = kind
        self.base = base
        self.raw_len = raw_len  # bytes the message occupies


def parse_datatype(b, pos) -> Datatype:
    cv = b[pos]
    cls, ver = cv & 0x0F, cv >> 4
    bits = b[pos + 1] | (b[pos + 2] << 8) | (b[pos + 3] << 16)
    size = int.from_bytes(b[pos + 4:pos + 8], "little")
    p = pos + 8
    if cls == 0:  # fixed-point
        order = ">" if bits & 1 else "<"
        signed = bool(bits & 0x08)
        dt = np.dtype(f"{order}{'i' if signed else 'u'}{size}")
        return Datatype(cls, size, dt, raw_len=12)
    if cls == 1:  # floating point
        order = ">" if bits & 1 else "<"
        dt = np.dtype(f"{order}f{size}")
        return Datatype(cls, size, dt, raw_len=20)
    if cls == 3:  # fixed-length string
        return Datatype(cls, size, np.dtype(f"S{size}"), kind="str", raw_len=8)
    if cls == 9:  # variable length
        base = parse_datatype(b, p)
        is_str = (bits & 0x0F) == 1
        return Datatype(cls, size, None, kind="vlen_str" if is_str else "vlen", base=base,
                        raw_len=8 + base.raw_len)
    if cls == 7:  # reference
        return Datatype(cls, size, np.dtype(f"V{size}"), kind="ref", raw_len=8)
    if cls == 8:  # enumeration (netCDF booleans): base type + names + values
        base = parse_datatype(b, p)
        nmemb = bits & 0xFFFF
        q = p + base.raw_len
        for _ in range(nmemb):
            e = b.find(b"\x00", q)
            n = e - q + 1
            q += n if ver >= 3 else (n + 7) // 8 * 8
        q += nmemb * base.size
        return Datatype(cls, size, base.dtype, kind="num", base=base, raw_len=q - pos)
    if cls == 6:  # compound: only skipped over (DIMENSION_LIST / REFERENCE_LIST attributes)
        nmemb = bits & 0xFFFF
        q = p
        for _ in range(nmemb):
            e = b.find(b"\x00", q)
            n = e - q + 1
            if ver >= 3:
                q += n
                osz = 1
                while (1 << (8 * osz)) <= size and osz < 8:
                    osz *= 2
                # member offset is stored in the minimum number of bytes for `size`
                nb = 1
                while size >= (1 << (8 * nb)):
                    nb += 1
                q += nb
            else:
                q += (n + 7) // 8 * 8
                q += 4
                if ver == 1:
                    q += 1 + 3 + 4 + 4 + 16
            m = parse_datatype(b, q)
            q += m.raw_len
        return Datatype(cls, size, np.dtype(f"V{size}"), kind="other", raw_len=q - pos)
    if cls == 10:  # array
        rank = b[p]
        q = p + (4 if ver < 3 else 1)
        dims = [int.from_bytes(b[q + 4 * i:q + 4 * i + 4], "little") for i in range(rank)]
        q += 4 * rank
        if ver < 3:
            q += 4 * rank  # permutation indices
        base = parse_datatype(b, q)
        dt = np.dtype((base.dtype, tuple(dims))) if base.dtype is not None else None
        return Datatype(cls, size, dt, kind="num" if dt is not None else "other", base=base,
                        raw_len=q + base.raw_len - pos)
    if cls in (2, 4, 5):  # time, bitfield, opaque
        extra = {2: 2, 4: 4, 5: (bits & 0xFF) + (-(bits & 0xFF)) % 8}[cls]
        return Datatype(cls, size, np.dtype(f"V{size}"), kind="other", raw_len=8 + extra)
    raise H5Error(f"HDF5 datatype class {cls}")


def parse_dataspace(r: _Reader, pos):
    """-> (shape tuple or None for a null dataspace, bytes used)."""
    b = r.b
    ver, rank, flags = b[pos], b[pos + 1], b[pos + 2]
    if ver == 1:
        p = pos + 8
    elif ver == 2:
        if b[pos + 3] == 2:
            return None, 4
        p = pos + 4
    else:
        raise H5Error(f"dataspace message version {ver}")
    dims = tuple(r.length(p + i * r.L) for i in range(rank))
    p += rank * r.L
    if flags & 1:
        p += rank * r.L
    if ver == 1 and flags & 2:
        p += rank * r.L
    return dims, p - pos


class Dataset:
    def __init__(self, f: "File", name, msgs):
        self.f = f
        self.name = name
        self._msgs = msgs
        self.shape = ()
        self.dt = None
        self.layout = None
        self.filters = []
        self._attrs = None
        r = f.r
        for t, pos, size, _ in msgs:
            if t == 0x0001:
                self.shape, _ = parse_dataspace(r, pos)
            elif t == 0x0003:
                self.dt = parse_datatype(r.b, pos)
            elif t == 0x0008:
                self.layout = self._parse_layout(pos)
            elif t == 0x000B:
                self.filters = self._parse_filters(pos)

    @property
    def dtype(self):
        return self.dt.dtype if self.dt is not None else None

    # ---- messages
    def _parse_layout(self, pos):
        r = self.f.r
        b = r.b
        ver = b[pos]
        if ver == 3:
            cls = b[pos + 1]
            if cls == 0:
                n = r.u(pos + 2, 2)
                return ("compact", pos + 4, n)
            if cls == 1:
                return ("contiguous", r.off(pos + 2), r.length(pos + 2 + r.O))
            if cls == 2:
                nd = b[pos + 2]
                addr = r.off(pos + 3)
                dims = [r.u(pos + 3 + r.O + 4 * i, 4) for i in range(nd)]
                return ("chunked", addr, dims)
            raise H5Error(f"data layout class {cls}")
        if ver in (1, 2):
            nd, cls = b[pos + 1], b[pos + 2]
            p = pos + 8
            addr = None
            if cls != 0:
                addr = r.off(p)
                p += r.O
            dims = [r.u(p + 4 * i, 4) for i in range(nd)]
            p += 4 * nd
            if cls == 0:
                n = r.u(p, 4)
                return ("compact", p + 4, n)
            if cls == 1:
                return ("contiguous", addr, None)
            return ("chunked", addr, dims)
        raise H5Error(f"data layout message version {ver} (written with libver='latest'?)")

    def _parse_filters(self, pos):
        r = self.f.r
        b = r.b
        ver, n = b[pos], b[pos + 1]
        out = []
        p = pos + (8 if ver == 1 else 2)
        for _ in range(n):
            fid = r.u(p, 2)
            p += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen = r.u(p, 2)
                p += 2
            p += 2  # flags
            ncd = r.u(p, 2)
            p += 2
            p += (nlen + 7) // 8 * 8 if ver == 1 else nlen
            cd = [r.u(p + 4 * i, 4) for i in range(ncd)]
            p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out

    # ---- data
    def _defilter(self, raw, mask=0):
        for i in range(len(self.filters) - 1, -1, -1):
            if mask & (1 << i):
                continue
            fid, cd = self.filters[i]
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:
                es = cd[0] if cd else self.dt.size
                a = np.frombuffer(raw, dtype=np.uint8)
                n = a.size // es
                if n * es == a.size:  # one transposing copy, no bytes object in between
                    raw = np.ascontiguousarray(a.reshape(es, n).T).reshape(-1)
                else:
                    raw = a[:n * es].reshape(es, n).T.tobytes() + bytes(a[n * es:])
            elif fid == 3:
                raw = raw[:-4]
            else:
                raise H5Error(f"HDF5 filter id {fid}")
        return raw

    def read(self):
        """The whole dataset as a numpy array (native byte order)."""
        if self.dt is None or self.layout is None:
            raise H5Error(f"{self.name}: not a simple dataset")
        if self.dt.kind not in ("num", "str"):
            raise H5Error(f"{self.name}: datatype class {self.dt.cls} is not supported for data")
        if self.shape is None:
            return np.zeros((0,), dtype=self.dt.dtype.newbyteorder("="))
        dt = self.dt.dtype
        n = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        b = self.f.r.b
        kind = self.layout[0]
        if kind == "compact":
            a = np.frombuffer(b[self.layout[1]:self.layout[1] + n * dt.itemsize], dtype=dt)
        elif kind == "contiguous":
            addr = self.layout[1]
            if addr is None:  # never written: fill value
                a = np.zeros(n, dtype=dt)
            else:
                addr += self.f.base
                a = np.frombuffer(b[addr:addr + n * dt.itemsize], dtype=dt)
        else:
            a = self._read_chunked(dt)
        a = a.reshape(self.shape)
        return a.astype(dt.newbyteorder("="), copy=True) if dt.kind != "S" else a.copy()

    def _read_chunked(self, dt):
        addr, cdims = self.layout[1], self.layout[2]
        rank = len(self.shape)
        chunk = cdims[:rank]
        out = np.zeros(self.shape, dtype=dt)
        if addr is None:
            return out
        def place(entry):
            offs, size, mask, caddr = entry
            raw = bytes(self.f.r.b[caddr + self.f.base:caddr + self.f.base + size])
            raw = self._defilter(raw, mask)
            c = np.frombuffer(raw, dtype=dt, count=int(np.prod(chunk))).reshape(chunk)
            sl_out, sl_c = [], []
            for d in range(rank):
                lo = offs[d]
                hi = min(lo + chunk[d], self.shape[d])
                sl_out.append(slice(lo, hi))
                sl_c.append(slice(0, hi - lo))
            out[tuple(sl_out)] = c[tuple(sl_c)]

        entries = list(self.f._chunk_btree(addr + self.f.base, rank))
        # Chunks land in disjoint parts of `out`, and inflating / unshuffling them
        # releases the GIL: large variables (a CONUS forcing file is 1.6 GB decoded)
        # are read by a few threads.
        if len(entries) > 1 and out.nbytes >= _THREAD_BYTES:
            from concurrent.futures import ThreadPoolExecutor

            with ThreadPoolExecutor(min(8, len(entries), os.cpu_count() or 1)) as pool:
                for _ in pool.map(place, entries):
                    pass
        else:
            for e in entries:
                place(e)
        return out

    # ---- attributes
    @property
    def attrs(self) -> dict:
        if self._attrs is None:
            self._attrs = self.f._attributes(self._msgs)
        return self._attrs

    @property
    def dims(self) -> tuple:
        """netCDF dimension names: the dimension scales DIMENSION_LIST points to;
        a coordinate variable / bare dimension is its own dimension."""
        a = self.attrs
        if a.get("CLASS") == "DIMENSION_SCALE" and len(self.shape or ()) == 1:
            return (self.name,)
        dl = a.get("DIMENSION_LIST")
        if dl is None:
            return tuple(f"dim_{i}" for i in range(len(self.shape or ())))
        by_addr = {v: k for k, v in self.f.links().items()}
        return tuple(by_addr.get(x[0], f"dim_{i}") if x else f"dim_{i}" for i, x in enumerate(dl))

    @property
    def is_bare_dimension(self) -> bool:
        """A netCDF dimension without a coordinate variable."""
        nm = self.attrs.get("NAME")
        return isinstance(nm, str) and nm.startswith("This is a netCDF dimension but not a netCDF variable")


class File:
    def __init__(self, path):
        self._fh = open(path, "rb")
        try:
            self._mm = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        except ValueError:
            self._fh.close()
            raise H5Error(f"{path}: empty file")
        b = self._mm
        pos = 0
        while b[pos:pos + 8] != SIG:  # the superblock may sit at 0, 512, 1024, ...
            pos = 512 if pos == 0 else pos * 2
            if pos + 8 > len(b):
                self.close()
                raise H5Error(f"{path}: not an HDF5 file")
        ver = b[pos + 8]
        self.base = 0
        if ver in (0, 1):
            O, L = b[pos + 13], b[pos + 14]
            self.r = _Reader(b, O, L)
            p = pos + 24 + (4 if ver == 1 else 0)
            self.base = self.r.u(p, O)
            p += 4 * O  # base, free-space, eof, driver info
            # root group symbol table entry: link name offset, object header address
            root = self.r.off(p + O)
        elif ver in (2, 3):
            O, L = b[pos + 9], b[pos + 10]
            self.r = _Reader(b, O, L)
            self.base = self.r.u(pos + 12, O)
            root = self.r.off(pos + 12 + 3 * O)
        else:
            self.close()
            raise H5Error(f"HDF5 superblock version {ver}")
        self.base += pos if self.base == 0 and pos else 0
        self._root_msgs = self._object_header(root + self.base)
        self._links = None
        self._cache = {}

    def close(self):
        try:
            self._mm.close()
        finally:
            self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- object headers
    def _object_header(self, addr):
        """-> list of (type, data position, size, flags)."""
        r = self.r
        b = r.b
        msgs = []
        if b[addr:addr + 4] == b"OHDR":
            flags = b[addr + 5]
            p = addr + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            nsz = 1 << (flags & 3)
            csize = r.u(p, nsz)
            p += nsz
            track = bool(flags & 0x04)
            blocks = [(p, p + csize)]
            while blocks:
                p, end = blocks.pop(0)
                while p + 4 + (2 if track else 0) <= end:
                    t = b[p]
                    size = r.u(p + 1, 2)
                    mflags = b[p + 3]
                    p += 4 + (2 if track else 0)
                    if t == 0x10:
                        caddr, clen = r.off(p), r.length(p + r.O)
                        caddr += self.base
                        if b[caddr:caddr + 4] != b"OCHK":
                            raise H5Error("bad object header continuation")
                        blocks.append((caddr + 4, caddr + clen - 4))
                    elif t != 0:
                        msgs.append((t, p, size, mflags))
                    p += size
            return msgs
        # version 1
        if b[addr] != 1:
            raise H5Error(f"object header version {b[addr]} at {addr}")
        nmsg = r.u(addr + 2, 2)
        hsize = r.u(addr + 8, 4)
        blocks = [(addr + 16, addr + 16 + hsize)]
        while blocks and len(msgs) < 4 * nmsg + 64:
            p, end = blocks.pop(0)
            while p + 8 <= end:
                t = r.u(p, 2)
                size = r.u(p + 2, 2)
                mflags = b[p + 4]
                p += 8
                if t == 0x10:
                    caddr, clen = r.off(p), r.length(p + r.O)
                    blocks.append((caddr + self.base, caddr + self.base + clen))
                elif t != 0:
                    msgs.append((t, p, size, mflags))
                p += size
        return msgs

    # ---- fractal heap: every managed object region, in heap order
    def _heap_blocks(self, addr):
        """Yield (start, end) byte ranges of the object area of every direct block."""
        r = self.r
        b = r.b
        if b[addr:addr + 4] != b"FRHP":
            raise H5Error("bad fractal heap header")
        p = addr + 5
        p += 2  # heap id length
        filt_len = r.u(p, 2)
        p += 2
        flags = b[p]
        p += 1
        p += 4  # max size of managed objects
        p += r.L + r.O  # next huge id, huge btree
        p += r.L + r.O  # free space, free space manager
        p += 4 * r.L    # managed space, allocated, iterator offset, number of managed objects
        p += 4 * r.L    # huge size/number, tiny size/number
        width = r.u(p, 2)
        p += 2
        start_size = r.length(p)
        p += r.L
        max_direct = r.length(p)
        p += r.L
        max_heap_bits = r.u(p, 2)
        p += 2
        p += 2  # starting rows in root indirect block
        root = r.off(p)
        p += r.O
        cur_rows = r.u(p, 2)
        if filt_len:
            raise H5Error("filtered fractal heap")
        if root is None:
            return
        off_bytes = (max_heap_bits + 7) // 8
        checksum = 4 if flags & 2 else 0
        hdr = 5 + r.O + off_bytes + checksum

        def row_size(row):
            return start_size if row < 2 else start_size << (row - 1)

        def direct(a, size):
            a += self.base
            if b[a:a + 4] != b"FHDB":
                raise H5Error("bad fractal heap direct block")
            yield (a + hdr, a + size)

        def indirect(a, nrows):
            a += self.base
            if b[a:a + 4] != b"FHIB":
                raise H5Error("bad fractal heap indirect block")
            q = a + 5 + r.O + off_bytes
            max_direct_rows = 2
            while row_size(max_direct_rows) <= max_direct:
                max_direct_rows += 1
            for row in range(nrows):
                size = row_size(row)
                for _ in range(width):
                    child = r.off(q)
                    q += r.O
                    if child is None:
                        continue
                    if row < max_direct_rows and size <= max_direct:
                        yield from direct(child, size)
                    else:
                        # rows of the child indirect block: it spans `size` bytes of heap
                        k = 0
                        tot = 0
                        while tot < size:
                            tot += width * row_size(k)
                            k += 1
                        yield from indirect(child, k)

        if cur_rows == 0:
            yield from direct(root, start_size)
        else:
            yield from indirect(root, cur_rows)

    # ---- groups
    def _parse_link(self, p):
        """Link message at p -> (name, object header address or None, bytes used)."""
        r = self.r
        b = r.b
        if b[p] != 1:
            return None
        flags = b[p + 1]
        q = p + 2
        ltype = 0
        if flags & 0x08:
            ltype = b[q]
            q += 1
        if flags & 0x04:
            q += 8
        if flags & 0x10:
            q += 1
        nsz = 1 << (flags & 3)
        nlen = r.u(q, nsz)
        q += nsz
        name = bytes(b[q:q + nlen]).decode("utf-8", "replace")
        q += nlen
        if ltype == 0:
            addr = r.off(q)
            q += r.O
        elif ltype == 1:  # soft link
            n = r.u(q, 2)
            q += 2 + n
            addr = None
        else:
            n = r.u(q, 2)
            q += 2 + n
            addr = None
        return name, addr, q - p

    def links(self) -> dict:
        """name -> object header address for the root group."""
        if self._links is not None:
            return self._links
        r = self.r
        b = r.b
        out = {}
        for t, pos, size, _ in self._root_msgs:
            if t == 0x0006:
                lk = self._parse_link(pos)
                if lk and lk[1] is not None:
                    out[lk[0]] = lk[1]
            elif t == 0x0002:  # link info -> dense storage
                flags = b[pos + 1]
                p = pos + 2 + (8 if flags & 1 else 0)
                heap = r.off(p)
                if heap is None:
                    continue
                for lo, hi in self._heap_blocks(heap + self.base):
                    p = lo
                    while p < hi and b[p] == 1:
                        lk = self._parse_link(p)
                        if lk is None or lk[2] <= 0 or not lk[0]:
                            break
                        if lk[1] is not None:
                            out[lk[0]] = lk[1]
                        p += lk[2]
            elif t == 0x0011:  # old-style symbol table
                btree, heap = r.off(pos), r.off(pos + r.O)
                out.update(self._symbol_table(btree + self.base, heap + self.base))
        self._links = out
        return out

    def _symbol_table(self, btree, heap):
        r = self.r
        b = r.b
        if b[heap:heap + 4] != b"HEAP":
            raise H5Error("bad local heap")
        data = r.off(heap + 8 + 2 * r.L) + self.base
        out = {}

        def node(a):
            if b[a:a + 4] == b"TREE":
                level, n = b[a + 5], r.u(a + 6, 2)
                p = a + 8 + 2 * r.O
                for i in range(n):
                    p += r.L  # key
                    node(r.off(p) + self.base)
                    p += r.O
            elif b[a:a + 4] == b"SNOD":
                n = r.u(a + 6, 2)
                p = a + 8
                for i in range(n):
                    noff = r.u(p, r.O)
                    oaddr = r.off(p + r.O)
                    e = b.find(b"\x00", data + noff)
                    out[bytes(b[data + noff:e]).decode("utf-8", "replace")] = oaddr
                    p += 2 * r.O + 4 + 4 + 16
            else:
                raise H5Error("bad group B-tree node")

        node(btree)
        return out

    # ---- chunk index (B-tree v1, node type 1)
    def _chunk_btree(self, addr, rank):
        r = self.r
        b = r.b
        if b[addr:addr + 4] != b"TREE":
            raise H5Error("bad chunk B-tree node")
        level, n = b[addr + 5], r.u(addr + 6, 2)
        p = addr + 8 + 2 * r.O
        ksize = 8 + 8 * (rank + 1)
        for i in range(n):
            size, mask = r.u(p, 4), r.u(p + 4, 4)
            offs = [r.u(p + 8 + 8 * d, 8) for d in range(rank)]
            child = r.off(p + ksize)
            p += ksize + r.O
            if level == 0:
                yield offs, size, mask, child
            else:
                yield from self._chunk_btree(child + self.base, rank)

    # ---- attributes
    def _parse_attribute(self, p):
        """Attribute message at p -> (name, value, bytes used)."""
        r = self.r
        b = r.b
        ver = b[p]
        if ver not in (1, 2, 3):
            return None
        nlen, dlen, slen = r.u(p + 2, 2), r.u(p + 4, 2), r.u(p + 6, 2)
        q = p + 8 + (1 if ver == 3 else 0)
        pad = (lambda n: (n + 7) // 8 * 8) if ver == 1 else (lambda n: n)
        name = bytes(b[q:q + nlen]).split(b"\x00")[0].decode("utf-8", "replace")
        q += pad(nlen)
        try:
            dt = parse_datatype(b, q)
        except H5Error:
            dt = None
        q += pad(dlen)
        shape, _ = parse_dataspace(r, q)
        q += pad(slen)
        n = 0 if shape is None else (int(np.prod(shape, dtype=np.int64)) if shape else 1)
        if dt is None:
            return name, None, None
        nbytes = n * dt.size
        val = None
        if dt.kind in ("num", "str") and dt.dtype is not None:
            a = np.frombuffer(b[q:q + nbytes], dtype=dt.dtype)
            if dt.kind == "str":
                s = [x.split(b"\x00")[0].decode("utf-8", "replace") for x in a.tolist()]
                val = s[0] if shape == () or len(s) == 1 else s
            else:
                a = a.astype(dt.dtype.newbyteorder("="))
                val = a[0] if shape == () else a.reshape(shape)
        elif dt.kind == "vlen_str":
            s = []
            for i in range(n):
                e = q + i * dt.size
                ln = r.u(e, 4)
                gaddr = r.off(e + 4)
                idx = r.u(e + 4 + r.O, 4)
                s.append(self._global_heap_object(gaddr, idx)[:ln].decode("utf-8", "replace")
                         if gaddr is not None else "")
            val = s[0] if shape == () or len(s) == 1 else s
        elif dt.kind == "vlen" and dt.base is not None and dt.base.kind == "ref":
            # DIMENSION_LIST: per dimension, the object addresses of its scales
            val = []
            for i in range(n):
                e = q + i * dt.size
                ln = r.u(e, 4)
                gaddr = r.off(e + 4)
                idx = r.u(e + 4 + r.O, 4)
                raw = self._global_heap_object(gaddr, idx) if gaddr is not None and ln else b""
                val.append([int.from_bytes(raw[j * dt.base.size:j * dt.base.size + r.O], "little")
                            for j in range(ln)])
        return name, val, q + nbytes - p

    def _global_heap_object(self, addr, idx):
        r = self.r
        b = r.b
        addr += self.base
        if b[addr:addr + 4] != b"GCOL":
            raise H5Error("bad global heap collection")
        size = r.length(addr + 8)
        p = addr + 8 + r.L
        end = addr + size
        while p + 8 + r.L <= end:
            i = r.u(p, 2)
            n = r.length(p + 8)
            if i == idx:
                return bytes(b[p + 8 + r.L:p + 8 + r.L + n])
            if i == 0:
                break
            p += 8 + r.L + (n + 7) // 8 * 8
        return b""

    def _attributes(self, msgs) -> dict:
        r = self.r
        b = r.b
        out = {}
        for t, pos, size, _ in msgs:
            if t == 0x000C:
                a = self._parse_attribute(pos)
                if a:
                    out[a[0]] = a[1]
            elif t == 0x0015:  # attribute info -> dense storage
                flags = b[pos + 1]
                p = pos + 2 + (2 if flags & 1 else 0)
                heap = r.off(p)
                if heap is None:
                    continue
                for lo, hi in self._heap_blocks(heap + self.base):
                    p = lo
                    while p + 8 < hi and b[p] in (1, 2, 3):
                        a = self._parse_attribute(p)
                        if a is None or a[2] is None or a[2] <= 0:
                            break
                        out[a[0]] = a[1]
                        p += a[2]
        return out

    # ---- public
    @property
    def attrs(self) -> dict:
        return self._attributes(self._root_msgs)

    def keys(self):
        return list(self.links().keys())

    def __contains__(self, name):
        return name in self.links()

    def __getitem__(self, name) -> Dataset:
        if name not in self._cache:
            addr = self.links()[name]
            self._cache[name] = Dataset(self, name, self._object_header(addr + self.base))
        return self._cache[name]

    def datasets(self) -> dict:
        """Every root-level object that is a dataset."""
        out = {}
        for k in self.keys():
            d = self[k]
            if d.dt is not None and d.layout is not None:
                out[k] = d
        return out
