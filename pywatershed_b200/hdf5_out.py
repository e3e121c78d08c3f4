"""A minimal streaming netCDF-4 (HDF5) writer for the model output.

The reference writes its results with the netCDF4 module: one `[time, unit]`
variable per file, zlib level 4, shuffle, chunks of one day
(pywatershed/utils/netcdf_utils.py:405-630).  That module (and libhdf5) is not part
of this image, so this file writes the same thing by hand, using only the
oldest structures of the HDF5 file format specification (the ones libhdf5 writes
by default and every reader understands):

    superblock v0 . object headers v1 . groups as symbol tables (B-tree v1 +
    local heap + SNOD) . dataspace v1 . datatype v1 . fill value v2 . layout v3
    (contiguous / chunked with a B-tree v1 chunk index) . filter pipeline v1
    (shuffle + deflate) . attribute v1 . global heap (variable-length data)

plus the netCDF-4 conventions on top (netCDF-C libhdf5/hdf5open.c): one dimension
scale per dimension (CLASS = "DIMENSION_SCALE", NAME, _Netcdf4Dimid), a
DIMENSION_LIST of object references on every variable, _NCProperties on the
root group.  (REFERENCE_LIST, the back pointer from a scale to its users that
only H5DSdetach_scale needs, is not written.)

Streaming: chunks are compressed (a few threads; zlib releases the GIL) and
appended as the blocks of days arrive; all metadata is written at close(),
behind the data, and the superblock last.  Same calling convention as
cdf_out.Cdf2Writer.

No libhdf5 in the build image can check these files: they are read back by
hdf5_min.py, which reads every netCDF-4 file the reference ships
(tests/test_host_layer.py), and compared structure by structure with files
libhdf5 wrote (tests/test_hdf5_out.py).  Opt-in (control option
`netcdf_output_format: netcdf4`); the default without the netCDF4 module stays
the classic streaming writer.
"""
from __future__ import annotations

import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
_K_LEAF, _K_INT, _K_CHUNK = 32, 16, 32   # symbol-table leaf / internal K (superblock), chunk B-tree K (v0: fixed 32)
_KINDS = {"f8": np.dtype("<f8"), "f4": np.dtype("<f4"), "i4": np.dtype("<i4"), "i8": np.dtype("<i8")}
_FILL = {"f8": 9.969209968386869e36, "f4": 9.969209968386869e36, "i4": -2147483647, "i8": -9223372036854775806}


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


# ---------------------------------------------------------------- datatypes / dataspaces
def _dt_num(dt: np.dtype) -> bytes:
    if dt.kind == "f":
        exp_bits, man_bits, bias = (11, 52, 1023) if dt.itemsize == 8 else (8, 23, 127)
        return struct.pack("<B3BI", 0x11, 0x20, dt.itemsize * 8 - 1, 0, dt.itemsize) + struct.pack(
            "<HHBBBBI", 0, dt.itemsize * 8, man_bits, exp_bits, 0, man_bits, bias)
    return struct.pack("<B3BI", 0x10, 0x08, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)


def _dt_string(n: int) -> bytes:      # fixed length, null terminated, ASCII
    return struct.pack("<B3BI", 0x13, 0x00, 0, 0, n)


_DT_REF = struct.pack("<B3BI", 0x17, 0x00, 0, 0, 8)                       # object reference
_DT_VLEN_REF = struct.pack("<B3BI", 0x19, 0x00, 0, 0, 16) + _DT_REF       # variable-length sequence of references


def _space(dims=None, maxdims=None) -> bytes:
    if dims is None:                  # scalar
        return struct.pack("<BBB5x", 1, 0, 0)
    flag = 1 if maxdims is not None else 0
    b = struct.pack("<BBB5x", 1, len(dims), flag) + b"".join(struct.pack("<Q", d) for d in dims)
    if maxdims is not None:
        b += b"".join(struct.pack("<Q", UNDEF if m is None else m) for m in maxdims)
    return b


def _msg(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _attr(name: str, dt: bytes, space: bytes, data: bytes) -> bytes:
    nm = name.encode() + b"\0"
    return _msg(0x000C, struct.pack("<BxHHH", 1, len(nm), len(dt), len(space)) + _pad8(nm) + _pad8(dt)
                + _pad8(space) + data)


def _attr_value(name: str, value) -> bytes:
    """An attribute as netCDF-C writes it: text as a fixed-length string in a
    scalar space, numbers as 1-D arrays."""
    if isinstance(value, (str, bytes)):
        v = value.encode() if isinstance(value, str) else value
        v = v + b"\0"
        return _attr(name, _dt_string(len(v)), _space(), v)
    a = np.atleast_1d(np.asarray(value))
    if a.dtype.kind == "f":
        a = a.astype("<f8" if a.dtype.itemsize == 8 else "<f4")
    elif a.dtype.kind in "iub":
        a = a.astype("<i4" if a.dtype.itemsize <= 4 else "<i8")
    else:
        return _attr_value(name, str(value))
    return _attr(name, _dt_num(a.dtype), _space([a.size]), a.tobytes())


def _object_header(msgs: list) -> bytes:
    body = b"".join(msgs)
    return struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body


class _Var:
    def __init__(self, w, name, kind, dims, atts):
        self.w, self.name, self.kind, self.dims = w, name, kind, tuple(dims)
        self.dtype = _KINDS[kind]
        self.atts = dict(atts)
        self.record = bool(dims) and w.dims[dims[0]] is None
        self.shape_rest = tuple(w.dims[d] for d in dims[1:]) if self.record else tuple(w.dims[d] for d in dims)
        self.chunks = []      # (first row of the chunk, address, bytes) for record variables
        self.data = None      # whole array of a fixed-size variable
        self.addr = None

    def setncattr(self, k, v):
        self.atts[k] = v

    def __setitem__(self, key, value):
        self.w.write_rows(self, key, value)


class H5Writer:
    """One netCDF-4 file; see the module docstring."""

    rec_dim = None

    def __init__(self, path, complevel: int = 4, shuffle: bool = True, threads: int = None):
        if threads is None:
            import os

            threads = max(1, min(16, os.cpu_count() or 1))
        self.f = open(path, "wb")
        self.f.write(b"\0" * 2048)            # superblock + room to spare; data starts 2 KiB in
        self.pos = 2048
        self.dims, self.vars, self.atts = {}, {}, {}
        self.complevel, self.shuffle = int(complevel), bool(shuffle)
        self.nrec = 0
        self._pool = ThreadPoolExecutor(max(1, threads)) if threads > 1 else None

    # ---- the calling convention of cdf_out.Cdf2Writer
    def setncattr(self, k, v):
        self.atts[k] = v

    def createDimension(self, name, size):
        self.dims[name] = None if size is None else int(size)
        if size is None:
            self.rec_dim = name

    def createVariable(self, name, kind, dims, **atts) -> _Var:
        v = _Var(self, name, kind, dims, atts)
        self.vars[name] = v
        return v

    def _append(self, b: bytes) -> int:
        addr = self.pos
        self.f.write(b)
        self.pos += len(b)
        pad = -self.pos % 8
        if pad:
            self.f.write(b"\0" * pad)
            self.pos += pad
        return addr

    def _filter(self, v: _Var, a: np.ndarray) -> bytes:
        if self.shuffle and v.dtype.itemsize > 1:   # bytes of equal significance together: one transposing copy
            raw = np.ascontiguousarray(a.reshape(-1).view(np.uint8).reshape(-1, v.dtype.itemsize).T).data
        else:
            raw = a.tobytes()
        return zlib.compress(raw, self.complevel) if self.complevel > 0 else bytes(raw)

    def write_rows(self, v: _Var, key, value):
        a = np.ascontiguousarray(np.asarray(value), dtype=v.dtype)
        if not v.record:
            v.data = a.reshape(v.shape_rest)
            return
        t0 = key.start if isinstance(key, slice) else key[0].start
        self.write_record_block(t0, t0 + a.reshape((-1,) + v.shape_rest).shape[0], {v.name: a})

    def write_record_block(self, t0, t1, arrays: dict):
        """Days [t0, t1) of the record variables in `arrays`: one chunk per day and
        variable (the reference's chunksizes=(1, n)), filtered by a few threads."""
        jobs = []
        for name, arr in arrays.items():
            v = self.vars[name]
            a = np.ascontiguousarray(np.asarray(arr), dtype=v.dtype).reshape((t1 - t0,) + v.shape_rest)
            if not v.shape_rest:          # the record coordinate itself (time): kept, written at close
                buf = getattr(v, "_rec1d", None)
                if buf is None or buf.size < t1:
                    nb = np.zeros(max(t1, 2 * (0 if buf is None else buf.size), 64), dtype=v.dtype)
                    if buf is not None:
                        nb[:buf.size] = buf
                    v._rec1d = buf = nb
                buf[t0:t1] = a
                continue
            jobs.extend((v, t0 + j, a[j]) for j in range(t1 - t0))
        run = (lambda j: (j[0], j[1], self._filter(j[0], j[2])))
        done = self._pool.map(run, jobs) if self._pool is not None and len(jobs) > 1 else map(run, jobs)
        for v, t, blob in done:
            v.chunks.append((t, self._append(blob), len(blob)))
        self.nrec = max(self.nrec, t1)

    # ---- metadata, at close
    def _chunk_btree(self, entries, rank, upper) -> int:
        """B-tree v1 (node type 1) over `entries` = [(offsets, address, nbytes)]
        sorted by offsets; returns the root's address."""
        def key(nbytes, offs):
            return struct.pack("<II", nbytes, 0) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", 0)

        level, nodes = 0, [(e[0], key(e[2], e[0]), e[1]) for e in entries]   # (first offsets, first key, child address)
        cap = 2 * _K_CHUNK
        while True:
            groups = [nodes[i:i + cap] for i in range(0, len(nodes), cap)] or [[]]
            out = []
            for gi, g in enumerate(groups):
                last = key(0, upper) if gi == len(groups) - 1 else groups[gi + 1][0][1]
                body = b"".join(k + struct.pack("<Q", child) for _, k, child in g) + last
                size = 24 + (cap + 1) * (16 + 8 * rank) + cap * 8   # keys: 8 + (rank + 1) x 8 bytes
                node = b"TREE" + struct.pack("<BBHQQ", 1, level, len(g), UNDEF, UNDEF) + body
                addr = self._append(node + b"\0" * (size - len(node)))
                out.append((g[0][0] if g else upper, g[0][1] if g else last, addr))
            if len(out) == 1:
                return out[0][2]
            nodes, level = out, level + 1

    def _dataset_header(self, v: _Var, dim_addr: dict, gheap) -> bytes:
        rank = len(v.dims)
        msgs = []
        if v.record:
            shape = (self.nrec,) + v.shape_rest
            msgs.append(_msg(0x0001, _space(shape, (None,) + v.shape_rest)))
        else:
            shape = v.shape_rest
            msgs.append(_msg(0x0001, _space(shape, shape)))
        msgs.append(_msg(0x0003, _dt_num(v.dtype), flags=1))
        fill = np.array(v.atts.get("_FillValue", _FILL[v.kind]), dtype=v.dtype).tobytes()
        if v.record:
            chunk = (1,) + v.shape_rest if v.shape_rest else (512,)
            msgs.append(_msg(0x0005, struct.pack("<BBBBI", 2, 3, 2, 1, len(fill)) + fill))   # incremental allocation
            entries = []
            filtered = bool(v.shape_rest) and (self.complevel > 0 or self.shuffle)
            if v.shape_rest:
                entries = [((t,) + (0,) * (rank - 1), addr, nb) for t, addr, nb in sorted(v.chunks)]
            else:
                buf = getattr(v, "_rec1d", np.zeros(0, dtype=v.dtype))
                for c0 in range(0, self.nrec, 512):
                    part = np.zeros(512, dtype=v.dtype)
                    n = min(512, self.nrec - c0)
                    part[:n] = buf[c0:c0 + n]
                    blob = part.tobytes()
                    entries.append(((c0,), self._append(blob), len(blob)))
            upper = (((self.nrec + chunk[0] - 1) // chunk[0]) * chunk[0],) + (0,) * (rank - 1)
            root = self._chunk_btree(entries, rank, upper)
            if filtered:
                filt = b""
                if self.shuffle:
                    filt += struct.pack("<HHHH", 2, 0, 1, 1) + struct.pack("<I4x", v.dtype.itemsize)
                if self.complevel > 0:
                    filt += struct.pack("<HHHH", 1, 0, 1, 1) + struct.pack("<I4x", self.complevel)
                msgs.append(_msg(0x000B, struct.pack("<BB6x", 1, int(self.shuffle) + int(self.complevel > 0)) + filt))
            msgs.append(_msg(0x0008, struct.pack("<BBBQ", 3, 2, rank + 1, root)
                             + b"".join(struct.pack("<I", c) for c in chunk) + struct.pack("<I", v.dtype.itemsize)))
        else:
            data = np.zeros(shape, dtype=v.dtype) if v.data is None else v.data
            blob = data.tobytes()
            msgs.append(_msg(0x0005, struct.pack("<BBBBI", 2, 2, 2, 1, len(fill)) + fill))   # late allocation
            msgs.append(_msg(0x0008, struct.pack("<BBQQ", 3, 1, self._append(blob), len(blob))))
        # netCDF-4: dimension scales and the variable's list of them
        is_scale = v.name in self.dims and v.dims == (v.name,)
        if is_scale:
            msgs.append(_attr_value("CLASS", "DIMENSION_SCALE"))
            msgs.append(_attr_value("NAME", v.name if not getattr(v, "bare", False) else
                                    "This is a netCDF dimension but not a netCDF variable.%10d" % (
                                        self.nrec if self.dims[v.name] is None else self.dims[v.name])))
            msgs.append(_attr("_Netcdf4Dimid", _dt_num(np.dtype("<i4")), _space(),
                              struct.pack("<i", list(self.dims).index(v.name))))
        elif rank:
            refs = b"".join(struct.pack("<IQI", 1, gheap["addr"], gheap["index"][d]) for d in v.dims)
            msgs.append(_attr("DIMENSION_LIST", _DT_VLEN_REF, _space([rank]), refs))
        for k, val in v.atts.items():
            if k == "_FillValue":
                msgs.append(_attr("_FillValue", _dt_num(v.dtype), _space([1]), fill))
            elif val is not None:
                msgs.append(_attr_value(k, val))
        return _object_header(msgs)

    def close(self):
        if self.f is None:
            return
        if self._pool is not None:
            self._pool.shutdown()
        # every dimension is a dataset: its coordinate variable when one exists, else a
        # bare scale (netCDF-C names those "This is a netCDF dimension but not a netCDF variable.")
        for d, n in self.dims.items():
            if d not in self.vars:
                v = self.createVariable(d, "f4", (d,))
                v.bare = True
        names = sorted(self.vars, key=lambda s: s.encode())
        # object headers need each other's addresses (DIMENSION_LIST holds references to the scales):
        # the scales go first, their addresses into a global heap collection, then the rest
        scales = [n for n in names if n in self.dims]
        others = [n for n in names if n not in self.dims]
        addr = {}
        for n in scales:
            addr[n] = self._append(self._dataset_header(self.vars[n], {}, None))
        objs = b""
        index = {}
        for i, d in enumerate(self.dims, start=1):
            index[d] = i
            objs += struct.pack("<HH4xQ", i, 1, 8) + struct.pack("<Q", addr[d])
        size = max(4096, 16 + len(objs) + 16)
        free = size - 16 - len(objs)
        gcol = b"GCOL" + struct.pack("<B3xQ", 1, size) + objs + struct.pack("<HH4xQ", 0, 0, free) + b"\0" * (free - 16)
        gheap = {"addr": self._append(gcol), "index": index}
        for n in others:
            addr[n] = self._append(self._dataset_header(self.vars[n], addr, gheap))
        # the root group: local heap of names, symbol table nodes, B-tree, object header
        heap = b"\0" * 8
        off = {}
        for n in names:
            off[n] = len(heap)
            heap += _pad8(n.encode() + b"\0")
        heap_data = self._append(heap)
        heap_addr = self._append(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), 1, heap_data))
        cap = 2 * _K_LEAF
        groups = [names[i:i + cap] for i in range(0, len(names), cap)] or [[]]
        assert len(groups) <= 2 * _K_INT, "too many variables for one group B-tree node"
        keys, children = [0], []
        for g in groups:
            ent = b"".join(struct.pack("<QQII16x", off[n], addr[n], 0, 0) for n in g)
            node = b"SNOD" + struct.pack("<BxH", 1, len(g)) + ent
            children.append(self._append(node + b"\0" * (8 + cap * 40 - len(node))))
            keys.append(off[g[-1]] if g else 0)
        body = b"".join(struct.pack("<QQ", k, c) for k, c in zip(keys, children)) + struct.pack("<Q", keys[-1])
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(children), UNDEF, UNDEF) + body
        tree_addr = self._append(tree + b"\0" * (24 + (2 * _K_INT + 1) * 8 + 2 * _K_INT * 8 - len(tree)))
        root_msgs = [_msg(0x0011, struct.pack("<QQ", tree_addr, heap_addr))]
        atts = dict(self.atts)
        atts.setdefault("_NCProperties", "version=2,pywatershed_b200=1,hdf5_layout=1.8")
        for k, val in atts.items():
            root_msgs.append(_attr_value(k, val))
        root_addr = self._append(_object_header(root_msgs))
        eof = self.pos
        sb = (b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _K_LEAF, _K_INT, 0)
              + struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
              + struct.pack("<QQII", 0, root_addr, 0, 0) + b"\0" * 16)  # (nothing cached in the entry)
        self.f.seek(0)
        self.f.write(sb)
        self.f.close()
        self.f = None

    def flush(self):
        self.f.flush()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
