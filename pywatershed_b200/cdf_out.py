"""Streaming writer for classic netCDF files (CDF-2, "64-bit offset"), used for
the model output when the `netCDF4` module is absent (netcdf_out.py).

Why not scipy.io.netcdf_file: in write mode it keeps every record of every
variable in memory and writes the file at close -- a 10-year CONUS run is
3.2 GB per variable.  Here the header is written once, each block of days is
turned into big-endian records and put in place with one `os.pwrite` (both
steps release the GIL, so several files are written by concurrent threads),
and `numrecs` in the header is brought up to date on flush / close.

File layout (the classic format specification, "NetCDF Classic Format",
Unidata): magic `CDF\\x02`, numrecs, dimension list, global attributes,
variable list (name, dim ids, attributes, type, vsize, 64-bit begin), then the
fixed-size variables in order, then the records: record r holds the r-th slab of
every record variable, each padded to 4 bytes.  The same dims, names, dtypes
and attributes as the reference's files (utils/netcdf_utils.py:356-697), so
xarray / netCDF4 / scipy open them like the reference's netCDF-4 files, minus
compression.
"""
from __future__ import annotations

import os
import struct

import numpy as np

NC_CHAR, NC_INT, NC_FLOAT, NC_DOUBLE = 2, 4, 5, 6
_TAG_DIM, _TAG_VAR, _TAG_ATT = 0x0A, 0x0B, 0x0C
_TYPES = {"f8": (NC_DOUBLE, ">f8"), "f4": (NC_FLOAT, ">f4"), "i4": (NC_INT, ">i4")}


def _pad4(n: int) -> int:
    return (n + 3) & ~3


def _name(s: str) -> bytes:
    b = s.encode()
    return struct.pack(">i", len(b)) + b + b"\0" * (_pad4(len(b)) - len(b))


def _att_list(atts: dict) -> bytes:
    if not atts:
        return struct.pack(">ii", 0, 0)
    out = struct.pack(">ii", _TAG_ATT, len(atts))
    for k, v in atts.items():
        if isinstance(v, (str, bytes)):
            b = v.encode() if isinstance(v, str) else v
            out += _name(k) + struct.pack(">ii", NC_CHAR, len(b)) + b + b"\0" * (_pad4(len(b)) - len(b))
        else:
            a = np.atleast_1d(np.asarray(v))
            if a.dtype.kind in "iub":
                t, a = NC_INT, a.astype(">i4")
            elif a.dtype == np.float32:
                t, a = NC_FLOAT, a.astype(">f4")
            else:
                t, a = NC_DOUBLE, a.astype(">f8")
            b = a.tobytes()
            out += _name(k) + struct.pack(">ii", t, a.size) + b + b"\0" * (_pad4(len(b)) - len(b))
    return out


class _Var:
    """Handle with the slice-assignment interface netcdf_out.py uses."""

    def __init__(self, w, name, kind, dims, atts):
        self.w, self.name, self.kind, self.dims, self.atts = w, name, kind, tuple(dims), dict(atts)
        self.nc_type, self.be = _TYPES[kind]
        self.itemsize = np.dtype(self.be).itemsize
        self.record = bool(dims) and dims[0] == w.rec_dim
        self.begin = 0      # file offset (fixed variables) / offset inside a record
        self.row_bytes = 0  # bytes of one slab along the first dimension
        self.vsize = 0

    def setncattr(self, k, v):
        if self.w.header_written:
            raise RuntimeError("attributes must be set before the first write")
        self.atts[k] = v

    def __setitem__(self, key, value):
        self.w.write_rows(self, key, value)


class Cdf2Writer:
    """One classic-format file: dimensions (at most one of them unlimited),
    global attributes, variables declared up front or until the first write."""

    def __init__(self, path):
        self.path = str(path)
        self.fd = os.open(self.path, os.O_CREAT | os.O_TRUNC | os.O_RDWR, 0o644)
        self.dims: dict[str, int | None] = {}
        self.rec_dim = None
        self.gatts: dict = {}
        self.vars: dict[str, _Var] = {}
        self.header_written = False
        self.numrecs = 0
        self.recsize = 0
        self.rec_start = 0

    # ---- definition phase
    def setncattr(self, k, v):
        if self.header_written:
            raise RuntimeError("attributes must be set before the first write")
        self.gatts[k] = v

    def createDimension(self, name, size):
        if self.header_written:
            raise RuntimeError("dimensions must be created before the first write")
        if size is None:
            if self.rec_dim is not None:
                raise ValueError("the classic format has one unlimited dimension")
            self.rec_dim = name
        self.dims[name] = size

    def createVariable(self, name, kind, dims, **atts) -> _Var:
        if self.header_written:
            raise RuntimeError("variables must be created before the first write")
        for d in dims[1:]:
            if self.dims[d] is None:
                raise ValueError("the unlimited dimension must come first")
        v = _Var(self, name, kind, dims, atts)
        self.vars[name] = v
        return v

    def _write_header(self):
        names = list(self.dims)
        fixed = [v for v in self.vars.values() if not v.record]
        recs = [v for v in self.vars.values() if v.record]
        for v in self.vars.values():
            inner = int(np.prod([self.dims[d] for d in v.dims[(1 if v.record else 0):]], dtype=np.int64))
            total = inner * v.itemsize
            v.row_bytes = total if v.record else (
                total // max(self.dims[v.dims[0]], 1) if v.dims else total)
            v.vsize = _pad4(total)
        # a lone record variable is not padded (format specification, "Note on padding")
        if len(recs) == 1:
            recs[0].vsize = recs[0].row_bytes

        def build(begin_of):
            out = b"CDF\x02" + struct.pack(">i", self.numrecs)
            out += struct.pack(">ii", _TAG_DIM, len(names)) if names else struct.pack(">ii", 0, 0)
            for n in names:
                out += _name(n) + struct.pack(">i", 0 if self.dims[n] is None else self.dims[n])
            out += _att_list(self.gatts)
            out += struct.pack(">ii", _TAG_VAR, len(self.vars)) if self.vars else struct.pack(">ii", 0, 0)
            for v in self.vars.values():
                out += _name(v.name) + struct.pack(">i", len(v.dims))
                out += b"".join(struct.pack(">i", names.index(d)) for d in v.dims)
                out += _att_list(v.atts)
                out += struct.pack(">iIq", v.nc_type, min(v.vsize, 0xFFFFFFFF), begin_of(v))
            return out

        hdr_len = len(build(lambda v: 0))
        pos = hdr_len
        for v in fixed:
            v.begin = pos
            pos += v.vsize
        self.rec_start = pos
        off = 0
        for v in recs:
            v.begin = off
            off += v.vsize
        self.recsize = off
        hdr = build(lambda v: v.begin if not v.record else self.rec_start + v.begin)
        assert len(hdr) == hdr_len
        os.pwrite(self.fd, hdr, 0)
        if fixed:  # fixed variables exist on disk even when never written (zeros)
            os.ftruncate(self.fd, max(self.rec_start, os.fstat(self.fd).st_size))
        self.header_written = True

    # ---- data phase
    def write_rows(self, v: _Var, key, value):
        if not self.header_written:
            self._write_header()
        if isinstance(key, tuple):
            if any(k != slice(None) for k in key[1:]):
                raise NotImplementedError("only whole rows are written")
            key = key[0]
        n0 = None if not v.dims else (self.dims[v.dims[0]] if not v.record else None)
        if isinstance(key, slice):
            t0 = key.start or 0
            a = np.asarray(value)
            t1 = key.stop if key.stop is not None else (n0 if n0 is not None else t0 + a.shape[0])
        else:
            t0, t1 = int(key), int(key) + 1
            a = np.asarray(value)[None]
        if a.dtype == np.bool_:
            a = a.astype(np.int32)
        data = np.ascontiguousarray(a, dtype=v.be)  # big-endian copy (releases the GIL)
        if data.size * v.itemsize != (t1 - t0) * v.row_bytes:
            raise ValueError(f"{v.name}[{t0}:{t1}]: got shape {a.shape}")
        if not v.record:
            os.pwrite(self.fd, data.data, v.begin + t0 * v.row_bytes)
            return
        self._write_records(t0, t1, {v: data})

    def write_record_block(self, t0, t1, arrays: dict):
        """Rows [t0, t1) of several record variables at once: when they are all of
        the file's record variables this is a single contiguous write."""
        if not self.header_written:
            self._write_header()
        n = t1 - t0
        recs = [v for v in self.vars.values() if v.record]
        for name, a in arrays.items():
            if np.asarray(a).size * self.vars[name].itemsize != n * self.vars[name].row_bytes:
                raise ValueError(f"{name}[{t0}:{t1}]: got shape {np.asarray(a).shape}")
        if len(recs) > 1 and set(arrays) == {v.name for v in recs}:
            # convert (cast + byte order) straight into the records: one pass, one write
            buf = np.empty((n, self.recsize), dtype=np.uint8)
            for v in recs:
                dst = buf[:, v.begin:v.begin + v.row_bytes].view(v.be)
                np.copyto(dst, np.asarray(arrays[v.name]).reshape(n, -1), casting="unsafe")
                if v.vsize > v.row_bytes:
                    buf[:, v.begin + v.row_bytes:v.begin + v.vsize] = 0
            os.pwrite(self.fd, buf.data, self.rec_start + t0 * self.recsize)
            self.numrecs = max(self.numrecs, t1)
            return
        conv = {}
        for name, a in arrays.items():
            v = self.vars[name]
            a = np.asarray(a)
            if a.dtype == np.bool_:
                a = a.astype(np.int32)
            conv[v] = np.ascontiguousarray(a, dtype=v.be)
        self._write_records(t0, t1, conv)

    def _write_records(self, t0, t1, conv: dict):
        n = t1 - t0
        recs = [v for v in self.vars.values() if v.record]
        base = self.rec_start + t0 * self.recsize
        if len(conv) == len(recs):
            if len(recs) == 1:
                (v, d), = conv.items()
                os.pwrite(self.fd, d.data, base)
            else:
                buf = np.zeros((n, self.recsize), dtype=np.uint8)
                for v, d in conv.items():
                    buf[:, v.begin:v.begin + v.row_bytes] = d.reshape(n, -1).view(np.uint8).reshape(n, v.row_bytes)
                os.pwrite(self.fd, buf.data, base)
        else:  # some of the record variables: row by row, the others stay as they are
            for v, d in conv.items():
                rows = d.reshape(n, -1).view(np.uint8).reshape(n, v.row_bytes)
                for r in range(n):
                    os.pwrite(self.fd, rows[r].data, base + r * self.recsize + v.begin)
        if t1 > self.numrecs:
            self.numrecs = t1

    def flush(self):
        if not self.header_written:
            self._write_header()
        os.pwrite(self.fd, struct.pack(">i", self.numrecs), 4)
        # the file covers every record up to numrecs even if some were skipped
        end = self.rec_start + self.numrecs * self.recsize
        if os.fstat(self.fd).st_size < end:
            os.ftruncate(self.fd, end)

    def close(self):
        if self.fd is not None:
            self.flush()
            os.close(self.fd)
            self.fd = None
