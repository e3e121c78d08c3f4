"""hdf5_out.H5Writer: netCDF-4 / HDF5 files written by hand (no libhdf5 in the
image).  Checked three ways: (1) read back by hdf5_min, the reader that reads every
netCDF-4 file the reference ships; (2) the message bodies a dataset header is made
of are byte-identical to the ones libhdf5 wrote into the reference's own files,
where both say the same thing; (3) through ModelNetcdfOutput with
`netcdf_output_format: netcdf4`."""
import pathlib as pl
import struct

import numpy as np
import pytest

from pywatershed_b200 import hdf5_min
from pywatershed_b200.hdf5_out import H5Writer

REF = pl.Path("/root/reference/test_data")
needs_ref_nc = pytest.mark.skipif(not (REF / "drb_2yr" / "prcp.nc").exists(), reason="reference files absent")


def _write(path, nd, n, blocks, complevel=4, extra_vars=0):
    w = H5Writer(path, complevel=complevel, threads=3)
    w.setncattr("Description", "pywatershed output data")
    w.setncattr("process class", "PRMSSnow")
    w.createDimension("time", None)
    w.createDimension("nhm_id", n)
    w.createVariable("time", "f4", ("time",), units="days since 1970-01-01 00:00:00")
    cv = w.createVariable("nhm_id", "i4", ("nhm_id",))
    v = w.createVariable("pkwater_equiv", "f8", ("time", "nhm_id"))
    v.setncattr("units", "inches")
    v.setncattr("_FillValue", np.array(9.969209968386869e36))
    iv = w.createVariable("newsnow", "i4", ("time", "nhm_id"))
    iv.setncattr("_FillValue", np.array(-2147483647, dtype=np.int32))
    for k in range(extra_vars):
        w.createVariable(f"extra_{k:03d}", "f8", ("time", "nhm_id"))
    cv[:] = np.arange(1, n + 1, dtype=np.int32)
    rng = np.random.default_rng(5)
    a = rng.normal(size=(nd, n))
    b = rng.integers(0, 2, size=(nd, n)).astype(np.int32)
    t = (3287 + np.arange(nd)).astype(np.float32)
    t0 = 0
    for nb in blocks:
        t1 = min(nd, t0 + nb)
        arrays = {"time": t[t0:t1], "pkwater_equiv": a[t0:t1], "newsnow": b[t0:t1]}
        arrays.update({f"extra_{k:03d}": a[t0:t1] + k for k in range(extra_vars)})
        w.write_record_block(t0, t1, arrays)
        t0 = t1
    assert t0 == nd
    w.close()
    return a, b, t


@pytest.mark.parametrize("nd,n,blocks,complevel", [(130, 765, (50, 1, 79), 4), (70, 12, (70,), 0),
                                                   (4200, 3, (1000,) * 5, 1)])
def test_round_trip_through_the_reader(tmp_path, nd, n, blocks, complevel):
    """One chunk per day and variable, shuffle + deflate; 130 days = a two-level chunk
    B-tree (64 entries per node), 4,200 days = three levels; complevel 0 = shuffle only."""
    p = tmp_path / "x.nc"
    a, b, t = _write(p, nd, n, blocks, complevel)
    with hdf5_min.File(p) as f:
        assert sorted(f.keys()) == ["newsnow", "nhm_id", "pkwater_equiv", "time"]
        assert f.attrs["Description"] == "pywatershed output data" and f.attrs["process class"] == "PRMSSnow"
        d = f["pkwater_equiv"]
        assert d.dims == ("time", "nhm_id") and d.shape == (nd, n)
        np.testing.assert_array_equal(d.read(), a)
        assert d.attrs["units"] == "inches" and d.attrs["_FillValue"][0] == pytest.approx(9.969209968386869e36)
        np.testing.assert_array_equal(f["newsnow"].read(), b)
        assert f["newsnow"].attrs["_FillValue"][0] == -2147483647
        np.testing.assert_array_equal(f["time"].read(), t)
        np.testing.assert_array_equal(f["nhm_id"].read(), np.arange(1, n + 1))
        tm = f["time"].attrs
        assert tm["CLASS"] == "DIMENSION_SCALE" and tm["NAME"] == "time" and tm["_Netcdf4Dimid"] == 0
        assert tm["units"] == "days since 1970-01-01 00:00:00"
        assert f["nhm_id"].attrs["_Netcdf4Dimid"] == 1
    if complevel:
        assert p.stat().st_size < 0.97 * (a.nbytes + b.nbytes) or n < 100


def test_many_variables_span_several_symbol_table_nodes(tmp_path):
    p = tmp_path / "many.nc"
    a, _, _ = _write(p, 5, 4, (5,), extra_vars=150)
    with hdf5_min.File(p) as f:
        assert len(f.keys()) == 154
        np.testing.assert_array_equal(f["extra_149"].read(), a + 149)
        np.testing.assert_array_equal(f["extra_000"].read(), a)


def test_superblock_and_group_structures(tmp_path):
    """The fixed-layout parts, field by field (HDF5 file format specification,
    'Disk Format: Level 0A / 1A / 1C / 1D')."""
    p = tmp_path / "s.nc"
    _write(p, 3, 5, (3,))
    b = p.read_bytes()
    assert b[:8] == b"\x89HDF\r\n\x1a\n"
    assert b[8:16] == bytes([0, 0, 0, 0, 0, 8, 8, 0])           # versions, 8-byte offsets and lengths
    leaf_k, int_k, flags = struct.unpack("<HHI", b[16:24])
    assert (leaf_k, int_k, flags) == (32, 16, 0)
    base, free, eof, drv = struct.unpack("<QQQQ", b[24:56])
    assert base == 0 and free == 2**64 - 1 and drv == 2**64 - 1 and eof == len(b)
    name_off, root, cache, _ = struct.unpack("<QQII", b[56:80])
    assert name_off == 0 and cache == 0 and root % 8 == 0
    ver, nmsg, refc, hsize = struct.unpack("<BxHII", b[root:root + 12])
    assert ver == 1 and refc == 1 and hsize % 8 == 0
    mtype, msize = struct.unpack("<HH", b[root + 16:root + 20])
    assert mtype == 0x11 and msize == 16                        # symbol table message
    tree, heap = struct.unpack("<QQ", b[root + 24:root + 40])
    assert b[tree:tree + 4] == b"TREE" and b[tree + 4] == 0 and b[tree + 5] == 0
    nent, left, right = struct.unpack("<HQQ", b[tree + 6:tree + 24])
    assert nent == 1 and left == right == 2**64 - 1
    assert b[heap:heap + 4] == b"HEAP"
    seg_size, free_off, seg = struct.unpack("<QQQ", b[heap + 8:heap + 32])
    assert free_off == 1 and seg_size % 8 == 0                  # H5HL_FREE_NULL: no free block
    key0, snod, key1 = struct.unpack("<QQQ", b[tree + 24:tree + 48])
    assert key0 == 0 and b[snod:snod + 4] == b"SNOD" and b[snod + 4] == 1
    nsym = struct.unpack("<H", b[snod + 6:snod + 8])[0]
    names = []
    for i in range(nsym):
        off, _ = struct.unpack("<QQ", b[snod + 8 + 40 * i:snod + 24 + 40 * i])
        names.append(b[seg + off:b.index(b"\0", seg + off)].decode())
    assert names == sorted(names) == ["newsnow", "nhm_id", "pkwater_equiv", "time"]
    assert b[seg + key1:b.index(b"\0", seg + key1)].decode() == "time"   # the node's largest name


def _raw_messages(path, name):
    with hdf5_min.File(path) as f:
        d = f[name]
        return [(t, bytes(f.r.b[p:p + s])) for t, p, s, _ in d._msgs]


@needs_ref_nc
def test_message_bodies_equal_what_libhdf5_wrote(tmp_path):
    """Datatype messages (float32 as in drb_2yr/prcp.nc, int32 as in a
    _Netcdf4Dimid), the CLASS / _Netcdf4Dimid attribute messages of a dimension
    scale (hru_1/cbh.nc: version-1 attribute messages) and the variable-length
    reference type of DIMENSION_LIST: the same bytes."""
    from pywatershed_b200 import hdf5_out as ho

    ref = dict((t, m) for t, m in _raw_messages(REF / "drb_2yr" / "prcp.nc", "prcp") if t == 0x3)
    assert ho._dt_num(np.dtype("<f4")) == ref[0x3][:20]
    dl = [m for t, m in _raw_messages(REF / "drb_2yr" / "prcp.nc", "prcp") if t == 0xC and b"DIMENSION_LIST" in m][0]
    assert ho._DT_VLEN_REF in dl
    scale = [m for t, m in _raw_messages(REF / "hru_1" / "cbh.nc", "hru") if t == 0xC]
    cls = [m for m in scale if m[8:13] == b"CLASS"][0]
    mine = ho._attr_value("CLASS", "DIMENSION_SCALE")[8:]       # without the 8-byte message header
    assert mine[:len(cls)] == cls
    dimid = [m for m in scale if m[8:21] == b"_Netcdf4Dimid"][0]
    mine = ho._attr("_Netcdf4Dimid", ho._dt_num(np.dtype("<i4")), ho._space(), struct.pack("<i", 1))[8:]
    assert mine[:len(dimid)] == dimid
    # dataspace of an unlimited 1-D variable, and layout messages (contiguous / chunked), version 3
    sp = [m for t, m in _raw_messages(REF / "hru_1" / "cbh.nc", "time") if t == 0x1][0]
    assert ho._space([0], [None]) == sp


def test_model_output_as_netcdf4(tmp_path, monkeypatch):
    """`netcdf_output_format: netcdf4`: the model's output files (a variable file, an
    integer flag, a soltab table on the `doy` axis, a budget file) as netCDF-4 from
    H5Writer, read back with the HDF5 reader; the engine is a stub."""
    import types

    import scenarios

    import pywatershed_b200 as pwsb
    from pywatershed_b200 import netcdf_out
    from pywatershed_b200.netcdf_in import read_time_variable

    if netcdf_out.nc4_module() is not None:
        pytest.skip("netCDF4 is the backend here")
    monkeypatch.setattr(netcdf_out, "_FORMAT", "classic")  # restored afterwards
    p, f, start = scenarios.load_drb()
    nhru, nseg, nd = 765, 456, 7
    t0 = np.datetime64(start, "s")
    control = pwsb.Control(t0, t0 + np.timedelta64(nd - 1, "D"), np.timedelta64(24, "h"),
                           options={"budget_type": "warn", "calc_method": "cuda",
                                    "netcdf_output_format": "netcdf4"})
    params = pwsb.Parameters(dims={"nhru": nhru, "nsegment": nseg},
                             coords={"nhm_id": np.arange(1, nhru + 1), "nhm_seg": np.arange(1, nseg + 1)},
                             data_vars=p)
    m = pwsb.Model([pwsb.PRMSSolarGeometry, pwsb.PRMSAtmosphere, pwsb.PRMSCanopy, pwsb.PRMSSnow,
                    pwsb.PRMSRunoff, pwsb.PRMSSoilzone, pwsb.PRMSGroundwater, pwsb.PRMSChannel],
                   control=control, parameters=params, find_input_files=False)
    m._engine = types.SimpleNamespace(nhru=nhru, nseg=nseg)
    rng = np.random.default_rng(3)
    snow_terms = [v for comp in m.processes["PRMSSnow"].budget.terms.values() for v in comp]
    blk = {v: rng.normal(size=(nd, nhru)) for v in set(snow_terms) | {"pkwater_equiv"}}
    blk["seg_outflow"] = rng.normal(size=(nd, nseg))
    blk["newsnow"] = rng.integers(0, 2, size=(nd, nhru)).astype(np.int32)
    m._block, m._block_t0 = blk, 0
    out = netcdf_out.ModelNetcdfOutput(m, tmp_path, True, ["pkwater_equiv", "seg_outflow", "newsnow"], None)
    out.write_days(m, 0, 3)
    out.write_days(m, 3, nd)
    out.close()
    assert (tmp_path / "pkwater_equiv.nc").read_bytes()[:8] == b"\x89HDF\r\n\x1a\n"
    for v in ("pkwater_equiv", "seg_outflow", "newsnow"):
        t, a = read_time_variable(tmp_path / f"{v}.nc", v)
        np.testing.assert_array_equal(a, blk[v])
        assert t[0] == t0 and len(t) == nd
    with hdf5_min.File(tmp_path / "seg_outflow.nc") as h:
        assert h["seg_outflow"].dims == ("time", "nhm_seg")
        assert h["seg_outflow"].attrs["units"] == "cfs"
        assert h.attrs["process class"] == "PRMSChannel"
    b = m.processes["PRMSSnow"].budget
    with hdf5_min.File(tmp_path / "PRMSSnow_budget.nc") as h:
        np.testing.assert_allclose(h["inputs_sum"].read(), sum(blk[v] for v in b.terms["inputs"]), atol=1e-12)
        assert h.attrs["Budget basis"] == "unit (unit or global)"
    with pytest.raises(ValueError, match="netcdf_output_format"):
        control.options["netcdf_output_format"] = "hdf4"
        netcdf_out.ModelNetcdfOutput(m, tmp_path / "x", True, ["pkwater_equiv"], None)
