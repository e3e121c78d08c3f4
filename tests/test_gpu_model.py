"""The reference-facing layer on the GPU: pywatershed_b200.Model driven the way
the reference's own tests drive pws.Model (autotest/test_model.py,
autotest/test_netcdf_output.py:155-198), compared with the oracle and with
the reference's golden numbers.  GPU only; everything goes through the C ABI.
"""
import numpy as np
import pytest

import parity
import prms_oracle as po
import scenarios

pytestmark = pytest.mark.gpu

NHM = ["PRMSSolarGeometry", "PRMSAtmosphere", "PRMSCanopy", "PRMSSnow", "PRMSRunoff",
       "PRMSSoilzone", "PRMSGroundwater", "PRMSChannel"]


def _control(start, nd, **opts):
    from pywatershed_b200 import Control

    t0 = np.datetime64(start, "s")
    o = {"budget_type": "error", "calc_method": "cuda"}
    o.update(opts)
    return Control(t0, t0 + np.timedelta64(nd - 1, "D"), np.timedelta64(24, "h"), options=o)


def _parameters(p):
    from pywatershed_b200 import Parameters

    nhru, nseg = p["hru_area"].shape[0], p["tosegment"].shape[0]
    coords = {"nhm_id": p.get("nhm_id", np.arange(1, nhru + 1)),
              "nhm_seg": p.get("nhm_seg", np.arange(1, nseg + 1))}
    data = {k: v for k, v in p.items() if k not in coords}
    return Parameters(dims={"nhru": nhru, "nsegment": nseg, "nmonth": 12, "ndoy": 366},
                      coords=coords, data_vars=data)


def _model(p, f, start, nd, procs=NHM, **opts):
    import pywatershed_b200 as pwsb

    control = _control(start, nd, **opts)
    m = pwsb.Model([getattr(pwsb, n) for n in procs], control=control,
                   parameters=_parameters(p), find_input_files=False)
    atm = m.processes["PRMSAtmosphere"]
    for k in ("prcp", "tmax", "tmin"):
        atm.set_input_to_adapter(k, f[k][:nd])
    return m


def test_model_run_matches_oracle_and_reference_known_answers(drb):
    p, f, start = drb
    nd = 731
    m = _model(p, f, start, nd)
    m.run(finalize=False)
    assert m.control.itime_step == nd - 1
    # the reference's last-day known answers
    assert m.processes["PRMSChannel"]["seg_outflow"].sum() == pytest.approx(214752.70363439058, rel=1e-9)
    assert m.processes["PRMSSnow"]["pkwater_equiv"].sum() == pytest.approx(186.00889550764475, rel=1e-6)
    for name in NHM[2:]:
        b = m.processes[name].budget
        assert b._zero_sum, name
        np.testing.assert_allclose(b._inputs_sum, b._outputs_sum + b._storage_changes_sum,
                                   rtol=1e-5, atol=1e-5)
    # all-time Atmosphere arrays hold the whole run, like the reference's
    atm = m.processes["PRMSAtmosphere"]
    assert atm["tmaxf"].data.shape == (nd, 765)
    o = po.Oracle(p, start=start)
    ref, _ = o.run(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd], ["tmaxf", "swrad", "potet"], ())
    for k in ("tmaxf", "swrad", "potet"):
        np.testing.assert_allclose(atm[k].data, ref[k], rtol=1e-9, atol=1e-10)
    m.finalize()


def test_stepwise_advance_calculate_equals_run(drb):
    """autotest/test_netcdf_output.py:155-198 drives advance/calculate/output
    one day at a time and reads state in between; time blocking must not show."""
    p, f, start = drb
    nd = 40
    m = _model(p, f, start, nd, block_days=16)
    ids = {}
    o = po.Oracle(p, start=start)
    names = ["pkwater_equiv", "soil_moist", "sroff", "pptmix", "gwres_stor", "hru_actet"]
    ref, _ = o.run(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd], names, ["seg_outflow"])
    owner = {"pkwater_equiv": "PRMSSnow", "soil_moist": "PRMSSoilzone", "sroff": "PRMSRunoff",
             "pptmix": "PRMSAtmosphere", "gwres_stor": "PRMSGroundwater", "hru_actet": "PRMSSoilzone"}
    for it in range(nd):
        m.advance()
        m.calculate()
        assert m.control.itime_step == it
        for k in names:
            v = m.processes[owner[k]][k]
            cur = v.current if hasattr(v, "current") else v
            np.testing.assert_allclose(np.asarray(cur, dtype=np.float64), ref[k][it],
                                       rtol=1e-9, atol=1e-10, err_msg=f"{k} day {it}")
            # stable identities (autotest/test_prms_atmosphere.py:82,145)
            ids.setdefault(k, id(cur))
            assert ids[k] == id(cur)
        np.testing.assert_allclose(m.processes["PRMSChannel"]["seg_outflow"], ref["seg_outflow"][it],
                                   rtol=1e-9, atol=1e-10)
        assert m.processes["PRMSSnow"].budget._zero_sum
    with pytest.raises(ValueError, match="End of time"):
        m.advance()


def test_netcdf_output_files(drb, tmp_path):
    """Same file names / dims / coordinate variables as the reference's
    NetCdfWrite (utils/netcdf_utils.py:405-630); values equal the in-memory run."""
    p, f, start = drb
    nd = 20
    out_vars = ["pkwater_equiv", "seg_outflow", "tmaxf", "newsnow", "soltab_potsw"]
    m = _model(p, f, start, nd, netcdf_output_var_names=out_vars)
    m.initialize_netcdf(tmp_path)
    m.run(finalize=True)
    for v in out_vars:
        assert (tmp_path / f"{v}.nc").exists(), v
    assert (tmp_path / "PRMSSnow_budget.nc").exists()
    o = po.Oracle(p, start=start)
    ref, _ = o.run(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd],
                   ["pkwater_equiv", "tmaxf", "newsnow"], ["seg_outflow"])
    from pywatershed_b200.netcdf_out import nc4_module

    nc4 = nc4_module()
    if nc4 is not None:
        def read(path, var):
            with nc4.Dataset(path) as ds:
                return (np.asarray(ds.variables[var][:]), list(ds.variables[var].dimensions),
                        np.asarray(ds.variables["time"][:]) if "time" in ds.variables else None)
    else:
        from scipy.io import netcdf_file

        def read(path, var):
            with netcdf_file(str(path), "r", mmap=False) as ds:
                return (np.array(ds.variables[var][:]), list(ds.variables[var].dimensions),
                        np.array(ds.variables["time"][:]) if "time" in ds.variables else None)
    a, dims, t = read(tmp_path / "pkwater_equiv.nc", "pkwater_equiv")
    assert dims == ["time", "nhm_id"] and a.shape == (nd, 765)
    np.testing.assert_allclose(a, ref["pkwater_equiv"], rtol=1e-9, atol=1e-10)
    day0 = (np.datetime64(start, "D") - np.datetime64("1970-01-01", "D")).astype(int)
    np.testing.assert_array_equal(t, (day0 + np.arange(nd)).astype(np.float32))
    a, dims, _ = read(tmp_path / "seg_outflow.nc", "seg_outflow")
    assert dims == ["time", "nhm_seg"] and a.shape == (nd, 456)
    np.testing.assert_allclose(a, ref["seg_outflow"], rtol=1e-9, atol=1e-10)
    a, _, _ = read(tmp_path / "newsnow.nc", "newsnow")
    np.testing.assert_array_equal(a, ref["newsnow"])
    a, dims, _ = read(tmp_path / "soltab_potsw.nc", "soltab_potsw")
    assert a.shape == (366, 765)


def test_netcdf4_output_files_from_our_writer(drb, tmp_path, monkeypatch):
    """`netcdf_output_format: netcdf4` without the netCDF4 module: the same files as
    netCDF-4 / HDF5 from hdf5_out.H5Writer (zlib 4, shuffle, a chunk per day), written
    block by block on the writer thread; read back with the HDF5 reader."""
    from pywatershed_b200 import hdf5_min, netcdf_out

    if netcdf_out.nc4_module() is not None:
        pytest.skip("netCDF4 is the backend here")
    monkeypatch.setattr(netcdf_out, "_FORMAT", "classic")
    p, f, start = drb
    nd = 40
    out_vars = ["pkwater_equiv", "seg_outflow", "newsnow", "soltab_potsw"]
    m = _model(p, f, start, nd, netcdf_output_var_names=out_vars, netcdf_output_format="netcdf4")
    m.block_days = 16
    m.initialize_netcdf(tmp_path)
    m.run(finalize=True)
    ref, _ = po.Oracle(p, start=start).run(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd],
                                           ["pkwater_equiv", "newsnow"], ["seg_outflow"])
    with hdf5_min.File(tmp_path / "pkwater_equiv.nc") as h:
        d = h["pkwater_equiv"]
        assert d.dims == ("time", "nhm_id") and d.shape == (nd, 765)
        np.testing.assert_allclose(d.read(), ref["pkwater_equiv"], rtol=1e-9, atol=1e-10)
        day0 = (np.datetime64(start, "D") - np.datetime64("1970-01-01", "D")).astype(int)
        np.testing.assert_array_equal(h["time"].read(), (day0 + np.arange(nd)).astype(np.float32))
        assert d.attrs["units"] == "inches" and h["time"].attrs["units"] == "days since 1970-01-01 00:00:00"
    with hdf5_min.File(tmp_path / "seg_outflow.nc") as h:
        np.testing.assert_allclose(h["seg_outflow"].read(), ref["seg_outflow"], rtol=1e-9, atol=1e-10)
        assert h["seg_outflow"].dims == ("time", "nhm_seg")
    with hdf5_min.File(tmp_path / "newsnow.nc") as h:
        np.testing.assert_array_equal(h["newsnow"].read(), ref["newsnow"])
    with hdf5_min.File(tmp_path / "soltab_potsw.nc") as h:
        assert h["soltab_potsw"].shape == (366, 765) and h["soltab_potsw"].dims == ("doy", "nhm_id")
        np.testing.assert_array_equal(h["doy"].read(), np.arange(1, 367))
    assert (tmp_path / "PRMSSnow_budget.nc").read_bytes()[:4] == b"\x89HDF"


def test_channel_only_model(drb):
    """examples/05_parameter_ensemble.ipynb cell 20: Model([PRMSChannel]) fed
    with the three volume fluxes."""
    import pywatershed_b200 as pwsb

    p, f, start = drb
    nd = 60
    o = po.Oracle(p, start=start)
    ref, _ = o.run(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd],
                   ["sroff_vol", "ssres_flow_vol", "gwres_flow_vol"], po.SEG_VARS)
    control = _control(start, nd)
    m = pwsb.Model([pwsb.PRMSChannel], control=control, parameters=_parameters(p),
                   find_input_files=False)
    ch = m.processes["PRMSChannel"]
    for k in ("sroff_vol", "ssres_flow_vol", "gwres_flow_vol"):
        ch.set_input_to_adapter(k, ref[k])
    m.run(finalize=False)
    np.testing.assert_allclose(ch["seg_outflow"], ref["seg_outflow"][-1], rtol=1e-9, atol=1e-10)


def test_reference_error_conventions(drb):
    import pywatershed_b200 as pwsb

    p, f, start = drb
    control = _control(start, 5)
    with pytest.raises(NameError):
        control.options["not_an_option"] = 1
    q = {k: v for k, v in p.items() if k != "carea_max"}
    with pytest.raises(ValueError, match="carea_max"):
        pwsb.PRMSRunoff(control, None, _parameters(q))
    with pytest.raises(ValueError, match="calc_method"):
        pwsb.PRMSCanopy(control, None, _parameters(p), calc_method="numpy")
    m = pwsb.Model([pwsb.PRMSSnow], control=control, parameters=_parameters(p),
                   find_input_files=False)
    with pytest.raises(ValueError, match="was not provided"):  # a single process needs its inputs
        m.advance()


def test_nodprst_model_matches_the_reference(drb):
    """nhm_no_dprst: the chain with PRMSRunoffNoDprst / PRMSSoilzoneNoDprst /
    PRMSGroundwaterNoDprst (autotest/test_prms_below_snow.py:31-45) against
    the reference's own outputs (tests/golden/drb_nodprst_golden.npz)."""
    import pywatershed_b200 as pwsb

    p, f, start = drb
    nd = 400
    names = ["PRMSSolarGeometry", "PRMSAtmosphere", "PRMSCanopy", "PRMSSnow", "PRMSRunoffNoDprst",
             "PRMSSoilzoneNoDprst", "PRMSGroundwaterNoDprst", "PRMSChannel"]
    q = {k: v for k, v in p.items() if not k.startswith("dprst") and k not in (
        "sro_to_dprst_perv", "sro_to_dprst_imperv", "va_open_exp", "va_clos_exp", "op_flow_thres")}
    control = _control(start, nd, dprst_flag=False)
    m = pwsb.Model([getattr(pwsb, n) for n in names], control=control, parameters=_parameters(q),
                   find_input_files=False)
    for k in ("prcp", "tmax", "tmin"):
        m.processes["PRMSAtmosphere"].set_input_to_adapter(k, f[k][:nd])
    g = np.load(scenarios.GOLDEN / "drb_nodprst_golden.npz")
    hs, ss = g["hru_sample"], g["seg_sample"]
    assert "dprst_stor_hru" not in m.processes["PRMSRunoffNoDprst"].get_variables()
    series = {}
    for it in range(nd):
        m.advance()
        m.calculate()
        for pname in names[4:]:
            proc = m.processes[pname]
            for v in proc.get_variables():
                a = np.asarray(proc[v], dtype=np.float64)
                series.setdefault(v, []).append(a[ss] if v in po.SEG_VARS else a[hs])
    got = {v: np.stack(rows) for v, rows in series.items()}
    ref = {v: np.asarray(g["ts_" + v], dtype=np.float64) for v in got}
    hru_names = [v for v in got if v not in po.SEG_VARS]
    rep = parity.compare_columns(got, ref, hru_names)
    # ghost-pack ties (tests/parity.py) can touch a sampled HRU; anything else is a mismatch
    assert len(rep["mismatches"]) <= 1, rep["mismatches"][:3]
    # Sampled segments: rtol 1e-9 unless a ghost-pack tie on one of the (unsampled) HRUs
    # upstream moved some water between days; then the flow summed over the run decides
    so_g, so_r = got["seg_outflow"], ref["seg_outflow"]
    close = np.isclose(so_g, so_r, rtol=1e-9, atol=1e-10)
    if not close.all():
        seg_ok = close.all(axis=0)
        cum = np.abs(so_g.sum(axis=0) - so_r.sum(axis=0)) / np.maximum(np.abs(so_r.sum(axis=0)), 1e-6)
        print(f"nodprst: {int((~seg_ok).sum())} of {seg_ok.size} sampled segments downstream of a tie, "
              f"cumulative flow off by at most {cum.max():.2e}")
        assert cum.max() <= parity.POST_TIE_CUM_RTOL
    for pname in names[4:7]:
        assert m.processes[pname].budget._zero_sum


def test_sagehen_model_through_the_api(sagehen):
    """The reference's second test domain through the Model API exactly as its
    sagehen_no_cascades_model.yaml composes it: seven processes, NoDprst
    variants, no PRMSChannel, a parameter set WITHOUT depression-storage or
    channel parameters; compared with the reference's outputs
    (tests/golden/sagehen_golden.npz) on the sampled HRUs, all 1,826 days."""
    import pywatershed_b200 as pwsb

    z = np.load(scenarios.GOLDEN / "sagehen_inputs.npz")
    raw = {k[2:]: z[k] for k in z.files if k.startswith("p_")}  # as in the parameter file
    _, f, start = sagehen
    nd = f["prcp"].shape[0]
    names = ["PRMSSolarGeometry", "PRMSAtmosphere", "PRMSCanopy", "PRMSSnow", "PRMSRunoffNoDprst",
             "PRMSSoilzoneNoDprst", "PRMSGroundwaterNoDprst"]
    nhru = raw["hru_area"].shape[0]
    params = pwsb.Parameters(dims={"nhru": nhru, "nmonth": 12, "ndoy": 366},
                             coords={"nhm_id": np.arange(1, nhru + 1)},
                             data_vars={k: v for k, v in raw.items() if k != "nhm_id"})
    control = _control(start, nd, dprst_flag=False)
    m = pwsb.Model([getattr(pwsb, n) for n in names], control=control, parameters=params,
                   find_input_files=False)
    for k in ("prcp", "tmax", "tmin"):
        m.processes["PRMSAtmosphere"].set_input_to_adapter(k, f[k])
    g = np.load(scenarios.GOLDEN / "sagehen_golden.npz")
    hs = g["hru_sample"]
    series = {}
    for it in range(nd):
        m.advance()
        m.calculate()
        for pname in names[2:]:
            proc = m.processes[pname]
            for v in proc.get_variables():
                series.setdefault(v, []).append(np.asarray(proc[v], dtype=np.float64)[hs])
    got = {v: np.stack(rows) for v, rows in series.items()}
    ref = {v: np.asarray(g["ts_" + v], dtype=np.float64) for v in got}
    rep = parity.compare_columns(got, ref, list(got), got["pkwater_equiv"], ref["pkwater_equiv"])
    assert rep["mismatches"] == [], rep["mismatches"][:3]
    assert got["pkwater_equiv"].max() > 30.0
    for pname in names[2:]:
        assert m.processes[pname].budget._zero_sum


def _write_cbh(path, name, start, values):
    """PRMS CBH ASCII (utils/cbh_utils.py:32-122): header, '####', then one
    line per day: year month day hour minute second v1 ... vn."""
    t = np.datetime64(start, "D") + np.arange(values.shape[0])
    with open(path, "w") as fh:
        fh.write("Written by tests/test_gpu_model.py\n")
        fh.write(f"{name} {values.shape[1]}\n")
        fh.write("########################################\n")
        for d, row in zip(t, values):
            y, m, dd = str(d).split("-")
            fh.write(f"{int(y)} {int(m)} {int(dd)} 0 0 0 " + " ".join(repr(float(v)) for v in row) + "\n")


def test_model_finds_cbh_input_files(drb, tmp_path):
    """Model(find_input_files=True): forcings come from <input_dir>/<name>.cbh
    (base/model.py:509-560), parsed by our PRMS ASCII reader."""
    import pywatershed_b200 as pwsb

    p, f, start = drb
    nd = 25
    for k in ("prcp", "tmax", "tmin"):
        _write_cbh(tmp_path / f"{k}.cbh", k, start, f[k][:nd + 3])  # file longer than the run
    control = _control(start, nd, input_dir=str(tmp_path))
    m = pwsb.Model([getattr(pwsb, n) for n in NHM], control=control, parameters=_parameters(p))
    m.run(finalize=True)
    o = po.Oracle(p, start=start)
    ref, _ = o.run(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd], ["soil_moist", "hru_ppt"], ["seg_outflow"])
    np.testing.assert_allclose(m.processes["PRMSSoilzone"]["soil_moist"], ref["soil_moist"][-1],
                               rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(m.processes["PRMSChannel"]["seg_outflow"], ref["seg_outflow"][-1],
                               rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(m.processes["PRMSAtmosphere"]["hru_ppt"].data, ref["hru_ppt"],
                               rtol=1e-9, atol=1e-10)


def test_model_from_yaml_runs(drb, tmp_path):
    """Model.from_yaml (base/model.py:595-664) on files laid out like the
    reference's test_data/drb_2yr/nhm_model.yaml: control yaml, parameter file,
    CBH forcings found in input_dir; run and compared with the oracle."""
    import pywatershed_b200 as pwsb
    from test_host_layer import _write_model_yaml

    p, f, start = drb
    nd = 40
    for k in ("prcp", "tmax", "tmin"):
        _write_cbh(tmp_path / f"{k}.cbh", k, start, f[k][:nd])
    m = pwsb.Model.from_yaml(_write_model_yaml(tmp_path, p, f, start, nd))
    m.run(finalize=True)
    o = po.Oracle(p, start=start)
    ref, _ = o.run(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd], ["pkwater_equiv", "gwres_flow"],
                   ["seg_outflow"])
    np.testing.assert_allclose(m.processes["PRMSSnow"]["pkwater_equiv"], ref["pkwater_equiv"][-1],
                               rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(m.processes["PRMSGroundwater"]["gwres_flow"], ref["gwres_flow"][-1],
                               rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(m.processes["PRMSChannel"]["seg_outflow"], ref["seg_outflow"][-1],
                               rtol=1e-9, atol=1e-10)


def test_run_accumulates_budgets_on_the_device_like_stepping(drb):
    """Model.run() sums the budget terms on the device (pws_select_accumulations)
    and leaves the all-time Atmosphere arrays on the device side until somebody
    reads them; a model stepped day by day sums on the host.  Same numbers."""
    p, f, start = drb
    nd = 50
    m1 = _model(p, f, start, nd, block_days=16)
    m1.run(finalize=False)
    atm = m1.processes["PRMSAtmosphere"]
    assert not atm._vars["tmaxf"].allocated  # nothing was drained for it
    m2 = _model(p, f, start, nd, block_days=16)
    for _ in range(nd):
        m2.advance()
        m2.calculate()
    for name in NHM[2:]:
        a1, a2 = m1.processes[name].budget.accumulations, m2.processes[name].budget.accumulations
        for comp in a1:
            for v in a1[comp]:
                np.testing.assert_allclose(a1[comp][v], a2[comp][v], rtol=1e-12, atol=1e-12,
                                           err_msg=f"{name} {comp} {v}")
    # the lazily replayed all-time arrays equal the ones filled block by block
    for k in ("tmaxf", "swrad", "potet", "pptmix", "transp_on"):
        np.testing.assert_array_equal(atm[k].data, m2.processes["PRMSAtmosphere"][k].data, err_msg=k)
        np.testing.assert_array_equal(atm[k].current, m2.processes["PRMSAtmosphere"][k].current)
    # and run() in pieces continues where it stopped
    m3 = _model(p, f, start, nd, block_days=16)
    m3.run(n_time_steps=20, finalize=False)
    m3.run(finalize=False)
    a3 = m3.processes["PRMSSoilzone"].budget.accumulations
    np.testing.assert_allclose(a3["outputs"]["perv_actet_hru"],
                               m1.processes["PRMSSoilzone"].budget.accumulations["outputs"]["perv_actet_hru"],
                               rtol=1e-12)
    for m in (m1, m2, m3):
        m.finalize()


def test_adjust_parameters_follows_the_reference(drb):
    """prms_soilzone.py:361-463 / prms_channel.py:246-260: "warn" announces the
    edits with the reference's messages, "error" raises after announcing, "no"
    is silent -- per process (the channel's choice does not switch off the
    soilzone's), and a control with dprst_flag=True runs the NoDprst classes."""
    import warnings

    import pywatershed_b200 as pwsb

    p, f, start = drb
    q = {k: np.array(v, copy=True) for k, v in p.items()}
    q["soil_rechr_init_frac"] = np.asarray(q["soil_rechr_init_frac"], dtype=np.float64).copy()
    q["soil_rechr_init_frac"][7] = 1.5       # soil_rechr_init > soil_rechr_max
    q["seg_slope"] = np.asarray(q["seg_slope"], dtype=np.float64).copy()
    q["seg_slope"][3] = 1e-9
    nd = 5

    def model(sz, ch):
        control = _control(start, nd)
        md = {"control": control, "dis": _parameters(q)}
        for n in NHM:
            kw = {"class": getattr(pwsb, n), "parameters": _parameters(q), "dis": "dis"}
            if n == "PRMSSoilzone":
                kw["adjust_parameters"] = sz
            if n == "PRMSChannel":
                kw["adjust_parameters"] = ch
            md[n] = kw
        md["model_order"] = list(NHM)
        m = pwsb.Model(md, find_input_files=False)
        for k in ("prcp", "tmax", "tmin"):
            m.processes["PRMSAtmosphere"].set_input_to_adapter(k, f[k][:nd])
        return m

    with pytest.warns(UserWarning) as rec:
        m = model("warn", "warn")
        m.run(finalize=True)
    msgs = [str(w.message) for w in rec]
    assert any("soil_rechr_init > soil_rechr_max" in s and "[7]" in s for s in msgs), msgs
    assert any("seg_slope < 1.0e-7" in s and "[3]" in s for s in msgs), msgs
    with pytest.warns(UserWarning), pytest.raises(ValueError, match="an error was requested"):
        model("error", "no").run()
    with pytest.warns(UserWarning), pytest.raises(ValueError, match="seg_slope parameter values"):
        model("no", "error").run()
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        model("no", "no").run(finalize=True)
    # "no" on the channel leaves the soilzone's adjustment on: same result as warn/warn
    with pytest.warns(UserWarning, match="soil_rechr_init"):
        m2 = model("warn", "no")
        m2.run(finalize=True)
    np.testing.assert_array_equal(m.processes["PRMSSoilzone"]["soil_rechr"],
                                  m2.processes["PRMSSoilzone"]["soil_rechr"])
    # a control that says dprst_flag=True with the NoDprst classes (the reference runs it)
    nod = ["PRMSSolarGeometry", "PRMSAtmosphere", "PRMSCanopy", "PRMSSnow", "PRMSRunoffNoDprst",
           "PRMSSoilzoneNoDprst", "PRMSGroundwaterNoDprst", "PRMSChannel"]
    m3 = _model(p, f, start, nd, procs=nod, dprst_flag=True)
    m3.run(finalize=True)
    assert m3.processes["PRMSRunoffNoDprst"]._dprst_flag is False
    assert m3.processes["PRMSCanopy"]._dprst_flag is None


def test_model_restart_round_trip(drb, tmp_path):
    """Model.save_state at a block boundary -> Model.load_state into a freshly
    built model -> run on: the last day and the budget accumulations equal the
    uninterrupted run bit for bit, PRMSChannel included."""
    p, f, start = drb
    nd = 60
    a = _model(p, f, start, nd)
    a.run(finalize=False)
    b = _model(p, f, start, nd)
    b.run(n_time_steps=25, finalize=False)
    b.save_state(tmp_path / "restart.npz")
    b.finalize()
    c = _model(p, f, start, nd)
    c.load_state(tmp_path / "restart.npz")
    assert c.control.itime_step == 24
    c.run(finalize=False)
    assert c.control.itime_step == nd - 1
    for name in NHM[2:]:
        pa, pc = a.processes[name], c.processes[name]
        for v in pa.get_variables():
            np.testing.assert_array_equal(np.asarray(pa[v]), np.asarray(pc[v]), err_msg=f"{name}.{v}")
        acc_a, acc_c = pa.budget.accumulations, pc.budget.accumulations
        for comp in acc_a:
            for v in acc_a[comp]:
                np.testing.assert_allclose(acc_a[comp][v], acc_c[comp][v], rtol=1e-13, atol=1e-13,
                                           err_msg=f"{name} {comp} {v}")
    np.testing.assert_array_equal(a.processes["PRMSAtmosphere"]["swrad"].data,
                                  c.processes["PRMSAtmosphere"]["swrad"].data)
    m = _model(p, f, start, nd, block_days=16)
    for _ in range(3):
        m.advance()
        m.calculate()
    with pytest.raises(RuntimeError, match="block boundary"):
        m.save_state()
    for x in (a, c, m):
        x.finalize()
