"""CPU tests of the reference-facing host layer (no compute calls): Control,
Parameters, the PRMS ASCII readers, Budget arithmetic, the process
descriptors and the netCDF writer.  Where the reference's own input files are
needed they are read from /root/reference and the test is skipped when that
tree is absent (the GPU box)."""
import pathlib as pl
import warnings

import numpy as np
import pytest

import scenarios

REF_DRB = pl.Path("/root/reference/pywatershed/data/drb_2yr")
needs_ref_files = pytest.mark.skipif(not REF_DRB.exists(), reason="reference input files absent")


def _control(nd=4, **opts):
    from pywatershed_b200 import Control

    t0 = np.datetime64("1979-01-03T00:00:00")
    return Control(t0, t0 + np.timedelta64(nd - 1, "D"), np.timedelta64(1, "D"), options=opts)


# ---------------------------------------------------------------- Control
def test_control_clock_follows_the_reference():
    """autotest/test_control.py: init_time = start - step, itime_step -1 -> 0.."""
    c = _control(4)
    assert c.n_times == 4 and c.itime_step == -1
    assert c.current_time == np.datetime64("1979-01-02T00:00:00")
    c.advance()
    assert c.itime_step == 0 and c.current_time == c.start_time
    assert (c.current_doy, c.current_dowy, c.current_month) == (3, 95, 1)
    assert c.previous_time == np.datetime64("1979-01-02T00:00:00")
    for _ in range(3):
        c.advance()
    assert c.current_time == c.end_time
    with pytest.raises(ValueError, match="End of time"):
        c.advance()
    with pytest.raises(NameError):
        c.options["nope"] = 1
    with pytest.raises(NameError):
        _control(2, nope=1)
    c.edit_n_time_steps(10)
    assert c.end_time == c.start_time + np.timedelta64(9, "D")


@needs_ref_files
def test_control_load_prms_drb():
    from pywatershed_b200 import Control

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        c = Control.load_prms(REF_DRB / "nhm.control", warn_unused_options=False)
    assert c.start_time == np.datetime64("1979-01-01T00:00:00")
    assert c.end_time == np.datetime64("1980-12-31T00:00:00")
    assert c.n_times == 731
    assert c.time_step == np.timedelta64(24, "h")
    names = c.options["netcdf_output_var_names"]
    import json

    default = json.loads((scenarios.GOLDEN / "nhm_default_vars.json").read_text())
    assert set(default) <= set(names)


# ---------------------------------------------------------------- readers
@needs_ref_files
def test_prms_parameter_and_cbh_readers_equal_the_reference_loaders(drb):
    """tests/golden/drb_inputs.npz was written from the reference's own
    PrmsParameters.load / cbh_files_to_np_dict (oracle/gen_golden.py)."""
    from pywatershed_b200 import PrmsParameters
    from pywatershed_b200.prms_files import read_cbh

    p, f, start = drb
    mine = PrmsParameters.load(REF_DRB / "myparam.param")
    d = mine.to_dict()
    checked = 0
    for k, v in p.items():
        if k not in d:
            continue
        np.testing.assert_array_equal(np.asarray(d[k]).squeeze(), np.asarray(v).squeeze(), err_msg=k)
        checked += 1
    assert checked > 100
    assert mine.dims["nhru"] == 765 and mine.dims["nsegment"] == 456
    with pytest.raises(ValueError):
        mine.parameters["hru_area"][0] = 1.0  # read-only, parameters.py:246-254
    for k in ("prcp", "tmax", "tmin"):
        times, vals = read_cbh(REF_DRB / f"{k}.cbh", 765)
        assert times[0] == np.datetime64(start) and len(times) == 731
        np.testing.assert_array_equal(vals, f[k])


# ---------------------------------------------------------------- Budget
class _FakeProc(dict):
    pass


def _budget(budget_type, stor):
    from pywatershed_b200 import Budget

    n = 5
    proc = _FakeProc(in1=np.ones(n), in2=np.ones(n), out1=np.zeros(n), out2=np.ones(n),
                     stor1=np.full(n, stor), stor2=np.zeros(n))
    terms = {"inputs": ["in1", "in2"], "outputs": ["out1", "out2"], "storage_changes": ["stor1", "stor2"]}
    c = _control(4)
    return c, Budget(c, proc, terms, budget_type, description="simple_test"), proc


def test_budget_arithmetic_as_reference_test_budget():
    """autotest/test_budget.py:25-118."""
    c, b, proc = _budget("error", 1.0)
    for comp in b.components:
        assert set(b[comp]) == set(b.terms[comp])
    c.advance()
    b.advance()
    b.calculate()
    assert (b._inputs_sum == 2).all() and (b._outputs_sum == 1).all()
    assert (b._storage_changes_sum == 1).all()
    assert (b.balance == 1).all() and b._zero_sum
    with pytest.raises(ValueError, match="twice"):
        b.calculate()
    b.accumulate_block({k: np.stack([v, v]) for k, v in proc.items()}, 1.0)
    assert (b.accumulations["inputs"]["in1"] == 2).all()


def test_budget_imbalance_error_and_warn():
    c, b, _ = _budget("error", 0.5)
    c.advance()
    with pytest.raises(ValueError, match="simple_test"):
        b.calculate()
    c, b, _ = _budget("warn", 0.5)
    c.advance()
    with pytest.warns(UserWarning, match="simple_test"):
        b.calculate()
    assert not b._zero_sum
    c, b, _ = _budget("error", 1.0)
    c.advance()
    with pytest.raises(ValueError):
        b.calculate(n_violations=3)  # the device's verdict


def test_budget_accumulation_surface_of_the_reference():
    """base/budget.py:204-237, 283-310, 422-648: set_initial_accumulations /
    reset_accumulations / _sum_component_accumulations and the report."""
    c, b, proc = _budget("error", 1.0)
    assert "only initialized" in repr(b)
    assert b.units is None and set(b.output_vars_desc) == {"inputs_sum", "outputs_sum", "storage_changes_sum"}
    b.set_initial_accumulations(None)
    assert (b.accumulations["inputs"]["in1"] == 0).all()
    c.advance()
    b.calculate()
    b.accumulate_block({k: np.stack([v, v, v]) for k, v in proc.items()}, 1.0)
    sums = b._sum_component_accumulations()
    assert (sums["inputs"] == 6).all() and (sums["outputs"] == 3).all() and (sums["storage_changes"] == 3).all()
    assert (b.inputs_sum == 2).all() and (b.outputs_sum == 1).all() and (b.storage_changes_sum == 1).all()
    text = repr(b)
    assert "Individual spatial unit budget for simple_test" in text
    assert "This timestep, rates:" in text and "Accumulated volumes" in text
    assert "in1: 5.e+00" in text and "in1: 1.5e+01" in text  # spatial sums of the rate and the volume
    assert "Balance: " in text and "!=!" not in text
    b.reset_accumulations()
    assert (b.accumulations["outputs"]["out2"] == 0).all()
    b.set_initial_accumulations({"inputs": {"in1": np.full(5, 7.0)}}, np.datetime64("1979-01-01"))
    assert (b.accumulations["inputs"]["in1"] == 7).all() and (b.accumulations["inputs"]["in2"] == 0).all()
    assert b._accum_start_time == np.datetime64("1979-01-01")
    # a global-basis budget sums over the units
    from pywatershed_b200 import Budget

    g = Budget(c, proc, b.terms, "warn", basis="global", description="glob")
    g.accumulate_block({k: np.stack([v, v]) for k, v in proc.items()}, 1.0)
    assert g._sum_component_accumulations()["inputs"] == 20.0
    assert "Global budget for glob" in repr(g)


def test_budget_units_come_from_the_terms_metadata():
    import pywatershed_b200 as pwsb

    c = _control(3, budget_type="warn")
    terms = pwsb.PRMSSnow.get_mass_budget_terms()
    proc = _FakeProc({v: np.zeros(2) for comp in terms.values() for v in comp})
    b = pwsb.Budget(c, proc, terms, "warn", description="PRMSSnow")
    assert b.units == "inches"
    assert b.meta["inputs_sum"]["desc"] == "Sum of input fluxes (unit)"


# ---------------------------------------------------------------- processes
def test_process_descriptors_and_variable_registry():
    import pywatershed_b200 as pwsb
    from pywatershed_b200._process_meta import PROCESS_META, VARIABLE_META

    import prms_oracle as po

    allv = []
    for name in ("PRMSSolarGeometry", "PRMSAtmosphere", "PRMSCanopy", "PRMSSnow", "PRMSRunoff",
                 "PRMSSoilzone", "PRMSGroundwater", "PRMSChannel"):
        cls = getattr(pwsb, name)
        assert cls.get_variables() == PROCESS_META[name]["variables"]
        assert set(cls.get_init_values()) == set(cls.get_variables())
        assert cls.description()["class_name"] == name
        allv += list(cls.get_variables())
    assert len(allv) == 151 == len(set(allv))
    hru = [v for v in allv if not v.startswith("soltab") and "nsegment" not in VARIABLE_META[v]["dims"]]
    assert sorted(hru) == sorted(po.HRU_VARS)
    assert pwsb.PRMSSnow.get_mass_budget_terms()["outputs"] == ["snow_evap", "snowmelt", "through_rain"]
    assert pwsb.PRMSChannel._budget_basis == "global"


@pytest.mark.skipif(not pl.Path("/root/reference/pywatershed").exists(), reason="reference absent")
def test_descriptors_equal_the_reference_classes():
    import sys

    sys.path.insert(0, str(pl.Path(__file__).resolve().parent.parent / "oracle"))
    import ref_harness as rh

    import pywatershed_b200 as pwsb

    pws = rh.import_reference()
    for name in rh.NHM_PROCESS_NAMES + ("PRMSEt",):
        ref, mine = getattr(pws, name), getattr(pwsb, name)
        assert tuple(ref.get_inputs()) == tuple(mine.get_inputs()), name
        assert tuple(ref.get_variables()) == tuple(mine.get_variables()), name
        assert tuple(ref.get_parameters()) == tuple(mine.get_parameters()), name
        if hasattr(ref, "get_mass_budget_terms"):
            assert {k: list(v) for k, v in ref.get_mass_budget_terms().items()} == \
                mine.get_mass_budget_terms(), name
        import inspect

        a = list(inspect.signature(ref.__init__).parameters)
        b = list(inspect.signature(mine.__init__).parameters)
        assert a == b, (name, a, b)


def test_model_builds_on_cpu_and_refuses_to_compute_without_a_gpu(drb):
    import torch

    import pywatershed_b200 as pwsb
    from pywatershed_b200 import _lib

    p, f, start = drb
    nhru, nseg = 765, 456
    params = pwsb.Parameters(dims={"nhru": nhru, "nsegment": nseg},
                             coords={"nhm_id": np.arange(1, nhru + 1), "nhm_seg": np.arange(1, nseg + 1)},
                             data_vars=p)
    c = _control(5, budget_type="warn")
    m = pwsb.Model([pwsb.PRMSSolarGeometry, pwsb.PRMSAtmosphere, pwsb.PRMSCanopy, pwsb.PRMSSnow,
                    pwsb.PRMSRunoff, pwsb.PRMSSoilzone, pwsb.PRMSGroundwater, pwsb.PRMSChannel],
                   control=c, parameters=params, find_input_files=False)
    assert list(m.processes) == ["PRMSSolarGeometry", "PRMSAtmosphere", "PRMSCanopy", "PRMSSnow",
                                 "PRMSRunoff", "PRMSSoilzone", "PRMSGroundwater", "PRMSChannel"]
    snow = m.processes["PRMSSnow"]
    assert snow["pkwater_equiv"].shape == (nhru,) and snow["newsnow"].dtype == np.int32
    assert snow["iasw"].dtype == np.bool_
    assert m.processes["PRMSChannel"]["seg_outflow"].shape == (nseg,)
    assert m.processes["PRMSAtmosphere"]["tmaxf"].data.shape == (5, nhru)
    assert m.processes["PRMSSolarGeometry"]["soltab_potsw"].data.shape == (366, nhru)
    if not torch.cuda.is_available():
        with pytest.raises(_lib.PwsLibraryError):  # no CPU fallback
            m.advance()


def test_netcdf_writer_layout(tmp_path):
    from pywatershed_b200.netcdf_out import _File

    f = _File(tmp_path / "pkwater_equiv.nc", ["pkwater_equiv"], "nhm_id", np.arange(1, 8), "PRMSSnow")
    a = np.arange(21, dtype=np.float64).reshape(3, 7)
    f.write(0, 2, [3287.0, 3288.0], {"pkwater_equiv": a[:2]})
    f.write(2, 3, [3289.0], {"pkwater_equiv": a[2:]})
    f.close()
    from pywatershed_b200.netcdf_out import nc4_module

    if nc4_module() is not None:
        ds = nc4_module().Dataset(tmp_path / "pkwater_equiv.nc")
    else:
        from scipy.io import netcdf_file

        ds = netcdf_file(str(tmp_path / "pkwater_equiv.nc"), "r", mmap=False)
    v = ds.variables["pkwater_equiv"]
    assert tuple(v.dimensions) == ("time", "nhm_id")
    np.testing.assert_array_equal(np.array(v[:]), a)
    np.testing.assert_array_equal(np.array(ds.variables["nhm_id"][:]), np.arange(1, 8))
    np.testing.assert_array_equal(np.array(ds.variables["time"][:]), [3287.0, 3288.0, 3289.0])
    units = ds.variables["time"].units
    assert (units.decode() if isinstance(units, bytes) else units) == "days since 1970-01-01 00:00:00"
    ds.close()


@pytest.mark.skipif(not pl.Path("/root/reference/pywatershed").exists(), reason="reference absent")
def test_model_accepts_the_reference_control_and_parameters():
    """Duck typing: a pywatershed.Control / PrmsParameters
    pair drives our Model; its clock and option dict are the ones used."""
    import sys

    sys.path.insert(0, str(pl.Path(__file__).resolve().parent.parent / "oracle"))
    import ref_harness as rh

    import pywatershed_b200 as pwsb

    pws = rh.import_reference()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        control, params, forcings, times = rh.load_drb(pws)
    for key in ("netcdf_output_dir", "netcdf_output_var_names"):
        if key in control.options:
            del control.options[key]
    control.options["budget_type"] = "error"
    if "calc_method" in control.options:
        del control.options["calc_method"]  # "numpy" / "numba" name the reference's CPU paths
    m = pwsb.Model([pws.PRMSSolarGeometry, pws.PRMSAtmosphere, pws.PRMSCanopy, pws.PRMSSnow,
                    pws.PRMSRunoff, pws.PRMSSoilzone, pws.PRMSGroundwater, pws.PRMSChannel],
                   control=control, parameters=params, find_input_files=False)
    assert m._mode == "chain+channel" and m.control is control
    assert m.processes["PRMSSnow"]["pkwater_equiv"].shape == (765,)
    assert m.processes["PRMSAtmosphere"]["tmaxf"].data.shape == (control.n_times, 765)
    # reference classes were mapped onto ours by name
    assert type(m.processes["PRMSRunoff"]).__module__.startswith("pywatershed_b200")
    np.testing.assert_array_equal(m._param_dict["hru_area"], params.parameters["hru_area"])
    assert m._param_dict["tmax_cbh_adj"].shape == (12, 765)


def test_parameters_netcdf_round_trip(drb, tmp_path):
    """Parameters.to_netcdf / from_netcdf (base/data_model.py:354-420)."""
    import pywatershed_b200 as pwsb

    p, f, start = drb
    keep = ("hru_area", "tmax_cbh_adj", "cov_type", "tosegment", "seg_length", "snarea_curve")
    params = pwsb.Parameters(
        dims={"nhru": 765, "nsegment": 456, "nmonth": 12, "ndeplval": 55},
        coords={"nhm_id": p["nhm_id"], "nhm_seg": p["nhm_seg"]},
        data_vars={k: p[k] for k in keep},
        metadata={"hru_area": {"dims": ("nhru",)}, "tmax_cbh_adj": {"dims": ("nmonth", "nhru")},
                  "cov_type": {"dims": ("nhru",)}, "tosegment": {"dims": ("nsegment",)},
                  "seg_length": {"dims": ("nsegment",)}, "snarea_curve": {"dims": ("ndeplval",)},
                  "nhm_id": {"dims": ("nhru",)}, "nhm_seg": {"dims": ("nsegment",)}})
    params.to_netcdf(tmp_path / "params.nc")
    back = pwsb.Parameters.from_netcdf(tmp_path / "params.nc")
    assert back.dims["nhru"] == 765 and back.dims["nmonth"] == 12
    for k in keep:
        np.testing.assert_array_equal(back.parameters[k], p[k], err_msg=k)
    np.testing.assert_array_equal(back.coords["nhm_id"], p["nhm_id"])
    sub = back.subset(["hru_area", "cov_type"])
    assert set(sub.parameters) == {"hru_area", "cov_type"}


def test_model_dict_construction(drb):
    """The model_dict style of base/model.py:56-120: a Control, named
    discretizations, one entry per process with its class, parameters and
    discretization, and a model_order."""
    import pywatershed_b200 as pwsb

    p, f, start = drb
    seg_keys = ("tosegment", "tosegment_nhm", "seg_length", "seg_slope", "seg_depth", "mann_n", "x_coef",
                "segment_type", "segment_flow_init", "nhm_seg")
    dis_hru = pwsb.Parameters(dims={"nhru": 765}, coords={"nhm_id": p["nhm_id"]},
                              data_vars={k: p[k] for k in ("hru_area", "hru_lat", "hru_slope", "hru_aspect",
                                                           "hru_type", "hru_in_to_cf")})
    dis_both = pwsb.Parameters(dims={"nhru": 765, "nsegment": 456},
                               coords={"nhm_id": p["nhm_id"], "nhm_seg": p["nhm_seg"]},
                               data_vars={k: p[k] for k in ("hru_area", "hru_segment", "hru_in_to_cf")
                                          + tuple(k for k in seg_keys if k in p and k != "nhm_seg")})
    rest = pwsb.Parameters(dims={"nhru": 765}, data_vars={
        k: v for k, v in p.items() if k not in dis_hru.parameters and k not in dis_both.parameters
        and k not in ("nhm_id", "nhm_seg")})
    control = _control(6, budget_type="warn")
    model_dict = {
        "control": control,
        "dis_hru": dis_hru,
        "dis_both": dis_both,
        "groundwater": {"class": pwsb.PRMSGroundwater, "parameters": rest, "dis": "dis_hru"},
        "channel": {"class": pwsb.PRMSChannel, "parameters": rest, "dis": "dis_both"},
        "model_order": ["groundwater", "channel"],
    }
    m = pwsb.Model(model_dict, find_input_files=False)
    assert list(m.processes) == ["PRMSGroundwater", "PRMSChannel"] and m._mode == "sub"
    assert m.control is control
    assert m._external_inputs() == ["soil_to_gw", "ssr_to_gw", "dprst_seep_hru", "sroff_vol", "ssres_flow_vol"]
    assert m._param_dict["tosegment"].shape == (456,) and m._param_dict["gwflow_coef"].shape == (765,)
    assert m.processes["PRMSChannel"].budget.basis == "global"


def test_control_yaml_round_trip(tmp_path):
    """Control.to_yaml / from_yaml (base/control.py:518-610): options flattened
    next to the four time keys, input_dir relative to the file."""
    from pywatershed_b200 import Control

    (tmp_path / "in").mkdir()
    (tmp_path / "control.yaml").write_text(
        "start_time: 1979-01-01T00:00:00\nend_time: 1979-01-10T00:00:00\ntime_step: 24\n"
        "time_step_units: h\nverbosity: 0\ninput_dir: in\nbudget_type: warn\ncalc_method: cuda\n"
        "netcdf_output_var_names:\n  - albedo\n  - seg_outflow\n")
    c = Control.from_yaml(tmp_path / "control.yaml")
    assert c.n_times == 10 and c.start_time == np.datetime64("1979-01-01T00:00:00")
    assert c.options["input_dir"] == (tmp_path / "in").resolve()
    assert c.options["netcdf_output_var_names"] == ["albedo", "seg_outflow"]
    c.to_yaml(tmp_path / "again.yaml")
    c2 = Control.from_yaml(tmp_path / "again.yaml")
    assert c2.n_times == 10 and c2.time_step == c.time_step and dict(c2.options) == dict(c.options)
    (tmp_path / "bad.yaml").write_text(
        "start_time: 1979-01-01T00:00:00\nend_time: 1979-01-10T00:00:00\ntime_step: 24\n"
        "time_step_units: h\nnot_an_option: 1\n")
    with pytest.raises(NameError):  # base/control.py:630-636
        Control.from_yaml(tmp_path / "bad.yaml")


def _write_model_yaml(tmp_path, p, f, start, nd):
    """A model yaml like the reference's test_data/drb_2yr/nhm_model.yaml."""
    scenarios.write_prms_param_file(tmp_path / "myparam.param", p)
    (tmp_path / "control.yaml").write_text(
        f"start_time: {start}T00:00:00\nend_time: {np.datetime64(start) + nd - 1}T00:00:00\n"
        "time_step: 24\ntime_step_units: h\nverbosity: 0\ninput_dir: .\nbudget_type: error\n"
        "calc_method: cuda\n")
    names = ["PRMSSolarGeometry", "PRMSAtmosphere", "PRMSCanopy", "PRMSSnow", "PRMSRunoff", "PRMSSoilzone",
             "PRMSGroundwater", "PRMSChannel"]
    txt = "control: control.yaml\ndis_both: myparam.param\n"
    for n in names:
        txt += f"{n[4:].lower()}:\n  class: {n}\n  parameters: myparam.param\n  dis: dis_both\n"
    txt += "model_order:\n" + "".join(f"  - {n[4:].lower()}\n" for n in names)
    (tmp_path / "nhm_model.yaml").write_text(txt)
    return tmp_path / "nhm_model.yaml"


def test_model_from_yaml_builds(drb, tmp_path):
    """Model.from_yaml / model_dict_from_yaml (base/model.py:541-664)."""
    import pywatershed_b200 as pwsb

    p, f, start = drb
    y = _write_model_yaml(tmp_path, p, f, start, 10)
    md = pwsb.Model.model_dict_from_yaml(y)
    assert md["snow"]["class"] is pwsb.PRMSSnow and md["model_order"][0] == "solargeometry"
    assert isinstance(md["control"], pwsb.Control) and md["dis_both"].dims["nsegment"] == 456
    m = pwsb.Model.from_yaml(y)
    assert m._mode == "chain+channel" and list(m.processes)[-1] == "PRMSChannel"
    assert m.control.n_times == 10
    np.testing.assert_array_equal(m._param_dict["soil_moist_max"], p["soil_moist_max"])
    np.testing.assert_array_equal(m._param_dict["tmax_cbh_adj"], p["tmax_cbh_adj"])
    (tmp_path / "bad.yaml").write_text("control: control.yaml\nx:\n  class: StarfitFlowNode\n  parameters: myparam.param\n")
    with pytest.raises(NotImplementedError):
        pwsb.Model.from_yaml(tmp_path / "bad.yaml")
    # PRMSEt is a process of its own (hydrology/prms_et.py), not mixed into the NHM chain
    (tmp_path / "et.yaml").write_text("control: control.yaml\nx:\n  class: PRMSEt\n  parameters: myparam.param\n")
    assert pwsb.Model.from_yaml(tmp_path / "et.yaml")._mode == "et"


# ---------------------------------------------------------------- netCDF-4 (HDF5) inputs
REF_TEST_DRB = pl.Path("/root/reference/test_data/drb_2yr")
needs_ref_nc = pytest.mark.skipif(not (REF_TEST_DRB / "parameters_PRMSSnow.nc").exists(),
                                  reason="reference netCDF files absent")


@needs_ref_nc
def test_hdf5_reader_on_the_reference_parameter_files(drb):
    """The reference's per-process parameter files (netCDF-4 = HDF5: superblock
    v2, dense links in a fractal heap, dense attributes) read by our own reader
    equal the PRMS ASCII parameter file value for value."""
    from pywatershed_b200 import Parameters
    from pywatershed_b200.netcdf_out import nc4_module

    assert nc4_module() is None  # this container has no netCDF4: the HDF5 reader is what runs
    p, f, start = drb
    checked = 0
    for fn in sorted(REF_TEST_DRB.glob("parameters_*.nc")):
        par = Parameters.from_netcdf(fn)
        assert par.dims.get("nhru", 765) == 765 and par.dims.get("nsegment", 456) == 456
        for k, v in par.to_dict().items():
            if k in p:
                np.testing.assert_array_equal(np.asarray(v).squeeze(), np.asarray(p[k]).squeeze(), err_msg=k)
                checked += 1
        if fn.name == "parameters_PRMSSnow.nc":
            assert par.metadata["tmax_allsnow"]["dims"] == ("nmonth", "nhru")
            assert par.metadata["snarea_curve"]["dims"] == ("ndeplval",)
    assert checked > 150


@needs_ref_nc
def test_hdf5_reader_on_the_reference_forcing_files(drb):
    """prcp.nc / tmax.nc / tmin.nc: chunked + shuffle + deflate float32 [time,
    nhm_id] with a CF time axis; equal to the CBH ASCII values rounded to f32."""
    from pywatershed_b200.netcdf_in import open_netcdf, read_time_variable

    p, f, start = drb
    for k in ("prcp", "tmax", "tmin"):
        times, vals = read_time_variable(REF_TEST_DRB / f"{k}.nc", k)
        assert times[0] == np.datetime64(start) and times[-1] == np.datetime64(start) + 730
        assert vals.dtype == np.float64 and vals.shape == (731, 765)
        np.testing.assert_array_equal(vals.astype(np.float32), f[k].astype(np.float32))
        # ... or as stored (what pws_run_host_f32 uploads): float32, the same numbers
        _, stored = read_time_variable(REF_TEST_DRB / f"{k}.nc", k, keep_f32=True)
        assert stored.dtype == np.float32
        np.testing.assert_array_equal(stored.astype(np.float64), vals)
    with open_netcdf(REF_TEST_DRB / "prcp.nc") as ds:
        assert ds.dims == {"nhm_id": 765, "time": 731}
        assert ds.variables["prcp"].dims == ("time", "nhm_id")
        assert ds.variables["prcp"].attrs["units"] == "in"
        np.testing.assert_array_equal(ds.variables["nhm_id"].read(), p["nhm_id"])
    # test_data/hru_1/cbh.nc: eight variables in one file, unlimited time axis, `datetime` coordinate
    hru1 = REF_TEST_DRB.parent / "hru_1" / "cbh.nc"
    times, vals = read_time_variable(hru1, "tmax")
    assert vals.shape == (14975, 1) and times[0] == np.datetime64("1979-01-01") and \
        times[-1] == np.datetime64("2019-12-31")
    z = np.load(scenarios.GOLDEN / "hru1_inputs.npz")
    np.testing.assert_array_equal(vals.astype(np.float32), z["tmax"].astype(np.float32))


@needs_ref_nc
def test_model_from_the_reference_yaml(drb):
    """Model.from_yaml on the reference's OWN test_data/drb_2yr/nhm_model.yaml
    (control yaml + netCDF-4 discretization / parameter files), unmodified."""
    import pywatershed_b200 as pwsb

    p, f, start = drb
    md = pwsb.Model.model_dict_from_yaml(REF_TEST_DRB / "nhm_model.yaml")
    md["control"].options["input_dir"] = REF_TEST_DRB  # as autotest/test_prms_above_snow.py:128-170 does
    with pytest.warns(UserWarning, match="calc_method='numba'"):  # the reference's control asks for numba
        m = pwsb.Model(md)
    assert m._mode == "chain+channel" and len(m.processes) == 8
    assert m.processes["PRMSSnow"]._calc_method == "cuda"
    # the forcings the engine would be given: input_dir/<name>.nc, sliced to the control period
    for k in ("prcp", "tmax", "tmin"):
        a = m._load_input(k, 765)
        assert a.dtype == np.float64 and a.shape == (731, 765)
        np.testing.assert_array_equal(a.astype(np.float32), f[k].astype(np.float32))
        a32 = m._load_input(k, 765, keep_f32=True)  # the full chain keeps the stored type
        assert a32.dtype == np.float32 and a32.flags.c_contiguous
        np.testing.assert_array_equal(a32.astype(np.float64), a)
    assert m.control.n_times == 731 and m.control.options["dprst_flag"] is True
    for k in ("soil_moist_max", "tmax_cbh_adj", "tosegment", "snarea_curve", "hru_segment", "mann_n"):
        np.testing.assert_array_equal(np.asarray(m._param_dict[k]).squeeze(), np.asarray(p[k]).squeeze(),
                                      err_msg=k)


@needs_ref_files
def test_control_prms_to_yaml_round_trip(tmp_path):
    """autotest/test_control.py:203-211."""
    from pywatershed_b200 import Control

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        c = Control.load_prms(REF_DRB / "nhm.control", warn_unused_options=False)
    c.to_yaml(tmp_path / "control.yaml")
    c2 = Control.from_yaml(tmp_path / "control.yaml")
    a, b = c.to_dict(), c2.to_dict()
    a["options"] = {k: (str(v) if isinstance(v, pl.Path) else v) for k, v in a["options"].items()}
    b["options"] = {k: (str(v) if isinstance(v, pl.Path) else v) for k, v in b["options"].items()}
    np.testing.assert_equal(a, b)


@needs_ref_nc
def test_hdf5_reader_sweeps_every_reference_netcdf_file():
    """Every .nc file shipped with the reference (87 files: superblock v0 and
    v2, compact / dense groups, contiguous / chunked / compressed data, enum and
    string variables) opens, and every variable reads with its declared shape."""
    from pywatershed_b200.netcdf_in import open_netcdf

    files = sorted(pl.Path("/root/reference").rglob("*.nc"))
    assert len(files) >= 80
    nvar = 0
    for fn in files:
        with open_netcdf(fn) as ds:
            for k, v in ds.variables.items():
                a = v.read()
                assert a.shape == tuple(v.shape) and len(v.dims) == a.ndim, (fn, k)
                assert all(ds.dims[d] >= n for d, n in zip(v.dims, a.shape)), (fn, k)
                nvar += 1
    assert nvar > 800


def test_untile_reads_the_tiled_result_layout():
    """Engine.tiled_view / untile against the layout include/pws_b200.h documents
    for pws_output_tile: element (day d, variable k, unit i) of a tiled buffer is
    at [d][i // H][k][i % H]."""
    from pywatershed_b200.engine import untile

    nd, nhru, nvar, H = 3, 765, 5, 128
    tiles = -(-nhru // H)
    rows = np.arange(nd * nvar * nhru, dtype=np.float64).reshape(nd, nvar, nhru)
    buf = np.full((nd, tiles, nvar, H), -1.0)
    for i in range(nhru):
        buf[:, i // H, :, i % H] = rows[:, :, i]
    for k in range(nvar):
        np.testing.assert_array_equal(untile(buf, k, nhru), rows[:, k, :])
    import torch

    np.testing.assert_array_equal(untile(torch.from_numpy(buf), 2, nhru).numpy(), rows[:, 2, :])


def test_streaming_classic_writer_against_scipy(tmp_path):
    """cdf_out.Cdf2Writer (the output backend without the netCDF4 module) read
    back by an independent implementation, scipy.io.netcdf_file: several record
    variables of different types in one file written block by block (one
    contiguous write per block), a block with one variable missing (row-wise
    path), blocks out of order, a fixed-size `doy` file, attributes with the
    reference's names, and numrecs brought up to date by flush()."""
    from scipy.io import netcdf_file

    from pywatershed_b200.netcdf_out import _File, nc4_module

    if nc4_module() is not None:
        pytest.skip("netCDF4 is the backend here")
    rng = np.random.default_rng(5)
    n, nt = 37, 11
    a = rng.normal(size=(nt, n))
    b = rng.integers(-5, 5, size=(nt, n)).astype(np.int32)
    c = rng.integers(0, 2, size=(nt, n)).astype(np.bool_)
    names = ["pkwater_equiv", "newsnow", "iasw"]  # float64, int32, bool in the metadata
    ids = np.arange(100, 100 + n)
    f = _File(tmp_path / "PRMSSnow.nc", names, "nhm_id", ids, "PRMSSnow")
    t = 3287.0 + np.arange(nt)
    f.write(4, 8, t[4:8], {"pkwater_equiv": a[4:8], "newsnow": b[4:8], "iasw": c[4:8]})  # out of order
    f.write(0, 4, t[0:4], {"pkwater_equiv": a[0:4], "newsnow": b[0:4], "iasw": c[0:4]})
    f.ds.flush()
    with netcdf_file(str(tmp_path / "PRMSSnow.nc"), "r", mmap=False) as ds:
        assert ds.variables["time"].shape == (8,)  # numrecs is in the header before close
    f.write(8, nt, t[8:], {"pkwater_equiv": a[8:], "iasw": c[8:]})  # newsnow missing: rows 8.. stay zero
    f.close()
    with netcdf_file(str(tmp_path / "PRMSSnow.nc"), "r", mmap=False) as ds:
        assert ds.Description == b"pywatershed output data"
        assert ds._attributes["process class"] == b"PRMSSnow"
        assert ds.dimensions == {"time": None, "nhm_id": n}
        v = ds.variables
        assert v["pkwater_equiv"].dimensions == ("time", "nhm_id") and v["pkwater_equiv"].units == b"inches"
        assert v["pkwater_equiv"][:].dtype == np.dtype(">f8") and v["newsnow"][:].dtype == np.dtype(">i4")
        np.testing.assert_array_equal(v["pkwater_equiv"][:], a)
        np.testing.assert_array_equal(v["newsnow"][:8], b[:8])
        np.testing.assert_array_equal(v["newsnow"][8:], 0)
        np.testing.assert_array_equal(v["iasw"][:], c.astype(np.int32))
        np.testing.assert_array_equal(v["time"][:], t.astype(np.float32))
        np.testing.assert_array_equal(v["nhm_id"][:], ids)
    # our own reader opens it too (classic files go through scipy there)
    from pywatershed_b200.netcdf_in import read_time_variable

    times, vals = read_time_variable(tmp_path / "PRMSSnow.nc", "pkwater_equiv")
    np.testing.assert_array_equal(vals, a)
    assert times[0] == np.datetime64("1979-01-01")
    # fixed-size file: the soltab tables on a `doy` axis
    g = _File(tmp_path / "soltab_potsw.nc", ["soltab_potsw"], "nhm_id", ids, "PRMSSolarGeometry",
              time_name="doy", n_doy=366)
    s = rng.normal(size=(366, n))
    g.write(0, 366, np.arange(1, 367), {"soltab_potsw": s})
    g.close()
    with netcdf_file(str(tmp_path / "soltab_potsw.nc"), "r", mmap=False) as ds:
        assert ds.dimensions == {"doy": 366, "nhm_id": n}
        np.testing.assert_array_equal(ds.variables["soltab_potsw"][:], s)
        np.testing.assert_array_equal(ds.variables["doy"][:], np.arange(1, 367, dtype=np.float32))
        np.testing.assert_array_equal(ds.variables["nhm_id"][:], ids)
    # a single record variable besides nothing else: the unpadded lone-record case
    from pywatershed_b200.cdf_out import Cdf2Writer

    w = Cdf2Writer(tmp_path / "lone.nc")
    w.createDimension("time", None)
    w.createDimension("x", 3)
    x = w.createVariable("x", "i4", ("x",))
    q = w.createVariable("q", "f4", ("time", "x"), units="cfs")
    x[:] = [7, 8, 9]
    q[0:2] = np.array([[1, 2, 3], [4, 5, 6]], dtype=np.float64)
    q[2] = [7.5, 8.5, 9.5]
    w.close()
    with netcdf_file(str(tmp_path / "lone.nc"), "r", mmap=False) as ds:
        np.testing.assert_array_equal(ds.variables["q"][:], [[1, 2, 3], [4, 5, 6], [7.5, 8.5, 9.5]])
        np.testing.assert_array_equal(ds.variables["x"][:], [7, 8, 9])
        assert ds.variables["q"].units == b"cfs"


def test_model_netcdf_output_on_a_stubbed_block(drb, tmp_path, monkeypatch):
    """ModelNetcdfOutput.write_days over a block of results that did not come
    from the GPU (the engine is a stub): the files and budget files of
    tests/test_gpu_model.py::test_netcdf_output_files, written block by block
    (in threads when the block is large) and read back with scipy."""
    import types

    from scipy.io import netcdf_file

    import pywatershed_b200 as pwsb
    from pywatershed_b200 import netcdf_out

    if netcdf_out.nc4_module() is not None:
        pytest.skip("netCDF4 is the backend here")
    p, f, start = drb
    nhru, nseg, nd = 765, 456, 9
    params = pwsb.Parameters(dims={"nhru": nhru, "nsegment": nseg},
                             coords={"nhm_id": np.arange(1, nhru + 1), "nhm_seg": np.arange(1, nseg + 1)},
                             data_vars=p)
    m = pwsb.Model([pwsb.PRMSSolarGeometry, pwsb.PRMSAtmosphere, pwsb.PRMSCanopy, pwsb.PRMSSnow,
                    pwsb.PRMSRunoff, pwsb.PRMSSoilzone, pwsb.PRMSGroundwater, pwsb.PRMSChannel],
                   control=_control(nd, budget_type="warn"), parameters=params, find_input_files=False)
    m._engine = types.SimpleNamespace(nhru=nhru, nseg=nseg)
    rng = np.random.default_rng(11)
    out_vars = ["pkwater_equiv", "seg_outflow", "newsnow", "iasw", "soltab_potsw"]
    snow_terms = [v for comp in m.processes["PRMSSnow"].budget.terms.values() for v in comp]
    blk = {v: rng.normal(size=(nd, nhru)) for v in set(snow_terms) | {"pkwater_equiv"}}
    blk["seg_outflow"] = rng.normal(size=(nd, 3, nseg))[:, 1, :]  # a strided view, as the engine hands out
    blk["newsnow"] = rng.integers(0, 2, size=(nd, nhru)).astype(np.int32)
    blk["iasw"] = rng.integers(0, 2, size=(nd, nhru)).astype(np.bool_)
    m._block, m._block_t0 = blk, 0
    monkeypatch.setattr(netcdf_out, "_THREAD_BYTES", 1)  # take the threaded path at this size too
    out = netcdf_out.ModelNetcdfOutput(m, tmp_path, True, out_vars, None)
    out.write_days(m, 0, 4)
    out.write_days(m, 4, nd)
    out.close()

    def read(name, var):
        with netcdf_file(str(tmp_path / name), "r", mmap=False) as ds:
            return np.array(ds.variables[var][:]), ds.variables[var].dimensions

    for v in ("pkwater_equiv", "seg_outflow", "newsnow", "iasw"):
        a, dims = read(f"{v}.nc", v)
        assert dims == ("time", "nhm_seg" if v == "seg_outflow" else "nhm_id")
        np.testing.assert_array_equal(a, np.asarray(blk[v]).astype(a.dtype), err_msg=v)
    a, dims = read("soltab_potsw.nc", "soltab_potsw")
    assert dims == ("doy", "nhm_id") and a.shape == (366, nhru)
    # coordinate and attribute conventions of the reference's NetCdfWrite
    # (utils/netcdf_utils.py:474-489, 600-616): doy is i4 "Day of year", every scalar
    # metadata item of a variable is an attribute, _FillValue is netCDF's default
    with netcdf_file(str(tmp_path / "soltab_potsw.nc"), "r", mmap=False) as ds:
        doy = ds.variables["doy"]
        assert doy.data.dtype.kind == "i" and doy.units == b"Day of year"
        np.testing.assert_array_equal(doy[:], np.arange(1, 367))
        assert ds.variables["soltab_potsw"].units == b"Langleys"
    with netcdf_file(str(tmp_path / "pkwater_equiv.nc"), "r", mmap=False) as ds:
        v = ds.variables["pkwater_equiv"]
        assert v.var_category == b"mass storage" and v.type == b"float64" and v.units == b"inches"
        assert v._FillValue == pytest.approx(9.969209968386869e36)
        assert ds.variables["time"].units == b"days since 1970-01-01 00:00:00"
    with netcdf_file(str(tmp_path / "newsnow.nc"), "r", mmap=False) as ds:
        assert ds.variables["newsnow"]._FillValue == -2147483647
    b = m.processes["PRMSSnow"].budget
    for comp, name in (("inputs", "inputs_sum"), ("outputs", "outputs_sum"),
                       ("storage_changes", "storage_changes_sum")):
        a, dims = read("PRMSSnow_budget.nc", name)
        assert dims == ("time", "nhm_id")
        np.testing.assert_allclose(a, sum(blk[v] for v in b.terms[comp]), rtol=0, atol=1e-12)
    t, _ = read("PRMSSnow_budget.nc", "time")
    day0 = (m.control.start_time.astype("datetime64[D]") - np.datetime64("1970-01-01", "D")).astype(int)
    np.testing.assert_array_equal(t, (day0 + np.arange(nd)).astype(np.float32))
    # budget_args (base/budget.py:661-706): chosen sums, and every term of every component
    out = netcdf_out.ModelNetcdfOutput(m, tmp_path / "b2", True, ["pkwater_equiv"],
                                       {"write_sum_vars": ["outputs_sum"], "write_individual_vars": True})
    out.write_days(m, 0, nd)
    out.close()
    with netcdf_file(str(tmp_path / "b2" / "PRMSSnow_budget.nc"), "r", mmap=False) as ds:
        assert "outputs_sum" in ds.variables and "inputs_sum" not in ds.variables
        assert ds.inputs == ("[" + ", ".join(b.terms["inputs"]) + "]").encode()
        assert getattr(ds, "Budget basis") == b"unit (unit or global)"
        for comp in b.components:
            for v in b.terms[comp]:
                np.testing.assert_array_equal(ds.variables[f"{comp}__{v}"][:], blk[v])
        assert ds.variables["outputs_sum"].units == b"inches"
    with pytest.warns(UserWarning, match="no requested output"):
        out = netcdf_out.ModelNetcdfOutput(m, tmp_path / "b3", True, ["pkwater_equiv"],
                                           {"write_sum_vars": False})
    assert not out.budget_files
    out.close()


@needs_ref_nc
def test_hdf5_reader_threads_give_the_same_arrays(monkeypatch):
    """Large chunked variables are inflated / unshuffled by several threads
    (hdf5_min._THREAD_BYTES); forced on for the reference's UCB forcing file
    (25 chunks, shuffle + deflate) it returns the serial reader's array."""
    from pywatershed_b200 import hdf5_min
    from pywatershed_b200.netcdf_in import read_time_variable

    path = REF_TEST_DRB.parent / "ucb_2yr" / "prcp.nc"
    monkeypatch.setattr(hdf5_min, "_THREAD_BYTES", 1 << 60)
    t0, serial = read_time_variable(path, "prcp", keep_f32=True)
    monkeypatch.setattr(hdf5_min, "_THREAD_BYTES", 1)
    t1, threaded = read_time_variable(path, "prcp", keep_f32=True)
    assert serial.shape == (731, 3851) and serial.dtype == np.float32
    np.testing.assert_array_equal(threaded, serial)
    np.testing.assert_array_equal(t0, t1)
