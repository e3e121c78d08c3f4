"""Single processes and partial chains, driven the way the reference's own
tests drive them (autotest/test_self_drive.py:83-115: every process alone,
its inputs read from the outputs of a full run; autotest/test_prms_snow.py:
122-132: control.advance(); proc.advance(); proc.calculate(1.0)).  The inputs
here are the oracle's full-chain outputs; each process must reproduce its own
variables of the full run.  GPU only, through the C ABI."""
import numpy as np
import pytest

import prms_oracle as po
from test_gpu_model import _control, _parameters

pytestmark = pytest.mark.gpu

ND = 90


@pytest.fixture(scope="module")
def full(drb):
    p, f, start = drb
    o = po.Oracle(p, start=start)
    ref, _ = o.run(f["prcp"][:ND], f["tmax"][:ND], f["tmin"][:ND], po.HRU_VARS, po.SEG_VARS)
    ref = {k: np.asarray(v, dtype=np.float64) for k, v in ref.items()}
    # Snow sees the pptmix Canopy edited in place (prms_canopy.py:500-505); the
    # variable Atmosphere publishes is the un-edited one
    ref["pptmix_for_snow"] = np.where((ref["hru_snow"] > 0) & (ref["net_snow"] == 0), 0.0, ref["pptmix"])
    return ref


def _run(drb, full, names, skip=(), stepwise=False):
    import pywatershed_b200 as pwsb

    p, f, start = drb
    control = _control(start, ND)
    classes = [getattr(pwsb, n) for n in names]
    m = pwsb.Model(classes, control=control, parameters=_parameters(p), find_input_files=False)
    produced = set()
    for proc in m.processes.values():
        produced.update(proc.get_variables())
    for proc in m.processes.values():
        for k in proc.get_inputs():
            if k in produced or k.startswith("soltab"):
                continue
            src = f[k][:ND] if k in f else full["pptmix_for_snow" if (
                k == "pptmix" and proc.name == "PRMSSnow") else k]
            proc.set_input_to_adapter(k, src)
    series = {}
    for it in range(ND):
        m.advance()
        m.calculate()
        for proc in m.processes.values():
            for v in proc.get_variables():
                if v in skip or v.startswith("soltab"):
                    continue
                a = proc[v]
                a = a.current if hasattr(a, "current") else a
                series.setdefault(v, []).append(np.array(a, dtype=np.float64))
    for v, rows in series.items():
        np.testing.assert_allclose(np.stack(rows), full[v], rtol=1e-9, atol=1e-10, err_msg=v)
    return m


@pytest.mark.parametrize("name,skip", [
    ("PRMSCanopy", ()), ("PRMSSnow", ()), ("PRMSRunoff", ("sroff", "sroff_vol")),
    ("PRMSSoilzone", ()), ("PRMSGroundwater", ())])
def test_each_process_alone_reproduces_the_full_run(drb, full, name, skip):
    m = _run(drb, full, [name], skip)
    assert m.processes[name].budget._zero_sum


def test_partial_chains_of_the_reference_examples(drb, full):
    # base/model.py:121-163 (GW + Channel) and :165-219 (the four upper processes)
    _run(drb, full, ["PRMSGroundwater", "PRMSChannel"])
    _run(drb, full, ["PRMSSolarGeometry", "PRMSAtmosphere", "PRMSCanopy", "PRMSSnow"],
         skip=("pptmix",))  # in memory the reference shows Canopy's edit; the file value is compared elsewhere


def test_process_stepped_on_its_own(drb, full):
    """autotest/test_prms_snow.py:122-132"""
    import pywatershed_b200 as pwsb

    p, f, start = drb
    control = _control(start, ND)
    gw = pwsb.PRMSGroundwater(control, None, _parameters(p), soil_to_gw=full["soil_to_gw"],
                              ssr_to_gw=full["ssr_to_gw"], dprst_seep_hru=full["dprst_seep_hru"],
                              budget_type="error")
    for it in range(20):
        control.advance()
        gw.advance()
        gw.calculate(1.0)
        np.testing.assert_allclose(gw["gwres_stor"], full["gwres_stor"][it], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(gw.gwres_flow_vol, full["gwres_flow_vol"][it], rtol=1e-9, atol=1e-10)
    assert gw.budget._zero_sum


def test_prms_et_alone(drb):
    """PRMSEt (hydrology/prms_et.py:22-181) driven on its own from the other
    processes' outputs, as autotest/test_prms_et.py:29-110 does: bit-identical to
    the reference's arithmetic restated in numpy (there is no transcendental in
    it), and its hru_actet agrees with PRMSSoilzone's own within the tolerance of
    the reference's test (:99-107)."""
    import pywatershed_b200 as pwsb
    from test_gpu_model import _control, _parameters

    p, f, start = drb
    nd = 120
    o = po.Oracle(p, start=start)
    names = list(pwsb.PRMSEt.get_inputs()) + ["hru_actet", "unused_potet"]
    ref, _ = o.run(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd], names, ())
    control = _control(start, nd, block_days=50)
    et = pwsb.PRMSEt(control, None, _parameters(p), budget_type=None,
                     **{k: ref[k] for k in pwsb.PRMSEt.get_inputs()})
    frac_perv = 1.0 - p["hru_percent_imperv"] - p["dprst_frac"]
    for it in range(nd):
        control.advance()
        et.advance()
        et.calculate(1.0)
        hpa = ref["perv_actet"][it] * frac_perv
        loss = ref["hru_intcpevap"][it].copy()
        for k in ("snow_evap", "dprst_evap_hru", "hru_impervevap"):
            loss += ref[k][it]
        loss += hpa
        unused = np.maximum(ref["potet"][it] - loss, 0.0)
        np.testing.assert_array_equal(et["hru_perv_actet"], hpa)
        np.testing.assert_array_equal(et["unused_potet"], unused)
        np.testing.assert_array_equal(et["hru_actet"], ref["potet"][it] - unused)
        np.testing.assert_allclose(et["hru_actet"], ref["hru_actet"][it], rtol=1e-6, atol=1e-9)
    assert id(et["hru_actet"]) == id(et.hru_actet)
