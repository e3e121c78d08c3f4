"""Extract the per-process descriptors of the reference (names only: inputs,
variables + initial values, parameters, dimensions, mass budget terms) and the
netCDF metadata of every public variable into pywatershed_b200/_process_meta.py.
Runs only in the build container (needs /root/reference)."""
import pathlib as pl
import pprint
import sys

import numpy as np

HERE = pl.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import ref_harness as rh  # noqa: E402

pws = rh.import_reference()
out = {}
varmeta = {}
NODPRST = ("PRMSRunoffNoDprst", "PRMSSoilzoneNoDprst", "PRMSGroundwaterNoDprst")
for name in rh.NHM_PROCESS_NAMES + NODPRST + ("PRMSEt",):
    cls = getattr(pws, name)
    init = {}
    for k, v in cls.get_init_values().items():
        v = float(v)
        init[k] = None if np.isnan(v) else v
    d = {
        "dimensions": tuple(cls.get_dimensions()),
        "parameters": tuple(cls.get_parameters()),
        "inputs": tuple(cls.get_inputs()),
        "variables": tuple(cls.get_variables()),
        "init_values": init,
    }
    if hasattr(cls, "get_mass_budget_terms"):
        d["mass_budget_terms"] = {k: list(v) for k, v in cls.get_mass_budget_terms().items()}
    if hasattr(cls, "get_restart_variables"):
        try:
            d["restart_variables"] = tuple(cls.get_restart_variables())
        except Exception:
            d["restart_variables"] = ()
    out[name] = d
    for v, m in pws.meta.find_variables(cls.get_variables()).items():
        varmeta[v] = {k: (list(x) if isinstance(x, (tuple, list)) else x) for k, x in m.items()
                      if k in ("desc", "units", "type", "dims", "var_category")}
# the drb control file's output list, intersected with the public variables
control = pws.Control.load_prms(rh.DRB_DIR / "nhm.control", warn_unused_options=False)
allv = set(v for d in out.values() for v in d["variables"])
nhm_default = [v for v in control.options["netcdf_output_var_names"] if v in allv]
# (the reference builds that option from a set: its order changes from run to run.  Keep the
# committed order when the content is unchanged, so fixtures and bench selections stay stable.)
_prev = HERE.parent / "tests" / "golden" / "nhm_default_vars.json"
if _prev.exists():
    import json as _json

    _old = _json.loads(_prev.read_text())
    if set(_old) == set(nhm_default):
        nhm_default = _old
txt = ('"""Machine-extracted (gen_process_meta.py, build container) from the reference\'s\n'
       'get_inputs()/get_variables()/get_init_values()/get_parameters()/\n'
       'get_mass_budget_terms() and static/metadata/variables.yaml.  Data only."""\n'
       "PROCESS_META = " + pprint.pformat(out, width=100, sort_dicts=False) + "\n\n"
       "VARIABLE_META = " + pprint.pformat(varmeta, width=100, sort_dicts=False) + "\n\n"
       "NHM_DEFAULT_OUTPUT_VARS = " + pprint.pformat(nhm_default, width=100) + "\n")
(HERE.parent / "pywatershed_b200" / "_process_meta.py").write_text(txt)
import json
(HERE.parent / "tests" / "golden" / "nhm_default_vars.json").write_text(json.dumps(nhm_default))
print({k: (len(v["inputs"]), len(v["variables"]), len(v["parameters"])) for k, v in out.items()}, len(nhm_default))
