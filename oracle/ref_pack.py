"""oracle/_ref: the UNMODIFIED reference package, packed for the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY (never imported by pywatershed_b200/).

`/root/reference` exists in the build container only.  north_star asks for the
reference's numpy and numba paths "timed on the GPU box's own host cores in the
same run", so `__graft_entry__.build()` calls `pack()` here, which copies -- does
not modify -- the reference's pure-Python package (its .py files, the metadata
YAML under static/, and the ASCII drb_2yr inputs: nhm.control, myparam.param,
{prcp,tmax,tmin}.cbh) into `oracle/_ref/pywatershed/`.  `oracle/_ref/` is
git-ignored (no reference source enters the history) but travels to the GPU box
with the snapshot, like the built .so files.  `bench.py --impl reference --ref
numpy|numba` then imports it under the same import stubs as
oracle/ref_harness.py (netCDF4 / xarray / ... are absent from the image and the
ASCII input path does not need them) and times `pws.Model.run()` on drb_2yr,
the recipe of asv_benchmarks/benchmarks/prms.py:146-171.
"""
from __future__ import annotations

import pathlib as pl
import shutil
import sys
import time

HERE = pl.Path(__file__).resolve().parent
REF_SRC = pl.Path("/root/reference")
REF_DST = HERE / "_ref"
DRB_FILES = ("nhm.control", "myparam.param", "prcp.cbh", "tmax.cbh", "tmin.cbh")


def available() -> bool:
    return (REF_DST / "pywatershed" / "__init__.py").exists() and all(
        (REF_DST / "pywatershed" / "data" / "drb_2yr" / f).exists() for f in DRB_FILES)


def pack(force: bool = False) -> bool:
    """Copy the reference package into oracle/_ref (only where /root/reference
    exists).  Returns True when oracle/_ref is usable afterwards."""
    src = REF_SRC / "pywatershed"
    if not (src / "__init__.py").exists():
        return available()
    if available() and not force:
        return True
    dst = REF_DST / "pywatershed"
    if dst.exists():
        shutil.rmtree(dst)
    for path in src.rglob("*"):
        rel = path.relative_to(src)
        if path.is_dir() or "__pycache__" in rel.parts:
            continue
        top = rel.parts[0]
        keep = path.suffix == ".py" or top == "static" and path.suffix in (".yaml", ".yml")
        if top == "data":
            keep = path.suffix == ".py" or (rel.parts[1:2] == ("drb_2yr",) and path.name in DRB_FILES)
        if keep:
            out = dst / rel
            out.parent.mkdir(parents=True, exist_ok=True)
            shutil.copyfile(path, out)
    for extra in ("LICENSE", "version.txt"):
        if (REF_SRC / extra).exists():
            shutil.copyfile(REF_SRC / extra, REF_DST / extra)
    return available()


def _import():
    import ref_harness as rh

    rh.REF_ROOT = REF_DST
    rh.DRB_DIR = REF_DST / "pywatershed" / "data" / "drb_2yr"
    return rh, rh.import_reference()


def time_drb(calc_method: str = "numpy", ndays: int = 731, warm: bool = False) -> dict:
    """Seconds of pws.Model.run() for the drb_2yr NHM chain (8 processes, budgets
    on, no netCDF output) with `calc_method`; `warm`: run 5 days first so numba's
    JIT compilation is outside the timed region."""
    if str(HERE) not in sys.path:
        sys.path.insert(0, str(HERE))
    rh, pws = _import()
    import numpy as np

    def model(n):
        control, params, forcings, times = rh.load_drb(pws)
        control.edit_n_time_steps(n)
        forcings = {k: v[:n] for k, v in forcings.items()}
        return rh.build_model(pws, control, params, forcings, times[:n], calc_method=calc_method)

    # numba compiles its kernels per process object on the first step: the first
    # `skip` days of the same model are run untimed ("JIT warm-up
    # excluded (run 5 days first, then time)"), for numpy too so the arms agree
    skip = 5 if warm else 0
    m = model(ndays)
    if skip:
        m.run(n_time_steps=skip, finalize=False)
    t0 = time.perf_counter()
    m.run(n_time_steps=ndays - skip, finalize=False)
    dt = time.perf_counter() - t0
    return {"seconds": dt, "ndays": ndays - skip, "calc_method": calc_method,
            "version": getattr(pws, "__version__", "?"),
            "sum_seg_outflow": float(np.sum(m.processes["PRMSChannel"]["seg_outflow"])),
            "sum_pkwater_equiv": float(np.sum(m.processes["PRMSSnow"]["pkwater_equiv"]))}


if __name__ == "__main__":
    print("packed:", pack(force="--force" in sys.argv))
    if "--time" in sys.argv:
        print(time_drb(sys.argv[sys.argv.index("--time") + 1], 60, True))
