"""bench.py -- simulated HRU-days/s of the full NHM chain on B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W]
                    [--workload conus|ensemble|ucb|drb|channel|channel_chained]
                    [--scaling strong|weak] [--outputs full|nhm_default|routing]
                    [--impl reference [--ref port|numpy|numba]]

A *step* is one pass of the hot path over the whole workload: reset to the
initial conditions, then every day of the run (time-blocked).  Default
workload `conus`: the synthetic CONUS-scale domain of BASELINE.json configs[2]
-- 144 DRB tiles in different climates = 110,160 HRUs / 65,664 segments,
3,653 days (1979-01-01 .. 1988-12-31), DRB forcings cycled with a per-tile
temperature offset / precipitation scale (tests/scenarios.py:tile).  It is the
configuration north_star's target is quoted on and fits one GPU; configs[1]
(3,851 HRUs) fills 30 of 148 SMs and is available as `--workload ucb`.

  value   HRU-days/s of the whole job, forcings already resident in HBM, every
          public variable (143 HRU + 5 segment) written to HBM each day
  e2e     the same domain through the host-facing entry point (pws_run_host_f32):
          forcings in pinned HOST memory, H2D every step, streamflow + budget
          verdicts drained D2H every step
  N > 1   STRONG scaling, as BASELINE.json configs[2] says ("basin-sharded
          1/2/4/8 GPU"): the ONE 110,160-HRU domain is partitioned by outlet tree
          (pywatershed_b200/sharding.py), every rank simulates its own basins
          with no data-path collective, NCCL gathers the budget verdicts and a
          streamflow checksum.  `--scaling weak` is the round-1 replica mode
          (every rank its own 144-tile set).
  extra   objects in the same JSON line: `ensemble` (configs[3]: the 1024-member
          DRB parameter ensemble, members perturbed as SURVEY.md 8d says, the
          members' common forcings uploaded once, sharded by member),
          `e2e_nhm_default` (the reference control file's 74 output variables
          drained to the host), `pcie` (the box's plain cudaMemcpyAsync ceiling),
          `e2e_model` (the Model API, N = 1).

`--impl reference` times the reference's CPU implementation on the host cores:
`--ref port` (default) the C restatement oracle/prms_oracle.c on every core;
`--ref numpy|numba` the unmodified pywatershed package packed into oracle/_ref
by oracle/pack_ref.py (drb_2yr, one process per core).
"""
from __future__ import annotations

import argparse
import json
import os
import pathlib as pl
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = pl.Path(__file__).resolve().parent
for p in (ROOT, ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

METRIC = "simulated HRU-days/sec, full NHM chain"
UNIT = "HRU-days/s"
BYTES_READ = 40  # prcp,tmax,tmin + 2 soltab rows, SURVEY.md 8(d)
NB_HRU, NB_SEG = 765, 456  # drb_2yr


def workload_spec(name):
    if name == "conus":
        return dict(ntile=144, nhru_keep=None, ndays=3653, label="synthetic CONUS-scale NHM domain "
                    "(144 DRB tiles: 110,160 HRUs, 65,664 segments, 3,653 days)")
    if name == "ucb":
        return dict(ntile=6, nhru_keep=3851, ndays=731, label="ucb_2yr-shaped synthetic "
                    "(3,851 HRUs, 2,736 segments, 731 days; real ucb_2yr inputs are HDF5 blobs)")
    if name in ("ensemble", "ensemble128"):
        nm = 1024 if name == "ensemble" else 128
        return dict(ntile=nm, nhru_keep=None, ndays=731, ensemble=True,
                    label=f"{nm}-member DRB parameter ensemble ({nm * NB_HRU:,} HRUs, {nm * NB_SEG:,} "
                          "segments member-major, perturbed soil/snow parameters, shared forcings, 731 days)")
    if name == "drb":
        return dict(ntile=1, nhru_keep=None, ndays=731, label="drb_2yr (765 HRUs, 456 segments, 731 days)")
    raise SystemExit(f"unknown workload {name}")


def build_domain(spec, rank=0):
    """Parameters (host, small) + base forcings of the whole tiled domain."""
    import scenarios

    p0, f0, start = scenarios.load_drb()
    ntile = spec["ntile"]
    # parameters tiled on the host; forcings tiled on the device (see below)
    dummy = {k: v[:2] for k, v in f0.items()}
    pt, _ = scenarios.tile(p0, dummy, ntile, nhru_keep=spec["nhru_keep"])
    return p0, f0, pt, start


def tile_offsets(ntile, rank):
    t = np.arange(ntile) + rank * ntile
    dtemp = -15.0 + 25.0 * ((t * 7) % 16) / 15.0
    pscale = 0.6 + 0.1 * ((t * 5) % 9)
    return dtemp, pscale


def device_forcings(f0, spec, rank, dev, hru_idx=None):
    """[ndays, n] float64 tensors on `dev`: DRB CBH cycled in time, tiled in
    space with the per-tile climate offsets of tests/scenarios.py:tile, rounded
    to float32-representable values: the reference's forcing files are float32
    netCDF (test_data/{drb_2yr,ucb_2yr}/prcp.nc), and the e2e leg uploads them
    in that type (pws_run_host_f32).  `hru_idx`: the HRUs of the tiled domain
    to build columns for (a basin shard); default all."""
    import torch

    nd, ntile = spec["ndays"], spec["ntile"]
    nb = f0["prcp"].shape[1]
    nh = spec["nhru_keep"] or nb * ntile
    if hru_idx is None:
        hru_idx = np.arange(nh)
    hru_idx = np.asarray(hru_idx, dtype=np.int64)
    dtemp, pscale = tile_offsets(ntile, rank)
    tile_of = torch.from_numpy(hru_idx // nb).to(dev)
    col_of = torch.from_numpy(hru_idx % nb).to(dev)
    dt_h = torch.from_numpy(dtemp).to(dev)[tile_of]
    ps_h = torch.from_numpy(pscale).to(dev)[tile_of]
    out = {}
    tidx = torch.arange(nd, device=dev) % f0["prcp"].shape[0]
    for k in ("prcp", "tmax", "tmin"):
        base = torch.from_numpy(f0[k]).to(dev)[tidx]  # [nd, nb]
        full = base[:, col_of]
        full = full * ps_h[None, :] if k == "prcp" else full + dt_h[None, :]
        out[k] = full.float().double().contiguous()
        del full, base
    return out


# member m scales these parameters by U(0.8, 1.2)
# (seed = m), clipped to [minimum, maximum] of the reference's
# static/metadata/parameters.yaml; forcings are shared by the members.
ENSEMBLE_PARAMS = {
    "soil_moist_max": (1e-05, 20.0), "soil_rechr_max_frac": (1e-05, 1.0), "soil2gw_max": (0.0, 5.0),
    "ssr2gw_rate": (0.0001, 1.0), "slowcoef_lin": (0.0, 1.0), "slowcoef_sq": (0.0, 1.0),
    "smidx_coef": (0.0, 1.0), "carea_max": (0.0, 1.0), "gwflow_coef": (0.0, 0.5),
    "tmax_allsnow": (-10.0, 40.0), "freeh2o_cap": (0.01, 0.2), "emis_noppt": (0.757, 1.0),
    "rad_trncf": (0.0, 1.0), "cecn_coef": (0.02, 20.0), "potet_sublim": (0.1, 0.75),
}


def ensemble_parameters(p0, members):
    """Member-major parameters of the ensemble members `members` (global member
    numbers): per-HRU / per-segment arrays repeated, the ENSEMBLE_PARAMS
    scaled per member and HRU, tosegment / hru_segment offset per member."""
    import scenarios

    members = np.asarray(members)
    nm = members.size
    dummy = {k: np.zeros((2, NB_HRU)) for k in ("prcp", "tmax", "tmin")}
    pt, _ = scenarios.tile(p0, dummy, nm, temp_shift=False)
    for k in ENSEMBLE_PARAMS:
        pt[k] = np.array(pt[k], dtype=np.float64, copy=True)
    for j, m in enumerate(members):
        rng = np.random.default_rng(int(m))  # one stream per member, parameters in table order
        sl = slice(j * NB_HRU, (j + 1) * NB_HRU)
        for k, (lo, hi) in ENSEMBLE_PARAMS.items():
            v = pt[k]
            v[..., sl] = np.clip(v[..., sl] * rng.uniform(0.8, 1.2, size=v[..., sl].shape), lo, hi)
    return pt


def output_sets(reg, mode):
    if mode == "full":
        return list(reg.hru_vars), list(reg.seg_vars)
    if mode == "routing":
        return [], list(reg.seg_vars)
    if mode == "nhm_default":
        names = json.loads((ROOT / "tests" / "golden" / "nhm_default_vars.json").read_text())
        return [v for v in reg.hru_vars if v in names], list(reg.seg_vars)
    raise SystemExit(f"unknown outputs {mode}")


def algorithmic_bytes_per_hru_day(reg, hru_vars):
    size = {np.float64: 8, np.int32: 4, np.bool_: 1}  # the reference's dtypes
    return BYTES_READ + sum(size[reg.dtype(v)] for v in hru_vars)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def parity_check(device=0, ndays=731):
    """The checker beside the CPU baseline: the reference's own drb_2yr case
    (BASELINE.json configs[0]) through the C ABI against the oracle -- every
    public variable of every day at rtol 1e-9 / atol 1e-10, ghost-pack ties
    (tests/parity.py) counted and audited.  The oracle is only the checker here."""
    for q in (ROOT / "oracle", ROOT / "tests"):
        if str(q) not in sys.path:
            sys.path.insert(0, str(q))
    import parity
    import prms_oracle as po
    import scenarios
    from pywatershed_b200 import Engine

    p, f, start = scenarios.load_drb()
    nd = min(ndays, f["prcp"].shape[0])
    eng = Engine(p, start, device=device)
    eng.select_outputs(eng.reg.hru_vars, eng.reg.seg_vars)
    got, viol = eng.run_host(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd])
    eng.close()
    ref, oviol = po.Oracle(p, start=start).run(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd],
                                               po.HRU_VARS, po.SEG_VARS)
    rep = parity.compare_columns(got, ref, po.HRU_VARS, got["pkwater_equiv"], ref["pkwater_equiv"])
    out = {"domain": f"drb_2yr: 765 HRUs x {nd} days x {len(po.HRU_VARS)} HRU + {len(po.SEG_VARS)} segment "
                     "variables, GPU (C ABI) against the C oracle pinned to the reference",
           "rtol": parity.RTOL, "atol": parity.ATOL, "ghost_pack_ties": len(rep["ties"]),
           "mismatches": len(rep["mismatches"]), "max_post_tie_cum_rel": 0.0,
           "budget_verdicts_equal": bool((np.asarray(viol) == np.asarray(oviol)).all())}
    if rep["ties"]:
        try:
            out["max_post_tie_cum_rel"] = parity.audit_ties(got, ref, rep, got["pkwater_equiv"],
                                                            ref["pkwater_equiv"])["max_cum_rel"]
        except AssertionError as exc:
            out["audit_failed"] = str(exc)[:200]
    tied = [t[0] for t in rep["ties"]]
    try:
        seg = parity.assert_segments_with_ties(got, ref, po.SEG_VARS, p["tosegment"], p["hru_segment"], tied)
        out["segments_within_tolerance"] = True
        out["max_cum_seg_rel_downstream_of_ties"] = seg["max_cum_seg_rel"]
    except AssertionError as exc:
        out["segments_within_tolerance"] = False
        out["segment_error"] = str(exc)[:200]
    return out


def cpu_oracle_rate(p0, f0, start, ntile_per_thread, ndays, nthreads):
    """HRU-days/s of the C oracle (oracle/prms_oracle.c): `nthreads` threads,
    each simulating `ntile_per_thread` DRB tiles for `ndays` days (ctypes
    releases the GIL, so the threads run on separate cores)."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import prms_oracle as po
    import scenarios

    nbase = f0["prcp"].shape[0]
    idx = np.arange(ndays) % nbase
    doms = []
    for th in range(nthreads):
        fsub = {k: v[idx] for k, v in f0.items()}
        pt, ft = scenarios.tile(p0, fsub, ntile_per_thread)
        dtemp, pscale = tile_offsets(ntile_per_thread, th)
        nb = f0["prcp"].shape[1]
        for t in range(ntile_per_thread):
            sl = slice(t * nb, (t + 1) * nb)
            ft["prcp"][:, sl] = fsub["prcp"] * pscale[t]
            ft["tmax"][:, sl] = fsub["tmax"] + dtemp[t]
            ft["tmin"][:, sl] = fsub["tmin"] + dtemp[t]
        doms.append((po.Oracle(pt, start=start), ft))
    seg_vars = po.SEG_VARS

    def work(o, ft):
        o.run(ft["prcp"], ft["tmax"], ft["tmin"], (), seg_vars, budgets=True)

    ths = [threading.Thread(target=work, args=d) for d in doms]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    units = nthreads * ntile_per_thread * f0["prcp"].shape[1] * ndays
    for o, _ in doms:
        o.close()
    return units / dt, units, dt


def _ref_worker(args):
    """One process of the numpy / numba reference arm: the unmodified
    pywatershed package from oracle/_ref, drb_2yr through pws.Model."""
    calc_method, ndays, warm = args
    os.dup2(2, 1)  # the reference prints ("... jit compiling with numba"): stdout is for the JSON line
    os.environ["TQDM_DISABLE"] = "1"
    sys.path.insert(0, str(ROOT / "oracle"))
    import ref_pack

    return ref_pack.time_drb(calc_method, ndays, warm)


def reference_python_rate(calc_method, ndays, nproc, passes=1, warmup=0):
    """HRU-days/s of the UNMODIFIED reference package (oracle/_ref), drb_2yr NHM
    chain through pws.Model (asv_benchmarks/benchmarks/prms.py:146-171), one
    process per host core as examples/05_parameter_ensemble.ipynb does, netCDF
    output off; the first 5 days of every model (numba's JIT) are untimed.
    Spawned processes: the caller may hold a CUDA context."""
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor

    rates, res = [], None
    for it in range(warmup + passes):
        with ProcessPoolExecutor(max_workers=nproc, mp_context=mp.get_context("spawn")) as ex:
            res = list(ex.map(_ref_worker, [(calc_method, ndays, True)] * nproc))
        wall = max(r["seconds"] for r in res)  # processes run side by side: the slowest one
        if it >= warmup:
            rates.append((nproc * NB_HRU * res[0]["ndays"], wall))
    units = sum(u for u, _ in rates)
    dt = sum(d for _, d in rates)
    one = res[0]
    sample = (f"{nproc} processes x drb_2yr (765 HRUs, 456 segments) x {one['ndays']} days (after 5 untimed "
              f"days: numba's JIT), pywatershed {one.get('version', '?')} calc_method={calc_method}, "
              f"pws.Model.run, no output; one process: {one['seconds']:.1f} s; last-day "
              f"sum(seg_outflow)={one['sum_seg_outflow']:.6f}")
    return units / dt, dt / max(len(rates), 1), sample


def reference_available():
    sys.path.insert(0, str(ROOT / "oracle"))
    import ref_pack

    return ref_pack.available()


def run_reference(args, spec):
    """--impl reference: the reference's CPU implementation on all host cores.
    Default path: the real reference package from oracle/_ref with its fastest
    calc_method (numba) when it was packed, else the C port; the line also
    carries one bounded pass of the other CPU paths (`other_cpu_paths`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import scenarios

    ncores = os.cpu_count() or 1
    nthreads = max(1, min(ncores, 64))
    have_ref = reference_available()
    ref = args.ref or ("numba" if have_ref else "port")
    if ref in ("numpy", "numba") and not have_ref:
        print(json.dumps({"impl": "reference", "unavailable":
                          "oracle/_ref (the packed reference package) is missing: run build() where "
                          "/root/reference exists"}))
        return
    p0, f0, start = scenarios.load_drb()

    def port_rate(passes, warmup):
        ndays = min(spec["ndays"], 3653)
        rates = []
        for it in range(warmup + passes):
            r, units, dt = cpu_oracle_rate(p0, f0, start, 1, ndays, nthreads)
            if it >= warmup:
                rates.append((units, dt))
        units, dt = sum(u for u, _ in rates), sum(d for _, d in rates)
        return units / dt, dt / len(rates), (
            f"{nthreads} threads x 1 DRB tile (765 HRUs, 456 segments) x {ndays} days per step, C "
            "restatement oracle/prms_oracle.c, full chain incl. routing and budgets, segment outputs kept")

    if ref == "port":
        value, step_s, sample = port_rate(args.steps, args.warmup)
        kind = "port"
    else:
        value, step_s, sample = reference_python_rate(ref, 731 if ref == "numba" else 200, nthreads,
                                                      args.steps, args.warmup)
        kind = "reference"
    others = {}
    for other in ("port", "numba", "numpy"):
        if other == ref or (other != "port" and not have_ref):
            continue
        if other == "port":
            v, _, smp = port_rate(1, 0)
        else:
            v, _, smp = reference_python_rate(other, 731 if other == "numba" else 120, nthreads)
        others[other] = {"value": v, "unit": UNIT, "cores": nthreads, "sample": smp}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * step_s,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": spec["label"], "note": "bounded sample of the workload on host cores: "
                   "independent drb_2yr basins, one per core", "cpu_path": ref},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": kind, "sample": sample},
        "other_cpu_paths": others,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def chain_tiles(p0, pt, ntile):
    """BASELINE.json configs[4], chained variant: every tile's main outlet drains
    into the farthest headwater of the next tile (depth ~ ntile x 72 levels)."""
    tos0 = np.asarray(p0["tosegment"], dtype=np.int64)
    nseg0 = tos0.shape[0]
    # distance to / id of the outlet by walking tosegment (456 segments: cheap)
    dist = np.zeros(nseg0, dtype=np.int64)
    outlet = np.zeros(nseg0, dtype=np.int64)
    for s0 in range(nseg0):
        j, n = s0, 0
        while tos0[j] > 0:
            j = tos0[j] - 1
            n += 1
        dist[s0], outlet[s0] = n, j
    main_outlet = int(np.argmax(np.bincount(outlet)))
    head = int(np.argmax(np.where(outlet == main_outlet, dist, -1)))
    tos = np.asarray(pt["tosegment"]).copy()
    for t in range(ntile - 1):
        tos[t * nseg0 + main_outlet] = (t + 1) * nseg0 + head + 1
    out = dict(pt)
    out["tosegment"] = tos
    return out, int(dist.max()) + 1


def run_channel_only(args):
    """--workload channel | channel_chained: BASELINE.json configs[4], PRMSChannel
    alone (prms_channel.py:388-585) over 123 DRB tiles of segments (56,088), lateral
    inflow LogNormal (seed 11) resident in HBM; metric = segment-days/s."""
    import torch
    import scenarios
    import __graft_entry__ as g

    g.build()
    from pywatershed_b200 import Engine

    dev = torch.device("cuda", 0)
    p0, f0, start = scenarios.load_drb()
    ntile, nd = 123, 3653
    pt, _ = scenarios.tile(p0, {k: v[:2] for k, v in f0.items()}, ntile)
    depth = 72
    if args.workload == "channel_chained":
        pt, d1 = chain_tiles(p0, pt, ntile)
        depth = d1 * ntile
    eng = Engine(pt, start, device=0)
    eng.select_outputs((), eng.reg.seg_vars)   # all five public variables are written each day
    nseg = eng.nseg
    T = max(1, min(args.block_days, nd))
    gen = torch.Generator(device=dev).manual_seed(11)
    lat = torch.exp(torch.randn((nd, nseg), dtype=torch.float64, device=dev, generator=gen))
    out = torch.empty((T, 5, nseg), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    def one_pass():
        eng.reset()
        torch.cuda.synchronize()
        eng.timer_mark(0)
        d = 0
        while d < nd:
            n = min(T, nd - d)
            eng.run_channel_device(n, lat.data_ptr() + d * nseg * 8, out.data_ptr(), sync=False)
            d += n
        eng.timer_mark(1)
        eng.sync()
        torch.cuda.synchronize()
        return eng.timer_elapsed_ms()

    for _ in range(args.warmup):
        one_pass()
    sampler = ClockSampler(0)
    sampler.start()
    ms = [one_pass() for _ in range(args.steps)]
    clocks = sampler.stop()
    t = sum(ms) / 1e3
    units = nseg * nd * args.steps
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    peak = float(peaks.get("hbm_gbs", 6548.8))
    ach = units * 48 / t / 1e9
    print(json.dumps({
        "metric": "routed segment-days/sec, PRMSChannel alone", "value": units / t, "unit": "segment-days/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"channel-only, {nseg:,} segments ({ntile} DRB tiles"
                               + (", chained outlet->headwater" if args.workload == "channel_chained" else "")
                               + f"), depth ~{depth} levels, {nd} days x 24 hourly sub-steps",
                   "block_days": T, "l2": "lateral inflow + outputs per block exceed L2"},
        "segment_substeps_per_s": units * 24 / t,
        "checksum_seg_outflow_last_day": float(out[(nd - 1) % T, eng.sel_seg.index("seg_outflow")].sum()),
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": None, "kernel": "k_route_chunks",
                     "note": "48 B per segment-day (8 read + 5x8 written); the kernel is bound by the "
                             "serial hourly recurrence along the network depth, not by HBM"},
        "gpu_launches": eng.launch_count(), "clocks": clocks}))


# --------------------------------------------------------------------------- the chain
class Dist:
    """torch.distributed plumbing of one rank (NCCL; a single process when N = 1)."""

    def __init__(self, shard_of=None):
        import torch

        self.shard_of = shard_of
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: pywatershed_b200 has no CPU path")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            import torch.distributed as dist

            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            # keep stdout to the one JSON line: NCCL prints its version banner to stdout at the
            # VERSION and WARN levels; anything more verbose the user asked for goes to stderr
            if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
                del os.environ["NCCL_DEBUG"]
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        import torch

        torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def max(self, x: float) -> float:
        import torch

        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            import torch.distributed as dist

            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x) -> np.ndarray:
        import torch

        t = torch.as_tensor(np.asarray(x, dtype=np.float64), device=self.dev).clone()
        if self.world > 1:
            import torch.distributed as dist

            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def close(self):
        if self.world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


def make_shard(spec, D: Dist, scaling: str):
    """This rank's part of the workload: parameters, the HRUs / segments of the
    whole domain it owns, and how many HRU-days the whole job is."""
    from pywatershed_b200 import sharding
    import scenarios

    world = D.shard_of or D.world  # --shard-of: rank 0's share of an N-way split, on one GPU
    if spec.get("ensemble"):
        p0, f0, start = scenarios.load_drb()
        nm = spec["ntile"]
        if scaling == "strong":
            members = sharding.partition_members(nm, world, D.rank)
            total = nm
        else:
            members = np.arange(nm) + D.rank * nm
            total = nm * D.world
        pt = ensemble_parameters(p0, members)
        return dict(p0=p0, f0=f0, start=start, params=pt, hru_idx=None, members=members,
                    forcing_period=NB_HRU, job_hrus=total * NB_HRU, job_segs=total * NB_SEG,
                    parallelism=f"member-sharded x{world} ({members.size} members on this rank), "
                                "no data-path collective")
    p0, f0, pt, start = build_domain(spec, D.rank)
    nhru, nseg = pt["hru_area"].shape[0], pt["tosegment"].shape[0]
    if scaling == "weak" or world == 1:
        return dict(p0=p0, f0=f0, start=start, params=pt, hru_idx=None, seg_idx=None,
                    forcing_rank=D.rank if scaling == "weak" else 0,
                    job_hrus=nhru * (D.world if scaling == "weak" else 1),
                    job_segs=nseg * (D.world if scaling == "weak" else 1),
                    parallelism=(f"replicas x{D.world}: one whole domain per GPU (weak scaling)"
                                 if D.world > 1 else "one GPU"))
    seg_rank, hru_rank = sharding.partition_basins(pt["tosegment"], pt["hru_segment"], world)
    q, hru_idx, seg_idx = sharding.shard_parameters(pt, seg_rank, hru_rank, D.rank)
    return dict(p0=p0, f0=f0, start=start, params=q, hru_idx=hru_idx, seg_idx=seg_idx, forcing_rank=0,
                job_hrus=nhru, job_segs=nseg,
                parallelism=f"basin-sharded x{world}: outlet trees of ONE domain dealt to the GPUs "
                            f"(this rank: {hru_idx.size:,} HRUs, {seg_idx.size:,} segments), "
                            "no data-path collective")


def shard_forcings(spec, shard, dev):
    import torch

    if spec.get("ensemble"):
        nd = spec["ndays"]
        f0 = shard["f0"]
        idx = np.arange(nd) % f0["prcp"].shape[0]
        return {k: torch.from_numpy(np.ascontiguousarray(f0[k][idx])).to(dev).float().double().contiguous()
                for k in ("prcp", "tmax", "tmin")}
    return device_forcings(shard["f0"], spec, shard["forcing_rank"], dev, hru_idx=shard["hru_idx"])


def time_device_passes(eng, forc, nd, T, bufs, viol, D, warmup, steps):
    """W + K passes with the forcings resident in HBM; returns per-step device
    ms (this rank) and the accumulated kernel timing."""
    ncol = eng.ncol

    def one_pass():
        eng.reset()
        D.barrier()
        eng.timer_mark(0)
        d = 0
        while d < nd:
            n = min(T, nd - d)
            o = d * ncol * 8
            eng.run_device(n, forc["prcp"].data_ptr() + o, forc["tmax"].data_ptr() + o,
                           forc["tmin"].data_ptr() + o, bufs[0].data_ptr(), bufs[1].data_ptr(),
                           bufs[2].data_ptr(), viol=viol[d:d + n], sync=False)
            d += n
        eng.timer_mark(1)
        eng.sync()
        D.barrier()
        return eng.timer_elapsed_ms()

    for _ in range(warmup):
        one_pass()
    acc = dict(step_ms=[], col_ms=0.0, ch_ms=0.0, ncol=0, nblk=0, lanes=0)
    for _ in range(steps):
        acc["step_ms"].append(one_pass())
        c, h = eng.timing()
        acc["col_ms"] += c
        acc["ch_ms"] += h
        acc["ncol"] += eng.last_launches()[0]
        nb, ln = eng.last_blocks()
        acc["nblk"] += nb
        acc["lanes"] = ln
    return acc


def alloc_outputs(eng, T, hru_vars, seg_vars, tiled, dev):
    import torch

    nf, ni = len(eng.sel_f64), len(eng.sel_i32)
    onhru = eng.output_tile()[1] if tiled else eng.nhru
    of = torch.empty((T, max(nf, 1), onhru), dtype=torch.float64, device=dev)
    oi = torch.empty((T, max(ni, 1), onhru), dtype=torch.int32, device=dev)
    osg = torch.empty((T, max(len(seg_vars), 1), max(eng.nseg, 1)), dtype=torch.float64, device=dev)
    return of, oi, osg


def pcie_ceiling(D: Dist, nbytes=1 << 30, reps=3):
    """Plain cudaMemcpyAsync between pinned host memory and the GPU, all ranks
    at the same time: what the box gives, aggregate GB/s over the ranks."""
    import torch

    h_up = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_dn = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_up = torch.empty(nbytes, dtype=torch.uint8, device=D.dev)
    d_dn = torch.empty(nbytes, dtype=torch.uint8, device=D.dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(up, dn):
        D.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s1):
                    d_up.copy_(h_up, non_blocking=True)
            if dn:
                with torch.cuda.stream(s2):
                    h_dn.copy_(d_dn, non_blocking=True)
        D.barrier()
        return D.max(time.perf_counter() - t0)

    run(True, True)
    res = {}
    for name, up, dn in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
        dt = run(up, dn)
        res[name + "_gbs"] = (reps * nbytes * D.world * (2 if up and dn else 1)) / dt / 1e9
    res["note"] = (f"{D.world} rank(s) x {reps} x {nbytes >> 20} MiB per direction, pinned host memory, "
                   "cudaMemcpyAsync, all ranks at once; 'both' = up and down at the same time (sum)")
    return res


def run_split_columns(args):
    """--workload conus_split | conus_split_chained: BASELINE.json configs[2] with its
    segment graph kept on ONE rank (as a real basin with one dominant outlet tree
    would need; `_chained`: the 144 tiles really are one tree, outlet -> headwater):
    HRU columns split over the ranks (sharding.partition_hrus), ONE exchange step per
    block of days -- an NCCL reduce of seg_lateral_inflow [T, nsegment] to the router --
    and PRMSChannel alone on the router (sharding.SplitColumnsOneRouter).  Strong
    scaling; `value` = HRU-days of the whole domain / max-over-ranks time."""
    import torch
    import __graft_entry__ as g

    D = Dist()
    if D.rank == 0:
        g.build()
    D.barrier()
    from pywatershed_b200 import sharding

    spec = workload_spec("conus")
    p0, f0, pt, start = build_domain(spec)
    chained = args.workload.endswith("_chained")
    if chained:
        pt, _ = chain_tiles(p0, pt, spec["ntile"])
    nhru, nseg, nd = pt["hru_area"].shape[0], pt["tosegment"].shape[0], spec["ndays"]
    T = max(1, min(args.block_days, nd))
    # the router also routes (about a fifth of the one-GPU step): its share of the HRUs
    share = args.router_share if args.router_share is not None else max(0.0, 1.0 - 0.2 * D.world)
    run = sharding.SplitColumnsOneRouter(pt, start, D.rank, D.world, D.local_rank, block_days=T,
                                         router_share=share, lanes=args.lanes,
                                         channel_chunk=args.channel_chunk)
    forc = device_forcings(f0, spec, 0, D.dev, hru_idx=run.hru_idx)
    seg = torch.zeros((nd, len(run.seg_vars), nseg), dtype=torch.float64, device=D.dev) \
        if run.route is not None else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def one_pass():
        run.reset()
        D.barrier()
        e0.record()
        run.run_device(forc["prcp"], forc["tmax"], forc["tmin"], seg_out=seg)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    for _ in range(args.warmup):
        one_pass()
    sampler = ClockSampler(D.local_rank)
    if D.rank == 0:
        sampler.start()
    ms = D.max(sum(one_pass() for _ in range(args.steps)))
    clocks = sampler.stop() if D.rank == 0 else None
    chk = float(seg[-1, run.seg_vars.index("seg_outflow")].sum()) if seg is not None else 0.0
    launches = run.cols.launch_count() + (run.route.launch_count() if run.route is not None else 0)
    nloc = run.cols.nhru
    if D.rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": nhru * nd * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": D.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": spec["label"] + (", tiles chained into ONE outlet tree" if chained else "")
                       + "; segment graph on rank 0, HRU columns split over the ranks",
                       "parallelism": f"hru-split x{D.world} + 1 router, one NCCL reduce of "
                                      f"seg_lateral_inflow [{T}, {nseg}] f64 per block",
                       "block_days": T, "router_share": share, "nhru_rank0": nloc, "job_hrus": nhru, "job_segments": nseg,
                       "outputs": "5 segment variables every day on the router (HRU variables: the three "
                                  "volume fluxes of the exchange, on each rank's device)",
                       "l2": "forcings and results exceed L2"},
            "exchange_bytes_per_block": T * nseg * 8,
            "checksum_seg_outflow_last_day": chk, "gpu_launches": launches, "clocks": clocks}))
    run.close()
    D.close()



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="conus")
    ap.add_argument("--scaling", choices=("strong", "weak"), default="strong",
                    help="N > 1: strong = ONE domain / ensemble split over the GPUs (BASELINE.json "
                         "configs[2], configs[3]); weak = every GPU its own copy of the workload")
    ap.add_argument("--outputs", default="full")
    ap.add_argument("--block-days", type=int, default=256)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--ref", choices=("port", "numpy", "numba"), default=None,
                    help="--impl reference: which CPU path is the line's value (default: numba from "
                         "oracle/_ref when packed, else the C port)")
    ap.add_argument("--no-e2e", action="store_true")
    # measured on the CONUS workload (float32 forcings, HRU-days/s): 64 -> 3.50e9, 128 -> 3.53e9,
    # 256 -> 3.88e9, 384 -> 3.76e9, 512 -> 3.57e9 (profiles/r01_e2e_block_days.jsonl)
    ap.add_argument("--e2e-block-days", type=int, default=256)
    ap.add_argument("--e2e-forcing", choices=("f32", "f64"), default="f32",
                    help="type of the host forcing arrays of the e2e leg: float32 as in the reference's "
                         "forcing files (widened on the device, exact) or float64")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the ensemble / e2e_nhm_default / pcie / e2e_model objects")
    ap.add_argument("--lanes", type=int, default=None, help="column launches per block (library option)")
    ap.add_argument("--overlap", type=int, default=None, choices=(0, 1),
                    help="routing beside the next block's column (library option; default: by domain size)")
    ap.add_argument("--channel-chunk", type=int, default=None,
                    help="segments per routing CTA (library option channel_chunk; default 512)")
    ap.add_argument("--column-hrus", type=int, default=None, choices=(128, 256),
                    help="HRUs per column CTA (library option column_hrus; default: by domain size)")
    ap.add_argument("--lib-opt", action="append", default=[], metavar="NAME=VALUE",
                    help="pws_set_option(NAME, VALUE) on every engine (tuning aid), e.g. wide_column=0")
    ap.add_argument("--router-share", type=float, default=None,
                    help="--workload conus_split: the router rank's share of the HRUs relative to "
                         "another rank's (default max(0, 1 - 0.2 N))")
    ap.add_argument("--shard-of", type=int, default=None,
                    help="tuning aid on ONE GPU: run rank 0's share of an N-way split; `value` is then "
                         "what N GPUs would give if every rank took as long as this one")
    ap.add_argument("--layout", choices=("tiled", "rows"), default="tiled",
                    help="device-resident results as [day][tile][variable][H] (pws_output_tile; needs "
                         "--outputs full) or as [day][variable][nhru] rows")
    args = ap.parse_args()
    if args.workload in ("channel", "channel_chained"):
        return run_channel_only(args)
    if args.workload in ("conus_split", "conus_split_chained"):
        return run_split_columns(args)
    spec = workload_spec(args.workload)
    if args.impl == "reference":
        return run_reference(args, spec)

    import torch

    D = Dist(args.shard_of)
    import __graft_entry__ as g

    if D.rank == 0:
        g.build()
    D.barrier()
    from pywatershed_b200 import Engine

    def engine_for(spec, shard):
        eng = Engine(shard["params"], shard["start"], device=D.local_rank,
                     forcing_period=shard.get("forcing_period"), lanes=args.lanes,
                     channel_chunk=args.channel_chunk)
        if args.column_hrus is not None:
            eng._opt("column_hrus", args.column_hrus)
        for kv in args.lib_opt:
            k, v = kv.split("=", 1)
            eng._opt(k, float(v))
        if args.overlap is not None:
            eng._opt("overlap_routing", args.overlap)
        return eng

    def device_leg(spec, outputs, layout, block_days, warmup, steps, clocks=False):
        """value-style measurement of one workload; returns a dict of results."""
        shard = make_shard(spec, D, args.scaling)
        eng = engine_for(spec, shard)
        hru_vars, seg_vars = output_sets(eng.reg, outputs)
        eng.select_outputs(hru_vars, seg_vars)
        nd = spec["ndays"]
        forc = shard_forcings(spec, shard, D.dev)
        tiled = layout == "tiled" and outputs == "full"
        if tiled:
            eng.set_tiled_outputs(True)
        # bound the result buffer of one block to ~35 GB
        per_day = (8 * len(eng.sel_f64) + 4 * len(eng.sel_i32)) * max(eng.nhru, 1) + 1
        T = int(max(1, min(block_days, nd, 35e9 // per_day)))
        bufs = alloc_outputs(eng, T, hru_vars, seg_vars, tiled, D.dev)
        viol_t = torch.zeros((nd, 6), dtype=torch.int32, pin_memory=True)
        viol = viol_t.numpy()
        torch.cuda.synchronize()
        sampler = ClockSampler(D.local_rank)
        l0 = eng.launch_count()
        if clocks and D.rank == 0:
            sampler.start()
        acc = time_device_passes(eng, forc, nd, T, bufs, viol, D, warmup, steps)
        clk = sampler.stop() if clocks and D.rank == 0 else None
        launches = eng.launch_count() - l0
        total_ms = D.max(sum(acc["step_ms"]))
        # the only communication of a sharded run: budget verdicts and a checksum of
        # the last day's streamflow, summed over the ranks (NCCL all-reduce)
        viol_total = int(D.sum(viol.sum(axis=0)).sum())
        seg_last = float(D.sum([float(bufs[2][(nd - 1) % T, eng.sel_seg.index("seg_outflow")].sum())
                                if "seg_outflow" in eng.sel_seg and eng.nseg else 0.0])[0])
        units_per_step = shard["job_hrus"] * nd
        res = dict(shard=shard, eng=eng, forc=forc, hru_vars=hru_vars, seg_vars=seg_vars, T=T, nd=nd,
                   tiled=tiled, acc=acc, total_ms=total_ms, launches=launches, clocks=clk,
                   viol_total=viol_total, seg_last=seg_last, units_per_step=units_per_step,
                   value=units_per_step * steps / (total_ms * 1e-3), bufs=bufs)
        return res

    main_leg = device_leg(spec, args.outputs, args.layout, args.block_days, args.warmup, args.steps, clocks=True)
    eng, shard, forc = main_leg["eng"], main_leg["shard"], main_leg["forc"]
    hru_vars, seg_vars, nd, T = main_leg["hru_vars"], main_leg["seg_vars"], main_leg["nd"], main_leg["T"]
    nhru, nseg = eng.nhru, eng.nseg
    units_per_step = main_leg["units_per_step"]
    del main_leg["bufs"]
    torch.cuda.empty_cache()

    # ---- e2e: host-facing call, pinned host buffers, H2D + D2H inside the region
    from pywatershed_b200.engine import calendar

    def host_leg(hsel, ssel, ndays, block_days, passes, f32=True):
        """pws_run_host[_f32] over the first `ndays` days with this output selection."""
        eng.set_tiled_outputs(False)
        eng.select_outputs(hsel, ssel)
        hf = {k: (v[:ndays].float() if f32 else v[:ndays]).cpu().pin_memory() for k, v in forc.items()}
        run_host = eng.L.pws_run_host_f32 if f32 else eng.L.pws_run_host
        eng.set_block_days(block_days)
        nf, ni = len(eng.sel_f64), len(eng.sel_i32)
        res_f = torch.empty((ndays, max(nf, 1), nhru if nf else 1), dtype=torch.float64, pin_memory=True)
        res_i = torch.empty((ndays, max(ni, 1), nhru if ni else 1), dtype=torch.int32, pin_memory=True)
        res_seg = torch.empty((ndays, max(len(ssel), 1), max(nseg, 1)), dtype=torch.float64, pin_memory=True)
        res_viol = np.zeros((ndays, 6), dtype=np.int32)
        cal = calendar(shard["start"], ndays)

        def one():
            eng.reset()
            D.barrier()
            t0 = time.perf_counter()
            eng._check(run_host(
                eng.h, ndays, cal.ctypes.data, hf["prcp"].data_ptr(), hf["tmax"].data_ptr(),
                hf["tmin"].data_ptr(), res_f.data_ptr(), res_i.data_ptr(),
                res_seg.data_ptr(), res_viol.ctypes.data))
            D.barrier()
            return time.perf_counter() - t0

        one()
        one()
        ts = [one() for _ in range(passes)]
        tot = D.max(sum(ts))
        h2d = int(D.sum([3 * ndays * eng.ncol * (4 if f32 else 8)])[0])
        d2h = int(D.sum([ndays * (nf * nhru * 8 + ni * nhru * 4 + len(ssel) * nseg * 8 + 24)])[0])
        chk = float(D.sum([float(res_seg[-1, ssel.index("seg_outflow")].sum()) if "seg_outflow" in ssel else 0.0])[0])
        return dict(value=shard["job_hrus"] * ndays * len(ts) / tot, seconds_per_pass=tot / len(ts),
                    h2d=h2d, d2h=d2h, checksum=chk, viol=int(D.sum([int(res_viol.sum())])[0]))

    e2e = None
    pcie = None
    if not args.no_e2e:
        if not args.no_extras:
            pcie = pcie_ceiling(D)
        f32 = args.e2e_forcing == "f32"
        r = host_leg([], ["seg_outflow"], nd, args.e2e_block_days, max(1, min(args.steps, 3)), f32)
        e2e = {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": r["h2d"],
               "d2h_bytes_per_step": r["d2h"],
               "forcing_dtype": "float32 (as stored in the reference's forcing files; widened on the "
                                "device)" if f32 else "float64",
               "outputs": "seg_outflow [ndays, nsegment] (streamflow at every segment, the product of an "
                          "NHM run) + 6 budget verdicts per day; the `value` leg writes all 148 variables "
                          "to HBM but drains none -- see e2e_nhm_default for HRU variables on the host",
               "block_days": args.e2e_block_days, "checksum_seg_outflow_last_day": r["checksum"],
               "budget_violations": r["viol"]}
        if pcie:
            e2e["pcie_gbs"] = (r["h2d"] + r["d2h"]) / r["seconds_per_pass"] / 1e9
            e2e["pcie_frac_of_ceiling"] = e2e["pcie_gbs"] / pcie["both_gbs"]

    # ---- extras: the same JSON line carries the other measurements north_star names
    extras = {}
    if not args.no_extras and args.workload == "conus" and not args.no_e2e:
        # the reference control file's output selection (74 HRU variables + the segment
        # variables) drained to pinned host memory: a bounded sample of 366 days
        names = json.loads((ROOT / "tests" / "golden" / "nhm_default_vars.json").read_text())
        hsel = [v for v in eng.reg.hru_vars if v in names]
        nd_s = 366
        r = host_leg(hsel, list(eng.reg.seg_vars), nd_s, 64, 2, True)
        x = {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": r["h2d"],
             "d2h_bytes_per_step": r["d2h"], "sample": f"first {nd_s} days of the workload, 64-day blocks",
             "outputs": f"{len(hsel)} HRU variables of pywatershed/data/drb_2yr/nhm.control + 5 segment "
                        "variables, every day, to pinned host memory",
             "bound": "PCIe (device-to-host drain)"}
        if pcie:
            x["pcie_gbs"] = (r["h2d"] + r["d2h"]) / r["seconds_per_pass"] / 1e9
            x["pcie_frac_of_d2h_ceiling"] = (r["d2h"] / r["seconds_per_pass"] / 1e9) / pcie["d2h_gbs"]
        extras["e2e_nhm_default"] = x
    eng.close()
    del forc, eng
    main_leg["eng"] = main_leg["forc"] = None
    torch.cuda.empty_cache()

    if not args.no_extras and args.workload == "conus":
        # BASELINE.json configs[3]: the 1024-member DRB parameter ensemble, by member
        espec = workload_spec("ensemble")
        el = device_leg(espec, "full", args.layout, 64, 2, max(1, min(args.steps, 3)))
        ea = el["acc"]
        ebpu = algorithmic_bytes_per_hru_day(el["eng"].reg, el["hru_vars"])
        e_units_rank = el["eng"].nhru * el["nd"] * len(ea["step_ms"])
        extras["ensemble"] = {
            "metric": METRIC, "value": el["value"], "unit": UNIT, "scaling": args.scaling,
            "ms_per_step": el["total_ms"] / len(ea["step_ms"]),
            "config": {"workload": espec["label"], "parallelism": el["shard"]["parallelism"],
                       "outputs": "full", "block_days": el["T"], "lanes": ea["lanes"],
                       "forcings": "[ndays, 765] shared by the members (option forcing_period): "
                                   "resident once, read through L2",
                       "budget_violations": el["viol_total"]},
            "column_hbm_frac": (e_units_rank * ebpu / max(ea["col_ms"] * 1e-3, 1e-12) / 1e9) /
                               float(_peak()[0]),
            "checksum_seg_outflow_last_day": el["seg_last"], "gpu_launches": int(el["launches"])}
        el["eng"].close()
        del el
        torch.cuda.empty_cache()

    if not args.no_extras and args.workload == "conus" and D.world == 1:
        try:
            extras["e2e_model"] = model_leg(spec)
        except Exception as exc:  # the Model leg must not take the bench line down
            extras["e2e_model"] = {"error": f"{type(exc).__name__}: {exc}"}

    if D.rank != 0:
        D.close()
        return

    # ---- roofline of the dominant kernel (k_hru_column), measured live.  The
    # lanes of a block run concurrently (and beside the previous block's routing),
    # so the kernel's time is the UNION of the launches' CUDA-event intervals on
    # their streams (pws_last_timing), and `achieved` = the bytes of all launches of
    # the timed region over that time.
    acc = main_leg["acc"]
    peak, peak_src = _peak()
    reg = __import__("pywatershed_b200").Registry.get()
    bpu = algorithmic_bytes_per_hru_day(reg, hru_vars)
    step_ms = acc["step_ms"]
    col_s = acc["col_ms"] * 1e-3
    units_rank = nhru * nd * args.steps  # HRU-days this rank's kernels processed
    achieved = units_rank * bpu / max(col_s, 1e-12) / 1e9
    units_per_launch = units_rank / max(acc["ncol"], 1)
    traffic, fp64 = None, None
    prof = ROOT / "profiles" / "r01_column_traffic.json"
    if prof.exists():
        try:
            pj = json.loads(prof.read_text())
            if args.outputs == "full":
                traffic = pj["dram_bytes_per_launch"] * units_per_launch / pj["hru_days_per_launch"]
            tf = pj["fp64_flops_per_hru_day"] * units_rank / max(col_s, 1e-12) / 1e12
            fp64 = {"achieved": tf, "peak": pj["fp64_peak_tflops"], "unit": "TFLOP/s",
                    "frac": tf / pj["fp64_peak_tflops"],
                    "flops_per_hru_day": pj["fp64_flops_per_hru_day"]}
        except Exception:
            traffic, fp64 = None, None
    whole = units_rank * bpu / (sum(step_ms) * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "kernel": "k_hru_column",
                "peak_source": peak_src, "bytes_per_hru_day": bpu,
                "hru_days_per_launch": units_per_launch,
                "avg_launch_ms": acc["col_ms"] / max(acc["ncol"], 1),
                "launches_per_block": acc["lanes"],
                "how": "algorithmic bytes of every k_hru_column launch of the timed region / union of "
                       "the launches' CUDA-event intervals (lanes run concurrently); rank 0's kernels",
                "share_of_step": acc["col_ms"] / max(sum(step_ms), 1e-9),
                "channel_share_of_step": acc["ch_ms"] / max(sum(step_ms), 1e-9),
                "whole_step_hbm_frac": whole / peak,
                "fp64": fp64}

    cpu = None
    if not args.no_cpu and D.world == 1:
        ncores = os.cpu_count() or 1
        nthreads = max(1, min(ncores, 64))
        r, units, dt = cpu_oracle_rate(shard["p0"], shard["f0"], shard["start"], 1, min(nd, 3653), nthreads)
        cpu = {"value": r, "unit": UNIT, "cores": nthreads, "kind": "port",
               "sample": f"{nthreads} threads x 1 DRB tile x {min(nd, 3653)} days "
                         f"({units:.3g} HRU-days, {dt:.1f} s), C oracle, full chain + routing + budgets"}
        if not args.no_extras:
            try:
                extras["parity"] = parity_check(D.local_rank)
            except Exception as exc:  # the checker must not take the bench line down
                extras["parity"] = {"error": f"{type(exc).__name__}: {exc}"}
        if reference_available() and not args.no_extras:
            # north_star: the reference's numpy and numba paths on this box's cores, same run
            for cm, ndr in (("numba", 731), ("numpy", 120)):
                try:
                    v, _, smp = reference_python_rate(cm, ndr, nthreads)
                    cpu["reference_" + cm] = {"value": v, "unit": UNIT, "cores": nthreads,
                                              "kind": "reference", "sample": smp}
                except Exception as exc:
                    cpu["reference_" + cm] = {"error": f"{type(exc).__name__}: {exc}"}

    kinfo = None
    line = {
        "metric": METRIC, "value": main_leg["value"], "unit": UNIT, "n_gpus": D.world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": main_leg["total_ms"] / args.steps, "higher_is_better": True,
        "scaling": args.scaling if D.world > 1 else "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": spec["label"], "outputs": args.outputs,
                   "layout": "tiled" if main_leg["tiled"] else "rows",
                   "n_output_vars": len(hru_vars) + len(seg_vars),
                   "block_days": T, "nhru_rank0": nhru, "nsegment_rank0": nseg, "ndays": nd,
                   "job_hrus": shard["job_hrus"], "job_segments": shard["job_segs"],
                   "l2": "inputs larger than L2 (forcings 3 x %.1f GB on rank 0, streamed once per step)"
                         % (nd * nhru * 8 / 1e9),
                   "parallelism": shard["parallelism"] + (
                       f" [EMULATED on one GPU with --shard-of {args.shard_of}: rank 0's share only]"
                       if args.shard_of else ""),
                   "budget_violations": main_leg["viol_total"],
                   "checksum_seg_outflow_last_day": main_leg["seg_last"]},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(main_leg["launches"]),
        "clocks": main_leg["clocks"],
    }
    if pcie:
        line["pcie"] = pcie
    line.update(extras)
    print(json.dumps(line))
    D.close()


def _peak():
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        return float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def model_leg(spec, ndays=731):
    """e2e through the API a pywatershed user calls: pywatershed_b200.Model over
    the whole CONUS-scale domain for `ndays` days, forcings as float32 numpy
    arrays in host memory (what the reference's forcing files hold), (i) no
    output files, (ii) the reference control file's output selection written to
    netCDF.  Timed: Model.run() (engine creation, parameter upload and init
    included in the first figure, excluded in the second)."""
    import tempfile

    import torch

    import pywatershed_b200 as pwsb

    p0, f0, pt, start = build_domain(spec)
    nhru = pt["hru_area"].shape[0]
    nseg = pt["tosegment"].shape[0]
    dev = torch.device("cuda", 0)
    sub = dict(spec)
    sub["ndays"] = ndays
    forc = {k: v.float().cpu().numpy() for k, v in device_forcings(f0, sub, 0, dev).items()}
    torch.cuda.empty_cache()
    procs = ["PRMSSolarGeometry", "PRMSAtmosphere", "PRMSCanopy", "PRMSSnow", "PRMSRunoff",
             "PRMSSoilzone", "PRMSGroundwater", "PRMSChannel"]

    def make(**opts):
        t0 = np.datetime64(start, "s")
        o = {"budget_type": "error", "calc_method": "cuda"}
        o.update(opts)
        control = pwsb.Control(t0, t0 + np.timedelta64(ndays - 1, "D"), np.timedelta64(24, "h"), options=o)
        coords = {"nhm_id": pt["nhm_id"], "nhm_seg": pt["nhm_seg"]}
        data = {k: v for k, v in pt.items() if k not in coords}
        params = pwsb.Parameters(dims={"nhru": nhru, "nsegment": nseg, "nmonth": 12, "ndoy": 366},
                                 coords=coords, data_vars=data)
        m = pwsb.Model([getattr(pwsb, n) for n in procs], control=control, parameters=params,
                       find_input_files=False)
        atm = m.processes["PRMSAtmosphere"]
        for k in ("prcp", "tmax", "tmin"):
            atm.set_input_to_adapter(k, forc[k])
        return m

    out = {"unit": UNIT, "ndays": ndays, "nhru": nhru,
           "api": "pywatershed_b200.Model([...8 processes...]).run(), forcings float32 numpy [ndays, nhru]"}
    units = nhru * ndays
    # (i) no output: first run (includes engine creation + init), then warm runs
    m = make()
    t0 = time.perf_counter()
    m.run(finalize=True)
    out["first_run_s"] = time.perf_counter() - t0
    chk = float(m.processes["PRMSChannel"]["seg_outflow"].sum())
    ts = []
    for _ in range(4):
        m = make()
        m._ensure_engine()  # construction cost reported separately
        t0 = time.perf_counter()
        m.run(finalize=True)
        ts.append(time.perf_counter() - t0)
    out["value"] = units / min(ts)
    out["run_s"] = min(ts)
    out["checksum_seg_outflow_last_day"] = chk
    out["h2d_bytes_per_step"] = int(3 * ndays * nhru * 4)
    # (ii) the reference control file's variables to netCDF: a bounded sample of days
    # (2 years of 74 variables at this size are 55 GB of files)
    names = json.loads((ROOT / "tests" / "golden" / "nhm_default_vars.json").read_text())
    nd_full, ndays = ndays, min(ndays, 92)
    units = nhru * ndays
    forc = {k: v[:ndays] for k, v in forc.items()}
    with tempfile.TemporaryDirectory(dir=os.environ.get("PWS_BENCH_TMP", None)) as td:
        m = make(netcdf_output_var_names=list(names), netcdf_output_dir=td)
        m._ensure_engine()
        m.initialize_netcdf(td)
        t0 = time.perf_counter()
        m.run(finalize=True)
        dt = time.perf_counter() - t0
        nbytes = sum(f.stat().st_size for f in pl.Path(td).glob("*.nc"))
    out["netcdf"] = {"value": units / dt, "unit": UNIT, "run_s": dt, "ndays": ndays, "bytes_written": int(nbytes),
                     "writer_gbs": nbytes / dt / 1e9, "n_vars": len(names),
                     "bound": "host: device-to-host drain + file writes"}
    return out


if __name__ == "__main__":
    main()
