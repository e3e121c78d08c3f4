"""bench.py -- simulated HRU-days/s of the full NHM chain on B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload conus|ucb|drb|ensemble|ensemble128|channel|channel_chained]
                    [--outputs full|nhm_default|routing] [--impl reference]

A *step* is one pass of the hot path over the whole workload: reset to the
initial conditions, then every day of the run (time-blocked).  Default
workload `conus`: the synthetic CONUS-scale domain of BASELINE.json configs[2]
-- 144 DRB tiles in different climates = 110,160 HRUs / 65,664 segments,
3,653 days (1979-01-01 .. 1988-12-31), DRB forcings cycled with a per-tile
temperature offset / precipitation scale (tests/scenarios.py:tile).  It is the
configuration north_star's target is quoted on and fits one GPU; configs[1]
(3,851 HRUs) fills 30 of 148 SMs and is available as `--workload ucb`.

  value   HRU-days/s, forcings already resident in HBM, every public variable
          (143 HRU + 5 segment) written to HBM each day ("full" outputs)
  e2e     the same run through the host-facing entry point (Engine.run_host ->
          pws_run_host): forcings in pinned HOST memory, H2D every step,
          segment outputs + budget verdicts drained D2H every step
  N > 1   weak scaling: every rank simulates its own 144-tile set (independent
          basins, no data-path collective); NCCL only reduces times/budgets

`--impl reference` times the CPU oracle port (oracle/prms_oracle.c, the
reference's algorithm restated in C; the Python reference cannot travel to the
GPU box) on the host cores, one tile set per thread.
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


def workload_spec(name):
    if name == "conus":
        return dict(ntile=144, nhru_keep=None, ndays=3653, label="synthetic CONUS-scale NHM domain "
                    "(144 DRB tiles: 110,160 HRUs, 65,664 segments, 3,653 days)")
    if name == "ucb":
        return dict(ntile=6, nhru_keep=3851, ndays=731, label="ucb_2yr-shaped synthetic "
                    "(3,851 HRUs, 2,736 segments, 731 days; real ucb_2yr inputs are HDF5 blobs)")
    if name in ("ensemble", "ensemble128"):
        nm = 1024 if name == "ensemble" else 128
        return dict(ntile=nm, nhru_keep=None, ndays=731, label=f"{nm}-member DRB parameter ensemble "
                    f"({nm * 765:,} HRUs, {nm * 456:,} segments member-major, 731 days)")
    if name == "drb":
        return dict(ntile=1, nhru_keep=None, ndays=731, label="drb_2yr (765 HRUs, 456 segments, 731 days)")
    raise SystemExit(f"unknown workload {name}")


def build_domain(spec, rank=0):
    """Parameters (host, small) + base forcings; tile offsets differ per rank."""
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


def device_forcings(f0, spec, rank, dev):
    """[ndays, nhru] float64 tensors on `dev`: DRB CBH cycled in time, tiled in
    space with the per-tile climate offsets of tests/scenarios.py:tile, rounded
    to float32-representable values: the reference's forcing files are float32
    netCDF (test_data/{drb_2yr,ucb_2yr}/prcp.nc), and the e2e leg uploads them
    in that type (pws_run_host_f32)."""
    import torch

    nd, ntile = spec["ndays"], spec["ntile"]
    nb = f0["prcp"].shape[1]
    nh = spec["nhru_keep"] or nb * ntile
    dtemp, pscale = tile_offsets(ntile, rank)
    out = {}
    tidx = torch.arange(nd, device=dev) % f0["prcp"].shape[0]
    for k in ("prcp", "tmax", "tmin"):
        base = torch.from_numpy(f0[k]).to(dev)[tidx]  # [nd, nb]
        full = torch.empty((nd, nb * ntile), dtype=torch.float64, device=dev)
        for t in range(ntile):
            sl = slice(t * nb, (t + 1) * nb)
            if k == "prcp":
                full[:, sl] = base * float(pscale[t])
            else:
                full[:, sl] = base + float(dtemp[t])
        out[k] = full[:, :nh].float().double().contiguous()
        del full, base
    return out


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


def run_reference(args, spec):
    """--impl reference: the CPU oracle port on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import scenarios

    p0, f0, start = scenarios.load_drb()
    ncores = os.cpu_count() or 1
    nthreads = max(1, min(ncores, 64))
    tiles_per_thread = 1
    ndays = min(spec["ndays"], 3653)
    rates = []
    for it in range(args.warmup + args.steps):
        r, units, dt = cpu_oracle_rate(p0, f0, start, tiles_per_thread, ndays, nthreads)
        if it >= args.warmup:
            rates.append((units, dt))
    units = sum(u for u, _ in rates)
    dt = sum(d for _, d in rates)
    value = units / dt
    sample = (f"{nthreads} threads x {tiles_per_thread} DRB tile(s) (765 HRUs, 456 segments) x "
              f"{ndays} days per step, full chain incl. routing and budgets, segment outputs kept")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(len(rates), 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": spec["label"], "note": "bounded sample of the workload on host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
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
        torch.cuda.synchronize()
        eng.timing()
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
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": None, "kernel": "k_route_chunks",
                     "note": "48 B per segment-day (8 read + 5x8 written); the kernel is bound by the "
                             "serial hourly recurrence along the network depth, not by HBM"},
        "gpu_launches": eng.launch_count(), "clocks": clocks}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="conus")
    ap.add_argument("--outputs", default="full")
    ap.add_argument("--block-days", type=int, default=256)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-e2e", action="store_true")
    # measured on the CONUS workload (float32 forcings, HRU-days/s): 64 -> 3.50e9, 128 -> 3.53e9,
    # 256 -> 3.88e9, 384 -> 3.76e9, 512 -> 3.57e9 (profiles/r01_e2e_block_days.jsonl)
    ap.add_argument("--e2e-block-days", type=int, default=256)
    ap.add_argument("--e2e-forcing", choices=("f32", "f64"), default="f32",
                    help="type of the host forcing arrays of the e2e leg: float32 as in the reference's "
                         "forcing files (widened on the device, exact) or float64")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--layout", choices=("tiled", "rows"), default="tiled",
                    help="device-resident results as [day][tile][variable][H] (pws_output_tile; needs "
                         "--outputs full) or as [day][variable][nhru] rows")
    args = ap.parse_args()
    if args.workload in ("channel", "channel_chained"):
        return run_channel_only(args)
    spec = workload_spec(args.workload)
    if args.impl == "reference":
        return run_reference(args, spec)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pywatershed_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the one JSON line: NCCL prints its version banner to stdout at the
        # VERSION and WARN levels; anything more verbose the user asked for goes to stderr
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            del os.environ["NCCL_DEBUG"]
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__ as g

    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    from pywatershed_b200 import Engine

    p0, f0, pt, start = build_domain(spec, rank)
    eng = Engine(pt, start, device=local_rank)
    hru_vars, seg_vars = output_sets(eng.reg, args.outputs)
    eng.select_outputs(hru_vars, seg_vars)
    nhru, nseg, nd = eng.nhru, eng.nseg, spec["ndays"]
    forc = device_forcings(f0, spec, rank, dev)
    T = max(1, min(args.block_days, nd))
    nf = len(eng.sel_f64)
    ni = len(eng.sel_i32)
    tiled = args.layout == "tiled" and args.outputs == "full"
    if tiled:
        eng.set_tiled_outputs(True)
    # (the tiled layout pads the last tile: [T, tiles, n, H] has the same size as [T, n, padded])
    onhru = eng.output_tile()[1] if tiled else nhru
    of = torch.empty((T, max(nf, 1), onhru), dtype=torch.float64, device=dev)
    oi = torch.empty((T, max(ni, 1), onhru), dtype=torch.int32, device=dev)
    osg = torch.empty((T, max(len(seg_vars), 1), max(nseg, 1)), dtype=torch.float64, device=dev)
    viol_t = torch.zeros((nd, 6), dtype=torch.int32).pin_memory()
    viol = viol_t.numpy()
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_pass(timed):
        """reset (untimed) + every block of days; returns device ms."""
        eng.reset()
        barrier()
        eng.timer_mark(0)
        d = 0
        while d < nd:
            n = min(T, nd - d)
            o = d * nhru * 8
            eng.run_device(n, forc["prcp"].data_ptr() + o, forc["tmax"].data_ptr() + o,
                           forc["tmin"].data_ptr() + o, of.data_ptr(), oi.data_ptr(),
                           osg.data_ptr(), viol=viol[d:d + n], sync=False)
            d += n
        eng.timer_mark(1)
        eng.sync()
        barrier()
        return eng.timer_elapsed_ms()

    for _ in range(args.warmup):
        one_pass(False)
    sampler = ClockSampler(local_rank)
    l0 = eng.launch_count()
    if rank == 0:
        sampler.start()
    step_ms, col_ms, ch_ms, ncol = [], 0.0, 0.0, 0
    for _ in range(args.steps):
        ms = one_pass(True)
        step_ms.append(ms)
        c, h = eng.timing()
        col_ms += c
        ch_ms += h
        ncol += eng.last_launches()[0]
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launch_count() - l0
    viol_total = int(viol.sum())
    t_dev = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    total_ms = float(t_dev.item())
    units_per_step = nhru * nd * world
    value = units_per_step * args.steps / (total_ms * 1e-3)

    # ---- e2e: host-facing call, pinned host buffers, H2D + D2H inside the region
    e2e = None
    if not args.no_e2e:
        e2e_seg = ["seg_outflow"]  # streamflow at every segment: the product of an NHM run
        if tiled:
            eng.set_tiled_outputs(False)  # a pws_run_device layout
        eng.select_outputs([], e2e_seg)
        f32 = args.e2e_forcing == "f32"
        hf = {k: (v.float() if f32 else v).cpu().pin_memory() for k, v in forc.items()}
        run_host = eng.L.pws_run_host_f32 if f32 else eng.L.pws_run_host
        e2e_T = args.e2e_block_days
        eng.set_block_days(e2e_T)
        # pin the result buffers too (the library drains with cudaMemcpyAsync)
        cal_nd = nd
        res_seg = torch.empty((cal_nd, len(e2e_seg), nseg), dtype=torch.float64).pin_memory()
        res_viol = np.zeros((cal_nd, 6), dtype=np.int32)
        from pywatershed_b200.engine import calendar

        cal = calendar(start, nd)
        dummy_f = np.empty(1)
        dummy_i = np.empty(1, dtype=np.int32)

        def e2e_pass():
            eng.reset()
            barrier()
            t0 = time.perf_counter()
            eng._check(run_host(
                eng.h, nd, cal.ctypes.data, hf["prcp"].data_ptr(), hf["tmax"].data_ptr(),
                hf["tmin"].data_ptr(), dummy_f.ctypes.data, dummy_i.ctypes.data,
                res_seg.data_ptr(), res_viol.ctypes.data))
            barrier()
            return time.perf_counter() - t0

        e2e_pass()
        e2e_pass()
        ts = [e2e_pass() for _ in range(max(1, min(args.steps, 3)))]
        te = torch.tensor([sum(ts)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": units_per_step * len(ts) / float(te.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(3 * nd * nhru * (4 if f32 else 8)) * world,
               "forcing_dtype": "float32 (as stored in the reference's forcing files; widened on the "
                                "device)" if f32 else "float64",
               "d2h_bytes_per_step": int(nd * (len(e2e_seg) * nseg * 8 + 24)) * world,
               "outputs": "seg_outflow [ndays, nsegment] + 6 budget verdicts per day",
               "block_days": e2e_T, "checksum_seg_outflow_last_day": float(res_seg[-1, 0].sum())}
        eng.select_outputs(hru_vars, seg_vars)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_hru_column), measured live
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    bpu = algorithmic_bytes_per_hru_day(eng.reg, hru_vars)
    avg_launch_s = (col_ms / max(ncol, 1)) * 1e-3
    units_per_launch = nhru * nd * args.steps / max(ncol, 1)
    achieved = units_per_launch * bpu / avg_launch_s / 1e9
    # DRAM bytes per launch from the committed ncu capture (profiles/), scaled to
    # this run's launch size; only meaningful for the output mode it was taken in
    traffic, fp64 = None, None
    prof = ROOT / "profiles" / "r01_column_traffic.json"
    if prof.exists():
        try:
            pj = json.loads(prof.read_text())
            if args.outputs == "full":
                traffic = pj["dram_bytes_per_launch"] * units_per_launch / pj["hru_days_per_launch"]
            # the other roofline of the kernel: FP64 work per HRU-day counted by ncu
            # (DADD + DMUL + 2 DFMA thread instructions) against the FP64 pipe's peak
            tf = pj["fp64_flops_per_hru_day"] * units_per_launch / avg_launch_s / 1e12
            fp64 = {"achieved": tf, "peak": pj["fp64_peak_tflops"], "unit": "TFLOP/s",
                    "frac": tf / pj["fp64_peak_tflops"],
                    "flops_per_hru_day": pj["fp64_flops_per_hru_day"]}
        except Exception:
            traffic, fp64 = None, None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "kernel": "k_hru_column",
                "peak_source": peak_src, "bytes_per_hru_day": bpu,
                "hru_days_per_launch": units_per_launch,
                "avg_launch_ms": avg_launch_s * 1e3,
                "share_of_step": col_ms / max(sum(step_ms), 1e-9),
                "channel_share_of_step": ch_ms / max(sum(step_ms), 1e-9),
                "fp64": fp64}

    cpu = None
    if not args.no_cpu and world == 1:
        ncores = os.cpu_count() or 1
        nthreads = max(1, min(ncores, 64))
        r, units, dt = cpu_oracle_rate(p0, f0, start, 1, min(nd, 3653), nthreads)
        cpu = {"value": r, "unit": UNIT, "cores": nthreads, "kind": "port",
               "sample": f"{nthreads} threads x 1 DRB tile x {min(nd, 3653)} days "
                         f"({units:.3g} HRU-days, {dt:.1f} s), C oracle, full chain + routing + budgets"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": spec["label"] + (f" per GPU x {world} GPUs" if world > 1 else ""),
                   "outputs": args.outputs, "layout": "tiled" if tiled else "rows", "n_output_vars": len(hru_vars) + len(seg_vars),
                   "block_days": T, "nhru_per_gpu": nhru, "nsegment_per_gpu": nseg, "ndays": nd,
                   "l2": "inputs larger than L2 (forcings 3 x %.1f GB per GPU, streamed once per step)"
                         % (nd * nhru * 8 / 1e9),
                   "parallelism": f"basin-sharded x{world}, no data-path collective",
                   "kernel": eng.kernel_info(), "budget_violations": viol_total},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
