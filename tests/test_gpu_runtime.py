"""The host runtime of libpwsb200 (round 2): column lanes and the routing stream,
shared-forcing ensembles, device budget accumulations, tiled results through
pws_run_host, restart including PRMSChannel's state, and the budget verdict's
positive control.  Everything is a size-independent property (two schedules /
layouts of the same work must agree bit for bit) or a comparison with the
oracle; all of it goes through the C ABI.  GPU only."""
import numpy as np
import pytest

import prms_oracle as po
import scenarios

pytestmark = pytest.mark.gpu


def _tiled(drb, ntile, nd):
    p, f, start = drb
    pt, ft = scenarios.tile(p, {k: v[:nd] for k, v in f.items()}, ntile)
    return pt, ft, start


def _run(pt, ft, start, hv, sv, block_days=None, **kw):
    from pywatershed_b200 import Engine

    opts = {k: kw.pop(k) for k in ("overlap_routing", "column_hrus", "wide_column", "one_cta_carveout",
                                   "nhm_sink") if k in kw}
    e = Engine(pt, start, **kw)
    for k, v in opts.items():
        e._opt(k, v)
    if opts:
        e._check(e.L.pws_init(e.h))
    e.select_outputs(hv, sv)
    if block_days:
        e.set_block_days(block_days)
    got, viol = e.run_host(ft["prcp"], ft["tmax"], ft["tmin"])
    info = (e.last_blocks(), e.last_launches())
    e.close()
    return got, viol, info


@pytest.mark.parametrize("lanes,overlap", [(1, 0), (3, 1), (8, 1), (0, 1)])
def test_lanes_and_routing_overlap_are_transparent(drb, lanes, overlap):
    """The same 5-tile domain over 5 blocks of days: one column launch per block
    with the blocks serialised (the round-1 schedule) against several lanes with
    the routing of block b beside the column of block b+1.  Bit-identical."""
    nd = 37
    pt, ft, start = _tiled(drb, 5, nd)
    hv = ["pkwater_equiv", "soil_moist", "sroff_vol", "gwres_flow_vol", "newsnow", "hru_actet"]
    sv = list(po.SEG_VARS)
    base, vb, _ = _run(pt, ft, start, hv, sv, block_days=8, lanes=1, overlap_routing=0, column_hrus=128)
    got, vg, (blocks, launches) = _run(pt, ft, start, hv, sv, block_days=8, lanes=lanes,
                                       overlap_routing=overlap, column_hrus=128)
    assert blocks[0] == 5
    if lanes:
        assert blocks[1] == lanes and launches[0] == 5 * lanes
    for k in hv + sv:
        np.testing.assert_array_equal(np.asarray(got[k]), np.asarray(base[k]), err_msg=k)
    np.testing.assert_array_equal(vg, vb)
    assert vb.sum() == 0


@pytest.mark.parametrize("outputs", ["full", "select", "none"])
def test_column_kernel_variants_agree(drb, outputs):
    """The instantiations of k_hru_column a domain may be run with -- 256 or 128
    HRUs per CTA, and for grids of at most one CTA per SM the uncapped-register
    `wide` variant with the one-CTA carve-out (the default at this size) or the
    capped one -- are the same arithmetic: bit-identical results in every sink mode."""
    nd = 45
    pt, ft, start = _tiled(drb, 2, nd)
    reg_h, reg_s = po.HRU_VARS, po.SEG_VARS
    hv = {"full": list(reg_h), "select": ["pkwater_equiv", "soil_moist", "newsnow", "sroff", "gwres_stor"],
          "none": []}[outputs]
    sv = list(reg_s)
    base, vb, _ = _run(pt, ft, start, hv, sv, block_days=16, column_hrus=256)
    for opts in (dict(column_hrus=128, wide_column=1), dict(column_hrus=128, wide_column=0),
                 dict(column_hrus=128, wide_column=0, one_cta_carveout=0), dict(column_hrus=192)):
        got, vg, _ = _run(pt, ft, start, hv, sv, block_days=16, **opts)
        for k in hv + sv:
            np.testing.assert_array_equal(np.asarray(got[k]), np.asarray(base[k]), err_msg=f"{k} {opts}")
        np.testing.assert_array_equal(vg, vb)


def test_full_chain_with_overlapped_routing_launches(drb):
    """The full chain over a deep (chained) network in 8-day blocks: routing launches
    of consecutive blocks on two streams (route_overlap 1, the default for networks
    deeper than 160 levels) against serialised ones -- results, budget verdicts and
    device accumulations bit-identical."""
    nd = 41
    pt, ft, start = _tiled(drb, 3, nd)
    p = drb[0]
    nseg0 = p["tosegment"].shape[0]
    tos0 = np.asarray(p["tosegment"], dtype=np.int64)
    dist = np.zeros(nseg0, dtype=np.int64)
    outlet = np.zeros(nseg0, dtype=np.int64)
    for s0 in range(nseg0):
        j, n = s0, 0
        while tos0[j] > 0:
            j, n = tos0[j] - 1, n + 1
        dist[s0], outlet[s0] = n, j
    main = int(np.argmax(np.bincount(outlet)))
    head = int(np.argmax(np.where(outlet == main, dist, -1)))
    tos = np.asarray(pt["tosegment"]).copy()
    for t in range(2):
        tos[t * nseg0 + main] = (t + 1) * nseg0 + head + 1
    pt = dict(pt)
    pt["tosegment"] = tos
    hv = ["sroff_vol", "gwres_flow_vol", "soil_moist"]
    sv = list(po.SEG_VARS)
    from pywatershed_b200 import Engine

    res = {}
    for overlap in (0, 1):
        e = Engine(pt, start)
        e._opt("route_overlap", overlap)
        e.select_outputs(hv, sv)
        e.select_accumulations(["sroff_vol"], ["seg_outflow", "seg_stor_change"])
        e.set_block_days(8)
        got, viol = e.run_host(ft["prcp"], ft["tmax"], ft["tmin"])
        res[overlap] = ({k: np.array(got[k]) for k in hv + sv}, np.array(viol), e.accumulations())
        e.close()
    for k in hv + sv:
        np.testing.assert_array_equal(res[1][0][k], res[0][0][k], err_msg=k)
    np.testing.assert_array_equal(res[1][1], res[0][1])
    for k, v in res[0][2].items():
        np.testing.assert_array_equal(res[1][2][k], v, err_msg=k)
    ref, _ = po.Oracle(pt, start=start).run(ft["prcp"], ft["tmax"], ft["tmin"], (), ("seg_outflow",))
    np.testing.assert_allclose(res[1][0]["seg_outflow"], ref["seg_outflow"], rtol=1e-9, atol=1e-10)


def test_default_output_selection_kernel_equals_the_general_one(drb):
    """The reference's default output selection (nhm.control's netcdf_output_var_names,
    71 HRU variables) in registry order runs k_hru_column<SINK_NHM>, which knows the
    selection at compile time and does not compute what nobody asked for: the same
    bits as the general SINK_SELECT kernel, in every launch shape; any other order or
    subset takes the general kernel."""
    import json
    import pathlib as pl

    from pywatershed_b200 import Engine

    nd = 45
    pt, ft, start = _tiled(drb, 2, nd)
    names = json.loads((pl.Path(__file__).parent / "golden" / "nhm_default_vars.json").read_text())
    reg = Engine(pt, start).reg
    hv = [v for v in reg.hru_vars if v in names]
    sv = [v for v in reg.seg_vars if v in names]
    assert len(hv) == 71
    base, vb, _ = _run(pt, ft, start, hv, sv, block_days=16, nhm_sink=0, column_hrus=256)
    for opts in (dict(column_hrus=256), dict(column_hrus=128, wide_column=1), dict(column_hrus=128, wide_column=0),
                 dict(column_hrus=192)):
        got, vg, _ = _run(pt, ft, start, hv, sv, block_days=16, nhm_sink=1, **opts)
        for k in hv + sv:
            np.testing.assert_array_equal(np.asarray(got[k]), np.asarray(base[k]), err_msg=f"{k} {opts}")
        np.testing.assert_array_equal(vg, vb)
    # reversed order: not the compile-time layout, the general kernel serves it
    got, _, _ = _run(pt, ft, start, hv[::-1], sv, block_days=16, nhm_sink=1)
    for k in hv:
        np.testing.assert_array_equal(np.asarray(got[k]), np.asarray(base[k]), err_msg=k)


def test_device_resident_blocks_overlap_safely(drb):
    """pws_run_device called block after block without a sync in between (what
    bench.py does): the scratch of a block (calendar, verdict counters, lateral
    inflow) is double-buffered, so results equal the synchronous sequence."""
    import torch

    from pywatershed_b200 import Engine

    nd, T = 40, 7
    pt, ft, start = _tiled(drb, 4, nd)
    dev = torch.device("cuda:0")
    forc = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in ft.items()}
    hv, sv = ["soil_moist", "pkwater_equiv", "transp_on"], ["seg_outflow", "seg_stor_change"]
    res = {}
    for mode in ("sync", "async"):
        e = Engine(pt, start, lanes=4)
        e.select_outputs(hv, sv)
        nblk = (nd + T - 1) // T
        of = torch.zeros((nblk, T, 2, e.nhru), dtype=torch.float64, device=dev)
        oi = torch.zeros((nblk, T, 1, e.nhru), dtype=torch.int32, device=dev)
        osg = torch.zeros((nblk, T, 2, e.nseg), dtype=torch.float64, device=dev)
        viol = torch.zeros((nd, 6), dtype=torch.int32, pin_memory=True).numpy()
        for b in range(nblk):
            d, n = b * T, min(T, nd - b * T)
            o = d * e.nhru * 8
            e.run_device(n, forc["prcp"].data_ptr() + o, forc["tmax"].data_ptr() + o,
                         forc["tmin"].data_ptr() + o, of[b].data_ptr(), oi[b].data_ptr(),
                         osg[b].data_ptr(), viol=viol[d:d + n], sync=mode == "sync")
        e.sync()
        res[mode] = (of.cpu().numpy(), oi.cpu().numpy(), osg.cpu().numpy(), viol.copy())
        e.close()
    for a, b in zip(res["sync"], res["async"]):
        np.testing.assert_array_equal(a, b)
    ref, _ = po.Oracle(pt, start=start).run(ft["prcp"], ft["tmax"], ft["tmin"], hv, sv)
    got = np.concatenate([res["async"][2][b, :min(T, nd - b * T), 0] for b in range((nd + T - 1) // T)])
    np.testing.assert_allclose(got, ref["seg_outflow"], rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("nbase,nd", [(765, 20), (764, 20), (300, 12)])
def test_shared_forcing_ensemble_equals_tiled_forcings(drb, nbase, nd):
    """forcing_period: a member-major ensemble whose members share the weather
    reads [ndays, P] forcings (HRU i -> column i mod P).  It must equal the same
    parameters run with the forcings repeated per member, bit for bit -- with
    an odd period (per-thread loads), an even one (TMA rows that wrap around the
    end of the period inside a CTA) and one shorter than a CTA."""
    import bench
    from pywatershed_b200 import Engine

    p, f, start = drb
    nm = 4
    keep = np.arange(nbase)
    pb = {k: (np.asarray(v)[..., keep] if np.asarray(v).ndim and np.asarray(v).shape[-1] == 765 else v)
          for k, v in p.items()}
    pb["hru_segment"] = np.asarray(pb["hru_segment"])
    dummy = {k: np.zeros((2, nbase)) for k in ("prcp", "tmax", "tmin")}
    pt, _ = scenarios.tile(pb, dummy, nm, temp_shift=False)
    rng = np.random.default_rng(5)
    for k in ("soil_moist_max", "carea_max", "freeh2o_cap", "gwflow_coef"):
        pt[k] = np.asarray(pt[k]) * rng.uniform(0.8, 1.2, size=np.asarray(pt[k]).shape)
    fb = {k: np.ascontiguousarray(v[:nd, :nbase]) for k, v in f.items()}
    hv = ["pkwater_equiv", "soil_moist", "sroff_vol", "swrad", "pptmix"]
    sv = ["seg_outflow"]
    e1 = Engine(pt, start, forcing_period=nbase)
    e1.select_outputs(hv, sv)
    e1.set_block_days(7)
    a, va = e1.run_host(fb["prcp"], fb["tmax"], fb["tmin"])
    e1.close()
    e2 = Engine(pt, start)
    e2.select_outputs(hv, sv)
    e2.set_block_days(7)
    ft = {k: np.ascontiguousarray(np.tile(v, (1, nm))) for k, v in fb.items()}
    b, vb = e2.run_host(ft["prcp"], ft["tmax"], ft["tmin"])
    e2.close()
    for k in hv + sv:
        np.testing.assert_array_equal(np.asarray(a[k]), np.asarray(b[k]), err_msg=k)
    np.testing.assert_array_equal(va, vb)
    # members differ (the parameters were perturbed), the weather does not
    assert not np.array_equal(a["soil_moist"][:, :nbase], a["soil_moist"][:, nbase:2 * nbase])
    np.testing.assert_array_equal(a["swrad"][:, :nbase], a["swrad"][:, nbase:2 * nbase])
    assert bench.NB_HRU == 765


def test_device_accumulations_are_the_reference_sums(drb):
    """pws_select_accumulations: budget terms summed per unit on the device, day
    by day in the reference's order (base/budget.py:262-269), without being
    drained.  Equal to the sequential numpy sum of the drained series, bit for bit."""
    from pywatershed_b200 import Engine

    p, f, start = drb
    nd = 45
    terms_h = ["net_rain", "snowmelt", "hru_actet", "soil_to_gw", "gwres_stor_change", "channel_sroff_vol"]
    terms_s = ["channel_outflow_vol", "seg_stor_change"]
    e = Engine(p, start)
    e.select_outputs(terms_h, terms_s)
    full, _ = e.run_host(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd])
    e.close()
    e = Engine(p, start)
    e.select_outputs(["pkwater_equiv", "net_rain"], ["seg_outflow"])  # drained; overlaps one term
    e.select_accumulations(terms_h, terms_s)
    e.set_block_days(16)
    got, _ = e.run_host(f["prcp"][:30], f["tmax"][:30], f["tmin"][:30])
    assert set(got) == {"pkwater_equiv", "net_rain", "seg_outflow"} and got["net_rain"].shape == (30, 765)
    np.testing.assert_array_equal(got["net_rain"], full["net_rain"][:30])
    e.set_accumulate(False)  # paused: days 30..37 do not count
    e.run_host(f["prcp"][30:38], f["tmax"][30:38], f["tmin"][30:38])
    e.set_accumulate(True)
    e.run_host(f["prcp"][38:nd], f["tmax"][38:nd], f["tmin"][38:nd])
    acc = e.accumulations()
    days = list(range(30)) + list(range(38, nd))
    for k in terms_h + terms_s:
        want = np.zeros(full[k].shape[1])
        for d in days:
            want = want + full[k][d]
        np.testing.assert_array_equal(acc[k], want, err_msg=k)
    e.reset()
    assert all((v == 0).all() for v in e.accumulations().values())
    e.close()


@pytest.mark.parametrize("ntile,column_hrus", [(1, 0), (3, 128), (2, 256), (2, 192)])
def test_tiled_results_through_run_host(drb, ntile, column_hrus):
    """The tiled result layout ([day][tile][variable][H]) drained by pws_run_host:
    every variable un-tiled on the host equals the rows layout."""
    from pywatershed_b200 import Engine

    nd = 19
    pt, ft, start = _tiled(drb, ntile, nd)
    e = Engine(pt, start)
    e.select_outputs(e.reg.hru_vars, e.reg.seg_vars)
    e.set_block_days(8)
    rows, vr = e.run_host(ft["prcp"], ft["tmax"], ft["tmin"])
    e.close()
    e = Engine(pt, start)
    if column_hrus:
        e._opt("column_hrus", column_hrus)
        e._check(e.L.pws_init(e.h))
    e.select_outputs(e.reg.hru_vars, e.reg.seg_vars)
    e.set_tiled_outputs(True)
    e.set_block_days(8)
    til, vt = e.run_host(ft["prcp"], ft["tmax"], ft["tmin"], pinned_outputs=True)
    for k in po.HRU_VARS + po.SEG_VARS:
        assert k in til
        np.testing.assert_array_equal(np.asarray(til[k]), np.asarray(rows[k]), err_msg=k)
    np.testing.assert_array_equal(vr, vt)
    e.close()


def test_restart_with_the_channel(drb):
    """save_state after 30 days -> load_state into a fresh engine -> 30 more days:
    identical to the uninterrupted run INCLUDING PRMSChannel, whose outflow_ts and
    seg_inflow carry over from day to day (prms_channel.py:384-386, 456-585)."""
    from pywatershed_b200 import Engine

    p, f, start = drb
    hv, sv = list(po.HRU_VARS), list(po.SEG_VARS)

    def eng():
        e = Engine(p, start, channel_chunk=100)  # trees cut into several chunks: cross-chunk edges too
        e.select_outputs(hv, sv)
        return e

    e = eng()
    a, va = e.run_host(f["prcp"][:60], f["tmax"][:60], f["tmin"][:60])
    e.close()
    e1 = eng()
    assert e1.seg_states == ["outflow_ts", "seg_inflow"]
    e1.run_host(f["prcp"][:30], f["tmax"][:30], f["tmin"][:30])
    st = e1.save_state()
    assert st["outflow_ts"].shape == (456,) and (st["outflow_ts"] > 0).any()
    e1.close()
    e2 = eng()
    e2.load_state(st)
    assert e2.day == 30
    b, vb = e2.run_host(f["prcp"][30:60], f["tmax"][30:60], f["tmin"][30:60])
    for k in hv + sv:
        np.testing.assert_array_equal(np.asarray(b[k]), np.asarray(a[k])[30:], err_msg=k)
    np.testing.assert_array_equal(vb, va[30:])
    # a state that is read back is what was set (segment order is the caller's)
    x = np.arange(456, dtype=np.float64)
    e2.set_state("seg_inflow", x)
    np.testing.assert_array_equal(e2.get_state("seg_inflow"), x)
    e2.close()


def _canopy_hrus(p, n):
    """HRUs with a real canopy in both seasons (an interception store that matters)."""
    ok = (np.asarray(p["covden_sum"]) > 0.2) & (np.asarray(p["covden_win"]) > 0.2) & \
         (np.asarray(p["hru_type"]) == 1) & (np.asarray(p["cov_type"]) > 0)
    return [int(i) for i in np.where(ok)[0][:n]]


def test_budget_verdict_fires_on_an_injected_imbalance(drb):
    """Positive control of the device-side budget check (budget_open, the ballot
    reduce, k_channel_budget): water put into a store behind the model's back
    must be counted on exactly the next day, for exactly the tampered units and
    processes, and nowhere else."""
    from pywatershed_b200 import Engine

    p, f, start = drb
    hrus = _canopy_hrus(p, 4)
    assert len(hrus) == 4
    e = Engine(p, start)
    e.select_outputs(["gwres_stor"], ["seg_outflow"])
    _, v0 = e.run_host(f["prcp"][:10], f["tmax"][:10], f["tmin"][:10])
    assert v0.sum() == 0
    # A store whose "old" value is re-read from the tampered state stays consistent: the
    # groundwater budget's storage change is gwres_stor(today) - gwres_stor(state)
    g = e.get_state("gwres_stor")
    g[hrus] += 5.0
    e.set_state("gwres_stor", g)
    _, v1 = e.run_host(f["prcp"][10:12], f["tmax"][10:12], f["tmin"][10:12])
    assert v1.sum() == 0
    # The channel's global balance (prms_channel.py:165-174) is an identity of the
    # routing scheme for finite flows; what the global check can catch is a
    # non-finite sum: one NaN in one headwater's outflow_ts reaches the outlet and
    # np.allclose(NaN, ...) is False (base/budget.py:388-416) from then on
    o = e.get_state("outflow_ts")
    head = int(np.where(~np.isin(np.arange(456) + 1, np.asarray(p["tosegment"])))[0][0])
    o[head] = np.nan
    e.set_state("outflow_ts", o)
    _, v2 = e.run_host(f["prcp"][12:15], f["tmax"][12:15], f["tmin"][12:15])
    e.close()
    assert (v2[:, 5] == 1).all() and v2[:, :5].sum() == 0, v2
    # The unit-basis check: canopy's hru_intcpstor_old is a state of its own, so water
    # added to intcp_stor alone is storage the day's fluxes do not explain
    e = Engine(p, start)
    e.select_outputs(["hru_intcpstor"], [])
    e.run_host(f["prcp"][:10], f["tmax"][:10], f["tmin"][:10])
    s = e.get_state("intcp_stor")
    s[hrus] += 0.5
    e.set_state("intcp_stor", s)
    _, v3 = e.run_host(f["prcp"][10:14], f["tmax"][10:14], f["tmin"][10:14])
    e.close()
    assert v3[0, 0] == len(hrus), v3[:2]
    assert v3[1:, 0].sum() == 0 and v3[:, 1:].sum() == 0, v3


def test_model_raises_on_the_day_of_an_imbalance(drb):
    """The same through the Model API: budget_type="error" raises the reference's
    ValueError (base/budget.py:380-383) on the day the device reports, "warn"
    warns and runs on."""
    import pywatershed_b200 as pwsb
    from test_gpu_model import NHM, _model

    p, f, start = drb
    for btype in ("error", "warn"):
        m = _model(p, f, start, 30, budget_type=btype)
        m.run(n_time_steps=10, finalize=False)
        eng = m._engine
        s = eng.get_state("intcp_stor")
        s[_canopy_hrus(p, 2)] += 0.5
        eng.set_state("intcp_stor", s)
        if btype == "error":
            with pytest.raises(ValueError, match="flux unit balance.*PRMSCanopy.*2 unit"):
                m.run(finalize=False)
            assert m.control.itime_step == 10  # the day after the 10 clean ones
        else:
            with pytest.warns(UserWarning, match="flux unit balance.*PRMSCanopy"):
                m.run(finalize=False)
            assert m.control.itime_step == 29
        m.finalize()
    assert pwsb.PRMSCanopy.__name__ in NHM


def test_cross_chunk_deadlock_guard_reports(drb):
    """k_route_chunks waits for upstream chunks through progress counters; the
    guard (PWS_ROUTE_MAX_IDLE -> error bit 16) must surface as an error instead of
    a hang.  Provoked with a library built with a tiny idle budget and a network
    cut into one-segment chunks, so that waits are long relative to it."""
    import os
    import pathlib as pl
    import subprocess
    import sys

    root = pl.Path(__file__).resolve().parent.parent
    lib = root / "tests" / "_build" / "libpwsb200_idle.so"
    lib.parent.mkdir(exist_ok=True)
    if not lib.exists():
        sys.path.insert(0, str(root))
        import __graft_entry__ as g

        subprocess.run(["nvcc", *g.NVCC_FLAGS, "-DPWS_ROUTE_MAX_IDLE=0", "-o", str(lib),
                        str(g.CSRC / "pws_b200.cu")], check=True)
    code = (
        "import sys; sys.path[:0] = [%r, %r]\n"
        "import numpy as np, scenarios\n"
        "from pywatershed_b200 import Engine\n"
        "p, f, start = scenarios.load_drb()\n"
        "e = Engine(p, start, channel_chunk=1)\n"
        "e.select_outputs([], ['seg_outflow'])\n"
        "try:\n"
        "    e.run_host(f['prcp'][:20], f['tmax'][:20], f['tmin'][:20])\n"
        "    print('NOERROR')\n"
        "except ValueError as exc:\n"
        "    print('ERR', exc)\n" % (str(root), str(root / "tests")))
    env = dict(os.environ, PWS_B200_LIB=str(lib))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "waited too long for its upstream segments" in r.stdout, r.stdout


def test_sanitize_case_runs_clean():
    """tools/sanitize_case.py -- the small case written for compute-sanitizer (every
    kernel variant: several column tiles, two lanes, routing beside the next block,
    cross-chunk edges, a shared-forcing ensemble in the tiled layout, accumulations,
    PRMSEt) -- checked against the oracle.  The sanitizer itself is closed on the
    build pool (profiles/r02_sanitizer_closed.log); this keeps the case alive."""
    import pathlib as pl
    import subprocess
    import sys

    root = pl.Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, str(root / "tools" / "sanitize_case.py")], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "sanitize case ok" in r.stdout
