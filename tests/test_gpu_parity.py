"""Parity of the CUDA path (through the C ABI, libpwsb200.so) with the CPU
oracle on the same inputs.  Tolerances are BASELINE.json's: rtol 1e-9 /
atol 1e-10 for every state and flux, flags bit-exact, disagreements counted
and traced to ghost-pack threshold ties (tests/parity.py).  GPU only.
"""
import numpy as np
import pytest

import parity
import prms_oracle as po
import scenarios

pytestmark = pytest.mark.gpu


def _engine(p, start, **kw):
    from pywatershed_b200 import Engine

    e = Engine(p, start, **kw)
    e.select_outputs(e.reg.hru_vars, e.reg.seg_vars)
    return e


def _oracle(p, f, start, nd):
    o = po.Oracle(p, start=start)
    st = o.soltabs()
    res, viol = o.run(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd], po.HRU_VARS, po.SEG_VARS)
    return st, res, viol


# Share of HRUs allowed to leave tolerance through a ghost-pack tie over a
# 1-2 year run (every one is checked to BE such a tie; anything else fails).
# Measured on the B200 with the correctly rounded tables of k_soltab
# (pws_crmath.cuh): 0 of 765 on DRB and on drbx in 731 days (with CUDA's own
# sin / cos / asin ... in k_soltab: 2 and 42), 2 of 765 on a tile of the
# 10-year CONUS domain (164 before); profiles/r02_tie_census.log.
MAX_TIE_FRACTION = 0.01
TIE_LOG = []  # (label, summary) of every audited comparison of this session


def _assert_chain_parity(got, ref, label=""):
    rep = parity.compare_columns(got, ref, po.HRU_VARS, got["pkwater_equiv"], ref["pkwater_equiv"])
    assert rep["mismatches"] == [], rep["mismatches"][:5]
    n = ref["pkwater_equiv"].shape[1]
    assert len(rep["ties"]) <= max(2, int(MAX_TIE_FRACTION * n)), rep["ties"]
    # tied HRUs stay in the comparison, with the bounds a ghost-pack tie explains
    audit = parity.audit_ties(got, ref, rep, got["pkwater_equiv"], ref["pkwater_equiv"])
    TIE_LOG.append((label, audit))
    print(f"ghost-pack ties {label}: {len(rep['ties'])} of {n} HRUs "
          f"({audit['existence']} pack present in one run only, {audit['residue']} residue differs); "
          f"after a tie: cumulative ET / flow off by at most {audit['max_cum_rel']:.2e}, "
          f"storages by at most {audit['max_store_abs']:.3g} in")
    clean = rep["first_bad"] < 0
    for k in po.INT_VARS + po.BOOL_VARS:
        assert (np.asarray(got[k], dtype=np.float64)[:, clean] == ref[k][:, clean]).all(), k
    return rep


def test_registry_matches_oracle():
    from pywatershed_b200 import Registry

    reg = Registry.get()
    assert reg.hru_vars == po.HRU_VARS and reg.seg_vars == po.SEG_VARS
    for v in po.INT_VARS:
        assert reg.dtype(v) is np.int32
    assert reg.dtype("iasw") is np.bool_


def test_soltab(drb):
    p, f, start = drb
    e = _engine(p, start)
    st = e.soltab()
    ref = po.Oracle(p, start=start).soltabs()
    differ = total = 0
    for k in st:
        np.testing.assert_allclose(st[k], ref[k], rtol=1e-9, atol=1e-10, err_msg=k)
        differ += int((np.asarray(st[k]).view(np.int64) != ref[k].view(np.int64)).sum())
        total += ref[k].size
    # k_soltab rounds every elementary function correctly (pws_crmath.cuh); the oracle's libm
    # does so for all but ~0.15 % of its calls, so nearly all table entries carry the same bits
    print(f"k_soltab vs the oracle's libm: {differ} of {total} table entries differ in bits")
    assert differ <= 0.01 * total
    g = np.load(scenarios.GOLDEN / "drb_golden.npz")
    for k in st:  # and directly against the reference's own tables
        np.testing.assert_allclose(st[k][:, g["hru_sample"]], g[k], rtol=1e-9, atol=1e-10)


def test_soltab_host_option(drb):
    """Option soltab_device 0: the PRMSSolarGeometry tables from host threads and
    the host libm (a debugging aid for last-bit comparisons; the default is the
    CUDA kernel k_soltab, which test_soltab and every other test here exercise).
    Both agree with the oracle at the north-star tolerance."""
    p, f, start = drb
    from pywatershed_b200 import Engine

    e = Engine(p, start, soltab_device=False)
    st = e.soltab()
    ref = po.Oracle(p, start=start).soltabs()
    for k in st:  # the same libm on the same machine: the same bits
        np.testing.assert_array_equal(st[k], ref[k], err_msg=k)
    e.select_outputs(["swrad", "potet"], [])
    got, _ = e.run_host(f["prcp"][:30], f["tmax"][:30], f["tmin"][:30])
    r, _ = po.Oracle(p, start=start).run(f["prcp"][:30], f["tmax"][:30], f["tmin"][:30], ["swrad", "potet"], ())
    for k in ("swrad", "potet"):
        np.testing.assert_allclose(got[k], r[k], rtol=1e-9, atol=1e-10, err_msg=k)
    d = _engine(p, start).soltab()
    worst = max(float(np.abs(d[k] - st[k]).max() / max(np.abs(st[k]).max(), 1.0)) for k in st)
    print(f"k_soltab vs host libm tables: largest difference {worst:.3g} of the table's range")
    assert worst < 1e-12


@pytest.mark.parametrize("scen", ["drb", "drbx"])
def test_full_chain_parity(scen, drb, drbx):
    p, f, start = drb if scen == "drb" else drbx
    nd = 731
    e = _engine(p, start)
    got, viol = e.run_host(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd])
    _, ref, oviol = _oracle(p, f, start, nd)
    rep = _assert_chain_parity(got, ref, scen)
    # budgets: same closure verdicts as the oracle (identical for tie-free HRUs)
    if not rep["ties"]:
        np.testing.assert_array_equal(viol, oviol)
        parity.assert_segments_close(got, ref, po.SEG_VARS)
    else:
        assert np.abs(viol.astype(int) - oviol).max() <= len(rep["ties"])
        # segments no tied HRU drains to keep rtol 1e-9; the others are held to their
        # cumulative flow (tests/parity.py)
        seg = parity.assert_segments_with_ties(got, ref, po.SEG_VARS, p["tosegment"], p["hru_segment"],
                                               [t[0] for t in rep["ties"]])
        print(f"segments {scen}: {seg}")
    # One-step parity: the oracle continued from the GPU's own previous day must give the
    # GPU's day, for EVERY HRU-day -- tied HRUs included (tests/parity.py)
    ff = {k: v[:nd] for k, v in f.items()}
    osp = parity.one_step_parity(po.Oracle(p, start=start), ff, got, po.HRU_VARS, p["hru_in_to_cf"])
    print(f"one-step parity {scen}: {len(osp['mismatching'])} of {osp['hru_days']} HRU-days differ "
          f"(all on a pack at the dnearzero threshold): {osp['mismatching'][:4]}")
    assert osp["unexplained"] == [], osp["unexplained"][:5]
    assert len(osp["mismatching"]) <= 1e-4 * osp["hru_days"]
    # ... and PRMSChannel given the GPU's own lateral inflows: every segment, every day,
    # whatever the ties upstream did (routing has no transcendental: bit for bit)
    ch = po.Oracle(p, start=start).run_channel(np.asarray(got["seg_lateral_inflow"]))
    for k in po.SEG_VARS:
        np.testing.assert_array_equal(np.asarray(got[k]), ch[k], err_msg=k)
    if scen == "drb":
        assert viol.sum() == 0  # all six budgets close, as in the reference
        # against the reference's own numbers
        assert got["seg_outflow"][-1].sum() == pytest.approx(214752.70363439058, rel=1e-9)
        g = np.load(scenarios.GOLDEN / "drb_golden.npz")
        hs = g["hru_sample"]
        rep2 = parity.compare_columns({k: got[k][:, hs] for k in po.HRU_VARS},
                                      {k: g["ts_" + k] for k in po.HRU_VARS}, po.HRU_VARS,
                                      got["pkwater_equiv"][:, hs], g["ts_pkwater_equiv"])
        assert rep2["mismatches"] == []


def test_sagehen_parity(sagehen):
    """Second domain, snow-dominated (the reference's test_data/sagehen_5yr as
    sagehen_no_cascades_model.yaml runs it: NoDprst processes, no channel; packs up
    to 79 inches, two depletion curves, 5 water years): the CUDA column against
    the oracle on every HRU and against the reference's own outputs."""
    p, f, start = sagehen
    nd = f["prcp"].shape[0]
    e = _engine(p, start, dprst_flag=False, channel=False)
    e.select_outputs(e.reg.hru_vars, ())
    got, viol = e.run_host(f["prcp"], f["tmax"], f["tmin"])
    o = po.Oracle(p, start=start, dprst_flag=False)
    ref, oviol = o.run(f["prcp"], f["tmax"], f["tmin"], po.HRU_VARS, ())
    names = [k for k in po.HRU_VARS if not k.startswith("channel_")]  # PRMSChannel is not in this model
    rep = parity.compare_columns(got, ref, names, got["pkwater_equiv"], ref["pkwater_equiv"])
    assert rep["mismatches"] == [], rep["mismatches"][:5]
    assert len(rep["ties"]) <= 6, rep["ties"]
    print(f"sagehen ghost-pack ties: {len(rep['ties'])} of {got['pkwater_equiv'].shape[1]} HRUs")
    assert got["pkwater_equiv"].max() > 70.0
    if not rep["ties"]:
        np.testing.assert_array_equal(viol[:, :5], oviol[:, :5])
    assert viol[:, :5].sum() <= len(rep["ties"])
    g = np.load(scenarios.GOLDEN / "sagehen_golden.npz")
    hs = g["hru_sample"]
    gnames = [k[3:] for k in g.files if k.startswith("ts_")]
    rep2 = parity.compare_columns({k: got[k][:, hs] for k in gnames}, {k: g["ts_" + k] for k in gnames},
                                  gnames, got["pkwater_equiv"][:, hs], g["ts_pkwater_equiv"])
    assert rep2["mismatches"] == [], rep2["mismatches"][:5]


def test_hru1_41_years(hru1):
    """The reference's test_data/hru_1: ONE HRU draining to ONE segment for
    14,975 days (1979-2019) through pws_run_host in blocks of days -- the
    single-unit edge of every kernel and 41 years without drift, against the
    oracle (all days) and the reference's own outputs (golden days)."""
    p, f, start = hru1
    e = _engine(p, start)
    got, viol = e.run_host(f["prcp"], f["tmax"], f["tmin"])
    o = po.Oracle(p, start=start)
    ref, oviol = o.run(f["prcp"], f["tmax"], f["tmin"], po.HRU_VARS, po.SEG_VARS)
    rep = parity.compare_columns(got, ref, po.HRU_VARS, got["pkwater_equiv"], ref["pkwater_equiv"])
    assert rep["mismatches"] == [], rep["mismatches"][:5]
    g = np.load(scenarios.GOLDEN / "hru1_golden.npz")
    days = g["days"]
    if not rep["ties"]:
        np.testing.assert_array_equal(viol, oviol)
        parity.assert_segments_close(got, ref, po.SEG_VARS)
        for k in [k[3:] for k in g.files if k.startswith("ts_")]:
            np.testing.assert_allclose(np.asarray(got[k], dtype=np.float64)[days],
                                       np.asarray(g["ts_" + k], dtype=np.float64),
                                       rtol=1e-9, atol=1e-10, err_msg=k)
        assert viol.sum() == 0
    else:
        print("hru_1: ghost-pack tie on day", rep["ties"][0][1])
        d = rep["ties"][0][1]
        for k in po.HRU_VARS:
            np.testing.assert_allclose(np.asarray(got[k], dtype=np.float64)[:d],
                                       np.asarray(ref[k], dtype=np.float64)[:d], rtol=1e-9, atol=1e-10, err_msg=k)


@pytest.mark.parametrize("tma", [1, 0])
def test_forcing_staging_paths_agree(drb, tma):
    """k_hru_column stages the forcing rows with TMA bulk copies when they are
    16-byte aligned (pws_run_host's padded ring: always; device-resident
    forcings: even nhru) and falls back to per-thread loads otherwise; both
    must give the oracle's numbers.  2 tiles = 1,530 HRUs: an even domain whose
    last CTA is partial."""
    import torch

    p, f, start = drb
    pt, ft = scenarios.tile(p, f, 2)
    nd = 120
    ref = _oracle(pt, ft, start, nd)[1]
    e = _engine(pt, start)
    e._opt("tma", tma)
    got, _ = e.run_host(ft["prcp"][:nd], ft["tmax"][:nd], ft["tmin"][:nd])
    _assert_chain_parity(got, ref)
    # device-resident forcings (row stride = nhru, even): same bits as the host path
    e2 = _engine(pt, start)
    e2._opt("tma", tma)
    dev = torch.device("cuda:0")
    t = {k: torch.from_numpy(np.ascontiguousarray(ft[k][:nd])).to(dev) for k in ("prcp", "tmax", "tmin")}
    nf, ni = len(e2.sel_f64), len(e2.sel_i32)
    of = torch.empty((nd, nf, e2.nhru), dtype=torch.float64, device=dev)
    oi = torch.empty((nd, ni, e2.nhru), dtype=torch.int32, device=dev)
    osg = torch.empty((nd, 5, e2.nseg), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    e2.run_device(nd, t["prcp"].data_ptr(), t["tmax"].data_ptr(), t["tmin"].data_ptr(),
                  of.data_ptr(), oi.data_ptr(), osg.data_ptr())
    of = of.cpu().numpy()
    for k, v in enumerate(e2.sel_f64):
        np.testing.assert_array_equal(of[:, k], got[v], err_msg=v)


def test_pinned_and_pageable_host_buffers_agree(drb):
    """pws_run_host uses pinned caller buffers directly and stages pageable
    ones (numpy arrays) through its own pinned slots: same bits, any block
    length (3+ blocks so that both stage slots are reused)."""
    import torch

    p, f, start = drb
    nd = 75
    e = _engine(p, start)
    e.set_block_days(20)
    a, va = e.run_host(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd])
    e2 = _engine(p, start)
    e2.set_block_days(20)
    pin = {k: torch.from_numpy(np.ascontiguousarray(f[k][:nd])).pin_memory().numpy()
           for k in ("prcp", "tmax", "tmin")}
    b, vb = e2.run_host(pin["prcp"], pin["tmax"], pin["tmin"], pinned_outputs=True)
    for k in po.HRU_VARS + po.SEG_VARS:
        np.testing.assert_array_equal(np.asarray(a[k]), np.asarray(b[k]), err_msg=k)
    np.testing.assert_array_equal(va, vb)
    # inputs and outputs are staged independently: the two mixed cases
    e3 = _engine(p, start)
    e3.set_block_days(20)
    c, vc = e3.run_host(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd], pinned_outputs=True)  # what Model does
    e4 = _engine(p, start)
    e4.set_block_days(20)
    d, vd = e4.run_host(pin["prcp"], pin["tmax"], pin["tmin"])
    for k in po.HRU_VARS + po.SEG_VARS:
        np.testing.assert_array_equal(np.asarray(a[k]), np.asarray(c[k]), err_msg=k)
        np.testing.assert_array_equal(np.asarray(a[k]), np.asarray(d[k]), err_msg=k)
    np.testing.assert_array_equal(va, vc)
    np.testing.assert_array_equal(va, vd)


def test_time_blocking_is_transparent(drb):
    """Block length must not change a single bit (state round-trips through
    HBM between blocks)."""
    p, f, start = drb
    nd = 90
    runs = []
    for T in (1, 7, 90):
        e = _engine(p, start)
        e.set_block_days(T)
        got, _ = e.run_host(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd])
        runs.append(got)
    for k in po.HRU_VARS + po.SEG_VARS:
        for r in runs[1:]:
            np.testing.assert_array_equal(np.asarray(r[k]), np.asarray(runs[0][k]), err_msg=k)


def test_output_selection_and_device_resident_run(drb):
    import torch

    p, f, start = drb
    nd = 40
    e = _engine(p, start)
    full, _ = e.run_host(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd])
    e2 = _engine(p, start)
    sel_h = ["pkwater_equiv", "newsnow", "soil_moist", "iasw", "sroff_vol"]
    sel_s = ["seg_outflow"]
    e2.select_outputs(sel_h, sel_s)
    dev = torch.device("cuda:0")
    t = {k: torch.from_numpy(np.ascontiguousarray(f[k][:nd])).to(dev) for k in ("prcp", "tmax", "tmin")}
    of = torch.empty((nd, 3, e2.nhru), dtype=torch.float64, device=dev)
    oi = torch.empty((nd, 2, e2.nhru), dtype=torch.int32, device=dev)
    osg = torch.empty((nd, 1, e2.nseg), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    e2.run_device(nd, t["prcp"].data_ptr(), t["tmax"].data_ptr(), t["tmin"].data_ptr(),
                  of.data_ptr(), oi.data_ptr(), osg.data_ptr())
    of, oi, osg = of.cpu().numpy(), oi.cpu().numpy(), osg.cpu().numpy()
    np.testing.assert_array_equal(of[:, 0], full["pkwater_equiv"])
    np.testing.assert_array_equal(of[:, 1], full["soil_moist"])
    np.testing.assert_array_equal(of[:, 2], full["sroff_vol"])
    np.testing.assert_array_equal(oi[:, 0], full["newsnow"])
    np.testing.assert_array_equal(oi[:, 1].astype(bool), full["iasw"])
    np.testing.assert_array_equal(osg[:, 0], full["seg_outflow"])
    assert e2.launch_count() > 0


def test_state_round_trip_is_a_restart(drb):
    """get_state after N days -> set_state into a fresh engine -> continue:
    identical to the uninterrupted run for the HRU column (restart, which the
    reference declares but does not implement, prms_snow.py:269-293)."""
    p, f, start = drb
    e = _engine(p, start, channel=False)
    e.select_outputs(e.reg.hru_vars, ())
    a, _ = e.run_host(f["prcp"][:60], f["tmax"][:60], f["tmin"][:60])
    e1 = _engine(p, start, channel=False)
    e1.select_outputs(e1.reg.hru_vars, ())
    e1.run_host(f["prcp"][:30], f["tmax"][:30], f["tmin"][:30])
    st = {s: e1.get_state(s) for s in e1.reg.states}
    e2 = _engine(p, start, channel=False)
    e2.select_outputs(e2.reg.hru_vars, ())
    for s, v in st.items():
        e2.set_state(s, v)
    e2.day = 30
    b, _ = e2.run_host(f["prcp"][30:60], f["tmax"][30:60], f["tmin"][30:60])
    for k in po.HRU_VARS:
        np.testing.assert_array_equal(np.asarray(b[k]), np.asarray(a[k])[30:], err_msg=k)


def test_channel_alone_bit_exact(drb):
    """PRMSChannel driven by given lateral inflows: integer-free but the
    routing recurrence has no transcendental, so the CUDA path must equal the
    oracle bit for bit."""
    import torch

    p, f, start = drb
    nd = 120
    rng = np.random.default_rng(11)
    nseg = p["tosegment"].shape[0]
    lat = rng.lognormal(0.0, 1.0, size=(nd, nseg))
    lat[:, ::7] = 0.0
    ref = po.Oracle(p, start=start).run_channel(lat)
    e = _engine(p, start)
    dev = torch.device("cuda:0")
    dl = torch.from_numpy(lat).to(dev)
    out = torch.empty((nd, 5, nseg), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    e.run_channel_device(nd, dl.data_ptr(), out.data_ptr())
    out = out.cpu().numpy()
    for k, v in enumerate(po.SEG_VARS):
        np.testing.assert_array_equal(out[:, k], ref[v], err_msg=v)


@pytest.mark.parametrize("chunk", [512, 100, 7, 1])
def test_channel_chunking_is_transparent(drb, chunk):
    """The routing kernel cuts the segment forest into chunks of <= `chunk`
    segments (one CTA each); edges inside a chunk hand over through shared
    memory, edges across chunks through global memory.  Any chunk size must
    give the oracle's bits (DRB's 434-segment tree is cut for chunk < 434)."""
    import torch

    p, f, start = drb
    nd = 45
    rng = np.random.default_rng(5)
    nseg = p["tosegment"].shape[0]
    lat = rng.lognormal(0.0, 1.0, size=(nd, nseg))
    ref = po.Oracle(p, start=start).run_channel(lat)
    e = _engine(p, start, channel_chunk=chunk)
    dev = torch.device("cuda:0")
    dl = torch.from_numpy(lat).to(dev)
    out = torch.empty((nd, 5, nseg), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    # two launches: state carried across blocks
    e.run_channel_device(20, dl.data_ptr(), out.data_ptr())
    e.run_channel_device(nd - 20, dl[20:].data_ptr(), out[20:].data_ptr())
    out = out.cpu().numpy()
    for k, v in enumerate(po.SEG_VARS):
        np.testing.assert_array_equal(out[:, k], ref[v], err_msg=v)


def test_deep_chained_network(drb):
    """BASELINE.json configs[4] in small: DRB tiles chained outlet -> headwater
    so the network is several hundred levels deep and crosses chunks."""
    import torch

    p, f, start = drb
    ntile = 4
    pt, _ = scenarios.tile(p, {k: v[:2] for k, v in f.items()}, ntile)
    nseg0 = p["tosegment"].shape[0]
    tos = pt["tosegment"].copy()
    main_outlet = int(np.argmax(np.bincount(_basin_of(p["tosegment"]))))  # outlet of the big tree
    head = int(np.argmax(np.where(_basin_of(p["tosegment"]) == main_outlet,
                                  _dist_to_outlet(p["tosegment"]), -1)))  # the farthest headwater
    for t in range(ntile - 1):
        tos[t * nseg0 + main_outlet] = (t + 1) * nseg0 + head + 1
    pt["tosegment"] = tos
    nd, nseg = 30, tos.shape[0]
    rng = np.random.default_rng(11)
    lat = rng.lognormal(0.0, 1.0, size=(nd, nseg))
    ref = po.Oracle(pt, start=start).run_channel(lat)
    e = _engine(pt, start)
    dev = torch.device("cuda:0")
    dl = torch.from_numpy(lat).to(dev)
    out = torch.empty((nd, 5, nseg), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    e.run_channel_device(nd, dl.data_ptr(), out.data_ptr())
    out = out.cpu().numpy()
    for k, v in enumerate(po.SEG_VARS):
        np.testing.assert_array_equal(out[:, k], ref[v], err_msg=v)


@pytest.mark.parametrize("chunk", [512, 100])
def test_overlapped_routing_launches_equal_one_launch(drb, chunk):
    """Deep networks pay their systolic fill per launch, so consecutive routing
    launches may overlap on two streams (option route_overlap; a chunk of launch b+1
    continues from the state the same chunk of launch b stored, d_chunk_done): the
    chained network in launches of 1..7 days, issued without a sync in between,
    equals the oracle bit for bit, with whole trees per chunk and with cut trees."""
    import torch

    p, f, start = drb
    ntile = 4
    pt, _ = scenarios.tile(p, {k: v[:2] for k, v in f.items()}, ntile)
    nseg0 = p["tosegment"].shape[0]
    tos = pt["tosegment"].copy()
    main_outlet = int(np.argmax(np.bincount(_basin_of(p["tosegment"]))))
    head = int(np.argmax(np.where(_basin_of(p["tosegment"]) == main_outlet,
                                  _dist_to_outlet(p["tosegment"]), -1)))
    for t in range(ntile - 1):
        tos[t * nseg0 + main_outlet] = (t + 1) * nseg0 + head + 1
    pt["tosegment"] = tos
    nd, nseg = 45, tos.shape[0]
    rng = np.random.default_rng(12)
    lat = rng.lognormal(0.0, 1.0, size=(nd, nseg))
    ref = po.Oracle(pt, start=start).run_channel(lat)
    dev = torch.device("cuda:0")
    dl = torch.from_numpy(lat).to(dev)
    for overlap in (1, 0):
        e = _engine(pt, start, channel_chunk=chunk)
        e._opt("route_overlap", overlap)
        out = torch.zeros((nd, 5, nseg), dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        d, k = 0, 0
        while d < nd:
            n = min((1, 7, 3, 5, 2)[k % 5], nd - d)
            e.run_channel_device(n, dl.data_ptr() + d * nseg * 8, out.data_ptr() + d * 5 * nseg * 8,
                                 sync=False)
            d += n
            k += 1
        e.sync()
        got = out.cpu().numpy()
        e.close()
        for j, v in enumerate(po.SEG_VARS):
            np.testing.assert_array_equal(got[:, j], ref[v], err_msg=f"{v} overlap={overlap}")


@pytest.mark.parametrize("seed,chunk", [(1, 512), (2, 96), (3, 17)])
def test_random_forests_route_bit_exact(drb, seed, chunk):
    """Random segment forests (fan-in up to 6, chains, isolated outlets, every
    Muskingum time-step class incl. pass-through and lake segments) cut into
    chunks of different sizes: routing must equal the oracle bit for bit."""
    import torch

    p, f, start = drb
    rng = np.random.default_rng(seed)
    nseg, nd = 1500, 40
    tos = np.zeros(nseg, dtype=np.int64)
    for i in range(nseg - 1):
        u = rng.random()
        if u < 0.02:
            tos[i] = 0                                    # an outlet
        elif u < 0.5:
            tos[i] = i + 2                                # chain to the next segment
        else:
            tos[i] = int(rng.integers(i + 2, min(nseg, i + 40) + 1))  # 1-based, downstream of i
    hub = nseg - 5
    tos[hub - 7:hub - 1] = hub + 1                        # fan-in of 6
    q = dict(p)
    q["tosegment"] = tos
    q["seg_length"] = rng.lognormal(9.0, 1.5, nseg)       # K from < 1 h to the 24 h cap
    q["seg_slope"] = rng.uniform(1e-4, 0.05, nseg)
    q["seg_depth"] = rng.uniform(0.3, 3.0, nseg)
    q["mann_n"] = rng.uniform(0.02, 0.08, nseg)
    q["x_coef"] = rng.uniform(0.0, 0.5, nseg)
    q["segment_type"] = np.where(rng.random(nseg) < 0.03, 2, 0)
    q["segment_flow_init"] = np.where(rng.random(nseg) < 0.3, rng.uniform(0, 5, nseg), 0.0)
    q["hru_segment"] = rng.integers(0, nseg + 1, p["hru_area"].shape[0])
    q["nhm_seg"] = np.arange(1, nseg + 1)
    lat = rng.lognormal(0.0, 1.0, size=(nd, nseg))
    lat[:, rng.random(nseg) < 0.2] = 0.0
    ref = po.Oracle(q, start=start).run_channel(lat)
    e = _engine(q, start, channel_chunk=chunk)
    dev = torch.device("cuda:0")
    dl = torch.from_numpy(lat).to(dev)
    out = torch.empty((nd, 5, nseg), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    e.run_channel_device(nd, dl.data_ptr(), out.data_ptr())
    out = out.cpu().numpy()
    for k, v in enumerate(po.SEG_VARS):
        np.testing.assert_array_equal(out[:, k], ref[v], err_msg=v)


def _dist_to_outlet(tosegment):
    tos = np.asarray(tosegment) - 1
    dist = np.zeros(len(tos), dtype=int)
    cur = np.arange(len(tos))
    for _ in range(len(tos)):
        nxt = tos[cur]
        step = nxt >= 0
        if not step.any():
            break
        dist += step
        cur = np.where(step, nxt, cur)
    return dist


def _basin_of(tosegment):
    """index of the outlet each segment drains to"""
    tos = np.asarray(tosegment) - 1
    out = np.arange(len(tos))
    for _ in range(len(tos)):
        nxt = np.where(tos[out] >= 0, tos[out], out)
        if (nxt == out).all():
            break
        out = nxt
    return out


def test_tiled_domain_and_ragged_sizes(drb):
    """3 DRB tiles in different climates, truncated to a ragged 1,531 HRUs
    (not a multiple of the warp or block size); every HRU within tolerance
    of the oracle, ties counted."""
    p, f, start = drb
    pt, ft = scenarios.tile(p, f, 3, nhru_keep=1531)
    # HRUs of the cut tile still point at segments that exist (3 full tiles of segments)
    nd = 400
    e = _engine(pt, start)
    got, viol = e.run_host(ft["prcp"][:nd], ft["tmax"][:nd], ft["tmin"][:nd])
    _, ref, oviol = _oracle(pt, ft, start, nd)
    rep = _assert_chain_parity(got, ref)
    if not rep["ties"]:
        parity.assert_segments_close(got, ref, po.SEG_VARS)


def _subset(p, f, hru_idx):
    """The HRU columns `hru_idx` of a domain, without segments."""
    nhru = p["hru_area"].shape[0]
    nseg = p["tosegment"].shape[0]
    q = {}
    for k, v in p.items():
        v = np.asarray(v)
        if k in ("tosegment", "tosegment_nhm", "nhm_seg") or (v.ndim >= 1 and v.shape[-1] == nseg and k in (
                "mann_n", "seg_depth", "seg_length", "seg_slope", "x_coef", "segment_flow_init", "segment_type")):
            q[k] = v
        elif v.ndim >= 1 and v.shape[-1] == nhru and k != "snarea_curve":
            q[k] = np.ascontiguousarray(v[..., hru_idx])
        else:
            q[k] = v
    return q, {k: np.ascontiguousarray(v[:, hru_idx]) for k, v in f.items()}


@pytest.mark.parametrize("n", [1, 31, 257])
def test_tiny_and_odd_domains(drbx, n):
    """hru_1-sized and warp / CTA-straddling domains (the reference's smallest
    test domain has one HRU): columns only, every variable against the oracle."""
    p, f, start = drbx
    idx = np.arange(n) * 2 + 2
    q, g = _subset(p, f, idx)
    q["hru_segment"] = np.zeros_like(q["hru_segment"])  # columns only: no HRU drains to a segment
    nd = 200
    e = _engine(q, start, channel=False)
    e.select_outputs(e.reg.hru_vars, ())
    got, viol = e.run_host(g["prcp"][:nd], g["tmax"][:nd], g["tmin"][:nd])
    o = po.Oracle(q, start=start)
    ref, oviol = o.run(g["prcp"][:nd], g["tmax"][:nd], g["tmin"][:nd], po.HRU_VARS, ())
    rep = parity.compare_columns(got, ref, po.HRU_VARS, got["pkwater_equiv"], ref["pkwater_equiv"])
    assert rep["mismatches"] == [], rep["mismatches"][:3]
    assert len(rep["ties"]) <= max(1, n // 30)
    # zero days is a no-op
    got0, _ = e.run_host(g["prcp"][:0], g["tmax"][:0], g["tmin"][:0])
    assert got0["pkwater_equiv"].shape == (0, n)


def test_ensemble_members_are_independent(drb):
    """BASELINE.json configs[3] in small: a parameter ensemble laid out
    member-major in one engine (perturbed soil / snow parameters per member);
    every member must equal the same member run alone, bit for bit."""
    p, f, start = drb
    nm, nd = 3, 150
    nhru = p["hru_area"].shape[0]
    pe, fe = scenarios.tile(p, {k: v[:nd] for k, v in f.items()}, nm, temp_shift=False)
    rng = np.random.default_rng(2026)
    scale = {k: rng.uniform(0.8, 1.2, nm) for k in
             ("soil_moist_max", "soil2gw_max", "ssr2gw_rate", "slowcoef_lin", "smidx_coef", "carea_max",
              "gwflow_coef", "freeh2o_cap", "emis_noppt", "rad_trncf", "potet_sublim")}
    for k, sc in scale.items():
        v = np.array(pe[k], dtype=np.float64, copy=True)
        for m in range(nm):
            v[..., m * nhru:(m + 1) * nhru] *= sc[m]
        if k in ("carea_max", "emis_noppt", "rad_trncf"):
            v = np.minimum(v, 1.0)
        pe[k] = v
    e = _engine(pe, start)
    allm, _ = e.run_host(fe["prcp"], fe["tmax"], fe["tmin"])
    nseg = p["tosegment"].shape[0]
    for m in range(nm):
        pm = dict(p)
        for k in scale:
            pm[k] = pe[k][..., m * nhru:(m + 1) * nhru]
        em = _engine(pm, start)
        one, _ = em.run_host(f["prcp"][:nd], f["tmax"][:nd], f["tmin"][:nd])
        for k in po.HRU_VARS:
            np.testing.assert_array_equal(np.asarray(allm[k])[:, m * nhru:(m + 1) * nhru],
                                          np.asarray(one[k]), err_msg=f"{k} member {m}")
        for k in po.SEG_VARS:
            np.testing.assert_array_equal(allm[k][:, m * nseg:(m + 1) * nseg], one[k], err_msg=k)


def test_missing_parameter_and_bad_option_raise(drb):
    from pywatershed_b200 import Engine

    p, f, start = drb
    q = dict(p)
    del q["carea_max"]
    with pytest.raises(KeyError):
        Engine(q, start)
    q = dict(p)
    q["snowpack_init"] = np.full_like(p["snowpack_init"], 0.5)
    with pytest.raises(ValueError, match="snowpack_init"):
        Engine(q, start)
    q = dict(p)
    q["albset_rna"] = np.full(p["hru_area"].shape[0], 0.8)
    with pytest.raises(ValueError, match="scalar"):
        Engine(q, start)


@pytest.mark.parametrize("ntile,nhru_keep,nd,outputs", [(1, 257, 7, "full"), (2, None, 5, "full"),
                                                       (1, 31, 9, "select"), (2, 1000, 3, "select")])
def test_output_writes_stay_inside_their_buffers(drb, ntile, nhru_keep, nd, outputs):
    """Guard bands around every device buffer the library writes into (HRU f64 /
    i32 outputs, segment outputs) and reads from (forcings): after a run the
    bands still hold their sentinel and the interior equals a run into exact-size
    buffers.  Odd sizes so that the last CTA, the last warp and the TMA fallback
    are all exercised."""
    import torch

    p, f, start = drb
    pt, ft = scenarios.tile(p, {k: v[:nd] for k, v in f.items()}, ntile, nhru_keep=nhru_keep)
    dev = torch.device("cuda:0")
    e = _engine(pt, start)
    if outputs == "select":
        e.select_outputs(["pkwater_equiv", "soil_moist", "transp_on", "sroff_vol", "pptmix"], ["seg_outflow"])
    nh, ns = e.nhru, e.nseg
    nf, ni, nsv = len(e.sel_f64), len(e.sel_i32), len(e.sel_seg)
    G = 4099  # guard elements on both sides (odd: the interior is not 16-byte aligned for i32)

    def guarded(n, dtype, fill):
        t = torch.full((n + 2 * G,), fill, dtype=dtype, device=dev)
        return t, t[G:G + n]

    SENT_F, SENT_I = -7.25e300, -1234567
    of_all, of = guarded(nd * nf * nh, torch.float64, SENT_F)
    oi_all, oi = guarded(nd * max(ni, 1) * nh, torch.int32, SENT_I)
    os_all, osg = guarded(nd * nsv * ns, torch.float64, SENT_F)
    forc = {}
    for k in ("prcp", "tmax", "tmin"):
        full, inner = guarded(nd * nh, torch.float64, float("nan"))  # a read outside would poison the results
        inner.copy_(torch.from_numpy(np.ascontiguousarray(ft[k])).reshape(-1))
        forc[k] = (full, inner)
    viol = np.zeros((nd, 6), dtype=np.int32)
    torch.cuda.synchronize()
    e.run_device(nd, forc["prcp"][1].data_ptr(), forc["tmax"][1].data_ptr(), forc["tmin"][1].data_ptr(),
                 of.data_ptr(), oi.data_ptr(), osg.data_ptr(), viol=viol)
    for t_all, sent in ((of_all, SENT_F), (oi_all, SENT_I), (os_all, SENT_F)):
        assert bool((t_all[:G] == sent).all()) and bool((t_all[-G:] == sent).all())
    assert not bool((of == SENT_F).any()) and not bool((osg == SENT_F).any())
    if ni:
        assert not bool((oi[:nd * ni * nh] == SENT_I).any())
    # the same run through the host entry point into exact-size buffers
    e2 = _engine(pt, start)
    if outputs == "select":
        e2.select_outputs(["pkwater_equiv", "soil_moist", "transp_on", "sroff_vol", "pptmix"], ["seg_outflow"])
    ref, rviol = e2.run_host(ft["prcp"], ft["tmax"], ft["tmin"])
    got_f = of.reshape(nd, nf, nh).cpu().numpy()
    for j, name in enumerate(e.sel_f64):
        np.testing.assert_array_equal(got_f[:, j], ref[name], err_msg=name)
    got_i = oi[:nd * ni * nh].reshape(nd, ni, nh).cpu().numpy() if ni else None
    for j, name in enumerate(e.sel_i32):
        np.testing.assert_array_equal(got_i[:, j], np.asarray(ref[name], dtype=np.int32), err_msg=name)
    got_s = osg.reshape(nd, nsv, ns).cpu().numpy()
    for j, name in enumerate(e.sel_seg):
        np.testing.assert_array_equal(got_s[:, j], ref[name], err_msg=name)
    np.testing.assert_array_equal(viol, rviol)


@pytest.mark.parametrize("hours", [2, 4, 12])
def test_route_hours_are_equivalent(drb, hours):
    """k_route_chunks advances 12 hours per systolic step on shallow segment
    graphs, 4 on deep ones and 2 on very deep ones (option route_hours); the
    schedule must not change a bit of the routing."""
    import torch

    p, f, start = drb
    nd, nseg = 40, p["tosegment"].shape[0]
    rng = np.random.default_rng(12)
    lat = rng.lognormal(0.0, 1.0, size=(nd, nseg))
    ref = po.Oracle(p, start=start).run_channel(lat)
    from pywatershed_b200 import Engine

    e = Engine(p, start, channel_chunk=96)  # several chunks per tree: cross-chunk edges too
    e._opt("route_hours", hours)
    e._check(e.L.pws_init(e.h))
    e.select_outputs(e.reg.hru_vars, e.reg.seg_vars)
    dev = torch.device("cuda:0")
    dl = torch.from_numpy(lat).to(dev)
    out = torch.empty((nd, 5, nseg), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    e.run_channel_device(nd, dl.data_ptr(), out.data_ptr())
    out = out.cpu().numpy()
    for k, v in enumerate(po.SEG_VARS):
        np.testing.assert_array_equal(out[:, k], ref[v], err_msg=v)


@pytest.mark.parametrize("ntile,column_hrus", [(1, 0), (2, 256), (3, 128)])
def test_tiled_result_layout(drb, ntile, column_hrus):
    """pws_run_device with `tiled_outputs` writes [day][tile][variable][H]
    (include/pws_b200.h: pws_output_tile): against the oracle, the same bits
    as the row layout, for both CTA sizes, a partial last tile whose padding
    stays untouched, two consecutive blocks, and refused by pws_run_host and
    for a partial selection."""
    import torch

    p, f, start = drb
    pt, ft = scenarios.tile(p, f, ntile) if ntile > 1 else (p, f)
    nd = 60
    ref = _oracle(pt, ft, start, nd)[1]
    rows, _ = _engine(pt, start).run_host(ft["prcp"][:nd], ft["tmax"][:nd], ft["tmin"][:nd])
    e = _engine(pt, start)
    if column_hrus:
        e._opt("column_hrus", column_hrus)
    e.set_tiled_outputs(True)
    H, padded = e.output_tile()
    assert H == (column_hrus or 128) and padded % H == 0 and 0 <= padded - e.nhru < H
    dev = torch.device("cuda:0")
    t = {k: torch.from_numpy(np.ascontiguousarray(ft[k][:nd])).to(dev) for k in ("prcp", "tmax", "tmin")}
    nf, ni = len(e.sel_f64), len(e.sel_i32)
    SENT_F, SENT_I = -7.25e300, -1234567
    of = torch.full((nd, padded // H, nf, H), SENT_F, dtype=torch.float64, device=dev)
    oi = torch.full((nd, padded // H, ni, H), SENT_I, dtype=torch.int32, device=dev)
    osg = torch.empty((nd, 5, e.nseg), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    n1 = 23  # two blocks: the second starts at a day offset of the same buffers
    for d0, n in ((0, n1), (n1, nd - n1)):
        o = d0 * e.nhru * 8
        e.run_device(n, t["prcp"].data_ptr() + o, t["tmax"].data_ptr() + o, t["tmin"].data_ptr() + o,
                     of[d0:].data_ptr(), oi[d0:].data_ptr(), osg[d0:].data_ptr())
    of, oi, osg = of.cpu().numpy(), oi.cpu().numpy(), osg.cpu().numpy()
    got = {}
    for v in e.sel_f64 + e.sel_i32:
        a = e.tiled_view(of if v in e.sel_f64 else oi, v)
        got[v] = a.astype(np.bool_) if e.reg.dtype(v) is np.bool_ else a
        np.testing.assert_array_equal(got[v], np.asarray(rows[v]), err_msg=v)
    for k, v in enumerate(e.sel_seg):
        got[v] = osg[:, k, :]
        np.testing.assert_array_equal(got[v], np.asarray(rows[v]), err_msg=v)
    _assert_chain_parity(got, ref)
    # the padding of the last tile is never written
    pad = padded - e.nhru
    if pad:
        assert (of[:, -1, :, H - pad:] == SENT_F).all() and (oi[:, -1, :, H - pad:] == SENT_I).all()
    e.select_outputs(["pkwater_equiv"], [])
    one = torch.empty((padded,), dtype=torch.float64, device=dev)
    with pytest.raises(ValueError, match="tiled_outputs"):
        e.run_device(1, t["prcp"].data_ptr(), t["tmax"].data_ptr(), t["tmin"].data_ptr(), one.data_ptr(), 0, 0)


def test_float32_forcings_are_widened_on_the_device(drb):
    """pws_run_host_f32: forcings in the storage type of the reference's
    forcing files (float32 netCDF) are uploaded as they are and widened on the
    device.  Same bits as pws_run_host on the widened arrays -- pageable and
    page-locked inputs, several blocks so that both ring slots are reused, an
    odd number of HRUs -- and the oracle's numbers."""
    import torch

    p, f, start = drb
    nd = 75
    f32 = {k: np.ascontiguousarray(f[k][:nd], dtype=np.float32) for k in ("prcp", "tmax", "tmin")}
    f64 = {k: v.astype(np.float64) for k, v in f32.items()}
    assert p["hru_area"].shape[0] % 2 == 1
    ref = _oracle(p, f64, start, nd)[1]
    e = _engine(p, start)
    e.set_block_days(20)
    l0 = e.launch_count()
    a, va = e.run_host(f64["prcp"], f64["tmax"], f64["tmin"])
    l1 = e.launch_count()
    _assert_chain_parity(a, ref)
    e.reset()
    l2 = e.launch_count()
    b, vb = e.run_host(f32["prcp"], f32["tmax"], f32["tmin"])
    assert e.launch_count() - l2 == (l1 - l0) + 4  # one widening kernel per block of 20 days
    e2 = _engine(p, start)
    e2.set_block_days(20)
    pin = {k: torch.from_numpy(v).pin_memory().numpy() for k, v in f32.items()}
    c, vc = e2.run_host(pin["prcp"], pin["tmax"], pin["tmin"], pinned_outputs=True)
    for k in po.HRU_VARS + po.SEG_VARS:
        np.testing.assert_array_equal(np.asarray(b[k]), np.asarray(a[k]), err_msg=k)
        np.testing.assert_array_equal(np.asarray(c[k]), np.asarray(a[k]), err_msg=k)
    np.testing.assert_array_equal(va, vb)
    np.testing.assert_array_equal(va, vc)
    # one float64 array among them: everything goes up as float64
    e3 = _engine(p, start)
    d, _ = e3.run_host(f32["prcp"], f64["tmax"], f32["tmin"])
    for k in ("pkwater_equiv", "soil_moist", "seg_outflow"):
        np.testing.assert_array_equal(np.asarray(d[k]), np.asarray(a[k]), err_msg=k)
