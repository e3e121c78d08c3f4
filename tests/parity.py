"""Parity bookkeeping shared by the oracle-vs-golden and CUDA-vs-oracle tests.

Tolerance (BASELINE.json north_star): every state/flux variable within
rtol 1e-9 / atol 1e-10; integer and flag outputs bit-exact, "with any
disagreements counted and traced to threshold ties".

The one tie source in this chain is PRMSSnow's *ghost pack*: when a pack melts
out, `pkwater_equiv` can be left at rounding-noise level (1e-16 .. 1e-8 inches,
the residue of a cancellation that has already lost most of its digits) and the day on which it finally drops below
`dnearzero = 2.23e-16` (prms_snow.py:797) depends on the last bits of the
transcendental functions.  Rain on such a ghost pack is routed as snowmelt
through check_capacity instead of perv_comp (prms_runoff.py:905-935), so a
1-ulp difference becomes a visible one for that HRU from that day on.
`compare_columns` therefore works per HRU: it finds the first day on which an
HRU leaves tolerance, requires a ghost pack (0 < pkwater_equiv < GHOST on
either side within the two days before) to call it a tie, and reports ties and
genuine mismatches separately.
"""
from __future__ import annotations

import numpy as np

RTOL = 1e-9
ATOL = 1e-10
GHOST = 1e-6  # = PRMS's own `nearzero` (inches of water)

# differences of two large numbers: compared with an absolute tolerance scaled
# by the magnitude of the operands (a flow of ~1e4 cfs x 86400 s)
DIFFERENCE_VARS = {"seg_stor_change": ("seg_outflow", 86400.0)}


def compare_columns(got: dict, ref: dict, names, pk_got=None, pk_ref=None,
                    rtol=RTOL, atol=ATOL):
    """got/ref: var -> [ndays, n].  Returns dict(first_bad=[n] day index or -1,
    ties=[...unit idx], mismatches=[(unit, day, var, got, ref)])."""
    names = list(names)
    nd, n = np.asarray(ref[names[0]]).shape
    first = np.full(n, nd, dtype=np.int64)
    which = [None] * n
    for k in names:
        r = np.asarray(ref[k], dtype=np.float64)
        g = np.asarray(got[k], dtype=np.float64)
        ok = np.isclose(g, r, rtol=rtol, atol=atol, equal_nan=True)
        if ok.all():
            continue
        bad_day = np.where((~ok).any(axis=0), (~ok).argmax(axis=0), nd)
        for u in np.where(bad_day < first)[0]:
            first[u] = bad_day[u]
            which[u] = k
    ties, mism = [], []
    for u in np.where(first < nd)[0]:
        d = int(first[u])
        k = which[u]
        is_tie = False
        if pk_got is not None:
            lo = max(d - 2, 0)
            w = np.concatenate([np.asarray(pk_got)[lo:d + 1, u], np.asarray(pk_ref)[lo:d + 1, u]])
            is_tie = bool(((w > 0) & (w < GHOST)).any())
        rec = (int(u), d, k, float(np.asarray(got[k])[d, u]), float(np.asarray(ref[k])[d, u]))
        (ties if is_tie else mism).append(rec)
    return {"first_bad": np.where(first < nd, first, -1), "ties": ties, "mismatches": mism}


def assert_segments_close(got: dict, ref: dict, names, rtol=RTOL, atol=ATOL):
    for k in names:
        g = np.asarray(got[k], dtype=np.float64)
        r = np.asarray(ref[k], dtype=np.float64)
        a = atol
        if k in DIFFERENCE_VARS:
            base, scale = DIFFERENCE_VARS[k]
            a = atol + rtol * scale * np.abs(np.asarray(ref[base], dtype=np.float64))
        bad = ~(np.abs(g - r) <= a + rtol * np.abs(r))
        bad &= ~(np.isnan(g) & np.isnan(r))
        assert not bad.any(), (
            f"{k}: {bad.sum()} values out of tolerance, first at {np.argwhere(bad)[0]}: "
            f"got {g[tuple(np.argwhere(bad)[0])]!r} ref {r[tuple(np.argwhere(bad)[0])]!r}")


# ---------------------------------------------------------------- after a tie
# A tied HRU is not dropped from the comparison: from its tie day on it is held
# to bounds that a ghost-pack tie can explain and a physics bug cannot.
#   * The two runs differ by how one day's rain on a ghost pack is split between
#     infiltration and surface runoff (prms_runoff.py:905-935), and by the
#     memory of a pack that holds 1e-16 inches in one run and nothing in the
#     other: `ai`, `pst`, `scrv` (the depletion curve's history) decide how much
#     of the HRU the NEXT snowfall covers, so the two runs -- both legitimate
#     trajectories of the reference's algorithm -- can stay apart for a season
#     (measured on drbx: up to 10 % of one HRU's surface runoff over 2 years).
#     The audit therefore bounds what a tie cannot do: the water an HRU gives
#     off (ET + flow to the stream) summed over the run agrees within
#     POST_TIE_CUM_RTOL of the precipitation-scale total, storages stay within
#     POST_TIE_STORE_ATOL inches, budgets keep closing (the caller checks the
#     verdict counts).  The tight statement about tied HRUs is one_step_parity.
#   * Segments that no tied HRU drains to (directly or through upstream
#     segments) keep the full tolerance; the others are compared on their
#     cumulative flow.
POST_TIE_CUM_RTOL = 5e-2
POST_TIE_STORE_ATOL = 2.0   # inches; soil / groundwater / pack storages
CUM_FLUXES = ("hru_actet", "sroff_vol", "ssres_flow_vol", "gwres_flow_vol")
STORAGES = ("soil_moist", "gwres_stor", "ssres_stor", "pkwater_equiv")


def downstream_mask(tosegment, hru_segment, hrus):
    """Segments that receive water from any of `hrus` (0-based), directly or from upstream."""
    tos = np.asarray(tosegment, dtype=np.int64) - 1
    hs = np.asarray(hru_segment, dtype=np.int64) - 1
    mask = np.zeros(tos.shape[0], dtype=bool)
    for h in hrus:
        j = hs[h]
        while j >= 0 and not mask[j]:
            mask[j] = True
            j = tos[j]
    return mask


def audit_ties(got: dict, ref: dict, rep: dict, pk_got, pk_ref):
    """Per tied HRU (see above).  Returns a summary dict; raises AssertionError
    when a tied HRU leaves the post-tie bounds."""
    out = {"n_ties": len(rep["ties"]), "existence": 0, "residue": 0, "max_cum_rel": 0.0,
           "max_store_abs": 0.0, "worst": None}
    pk_got, pk_ref = np.asarray(pk_got), np.asarray(pk_ref)
    for (u, d, k, gv, rv) in rep["ties"]:
        lo = max(d - 3, 0)
        differ = ((pk_got[lo:d + 1, u] > 0) != (pk_ref[lo:d + 1, u] > 0)).any()
        out["existence" if differ else "residue"] += 1
        tot_g = tot_r = 0.0
        for name in ("hru_actet", "sroff", "ssres_flow", "gwres_flow"):  # inches, comparable
            if name in got and name in ref:
                tot_g += float(np.asarray(got[name], dtype=np.float64)[:, u].sum())
                tot_r += float(np.asarray(ref[name], dtype=np.float64)[:, u].sum())
        rel = abs(tot_g - tot_r) / max(abs(tot_r), 1e-6)
        if rel > out["max_cum_rel"]:
            out["max_cum_rel"], out["worst"] = rel, (u, d, tot_g, tot_r)
        assert rel <= POST_TIE_CUM_RTOL, (
            f"HRU {u} (tie on day {d} in {k}): ET + flows over the run {tot_g!r} vs {tot_r!r}")
        for name in STORAGES:
            if name not in got or name not in ref:
                continue
            dev = float(np.abs(np.asarray(got[name], dtype=np.float64)[:, u]
                               - np.asarray(ref[name], dtype=np.float64)[:, u]).max())
            out["max_store_abs"] = max(out["max_store_abs"], dev)
            assert dev <= POST_TIE_STORE_ATOL, f"HRU {u} (tie on day {d}): {name} deviates by {dev}"
    return out


def assert_segments_with_ties(got: dict, ref: dict, names, tosegment, hru_segment, tied_hrus,
                              rtol=RTOL, atol=ATOL):
    """Full tolerance on the segments no tied HRU drains to; cumulative flow on the others."""
    mask = downstream_mask(tosegment, hru_segment, tied_hrus)
    clean = ~mask
    sub = lambda d: {k: np.asarray(d[k], dtype=np.float64)[:, clean] for k in set(names) | {"seg_outflow"}}  # noqa: E731
    assert_segments_close(sub(got), sub(ref), names, rtol=rtol, atol=atol)
    worst = 0.0
    if mask.any() and "seg_outflow" in names:
        g = np.asarray(got["seg_outflow"], dtype=np.float64)[:, mask].sum(axis=0)
        r = np.asarray(ref["seg_outflow"], dtype=np.float64)[:, mask].sum(axis=0)
        rel = np.abs(g - r) / np.maximum(np.abs(r), 1e-6)
        worst = float(rel.max())
        assert worst <= POST_TIE_CUM_RTOL, f"cumulative seg_outflow downstream of a tie off by {worst}"
    return {"segments_downstream_of_ties": int(mask.sum()), "max_cum_seg_rel": worst}


# ---------------------------------------------------------------- one-step parity
# The free-running comparison above cannot say much about an HRU after a
# ghost-pack tie: a pack that exists in one run and not in the other changes the
# depletion-curve memory (`ai`, `pst`, `scrv` ...) the next snowfall lands on, and
# the two trajectories -- both legitimate trajectories of the reference's
# algorithm -- stay apart for a season.  The one-step check has no such blind
# spot: the oracle is made to continue from the OTHER implementation's values
# of day d-1 (every prognostic quantity of the column is a public variable) and
# its day d must equal the other implementation's day d, for every HRU and every
# day, ties or not.  What can still differ is a threshold decided by the last bit
# INSIDE one day (a pack melting to 2.2e-16 vs 2.3e-16 inches against
# dnearzero = 2.23e-16): such HRU-days are counted and each is required to be that.
def volume_atol(name, hru_in_to_cf, atol=ATOL):
    """atol of a variable in ITS units: `*_vol` variables are inches x hru_in_to_cf."""
    if name.endswith("_vol"):
        return atol * np.maximum(np.asarray(hru_in_to_cf, dtype=np.float64), 1.0)
    return atol


def one_step_parity(oracle, forcings, got: dict, names, hru_in_to_cf, rtol=RTOL, atol=ATOL):
    """oracle: a fresh prms_oracle.Oracle of the same parameters.  Returns
    dict(hru_days, mismatching=[(hru, day, vars)], unexplained=[...])."""
    f = forcings
    one = oracle.run_forced(f["prcp"], f["tmax"], f["tmin"], got)
    nd, n = one[names[0]].shape
    bad = np.zeros((nd, n), dtype=bool)
    which = {}
    for k in names:
        g = np.asarray(got[k], dtype=np.float64)
        r = one[k]
        a = volume_atol(k, hru_in_to_cf, atol)
        ok = (np.abs(g - r) <= a + rtol * np.abs(r)) | (np.isnan(g) & np.isnan(r))
        if not ok.all():
            bad |= ~ok
            for d, u in zip(*np.where(~ok)):
                which.setdefault((int(u), int(d)), []).append(k)
    pk_g = np.asarray(got["pkwater_equiv"], dtype=np.float64)
    pk_r = one["pkwater_equiv"]
    mism, unexplained = [], []
    for (u, d), ks in sorted(which.items()):
        lo = max(d - 1, 0)
        w = np.concatenate([pk_g[lo:d + 1, u], pk_r[lo:d + 1, u]])
        rec = (u, d, ks[:6], float(np.asarray(got[ks[0]])[d, u]), float(one[ks[0]][d, u]))
        mism.append(rec)
        if not ((w > 0) & (w < GHOST)).any():
            unexplained.append(rec)
    return {"hru_days": nd * n, "mismatching": mism, "unexplained": unexplained}
