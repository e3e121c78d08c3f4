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
#     infiltration and surface runoff (prms_runoff.py:905-935) -- water that
#     leaves the HRU either way -- and by the diagnostic state of a pack that
#     holds 1e-16 inches in one run and nothing in the other.  So the HRU's
#     water leaving as evapotranspiration and as flow to the stream, summed
#     over the run, must agree closely (POST_TIE_CUM_RTOL), its storages must
#     stay within POST_TIE_STORE_ATOL inches of each other, and its budgets
#     must keep closing (checked by the caller through the verdict counts).
#   * Segments that no tied HRU drains to (directly or through upstream
#     segments) keep the full tolerance; the others are compared on their
#     cumulative flow.
POST_TIE_CUM_RTOL = 2e-2
POST_TIE_STORE_ATOL = 0.5   # inches; soil / groundwater / pack storages
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
        for name in CUM_FLUXES:
            if name not in got or name not in ref:
                continue
            g = float(np.asarray(got[name], dtype=np.float64)[:, u].sum())
            r = float(np.asarray(ref[name], dtype=np.float64)[:, u].sum())
            rel = abs(g - r) / max(abs(r), 1e-6)
            if abs(g - r) > 1e-6 and rel > out["max_cum_rel"]:
                out["max_cum_rel"], out["worst"] = rel, (u, d, name, g, r)
            assert abs(g - r) <= 1e-6 + POST_TIE_CUM_RTOL * abs(r), (
                f"HRU {u} (tie on day {d} in {k}): cumulative {name} {g!r} vs {r!r}")
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
