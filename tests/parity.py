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
