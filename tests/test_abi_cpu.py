"""CPU-side checks of the boundary: the C-ABI library loads, exports every
symbol include/pws_b200.h declares, agrees with the oracle's registry, and
refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import pathlib as pl
import re

import numpy as np
import pytest

import prms_oracle as po

ROOT = pl.Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def L():
    from pywatershed_b200 import _lib

    if not _lib.LIB_PATH.exists():
        import __graft_entry__ as g

        g.build()
    return _lib.lib()


def test_header_symbols_are_exported(L):
    hdr = (ROOT / "include" / "pws_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pws_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    for sym in declared:
        assert hasattr(L, sym), sym
    from pywatershed_b200 import _lib

    assert declared == set(_lib.EXPORTED_SYMBOLS)


def test_registry_is_the_references(L):
    names = [L.pws_hru_var_name(i).decode() for i in range(L.pws_n_hru_vars())]
    assert names == po.HRU_VARS  # 143 public [nhru] variables, process order
    assert [L.pws_seg_var_name(i).decode() for i in range(L.pws_n_seg_vars())] == po.SEG_VARS
    kinds = {n: L.pws_hru_var_kind(i) for i, n in enumerate(names)}
    assert sorted(n for n, k in kinds.items() if k == 1) == sorted(po.INT_VARS)
    assert [n for n, k in kinds.items() if k == 2] == list(po.BOOL_VARS)
    assert L.pws_hru_var_name(L.pws_n_hru_vars()) is None


def test_no_cpu_fallback(L):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert L.pws_create(C.byref(h), 0, 8, 0, 11) != 0
    assert b"no CUDA device" in L.pws_last_error(None)
    from pywatershed_b200 import Engine, _lib

    import scenarios

    p, f, start = scenarios.load_drb()
    with pytest.raises(_lib.PwsLibraryError):
        Engine(p, start)


def test_product_does_not_touch_the_oracle():
    for path in (ROOT / "pywatershed_b200").rglob("*"):
        if path.suffix in (".py", ".cu", ".cuh", ".h", ".cpp"):
            txt = path.read_text()
            assert "prms_oracle" not in txt and "oracle/" not in txt, path


def test_calendar_matches_oracle_calendar():
    from pywatershed_b200 import calendar

    for start, n in (("1979-01-01", 1500), ("1999-09-28", 800)):
        np.testing.assert_array_equal(calendar(start, n), po.calendar(start, n))
