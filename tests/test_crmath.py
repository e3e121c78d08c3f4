"""pws_crmath.cuh (the elementary functions of k_soltab) compiled for the host:
every result is the correctly rounded one (checked against mpmath at 200 bits
on sampled arguments), and the machine's libm agrees with it except on a few
arguments in a thousand -- which is why the device tables of
PRMSSolarGeometry follow a host libm's almost everywhere."""
import ctypes as C
import pathlib as pl
import subprocess

import numpy as np
import pytest

ROOT = pl.Path(__file__).resolve().parent.parent
FUNCS = ("sin", "cos", "tan", "atan", "asin", "acos")


@pytest.fixture(scope="module")
def crlib():
    out = ROOT / "tests" / "_build" / "libcrmath_host.so"
    src = ROOT / "tests" / "_build" / "crmath_host.cpp"
    out.parent.mkdir(exist_ok=True)
    hdr = ROOT / "pywatershed_b200" / "csrc" / "pws_crmath.cuh"
    body = "\n".join(
        f"      case {k}: y_cr[i] = pws::cr::{f}(x[i]); y_libm[i] = ::{f}(x[i]); break;" for k, f in enumerate(FUNCS))
    src.write_text(f'#include "{hdr}"\nextern "C" void cr_eval(int fn, long n, const double* x, double* y_cr, '
                   f"double* y_libm) {{\n  for (long i = 0; i < n; ++i) {{\n    switch (fn) {{\n{body}\n    }}\n  }}\n}}\n")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-mfma", "-shared", "-fPIC", "-o", str(out), str(src)],
                   check=True)
    return C.CDLL(str(out))


def _eval(lib, fn, x):
    a, b = np.empty_like(x), np.empty_like(x)
    lib.cr_eval(fn, C.c_long(x.size), x.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p),
                b.ctypes.data_as(C.c_void_p))
    return a, b


def _args(name, n, rng):
    if name in ("asin", "acos"):
        return np.concatenate([rng.uniform(-1, 1, n - 6), [-1.0, 1.0, 0.0, 1 - 2.0 ** -53, -1 + 2.0 ** -53, 1e-300]])
    if name == "atan":
        return np.concatenate([rng.uniform(-3, 3, n // 2), rng.standard_cauchy(n - n // 2)])
    return np.concatenate([rng.uniform(-7, 13, n - 3), [0.0, np.pi / 2, np.pi]])


@pytest.mark.parametrize("fn", range(len(FUNCS)))
def test_correctly_rounded(crlib, fn):
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    name = FUNCS[fn]
    x = _args(name, 1500, np.random.default_rng(fn))
    got, _ = _eval(crlib, fn, x)
    f = getattr(mp, name)
    for xi, yi in zip(x, got):
        t = f(mp.mpf(float(xi)))
        e = abs(mp.mpf(float(yi)) - t)
        for nb in (np.nextafter(yi, np.inf), np.nextafter(yi, -np.inf)):
            assert e <= abs(mp.mpf(float(nb)) - t), (name, xi, yi)


@pytest.mark.parametrize("fn", range(len(FUNCS)))
def test_host_libm_agrees_almost_everywhere(crlib, fn):
    x = _args(FUNCS[fn], 400_000, np.random.default_rng(100 + fn))
    got, libm = _eval(crlib, fn, x)
    differ = got.view(np.int64) != libm.view(np.int64)
    assert np.abs(got - libm)[differ].max(initial=0.0) <= np.abs(np.spacing(got[differ])).max(initial=0.0)  # 1 ulp
    assert differ.mean() < 1e-2, (FUNCS[fn], differ.mean())
