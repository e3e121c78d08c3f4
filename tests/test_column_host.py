"""The device physics (pywatershed_b200/csrc/hru_column.cuh) compiled for the
HOST by tests/host_shim and run against the oracle: same inputs, every public
HRU variable of every day bit-identical.  This is a logic check that runs
without a GPU; the shipped library contains no such host path, and the GPU
parity proper is tests/test_gpu_parity.py."""
import ctypes as C
import pathlib as pl
import subprocess

import numpy as np
import pytest

import prms_oracle as po
import scenarios

ROOT = pl.Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def shim():
    out = ROOT / "tests" / "_build" / "libcolumn_host.so"
    out.parent.mkdir(exist_ok=True)
    src = ROOT / "tests" / "host_shim" / "column_host.cpp"
    subprocess.run(["g++", "-O2", "-fPIC", "-std=c++17", "-ffp-contract=off", "-shared",
                    "-o", str(out), str(src)], check=True)
    L = C.CDLL(str(out))
    L.shim_create.restype = C.c_void_p
    L.shim_create.argtypes = [C.c_int64, C.c_int]
    L.shim_set_f64.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]
    L.shim_set_i32.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]
    L.shim_set_opt.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    L.shim_init.argtypes = [C.c_void_p]
    L.shim_run.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6
    L.shim_var_name.restype = C.c_char_p
    L.shim_destroy.argtypes = [C.c_void_p]
    return L


def _run(L, P, F, start, nd, dprst_flag=True):
    n = P["hru_area"].shape[0]
    h = L.shim_create(n, P["snarea_curve"].size)
    for k in po.HRU_F8 + po.MON_F8 + ["snarea_curve"]:
        a = np.ascontiguousarray(np.asarray(P[k], dtype=np.float64).ravel())
        L.shim_set_f64(h, k.encode(), a.ctypes.data, a.size)
    for k in po.HRU_I4 + po.MON_I4:
        a = np.ascontiguousarray(np.asarray(P[k]).ravel().astype(np.int32))
        L.shim_set_i32(h, k.encode(), a.ctypes.data, a.size)
    for k in po.SCALAR_F8:
        L.shim_set_opt(h, k.encode(), float(np.ravel(P[k])[0]))
    cal = po.calendar(start, nd)
    L.shim_set_opt(h, b"dprst_flag", 1.0 if dprst_flag else 0.0)
    L.shim_set_opt(h, b"temp_units", float(np.ravel(P["temp_units"])[0]))
    L.shim_set_opt(h, b"start_month", float(cal[0, 2]))
    L.shim_set_opt(h, b"start_doy", float(cal[0, 0]))
    assert L.shim_init(h) == 0
    names = [L.shim_var_name(i).decode() for i in range(L.shim_n_vars())]
    out = np.empty((nd, len(names), n))
    viol = np.zeros((nd, 6), np.int32)
    pr, tx, tn = [np.ascontiguousarray(F[k][:nd]) for k in ("prcp", "tmax", "tmin")]
    assert L.shim_run(h, nd, cal.ctypes.data, pr.ctypes.data, tx.ctypes.data, tn.ctypes.data,
                      out.ctypes.data, viol.ctypes.data) == 0
    L.shim_destroy(h)
    return {k: out[:, i, :] for i, k in enumerate(names)}, viol


@pytest.mark.parametrize("scen", ["drb", "drbx"])
def test_device_physics_bit_identical_to_oracle(shim, scen, drb, drbx):
    P, F, start = drb if scen == "drb" else drbx
    nd = 731
    got, viol = _run(shim, P, F, start, nd)
    o = po.Oracle(P, start=start)
    ref, oviol = o.run(F["prcp"][:nd], F["tmax"][:nd], F["tmin"][:nd], po.HRU_VARS, ())
    for k in po.HRU_VARS:
        np.testing.assert_array_equal(got[k], ref[k], err_msg=k)
    np.testing.assert_array_equal(viol[:, :5], oviol[:, :5])


def test_device_physics_on_sagehen(shim, sagehen):
    """The snow-dominated second domain (deep packs, two depletion curves),
    NoDprst variant: device physics on the host == oracle, bit for bit."""
    P, F, start = sagehen
    nd = F["prcp"].shape[0]
    got, viol = _run(shim, P, F, start, nd, dprst_flag=False)
    o = po.Oracle(P, start=start, dprst_flag=False)
    ref, oviol = o.run(F["prcp"], F["tmax"], F["tmin"], po.HRU_VARS, ())
    for k in po.HRU_VARS:
        if k.startswith("channel_"):
            continue  # PRMSChannel's per-HRU variables: the process is not part of this model
        np.testing.assert_array_equal(got[k], ref[k], err_msg=k)
    np.testing.assert_array_equal(viol[:, :5], oviol[:, :5])


def test_device_physics_on_hru1_41_years(shim, hru1):
    """One HRU for 14,975 days: device physics on the host == oracle, bit for bit."""
    P, F, start = hru1
    nd = F["prcp"].shape[0]
    got, viol = _run(shim, P, F, start, nd)
    o = po.Oracle(P, start=start)
    ref, oviol = o.run(F["prcp"], F["tmax"], F["tmin"], po.HRU_VARS, ())
    for k in po.HRU_VARS:
        np.testing.assert_array_equal(got[k], ref[k], err_msg=k)
    np.testing.assert_array_equal(viol[:, :5], oviol[:, :5])


def test_div_by_ts_is_the_division(shim):
    """k_route_chunks divides by the Muskingum sub-daily step with
    q = x r, q + fma(-q, ts, x) r (r = correctly rounded 1/ts): must be the
    IEEE quotient for every ts PRMSChannel can produce (prms_channel.py:286-317)
    and for the daily mean's 24."""
    rng = np.random.default_rng(3)
    n = 4_000_000
    x = np.concatenate([
        rng.lognormal(0.0, 12.0, n), rng.random(n) * 1e-3, rng.random(n) * 1e6,
        np.array([0.0, 1.0, 3.0, 1e-300, 1e300, 5e-324, 2.2250738585072014e-308])])
    out = np.empty_like(x)
    shim.shim_div_by_ts.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_void_p]
    for ts in (2.0, 3.0, 4.0, 6.0, 8.0, 12.0, 24.0):
        shim.shim_div_by_ts(x.ctypes.data, x.size, ts, out.ctypes.data)
        ok = out == x / ts
        # the reformulation needs x / ts and the remainder to stay normal
        assert ok[x > 1e-290].all(), (ts, x[~ok][:5])


def test_div_r_is_the_division(shim):
    """column_lower divides by hru_area ~20 times a day through one true
    reciprocal and q = a r, q + fma(-q, b, a) r: must be the IEEE quotient."""
    rng = np.random.default_rng(4)
    n = 8_000_000
    a = rng.lognormal(0.0, 10.0, n) * rng.choice([-1.0, 1.0], n)
    b = rng.lognormal(0.0, 8.0, n)
    b[: n // 4] = np.nextafter(2.0 ** rng.integers(-20, 20, n // 4), 0.0)  # all-ones significands
    a[::1000] = 0.0
    b[1::7] = 86400.0  # seconds per day: the lateral-inflow conversion (prms_channel.py:423-425)
    out = np.empty_like(a)
    shim.shim_div_r.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    shim.shim_div_r(a.ctypes.data, b.ctypes.data, n, out.ctypes.data)
    np.testing.assert_array_equal(out, a / b)
    assert np.array_equal(np.signbit(out[a == 0.0]), np.signbit((a / b)[a == 0.0]))


def test_div0_is_the_division(shim):
    """div0 skips the division for a zero numerator and a positive divisor:
    same bits as IEEE division for zeros of either sign, zero / negative /
    infinite / NaN divisors, and ordinary operands."""
    rng = np.random.default_rng(5)
    n = 200_000
    a = rng.lognormal(0.0, 10.0, n) * rng.choice([-1.0, 1.0], n)
    b = rng.lognormal(0.0, 8.0, n) * rng.choice([-1.0, 1.0], n)
    a[::3] = 0.0
    a[1::30] = -0.0
    b[::11] = 0.0
    b[5::101] = -0.0
    b[7::103] = np.inf
    b[9::107] = -np.inf
    b[13::109] = np.nan
    a[17::113] = np.nan
    a[19::127] = np.inf
    out = np.empty_like(a)
    shim.shim_div0.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    shim.shim_div0(a.ctypes.data, b.ctypes.data, n, out.ctypes.data)
    with np.errstate(all="ignore"):
        ref = a / b
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    ok = ~np.isnan(ref)
    np.testing.assert_array_equal(out[ok].view(np.uint64), ref[ok].view(np.uint64))  # incl. the sign of zero
