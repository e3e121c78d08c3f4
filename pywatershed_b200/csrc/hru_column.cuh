// hru_column.cuh -- the fused PRMS/NHM vertical column for ONE HRU and ONE day:
// atmosphere -> canopy -> snow -> runoff -> soilzone -> groundwater
// (+ the HRU's lateral inflow to its segment), written for one GPU thread per
// HRU with every prognostic state held in registers (struct HruState) across
// a block of days.  fp64 throughout, the reference's operation order kept so
// that results stay within rtol 1e-9 of pywatershed's numpy calc_method.
//
// What each part reproduces (paths relative to /root/reference/pywatershed/):
//   atm_day      atmosphere/prms_atmosphere.py:287-765
//   canopy_day   hydrology/prms_canopy.py:355-581
//   snow_day     hydrology/prms_snow.py:623-3222
//   runoff_day   hydrology/prms_runoff.py:543-1287
//   soil_day     hydrology/prms_soilzone.py:712-1456
//   gw_day       hydrology/prms_groundwater.py:231-275
//   lateral      hydrology/prms_channel.py:396-421
//   budgets      base/budget.py:336-386 (unit basis, rtol = atol = 1e-5)
// Cross-process coupling that the reference does through shared numpy buffers
// (Canopy zeroing Atmosphere's pptmix, Soilzone adding Dunnian flow into
// Runoff's sroff) happens here in registers.
//
// The file is plain C++ apart from the PWS_DEV / load macros, so tests can
// also compile it for the host (tests/host_shim) to debug the physics without
// a GPU; the shipped library only ever instantiates it inside CUDA kernels.
#pragma once
#include <math.h>
#include <stdint.h>

#include "pws_registry.h"
#include "pws_crmath.cuh"

#if defined(__CUDACC__)
#define PWS_DEV __device__ __forceinline__
#define PWS_HD __host__ __device__ __forceinline__
#define PWS_HDM __host__ __device__ __forceinline__
#define PWS_LD(p) (*(p))
#define PWS_TABLE static __device__ const
#define PWS_EXP10(x) exp10(x)
#else
#define PWS_DEV static inline
#define PWS_HD static inline
#define PWS_HDM inline
#define PWS_LD(p) (*(p))
#define PWS_TABLE static const
#define PWS_EXP10(x) pow(10.0, (x))
#endif

namespace pws {

// a / b that skips the division when a is zero.  IEEE: (+-0) / b = +-0 for
// every b > 0 (NaN and b <= 0 fall through to the division), so the result is
// bit-identical.  CUDA's fp64 division leaves its inline fast path for an
// out-of-line subroutine (~75 instructions) when the numerator is zero, and on
// dry days most numerators of the runoff / soilzone fluxes are: ncu counted 5
// such calls per HRU-day before this guard.
PWS_HD double div0(double a, double b) {
  if (a == 0.0 && b > 0.0) return a;
#ifdef __CUDA_ARCH__
  // volatile: otherwise the compiler divides unconditionally and selects afterwards
  double q;
  asm volatile("div.rn.f64 %0, %1, %2;" : "=d"(q) : "d"(a), "d"(b));
  return q;
#else
  return a / b;
#endif
}


// constants.py:30-44
#define PWS_NEARZERO 1.0e-6
#define PWS_DNEARZERO 2.23e-16
#define PWS_CLOSEZERO 1.20e-07
#define PWS_INCH2CM 2.54

enum { HRU_INACTIVE = 0, HRU_LAND = 1, HRU_LAKE = 2, HRU_SWALE = 3 };
enum { BUD_CANOPY = 0, BUD_SNOW, BUD_RUNOFF, BUD_SOILZONE, BUD_GW, BUD_CHANNEL, BUD_N };

// solar_constants.py:45-74
PWS_TABLE double k_solf[26] = {0.20,  0.35,  0.45,  0.51,  0.56,  0.59, 0.62,
                               0.64,  0.655, 0.67,  0.682, 0.69,  0.70, 0.71,
                               0.715, 0.72,  0.722, 0.724, 0.726, 0.728, 0.73,
                               0.734, 0.738, 0.742, 0.746, 0.75};
// prms_snow.py:26-64
PWS_TABLE double k_acum[15] = {0.80, 0.77, 0.75, 0.72, 0.70, 0.69, 0.68, 0.67,
                               0.66, 0.65, 0.64, 0.63, 0.62, 0.61, 0.60};
PWS_TABLE double k_amlt[15] = {0.72, 0.65, 0.60, 0.58, 0.56, 0.54, 0.52, 0.50,
                               0.48, 0.46, 0.44, 0.43, 0.42, 0.41, 0.40};

// SoA views of everything static; [12][stride] for monthly arrays
struct HruParams {
  int64_t nhru;
  int64_t stride;
#define X(n) const double* n;
  PWS_HRU_F64_PARAMS(X)
  PWS_MON_F64_PARAMS(X)
  PWS_HRU_F64_DERIVED(X)
#undef X
#define X(n) const int32_t* n;
  PWS_HRU_I32_PARAMS(X)
  PWS_MON_I32_PARAMS(X)
#undef X
  const double* tmax_allsnow_c;      // [12][stride], (tmax_allsnow-32)/1.8
  const double* snarea_curve;        // [ndeplval]
  const double* soltab_potsw;        // [366][stride]
  const double* soltab_horad_potsw;  // [366][stride]
  double albset_rna, albset_rnm, albset_sna, albset_snm;
  int dprst_flag, temp_units;
};

#define PWS_DECL_F(name) double name;
#define PWS_DECL_I(name) int name;
struct HruState {
#define X(name, kind) PWS_DECL_##kind(name)
  PWS_HRU_STATE(X)
#undef X
};

struct DayIn {
  double prcp, tmax, tmin, potsw, horad;
  int doy, dowy, month, dom;
  const int* cal = nullptr;  // {doy, dowy, month, dom} in memory, see PWS_REREAD
};

// Small integers that are needed late in the day (calendar fields, HRU / cover
// type, the transpiration flag handed to the lower half) live in a register
// from the top of the column to their last use, and such registers are what
// the compiler spills to local memory.  A sink with kReread (k_hru_column) has
// them in shared memory, so the column can read them again where it uses them
// (a volatile LDS) instead.  Measured per kind on the CONUS domain (column
// kernel ms per 10-year pass, none = 103.9): parameters 103.0, lower-half
// flags 104.9, calendar 107.0 (the re-read sits right in front of a branch
// and cannot be scheduled early) -- so only the parameters are re-read.
#ifdef __CUDA_ARCH__
#define PWS_LDV(p) (*(const volatile int32_t*)(p))
#define PWS_LDVB(p) ((int)*(const volatile unsigned char*)(p))
#else
#define PWS_LDV(p) (*(p))
#define PWS_LDVB(p) ((int)*(p))
#endif
#ifndef PWS_REREAD_CAL
#define PWS_REREAD_CAL 0
#endif
#ifndef PWS_REREAD_PARAM
#define PWS_REREAD_PARAM 1
#endif
#ifndef PWS_REREAD_MID
#define PWS_REREAD_MID 0
#endif
#define PWS_CAL(k, field) ((PWS_REREAD_CAL && Sink::kReread) ? (int)PWS_LDV(in.cal + (k)) : in.field)
#define PWS_IN_DOY PWS_CAL(0, doy)
#define PWS_IN_DOWY PWS_CAL(1, dowy)
#define PWS_IN_MONTH PWS_CAL(2, month)
#define PWS_IN_DOM PWS_CAL(3, dom)
#define PWS_PARAM_I(name) ((PWS_REREAD_PARAM && Sink::kReread) ? (int)PWS_LDV(P.name + i) : (int)PWS_LD(P.name + i))

// ---------------------------------------------------------------- helpers
PWS_DEV bool budget_open(double in, double out_plus_ds) {
  // np.allclose(in, out + dstor, rtol=1e-5, atol=1e-5)  base/budget.py:340-346
  return !(fabs(in - out_plus_ds) <= 1e-5 + 1e-5 * fabs(out_plus_ds));
}

// prms_canopy.py:575-581
PWS_DEV void intercept(double precip, double stor_max, double cov,
                       double& intcp_stor, double& net_precip) {
  net_precip = precip * (1.0 - cov);
  intcp_stor = intcp_stor + precip;
  if (intcp_stor > stor_max) {
    net_precip = net_precip + (intcp_stor - stor_max) * cov;
    intcp_stor = stor_max;
  }
}

// prms_snow.py:1580-1767
PWS_DEV void calc_calin(HruState& s, double& snowmelt, double cal,
                        double den_max, double denmaxinv, double freeh2o_cap) {
  const double dif = cal - s.pk_def;
  if (dif < 0.0) {
    s.pk_def = s.pk_def - cal;
    s.pk_temp = -1 * s.pk_def / (s.pkwater_equiv * 1.27);
  } else if (dif > 0.0) {
    const double pmlt = dif / 203.2;
    const double apmlt = pmlt * s.snowcov_area;
    s.pk_def = 0.0;
    s.pk_temp = 0.0;
    const double apk_ice = s.snowcov_area > 0.0 ? s.pk_ice / s.snowcov_area : 0.0;
    if (pmlt > apk_ice) {
      snowmelt = snowmelt + s.pkwater_equiv;
      s.pkwater_equiv = 0.0;
      s.iasw = 0;
      s.pk_ice = 0.0;
      s.freeh2o = 0.0;
      s.pk_depth = 0.0;
      s.pss = 0.0;
      s.pst = 0.0;
      s.pk_den = 0.0;
    } else {
      s.pk_ice = s.pk_ice - apmlt;
      s.freeh2o = s.freeh2o + apmlt;
      const double pwcap = freeh2o_cap * s.pk_ice;
      double dif_water = s.freeh2o - pwcap;
      if (dif_water > 0.0) {
        if (dif_water > s.pkwater_equiv) dif_water = s.pkwater_equiv;
        s.pkwater_equiv = s.pkwater_equiv - dif_water;
        s.freeh2o = pwcap;
        if (s.pk_den > 0.0) {
          s.pk_depth = s.pkwater_equiv / s.pk_den;
        } else {
          s.pk_den = den_max;
          s.pk_depth = s.pkwater_equiv * denmaxinv;
        }
        snowmelt = snowmelt + dif_water;
        s.pss = s.pkwater_equiv;
      }
    }
  } else {
    s.pk_temp = 0.0;
    s.pk_def = 0.0;
  }
  if (!(s.pkwater_equiv > 0.0)) s.pk_den = 0.0;
}

// prms_snow.py:1770-1848
PWS_DEV void calc_caloss(HruState& s, double cal) {
  if (s.freeh2o < PWS_CLOSEZERO) {
    s.pk_def = s.pk_def - cal;
  } else {
    const double calnd = s.freeh2o * 203.2;
    const double dif = cal + calnd;
    if (dif > 0.0) {
      s.pk_ice = s.pk_ice - (cal / 203.2);
      s.freeh2o = s.freeh2o + (cal / 203.2);
      return;
    }
    if (dif < 0.0) s.pk_def = -dif;
    s.pk_ice = s.pk_ice + s.freeh2o;
    s.freeh2o = 0.0;
  }
  if (s.pkwater_equiv > 0.0)
    s.pk_temp = -1 * s.pk_def / (s.pkwater_equiv * 1.27);
  else if (s.pkwater_equiv < 0.0)
    s.pkwater_equiv = 0.0;
}

// x**4 rounded once: x*x and (x*x)*(x*x) are split exactly with fma, so the
// result is the correctly rounded fourth power -- what a < 1 ulp host pow(x, 4.0)
// returns (prms_snow.py:2768 evaluates `(temp + 273.16) ** 4.0` through libm),
// where CUDA's pow() may be 2 ulp off.
PWS_DEV double pow4(double x) {
  const double t = x * x;
  const double e = fma(x, x, -t);
  const double p = t * t;
  const double pe = fma(t, t, -p);
  return p + (pe + 2.0 * t * e);
}

// prms_snow.py:2734-3093; one half-day energy balance, returns cal
PWS_DEV double calc_snowbal(HruState& s, double& snowmelt, int niteda, double cec,
                            double cst, double esv, double sw, double temp,
                            double trd, double canopy_covden, double den_max,
                            double denmaxinv, double emis_noppt,
                            double freeh2o_cap, double hru_ppt, int tstorm_mo) {
  const double tk = temp + 273.16;
  const double air = 0.585e-7 * pow4(tk);
  double emis = esv;
  double ts, sno;
  if (temp < 0.0) {
    ts = temp;
    sno = air;
  } else {
    ts = 0.0;
    sno = 325.7;
  }
  if (hru_ppt > 0.0 && tstorm_mo == 1) {
    if (niteda == 1) {
      emis = 0.85;
      if (trd > (1.0 / 3.0)) emis = emis_noppt;
    } else {
      if (trd > (1.0 / 3.0)) emis = 1.29 - (0.882 * trd);
      if (trd >= 0.5) emis = 0.95 - (0.2 * trd);
    }
  }
  const double sky = (1.0 - canopy_covden) * ((emis * air) - sno);
  const double can = canopy_covden * (air - sno);
  double cecsub = 0.0;
  if (temp > 0.0 && hru_ppt > 0.0) cecsub = cec * temp;
  const double cal = sky + can + cecsub + sw;
  bool do_calin = ts >= 0.0 && cal > 0.0;
  if (!do_calin) {
    const double qcond = cst * (ts - s.pk_temp);
    if (qcond < 0.0) {
      if (s.pk_temp < 0.0) {
        s.pk_def = s.pk_def - qcond;
        s.pk_temp = -1 * s.pk_def / (s.pkwater_equiv * 1.27);
      } else {
        calc_caloss(s, qcond);
      }
    } else if (qcond < PWS_CLOSEZERO) {
      do_calin = s.pk_temp >= 0.0 && cal > 0.0;
    } else if (ts >= 0.0) {
      const double pk_defsub = s.pk_def - qcond;
      if (pk_defsub < 0.0) {
        s.pk_def = 0.0;
        s.pk_temp = 0.0;
      } else {
        s.pk_def = pk_defsub;
        s.pk_temp = -pk_defsub / (s.pkwater_equiv * 1.27);
      }
    } else {
      const double pkt = -ts * (s.pkwater_equiv * 1.27);
      const double pks = s.pk_def - pkt;
      const double pk_defsub = pks - qcond;
      if (pk_defsub < 0.0) {
        s.pk_def = pkt;
        s.pk_temp = ts;
      } else {
        s.pk_def = pk_defsub + pkt;
        s.pk_temp = -1 * s.pk_def / (s.pkwater_equiv * 1.27);
      }
    }
  }
  // one call site for both the direct and the qcond ~ 0 melt paths (:2812, :2957)
  if (do_calin) calc_calin(s, snowmelt, cal, den_max, denmaxinv, freeh2o_cap);
  return cal;
}

// prms_runoff.py:1231-1259
PWS_DEV void perv_comp(double soil_moist_prev, double carea_max, double smidx_coef,
                       double smidx_exp, double pptp, double ptc, double& infil,
                       double& srp, double& contrib_fraction) {
  const double smidx = soil_moist_prev + 0.5 * ptc;
  double ca_fraction;
  if (smidx > 25.0) ca_fraction = carea_max;
  else ca_fraction = smidx_coef * PWS_EXP10(smidx_exp * smidx);
  if (ca_fraction > carea_max) ca_fraction = carea_max;
  const double srpp = ca_fraction * pptp;
  infil = infil - srpp;
  srp = srp + srpp;
  contrib_fraction = ca_fraction;
}

// prms_runoff.py:1262-1275
PWS_DEV void check_capacity(double soil_moist_prev, double soil_moist_max,
                            double snowinfil_max, double& infil, double& srp) {
  const double capacity = soil_moist_max - soil_moist_prev;
  const double excess = infil - capacity;
  if (excess > snowinfil_max) {
    srp = srp + excess - snowinfil_max;
    infil = snowinfil_max + capacity;
  }
}

// depression surface area from volume, prms_runoff.py:1100-1132
PWS_DEV double dprst_area(double vol, double vol_max, double area_max, double va_exp) {
  double area = 0.0;
  if (vol > 0.0) {
    const double r = div0(vol, vol_max);
    double frac;
    if (r < PWS_NEARZERO) frac = 0.0;
    else if (r > 1.0) frac = 1.0;
    else frac = exp(va_exp * log(r));
    area = area_max * frac;
    if (area > area_max) area = area_max;
  }
  return area;
}

// prms_soilzone.py:1260-1312; returns false on the reference's RuntimeError
PWS_DEV bool compute_interflow(double coef_lin, double coef_sq, double ssres_in,
                               double& storage, double& inter_flow) {
  double flow;
  if (coef_lin <= 0.0 && ssres_in <= 0.0) {
    const double c1 = coef_sq * storage;
    flow = storage * (c1 / (1.0 + c1));
  } else if (coef_lin > 0.0 && coef_sq <= 0.0) {
    const double c2 = 1.0 - exp(-coef_lin);
    flow = ssres_in * (1.0 - c2 / coef_lin) + storage * c2;
  } else if (coef_sq > 0.0) {
    const double c3 = sqrt(coef_lin * coef_lin + 4.0 * coef_sq * ssres_in);
    const double sos = storage - div0(c3 - coef_lin, 2.0 * coef_sq);
    if (c3 == 0.0) return false;
    const double c1 = coef_sq * sos / c3;
    const double c2 = 1.0 - exp(-c3);
    if ((1.0 + c1 * c2) > 0.0)
      flow = ssres_in + ((sos * (1.0 + c1) * c2) / (1.0 + c1 * c2));
    else
      flow = ssres_in;
  } else {
    flow = 0.0;
  }
  if (flow < 0.0) flow = 0.0;
  else if (flow > storage) flow = storage;
  storage = storage - flow;
  inter_flow = flow;
  return true;
}

// x / ts for ts in {2, 3, 4, 6, 8, 12, 24} with the correctly rounded
// reciprocal r: q = RN(x r), q' = RN(q + RN(x - ts q) r) is the correctly
// rounded quotient (Markstein), i.e. bit-identical to the reference's
// division, in three dependent instructions instead of the division
// subroutine (checked exhaustively-at-random in tests/test_column_host.py::test_div_by_ts_is_the_division).
PWS_HD double div_by_ts(double x, double ts, double r) {
  const double q = x * r;
  const double rem = fma(-q, ts, x);
  return fma(rem, r, q);
}
// a / b given r = RN(1 / b) (a true IEEE division): the same Markstein
// sequence, correctly rounded for any finite normal a, b (1.8e10 random and
// adversarial operand pairs checked against the division; zero and sign are
// preserved; an infinite a would give NaN, which this path never sees).
PWS_HD double div_r(double a, double b, double r) {
  const double q = a * r;
  const double rem = fma(-q, b, a);
  return fma(rem, r, q);
}

constexpr double kPi = 3.141592653589793;

// ---------------------------------------------------------------- soltab
// PRMSSolarGeometry (atmosphere/prms_solar_geometry.py:176-408), split into the
// per-HRU part (slope/aspect/latitude geometry) and the per-(doy, HRU) part.
// Host+device, with the elementary functions as a policy M: k_soltab (the
// default) instantiates it with CrMath -- correctly rounded sin / cos / tan /
// asin / acos / atan (pws_crmath.cuh), because the reference's own tables come
// from a host libm and PRMSSnow's melt-out is sensitive to their last bit
// (tests/parity.py) -- and the host variant (option "soltab_device" 0, and the
// host shim of tests/test_column_host.py) with LibmMath, the machine's libm.
struct LibmMath {
  static PWS_HDM double sin(double x) { return ::sin(x); }
  static PWS_HDM double cos(double x) { return ::cos(x); }
  static PWS_HDM double tan(double x) { return ::tan(x); }
  static PWS_HDM double asin(double x) { return ::asin(x); }
  static PWS_HDM double acos(double x) { return ::acos(x); }
  static PWS_HDM double atan(double x) { return ::atan(x); }
};
struct CrMath {
  static PWS_HDM double sin(double x) { return cr::sin(x); }
  static PWS_HDM double cos(double x) { return cr::cos(x); }
  static PWS_HDM double tan(double x) { return cr::tan(x); }
  static PWS_HDM double asin(double x) { return cr::asin(x); }
  static PWS_HDM double acos(double x) { return cr::acos(x); }
  static PWS_HDM double atan(double x) { return cr::atan(x); }
};
struct SolHru {
  double sl, x0, x1, x2, sin_x1, cos_x1, tan_x1, sin_x0, cos_x0, tan_x0;
};

template <class M = LibmMath>
PWS_HD SolHru sol_hru(double slope, double aspect, double lat) {
  const double deg2rad = kPi / 180.0;
  SolHru h;
  h.sl = M::atan(slope);
  const double sl_sin = M::sin(h.sl), sl_cos = M::cos(h.sl);
  const double aspects_rad = aspect * deg2rad;
  const double aspects_cos = M::cos(aspects_rad);
  h.x0 = lat * deg2rad;
  h.cos_x0 = M::cos(h.x0);
  h.sin_x0 = M::sin(h.x0);
  // x1: latitude of the equivalent slope (Lee 1963 eq. 13), :207
  h.x1 = M::asin(sl_cos * h.sin_x0 + sl_sin * h.cos_x0 * aspects_cos);
  double d1 = sl_cos * h.cos_x0 - sl_sin * h.sin_x0 * aspects_cos;  // :210-211
  if (fabs(d1) < PWS_DNEARZERO) d1 = PWS_DNEARZERO;
  h.x2 = M::atan(sl_sin * M::sin(aspects_rad) / d1);  // :216
  if (d1 < 0.0) h.x2 = h.x2 + kPi;
  h.sin_x1 = M::sin(h.x1);
  h.cos_x1 = M::cos(h.x1);
  h.tan_x1 = M::tan(h.x1);
  h.tan_x0 = M::tan(h.x0);
  return h;
}

// compute_t :319-358 with tan(lat), tan(decl) hoisted
template <class M = LibmMath>
PWS_HD double sol_compute_t(double tan_lat, double tan_decl) {
  const double tx = (-1 * tan_lat) * tan_decl;
  if (tx < -1.0) return kPi;
  if (tx > 1.0) return 0.0;
  return M::acos(tx);
}

// func3 :361-408 with sin/cos of the declination and of w hoisted
template <class M = LibmMath>
PWS_HD double sol_func3(double v, double sin_w, double cos_w, double x, double y, double r1,
                        double sin_d, double cos_d) {
  const double pi_12 = 12 / kPi;
  return r1 * pi_12 * (sin_d * sin_w * (x - y) + cos_d * cos_w * (M::sin(x + v) - M::sin(y + v)));
}

// One (doy, HRU) of compute_soltab :230-316.  t3/t7 and t2/t6 are ALIASES in
// the reference (:255-272), so the clamped values feed t6' and t7'.
template <class M = LibmMath>
PWS_HD void sol_day(const SolHru& h, double sin_d, double cos_d, double tan_d, double r1,
                    double& solt, double& sunh) {
  const double two_pi = 2 * kPi, pi_12 = 12 / kPi;
  double tt = sol_compute_t<M>(h.tan_x1, tan_d);
  double t6 = (-1 * tt) - h.x2;
  double t7 = tt - h.x2;
  tt = sol_compute_t<M>(h.tan_x0, tan_d);
  const double t0 = -1 * tt, t1 = tt;
  if (t7 > t1) t7 = t1;
  double t3 = t7;
  if (t6 < t0) t6 = t0;
  double t2 = t6;
  t6 = t6 + two_pi;
  t7 = t7 - two_pi;
  if (t3 < t2) {
    t2 = 0.0;
    t3 = 0.0;
  }
  const double base = sol_func3<M>(h.x2, h.sin_x1, h.cos_x1, t3, t2, r1, sin_d, cos_d);
  solt = base;
  sunh = (t3 - t2) * pi_12;
  if (t7 > t0) {
    solt = base + sol_func3<M>(h.x2, h.sin_x1, h.cos_x1, t7, t0, r1, sin_d, cos_d);
    sunh = (t3 - t2 + t7 - t0) * pi_12;
  }
  if (t6 < t1) {
    solt = base + sol_func3<M>(h.x2, h.sin_x1, h.cos_x1, t1, t6, r1, sin_d, cos_d);
    sunh = (t3 - t2 + t1 - t6) * pi_12;
  }
  if (fabs(h.sl) < PWS_DNEARZERO) {
    solt = sol_func3<M>(0.0, h.sin_x0, h.cos_x0, t1, t0, r1, sin_d, cos_d);
    sunh = (t1 - t0) * pi_12;
  }
  if (sunh < PWS_DNEARZERO) sunh = 0.0;
  if (solt < 0.0) solt = 0.0;
}

// solar_constants.py:10-43 for julian day j+1: declination and r1
template <class M = LibmMath>
PWS_HD void sol_constants(int j, double& decl, double& r1) {
  const double rad_day = (2 * kPi) / 365.242;
  const double jd = (double)(j + 1);
  const double obliquity = 1 - (0.01671 * M::cos((jd - 3) * rad_day));
  const double yy = (jd - 1) * rad_day, yy2 = yy * 2, yy3 = yy * 3;
  decl = 0.006918 - 0.399912 * M::cos(yy) + 0.070257 * M::sin(yy) - 0.006758 * M::cos(yy2) +
         0.000907 * M::sin(yy2) - 0.002697 * M::cos(yy3) + 0.00148 * M::sin(yy3);
  r1 = (60.0 * 2.0) / (obliquity * obliquity);
}

struct HruDerivedMut {
#define X(n) double* n;
  PWS_HRU_F64_DERIVED(X)
#undef X
  double* tmax_allsnow_c;
};

// One HRU: what the reference derives in the constructors (see pws_b200.h:
// pws_init) and the initial prognostic state.  Returns an error mask
// (2: snowpack_init > 0, 4: dprst_depth_avg == 0, 8: pref_flow_infil_frac range).
PWS_DEV int hru_init_one(const HruParams& P, const HruDerivedMut& D, double* state,
                         const int64_t i, int start_month, int start_doy,
                         int adjust_parameters) {
  int errmask = 0;
  const int64_t S = P.stride;
  auto ST = [&](int k) -> double& { return state[(int64_t)k * S + i]; };
#define X(name, kind) ST(ST_##name) = 0.0;
  PWS_HRU_STATE(X)
#undef X
  ST(ST_int_alb) = 1.0;
  ST(ST_iso) = 1.0;
  ST(ST_mso) = 1.0;
  const int ht = P.hru_type[i];
  const double harea = P.hru_area[i];
  // --- atmosphere: prms_atmosphere.py:519, 690-716
  // hru_cossl = cos(atan(hru_slope)) is filled by the caller (host libm, see pws_init)
  {
    const int beg = P.transp_beg[i], end = P.transp_end[i];
    const int motmp = start_month + 12;
    int on = 0, check = 0;
    if (start_month == beg) {
      if (start_doy > 10) on = 1;
      else check = 1;
    } else if (end > beg) {
      if (start_month > beg && start_month < end) on = 1;
    } else {
      if (start_month > beg || motmp < end + 12) on = 1;
    }
    ST(ST_transp_on) = on;
    ST(ST_transp_check) = check;
  }
  // --- snow: prms_snow.py:295-431 (snowpack_init > 0 cannot run upstream)
  for (int m = 0; m < 12; ++m)
    D.tmax_allsnow_c[(int64_t)m * S + i] = (P.tmax_allsnow[(int64_t)m * S + i] - 32.0) / 1.8;
  D.deninv[i] = 1.0 / P.den_init[i];
  D.denmaxinv[i] = 1.0 / P.den_max[i];
  if (P.snowpack_init[i] > 0.0) errmask |= 2;
  // --- groundwater: prms_groundwater.py:145-149
  ST(ST_gwres_stor) = P.gwstor_init[i];
  // --- runoff basin_init: prms_runoff.py:253-291
  {
    double perv_area = harea, imperv = 0.0;
    if (P.hru_percent_imperv[i] > 0.0) {
      imperv = P.hru_percent_imperv[i] * harea;
      perv_area = perv_area - imperv;
    }
    double area_max = 0.0, open_max = 0.0, clos_max = 0.0, frac_clos = 0.0;
    if (P.dprst_flag) {
      area_max = P.dprst_frac[i] * harea;
      if (area_max > 0.0) {
        open_max = area_max * P.dprst_frac_open[i];
        frac_clos = 1.0 - P.dprst_frac_open[i];
        clos_max = area_max - open_max;
        perv_area = perv_area - area_max;
      }
    }
    D.ro_hru_imperv[i] = imperv;
    D.ro_hru_perv[i] = perv_area;
    D.ro_hru_frac_perv[i] = perv_area / harea;
    D.dprst_area_max[i] = area_max;
    D.dprst_area_open_max[i] = open_max;
    D.dprst_area_clos_max[i] = clos_max;
    D.dprst_frac_clos[i] = frac_clos;
    // dprst_init: prms_runoff.py:293-389
    double vol_open_max = 0.0, vol_clos_max = 0.0, thres = 0.0, area_clos = 0.0;
    if (P.dprst_flag && P.dprst_frac[i] > 0.0) {
      if (P.dprst_depth_avg[i] == 0.0) errmask |= 4;
      vol_clos_max = clos_max * P.dprst_depth_avg[i];
      vol_open_max = open_max * P.dprst_depth_avg[i];
      const double vol_open = P.dprst_frac_init[i] * vol_open_max;
      const double vol_clos = P.dprst_frac_init[i] * vol_clos_max;
      thres = P.op_flow_thres[i] * vol_open_max;
      ST(ST_dprst_vol_open) = vol_open;
      ST(ST_dprst_vol_clos) = vol_clos;
      ST(ST_dprst_area_open) = dprst_area(vol_open, vol_open_max, open_max, P.va_open_exp[i]);
      area_clos = dprst_area(vol_clos, vol_clos_max, clos_max, P.va_clos_exp[i]);
    }
    D.dprst_vol_open_max[i] = vol_open_max;
    D.dprst_vol_clos_max[i] = vol_clos_max;
    D.dprst_vol_thres_open[i] = thres;
    D.dprst_area_clos_init[i] = area_clos;
  }
  // --- soilzone: prms_soilzone.py:261-536
  {
    const double hru_area_imperv = P.hru_percent_imperv[i] * harea;
    double hru_area_perv = harea - hru_area_imperv;
    double soil_moist_max = P.soil_moist_max[i];
    double soil_rechr_max = P.soil_rechr_max_frac[i] * soil_moist_max;
    double hru_frac_perv = 1.0 - P.hru_percent_imperv[i];
    if (ht != HRU_INACTIVE) {
      if (P.dprst_flag) hru_area_perv = hru_area_perv - P.dprst_frac[i] * harea;
      hru_frac_perv = hru_area_perv / harea;
    }
    const bool off = ht == HRU_INACTIVE || ht == HRU_LAKE;
    const double sat_threshold = off ? 0.0 : P.sat_threshold[i];
    const double pref_flow_den = ht == HRU_LAND ? P.pref_flow_den[i] : 0.0;
    if (P.pref_flow_infil_frac[i] < 0.0 || P.pref_flow_infil_frac[i] > 1.0) errmask |= 8;
    double soil_moist = P.soil_moist_init_frac[i] * soil_moist_max;
    double soil_rechr = P.soil_rechr_init_frac[i] * soil_rechr_max;
    double ssres_stor = off ? 0.0 : P.ssstor_init_frac[i] * sat_threshold;
    if (adjust_parameters) {  // :361-463, the "warn" behaviour
      if (soil_moist_max < 1.0e-5) soil_moist_max = 1.0e-5;
      if (soil_rechr_max < 1.0e-5) soil_rechr_max = 1.0e-5;
      if (soil_rechr_max > soil_moist_max) soil_rechr_max = soil_moist_max;
      if (soil_rechr > soil_rechr_max) soil_rechr = soil_rechr_max;
      if (soil_moist > soil_moist_max) soil_moist = soil_moist_max;
      if (soil_rechr > soil_moist) soil_rechr = soil_moist;
      if (ssres_stor > sat_threshold) ssres_stor = sat_threshold;
    }
    double thrsh = 0.0, pfmax = 0.0;
    if (ht == HRU_SWALE) thrsh = sat_threshold;
    if (ht == HRU_LAND) {
      thrsh = sat_threshold * (1.0 - pref_flow_den);
      pfmax = sat_threshold - thrsh;
    }
    if (ht == HRU_LAND || ht == HRU_SWALE) {
      const double slow = fmin(ssres_stor, thrsh);
      ST(ST_slow_stor) = slow;
      ST(ST_pref_flow_stor) = ssres_stor - slow;
    }
    ST(ST_soil_moist) = soil_moist;
    ST(ST_soil_rechr) = soil_rechr;
    D.sz_hru_frac_perv[i] = hru_frac_perv;
    D.sz_soil_rechr_max[i] = soil_rechr_max;
    D.sz_soil_moist_max[i] = soil_moist_max;
    D.sz_sat_threshold[i] = sat_threshold;
    D.sz_pref_flow_den[i] = pref_flow_den;
    D.pref_flow_thrsh[i] = thrsh;
    D.pref_flow_max[i] = pfmax;
    D.soil_zone_max[i] = sat_threshold + soil_moist_max * hru_frac_perv;
    const double lower_max = soil_moist_max - soil_rechr_max;
    D.soil_lower_max[i] = lower_max;
    D.soil_lower_ratio_init[i] = lower_max > 0.0 ? (soil_moist - soil_rechr) / lower_max : 0.0;
  }
  return errmask;
}

// ------------------------------------------------------------------------
// One HRU, one day.  `i` indexes the SoA arrays.  `out` receives every public
// variable (PWS_F / PWS_I: the variable is a template argument of the sink's
// f<V_x>() / i<V_x>(), so a sink can turn it into a compile-time address),
// budget closure flags (out.viol), the lateral inflow contribution
// (out.lateral) and device-side error flags (out.error).  out.sync() marks the
// process boundaries and out.sync2() finer points inside the large processes:
// CTA-wide barriers on the GPU when PWS_COLUMN_SYNC is 1 / 2, so that the
// warps of a CTA walk the code together and share instruction fetches (every
// such point is in convergent code: all lanes of the CTA reach it).
// ------------------------------------------------------------------------
#define PWS_F(var, val) out.template f<V_##var>(val)
#define PWS_I(var, val) out.template i<V_##var>(val)
// Input override points: when a process of the chain is NOT part of the model
// (a sub-chain, or a single process driven from files as the reference's own
// tests do, autotest/test_self_drive.py), the variables it would have produced
// are taken from the caller's input arrays instead.  For the fused full-chain
// kernels out.ov<>() is the identity and compiles to nothing.
#define PWS_OV(name, var) var = out.template ov<OV_##name>(var)
#define PWS_OVI(name, var) var = out.template ovi<OV_##name>(var)
// What the upper half of the column (atmosphere, canopy, snow) hands to the
// lower half (runoff, soilzone, groundwater) for one HRU-day.
struct ColumnMid {
  double potet, net_rain, net_snow, hru_intcpevap, intcp_changeover;
  double snowmelt, snow_evap, pkwater_equiv, snowcov_area, through_rain;
  int transp_on, pptmix_nopack;
  const unsigned char* flags = nullptr;  // bit 0 transp_on, bit 1 pptmix_nopack (PWS_REREAD)
};

// ---- upper half: PRMSAtmosphere, PRMSCanopy, PRMSSnow
template <class Sink>
PWS_DEV void column_upper(const HruParams& P, const int64_t i, HruState& s,
                          const DayIn& in, Sink& out, ColumnMid& mid) {
  const int64_t mi = (int64_t)(in.month - 1) * P.stride + i;
  // hru_type / cov_type: PWS_PARAM_I at their uses

  // ===================== PRMSAtmosphere ==================================
  // adjust_temperature prms_atmosphere.py:287-317
  const double tmaxf = in.tmax + PWS_LD(P.tmax_cbh_adj + mi);
  const double tminf = in.tmin + PWS_LD(P.tmin_cbh_adj + mi);
  double tminc = (tminf - 32.0) * (5.0 / 9.0);
  double tmaxc = (tmaxf - 32.0) * (5.0 / 9.0);
  double tavgc = (tmaxc + tminc) / 2.0;
  // adjust_precip :319-430 (mixed everywhere, then all-rain, then all-snow)
  const double allsnow = PWS_LD(P.tmax_allsnow + mi);
  const double allrain_off = PWS_LD(P.tmax_allrain_offset + mi);
  const double allrain = allsnow + allrain_off;
  double tdiff = tmaxf - tminf;
  if (tdiff < PWS_NEARZERO) tdiff = 1.0e-4;
  double prmx = ((tmaxf - allsnow) / tdiff) * PWS_LD(P.adjmix_rain + mi);
  if (prmx < 0.0) prmx = 0.0;
  if (prmx > 1.0) prmx = 1.0;
  if (tminf > allsnow || tmaxf >= allrain) prmx = 1.0;
  if (tmaxf <= allsnow) prmx = 0.0;
  if (in.prcp <= 0.0) prmx = 0.0;
  double hru_ppt = in.prcp * PWS_LD(P.snow_cbh_adj + mi);
  double hru_rain = prmx * hru_ppt;
  double hru_snow = hru_ppt - hru_rain;
  if (prmx <= 0.0) {
    hru_snow = hru_ppt;
    hru_rain = 0.0;
  }
  if (prmx >= 1.0) {
    hru_ppt = in.prcp * PWS_LD(P.rain_cbh_adj + mi);
    hru_rain = hru_ppt;
    hru_snow = 0.0;
  }
  int pptmix = (hru_ppt > 0.0) && (tmaxf > allsnow) && (tminf <= allsnow) &&
               (tmaxf < allrain) && (prmx < 1.0);
  // _ddsolrad_run :475-620
  double dday = (PWS_LD(P.dday_slope + mi) * tmaxf) + PWS_LD(P.dday_intcp + mi) + 1.0;
  if (dday < 1.0) dday = 1.0;
  const double radmax = PWS_LD(P.radmax + mi);
  double radadj = radmax;
  if (dday < 26.0) {
    const int kp = (int)dday;
    radadj = k_solf[kp - 1] + ((k_solf[kp] - k_solf[kp - 1]) * (dday - kp));
    if (radadj > radmax) radadj = radmax;
  }
  double pptadj = 1.0;
  if (hru_ppt > PWS_LD(P.ppt_rad_adj + mi)) {
    const double tmax_index = PWS_LD(P.tmax_index + mi);
    pptadj = PWS_LD(P.radadj_intcp + mi) +
             PWS_LD(P.radadj_slope + mi) * (tmaxf - tmax_index);
    if (pptadj > 1.0) pptadj = 1.0;
    if (tmaxf < tmax_index) {
      const double allrain2 = allrain_off + allsnow;
      const bool is_summer = (PWS_IN_DOY >= 79) && (PWS_IN_DOY <= 265);
      pptadj = PWS_LD(P.radj_sppt + i);
      if (tmaxf < allrain2 || !is_summer) pptadj = PWS_LD(P.radj_wppt + i);
    }
  }
  radadj = radadj * pptadj;
  if (radadj < 0.2) radadj = 0.2;
  double swrad = in.potsw * radadj / PWS_LD(P.hru_cossl + i);
  double orad_hru = radadj * in.horad;
  // _potet_jh_run :643-668
  const double tavgf = (tavgc * 9 / 5) + 32;
  const double elh = (597.3 - (0.5653 * tavgc)) * PWS_INCH2CM;
  double potet = PWS_LD(P.jh_coef + mi) * (tavgf - PWS_LD(P.jh_coef_hru + i)) * swrad / elh;
  if (potet < 0.0) potet = 0.0;
  // calculate_transp_tindex :718-765 (the per-day part of the state machine)
  {
    double transp_tmax_f = PWS_LD(P.transp_tmax + i);
    if (P.temp_units != 0) transp_tmax_f = (transp_tmax_f * (9.0 / 5.0)) + 32.0;
    if (PWS_IN_DOM == 1) {
      if (PWS_IN_MONTH == PWS_LD(P.transp_end + i)) {
        s.transp_on = 0;
        s.transp_check = 0;
        s.tmax_sum = 0.0;
      }
      if (PWS_IN_MONTH == PWS_LD(P.transp_beg + i)) {
        s.transp_check = 1;
        s.tmax_sum = 0.0;
      }
    }
    if (s.transp_check == 1) {
      if (tmaxf > 32.0) s.tmax_sum = s.tmax_sum + tmaxf;
      if (s.tmax_sum > transp_tmax_f) {
        s.transp_on = 1;
        s.transp_check = 0;
        s.tmax_sum = 0.0;
      }
    }
  }
  int transp_on = s.transp_on;
  PWS_F(tmaxf, tmaxf); PWS_F(tminf, tminf); PWS_F(prmx, prmx);
  PWS_F(hru_ppt, hru_ppt); PWS_F(hru_rain, hru_rain); PWS_F(hru_snow, hru_snow);
  PWS_F(swrad, swrad); PWS_F(potet, potet); PWS_I(transp_on, transp_on);
  PWS_F(tmaxc, tmaxc); PWS_F(tavgc, tavgc); PWS_F(tminc, tminc);
  PWS_I(pptmix, pptmix);  // the all-time value Atmosphere writes to file
  PWS_F(orad_hru, orad_hru);
  PWS_OV(tmaxc, tmaxc); PWS_OV(tminc, tminc); PWS_OV(tavgc, tavgc); PWS_OV(prmx, prmx);
  PWS_OV(hru_ppt, hru_ppt); PWS_OV(hru_rain, hru_rain); PWS_OV(hru_snow, hru_snow);
  PWS_OV(swrad, swrad); PWS_OV(orad_hru, orad_hru); PWS_OV(potet, potet);
  PWS_OVI(transp_on, transp_on); PWS_OVI(pptmix, pptmix);

  out.sync();
  // ===================== PRMSCanopy ======================================
  const double covden_sum = PWS_LD(P.covden_sum + i);
  const double covden_win = PWS_LD(P.covden_win + i);
  const double hru_intcpstor_old = s.hru_intcpstor;  // _advance_variables :277
  double pk_ice_prev = s.pk_ice;               // prms_snow.py:483-486
  double freeh2o_prev = s.freeh2o;
  PWS_OV(pk_ice_prev, pk_ice_prev); PWS_OV(freeh2o_prev, freeh2o_prev);
  double net_rain = hru_rain, net_snow = hru_snow;
  double hru_intcpevap, intcp_changeover;
  {
    double cov, stor_max_rain;
    if (transp_on == 1) {
      cov = covden_sum;
      stor_max_rain = PWS_LD(P.srain_intcp + i);
    } else {
      cov = covden_win;
      stor_max_rain = PWS_LD(P.wrain_intcp + i);
    }
    const int intcp_form = hru_snow > 0.0 ? 1 : 0;
    double intcpstor = s.intcp_stor;
    double intcpevap = 0.0, changeover = 0.0, extra_water = 0.0;
    // hru_type is forced to LAND in the reference (prms_canopy.py:92)
    if (PWS_PARAM_I(cov_type) == 0) {
      if (intcpstor > 0.0) extra_water = s.intcp_stor;
      intcpstor = 0.0;
    }
    if (transp_on == 0 && s.intcp_transp_on == 1) {
      s.intcp_transp_on = 0;
      if (intcpstor > 0.0) {
        const double diff = covden_sum - cov;
        changeover = intcpstor * diff;
        if (cov > 0.0) {
          if (changeover < 0.0) {
            intcpstor = intcpstor * covden_sum / cov;
            changeover = 0.0;
          }
        } else {
          intcpstor = 0.0;
        }
      }
    } else if (transp_on == 1 && s.intcp_transp_on == 0) {
      s.intcp_transp_on = 1;
      if (intcpstor > 0.0) {
        const double diff = covden_win - cov;
        changeover = intcpstor * diff;
        if (cov > 0.0) {
          if (changeover < 0.0) {
            intcpstor = intcpstor * covden_win / cov;
            changeover = 0.0;
          }
        } else {
          intcpstor = 0.0;
        }
      }
    }
    if (PWS_PARAM_I(cov_type) != 0 && cov > 0.0) {
      if (hru_rain > 0.0) {
        if (PWS_PARAM_I(cov_type) > 1) {
          intercept(hru_rain, stor_max_rain, cov, intcpstor, net_rain);
        } else if ((pk_ice_prev + freeh2o_prev) < PWS_DNEARZERO && net_snow < PWS_NEARZERO) {
          intercept(hru_rain, stor_max_rain, cov, intcpstor, net_rain);
        }
      }
    }
    if (hru_snow > 0.0 && cov > 0.0 && PWS_PARAM_I(cov_type) > 1) {
      intercept(hru_snow, PWS_LD(P.snow_intcp + i), cov, intcpstor, net_snow);
      if (net_snow < PWS_NEARZERO) {
        net_rain = net_rain + net_snow;
        net_snow = 0.0;
        pptmix = 0;  // prms_canopy.py:500-505: Snow sees the edited flag
      }
    }
    if (intcpstor > 0.0 && hru_ppt < PWS_NEARZERO) {
      const double evrn = potet / 1.0;
      const double evsn = potet * PWS_LD(P.potet_sublim + i);
      const double ev = intcp_form == 1 ? evsn : evrn;
      const double z = intcpstor - ev;
      if (z > 0.0) {
        intcpstor = z;
        intcpevap = ev;
      } else {
        intcpevap = intcpstor;
        intcpstor = 0.0;
      }
    }
    if (intcpevap * cov > potet) {
      const double last = intcpevap;
      intcpevap = cov > 0.0 ? potet / cov : 0.0;
      intcpstor = intcpstor + last - intcpevap;
    }
    s.intcp_stor = intcpstor;
    s.hru_intcpstor = intcpstor * cov;
    hru_intcpevap = intcpevap * cov;
    intcp_changeover = changeover + extra_water;
    PWS_F(intcp_evap, intcpevap);
    PWS_I(intcp_form, intcp_form);
  }
  double net_ppt = net_rain + net_snow;
  const double hru_intcpstor_change = s.hru_intcpstor - hru_intcpstor_old;
  PWS_F(net_ppt, net_ppt); PWS_F(net_rain, net_rain); PWS_F(net_snow, net_snow);
  PWS_F(intcp_changeover, intcp_changeover); PWS_F(intcp_stor, s.intcp_stor);
  PWS_I(intcp_transp_on, s.intcp_transp_on); PWS_F(hru_intcpevap, hru_intcpevap);
  PWS_F(hru_intcpstor, s.hru_intcpstor);
  PWS_F(hru_intcpstor_change, hru_intcpstor_change);
  PWS_F(hru_intcpstor_old, hru_intcpstor_old);
  out.viol(BUD_CANOPY,
           budget_open(0 + hru_rain + hru_snow,
                       (0 + net_rain + net_snow + hru_intcpevap + intcp_changeover) +
                           (0 + hru_intcpstor_change)));
  PWS_OV(net_ppt, net_ppt); PWS_OV(net_rain, net_rain); PWS_OV(net_snow, net_snow);
  PWS_OV(hru_intcpevap, hru_intcpevap); PWS_OV(intcp_changeover, intcp_changeover);

  out.sync();
  // ===================== PRMSSnow ========================================
  double snowmelt = 0.0, snow_evap = 0.0, tcal = 0.0, pk_precip = 0.0;
  double ai = 0.0, frac_swe = 0.0;
  int pptmix_nopack = 0;
  const int newsnow = net_snow > 0.0;
  {
    bool active = PWS_PARAM_I(hru_type) != HRU_LAKE;
    if (active) {
      if (PWS_IN_DOWY == 1) {
        s.pss = 0.0;
        s.iso = 1;
        s.mso = 1;
        s.lso = 0;
      }
      if (PWS_IN_DOY == PWS_LD(P.melt_force + i)) s.iso = 2;
      if (PWS_IN_DOY == PWS_LD(P.melt_look + i)) s.mso = 2;
      if (s.pkwater_equiv < PWS_DNEARZERO && !newsnow) {
        s.snowcov_area = 0.0;
        active = false;
      }
    }
    const double den_max = PWS_LD(P.den_max + i);
    const double denmaxinv = PWS_LD(P.denmaxinv + i);
    const double freeh2o_cap = PWS_LD(P.freeh2o_cap + i);
    if (active) {
      if (newsnow && s.pkwater_equiv < PWS_DNEARZERO) s.snowcov_area = 1.0;
      // ---- step 1: precipitation into the pack, prms_snow.py:1204-1577
      if ((s.pkwater_equiv > 0.0 && net_ppt > 0.0) || net_snow > 0.0) {
        const double tmax_allsnow_c = PWS_LD(P.tmax_allsnow_c + mi);
        double tsnow = tavgc, train;
        if (pptmix == 1) {
          train = (tmaxc + tmax_allsnow_c) * 0.5;
          if (s.pkwater_equiv > 0.0) tsnow = (tminc + tmax_allsnow_c) * 0.5;
          else if (s.pkwater_equiv < 0.0) s.pkwater_equiv = 0.0;
        } else {
          train = tavgc;
          if (train < PWS_CLOSEZERO) train = (tmaxc + tmax_allsnow_c) * 0.5;
        }
        if (train < 0.0) train = 0.0;
        if (tsnow > 0.0) tsnow = 0.0;
        if (s.pkwater_equiv > 0.0) {
          if (net_rain > 0.0) {
            s.pkwater_equiv = s.pkwater_equiv + net_rain;
            pk_precip = pk_precip + net_rain;
            double calpr = 0.0;
            bool do_calin = false;
            if (s.pk_def > 0.0) {
              const double caln = (80.000 + train) * PWS_INCH2CM;
              const double pndz = s.pk_def / caln;
              if (fabs(net_rain - pndz) < PWS_CLOSEZERO) {
                s.pk_def = 0.0;
                s.pk_temp = 0.0;
                s.pk_ice = s.pk_ice + net_rain;
              } else if (net_rain < pndz) {
                s.pk_def = s.pk_def - (caln * net_rain);
                s.pk_temp = -1 * s.pk_def / (s.pkwater_equiv * 1.27);
                s.pk_ice = s.pk_ice + net_rain;
              } else {
                s.pk_def = 0.0;
                s.pk_temp = 0.0;
                s.pk_ice = s.pk_ice + pndz;
                s.freeh2o = net_rain - pndz;
                calpr = train * (net_rain - pndz) * PWS_INCH2CM;
                do_calin = true;
              }
            } else {
              s.freeh2o = s.freeh2o + net_rain;
              calpr = train * net_rain * PWS_INCH2CM;
              do_calin = true;
            }
            if (do_calin) calc_calin(s, snowmelt, calpr, den_max, denmaxinv, freeh2o_cap);
          }
        } else if (net_rain > 0.0) {
          pptmix_nopack = 1;
        }
        if (net_snow > 0.0) {
          s.pkwater_equiv = s.pkwater_equiv + net_snow;
          pk_precip = pk_precip + net_snow;
          s.pk_ice = s.pk_ice + net_snow;
          if (tsnow >= 0.0) {
            s.pk_temp = -1 * s.pk_def / (s.pkwater_equiv * 1.27);
          } else {
            const double calps = tsnow * net_snow * 1.27;
            if (s.freeh2o > 0.0) {
              calc_caloss(s, calps);
            } else {
              s.pk_def = s.pk_def - calps;
              s.pk_temp = -1 * s.pk_def / (s.pkwater_equiv * 1.27);
            }
          }
        }
      }
    }
    out.sync2();
    if (active) {
      if (s.pkwater_equiv > 0.0) {
        // ---- step 2: snow-covered area, prms_snow.py:1851-2128 + :1177-1201
        {
          const double* curve = P.snarea_curve + 11 * (PWS_LD(P.hru_deplcrv + i) - 1);
          const double snowcov_area_ante = s.snowcov_area;
          s.snowcov_area = PWS_LD(curve + 10);
          if (s.pkwater_equiv > s.pst) s.pst = s.pkwater_equiv;
          ai = s.pst;
          const double snarea_thresh = PWS_LD(P.snarea_thresh + i);
          if (ai > snarea_thresh) ai = snarea_thresh;
          if (ai > PWS_DNEARZERO) frac_swe = fmin(1.0, s.pkwater_equiv / ai);
          else frac_swe = 0.0;
          bool from_curve = false;
          if (s.pkwater_equiv >= ai) {
            s.iasw = 0;
          } else if (newsnow) {
            if (s.iasw) {
              s.scrv = s.scrv + (0.75 * net_snow);
            } else {
              s.iasw = 1;
              s.snowcov_areasv = snowcov_area_ante;
              s.pksv = s.pkwater_equiv - net_snow;
              s.scrv = s.pkwater_equiv - (0.25 * net_snow);
            }
          } else if (s.iasw) {
            if (s.pkwater_equiv > s.scrv) {
              // keep the maximum
            } else if (s.pkwater_equiv >= s.pksv) {
              const double difx = s.snowcov_area - s.snowcov_areasv;
              const double dify = s.scrv - s.pksv;
              double fracy = 0.0;
              if (dify > 0.0) fracy = (s.pkwater_equiv - s.pksv) / dify;
              s.snowcov_area = s.snowcov_areasv + fracy * difx;
            } else {
              s.iasw = 0;
              from_curve = true;
            }
          } else {
            from_curve = true;
          }
          if (from_curve) {
            if (frac_swe > 1.0) {
              s.snowcov_area = PWS_LD(curve + 10);
            } else {
              int idx = (int)(10.0 * (frac_swe + 0.2));
              const int jdx = idx - 1;
              if (idx > 11) idx = 11;
              const double dify = (frac_swe * 10.0) - (double)(jdx - 1);
              int a = idx - 1, b = jdx - 1;
              if (a < 0) a += 11;  // python negative index
              if (b < 0) b += 11;
              const double cb = PWS_LD(curve + b);
              const double difx = PWS_LD(curve + a) - cb;
              s.snowcov_area = cb + dify * difx;
            }
          }
        }
        // ---- step 3: albedo, prms_snow.py:2131-2433
        if (!newsnow) {
          if (s.lst) {
            s.slst = s.salb - 3.0;
            if (s.slst < 1.0) s.slst = 1.0;
            if (s.iso != 2 && s.slst > 5.0) s.slst = 5.0;
            s.lst = 0;
            s.snsv = 0.0;
          }
        } else if (s.iso == 2) {
          if (prmx < P.albset_rnm) {
            if (net_snow > P.albset_snm) {
              s.slst = 0.0;
              s.lst = 0;
              s.snsv = 0.0;
            } else {
              s.snsv = s.snsv + net_snow;
              if (s.snsv > P.albset_snm) {
                s.slst = 0.0;
                s.lst = 0;
                s.snsv = 0.0;
              } else {
                if (!s.lst) s.salb = s.slst;
                s.slst = 0.0;
                s.lst = 1;
              }
            }
          }
        } else {
          if (pptmix == 0) {
            s.slst = 0.0;
            s.lst = 0;
          } else if (prmx >= P.albset_rna) {
            s.lst = 0;
          } else if (net_snow >= P.albset_sna) {
            s.slst = 0.0;
            s.lst = 0;
          } else {
            s.slst = s.slst - 3.0;
            if (s.slst < 0.0) s.slst = 0.0;
            if (s.slst > 5.0) s.slst = 5.0;
            s.lst = 0;
          }
          s.snsv = 0.0;
        }
        {
          int ll = (int)(s.slst + 0.5);
          s.slst = s.slst + 1.0;
          if (ll > 0) {
            if (s.int_alb == 2) {
              if (ll > 15) ll = 15;
              s.albedo = k_amlt[ll - 1];
            } else if (ll <= 15) {
              s.albedo = k_acum[ll - 1];
            } else {
              ll = ll - 12;
              if (ll > 15) ll = 15;
              s.albedo = k_amlt[ll - 1];
            }
          } else if (s.iso == 2) {
            s.albedo = 0.72;
            s.int_alb = 2;
          } else {
            s.albedo = 0.91;
            s.int_alb = 1;
          }
        }
      }
    }
    out.sync2();
    if (active) {
      if (s.pkwater_equiv > 0.0) {
        // ---- step 4: night + day energy balance, prms_snow.py:2436-2731
        const double emis_noppt = PWS_LD(P.emis_noppt + i);
        const double esv = hru_ppt > 0.0 ? 1.0 : emis_noppt;
        double cec = PWS_LD(P.cecn_coef + mi) * 0.5;
        if (PWS_PARAM_I(cov_type) > 2) cec = cec * 0.5;
        if (s.iso == 1 && s.mso == 2) {
          if (s.pk_temp >= 0.0) {
            s.lso = s.lso + 1;
            if (s.lso > 4) {
              s.iso = 2;
              s.lso = 0;
            }
          } else {
            s.lso = 0;
          }
        }
        const double canopy_covden = transp_on ? covden_sum : covden_win;
        const double trd = orad_hru / in.horad;
        const int tstorm_mo = PWS_LD(P.tstorm_mo + mi);
        const double swn = swrad * (1.0 - s.albedo) * PWS_LD(P.rad_trncf + i);
        s.pss = s.pss + net_snow;
        const double dpt_before_settle = s.pk_depth + net_snow * PWS_LD(P.deninv + i);
        const double dpt1 = dpt_before_settle +
                            PWS_LD(P.settle_const + i) *
                                ((s.pss * denmaxinv) - dpt_before_settle);
        s.pk_depth = dpt1;
        if (dpt1 > 0.0) s.pk_den = s.pkwater_equiv / dpt1;
        else s.pk_den = 0.0;
        const double effk = 0.0154 * s.pk_den;
        const double cst = s.pk_den * (sqrt(effk * 13751.0));
        // night (niteda 1) then day (niteda 2) through ONE inlined copy of the
        // energy balance (instruction-cache footprint)
#pragma unroll 1
        for (int niteda = 1; niteda <= 2; ++niteda) {
          if (!(s.pkwater_equiv > 0.0)) break;
          const double temp = (niteda == 1 ? tminc : tmaxc);
          const double cal =
              calc_snowbal(s, snowmelt, niteda, cec, cst, esv, niteda == 1 ? 0.0 : swn,
                           (temp + tavgc) * 0.5, trd, canopy_covden, den_max, denmaxinv,
                           emis_noppt, freeh2o_cap, hru_ppt, tstorm_mo);
          tcal = niteda == 1 ? cal : tcal + cal;
        }
        // ---- step 5: sublimation, prms_snow.py:3096-3222
        if (s.pkwater_equiv > 0.0) {
          if (!transp_on || PWS_PARAM_I(cov_type) < 2) {
            const double ez = PWS_LD(P.potet_sublim + i) * potet * s.snowcov_area - hru_intcpevap;
            if (ez < PWS_CLOSEZERO) {
              snow_evap = 0.0;
            } else if (ez >= s.pkwater_equiv) {
              snow_evap = s.pkwater_equiv;
              s.pkwater_equiv = 0.0;
              s.pk_ice = 0.0;
              s.freeh2o = 0.0;
              s.pk_def = 0.0;
              s.pk_temp = 0.0;
            } else {
              s.pk_ice = s.pk_ice - ez;
              if (s.pk_ice < 0.0) {
                s.freeh2o = s.freeh2o + s.pk_ice;
                s.pk_ice = 0.0;
                s.pk_def = 0.0;
                s.pk_temp = 0.0;
              } else {
                const double cal = s.pk_temp * ez * 1.27;
                s.pk_def = s.pk_def + cal;
              }
              s.pkwater_equiv = s.pkwater_equiv - ez;
              snow_evap = ez;
            }
            if (snow_evap < 0.0) {
              s.pkwater_equiv = s.pkwater_equiv - snow_evap;
              if (s.pkwater_equiv < 0.0) s.pkwater_equiv = 0.0;
              snow_evap = 0.0;
            }
            const double avail_et = potet - hru_intcpevap - snow_evap;
            if (avail_et < 0.0) {
              snow_evap = snow_evap + avail_et;
              s.pkwater_equiv = s.pkwater_equiv - avail_et;
              if (snow_evap < 0.0) {
                s.pkwater_equiv = s.pkwater_equiv - snow_evap;
                if (s.pkwater_equiv < 0.0) s.pkwater_equiv = 0.0;
                snow_evap = 0.0;
              }
            }
          }
        } else if (s.pkwater_equiv < 0.0) {
          s.pkwater_equiv = 0.0;
        }
        if (s.pkwater_equiv > 0.0) {
          if (s.pk_den > 0.0) {
            s.pk_depth = s.pkwater_equiv / s.pk_den;
          } else {
            s.pk_den = den_max;
            s.pk_depth = s.pkwater_equiv * denmaxinv;
          }
          s.pss = s.pkwater_equiv;
          if (s.lst > 0) {
            s.snsv = s.snsv - snowmelt;
            if (s.snsv < 0.0) s.snsv = 0.0;
          }
        }
      }
      if (s.pkwater_equiv <= PWS_DNEARZERO) {  // prms_snow.py:1064-1096
        s.pkwater_equiv = 0.0; s.pk_depth = 0.0; s.pss = 0.0; s.snsv = 0.0;
        s.lst = 0; s.pst = 0.0; s.iasw = 0; s.albedo = 0.0; s.pk_den = 0.0;
        s.snowcov_area = 0.0; s.pk_def = 0.0; s.pk_temp = 0.0; s.pk_ice = 0.0;
        s.freeh2o = 0.0; s.snowcov_areasv = 0.0; s.scrv = 0.0; s.pksv = 0.0;
        ai = 0.0;
        frac_swe = 0.0;
      }
    }
  }
  const double freeh2o_change = s.freeh2o - freeh2o_prev;
  const double pk_ice_change = s.pk_ice - pk_ice_prev;
  double through_rain;
  {  // prms_snow.py:1101-1137
    const bool cond1 = net_ppt > 0.0;
    const bool cond2 = pptmix_nopack != 0;
    const bool cond3 = snowmelt < PWS_NEARZERO;
    const bool cond4 = s.pkwater_equiv < PWS_DNEARZERO;
    const bool cond5 = snow_evap < PWS_NEARZERO;
    const bool cond6 = net_snow < PWS_NEARZERO;
    const bool cond7 = snow_evap > (-1 * (pk_ice_change + freeh2o_change));
    through_rain = (cond1 && cond3 && cond4 && cond6) ? net_rain : 0.0;
    if (cond1 && cond3 && cond4 && cond5) through_rain = net_ppt;
    if (cond1 && cond2) through_rain = net_rain;
    if (cond1 && cond6 && cond7) through_rain = 0.0;
  }
  PWS_F(ai, ai); PWS_F(albedo, s.albedo); PWS_F(frac_swe, frac_swe);
  PWS_F(freeh2o, s.freeh2o); PWS_F(freeh2o_change, freeh2o_change);
  PWS_F(freeh2o_prev, freeh2o_prev); PWS_I(iasw, s.iasw);
  PWS_I(int_alb, s.int_alb); PWS_I(iso, s.iso); PWS_I(lso, s.lso);
  PWS_I(lst, s.lst); PWS_I(mso, s.mso); PWS_I(newsnow, newsnow);
  PWS_F(pk_def, s.pk_def); PWS_F(pk_den, s.pk_den); PWS_F(pk_depth, s.pk_depth);
  PWS_F(pk_ice, s.pk_ice); PWS_F(pk_ice_change, pk_ice_change);
  PWS_F(pk_ice_prev, pk_ice_prev); PWS_F(pk_precip, pk_precip);
  PWS_F(pk_temp, s.pk_temp); PWS_F(pksv, s.pksv);
  PWS_F(pkwater_equiv, s.pkwater_equiv); PWS_I(pptmix_nopack, pptmix_nopack);
  PWS_F(pss, s.pss); PWS_F(pst, s.pst); PWS_F(salb, s.salb);
  PWS_F(scrv, s.scrv); PWS_F(slst, s.slst); PWS_F(snow_evap, snow_evap);
  PWS_F(snowcov_area, s.snowcov_area); PWS_F(snowcov_areasv, s.snowcov_areasv);
  PWS_F(snowmelt, snowmelt); PWS_F(snsv, s.snsv); PWS_F(tcal, tcal);
  PWS_F(through_rain, through_rain);
  out.viol(BUD_SNOW, budget_open(0 + net_rain + net_snow,
                                 (0 + snow_evap + snowmelt + through_rain) +
                                     (0 + freeh2o_change + pk_ice_change)));

  mid.potet = potet;
  mid.net_rain = net_rain;
  mid.net_snow = net_snow;
  mid.hru_intcpevap = hru_intcpevap;
  mid.intcp_changeover = intcp_changeover;
  mid.snowmelt = snowmelt;
  mid.snow_evap = snow_evap;
  mid.pkwater_equiv = s.pkwater_equiv;
  mid.snowcov_area = s.snowcov_area;
  mid.through_rain = through_rain;
  mid.transp_on = transp_on;
  mid.pptmix_nopack = pptmix_nopack;
  PWS_OV(snowmelt, mid.snowmelt); PWS_OV(snow_evap, mid.snow_evap);
  PWS_OV(pkwater_equiv, mid.pkwater_equiv); PWS_OV(snowcov_area, mid.snowcov_area);
  PWS_OV(through_rain, mid.through_rain); PWS_OVI(pptmix_nopack, mid.pptmix_nopack);
}

// ---- lower half: PRMSRunoff, PRMSSoilzone, PRMSGroundwater, lateral inflow
template <class State, class Sink>
PWS_DEV void column_lower(const HruParams& P, const int64_t i, State& s,
                          const ColumnMid& mid, Sink& out) {
  // hru_type / cov_type: PWS_PARAM_I at their uses
  const double hru_area = PWS_LD(P.hru_area + i);
  const double hru_in_to_cf = PWS_LD(P.hru_in_to_cf + i);
  // x / hru_area occurs ~20 times a day: one true division for the correctly
  // rounded reciprocal, then the exactly rounded 3-instruction quotient
  const double r_hru_area = 1.0 / hru_area;
#define PER_AREA(x) div_r((x), hru_area, r_hru_area)
  const double potet = mid.potet, net_rain = mid.net_rain, net_snow = mid.net_snow;
  double net_ppt = net_rain + net_snow;
  PWS_OV(net_ppt, net_ppt);
  const double hru_intcpevap = mid.hru_intcpevap, intcp_changeover = mid.intcp_changeover;
  const double snowmelt = mid.snowmelt, snow_evap = mid.snow_evap;
  const double pkwater_equiv = mid.pkwater_equiv, snowcov_area = mid.snowcov_area;
  const double through_rain = mid.through_rain;
#define PWS_MID_FLAG(bit, field) ((PWS_REREAD_MID && Sink::kReread) ? ((PWS_LDVB(mid.flags) >> (bit)) & 1) : mid.field)
  // ===================== PRMSRunoff ======================================
  double soil_rechr_prev = s.soil_rechr;
  double soil_lower_prev = s.soil_moist - s.soil_rechr;  // = yesterday's soil_lower
  PWS_OV(soil_rechr_prev, soil_rechr_prev); PWS_OV(soil_lower_prev, soil_lower_prev);
  const double perv_frac = PWS_LD(P.ro_hru_frac_perv + i);
  const double hruarea_imperv = PWS_LD(P.ro_hru_imperv + i);
  const double hru_impervstor_old = s.hru_impervstor;
  const double dprst_area_max = PWS_LD(P.dprst_area_max + i);
  const bool dprst_on = P.dprst_flag && dprst_area_max > 0.0;
  const double dprst_stor_hru_old =
      dprst_on ? PER_AREA(s.dprst_vol_open + s.dprst_vol_clos) : 0.0;
  double infil = 0.0, srp = 0.0, sri = 0.0, contrib_fraction = 0.0;
  double sroff, hru_sroffp = 0.0, hru_sroffi = 0.0, imperv_evap = 0.0, hru_impervevap = 0.0;
  double dprst_sroff_hru = 0.0, dprst_seep_hru = 0.0, dprst_evap_hru = 0.0;
  double dprst_insroff_hru = 0.0, dprst_stor_hru = 0.0;
  double dprst_vol_open_frac = 0.0, dprst_vol_clos_frac = 0.0, dprst_vol_frac = 0.0;
  double dprst_area_clos = PWS_LD(P.dprst_area_clos_init + i);
  {
    const double soil_moist_prev = soil_lower_prev + soil_rechr_prev;
    const double imperv_frac = hruarea_imperv > 0.0 ? PWS_LD(P.hru_percent_imperv + i) : 0.0;
    double avail_et = potet - snow_evap - hru_intcpevap;
    const double availh2o = intcp_changeover + net_rain;
    const bool land = PWS_PARAM_I(hru_type) == HRU_LAND;
    // compute_infil prms_runoff.py:819-989
    double avail_water = 0.0;
    const double carea_max = PWS_LD(P.carea_max + i);
    const double smidx_coef = PWS_LD(P.smidx_coef + i);
    const double smidx_exp = PWS_LD(P.smidx_exp + i);
    // The reference calls perv_comp / check_capacity from up to three places in
    // sequence (:868-963); the three stages run through one loop body so the
    // pow() and the capacity check are inlined once.
    const double soil_moist_max_ro = PWS_LD(P.soil_moist_max + i);
    const double snowinfil_max = PWS_LD(P.snowinfil_max + i);
#pragma unroll 1
    for (int stage = 0; stage < 3; ++stage) {
      int kind = 0;  // 1 perv_comp(pptp, ptc), 2 check_capacity
      double pptp = 0.0, ptc = 0.0;
      if (stage == 0) {
        if (intcp_changeover > 0.0) {
          avail_water = avail_water + intcp_changeover;
          infil = infil + intcp_changeover;
          kind = land ? 1 : 0;
          pptp = ptc = intcp_changeover;
        }
      } else if (stage == 1) {
        if (PWS_MID_FLAG(1, pptmix_nopack) != 0) {
          avail_water = avail_water + through_rain;
          infil = infil + through_rain;
          kind = land ? 1 : 0;
          pptp = ptc = through_rain;
        }
      } else if (snowmelt > 0.0) {
        avail_water = avail_water + snowmelt;
        infil = infil + snowmelt;
        if (land) {
          if (pkwater_equiv > 0.0 || net_rain < PWS_NEARZERO) {
            kind = 2;
          } else {
            kind = 1;
            pptp = snowmelt;
            ptc = net_ppt;
          }
        }
      } else if (pkwater_equiv < PWS_DNEARZERO) {
        if (net_snow < PWS_NEARZERO && through_rain > 0.0) {
          avail_water = avail_water + through_rain;
          infil = infil + through_rain;
          kind = land ? 1 : 0;
          pptp = ptc = through_rain;
        }
      } else if (infil > 0.0) {
        kind = land ? 2 : 0;
      }
      if (kind == 1)
        perv_comp(soil_moist_prev, carea_max, smidx_coef, smidx_exp, pptp, ptc, infil, srp,
                  contrib_fraction);
      else if (kind == 2)
        check_capacity(soil_moist_prev, soil_moist_max_ro, snowinfil_max, infil, srp);
    }
    out.sync2();
    if (hruarea_imperv > 0.0) {
      s.imperv_stor = s.imperv_stor + avail_water;
      if (land) {
        const double imperv_stor_max = PWS_LD(P.imperv_stor_max + i);
        if (s.imperv_stor > imperv_stor_max) {
          sri = s.imperv_stor - imperv_stor_max;
          s.imperv_stor = imperv_stor_max;
        }
      }
    }
    double runoff = 0.0;
    if (dprst_on) {
      // dprst_comp prms_runoff.py:992-1228 (net_rain argument = availh2o, :716)
      const double open_max = PWS_LD(P.dprst_area_open_max + i);
      const double clos_max = PWS_LD(P.dprst_area_clos_max + i);
      const double vol_open_max = PWS_LD(P.dprst_vol_open_max + i);
      const double vol_clos_max = PWS_LD(P.dprst_vol_clos_max + i);
      double inflow = 0.0;
      if (PWS_MID_FLAG(1, pptmix_nopack)) inflow = inflow + availh2o;
      if (snowmelt > 0.0) {
        inflow = inflow + snowmelt;
      } else if (pkwater_equiv < PWS_DNEARZERO) {
        if (net_snow < PWS_NEARZERO && availh2o > 0.0) inflow = inflow + availh2o;
      }
      if (open_max > 0.0) s.dprst_vol_open = s.dprst_vol_open + inflow * open_max;
      if (clos_max > 0.0) s.dprst_vol_clos = s.dprst_vol_clos + inflow * clos_max;
      double dprst_srp = 0.0, dprst_sri = 0.0;
      if (srp > 0.0) {
        const double tmp = srp * perv_frac * PWS_LD(P.sro_to_dprst_perv + i) * hru_area;
        if (open_max > 0.0) {
          const double v = tmp * PWS_LD(P.dprst_frac_open + i);
          dprst_srp = PER_AREA(v);
          s.dprst_vol_open = s.dprst_vol_open + v;
        }
        if (clos_max > 0.0) {
          const double v = tmp * PWS_LD(P.dprst_frac_clos + i);
          dprst_srp = dprst_srp + PER_AREA(v);
          s.dprst_vol_clos = s.dprst_vol_clos + v;
        }
        srp = srp - div0(dprst_srp, perv_frac);
        if (srp < 0.0) srp = 0.0;
      }
      if (sri > 0.0) {
        const double tmp = sri * imperv_frac * PWS_LD(P.sro_to_dprst_imperv + i) * hru_area;
        if (open_max > 0.0) {
          const double v = tmp * PWS_LD(P.dprst_frac_open + i);
          dprst_sri = PER_AREA(v);
          s.dprst_vol_open = s.dprst_vol_open + v;
        }
        if (clos_max > 0.0) {
          const double v = tmp * PWS_LD(P.dprst_frac_clos + i);
          dprst_sri = dprst_sri + PER_AREA(v);
          s.dprst_vol_clos = s.dprst_vol_clos + v;
        }
        sri = sri - div0(dprst_sri, imperv_frac);
        if (sri < 0.0) sri = 0.0;
      }
      dprst_insroff_hru = dprst_srp + dprst_sri;
      s.dprst_area_open =
          dprst_area(s.dprst_vol_open, vol_open_max, open_max, PWS_LD(P.va_open_exp + i));
      // closed area: recomputed locally only (not returned, prms_runoff.py:1120-1132)
      if (clos_max > 0.0)
        dprst_area_clos =
            dprst_area(s.dprst_vol_clos, vol_clos_max, clos_max, PWS_LD(P.va_clos_exp + i));
      double unsatisfied_et = avail_et;
      const double dprst_avail_et =
          potet * (1.0 - snowcov_area) * PWS_LD(P.dprst_et_coef + i);
      if (dprst_avail_et > 0.0) {
        double evap_open = 0.0, evap_clos = 0.0;
        if (s.dprst_area_open > 0.0) {
          evap_open = fmin(s.dprst_area_open * dprst_avail_et, s.dprst_vol_open);
          if (PER_AREA(evap_open) > unsatisfied_et) evap_open = unsatisfied_et * hru_area;
          if (evap_open > s.dprst_vol_open) evap_open = s.dprst_vol_open;
          unsatisfied_et = unsatisfied_et - PER_AREA(evap_open);
          s.dprst_vol_open = s.dprst_vol_open - evap_open;
        }
        if (dprst_area_clos > 0.0) {
          evap_clos = fmin(dprst_area_clos * dprst_avail_et, s.dprst_vol_clos);
          if (PER_AREA(evap_clos) > unsatisfied_et) evap_clos = unsatisfied_et * hru_area;
          if (evap_clos > s.dprst_vol_clos) evap_clos = s.dprst_vol_clos;
          s.dprst_vol_clos = s.dprst_vol_clos - evap_clos;
        }
        dprst_evap_hru = PER_AREA(evap_open + evap_clos);
      }
      if (s.dprst_vol_open > 0.0) {
        double seep_open = s.dprst_vol_open * PWS_LD(P.dprst_seep_rate_open + i);
        s.dprst_vol_open = s.dprst_vol_open - seep_open;
        if (s.dprst_vol_open < 0.0) {
          seep_open = seep_open + s.dprst_vol_open;
          s.dprst_vol_open = 0.0;
        }
        dprst_seep_hru = PER_AREA(seep_open);
      }
      if (s.dprst_vol_open > 0.0) {
        dprst_sroff_hru = fmax(0.0, s.dprst_vol_open - vol_open_max);
        dprst_sroff_hru =
            dprst_sroff_hru +
            fmax(0.0, (s.dprst_vol_open - dprst_sroff_hru - PWS_LD(P.dprst_vol_thres_open + i)) *
                          PWS_LD(P.dprst_flow_coef + i));
        s.dprst_vol_open = s.dprst_vol_open - dprst_sroff_hru;
        dprst_sroff_hru = PER_AREA(dprst_sroff_hru);
        if (s.dprst_vol_open < 0.0) s.dprst_vol_open = 0.0;
      }
      if (clos_max > 0.0 && dprst_area_clos > PWS_NEARZERO) {
        double seep_clos = s.dprst_vol_clos * PWS_LD(P.dprst_seep_rate_clos + i);
        s.dprst_vol_clos = s.dprst_vol_clos - seep_clos;
        if (s.dprst_vol_clos < 0.0) {
          seep_clos = seep_clos + s.dprst_vol_clos;
          s.dprst_vol_clos = 0.0;
        }
        dprst_seep_hru = dprst_seep_hru + PER_AREA(seep_clos);
      }
      avail_et = avail_et - dprst_evap_hru;
      if (vol_open_max > 0.0) dprst_vol_open_frac = div0(s.dprst_vol_open, vol_open_max);
      if (vol_clos_max > 0.0) dprst_vol_clos_frac = div0(s.dprst_vol_clos, vol_clos_max);
      if (vol_open_max + vol_clos_max > 0.0)
        dprst_vol_frac = div0(s.dprst_vol_open + s.dprst_vol_clos, vol_open_max + vol_clos_max);
      dprst_stor_hru = PER_AREA(s.dprst_vol_open + s.dprst_vol_clos);
      runoff = runoff + dprst_sroff_hru * hru_area;
      // the public dprst_area_clos never changes after init
      dprst_area_clos = PWS_LD(P.dprst_area_clos_init + i);
    }
    out.sync2();
    double srunoff = 0.0;
    if (land) {
      runoff = runoff + srp * PWS_LD(P.ro_hru_perv + i) + sri * hruarea_imperv;
      srunoff = PER_AREA(runoff);
      hru_sroffp = srp * perv_frac;
    }
    if (hruarea_imperv > 0.0) {
      if (s.imperv_stor > 0.0) {
        // imperv_et prms_runoff.py:1278-1287
        if (snowcov_area < 1.0) {
          if (potet < s.imperv_stor) imperv_evap = potet * (1.0 - snowcov_area);
          else imperv_evap = s.imperv_stor * (1.0 - snowcov_area);
          if (imperv_evap * imperv_frac > avail_et) imperv_evap = div0(avail_et, imperv_frac);
          s.imperv_stor = s.imperv_stor - imperv_evap;
        }
        hru_impervevap = imperv_evap * imperv_frac;
        avail_et = avail_et - hru_impervevap;
        if (avail_et < 0.0) {
          hru_impervevap = hru_impervevap + avail_et;
          if (hru_impervevap < 0.0) hru_impervevap = 0.0;
          imperv_evap = imperv_evap / imperv_frac;
          s.imperv_stor = s.imperv_stor - div0(avail_et, imperv_frac);
          avail_et = 0.0;
        }
        s.hru_impervstor = s.imperv_stor * imperv_frac;
      }
      hru_sroffi = sri * imperv_frac;
    }
    sroff = srunoff;
  }
  double infil_hru = infil * perv_frac;
  const double hru_impervstor_change = s.hru_impervstor - hru_impervstor_old;
  const double dprst_stor_hru_change = dprst_stor_hru - dprst_stor_hru_old;
  PWS_F(contrib_fraction, contrib_fraction); PWS_F(infil, infil);
  PWS_F(infil_hru, infil_hru); PWS_F(hru_sroffp, hru_sroffp);
  PWS_F(hru_sroffi, hru_sroffi); PWS_F(imperv_stor, s.imperv_stor);
  PWS_F(imperv_evap, imperv_evap); PWS_F(hru_impervevap, hru_impervevap);
  PWS_F(hru_impervstor, s.hru_impervstor);
  PWS_F(hru_impervstor_old, hru_impervstor_old);
  PWS_F(hru_impervstor_change, hru_impervstor_change);
  PWS_F(dprst_vol_frac, dprst_on ? dprst_vol_frac : 0.0);
  PWS_F(dprst_vol_clos, s.dprst_vol_clos); PWS_F(dprst_vol_open, s.dprst_vol_open);
  PWS_F(dprst_vol_clos_frac, dprst_vol_clos_frac);
  PWS_F(dprst_vol_open_frac, dprst_vol_open_frac);
  PWS_F(dprst_area_clos, dprst_area_clos); PWS_F(dprst_area_open, s.dprst_area_open);
  PWS_F(dprst_area_clos_max, PWS_LD(P.dprst_area_clos_max + i));
  PWS_F(dprst_area_open_max, PWS_LD(P.dprst_area_open_max + i));
  PWS_F(dprst_sroff_hru, dprst_sroff_hru); PWS_F(dprst_seep_hru, dprst_seep_hru);
  PWS_F(dprst_evap_hru, dprst_evap_hru); PWS_F(dprst_insroff_hru, dprst_insroff_hru);
  PWS_F(dprst_stor_hru, dprst_stor_hru); PWS_F(dprst_stor_hru_old, dprst_stor_hru_old);
  PWS_F(dprst_stor_hru_change, dprst_stor_hru_change);
  out.viol(BUD_RUNOFF,
           budget_open(0 + through_rain + snowmelt + intcp_changeover,
                       (0 + hru_sroffi + hru_sroffp + dprst_sroff_hru + infil_hru +
                        hru_impervevap + dprst_seep_hru + dprst_evap_hru) +
                           (0 + hru_impervstor_change + dprst_stor_hru_change)));
  PWS_OV(infil_hru, infil_hru); PWS_OV(sroff, sroff); PWS_OV(dprst_evap_hru, dprst_evap_hru);
  PWS_OV(dprst_seep_hru, dprst_seep_hru); PWS_OV(hru_impervevap, hru_impervevap);

  out.sync();
  // ===================== PRMSSoilzone ====================================
  const double hru_frac_perv = PWS_LD(P.sz_hru_frac_perv + i);
  const double soil_moist_max = PWS_LD(P.sz_soil_moist_max + i);
  const double soil_rechr_max = PWS_LD(P.sz_soil_rechr_max + i);
  const double pref_flow_den = PWS_LD(P.sz_pref_flow_den + i);
  const double pref_flow_max = PWS_LD(P.pref_flow_max + i);
  const double pref_flow_stor_prev = s.pref_flow_stor;
  const double slow_stor_prev = s.slow_stor;
  double soil_to_gw = 0.0, soil_to_ssr = 0.0, ssr_to_gw = 0.0, slow_flow = 0.0;
  double ssres_flow = 0.0, potet_rechr = 0.0, potet_lower = 0.0, perv_actet = 0.0;
  double pref_flow_infil = 0.0, pref_flow_in = 0.0, pref_flow = 0.0;
  double dunnian_flow = 0.0, swale_actet = 0.0, cap_waterin, cap_infil_tot, ssres_stor;
  double hru_actet = hru_impervevap + hru_intcpevap + snow_evap;
  if (P.dprst_flag) hru_actet = hru_actet + dprst_evap_hru;
  {
    const double snow_free = 1.0 - snowcov_area;
    double dunnianflw_pfr = 0.0, dunnianflw_gvr = 0.0, prefflow = 0.0;
    double avail_potet = fmax(0.0, potet - hru_actet);
    double capwater_maxin = div0(infil_hru, hru_frac_perv);
    const double pref_flow_infil_frac = PWS_LD(P.pref_flow_infil_frac + i);
    if (pref_flow_infil_frac != 0.0) {
      if (capwater_maxin > 0.0) {
        double pref_flow_maxin = capwater_maxin * pref_flow_infil_frac;
        capwater_maxin = capwater_maxin - pref_flow_maxin;
        pref_flow_maxin = pref_flow_maxin * hru_frac_perv;
        s.pref_flow_stor = s.pref_flow_stor + pref_flow_maxin;
        dunnianflw_pfr = fmax(0.0, s.pref_flow_stor - pref_flow_max);
        if (dunnianflw_pfr > 0.0) s.pref_flow_stor = pref_flow_max;
        pref_flow_infil = pref_flow_maxin - dunnianflw_pfr;
      }
    }
    cap_infil_tot = capwater_maxin * hru_frac_perv;
    cap_waterin = capwater_maxin;
    if ((capwater_maxin + s.soil_moist) > 0.0) {
      // _compute_soilmoist prms_soilzone.py:1190-1257
      double wat = cap_waterin;
      s.soil_rechr = fmin(s.soil_rechr + wat, soil_rechr_max);
      double excess = s.soil_moist + wat;
      s.soil_moist = fmin(excess, soil_moist_max);
      excess = (excess - soil_moist_max) * hru_frac_perv;
      if (excess > 0.0) {
        const double soil2gw_max = PWS_LD(P.soil2gw_max + i);
        if (soil2gw_max > 0.0) {
          soil_to_gw = fmin(soil2gw_max, excess);
          excess = excess - soil_to_gw;
        }
        if (excess > (wat * hru_frac_perv)) wat = 0.0;
        else wat = wat - div0(excess, hru_frac_perv);
        soil_to_ssr = fmax(0.0, excess);
      }
      cap_waterin = wat * hru_frac_perv;
    }
    out.sync2();
    double topfr = 0.0;
    double availh2o = s.slow_stor + soil_to_ssr;
    if (PWS_PARAM_I(hru_type) == HRU_LAND) {
      topfr = fmax(0.0, availh2o - PWS_LD(P.pref_flow_thrsh + i));
      const double ssresin = soil_to_ssr - topfr;
      s.slow_stor = fmax(0.0, availh2o - topfr);
      if (s.slow_stor > 0.0) {
        if (!compute_interflow(PWS_LD(P.slowcoef_lin + i), PWS_LD(P.slowcoef_sq + i), ssresin,
                               s.slow_stor, slow_flow))
          out.error(1);
      }
    } else if (PWS_PARAM_I(hru_type) == HRU_SWALE) {
      s.slow_stor = availh2o;
    }
    {
      const double ssr2gw_rate = PWS_LD(P.ssr2gw_rate + i);
      if (s.slow_stor > 0.0 && ssr2gw_rate > 0.0) {
        // _compute_gwflow :1315-1320
        const double ssr2gw_exp = PWS_LD(P.ssr2gw_exp + i);
        // pow(x, 1.0) == x exactly; the NHM default exponent is 1
        const double sp = ssr2gw_exp == 1.0 ? s.slow_stor : pow(s.slow_stor, ssr2gw_exp);
        ssr_to_gw = fmax(0.0, ssr2gw_rate * sp);
        ssr_to_gw = fmin(ssr_to_gw, s.slow_stor);
        s.slow_stor = s.slow_stor - ssr_to_gw;
      }
    }
    if (pref_flow_den > 0.0) {
      availh2o = s.pref_flow_stor + topfr;
      dunnianflw_gvr = fmax(0.0, availh2o - pref_flow_max);
      if (dunnianflw_gvr > 0.0) topfr = fmax(0.0, topfr - dunnianflw_gvr);
      pref_flow_in = pref_flow_infil + topfr;
      s.pref_flow_stor = s.pref_flow_stor + topfr;
      if (s.pref_flow_stor > 0.0) {
        if (!compute_interflow(PWS_LD(P.fastcoef_lin + i), PWS_LD(P.fastcoef_sq + i),
                               pref_flow_in, s.pref_flow_stor, prefflow))
          out.error(1);
      }
    } else if (PWS_PARAM_I(hru_type) == HRU_LAND) {
      dunnianflw_gvr = topfr;
    }
    out.sync2();
    if (s.soil_moist > 0.0) {
      // _compute_szactet :1323-1456; et_type 1 default, 2 evap only, 3 evap+transp
      int et_type;
      if (avail_potet < PWS_NEARZERO) {
        et_type = 1;
        avail_potet = 0.0;
      } else if (!PWS_MID_FLAG(0, transp_on)) {
        et_type = snow_free < 0.01 ? 1 : 2;
      } else if (PWS_PARAM_I(cov_type) > 0) {
        et_type = 3;
      } else {
        et_type = snow_free < 0.01 ? 1 : 2;
      }
      if (et_type != 1) {
        const double pcts = div0(s.soil_moist, soil_moist_max);
        const double pctr = div0(s.soil_rechr, soil_rechr_max);
        potet_lower = avail_potet;
        potet_rechr = avail_potet;
        const int soil_type = PWS_LD(P.soil_type + i);
        if (soil_type == 1) {
          if (pcts < 0.25) potet_lower = 0.5 * pcts * avail_potet;
          if (pctr < 0.25) potet_rechr = 0.5 * pctr * avail_potet;
        } else if (soil_type == 2) {
          if (pcts < 0.5) potet_lower = pcts * avail_potet;
          if (pctr < 0.5) potet_rechr = pctr * avail_potet;
        } else if (soil_type == 3) {
          if (pcts < (2.0 / 3.0) && pcts > (1.0 / 3.0)) potet_lower = pcts * avail_potet;
          else if (pcts <= (1.0 / 3.0)) potet_lower = 0.5 * pcts * avail_potet;
          if (pctr < (2.0 / 3.0) && pctr > (1.0 / 3.0)) potet_rechr = pctr * avail_potet;
          else if (pctr <= (1.0 / 3.0)) potet_rechr = 0.5 * pctr * avail_potet;
        }
        if (et_type == 2) potet_rechr = potet_rechr * snow_free;
        if (potet_rechr > s.soil_rechr) {
          potet_rechr = s.soil_rechr;
          s.soil_rechr = 0.0;
        } else {
          s.soil_rechr = s.soil_rechr - potet_rechr;
        }
        double et;
        if (et_type == 2 || potet_rechr >= potet_lower) {
          if (potet_rechr > s.soil_moist) {
            potet_rechr = s.soil_moist;
            s.soil_moist = 0.0;
          } else {
            s.soil_moist = s.soil_moist - potet_rechr;
          }
          et = potet_rechr;
        } else if (potet_lower > s.soil_moist) {
          et = s.soil_moist;
          s.soil_moist = 0.0;
        } else {
          s.soil_moist = s.soil_moist - potet_lower;
          et = potet_lower;
        }
        if (s.soil_rechr > s.soil_moist) s.soil_rechr = s.soil_moist;
        perv_actet = et;
      }
    }
    out.sync2();
    hru_actet = hru_actet + perv_actet * hru_frac_perv;
    if (PWS_PARAM_I(hru_type) == HRU_LAND) {
      dunnian_flow = dunnianflw_gvr + dunnianflw_pfr;
      ssres_flow = slow_flow;
      if (pref_flow_den > 0.0) {
        pref_flow = prefflow;
        ssres_flow = ssres_flow + prefflow;
      }
      sroff = sroff + dunnian_flow;  // prms_soilzone.py:1091, in Runoff's buffer
      ssres_stor = s.slow_stor + s.pref_flow_stor;
    } else {
      const double avail = s.slow_stor - PWS_LD(P.sz_sat_threshold + i);
      if (avail > 0.0) {
        const double unsatisfied_et = potet - hru_actet;
        if (unsatisfied_et > 0.0) {
          swale_actet = fmin(avail, unsatisfied_et);
          hru_actet = hru_actet + swale_actet;
          s.slow_stor = s.slow_stor - swale_actet;
        }
      }
      ssres_stor = s.slow_stor;
    }
  }
  const double soil_lower = s.soil_moist - s.soil_rechr;
  const double soil_lower_max = PWS_LD(P.soil_lower_max + i);
  const double soil_lower_change = soil_lower - soil_lower_prev;
  const double soil_rechr_change = s.soil_rechr - soil_rechr_prev;
  const double slow_stor_change = s.slow_stor - slow_stor_prev;
  const double pref_flow_stor_change = s.pref_flow_stor - pref_flow_stor_prev;
  const double soil_lower_change_hru = soil_lower_change * hru_frac_perv;
  const double soil_rechr_change_hru = soil_rechr_change * hru_frac_perv;
  const double perv_actet_hru = perv_actet * hru_frac_perv;
  double ssres_flow_vol = ssres_flow * hru_in_to_cf;
  double sroff_vol = sroff * hru_in_to_cf;  // prms_soilzone.py:707
  double recharge = soil_to_gw + ssr_to_gw;
  if (P.dprst_flag) recharge = recharge + dprst_seep_hru;
  PWS_F(sroff, sroff); PWS_F(sroff_vol, sroff_vol);
  PWS_F(cap_infil_tot, cap_infil_tot); PWS_F(cap_waterin, cap_waterin);
  PWS_F(dunnian_flow, dunnian_flow); PWS_F(hru_actet, hru_actet);
  PWS_F(perv_actet, perv_actet); PWS_F(perv_actet_hru, perv_actet_hru);
  PWS_F(potet_lower, potet_lower); PWS_F(potet_rechr, potet_rechr);
  PWS_F(pref_flow, pref_flow); PWS_F(pref_flow_in, pref_flow_in);
  PWS_F(pref_flow_infil, pref_flow_infil); PWS_F(pref_flow_max, pref_flow_max);
  PWS_F(pref_flow_stor, s.pref_flow_stor);
  PWS_F(pref_flow_stor_change, pref_flow_stor_change);
  PWS_F(pref_flow_stor_prev, pref_flow_stor_prev);
  PWS_F(pref_flow_thrsh, PWS_LD(P.pref_flow_thrsh + i));
  PWS_F(recharge, recharge); PWS_F(slow_flow, slow_flow);
  PWS_F(slow_stor, s.slow_stor); PWS_F(slow_stor_change, slow_stor_change);
  PWS_F(slow_stor_prev, slow_stor_prev); PWS_F(soil_lower, soil_lower);
  PWS_F(soil_lower_change, soil_lower_change);
  PWS_F(soil_lower_change_hru, soil_lower_change_hru);
  PWS_F(soil_lower_prev, soil_lower_prev);
  PWS_F(soil_lower_ratio, soil_lower_max > 0.0 ? div0(soil_lower, soil_lower_max)
                                                 : PWS_LD(P.soil_lower_ratio_init + i));
  PWS_F(soil_lower_max, soil_lower_max); PWS_F(soil_moist, s.soil_moist);
  PWS_F(soil_moist_tot, ssres_stor + s.soil_moist * hru_frac_perv);
  PWS_F(soil_rechr, s.soil_rechr); PWS_F(soil_rechr_change, soil_rechr_change);
  PWS_F(soil_rechr_change_hru, soil_rechr_change_hru);
  PWS_F(soil_rechr_prev, soil_rechr_prev); PWS_F(soil_to_gw, soil_to_gw);
  PWS_F(soil_to_ssr, soil_to_ssr); PWS_F(soil_zone_max, PWS_LD(P.soil_zone_max + i));
  PWS_F(ssr_to_gw, ssr_to_gw); PWS_F(ssres_flow, ssres_flow);
  PWS_F(ssres_flow_vol, ssres_flow_vol);
  PWS_F(ssres_in, soil_to_ssr + pref_flow_infil + 0.0); PWS_F(ssres_stor, ssres_stor);
  PWS_F(swale_actet, swale_actet); PWS_F(unused_potet, potet - hru_actet);
  out.viol(BUD_SOILZONE,
           budget_open(0 + infil_hru,
                       (0 + perv_actet_hru + soil_to_gw + ssr_to_gw + slow_flow +
                        dunnian_flow + pref_flow) +
                           (0 + soil_rechr_change_hru + soil_lower_change_hru +
                            slow_stor_change + pref_flow_stor_change)));
  PWS_OV(soil_to_gw, soil_to_gw); PWS_OV(ssr_to_gw, ssr_to_gw);

  out.sync();
  // ===================== PRMSGroundwater =================================
  const double gwres_stor_old = s.gwres_stor;
  double gwres_flow, gwres_sink, gwres_stor_change, gwres_flow_vol;
  {
    const double soil_to_gw_vol = soil_to_gw * hru_area;
    const double ssr_to_gw_vol = ssr_to_gw * hru_area;
    const double dprst_seep_hru_vol = dprst_seep_hru * hru_area;
    double stor = s.gwres_stor * hru_area;
    stor += soil_to_gw_vol + ssr_to_gw_vol + dprst_seep_hru_vol;
    const double flow = stor * PWS_LD(P.gwflow_coef + i);
    stor -= flow;
    double sink = stor * PWS_LD(P.gwsink_coef + i);
    if (sink > stor) sink = stor;
    stor -= sink;
    s.gwres_stor = PER_AREA(stor);
    gwres_flow = PER_AREA(flow);
    gwres_sink = PER_AREA(sink);
    gwres_stor_change = s.gwres_stor - gwres_stor_old;
    gwres_flow_vol = gwres_flow * hru_in_to_cf;
  }
  PWS_F(gwres_flow, gwres_flow); PWS_F(gwres_flow_vol, gwres_flow_vol);
  PWS_F(gwres_sink, gwres_sink); PWS_F(gwres_stor, s.gwres_stor);
  PWS_F(gwres_stor_old, gwres_stor_old); PWS_F(gwres_stor_change, gwres_stor_change);
  out.viol(BUD_GW, budget_open(0 + soil_to_gw + ssr_to_gw + dprst_seep_hru,
                               (0 + gwres_flow) + (0 + gwres_stor_change)));

  // ===================== PRMSChannel, HRU side ===========================
  // prms_channel.py:396-421: HRUs without a segment contribute nothing
  PWS_OV(sroff_vol, sroff_vol); PWS_OV(ssres_flow_vol, ssres_flow_vol);
  PWS_OV(gwres_flow_vol, gwres_flow_vol);
  const bool has_seg = PWS_LD(P.hru_segment + i) > 0;
  const double ch_sroff = has_seg ? sroff_vol : 0.0;
  const double ch_ssres = has_seg ? ssres_flow_vol : 0.0;
  const double ch_gwres = has_seg ? gwres_flow_vol : 0.0;
  PWS_F(channel_sroff_vol, ch_sroff); PWS_F(channel_ssres_flow_vol, ch_ssres);
  PWS_F(channel_gwres_flow_vol, ch_gwres);
  out.lateral(div_r(ch_sroff + ch_ssres + ch_gwres, 86400.0, 1.0 / 86400.0));
}

#undef PER_AREA

// One HRU, one day: the whole column (what the reference's Model.calculate
// does for one unit, base/model.py:739-743).
template <class Sink>
PWS_DEV void column_day(const HruParams& P, const int64_t i, HruState& s,
                        const DayIn& in, Sink& out) {
  ColumnMid mid;
  column_upper(P, i, s, in, out, mid);
  out.sync();
  column_lower(P, i, s, mid, out);
}

}  // namespace pws
