/*
 * prms_oracle.c -- CPU restatement (plain C, scalar loops over [nhru] arrays,
 * one process after another per day, exactly the reference's structure) of
 * pywatershed's per-timestep PRMS/NHM process chain.
 *
 * TEST INFRASTRUCTURE ONLY -- see prms_oracle.h.  The product
 * (pywatershed_b200/) never links, includes or imports this file.
 *
 * Parity status: PINNED against the reference's numpy calc_method run in the
 * build container (oracle/gen_golden.py -> tests/golden/ (npz files)).
 *
 * Every function cites the reference file:line it follows; paths are relative
 * to /root/reference/pywatershed/.  Floating point expressions keep the
 * reference's operation order (compile with -ffp-contract=off).
 */
#include "prms_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NDOY 366
#define NMONTH 12
#define PI 3.141592653589793 /* math.pi */

/* constants.py:30-44 */
static const double nearzero = 1.0e-6;
static const double dnearzero = 2.23e-16;
static const double closezero = 1.20e-07;
static const double inch2cm = 2.54;

enum { HRU_INACTIVE = 0, HRU_LAND = 1, HRU_LAKE = 2, HRU_SWALE = 3 };
enum { COV_BARESOIL = 0, COV_GRASSES = 1 };
enum { SOIL_SAND = 1, SOIL_LOAM = 2, SOIL_CLAY = 3 };
enum { SEG_LAKE = 2 };

struct Orc {
  int nhru, nseg, ndeplval;
  char err[256];
  int inited;
  /* ---- parameters ---- */
#define X(n) double *n;
  ORC_HRU_F8_PARAMS(X)
  ORC_MON_F8_PARAMS(X)
  ORC_SEG_F8_PARAMS(X)
#undef X
#define X(n) int32_t *n;
  ORC_HRU_I4_PARAMS(X)
  ORC_MON_I4_PARAMS(X)
  ORC_SEG_I4_PARAMS(X)
#undef X
  double *snarea_curve; /* [ndeplval] */
  double albset_rna, albset_rnm, albset_sna, albset_snm;
  int dprst_flag, temp_units, start_month, start_doy, adjust_parameters;
  /* ---- public variables ---- */
  double *v[ORC_N_HRU_VARS];
  double *s[ORC_N_SEG_VARS];
  /* ---- soltab ---- */
  double *soltab_potsw, *soltab_horad_potsw, *soltab_sunhrs;
  double solar_declination[NDOY], r1[NDOY];
  /* ---- derived / private ---- */
  double *hru_cossl;
  double *transp_check, *tmax_sum; /* prms_atmosphere.py:690-694 */
  double *pptmix_cur;              /* Atmosphere's .current row as edited by Canopy */
  double *tmax_allsnow_c;          /* [12][nhru] prms_snow.py:297 */
  double *deninv, *denmaxinv;
  /* runoff private (prms_runoff.py:123-137, 253-291) */
  double *ro_hru_perv, *ro_hru_frac_perv, *ro_hru_imperv, *dprst_area_max,
      *dprst_frac_clos, *dprst_vol_open_max, *dprst_vol_clos_max,
      *dprst_vol_thres_open, *dprst_in;
  /* soilzone private (prms_soilzone.py:261-536) */
  double *sz_hru_frac_perv, *sz_soil_rechr_max, *sz_soil_moist_max,
      *sz_sat_threshold, *sz_pref_flow_den, *sz_snow_free, *sz_swale_limit;
  int32_t *sz_soil2gw_flag;
  /* channel private (prms_channel.py:189-342) */
  int32_t *ch_order, *ch_tsi;
  double *ch_ts, *ch_c0, *ch_c1, *ch_c2, *ch_seg_inflow, *ch_seg_inflow0,
      *ch_inflow_ts, *ch_outflow_ts, *ch_seg_current_sum, *ch_up;
  int fg_started; /* orc_run_flow_graph has zeroed the node states */
  int32_t *ch_outflow_mask;
  /* budget scratch */
  int64_t itime;
};

static double *dalloc(int64_t n, double fill) {
  double *p = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  for (int64_t i = 0; i < n; ++i) p[i] = fill;
  return p;
}
static int32_t *ialloc(int64_t n, int32_t fill) {
  int32_t *p = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  for (int64_t i = 0; i < n; ++i) p[i] = fill;
  return p;
}

static const char *const hru_var_names[] = {
#define X(name, init) #name,
    ORC_HRU_VARS(X)
#undef X
};
static const char *const seg_var_names[] = {
#define X(name, init) #name,
    ORC_SEG_VARS(X)
#undef X
};

int orc_n_hru_vars(void) { return ORC_N_HRU_VARS; }
int orc_n_seg_vars(void) { return ORC_N_SEG_VARS; }
const char *orc_hru_var_name(int i) {
  return (i >= 0 && i < ORC_N_HRU_VARS) ? hru_var_names[i] : NULL;
}
const char *orc_seg_var_name(int i) {
  return (i >= 0 && i < ORC_N_SEG_VARS) ? seg_var_names[i] : NULL;
}
const char *orc_last_error(Orc *o) { return o->err; }
const double *orc_hru_var(Orc *o, int idx) { return o->v[idx]; }
/* one-step parity (tests/parity.py): continue from the caller's values */
int orc_set_hru_var(Orc *o, int idx, const double *v) {
  if (idx < 0 || idx >= ORC_N_HRU_VARS) return -1;
  memcpy(o->v[idx], v, sizeof(double) * (size_t)o->nhru);
  return 0;
}
const double *orc_seg_var(Orc *o, int idx) { return o->s[idx]; }
const double *orc_soltab_potsw(Orc *o) { return o->soltab_potsw; }
const double *orc_soltab_horad_potsw(Orc *o) { return o->soltab_horad_potsw; }
const double *orc_soltab_sunhrs(Orc *o) { return o->soltab_sunhrs; }

Orc *orc_create(int nhru, int nseg, int ndeplval) {
  Orc *o = (Orc *)calloc(1, sizeof(Orc));
  o->nhru = nhru;
  o->nseg = nseg;
  o->ndeplval = ndeplval;
  o->dprst_flag = 1;
  o->temp_units = 0;
  o->start_month = 1;
  o->start_doy = 1;
  o->adjust_parameters = 1;
#define X(n) o->n = dalloc(nhru, 0.0);
  ORC_HRU_F8_PARAMS(X)
#undef X
#define X(n) o->n = ialloc(nhru, 0);
  ORC_HRU_I4_PARAMS(X)
#undef X
#define X(n) o->n = dalloc((int64_t)NMONTH * nhru, 0.0);
  ORC_MON_F8_PARAMS(X)
#undef X
#define X(n) o->n = ialloc((int64_t)NMONTH * nhru, 0);
  ORC_MON_I4_PARAMS(X)
#undef X
#define X(n) o->n = dalloc(nseg, 0.0);
  ORC_SEG_F8_PARAMS(X)
#undef X
#define X(n) o->n = ialloc(nseg, 0);
  ORC_SEG_I4_PARAMS(X)
#undef X
  o->snarea_curve = dalloc(ndeplval, 0.0);
  {
    int k = 0;
#define X(name, init) o->v[k++] = dalloc(nhru, (double)(init));
    ORC_HRU_VARS(X)
#undef X
    k = 0;
#define X(name, init) o->s[k++] = dalloc(nseg, (double)(init));
    ORC_SEG_VARS(X)
#undef X
  }
  o->soltab_potsw = dalloc((int64_t)NDOY * nhru, 0.0);
  o->soltab_horad_potsw = dalloc((int64_t)NDOY * nhru, 0.0);
  o->soltab_sunhrs = dalloc((int64_t)NDOY * nhru, 0.0);
  o->hru_cossl = dalloc(nhru, 0.0);
  o->transp_check = dalloc(nhru, 0.0);
  o->tmax_sum = dalloc(nhru, 0.0);
  o->pptmix_cur = dalloc(nhru, 0.0);
  o->tmax_allsnow_c = dalloc((int64_t)NMONTH * nhru, 0.0);
  o->deninv = dalloc(nhru, 0.0);
  o->denmaxinv = dalloc(nhru, 0.0);
  o->ro_hru_perv = dalloc(nhru, 0.0);
  o->ro_hru_frac_perv = dalloc(nhru, 0.0);
  o->ro_hru_imperv = dalloc(nhru, 0.0);
  o->dprst_area_max = dalloc(nhru, 0.0);
  o->dprst_frac_clos = dalloc(nhru, 0.0);
  o->dprst_vol_open_max = dalloc(nhru, 0.0);
  o->dprst_vol_clos_max = dalloc(nhru, 0.0);
  o->dprst_vol_thres_open = dalloc(nhru, 0.0);
  o->dprst_in = dalloc(nhru, 0.0);
  o->sz_hru_frac_perv = dalloc(nhru, 0.0);
  o->sz_soil_rechr_max = dalloc(nhru, 0.0);
  o->sz_soil_moist_max = dalloc(nhru, 0.0);
  o->sz_sat_threshold = dalloc(nhru, 0.0);
  o->sz_pref_flow_den = dalloc(nhru, 0.0);
  o->sz_snow_free = dalloc(nhru, 0.0);
  o->sz_swale_limit = dalloc(nhru, 0.0);
  o->sz_soil2gw_flag = ialloc(nhru, 0);
  o->ch_order = ialloc(nseg, 0);
  o->ch_tsi = ialloc(nseg, 1);
  o->ch_ts = dalloc(nseg, 1.0);
  o->ch_c0 = dalloc(nseg, 0.0);
  o->ch_c1 = dalloc(nseg, 0.0);
  o->ch_c2 = dalloc(nseg, 0.0);
  o->ch_seg_inflow = dalloc(nseg, 0.0);
  o->ch_seg_inflow0 = dalloc(nseg, 0.0);
  o->ch_inflow_ts = dalloc(nseg, 0.0);
  o->ch_outflow_ts = dalloc(nseg, 0.0);
  o->ch_seg_current_sum = dalloc(nseg, 0.0);
  o->ch_up = dalloc(nseg, 0.0);
  o->ch_outflow_mask = ialloc(nseg, 0);
  return o;
}

void orc_destroy(Orc *o) {
  if (!o) return;
#define X(n) free(o->n);
  ORC_HRU_F8_PARAMS(X)
  ORC_HRU_I4_PARAMS(X)
  ORC_MON_F8_PARAMS(X)
  ORC_MON_I4_PARAMS(X)
  ORC_SEG_F8_PARAMS(X)
  ORC_SEG_I4_PARAMS(X)
#undef X
  free(o->snarea_curve);
  for (int k = 0; k < ORC_N_HRU_VARS; ++k) free(o->v[k]);
  for (int k = 0; k < ORC_N_SEG_VARS; ++k) free(o->s[k]);
  free(o->soltab_potsw); free(o->soltab_horad_potsw); free(o->soltab_sunhrs);
  free(o->hru_cossl); free(o->transp_check); free(o->tmax_sum);
  free(o->pptmix_cur); free(o->tmax_allsnow_c); free(o->deninv);
  free(o->denmaxinv); free(o->ro_hru_perv); free(o->ro_hru_frac_perv);
  free(o->ro_hru_imperv); free(o->dprst_area_max); free(o->dprst_frac_clos);
  free(o->dprst_vol_open_max); free(o->dprst_vol_clos_max);
  free(o->dprst_vol_thres_open); free(o->dprst_in); free(o->sz_hru_frac_perv);
  free(o->sz_soil_rechr_max); free(o->sz_soil_moist_max);
  free(o->sz_sat_threshold); free(o->sz_pref_flow_den); free(o->sz_snow_free);
  free(o->sz_swale_limit); free(o->sz_soil2gw_flag); free(o->ch_order);
  free(o->ch_tsi); free(o->ch_ts); free(o->ch_c0); free(o->ch_c1);
  free(o->ch_c2); free(o->ch_seg_inflow); free(o->ch_seg_inflow0);
  free(o->ch_inflow_ts); free(o->ch_outflow_ts); free(o->ch_seg_current_sum);
  free(o->ch_up); free(o->ch_outflow_mask);
  free(o);
}

/* n == len copies; n == 1 broadcasts (prms_snow.py:302-320 does that for the
 * scalar-shaped den_init/den_max/settle_const of the shipped domains). */
static int set_f8(double *dst, int64_t len, const double *v, int64_t n) {
  if (n == len) {
    memcpy(dst, v, sizeof(double) * (size_t)len);
    return 0;
  }
  if (n == 1) {
    for (int64_t i = 0; i < len; ++i) dst[i] = v[0];
    return 0;
  }
  return -1;
}
static int set_i4(int32_t *dst, int64_t len, const int32_t *v, int64_t n) {
  if (n == len) {
    memcpy(dst, v, sizeof(int32_t) * (size_t)len);
    return 0;
  }
  if (n == 1) {
    for (int64_t i = 0; i < len; ++i) dst[i] = v[0];
    return 0;
  }
  return -1;
}

int orc_set_f8(Orc *o, const char *name, const double *v, int64_t n) {
  int rc = -2;
#define X(p) \
  if (rc == -2 && !strcmp(name, #p)) rc = set_f8(o->p, o->nhru, v, n);
  ORC_HRU_F8_PARAMS(X)
#undef X
#define X(p) \
  if (rc == -2 && !strcmp(name, #p)) \
    rc = set_f8(o->p, (int64_t)NMONTH * o->nhru, v, n);
  ORC_MON_F8_PARAMS(X)
#undef X
#define X(p) \
  if (rc == -2 && !strcmp(name, #p)) rc = set_f8(o->p, o->nseg, v, n);
  ORC_SEG_F8_PARAMS(X)
#undef X
  if (rc == -2 && !strcmp(name, "snarea_curve"))
    rc = set_f8(o->snarea_curve, o->ndeplval, v, n);
  if (rc == -2 && n == 1) {
    rc = 0;
    if (!strcmp(name, "albset_rna")) o->albset_rna = v[0];
    else if (!strcmp(name, "albset_rnm")) o->albset_rnm = v[0];
    else if (!strcmp(name, "albset_sna")) o->albset_sna = v[0];
    else if (!strcmp(name, "albset_snm")) o->albset_snm = v[0];
    else if (!strcmp(name, "dprst_flag")) o->dprst_flag = (int)v[0];
    else if (!strcmp(name, "temp_units")) o->temp_units = (int)v[0];
    else if (!strcmp(name, "start_month")) o->start_month = (int)v[0];
    else if (!strcmp(name, "start_doy")) o->start_doy = (int)v[0];
    else if (!strcmp(name, "adjust_parameters")) o->adjust_parameters = (int)v[0];
    else rc = -2;
  }
  if (rc) snprintf(o->err, sizeof o->err, "orc_set_f8(%s, n=%lld): %s", name,
                   (long long)n, rc == -2 ? "unknown name" : "bad size");
  return rc ? -1 : 0;
}

int orc_set_i4(Orc *o, const char *name, const int32_t *v, int64_t n) {
  int rc = -2;
#define X(p) \
  if (rc == -2 && !strcmp(name, #p)) rc = set_i4(o->p, o->nhru, v, n);
  ORC_HRU_I4_PARAMS(X)
#undef X
#define X(p) \
  if (rc == -2 && !strcmp(name, #p)) \
    rc = set_i4(o->p, (int64_t)NMONTH * o->nhru, v, n);
  ORC_MON_I4_PARAMS(X)
#undef X
#define X(p) \
  if (rc == -2 && !strcmp(name, #p)) rc = set_i4(o->p, o->nseg, v, n);
  ORC_SEG_I4_PARAMS(X)
#undef X
  if (rc) snprintf(o->err, sizeof o->err, "orc_set_i4(%s, n=%lld): %s", name,
                   (long long)n, rc == -2 ? "unknown name" : "bad size");
  return rc ? -1 : 0;
}

#define VAR(name) (o->v[ORC_V_##name])
#define SEG(name) (o->s[ORC_S_##name])

/* ====================================================================== */
/* PRMSSolarGeometry: atmosphere/solar_constants.py:10-43 and              */
/* atmosphere/prms_solar_geometry.py:176-408                               */
/* ====================================================================== */
static void solar_tables(Orc *o) {
  const double two_pi = 2 * PI;
  const double rad_day = two_pi / 365.242;
  for (int j = 0; j < NDOY; ++j) {
    double jd = (double)(j + 1);
    double obliquity = 1 - (0.01671 * cos((jd - 3) * rad_day));
    double yy = (jd - 1) * rad_day;
    double yy2 = yy * 2, yy3 = yy * 3;
    o->solar_declination[j] = 0.006918 - 0.399912 * cos(yy) +
                              0.070257 * sin(yy) - 0.006758 * cos(yy2) +
                              0.000907 * sin(yy2) - 0.002697 * cos(yy3) +
                              0.00148 * sin(yy3);
    o->r1[j] = (60.0 * 2.0) / (obliquity * obliquity);
  }
}

/* prms_solar_geometry.py:319-358 */
static double compute_t(double lat, double decl) {
  double tx = (-1 * tan(lat)) * tan(decl);
  if (tx < -1.0) return PI;
  if (tx > 1.0) return 0.0;
  return acos(tx);
}

/* prms_solar_geometry.py:361-408 */
static double func3(double v, double w, double x, double y, double r1,
                    double decl) {
  const double pi_12 = 12 / PI;
  return r1 * pi_12 *
         (sin(decl) * sin(w) * (x - y) +
          cos(decl) * cos(w) * (sin(x + v) - sin(y + v)));
}

/* prms_solar_geometry.py:176-316; the aliasing of t3/t7 and t2/t6 at
 * :255-272 is reproduced. */
static void compute_soltab(Orc *o, const double *slopes, const double *aspects,
                           const double *lats, double *solt_out,
                           double *sunh_out) {
  const double two_pi = 2 * PI, pi_12 = 12 / PI;
  const double deg2rad = PI / 180.0; /* np.radians */
  const int nhru = o->nhru;
  for (int i = 0; i < nhru; ++i) {
    double slope = slopes ? slopes[i] : 0.0;
    double aspect = aspects ? aspects[i] : 0.0;
    double sl = atan(slope);
    double sl_sin = sin(sl), sl_cos = cos(sl);
    double aspects_rad = aspect * deg2rad;
    double aspects_cos = cos(aspects_rad);
    double x0 = lats[i] * deg2rad;
    double x0_cos = cos(x0);
    double x1 = asin(sl_cos * sin(x0) + sl_sin * x0_cos * aspects_cos);
    double d1 = sl_cos * x0_cos - sl_sin * sin(x0) * aspects_cos;
    if (fabs(d1) < dnearzero) d1 = dnearzero;
    double x2 = atan(sl_sin * sin(aspects_rad) / d1);
    if (d1 < 0.0) x2 = x2 + PI;
    for (int j = 0; j < NDOY; ++j) {
      double decl = o->solar_declination[j], r1 = o->r1[j];
      double tt = compute_t(x1, decl);
      double t6 = (-1 * tt) - x2;
      double t7 = tt - x2;
      tt = compute_t(x0, decl);
      double t0 = -1 * tt, t1 = tt;
      /* t3 aliases t7, t2 aliases t6 */
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
      double solt = func3(x2, x1, t3, t2, r1, decl);
      double sunh = (t3 - t2) * pi_12;
      if (t7 > t0) {
        solt = func3(x2, x1, t3, t2, r1, decl) + func3(x2, x1, t7, t0, r1, decl);
        sunh = (t3 - t2 + t7 - t0) * pi_12;
      }
      if (t6 < t1) {
        solt = func3(x2, x1, t3, t2, r1, decl) + func3(x2, x1, t1, t6, r1, decl);
        sunh = (t3 - t2 + t1 - t6) * pi_12;
      }
      if (fabs(sl) < dnearzero) {
        solt = func3(0.0, x0, t1, t0, r1, decl);
        sunh = (t1 - t0) * pi_12;
      }
      if (sunh < dnearzero) sunh = 0.0;
      if (solt < 0.0) solt = 0.0;
      solt_out[(int64_t)j * nhru + i] = solt;
      if (sunh_out) sunh_out[(int64_t)j * nhru + i] = sunh;
    }
  }
}

/* ====================================================================== */
/* init: everything the reference derives at construction time            */
/* ====================================================================== */

/* hydrology/prms_runoff.py:253-291 (basin_init) and :293-389 (dprst_init) */
static void runoff_init(Orc *o) {
  const int nhru = o->nhru;
  double *dprst_area_open_max = VAR(dprst_area_open_max);
  double *dprst_area_clos_max = VAR(dprst_area_clos_max);
  double *dprst_vol_open = VAR(dprst_vol_open);
  double *dprst_vol_clos = VAR(dprst_vol_clos);
  double *dprst_area_open = VAR(dprst_area_open);
  double *dprst_area_clos = VAR(dprst_area_clos);
  for (int i = 0; i < nhru; ++i) {
    double harea = o->hru_area[i];
    double perv_area = harea;
    if (o->hru_percent_imperv[i] > 0.0) {
      o->ro_hru_imperv[i] = o->hru_percent_imperv[i] * harea;
      perv_area = perv_area - o->ro_hru_imperv[i];
    }
    if (o->dprst_flag) {
      o->dprst_area_max[i] = o->dprst_frac[i] * harea;
      if (o->dprst_area_max[i] > 0.0) {
        dprst_area_open_max[i] = o->dprst_area_max[i] * o->dprst_frac_open[i];
        o->dprst_frac_clos[i] = 1.0 - o->dprst_frac_open[i];
        dprst_area_clos_max[i] = o->dprst_area_max[i] - dprst_area_open_max[i];
        perv_area = perv_area - o->dprst_area_max[i];
      }
    }
    o->ro_hru_perv[i] = perv_area;
    o->ro_hru_frac_perv[i] = perv_area / harea;
  }
  if (!o->dprst_flag) return;
  for (int i = 0; i < nhru; ++i) {
    if (!(o->dprst_frac[i] > 0.0)) continue;
    o->dprst_vol_clos_max[i] = dprst_area_clos_max[i] * o->dprst_depth_avg[i];
    o->dprst_vol_open_max[i] = dprst_area_open_max[i] * o->dprst_depth_avg[i];
    dprst_vol_open[i] = o->dprst_frac_init[i] * o->dprst_vol_open_max[i];
    dprst_vol_clos[i] = o->dprst_frac_init[i] * o->dprst_vol_clos_max[i];
    o->dprst_vol_thres_open[i] = o->op_flow_thres[i] * o->dprst_vol_open_max[i];
    if (dprst_vol_open[i] > 0.0) {
      double open_vol_r = dprst_vol_open[i] / o->dprst_vol_open_max[i];
      double frac_op_ar;
      if (open_vol_r < nearzero) frac_op_ar = 0.0;
      else if (open_vol_r > 1.0) frac_op_ar = 1.0;
      else frac_op_ar = exp(o->va_open_exp[i] * log(open_vol_r));
      dprst_area_open[i] = dprst_area_open_max[i] * frac_op_ar;
      if (dprst_area_open[i] > dprst_area_open_max[i])
        dprst_area_open[i] = dprst_area_open_max[i];
    }
    if (dprst_vol_clos[i] > 0.0) {
      double clos_vol_r = dprst_vol_clos[i] / o->dprst_vol_clos_max[i];
      double frac_cl_ar;
      if (clos_vol_r < nearzero) frac_cl_ar = 0.0;
      else if (clos_vol_r > 1.0) frac_cl_ar = 1.0;
      else frac_cl_ar = exp(o->va_clos_exp[i] * log(clos_vol_r));
      dprst_area_clos[i] = dprst_area_clos_max[i] * frac_cl_ar;
      if (dprst_area_clos[i] > dprst_area_clos_max[i])
        dprst_area_clos[i] = dprst_area_clos_max[i];
    }
    VAR(dprst_stor_hru)[i] = (dprst_vol_open[i] + dprst_vol_clos[i]) / o->hru_area[i];
    VAR(dprst_stor_hru_old)[i] = VAR(dprst_stor_hru)[i];
    if (o->dprst_vol_open_max[i] + o->dprst_vol_clos_max[i] > 0.0)
      VAR(dprst_vol_frac)[i] = (dprst_vol_open[i] + dprst_vol_clos[i]) /
                               (o->dprst_vol_open_max[i] + o->dprst_vol_clos_max[i]);
  }
}

/* hydrology/prms_soilzone.py:261-536 (_initialize_soilzone_data) */
static void soilzone_init(Orc *o) {
  const int nhru = o->nhru;
  const int adj = o->adjust_parameters;
  for (int i = 0; i < nhru; ++i) {
    int ht = o->hru_type[i];
    double hru_area_imperv = o->hru_percent_imperv[i] * o->hru_area[i];
    double hru_area_perv = o->hru_area[i] - hru_area_imperv;
    double soil_rechr_max = o->soil_rechr_max_frac[i] * o->soil_moist_max[i];
    double soil_moist_max = o->soil_moist_max[i];
    double hru_frac_perv = 1.0 - o->hru_percent_imperv[i];
    if (ht != HRU_INACTIVE) {
      if (o->dprst_flag) {
        double dprst_area_max = o->dprst_frac[i] * o->hru_area[i];
        hru_area_perv = hru_area_perv - dprst_area_max;
      }
      hru_frac_perv = hru_area_perv / o->hru_area[i];
    }
    double sat_threshold = o->sat_threshold[i];
    if (ht == HRU_INACTIVE || ht == HRU_LAKE) sat_threshold = 0.0;
    double pref_flow_den = o->pref_flow_den[i];
    if (ht != HRU_LAND) pref_flow_den = 0.0;
    double soil_moist = o->soil_moist_init_frac[i] * soil_moist_max;
    double soil_rechr = o->soil_rechr_init_frac[i] * soil_rechr_max;
    double ssres_stor = o->ssstor_init_frac[i] * sat_threshold;
    if (ht == HRU_INACTIVE || ht == HRU_LAKE) ssres_stor = 0.0;
    if (adj) { /* :361-463 parameter range fixes ("warn" behaviour) */
      if (soil_moist_max < 1.0e-5) soil_moist_max = 1.0e-5;
      if (soil_rechr_max < 1.0e-5) soil_rechr_max = 1.0e-5;
      if (soil_rechr_max > soil_moist_max) soil_rechr_max = soil_moist_max;
      if (soil_rechr > soil_rechr_max) soil_rechr = soil_rechr_max;
      if (soil_moist > soil_moist_max) soil_moist = soil_moist_max;
      if (soil_rechr > soil_moist) soil_rechr = soil_moist;
      if (ssres_stor > sat_threshold) ssres_stor = sat_threshold;
    }
    double pref_flow_thrsh = 0.0, pref_flow_max = 0.0;
    o->sz_swale_limit[i] = 0.0;
    if (ht == HRU_SWALE) {
      o->sz_swale_limit[i] = 3.0 * sat_threshold;
      pref_flow_thrsh = sat_threshold;
    }
    if (ht == HRU_LAND) {
      pref_flow_thrsh = sat_threshold * (1.0 - pref_flow_den);
      pref_flow_max = sat_threshold - pref_flow_thrsh;
    }
    if (ht == HRU_LAND || ht == HRU_SWALE) {
      VAR(slow_stor)[i] = fmin(ssres_stor, pref_flow_thrsh);
      VAR(pref_flow_stor)[i] = ssres_stor - VAR(slow_stor)[i];
    }
    o->sz_soil2gw_flag[i] = o->soil2gw_max[i] > 0.0;
    o->sz_hru_frac_perv[i] = hru_frac_perv;
    o->sz_soil_rechr_max[i] = soil_rechr_max;
    o->sz_soil_moist_max[i] = soil_moist_max;
    o->sz_sat_threshold[i] = sat_threshold;
    o->sz_pref_flow_den[i] = pref_flow_den;
    VAR(pref_flow_thrsh)[i] = pref_flow_thrsh;
    VAR(pref_flow_max)[i] = pref_flow_max;
    VAR(soil_moist)[i] = soil_moist;
    VAR(soil_rechr)[i] = soil_rechr;
    VAR(ssres_stor)[i] = ssres_stor;
    VAR(soil_zone_max)[i] = sat_threshold + soil_moist_max * hru_frac_perv;
    VAR(soil_moist_tot)[i] = ssres_stor + soil_moist * hru_frac_perv;
    VAR(soil_lower)[i] = soil_moist - soil_rechr;
    VAR(soil_lower_max)[i] = soil_moist_max - soil_rechr_max;
    if (VAR(soil_lower_max)[i] > 0.0)
      VAR(soil_lower_ratio)[i] = VAR(soil_lower)[i] / VAR(soil_lower_max)[i];
  }
}

/* Kahn topological order: outlets that appear in no edge first, then a valid
 * topological order (hydrology/prms_channel.py:213-231 uses networkx; any
 * valid order gives identical sums for in-degree <= 2). */
static int channel_order(Orc *o) {
  const int nseg = o->nseg;
  int32_t *indeg = ialloc(nseg, 0);
  int32_t *inedge = ialloc(nseg, 0);
  int n = 0;
  for (int i = 0; i < nseg; ++i) {
    int to = o->tosegment[i] - 1;
    if (to >= 0) {
      indeg[to]++;
      inedge[to] = 1;
      inedge[i] = 1;
    }
  }
  for (int i = 0; i < nseg; ++i)
    if (!inedge[i]) o->ch_order[n++] = i; /* isolated outlets first */
  int head = n;
  for (int i = 0; i < nseg; ++i)
    if (inedge[i] && indeg[i] == 0) o->ch_order[n++] = i;
  while (head < n) {
    int i = o->ch_order[head++];
    int to = o->tosegment[i] - 1;
    if (to >= 0 && --indeg[to] == 0) o->ch_order[n++] = to;
  }
  free(indeg);
  free(inedge);
  if (n != nseg) {
    snprintf(o->err, sizeof o->err, "segment graph has a cycle");
    return -1;
  }
  return 0;
}

/* hydrology/prms_channel.py:186-342 */
static int channel_init(Orc *o) {
  const int nseg = o->nseg;
  if (nseg == 0) return 0;
  if (channel_order(o)) return -1;
  for (int i = 0; i < nseg; ++i) {
    SEG(seg_outflow)[i] = o->segment_flow_init[i];
    o->ch_outflow_mask[i] = (o->tosegment[i] - 1) < 0;
    /* velocity from the UNADJUSTED slope (:234-242) */
    double velocity = ((1.0 / o->mann_n[i]) * sqrt(o->seg_slope[i]) *
                       pow(o->seg_depth[i], 2.0 / 3.0)) * 60.0 * 60.0;
    double K = 24.0;
    if (velocity > 0.0) K = o->seg_length[i] / velocity;
    if (o->segment_type[i] == SEG_LAKE) K = 24.0;
    if (K < 0.01) K = 0.01;
    if (K > 24.0) K = 24.0;
    double ts = 1.0;
    int tsi = 1;
    if (K < 1.0) tsi = -1;
    else if (K < 2.0) { ts = 1.0; tsi = 1; }
    else if (K < 3.0) { ts = 2.0; tsi = 2; }
    else if (K < 4.0) { ts = 3.0; tsi = 3; }
    else if (K < 6.0) { ts = 4.0; tsi = 4; }
    else if (K < 8.0) { ts = 6.0; tsi = 6; }
    else if (K < 12.0) { ts = 8.0; tsi = 8; }
    else if (K < 24.0) { ts = 12.0; tsi = 12; }
    else { ts = 24.0; tsi = 24; }
    o->ch_ts[i] = ts;
    o->ch_tsi[i] = tsi;
    double x = o->x_coef[i];
    double d = K - (K * x) + (0.5 * ts);
    if (fabs(d) < 1e-6) d = 0.0001;
    double c0 = (-(K * x) + (0.5 * ts)) / d;
    double c1 = ((K * x) + (0.5 * ts)) / d;
    double c2 = (K - (K * x) - (0.5 * ts)) / d;
    if (c2 < 0.0) { c1 += c2; c2 = 0.0; }
    if (c0 < 0.0) { c1 += c0; c0 = 0.0; }
    o->ch_c0[i] = c0; o->ch_c1[i] = c1; o->ch_c2[i] = c2;
    o->ch_seg_inflow[i] = 0.0;
    o->ch_seg_inflow0[i] = NAN;
    o->ch_inflow_ts[i] = 0.0;
    o->ch_outflow_ts[i] = 0.0;
    o->ch_seg_current_sum[i] = 0.0;
  }
  for (int i = 0; i < nseg; ++i) { /* :336-340 assignment, not sum */
    int j = o->tosegment[i] - 1;
    if (j < 0) continue;
    o->ch_seg_inflow[j] = SEG(seg_outflow)[i];
  }
  return 0;
}

int orc_init(Orc *o) {
  const int nhru = o->nhru;
  /* reset public variables to get_init_values() */
  {
    int k = 0;
#define X(name, init) \
  for (int i = 0; i < nhru; ++i) o->v[k][i] = (double)(init); \
  ++k;
    ORC_HRU_VARS(X)
#undef X
    k = 0;
#define X(name, init) \
  for (int i = 0; i < o->nseg; ++i) o->s[k][i] = (double)(init); \
  ++k;
    ORC_SEG_VARS(X)
#undef X
  }
  solar_tables(o);
  /* prms_solar_geometry.py:138-161 */
  compute_soltab(o, NULL, NULL, o->hru_lat, o->soltab_horad_potsw, NULL);
  compute_soltab(o, o->hru_slope, o->hru_aspect, o->hru_lat, o->soltab_potsw,
                 o->soltab_sunhrs);
  for (int i = 0; i < nhru; ++i) {
    o->hru_cossl[i] = cos(atan(o->hru_slope[i]));
    /* prms_atmosphere.py:690-716 transp init */
    o->transp_check[i] = 0.0;
    o->tmax_sum[i] = 0.0;
    VAR(transp_on)[i] = 0.0;
    int beg = o->transp_beg[i], end = o->transp_end[i];
    int motmp = o->start_month + NMONTH;
    if (o->start_month == beg) {
      if (o->start_doy > 10) VAR(transp_on)[i] = 1.0;
      else o->transp_check[i] = 1.0;
    } else if (end > beg) {
      if (o->start_month > beg && o->start_month < end) VAR(transp_on)[i] = 1.0;
    } else {
      if (o->start_month > beg || motmp < end + NMONTH) VAR(transp_on)[i] = 1.0;
    }
    /* prms_snow.py:295-431 with snowpack_init == 0 */
    for (int m = 0; m < NMONTH; ++m)
      o->tmax_allsnow_c[(int64_t)m * nhru + i] =
          (o->tmax_allsnow[(int64_t)m * nhru + i] - 32.0) / 1.8;
    o->deninv[i] = 1.0 / o->den_init[i];
    o->denmaxinv[i] = 1.0 / o->den_max[i];
    if (o->snowpack_init[i] > 0.0) {
      snprintf(o->err, sizeof o->err,
               "snowpack_init > 0 is not runnable in the reference "
               "(prms_snow.py:413 calls a missing method)");
      return -1;
    }
    VAR(pkwater_equiv)[i] = o->snowpack_init[i];
    VAR(pss)[i] = VAR(pkwater_equiv)[i];
    VAR(pst)[i] = VAR(pkwater_equiv)[i];
    /* prms_groundwater.py:145-149 */
    VAR(gwres_stor)[i] = o->gwstor_init[i];
    VAR(gwres_stor_old)[i] = o->gwstor_init[i];
  }
  runoff_init(o);
  soilzone_init(o);
  if (channel_init(o)) return -1;
  o->itime = 0;
  o->inited = 1;
  return 0;
}

/* ====================================================================== */
/* PRMSAtmosphere, one day (atmosphere/prms_atmosphere.py:287-765)        */
/* ====================================================================== */
static const double solf[26] = {0.20,  0.35,  0.45,  0.51,  0.56,  0.59, 0.62,
                                0.64,  0.655, 0.67,  0.682, 0.69,  0.70, 0.71,
                                0.715, 0.72,  0.722, 0.724, 0.726, 0.728, 0.73,
                                0.734, 0.738, 0.742, 0.746, 0.75};

static void atmosphere_day(Orc *o, int first_day, int doy, int month, int dom,
                           const double *prcp, const double *tmax,
                           const double *tmin) {
  const int nhru = o->nhru;
  const int64_t m0 = (int64_t)(month - 1) * nhru;
  const int is_summer = (doy >= 79) && (doy <= 265);
  const double *potsw = o->soltab_potsw + (int64_t)(doy - 1) * nhru;
  const double *horad = o->soltab_horad_potsw + (int64_t)(doy - 1) * nhru;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nhru; ++i) {
    /* adjust_temperature :287-317 */
    double tmaxf = tmax[i] + o->tmax_cbh_adj[m0 + i];
    double tminf = tmin[i] + o->tmin_cbh_adj[m0 + i];
    double tminc = (tminf - 32.0) * (5.0 / 9.0);
    double tmaxc = (tmaxf - 32.0) * (5.0 / 9.0);
    double tavgc = (tmaxc + tminc) / 2.0;
    /* adjust_precip :319-430 */
    double allsnow = o->tmax_allsnow[m0 + i];
    double allrain = allsnow + o->tmax_allrain_offset[m0 + i];
    double tdiff = tmaxf - tminf;
    if (tdiff < nearzero) tdiff = 1.0e-4;
    double prmx = ((tmaxf - allsnow) / tdiff) * o->adjmix_rain[m0 + i];
    if (prmx < 0.0) prmx = 0.0;
    if (prmx > 1.0) prmx = 1.0;
    if (tminf > allsnow || tmaxf >= allrain) prmx = 1.0;
    if (tmaxf <= allsnow) prmx = 0.0;
    if (prcp[i] <= 0.0) prmx = 0.0;
    double hru_ppt = prcp[i] * o->snow_cbh_adj[m0 + i];
    double hru_rain = prmx * hru_ppt;
    double hru_snow = hru_ppt - hru_rain;
    if (prmx <= 0.0) {
      hru_ppt = prcp[i] * o->snow_cbh_adj[m0 + i];
      hru_snow = hru_ppt;
      hru_rain = 0.0;
    }
    if (prmx >= 1.0) {
      hru_ppt = prcp[i] * o->rain_cbh_adj[m0 + i];
      hru_rain = hru_ppt;
      hru_snow = 0.0;
    }
    int pptmix = (hru_ppt > 0.0) && (tmaxf > allsnow) &&
                 ((tminf <= allsnow) && (tmaxf < allrain)) && (prmx < 1.0);
    /* _ddsolrad_run :475-620 */
    double dday = (o->dday_slope[m0 + i] * tmaxf) + o->dday_intcp[m0 + i] + 1.0;
    if (dday < 1.0) dday = 1.0;
    double radmax = o->radmax[m0 + i];
    double radadj = radmax;
    if (dday < 26.0) {
      int kp = (int)dday;
      radadj = solf[kp - 1] + ((solf[kp] - solf[kp - 1]) * (dday - kp));
      if (radadj > radmax) radadj = radmax;
    }
    double pptadj = 1.0;
    if (hru_ppt > o->ppt_rad_adj[m0 + i]) {
      double tmax_index = o->tmax_index[m0 + i];
      pptadj = o->radadj_intcp[m0 + i] +
               o->radadj_slope[m0 + i] * (tmaxf - tmax_index);
      if (pptadj > 1.0) pptadj = 1.0;
      if (tmaxf < tmax_index) {
        double allrain2 = o->tmax_allrain_offset[m0 + i] + allsnow;
        pptadj = o->radj_sppt[i];
        if (tmaxf < allrain2) pptadj = o->radj_wppt[i];
        if (tmaxf >= allrain2 && !is_summer) pptadj = o->radj_wppt[i];
      }
    }
    radadj = radadj * pptadj;
    if (radadj < 0.2) radadj = 0.2;
    double swrad = potsw[i] * radadj / o->hru_cossl[i];
    double orad_hru = radadj * horad[i];
    /* _potet_jh_run :643-668 */
    double tavgf = (tavgc * 9 / 5) + 32;
    double elh = (597.3 - (0.5653 * tavgc)) * inch2cm;
    double potet = o->jh_coef[m0 + i] * (tavgf - o->jh_coef_hru[i]) * swrad / elh;
    if (potet < 0.0) potet = 0.0;
    /* calculate_transp_tindex :680-765 (RUN part, one day) */
    double transp_tmax_f = o->temp_units == 0
                               ? o->transp_tmax[i]
                               : (o->transp_tmax[i] * (9.0 / 5.0)) + 32.0;
    double transp_on = VAR(transp_on)[i]; /* carried from yesterday/init */
    (void)first_day;
    if (dom == 1) {
      if (month == o->transp_end[i]) {
        transp_on = 0.0;
        o->transp_check[i] = 0.0;
        o->tmax_sum[i] = 0.0;
      }
      if (month == o->transp_beg[i]) {
        o->transp_check[i] = 1.0;
        o->tmax_sum[i] = 0.0;
      }
    }
    if (o->transp_check[i] == 1.0) {
      if (tmaxf > 32.0) o->tmax_sum[i] = o->tmax_sum[i] + tmaxf;
      if (o->tmax_sum[i] > transp_tmax_f) {
        transp_on = 1.0;
        o->transp_check[i] = 0.0;
        o->tmax_sum[i] = 0.0;
      }
    }
    VAR(tmaxf)[i] = tmaxf; VAR(tminf)[i] = tminf; VAR(prmx)[i] = prmx;
    VAR(hru_ppt)[i] = hru_ppt; VAR(hru_rain)[i] = hru_rain;
    VAR(hru_snow)[i] = hru_snow; VAR(swrad)[i] = swrad; VAR(potet)[i] = potet;
    VAR(transp_on)[i] = transp_on; VAR(tmaxc)[i] = tmaxc; VAR(tavgc)[i] = tavgc;
    VAR(tminc)[i] = tminc; VAR(pptmix)[i] = (double)pptmix;
    VAR(orad_hru)[i] = orad_hru;
    o->pptmix_cur[i] = (double)pptmix;
  }
}

/* ====================================================================== */
/* PRMSCanopy (hydrology/prms_canopy.py:285-581)                          */
/* ====================================================================== */
static void intercept(double precip, double stor_max, double cov,
                      double *intcp_stor, double *net_precip) {
  /* prms_canopy.py:575-581 */
  *net_precip = precip * (1.0 - cov);
  *intcp_stor = *intcp_stor + precip;
  if (*intcp_stor > stor_max) {
    *net_precip = *net_precip + (*intcp_stor - stor_max) * cov;
    *intcp_stor = stor_max;
  }
}

static void canopy_day(Orc *o) {
  const int nhru = o->nhru;
  const double *pk_ice_prev = VAR(pk_ice_prev), *freeh2o_prev = VAR(freeh2o_prev);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nhru; ++i) {
    const double hru_rain = VAR(hru_rain)[i], hru_snow = VAR(hru_snow)[i];
    const double hru_ppt = VAR(hru_ppt)[i], potet = VAR(potet)[i];
    const int transp_on = (int)VAR(transp_on)[i];
    const int cov_type = o->cov_type[i];
    const int hru_type = HRU_LAND; /* prms_canopy.py:92 forces LAND */
    double netrain = hru_rain, netsnow = hru_snow;
    double cov, stor_max_rain;
    if (transp_on == 1) {
      cov = o->covden_sum[i];
      stor_max_rain = o->srain_intcp[i];
    } else {
      cov = o->covden_win[i];
      stor_max_rain = o->wrain_intcp[i];
    }
    int intcp_form = 0; /* RAIN */
    if (hru_snow > 0.0) intcp_form = 1; /* SNOW */
    double intcpstor = VAR(intcp_stor)[i];
    double intcpevap = 0.0, changeover = 0.0, extra_water = 0.0;
    if (hru_type == HRU_LAKE || cov_type == COV_BARESOIL) {
      if (cov_type == COV_BARESOIL && intcpstor > 0.0)
        extra_water = VAR(intcp_stor)[i];
      intcpstor = 0.0;
    }
    int intcp_transp_on = (int)VAR(intcp_transp_on)[i];
    if (transp_on == 0 && intcp_transp_on == 1) {
      intcp_transp_on = 0;
      if (intcpstor > 0.0) {
        double diff = o->covden_sum[i] - cov;
        changeover = intcpstor * diff;
        if (cov > 0.0) {
          if (changeover < 0.0) {
            intcpstor = intcpstor * o->covden_sum[i] / cov;
            changeover = 0.0;
          }
        } else {
          intcpstor = 0.0;
        }
      }
    } else if (transp_on == 1 && intcp_transp_on == 0) {
      intcp_transp_on = 1;
      if (intcpstor > 0.0) {
        double diff = o->covden_win[i] - cov;
        changeover = intcpstor * diff;
        if (cov > 0.0) {
          if (changeover < 0.0) {
            intcpstor = intcpstor * o->covden_win[i] / cov;
            changeover = 0.0;
          }
        } else {
          intcpstor = 0.0;
        }
      }
    }
    if (hru_type != HRU_LAKE && cov_type != COV_BARESOIL) {
      if (hru_rain > 0.0) {
        if (cov > 0.0) {
          if (cov_type > COV_GRASSES) {
            intercept(hru_rain, stor_max_rain, cov, &intcpstor, &netrain);
          } else if (cov_type == COV_GRASSES) {
            if ((pk_ice_prev[i] + freeh2o_prev[i]) < dnearzero &&
                netsnow < nearzero)
              intercept(hru_rain, stor_max_rain, cov, &intcpstor, &netrain);
          }
        }
      }
    }
    if (hru_snow > 0.0) {
      if (cov > 0.0) {
        if (cov_type > COV_GRASSES) {
          intercept(hru_snow, o->snow_intcp[i], cov, &intcpstor, &netsnow);
          if (netsnow < nearzero) {
            netrain = netrain + netsnow;
            netsnow = 0.0;
            o->pptmix_cur[i] = 0.0; /* :500-505 writes Atmosphere's buffer */
          }
        }
      }
    }
    if (intcpstor > 0.0) {
      if (hru_ppt < nearzero) {
        double epan_coef = 1.0;
        double evrn = potet / epan_coef;
        double evsn = potet * o->potet_sublim[i];
        if (intcp_form == 1) {
          double z = intcpstor - evsn;
          if (z > 0) {
            intcpstor = z;
            intcpevap = evsn;
          } else {
            intcpevap = intcpstor;
            intcpstor = 0.0;
          }
        } else {
          double d = intcpstor - evrn;
          if (d > 0.0) {
            intcpstor = d;
            intcpevap = evrn;
          } else {
            intcpevap = intcpstor;
            intcpstor = 0.0;
          }
        }
      }
    }
    if (intcpevap * cov > potet) {
      double last = intcpevap;
      if (cov > 0.0) intcpevap = potet / cov;
      else intcpevap = 0.0;
      intcpstor = intcpstor + last - intcpevap;
    }
    VAR(intcp_evap)[i] = intcpevap;
    VAR(intcp_stor)[i] = intcpstor;
    VAR(net_rain)[i] = netrain;
    VAR(net_snow)[i] = netsnow;
    VAR(net_ppt)[i] = netrain + netsnow;
    VAR(hru_intcpstor)[i] = intcpstor * cov;
    VAR(hru_intcpevap)[i] = intcpevap * cov;
    VAR(intcp_changeover)[i] = changeover + extra_water;
    VAR(intcp_form)[i] = (double)intcp_form;
    VAR(intcp_transp_on)[i] = (double)intcp_transp_on;
    /* :348-350 */
    VAR(hru_intcpstor_change)[i] = VAR(hru_intcpstor)[i] - VAR(hru_intcpstor_old)[i];
  }
}

/* ====================================================================== */
/* PRMSSnow (hydrology/prms_snow.py:623-3222)                             */
/* ====================================================================== */
static const double acum_init[15] = {0.80, 0.77, 0.75, 0.72, 0.70, 0.69, 0.68, 0.67,
                                     0.66, 0.65, 0.64, 0.63, 0.62, 0.61, 0.60};
static const double amlt_init[15] = {0.72, 0.65, 0.60, 0.58, 0.56, 0.54, 0.52, 0.50,
                                     0.48, 0.46, 0.44, 0.43, 0.42, 0.41, 0.40};
#define MAXALB 15
#define ONETHIRD (1.0 / 3.0)

/* one HRU's snow scalars, read/written through this struct */
typedef struct {
  double freeh2o, pk_def, pk_den, pk_depth, pk_ice, pk_precip, pk_temp,
      pkwater_equiv, pss, pst, snowmelt;
  int iasw, pptmix_nopack;
} SnowPack;

/* prms_snow.py:1580-1767 */
static void calc_calin(SnowPack *s, double cal, double den_max, double denmaxinv,
                       double freeh2o_cap, double snowcov_area) {
  double dif = cal - s->pk_def;
  if (dif < 0.0) {
    s->pk_def = s->pk_def - cal;
    s->pk_temp = -1 * s->pk_def / (s->pkwater_equiv * 1.27);
  } else if (dif > 0.0) {
    double pmlt = dif / 203.2;
    double apmlt = pmlt * snowcov_area;
    s->pk_def = 0.0;
    s->pk_temp = 0.0;
    double apk_ice;
    if (snowcov_area > 0.0) apk_ice = s->pk_ice / snowcov_area;
    else apk_ice = 0.0;
    if (pmlt > apk_ice) {
      s->snowmelt = s->snowmelt + s->pkwater_equiv;
      s->pkwater_equiv = 0.0;
      s->iasw = 0;
      s->pk_def = 0.0;
      s->pk_temp = 0.0;
      s->pk_ice = 0.0;
      s->freeh2o = 0.0;
      s->pk_depth = 0.0;
      s->pss = 0.0;
      s->pst = 0.0;
      s->pk_den = 0.0;
    } else {
      s->pk_ice = s->pk_ice - apmlt;
      s->freeh2o = s->freeh2o + apmlt;
      double pwcap = freeh2o_cap * s->pk_ice;
      double dif_water = s->freeh2o - pwcap;
      if (dif_water > 0.0) {
        if (dif_water > s->pkwater_equiv) dif_water = s->pkwater_equiv;
        s->pkwater_equiv = s->pkwater_equiv - dif_water;
        s->freeh2o = pwcap;
        if (s->pk_den > 0.0) {
          s->pk_depth = s->pkwater_equiv / s->pk_den;
        } else {
          s->pk_den = den_max;
          s->pk_depth = s->pkwater_equiv * denmaxinv;
        }
        s->snowmelt = s->snowmelt + dif_water;
        s->pss = s->pkwater_equiv;
      }
    }
  } else {
    s->pk_temp = 0.0;
    s->pk_def = 0.0;
  }
  if (!(s->pkwater_equiv > 0.0)) s->pk_den = 0.0;
}

/* prms_snow.py:1770-1848 */
static void calc_caloss(SnowPack *s, double cal) {
  if (s->freeh2o < closezero) {
    s->pk_def = s->pk_def - cal;
  } else {
    double calnd = s->freeh2o * 203.2;
    double dif = cal + calnd;
    if (dif > 0.0) {
      s->pk_ice = s->pk_ice - (cal / 203.2);
      s->freeh2o = s->freeh2o + (cal / 203.2);
      return;
    } else {
      if (dif < 0.0) s->pk_def = -dif;
      s->pk_ice = s->pk_ice + s->freeh2o;
      s->freeh2o = 0.0;
    }
  }
  if (s->pkwater_equiv > 0.0)
    s->pk_temp = -1 * s->pk_def / (s->pkwater_equiv * 1.27);
  else if (s->pkwater_equiv < 0.0)
    s->pkwater_equiv = 0.0;
}

/* prms_snow.py:1204-1577 */
static void calc_ppt_to_pack(SnowPack *s, double den_max, double denmaxinv,
                             double freeh2o_cap, double net_ppt, double net_rain,
                             double net_snow, int pptmix, double snowcov_area,
                             double tavgc, double tmax_allsnow_c, double tmaxc,
                             double tminc) {
  int ppt_through = !((s->pkwater_equiv > 0.0 && net_ppt > 0.0) || net_snow > 0.0);
  if (ppt_through) return;
  double tsnow = tavgc, train;
  if (pptmix == 1) {
    train = (tmaxc + tmax_allsnow_c) * 0.5;
    if (s->pkwater_equiv > 0.0) tsnow = (tminc + tmax_allsnow_c) * 0.5;
    else if (s->pkwater_equiv < 0.0) s->pkwater_equiv = 0.0;
  } else {
    train = tavgc;
    if (train < closezero) train = (tmaxc + tmax_allsnow_c) * 0.5;
  }
  if (train < 0.0) train = 0.0;
  if (tsnow > 0.0) tsnow = 0.0;
  if (s->pkwater_equiv > 0.0) {
    if (net_rain > 0.0) {
      s->pkwater_equiv = s->pkwater_equiv + net_rain;
      s->pk_precip = s->pk_precip + net_rain;
      if (s->pk_def > 0.0) {
        double caln = (80.000 + train) * inch2cm;
        double pndz = s->pk_def / caln;
        if (fabs(net_rain - pndz) < closezero) {
          s->pk_def = 0.0;
          s->pk_temp = 0.0;
          s->pk_ice = s->pk_ice + net_rain;
        } else if (net_rain < pndz) {
          s->pk_def = s->pk_def - (caln * net_rain);
          s->pk_temp = -1 * s->pk_def / (s->pkwater_equiv * 1.27);
          s->pk_ice = s->pk_ice + net_rain;
        } else {
          s->pk_def = 0.0;
          s->pk_temp = 0.0;
          s->pk_ice = s->pk_ice + pndz;
          s->freeh2o = net_rain - pndz;
          double calpr = train * (net_rain - pndz) * inch2cm;
          calc_calin(s, calpr, den_max, denmaxinv, freeh2o_cap, snowcov_area);
        }
      } else {
        s->freeh2o = s->freeh2o + net_rain;
        double calpr = train * net_rain * inch2cm;
        calc_calin(s, calpr, den_max, denmaxinv, freeh2o_cap, snowcov_area);
      }
    }
  } else if (net_rain > 0.0) {
    s->pptmix_nopack = 1;
  }
  if (net_snow > 0.0) {
    s->pkwater_equiv = s->pkwater_equiv + net_snow;
    s->pk_precip = s->pk_precip + net_snow;
    s->pk_ice = s->pk_ice + net_snow;
    if (tsnow >= 0.0) {
      s->pk_temp = -1 * s->pk_def / (s->pkwater_equiv * 1.27);
    } else {
      double calps = tsnow * net_snow * 1.27;
      if (s->freeh2o > 0.0) {
        calc_caloss(s, calps);
      } else {
        s->pk_def = s->pk_def - calps;
        s->pk_temp = -1 * s->pk_def / (s->pkwater_equiv * 1.27);
      }
    }
  }
}

/* prms_snow.py:1177-1201; python negative indices wrap */
static double sca_deplcrv(const double *c, double frac_swe) {
  if (frac_swe > 1.0) return c[10];
  int idx = (int)(10.0 * (frac_swe + 0.2));
  int jdx = idx - 1;
  if (idx > 11) idx = 11;
  double dify = (frac_swe * 10.0) - (double)(jdx - 1);
  int a = idx - 1, b = jdx - 1;
  if (a < 0) a += 11;
  if (b < 0) b += 11;
  double difx = c[a] - c[b];
  return c[b] + dify * difx;
}

/* prms_snow.py:1851-2128 */
static void calc_snowcov(double *ai, double *frac_swe, int *iasw, double *pksv,
                         double *pst, double *scrv, double *snowcov_area,
                         double *snowcov_areasv, double net_snow, int newsnow,
                         double pkwater_equiv, const double *curve,
                         double snarea_thresh) {
  double snowcov_area_ante = *snowcov_area;
  *snowcov_area = curve[11 - 1];
  if (pkwater_equiv > *pst) *pst = pkwater_equiv;
  *ai = *pst;
  if (*ai > snarea_thresh) *ai = snarea_thresh;
  if (*ai > dnearzero) {
    *frac_swe = pkwater_equiv / *ai;
    *frac_swe = fmin(1.0, *frac_swe);
  } else {
    *frac_swe = 0.0;
  }
  if (pkwater_equiv >= *ai) {
    *iasw = 0;
  } else {
    if (newsnow) {
      if (*iasw) {
        *scrv = *scrv + (0.75 * net_snow);
      } else {
        *iasw = 1;
        *snowcov_areasv = snowcov_area_ante;
        *pksv = pkwater_equiv - net_snow;
        *scrv = pkwater_equiv - (0.25 * net_snow);
      }
      return;
    } else if (*iasw) {
      if (pkwater_equiv > *scrv) return;
      if (pkwater_equiv >= *pksv) {
        double difx = *snowcov_area - *snowcov_areasv;
        double dify = *scrv - *pksv;
        double fracy = 0.0;
        if (dify > 0.0) fracy = (pkwater_equiv - *pksv) / dify;
        *snowcov_area = *snowcov_areasv + fracy * difx;
        return;
      } else {
        *iasw = 0;
      }
    }
    *snowcov_area = sca_deplcrv(curve, *frac_swe);
  }
}

/* prms_snow.py:2131-2433 */
static void calc_snalbedo(const Orc *o, double *albedo, int *int_alb, int iso,
                          int *lst, double net_snow, int newsnow, int pptmix,
                          double prmx, double *salb, double *slst, double *snsv) {
  if (!newsnow) {
    if (*lst) {
      *slst = *salb - 3.0;
      if (*slst < 1.0) *slst = 1.0;
      if (iso != 2) {
        if (*slst > 5.0) *slst = 5.0;
      }
      *lst = 0;
      *snsv = 0.0;
    }
  } else if (iso == 2) {
    if (prmx < o->albset_rnm) {
      if (net_snow > o->albset_snm) {
        *slst = 0.0;
        *lst = 0;
        *snsv = 0.0;
      } else {
        *snsv = *snsv + net_snow;
        if (*snsv > o->albset_snm) {
          *slst = 0.0;
          *lst = 0;
          *snsv = 0.0;
        } else {
          if (!*lst) *salb = *slst;
          *slst = 0.0;
          *lst = 1;
        }
      }
    }
  } else {
    if (pptmix == 0) {
      *slst = 0.0;
      *lst = 0;
    } else if (prmx >= o->albset_rna) {
      *lst = 0;
    } else if (net_snow >= o->albset_sna) {
      *slst = 0.0;
      *lst = 0;
    } else {
      *slst = *slst - 3.0;
      if (*slst < 0.0) *slst = 0.0;
      if (*slst > 5.0) *slst = 5.0;
      *lst = 0;
    }
    *snsv = 0.0;
  }
  int ll = (int)(*slst + 0.5);
  *slst = *slst + 1.0;
  if (ll > 0) {
    if (*int_alb == 2) {
      if (ll > MAXALB) ll = MAXALB;
      *albedo = amlt_init[ll - 1];
    } else if (ll <= MAXALB) {
      *albedo = acum_init[ll - 1];
    } else {
      ll = ll - 12;
      if (ll > MAXALB) ll = MAXALB;
      *albedo = amlt_init[ll - 1];
    }
  } else if (iso == 2) {
    *albedo = 0.72;
    *int_alb = 2;
  } else {
    *albedo = 0.91;
    *int_alb = 1;
  }
}

/* x**4 rounded once (error-free products via fma).  prms_snow.py:2768 writes
 * `(temp + 273.16) ** 4.0`, which numpy hands to the platform libm's pow; a
 * < 1 ulp pow returns the correctly rounded power except when the exact value
 * sits within a few hundredths of an ulp of a rounding boundary (glibc 2.39:
 * 0.086 % of arguments in [230, 310] K differ, by one ulp).  The oracle uses
 * the correctly rounded value so that it does not depend on the libm build;
 * the golden-vector tests pin the result against the reference run. */
static double pow4(double x) {
  const double t = x * x;
  const double e = fma(x, x, -t);
  const double p = t * t;
  const double pe = fma(t, t, -p);
  return p + (pe + 2.0 * t * e);
}

/* prms_snow.py:2734-3093; returns cal */
static double calc_snowbal(SnowPack *s, int niteda, double cec, double cst,
                           double esv, double sw, double temp, double trd,
                           double canopy_covden, double den_max,
                           double denmaxinv, double emis_noppt,
                           double freeh2o_cap, double hru_ppt,
                           double snowcov_area, int tstorm_mo) {
  double air = 0.585e-7 * pow4(temp + 273.16);
  double emis = esv;
  double ts, sno;
  if (temp < 0.0) {
    ts = temp;
    sno = air;
  } else {
    ts = 0.0;
    sno = 325.7;
  }
  if (hru_ppt > 0.0) {
    if (tstorm_mo == 1) {
      if (niteda == 1) {
        emis = 0.85;
        if (trd > ONETHIRD) emis = emis_noppt;
      } else {
        if (trd > ONETHIRD) emis = 1.29 - (0.882 * trd);
        if (trd >= 0.5) emis = 0.95 - (0.2 * trd);
      }
    }
  }
  double sky = (1.0 - canopy_covden) * ((emis * air) - sno);
  double can = canopy_covden * (air - sno);
  double cecsub = 0.0;
  if (temp > 0.0 && hru_ppt > 0.0) cecsub = cec * temp;
  double cal = sky + can + cecsub + sw;
  if (ts >= 0.0 && cal > 0.0) {
    calc_calin(s, cal, den_max, denmaxinv, freeh2o_cap, snowcov_area);
    return cal;
  }
  double qcond = cst * (ts - s->pk_temp);
  if (qcond < 0.0) {
    if (s->pk_temp < 0.0) {
      s->pk_def = s->pk_def - qcond;
      s->pk_temp = -1 * s->pk_def / (s->pkwater_equiv * 1.27);
    } else {
      calc_caloss(s, qcond);
    }
  } else if (qcond < closezero) {
    if (s->pk_temp >= 0.0) {
      if (cal > 0.0)
        calc_calin(s, cal, den_max, denmaxinv, freeh2o_cap, snowcov_area);
    }
  } else if (ts >= 0.0) {
    double pk_defsub = s->pk_def - qcond;
    if (pk_defsub < 0.0) {
      s->pk_def = 0.0;
      s->pk_temp = 0.0;
    } else {
      s->pk_def = pk_defsub;
      s->pk_temp = -pk_defsub / (s->pkwater_equiv * 1.27);
    }
  } else {
    double pkt = -ts * (s->pkwater_equiv * 1.27);
    double pks = s->pk_def - pkt;
    double pk_defsub = pks - qcond;
    if (pk_defsub < 0.0) {
      s->pk_def = pkt;
      s->pk_temp = ts;
    } else {
      s->pk_def = pk_defsub + pkt;
      s->pk_temp = -1 * s->pk_def / (s->pkwater_equiv * 1.27);
    }
  }
  return cal;
}

/* prms_snow.py:3096-3222 */
static void calc_snowevap(SnowPack *s, double *snow_evap, double hru_intcpevap,
                          double potet, double potet_sublim,
                          double snowcov_area) {
  double ez = potet_sublim * potet * snowcov_area - hru_intcpevap;
  if (ez < closezero) {
    *snow_evap = 0.0;
  } else if (ez >= s->pkwater_equiv) {
    *snow_evap = s->pkwater_equiv;
    s->pkwater_equiv = 0.0;
    s->pk_ice = 0.0;
    s->freeh2o = 0.0;
    s->pk_def = 0.0;
    s->pk_temp = 0.0;
  } else {
    s->pk_ice = s->pk_ice - ez;
    if (s->pk_ice < 0.0) {
      s->freeh2o = s->freeh2o + s->pk_ice;
      s->pk_ice = 0.0;
      s->pk_def = 0.0;
      s->pk_temp = 0.0;
    } else {
      double cal = s->pk_temp * ez * 1.27;
      s->pk_def = s->pk_def + cal;
    }
    s->pkwater_equiv = s->pkwater_equiv - ez;
    *snow_evap = ez;
  }
  if (*snow_evap < 0.0) {
    s->pkwater_equiv = s->pkwater_equiv - *snow_evap;
    if (s->pkwater_equiv < 0.0) s->pkwater_equiv = 0.0;
    *snow_evap = 0.0;
  }
  double avail_et = potet - hru_intcpevap - *snow_evap;
  if (avail_et < 0.0) {
    *snow_evap = *snow_evap + avail_et;
    s->pkwater_equiv = s->pkwater_equiv - avail_et;
    if (*snow_evap < 0.0) {
      s->pkwater_equiv = s->pkwater_equiv - *snow_evap;
      if (s->pkwater_equiv < 0.0) s->pkwater_equiv = 0.0;
      *snow_evap = 0.0;
    }
  }
}

/* prms_snow.py:623-1174 */
static void snow_day(Orc *o, int doy, int dowy, int month) {
  const int nhru = o->nhru;
  const int64_t m0 = (int64_t)(month - 1) * nhru;
  const double *horad = o->soltab_horad_potsw + (int64_t)(doy - 1) * nhru;
#pragma omp parallel for schedule(static)
  for (int jj = 0; jj < nhru; ++jj) {
    const double net_snow = VAR(net_snow)[jj], net_rain = VAR(net_rain)[jj];
    const double net_ppt = VAR(net_ppt)[jj];
    const int transp_on = (int)VAR(transp_on)[jj];
    const int pptmix = (int)o->pptmix_cur[jj];
    const double canopy_covden = transp_on ? o->covden_sum[jj] : o->covden_win[jj];
    const int newsnow = net_snow > 0.0;
    const double trd = VAR(orad_hru)[jj] / horad[jj];
    VAR(newsnow)[jj] = (double)newsnow;
    VAR(frac_swe)[jj] = 0.0;
    VAR(pk_precip)[jj] = 0.0;
    VAR(snowmelt)[jj] = 0.0;
    VAR(snow_evap)[jj] = 0.0;
    VAR(tcal)[jj] = 0.0;
    VAR(ai)[jj] = 0.0;
    int skip = 0;
    if (o->hru_type[jj] == HRU_LAKE) skip = 1;
    if (!skip) {
      if (dowy == 1) {
        VAR(pss)[jj] = 0.0;
        VAR(iso)[jj] = 1;
        VAR(mso)[jj] = 1;
        VAR(lso)[jj] = 0;
      }
      VAR(pptmix_nopack)[jj] = 0;
      if (doy == o->melt_force[jj]) VAR(iso)[jj] = 2;
      if (doy == o->melt_look[jj]) VAR(mso)[jj] = 2;
      if (VAR(pkwater_equiv)[jj] < dnearzero && newsnow == 0) {
        VAR(snowcov_area)[jj] = 0.0;
        skip = 1;
      }
    }
    if (!skip) {
      if (newsnow && VAR(pkwater_equiv)[jj] < dnearzero) VAR(snowcov_area)[jj] = 1.0;
      SnowPack s;
      s.freeh2o = VAR(freeh2o)[jj]; s.pk_def = VAR(pk_def)[jj];
      s.pk_den = VAR(pk_den)[jj]; s.pk_depth = VAR(pk_depth)[jj];
      s.pk_ice = VAR(pk_ice)[jj]; s.pk_precip = VAR(pk_precip)[jj];
      s.pk_temp = VAR(pk_temp)[jj]; s.pkwater_equiv = VAR(pkwater_equiv)[jj];
      s.pss = VAR(pss)[jj]; s.pst = VAR(pst)[jj]; s.snowmelt = VAR(snowmelt)[jj];
      s.iasw = (int)VAR(iasw)[jj]; s.pptmix_nopack = (int)VAR(pptmix_nopack)[jj];
      double snowcov_area = VAR(snowcov_area)[jj];
      double ai = VAR(ai)[jj], frac_swe = VAR(frac_swe)[jj];
      double pksv = VAR(pksv)[jj], scrv = VAR(scrv)[jj];
      double snowcov_areasv = VAR(snowcov_areasv)[jj];
      double albedo = VAR(albedo)[jj], salb = VAR(salb)[jj], slst = VAR(slst)[jj];
      double snsv = VAR(snsv)[jj], tcal = VAR(tcal)[jj], snow_evap = 0.0;
      int int_alb = (int)VAR(int_alb)[jj], lst = (int)VAR(lst)[jj];
      int iso = (int)VAR(iso)[jj], lso = (int)VAR(lso)[jj], mso = (int)VAR(mso)[jj];
      const double den_max = o->den_max[jj], denmaxinv = o->denmaxinv[jj];
      const double freeh2o_cap = o->freeh2o_cap[jj];
      const double tavgc = VAR(tavgc)[jj], tmaxc = VAR(tmaxc)[jj], tminc = VAR(tminc)[jj];
      const double hru_ppt = VAR(hru_ppt)[jj];

      /* step 1 */
      calc_ppt_to_pack(&s, den_max, denmaxinv, freeh2o_cap, net_ppt, net_rain,
                       net_snow, pptmix, snowcov_area, tavgc,
                       o->tmax_allsnow_c[m0 + jj], tmaxc, tminc);
      if (s.pkwater_equiv > 0.0) {
        /* step 2 */
        const double *curve = o->snarea_curve + 11 * (o->hru_deplcrv[jj] - 1);
        calc_snowcov(&ai, &frac_swe, &s.iasw, &pksv, &s.pst, &scrv,
                     &snowcov_area, &snowcov_areasv, net_snow, newsnow,
                     s.pkwater_equiv, curve, o->snarea_thresh[jj]);
        /* step 3 */
        calc_snalbedo(o, &albedo, &int_alb, iso, &lst, net_snow, newsnow,
                      pptmix, VAR(prmx)[jj], &salb, &slst, &snsv);
      }
      if (s.pkwater_equiv > 0.0) {
        /* step 4: prms_snow.py:2436-2731 */
        double emis = o->emis_noppt[jj];
        if (hru_ppt > 0.0) emis = 1.0;
        double esv = emis;
        double cec = o->cecn_coef[m0 + jj] * 0.5;
        if (o->cov_type[jj] > 2) cec = cec * 0.5;
        if (iso == 1) {
          if (mso == 2) {
            if (s.pk_temp >= 0.0) {
              lso = lso + 1;
              if (lso > 4) {
                iso = 2;
                lso = 0;
              }
            } else {
              lso = 0;
            }
          }
        }
        double temp = (tminc + tavgc) * 0.5;
        double swn = 0.0, cst = 0.0;
        const int tstorm = o->tstorm_mo[m0 + jj];
        if (s.pkwater_equiv > 0.0) {
          swn = VAR(swrad)[jj] * (1.0 - albedo) * o->rad_trncf[jj];
          s.pss = s.pss + net_snow;
          double dpt_before_settle = s.pk_depth + net_snow * o->deninv[jj];
          double dpt1 = dpt_before_settle +
                        o->settle_const[jj] * ((s.pss * denmaxinv) - dpt_before_settle);
          s.pk_depth = dpt1;
          if (dpt1 > 0.0) s.pk_den = s.pkwater_equiv / dpt1;
          else s.pk_den = 0.0;
          double effk = 0.0154 * s.pk_den;
          cst = s.pk_den * (sqrt(effk * 13751.0));
          tcal = calc_snowbal(&s, 1, cec, cst, esv, 0.0, temp, trd, canopy_covden,
                              den_max, denmaxinv, o->emis_noppt[jj], freeh2o_cap,
                              hru_ppt, snowcov_area, tstorm);
        }
        temp = (tmaxc + tavgc) * 0.5;
        if (s.pkwater_equiv > 0.0) {
          double cals = calc_snowbal(&s, 2, cec, cst, esv, swn, temp, trd,
                                     canopy_covden, den_max, denmaxinv,
                                     o->emis_noppt[jj], freeh2o_cap, hru_ppt,
                                     snowcov_area, tstorm);
          tcal = tcal + cals;
        }
        /* step 5 */
        if (s.pkwater_equiv > 0.0) {
          if (!transp_on || (transp_on && o->cov_type[jj] < 2))
            calc_snowevap(&s, &snow_evap, VAR(hru_intcpevap)[jj], VAR(potet)[jj],
                          o->potet_sublim[jj], snowcov_area);
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
          if (lst > 0) {
            snsv = snsv - s.snowmelt;
            if (snsv < 0.0) snsv = 0.0;
          }
        }
      }
      if (s.pkwater_equiv <= dnearzero) {
        s.pkwater_equiv = 0.0;
        s.pk_depth = 0.0; s.pss = 0.0; snsv = 0.0; lst = 0; s.pst = 0.0;
        s.iasw = 0; albedo = 0.0; s.pk_den = 0.0; snowcov_area = 0.0;
        s.pk_def = 0.0; s.pk_temp = 0.0; s.pk_ice = 0.0; s.freeh2o = 0.0;
        snowcov_areasv = 0.0; ai = 0.0; frac_swe = 0.0; scrv = 0.0; pksv = 0.0;
      }
      VAR(freeh2o)[jj] = s.freeh2o; VAR(pk_def)[jj] = s.pk_def;
      VAR(pk_den)[jj] = s.pk_den; VAR(pk_depth)[jj] = s.pk_depth;
      VAR(pk_ice)[jj] = s.pk_ice; VAR(pk_precip)[jj] = s.pk_precip;
      VAR(pk_temp)[jj] = s.pk_temp; VAR(pkwater_equiv)[jj] = s.pkwater_equiv;
      VAR(pss)[jj] = s.pss; VAR(pst)[jj] = s.pst; VAR(snowmelt)[jj] = s.snowmelt;
      VAR(iasw)[jj] = s.iasw; VAR(pptmix_nopack)[jj] = s.pptmix_nopack;
      VAR(snowcov_area)[jj] = snowcov_area; VAR(ai)[jj] = ai;
      VAR(frac_swe)[jj] = frac_swe; VAR(pksv)[jj] = pksv; VAR(scrv)[jj] = scrv;
      VAR(snowcov_areasv)[jj] = snowcov_areasv; VAR(albedo)[jj] = albedo;
      VAR(salb)[jj] = salb; VAR(slst)[jj] = slst; VAR(snsv)[jj] = snsv;
      VAR(tcal)[jj] = tcal; VAR(snow_evap)[jj] = snow_evap;
      VAR(int_alb)[jj] = int_alb; VAR(lst)[jj] = lst; VAR(iso)[jj] = iso;
      VAR(lso)[jj] = lso; VAR(mso)[jj] = mso;
    }
    /* :1098-1137 vectorised tail */
    VAR(freeh2o_change)[jj] = VAR(freeh2o)[jj] - VAR(freeh2o_prev)[jj];
    VAR(pk_ice_change)[jj] = VAR(pk_ice)[jj] - VAR(pk_ice_prev)[jj];
    int cond1 = net_ppt > 0.0;
    int cond2 = VAR(pptmix_nopack)[jj] != 0;
    int cond3 = VAR(snowmelt)[jj] < nearzero;
    int cond4 = VAR(pkwater_equiv)[jj] < dnearzero;
    int cond5 = VAR(snow_evap)[jj] < nearzero;
    int cond6 = net_snow < nearzero;
    int cond7 = VAR(snow_evap)[jj] >
                (-1 * (VAR(pk_ice_change)[jj] + VAR(freeh2o_change)[jj]));
    double through_rain = (cond1 && cond3 && cond4 && cond6) ? net_rain : 0.0;
    if (cond1 && cond3 && cond4 && cond5) through_rain = net_ppt;
    if (cond1 && cond2) through_rain = net_rain;
    if (cond1 && cond6 && cond7) through_rain = 0.0;
    VAR(through_rain)[jj] = through_rain;
  }
}

/* ====================================================================== */
/* PRMSRunoff (hydrology/prms_runoff.py:433-1287)                         */
/* ====================================================================== */

/* :1231-1259 */
static void perv_comp(double soil_moist_prev, double carea_max, double smidx_coef,
                      double smidx_exp, double pptp, double ptc, double *infil,
                      double *srp, double *contrib_fraction) {
  double smidx = soil_moist_prev + 0.5 * ptc;
  double ca_fraction;
  if (smidx > 25.0) ca_fraction = carea_max;
  else ca_fraction = smidx_coef * pow(10.0, smidx_exp * smidx);
  if (ca_fraction > carea_max) ca_fraction = carea_max;
  double srpp = ca_fraction * pptp;
  *infil = *infil - srpp;
  *srp = *srp + srpp;
  *contrib_fraction = ca_fraction;
}

/* :1262-1275 */
static void check_capacity(double soil_moist_prev, double soil_moist_max,
                           double snowinfil_max, double *infil, double *srp) {
  double capacity = soil_moist_max - soil_moist_prev;
  double excess = *infil - capacity;
  if (excess > snowinfil_max) {
    *srp = *srp + excess - snowinfil_max;
    *infil = snowinfil_max + capacity;
  }
}

static void runoff_day(Orc *o) {
  const int nhru = o->nhru;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nhru; ++i) {
    const double soil_moist_prev = VAR(soil_lower_prev)[i] + VAR(soil_rechr_prev)[i];
    const double hruarea = o->hru_area[i];
    const double perv_area = o->ro_hru_perv[i], perv_frac = o->ro_hru_frac_perv[i];
    const double hruarea_imperv = o->ro_hru_imperv[i];
    const int hru_type = o->hru_type[i];
    const double potet = VAR(potet)[i], net_rain = VAR(net_rain)[i];
    const double net_ppt = VAR(net_ppt)[i], net_snow = VAR(net_snow)[i];
    const double snowmelt = VAR(snowmelt)[i], pkwater_equiv = VAR(pkwater_equiv)[i];
    const double through_rain = VAR(through_rain)[i];
    const double intcp_changeover = VAR(intcp_changeover)[i];
    const int pptmix_nopack = (int)VAR(pptmix_nopack)[i];
    const double snowcov_area = VAR(snowcov_area)[i];
    double runoff = 0.0, srp = 0.0, sri = 0.0;
    double infil = 0.0; /* :618 infil[:] = 0 */
    VAR(hru_sroffp)[i] = 0.0;
    double contrib_fraction = 0.0;
    double imperv_frac = 0.0;
    if (hruarea_imperv > 0.0) {
      imperv_frac = o->hru_percent_imperv[i];
      VAR(hru_sroffi)[i] = 0.0;
      VAR(imperv_evap)[i] = 0.0;
      VAR(hru_impervevap)[i] = 0.0;
    }
    double avail_et = potet - VAR(snow_evap)[i] - VAR(hru_intcpevap)[i];
    double availh2o = intcp_changeover + net_rain;
    double imperv_stor = VAR(imperv_stor)[i];
    /* compute_infil :819-989 */
    {
      int hru_flag = (hru_type == HRU_LAND);
      double avail_water = 0.0;
      if (intcp_changeover > 0.0) {
        avail_water = avail_water + intcp_changeover;
        infil = infil + intcp_changeover;
        if (hru_flag)
          perv_comp(soil_moist_prev, o->carea_max[i], o->smidx_coef[i],
                    o->smidx_exp[i], intcp_changeover, intcp_changeover, &infil,
                    &srp, &contrib_fraction);
      }
      if (pptmix_nopack != 0) {
        avail_water = avail_water + through_rain;
        infil = infil + through_rain;
        if (hru_flag)
          perv_comp(soil_moist_prev, o->carea_max[i], o->smidx_coef[i],
                    o->smidx_exp[i], through_rain, through_rain, &infil, &srp,
                    &contrib_fraction);
      }
      int cond4 = pkwater_equiv < dnearzero;
      int cond6 = net_snow < nearzero;
      if (snowmelt > 0.0) {
        avail_water = avail_water + snowmelt;
        infil = infil + snowmelt;
        if (hru_flag) {
          if (pkwater_equiv > 0.0 || net_rain < nearzero)
            check_capacity(soil_moist_prev, o->soil_moist_max[i],
                           o->snowinfil_max[i], &infil, &srp);
          else
            perv_comp(soil_moist_prev, o->carea_max[i], o->smidx_coef[i],
                      o->smidx_exp[i], snowmelt, net_ppt, &infil, &srp,
                      &contrib_fraction);
        }
      } else if (cond4) {
        if (cond6 && through_rain > 0.0) {
          avail_water = avail_water + through_rain;
          infil = infil + through_rain;
          if (hru_flag)
            perv_comp(soil_moist_prev, o->carea_max[i], o->smidx_coef[i],
                      o->smidx_exp[i], through_rain, through_rain, &infil, &srp,
                      &contrib_fraction);
        }
      } else if (infil > 0.0) {
        if (hru_flag)
          check_capacity(soil_moist_prev, o->soil_moist_max[i],
                         o->snowinfil_max[i], &infil, &srp);
      }
      if (hruarea_imperv > 0.0) {
        imperv_stor = imperv_stor + avail_water;
        if (hru_flag) {
          if (imperv_stor > o->imperv_stor_max[i]) {
            sri = imperv_stor - o->imperv_stor_max[i];
            imperv_stor = o->imperv_stor_max[i];
          }
        }
      }
    }
    int dprst_chk = 0;
    if (o->dprst_flag) {
      o->dprst_in[i] = 0.0;
      if (o->dprst_area_max[i] > 0.0) {
        dprst_chk = 1;
        /* dprst_comp :992-1228 */
        const double dprst_area_open_max = VAR(dprst_area_open_max)[i];
        const double dprst_area_clos_max = VAR(dprst_area_clos_max)[i];
        const double dprst_vol_open_max = o->dprst_vol_open_max[i];
        const double dprst_vol_clos_max = o->dprst_vol_clos_max[i];
        double dprst_vol_open = VAR(dprst_vol_open)[i];
        double dprst_vol_clos = VAR(dprst_vol_clos)[i];
        double dprst_area_clos = VAR(dprst_area_clos)[i]; /* local only, :1120-1132 */
        const double rain = availh2o; /* net_rain=availh2o in the call :716 */
        double inflow = 0.0;
        if (pptmix_nopack) inflow = inflow + rain;
        if (snowmelt > 0.0) {
          inflow = inflow + snowmelt;
        } else if (pkwater_equiv < dnearzero) {
          if (net_snow < nearzero && rain > 0.0) inflow = inflow + rain;
        }
        double dprst_in = 0.0;
        if (dprst_area_open_max > 0.0) {
          dprst_in = inflow * dprst_area_open_max;
          dprst_vol_open = dprst_vol_open + dprst_in;
        }
        if (dprst_area_clos_max > 0.0) {
          double tmp1 = inflow * dprst_area_clos_max;
          dprst_vol_clos = dprst_vol_clos + tmp1;
          dprst_in = dprst_in + tmp1;
        }
        dprst_in = dprst_in / hruarea;
        double dprst_srp = 0.0, dprst_sri = 0.0;
        if (srp > 0.0) {
          double tmp = srp * perv_frac * o->sro_to_dprst_perv[i] * hruarea;
          if (dprst_area_open_max > 0.0) {
            double dprst_srp_open = tmp * o->dprst_frac_open[i];
            dprst_srp = dprst_srp_open / hruarea;
            dprst_vol_open = dprst_vol_open + dprst_srp_open;
          }
          if (dprst_area_clos_max > 0.0) {
            double dprst_srp_clos = tmp * o->dprst_frac_clos[i];
            dprst_srp = dprst_srp + dprst_srp_clos / hruarea;
            dprst_vol_clos = dprst_vol_clos + dprst_srp_clos;
          }
          srp = srp - dprst_srp / perv_frac;
          if (srp < 0.0) srp = 0.0;
        }
        if (sri > 0.0) {
          double tmp = sri * imperv_frac * o->sro_to_dprst_imperv[i] * hruarea;
          if (dprst_area_open_max > 0.0) {
            double dprst_sri_open = tmp * o->dprst_frac_open[i];
            dprst_sri = dprst_sri_open / hruarea;
            dprst_vol_open = dprst_vol_open + dprst_sri_open;
          }
          if (dprst_area_clos_max > 0.0) {
            double dprst_sri_clos = tmp * o->dprst_frac_clos[i];
            dprst_sri = dprst_sri + dprst_sri_clos / hruarea;
            dprst_vol_clos = dprst_vol_clos + dprst_sri_clos;
          }
          sri = sri - dprst_sri / imperv_frac;
          if (sri < 0.0) sri = 0.0;
        }
        VAR(dprst_insroff_hru)[i] = dprst_srp + dprst_sri;
        double dprst_area_open = 0.0;
        if (dprst_vol_open > 0.0) {
          double open_vol_r = dprst_vol_open / dprst_vol_open_max;
          double frac_op_ar;
          if (open_vol_r < nearzero) frac_op_ar = 0.0;
          else if (open_vol_r > 1.0) frac_op_ar = 1.0;
          else frac_op_ar = exp(o->va_open_exp[i] * log(open_vol_r));
          dprst_area_open = dprst_area_open_max * frac_op_ar;
          if (dprst_area_open > dprst_area_open_max)
            dprst_area_open = dprst_area_open_max;
        }
        if (dprst_area_clos_max > 0.0) {
          dprst_area_clos = 0.0;
          if (dprst_vol_clos > 0.0) {
            double clos_vol_r = dprst_vol_clos / dprst_vol_clos_max;
            double frac_cl_ar;
            if (clos_vol_r < nearzero) frac_cl_ar = 0.0;
            else if (clos_vol_r > 1.0) frac_cl_ar = 1.0;
            else frac_cl_ar = exp(o->va_clos_exp[i] * log(clos_vol_r));
            dprst_area_clos = dprst_area_clos_max * frac_cl_ar;
            if (dprst_area_clos > dprst_area_clos_max)
              dprst_area_clos = dprst_area_clos_max;
          }
        }
        double unsatisfied_et = avail_et;
        double dprst_avail_et = potet * (1.0 - snowcov_area) * o->dprst_et_coef[i];
        double dprst_evap_hru = 0.0;
        if (dprst_avail_et > 0.0) {
          double dprst_evap_open = 0.0, dprst_evap_clos = 0.0;
          if (dprst_area_open > 0.0) {
            dprst_evap_open = fmin(dprst_area_open * dprst_avail_et, dprst_vol_open);
            if (dprst_evap_open / hruarea > unsatisfied_et)
              dprst_evap_open = unsatisfied_et * hruarea;
            if (dprst_evap_open > dprst_vol_open) dprst_evap_open = dprst_vol_open;
            unsatisfied_et = unsatisfied_et - dprst_evap_open / hruarea;
            dprst_vol_open = dprst_vol_open - dprst_evap_open;
          }
          if (dprst_area_clos > 0.0) {
            dprst_evap_clos = fmin(dprst_area_clos * dprst_avail_et, dprst_vol_clos);
            if (dprst_evap_clos / hruarea > unsatisfied_et)
              dprst_evap_clos = unsatisfied_et * hruarea;
            if (dprst_evap_clos > dprst_vol_clos) dprst_evap_clos = dprst_vol_clos;
            dprst_vol_clos = dprst_vol_clos - dprst_evap_clos;
          }
          dprst_evap_hru = (dprst_evap_open + dprst_evap_clos) / hruarea;
        }
        double dprst_seep_hru = 0.0;
        if (dprst_vol_open > 0.0) {
          double seep_open = dprst_vol_open * o->dprst_seep_rate_open[i];
          dprst_vol_open = dprst_vol_open - seep_open;
          if (dprst_vol_open < 0.0) {
            seep_open = seep_open + dprst_vol_open;
            dprst_vol_open = 0.0;
          }
          dprst_seep_hru = seep_open / hruarea;
        }
        double dprst_sroff_hru = 0.0;
        if (dprst_vol_open > 0.0) {
          dprst_sroff_hru = fmax(0.0, dprst_vol_open - dprst_vol_open_max);
          dprst_sroff_hru =
              dprst_sroff_hru +
              fmax(0.0, (dprst_vol_open - dprst_sroff_hru - o->dprst_vol_thres_open[i]) *
                            o->dprst_flow_coef[i]);
          dprst_vol_open = dprst_vol_open - dprst_sroff_hru;
          dprst_sroff_hru = dprst_sroff_hru / hruarea;
          if (dprst_vol_open < 0.0) dprst_vol_open = 0.0;
        }
        if (dprst_area_clos_max > 0.0) {
          if (dprst_area_clos > nearzero) {
            double seep_clos = dprst_vol_clos * o->dprst_seep_rate_clos[i];
            dprst_vol_clos = dprst_vol_clos - seep_clos;
            if (dprst_vol_clos < 0.0) {
              seep_clos = seep_clos + dprst_vol_clos;
              dprst_vol_clos = 0.0;
            }
            dprst_seep_hru = dprst_seep_hru + seep_clos / hruarea;
          }
        }
        avail_et = avail_et - dprst_evap_hru;
        if (dprst_vol_open_max > 0.0)
          VAR(dprst_vol_open_frac)[i] = dprst_vol_open / dprst_vol_open_max;
        if (dprst_vol_clos_max > 0.0)
          VAR(dprst_vol_clos_frac)[i] = dprst_vol_clos / dprst_vol_clos_max;
        if (dprst_vol_open_max + dprst_vol_clos_max > 0.0)
          VAR(dprst_vol_frac)[i] = (dprst_vol_open + dprst_vol_clos) /
                                   (dprst_vol_open_max + dprst_vol_clos_max);
        VAR(dprst_stor_hru)[i] = (dprst_vol_open + dprst_vol_clos) / hruarea;
        o->dprst_in[i] = dprst_in;
        VAR(dprst_vol_open)[i] = dprst_vol_open;
        VAR(dprst_area_open)[i] = dprst_area_open;
        VAR(dprst_vol_clos)[i] = dprst_vol_clos;
        VAR(dprst_sroff_hru)[i] = dprst_sroff_hru;
        VAR(dprst_evap_hru)[i] = dprst_evap_hru;
        VAR(dprst_seep_hru)[i] = dprst_seep_hru;
        runoff = runoff + dprst_sroff_hru * hruarea;
      }
    }
    double srunoff = 0.0;
    if (hru_type == HRU_LAND) {
      runoff = runoff + srp * perv_area + sri * hruarea_imperv;
      srunoff = runoff / hruarea;
      VAR(hru_sroffp)[i] = srp * perv_frac;
    }
    if (hruarea_imperv > 0.0) {
      if (imperv_stor > 0.0) {
        /* imperv_et :1278-1287 */
        double imperv_evap = VAR(imperv_evap)[i];
        if (snowcov_area < 1.0) {
          if (potet < imperv_stor) imperv_evap = potet * (1.0 - snowcov_area);
          else imperv_evap = imperv_stor * (1.0 - snowcov_area);
          if (imperv_evap * imperv_frac > avail_et) imperv_evap = avail_et / imperv_frac;
          imperv_stor = imperv_stor - imperv_evap;
        }
        double hru_impervevap = imperv_evap * imperv_frac;
        avail_et = avail_et - hru_impervevap;
        if (avail_et < 0.0) {
          hru_impervevap = hru_impervevap + avail_et;
          if (hru_impervevap < 0.0) hru_impervevap = 0.0;
          imperv_evap = imperv_evap / imperv_frac;
          imperv_stor = imperv_stor - avail_et / imperv_frac;
          avail_et = 0.0;
        }
        VAR(imperv_evap)[i] = imperv_evap;
        VAR(hru_impervevap)[i] = hru_impervevap;
        VAR(hru_impervstor)[i] = imperv_stor * imperv_frac;
      }
      VAR(hru_sroffi)[i] = sri * imperv_frac;
    }
    if (dprst_chk)
      VAR(dprst_stor_hru)[i] = (VAR(dprst_vol_open)[i] + VAR(dprst_vol_clos)[i]) / hruarea;
    VAR(imperv_stor)[i] = imperv_stor;
    VAR(infil)[i] = infil;
    VAR(contrib_fraction)[i] = contrib_fraction;
    VAR(sroff)[i] = srunoff;
    /* :529-538 */
    VAR(infil_hru)[i] = infil * perv_frac;
    VAR(hru_impervstor_change)[i] = VAR(hru_impervstor)[i] - VAR(hru_impervstor_old)[i];
    VAR(dprst_stor_hru_change)[i] = VAR(dprst_stor_hru)[i] - VAR(dprst_stor_hru_old)[i];
    VAR(sroff_vol)[i] = srunoff * o->hru_in_to_cf[i];
  }
}

/* ====================================================================== */
/* PRMSSoilzone (hydrology/prms_soilzone.py:588-1456)                     */
/* ====================================================================== */

/* :1260-1312; returns -1 on the reference's RuntimeError */
static int compute_interflow(double coef_lin, double coef_sq, double ssres_in,
                             double *storage, double *inter_flow) {
  double flow;
  if (coef_lin <= 0.0 && ssres_in <= 0.0) {
    double c1 = coef_sq * *storage;
    flow = *storage * (c1 / (1.0 + c1));
  } else if (coef_lin > 0.0 && coef_sq <= 0.0) {
    double c2 = 1.0 - exp(-coef_lin);
    flow = ssres_in * (1.0 - c2 / coef_lin) + *storage * c2;
  } else if (coef_sq > 0.0) {
    double c3 = sqrt(pow(coef_lin, 2.0) + 4.0 * coef_sq * ssres_in);
    double sos = *storage - ((c3 - coef_lin) / (2.0 * coef_sq));
    if (c3 == 0.0) return -1;
    double c1 = coef_sq * sos / c3;
    double c2 = 1.0 - exp(-c3);
    if ((1.0 + c1 * c2) > 0.0)
      flow = ssres_in + ((sos * (1.0 + c1) * c2) / (1.0 + c1 * c2));
    else
      flow = ssres_in;
  } else {
    flow = 0.0;
  }
  if (flow < 0.0) flow = 0.0;
  else if (flow > *storage) flow = *storage;
  *storage = *storage - flow;
  *inter_flow = flow;
  return 0;
}

static int soilzone_day(Orc *o) {
  const int nhru = o->nhru;
  int err = 0;
#pragma omp parallel for schedule(static) reduction(| : err)
  for (int hh = 0; hh < nhru; ++hh) {
    const int hru_type = o->hru_type[hh];
    const double hru_frac_perv = o->sz_hru_frac_perv[hh];
    const double potet = VAR(potet)[hh];
    const double soil_moist_max = o->sz_soil_moist_max[hh];
    const double soil_rechr_max = o->sz_soil_rechr_max[hh];
    const double pref_flow_den = o->sz_pref_flow_den[hh];
    const double pref_flow_max = VAR(pref_flow_max)[hh];
    const double gwin = 0.0;
    double soil_to_gw = 0.0, soil_to_ssr = 0.0, ssr_to_gw = 0.0, slow_flow = 0.0;
    double ssres_flow = 0.0, potet_rechr = 0.0, potet_lower = 0.0;
    const double snow_free = 1.0 - VAR(snowcov_area)[hh];
    o->sz_snow_free[hh] = snow_free;
    double hru_actet = VAR(hru_impervevap)[hh] + VAR(hru_intcpevap)[hh] + VAR(snow_evap)[hh];
    if (o->dprst_flag) hru_actet = hru_actet + VAR(dprst_evap_hru)[hh];
    double dunnianflw = 0.0, dunnianflw_pfr = 0.0, dunnianflw_gvr = 0.0, prefflow = 0.0;
    double avail_potet = fmax(0.0, potet - hru_actet);
    double capwater_maxin = VAR(infil_hru)[hh] / hru_frac_perv;
    double pref_flow_stor = VAR(pref_flow_stor)[hh];
    double soil_moist = VAR(soil_moist)[hh], soil_rechr = VAR(soil_rechr)[hh];
    double slow_stor = VAR(slow_stor)[hh];
    if (o->pref_flow_infil_frac[hh] != 0.0) {
      double pref_flow_maxin = 0.0;
      VAR(pref_flow_infil)[hh] = 0.0;
      if (capwater_maxin > 0.0) {
        pref_flow_maxin = capwater_maxin * o->pref_flow_infil_frac[hh];
        capwater_maxin = capwater_maxin - pref_flow_maxin;
        pref_flow_maxin = pref_flow_maxin * hru_frac_perv;
        pref_flow_stor = pref_flow_stor + pref_flow_maxin;
        dunnianflw_pfr = fmax(0.0, pref_flow_stor - pref_flow_max);
        if (dunnianflw_pfr > 0.0) pref_flow_stor = pref_flow_max;
        VAR(pref_flow_infil)[hh] = pref_flow_maxin - dunnianflw_pfr;
      }
    }
    VAR(cap_infil_tot)[hh] = capwater_maxin * hru_frac_perv;
    double cap_waterin = capwater_maxin;
    if ((capwater_maxin + soil_moist) > 0.0) {
      /* _compute_soilmoist :1190-1257 */
      double infil = cap_waterin;
      soil_rechr = fmin(soil_rechr + infil, soil_rechr_max);
      double excess = soil_moist + infil;
      soil_moist = fmin(excess, soil_moist_max);
      excess = (excess - soil_moist_max) * hru_frac_perv;
      if (excess > 0.0) {
        if (o->sz_soil2gw_flag[hh]) {
          soil_to_gw = fmin(o->soil2gw_max[hh], excess);
          excess = excess - soil_to_gw;
        }
        if (excess > (infil * hru_frac_perv)) infil = 0.0;
        else infil = infil - (excess / hru_frac_perv);
        soil_to_ssr = fmax(0.0, excess);
      }
      cap_waterin = infil * hru_frac_perv;
    }
    VAR(cap_waterin)[hh] = cap_waterin;
    double topfr = 0.0;
    double availh2o = slow_stor + soil_to_ssr;
    if (hru_type == HRU_LAND) {
      topfr = fmax(0.0, availh2o - VAR(pref_flow_thrsh)[hh]);
      double ssresin = soil_to_ssr - topfr;
      slow_stor = fmax(0.0, availh2o - topfr);
      if (slow_stor > 0.0)
        err |= compute_interflow(o->slowcoef_lin[hh], o->slowcoef_sq[hh], ssresin,
                                 &slow_stor, &slow_flow) != 0;
    } else if (hru_type == HRU_SWALE) {
      slow_stor = availh2o;
    }
    if (slow_stor > 0.0 && o->ssr2gw_rate[hh] > 0.0) {
      /* _compute_gwflow :1315-1320 */
      ssr_to_gw = fmax(0.0, o->ssr2gw_rate[hh] * pow(slow_stor, o->ssr2gw_exp[hh]));
      ssr_to_gw = fmin(ssr_to_gw, slow_stor);
      slow_stor = slow_stor - ssr_to_gw;
    }
    if (pref_flow_den > 0.0) {
      availh2o = pref_flow_stor + topfr;
      dunnianflw_gvr = fmax(0.0, availh2o - pref_flow_max);
      if (dunnianflw_gvr > 0.0) topfr = fmax(0.0, topfr - dunnianflw_gvr);
      VAR(pref_flow_in)[hh] = VAR(pref_flow_infil)[hh] + topfr;
      pref_flow_stor = pref_flow_stor + topfr;
      if (pref_flow_stor > 0.0)
        err |= compute_interflow(o->fastcoef_lin[hh], o->fastcoef_sq[hh],
                                 VAR(pref_flow_in)[hh], &pref_flow_stor,
                                 &prefflow) != 0;
    } else if (hru_type == HRU_LAND) {
      dunnianflw_gvr = topfr;
    }
    double perv_actet = 0.0;
    if (soil_moist > 0.0) {
      /* _compute_szactet :1323-1456 */
      int et_type;
      const int transp_on = (int)VAR(transp_on)[hh];
      if (avail_potet < nearzero) {
        et_type = 1;
        avail_potet = 0.0;
      } else if (!transp_on) {
        if (snow_free < 0.01) et_type = 1;
        else et_type = 2;
      } else if (o->cov_type[hh] > 0) {
        et_type = 3;
      } else if (snow_free < 0.01) {
        et_type = 1;
      } else {
        et_type = 2;
      }
      double et;
      if (et_type == 2 || et_type == 3) {
        double pcts = soil_moist / soil_moist_max;
        double pctr = soil_rechr / soil_rechr_max;
        potet_lower = avail_potet;
        potet_rechr = avail_potet;
        const int soil_type = o->soil_type[hh];
        const double TWOTHIRDS = 2.0 / 3.0;
        if (soil_type == SOIL_SAND) {
          if (pcts < 0.25) potet_lower = 0.5 * pcts * avail_potet;
          if (pctr < 0.25) potet_rechr = 0.5 * pctr * avail_potet;
        } else if (soil_type == SOIL_LOAM) {
          if (pcts < 0.5) potet_lower = pcts * avail_potet;
          if (pctr < 0.5) potet_rechr = pctr * avail_potet;
        } else if (soil_type == SOIL_CLAY) {
          if (pcts < TWOTHIRDS && pcts > ONETHIRD) potet_lower = pcts * avail_potet;
          else if (pcts <= ONETHIRD) potet_lower = 0.5 * pcts * avail_potet;
          if (pctr < TWOTHIRDS && pctr > ONETHIRD) potet_rechr = pctr * avail_potet;
          else if (pctr <= ONETHIRD) potet_rechr = 0.5 * pctr * avail_potet;
        }
        if (et_type == 2) potet_rechr = potet_rechr * snow_free;
        if (potet_rechr > soil_rechr) {
          potet_rechr = soil_rechr;
          soil_rechr = 0.0;
        } else {
          soil_rechr = soil_rechr - potet_rechr;
        }
        if (et_type == 2 || potet_rechr >= potet_lower) {
          if (potet_rechr > soil_moist) {
            potet_rechr = soil_moist;
            soil_moist = 0.0;
          } else {
            soil_moist = soil_moist - potet_rechr;
          }
          et = potet_rechr;
        } else if (potet_lower > soil_moist) {
          et = soil_moist;
          soil_moist = 0.0;
        } else {
          soil_moist = soil_moist - potet_lower;
          et = potet_lower;
        }
        if (soil_rechr > soil_moist) soil_rechr = soil_moist;
      } else {
        et = 0.0;
      }
      perv_actet = et;
    }
    hru_actet = hru_actet + perv_actet * hru_frac_perv;
    avail_potet = potet - hru_actet;
    double soil_lower = soil_moist - soil_rechr;
    double ssres_stor;
    if (hru_type == HRU_LAND) {
      dunnianflw = dunnianflw_gvr + dunnianflw_pfr;
      VAR(dunnian_flow)[hh] = dunnianflw;
      ssres_flow = slow_flow;
      if (pref_flow_den > 0.0) {
        VAR(pref_flow)[hh] = prefflow;
        ssres_flow = ssres_flow + prefflow;
      }
      VAR(sroff)[hh] = VAR(sroff)[hh] + VAR(dunnian_flow)[hh];
      ssres_stor = slow_stor + pref_flow_stor;
    } else {
      availh2o = slow_stor - o->sz_sat_threshold[hh];
      VAR(swale_actet)[hh] = 0.0;
      if (availh2o > 0.0) {
        double unsatisfied_et = potet - hru_actet;
        if (unsatisfied_et > 0.0) {
          availh2o = fmin(availh2o, unsatisfied_et);
          VAR(swale_actet)[hh] = availh2o;
          hru_actet = hru_actet + VAR(swale_actet)[hh];
          slow_stor = slow_stor - VAR(swale_actet)[hh];
        }
      }
      ssres_stor = slow_stor;
    }
    VAR(ssres_in)[hh] = soil_to_ssr + VAR(pref_flow_infil)[hh] + gwin;
    VAR(unused_potet)[hh] = potet - hru_actet;
    VAR(soil_to_gw)[hh] = soil_to_gw; VAR(soil_to_ssr)[hh] = soil_to_ssr;
    VAR(ssr_to_gw)[hh] = ssr_to_gw; VAR(slow_flow)[hh] = slow_flow;
    VAR(ssres_flow)[hh] = ssres_flow; VAR(potet_rechr)[hh] = potet_rechr;
    VAR(potet_lower)[hh] = potet_lower; VAR(soil_moist)[hh] = soil_moist;
    VAR(soil_rechr)[hh] = soil_rechr; VAR(hru_actet)[hh] = hru_actet;
    VAR(slow_stor)[hh] = slow_stor; VAR(pref_flow_stor)[hh] = pref_flow_stor;
    VAR(perv_actet)[hh] = perv_actet; VAR(soil_lower)[hh] = soil_lower;
    VAR(ssres_stor)[hh] = ssres_stor;
    /* :1101-1126 */
    if (VAR(soil_lower_max)[hh] > 0.0)
      VAR(soil_lower_ratio)[hh] = soil_lower / VAR(soil_lower_max)[hh];
    VAR(soil_moist_tot)[hh] = ssres_stor + soil_moist * hru_frac_perv;
    double recharge = soil_to_gw + ssr_to_gw;
    if (o->dprst_flag) recharge = recharge + VAR(dprst_seep_hru)[hh];
    VAR(recharge)[hh] = recharge;
    VAR(pref_flow_stor_change)[hh] = pref_flow_stor - VAR(pref_flow_stor_prev)[hh];
    VAR(soil_lower_change)[hh] = soil_lower - VAR(soil_lower_prev)[hh];
    VAR(soil_rechr_change)[hh] = soil_rechr - VAR(soil_rechr_prev)[hh];
    VAR(slow_stor_change)[hh] = slow_stor - VAR(slow_stor_prev)[hh];
    VAR(soil_lower_change_hru)[hh] = VAR(soil_lower_change)[hh] * hru_frac_perv;
    VAR(soil_rechr_change_hru)[hh] = VAR(soil_rechr_change)[hh] * hru_frac_perv;
    VAR(perv_actet_hru)[hh] = perv_actet * hru_frac_perv;
    VAR(ssres_flow_vol)[hh] = ssres_flow * o->hru_in_to_cf[hh];
    /* :707 */
    VAR(sroff_vol)[hh] = VAR(sroff)[hh] * o->hru_in_to_cf[hh];
  }
  if (err) {
    snprintf(o->err, sizeof o->err,
             "ERROR, in compute_interflow sos=0 (prms_soilzone.py:1282-1287)");
    return -1;
  }
  return 0;
}

/* ====================================================================== */
/* PRMSGroundwater (hydrology/prms_groundwater.py:231-275)                */
/* ====================================================================== */
static void groundwater_day(Orc *o) {
  const int nhru = o->nhru;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nhru; ++i) {
    double gwarea = o->hru_area[i];
    double soil_to_gw_vol = VAR(soil_to_gw)[i] * gwarea;
    double ssr_to_gw_vol = VAR(ssr_to_gw)[i] * gwarea;
    double dprst_seep_hru_vol = VAR(dprst_seep_hru)[i] * gwarea;
    double stor = VAR(gwres_stor)[i] * gwarea;
    stor += soil_to_gw_vol + ssr_to_gw_vol + dprst_seep_hru_vol;
    double flow = stor * o->gwflow_coef[i];
    stor -= flow;
    double sink = stor * o->gwsink_coef[i];
    if (sink > stor) sink = stor;
    stor -= sink;
    VAR(gwres_stor)[i] = stor / gwarea;
    VAR(gwres_flow)[i] = flow / gwarea;
    VAR(gwres_sink)[i] = sink / gwarea;
    VAR(gwres_stor_change)[i] = VAR(gwres_stor)[i] - VAR(gwres_stor_old)[i];
    VAR(gwres_flow_vol)[i] = VAR(gwres_flow)[i] * o->hru_in_to_cf[i];
  }
}

/* ====================================================================== */
/* PRMSChannel (hydrology/prms_channel.py:388-585)                        */
/* ====================================================================== */
static void muskingum_day(Orc *o) {
  const int nseg = o->nseg;
  double *seg_inflow = o->ch_seg_inflow, *seg_inflow0 = o->ch_seg_inflow0;
  double *inflow_ts = o->ch_inflow_ts, *outflow_ts = o->ch_outflow_ts;
  double *seg_current_sum = o->ch_seg_current_sum, *up = o->ch_up;
  double *seg_outflow = SEG(seg_outflow);
  const double *lat = SEG(seg_lateral_inflow);
  for (int j = 0; j < nseg; ++j) {
    seg_inflow[j] = seg_inflow0[j] * 0.0;
    seg_outflow[j] = seg_inflow0[j] * 0.0;
    inflow_ts[j] = seg_inflow0[j] * 0.0;
    seg_current_sum[j] = seg_inflow0[j] * 0.0;
  }
  for (int ihr = 0; ihr < 24; ++ihr) {
    for (int j = 0; j < nseg; ++j) up[j] = seg_inflow[j] * 0.0;
    for (int k = 0; k < nseg; ++k) {
      int j = o->ch_order[k];
      double cur = lat[j] + up[j];
      seg_inflow[j] += cur;
      inflow_ts[j] += cur;
      seg_current_sum[j] += up[j];
      int tsi = o->ch_tsi[j];
      /* python: (ihr + 1) % tsi == 0; tsi == -1 always true */
      int rem = (tsi < 0) ? 0 : (ihr + 1) % tsi;
      if (rem == 0) {
        inflow_ts[j] /= o->ch_ts[j];
        if (tsi > 0)
          outflow_ts[j] = inflow_ts[j] * o->ch_c0[j] + seg_inflow0[j] * o->ch_c1[j] +
                          outflow_ts[j] * o->ch_c2[j];
        else
          outflow_ts[j] = inflow_ts[j];
        seg_inflow0[j] = inflow_ts[j];
        inflow_ts[j] = 0.0;
      }
      seg_outflow[j] += outflow_ts[j];
      int to = o->tosegment[j] - 1;
      if (to >= 0) up[to] += outflow_ts[j];
    }
  }
  for (int j = 0; j < nseg; ++j) {
    seg_outflow[j] /= 24.0;
    seg_inflow[j] /= 24.0;
    SEG(seg_upstream_inflow)[j] = seg_current_sum[j] / 24.0;
    SEG(seg_stor_change)[j] = (seg_inflow[j] - seg_outflow[j]) * 86400.0;
    SEG(channel_outflow_vol)[j] =
        (o->ch_outflow_mask[j] ? seg_outflow[j] : 0.0) * 86400.0;
  }
}

static void channel_day(Orc *o) {
  const int nhru = o->nhru, nseg = o->nseg;
  const double s_per_time = 86400.0;
  double *lat = SEG(seg_lateral_inflow);
  for (int j = 0; j < nseg; ++j) lat[j] = 0.0;
  for (int i = 0; i < nhru; ++i) {
    int iseg = o->hru_segment[i] - 1;
    if (iseg < 0) {
      VAR(channel_sroff_vol)[i] = 0.0;
      VAR(channel_ssres_flow_vol)[i] = 0.0;
      VAR(channel_gwres_flow_vol)[i] = 0.0;
      continue;
    }
    VAR(channel_sroff_vol)[i] = VAR(sroff_vol)[i];
    VAR(channel_ssres_flow_vol)[i] = VAR(ssres_flow_vol)[i];
    VAR(channel_gwres_flow_vol)[i] = VAR(gwres_flow_vol)[i];
    double lateral = (VAR(channel_sroff_vol)[i] + VAR(channel_ssres_flow_vol)[i] +
                      VAR(channel_gwres_flow_vol)[i]) / s_per_time;
    lat[iseg] += lateral;
  }
  muskingum_day(o);
}

/* ====================================================================== */
/* Budget (base/budget.py:253-416): count units that fail                  */
/* np.allclose(in, out + dstor, rtol=1e-5, atol=1e-5)                      */
/* ====================================================================== */
static int notclose(double a, double b) {
  return !(fabs(a - b) <= 1e-5 + 1e-5 * fabs(b));
}

static void budgets_day(Orc *o, int32_t *viol) {
  const int nhru = o->nhru, nseg = o->nseg;
  int32_t c[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < nhru; ++i) {
    /* canopy prms_canopy.py:149-159 */
    double in = 0 + VAR(hru_rain)[i] + VAR(hru_snow)[i];
    double out = 0 + VAR(net_rain)[i] + VAR(net_snow)[i] + VAR(hru_intcpevap)[i] +
                 VAR(intcp_changeover)[i];
    double ds = 0 + VAR(hru_intcpstor_change)[i];
    c[0] += notclose(in, out + ds);
    /* snow prms_snow.py:257-266 */
    in = 0 + VAR(net_rain)[i] + VAR(net_snow)[i];
    out = 0 + VAR(snow_evap)[i] + VAR(snowmelt)[i] + VAR(through_rain)[i];
    ds = 0 + VAR(freeh2o_change)[i] + VAR(pk_ice_change)[i];
    c[1] += notclose(in, out + ds);
    /* runoff prms_runoff.py:228-251 */
    in = 0 + VAR(through_rain)[i] + VAR(snowmelt)[i] + VAR(intcp_changeover)[i];
    out = 0 + VAR(hru_sroffi)[i] + VAR(hru_sroffp)[i] + VAR(dprst_sroff_hru)[i] +
          VAR(infil_hru)[i] + VAR(hru_impervevap)[i] + VAR(dprst_seep_hru)[i] +
          VAR(dprst_evap_hru)[i];
    ds = 0 + VAR(hru_impervstor_change)[i] + VAR(dprst_stor_hru_change)[i];
    c[2] += notclose(in, out + ds);
    /* soilzone prms_soilzone.py:236-255 */
    in = 0 + VAR(infil_hru)[i];
    out = 0 + VAR(perv_actet_hru)[i] + VAR(soil_to_gw)[i] + VAR(ssr_to_gw)[i] +
          VAR(slow_flow)[i] + VAR(dunnian_flow)[i] + VAR(pref_flow)[i];
    ds = 0 + VAR(soil_rechr_change_hru)[i] + VAR(soil_lower_change_hru)[i] +
         VAR(slow_stor_change)[i] + VAR(pref_flow_stor_change)[i];
    c[3] += notclose(in, out + ds);
    /* groundwater prms_groundwater.py:119-132 */
    in = 0 + VAR(soil_to_gw)[i] + VAR(ssr_to_gw)[i] + VAR(dprst_seep_hru)[i];
    out = 0 + VAR(gwres_flow)[i];
    ds = 0 + VAR(gwres_stor_change)[i];
    c[4] += notclose(in, out + ds);
  }
  if (nseg > 0) {
    /* channel, global basis prms_channel.py:115,165-174 */
    double a = 0, b = 0, cc = 0, outv = 0, ds = 0;
    for (int i = 0; i < nhru; ++i) {
      a += VAR(channel_sroff_vol)[i];
      b += VAR(channel_ssres_flow_vol)[i];
      cc += VAR(channel_gwres_flow_vol)[i];
    }
    for (int j = 0; j < nseg; ++j) {
      outv += SEG(channel_outflow_vol)[j];
      ds += SEG(seg_stor_change)[j];
    }
    double in = 0 + a + b + cc;
    c[5] += notclose(in, outv + ds);
  }
  for (int k = 0; k < 6; ++k) viol[k] = c[k];
}

/* ====================================================================== */
/* drivers                                                                */
/* ====================================================================== */
int orc_run(Orc *o, int ndays, const int32_t *cal, const double *prcp,
            const double *tmax, const double *tmin, int n_hru_out,
            const int32_t *hru_idx, double *hru_out, int n_seg_out,
            const int32_t *seg_idx, double *seg_out, int32_t *budget_viol) {
  if (!o->inited) {
    snprintf(o->err, sizeof o->err, "orc_run before orc_init");
    return -1;
  }
  const int nhru = o->nhru, nseg = o->nseg;
  for (int t = 0; t < ndays; ++t) {
    const int doy = cal[4 * t + 0], dowy = cal[4 * t + 1];
    const int month = cal[4 * t + 2], dom = cal[4 * t + 3];
    const double *p = prcp + (int64_t)t * nhru;
    const double *tx = tmax + (int64_t)t * nhru;
    const double *tn = tmin + (int64_t)t * nhru;
    /* ---- Model.advance (base/model.py:723-737): every process's
     * _advance_variables, in process order ---- */
    atmosphere_day(o, o->itime == 0, doy, month, dom, p, tx, tn);
    for (int i = 0; i < nhru; ++i) {
      VAR(hru_intcpstor_old)[i] = VAR(hru_intcpstor)[i];   /* prms_canopy.py:277 */
      VAR(freeh2o_prev)[i] = VAR(freeh2o)[i];               /* prms_snow.py:484 */
      VAR(pk_ice_prev)[i] = VAR(pk_ice)[i];
      VAR(hru_impervstor_old)[i] = VAR(hru_impervstor)[i]; /* prms_runoff.py:429 */
      VAR(dprst_stor_hru_old)[i] = VAR(dprst_stor_hru)[i];
      VAR(pref_flow_stor_prev)[i] = VAR(pref_flow_stor)[i]; /* prms_soilzone.py:581 */
      VAR(soil_rechr_prev)[i] = VAR(soil_rechr)[i];
      VAR(soil_lower_prev)[i] = VAR(soil_lower)[i];
      VAR(slow_stor_prev)[i] = VAR(slow_stor)[i];
      VAR(gwres_stor_old)[i] = VAR(gwres_stor)[i];          /* prms_groundwater.py:206 */
    }
    for (int j = 0; j < nseg; ++j)
      o->ch_seg_inflow0[j] = o->ch_seg_inflow[j];           /* prms_channel.py:385 */
    /* ---- Model.calculate (base/model.py:739-743) ---- */
    canopy_day(o);
    snow_day(o, doy, dowy, month);
    runoff_day(o);
    if (soilzone_day(o)) return -1;
    groundwater_day(o);
    if (nseg > 0) channel_day(o);
    if (budget_viol) budgets_day(o, budget_viol + 6 * t);
    for (int k = 0; k < n_hru_out; ++k)
      memcpy(hru_out + ((int64_t)t * n_hru_out + k) * nhru, o->v[hru_idx[k]],
             sizeof(double) * (size_t)nhru);
    for (int k = 0; k < n_seg_out; ++k)
      memcpy(seg_out + ((int64_t)t * n_seg_out + k) * nseg, o->s[seg_idx[k]],
             sizeof(double) * (size_t)nseg);
    o->itime++;
  }
  return 0;
}

int orc_run_channel(Orc *o, int ndays, const double *seg_lateral_inflow,
                    int n_seg_out, const int32_t *seg_idx, double *seg_out) {
  if (!o->inited) {
    snprintf(o->err, sizeof o->err, "orc_run_channel before orc_init");
    return -1;
  }
  const int nseg = o->nseg;
  for (int t = 0; t < ndays; ++t) {
    for (int j = 0; j < nseg; ++j) {
      o->ch_seg_inflow0[j] = o->ch_seg_inflow[j];
      SEG(seg_lateral_inflow)[j] = seg_lateral_inflow[(int64_t)t * nseg + j];
    }
    muskingum_day(o);
    for (int k = 0; k < n_seg_out; ++k)
      memcpy(seg_out + ((int64_t)t * n_seg_out + k) * nseg, o->s[seg_idx[k]],
             sizeof(double) * (size_t)nseg);
  }
  return 0;
}

/* Test aid: replace the Muskingum coefficients channel_init derived (tsi, ts,
 * c0, c1, c2 per segment) with given ones.  The reference computes
 * seg_depth ** (2/3) with numpy's vectorised power, which is not the libm pow
 * of channel_init in the last bit for some segments; with the reference's own
 * coefficients the routing restatements can be held to its results bit for bit. */
int orc_set_channel_coefficients(Orc *o, const int32_t *tsi, const double *ts, const double *c0,
                                 const double *c1, const double *c2) {
  if (!o->inited) {
    snprintf(o->err, sizeof o->err, "orc_set_channel_coefficients before orc_init");
    return -1;
  }
  for (int j = 0; j < o->nseg; ++j) {
    o->ch_tsi[j] = tsi[j];
    o->ch_ts[j] = ts[j];
    o->ch_c0[j] = c0[j];
    o->ch_c1[j] = c1[j];
    o->ch_c2[j] = c2[j];
  }
  return 0;
}

/* ====================================================================== */
/* FlowGraph (base/flow_graph.py:576-657) over PRMSChannelFlowNode          */
/* (hydrology/prms_channel_flow_graph.py:30-160, 367-414), PassThroughFlowNode */
/* (hydrology/pass_through_flow_node.py:6-53) and ObsInFlowNode             */
/* (hydrology/obsin_flow_node.py:8-82).  The nodes are this handle's        */
/* segments (tosegment = to_graph_index + 1; the Muskingum coefficients of  */
/* channel_init = PRMSChannelFlowNodeMaker._init_data :230-330); kind[j]:   */
/* 0 Muskingum-Mann, 1 pass-through, 2 observed flow.  obs[t][j] is the day's */
/* observation of an ObsIn node (negative = pass the inflow through).       */
/* out: [ndays][7][nnodes] in get_init_values order (flow_graph.py:425-436): */
/* outflows, node_upstream_inflows, node_outflows, node_storage_changes,    */
/* node_storages, node_sink_source, node_negative_sink_source.              */
/* ====================================================================== */
int orc_run_flow_graph(Orc *o, int ndays, const double *inflows, const int32_t *kind,
                       const double *obs, const double *flow_min, double *out) {
  if (!o->inited) {
    snprintf(o->err, sizeof o->err, "orc_run_flow_graph before orc_init");
    return -1;
  }
  const int nn = o->nseg;
  double *seg_inflow = o->ch_seg_inflow, *seg_inflow0 = o->ch_seg_inflow0;
  double *inflow_ts = o->ch_inflow_ts, *outflow_ts = o->ch_outflow_ts;
  double *up_acc = o->ch_seg_current_sum, *up = o->ch_up;
  double *seg_outflow = dalloc(nn, 0.0), *sub = dalloc(nn, 0.0), *ss_sum = dalloc(nn, 0.0);
  double *sink = dalloc(nn, 0.0);
  if (!o->fg_started) { /* PRMSChannelFlowNode.__init__ :63-69: every node state starts at zero */
    for (int j = 0; j < nn; ++j) {
      seg_inflow[j] = 0.0;
      seg_inflow0[j] = 0.0;
      outflow_ts[j] = 0.0;
      inflow_ts[j] = 0.0;
    }
    o->fg_started = 1;
  }
  for (int t = 0; t < ndays; ++t) {
    const double *lat = inflows + (int64_t)t * nn;
    const double *ob = obs ? obs + (int64_t)t * nn : NULL;
    for (int j = 0; j < nn; ++j) {
      /* FlowGraph._advance_variables -> node.advance() (:569-573) */
      if (kind[j] == 0) seg_inflow0[j] = seg_inflow[j];
      /* node.prepare_timestep() (:579-580) */
      seg_inflow[j] = 0.0;  /* MM: _seg_inflow; pass-through: _accum_inflow */
      seg_outflow[j] = kind[j] == 2 ? ob[j] : 0.0; /* ObsIn: _seg_outflow = the day's observation */
      /* SourceSink (kind 3): ob[j] is the day's requested source (+) / sink (-),
       * source_sink_flow_node.py:54-58 */
      inflow_ts[j] = 0.0;
      ss_sum[j] = 0.0;
      sink[j] = 0.0;
      up_acc[j] = 0.0;
    }
    for (int ihr = 0; ihr < 24; ++ihr) {
      for (int j = 0; j < nn; ++j) up[j] = 0.0;
      for (int k = 0; k < nn; ++k) {
        const int j = o->ch_order[k];
        if (kind[j] == 0) { /* _calculate_subtimestep_numpy :333-364 */
          const double cur = lat[j] + up[j];
          seg_inflow[j] += cur;
          inflow_ts[j] += cur;
          const int tsi = o->ch_tsi[j];
          const int rem = (tsi < 0) ? 0 : (ihr + 1) % tsi;
          if (rem == 0) {
            inflow_ts[j] /= o->ch_ts[j];
            if (tsi > 0)
              outflow_ts[j] = inflow_ts[j] * o->ch_c0[j] + seg_inflow0[j] * o->ch_c1[j] +
                              outflow_ts[j] * o->ch_c2[j];
            else
              outflow_ts[j] = inflow_ts[j];
            seg_inflow0[j] = inflow_ts[j];
            inflow_ts[j] = 0.0;
          }
          seg_outflow[j] += outflow_ts[j];
          sub[j] = outflow_ts[j];
        } else if (kind[j] == 1) { /* pass_through_flow_node.py:24-33 */
          const double in = up[j] + lat[j];
          seg_inflow[j] += in;
          seg_outflow[j] = seg_inflow[j] / (double)(ihr + 1);
          sub[j] = in;
        } else if (kind[j] == 3) { /* source_sink_flow_node.py:60-92 */
          const double in = up[j] + lat[j];
          double ss = ob[j];
          const double fmin = flow_min ? flow_min[j] : 0.0;
          double outq = in;
          if (ss >= 0.0) { /* a source is always applied */
            outq = in + ss;
          } else if (ss < 0.0 && in < fmin) { /* no sink below the minimum flow */
            outq = in;
            ss = 0.0;
          } else if (ss < 0.0 && in >= fmin) {
            if (in + ss < fmin) {
              ss = fmin - in;
              outq = fmin;
            } else {
              outq = in + ss;
            }
          }
          seg_outflow[j] = outq; /* the node's outflow is the LAST substep's (:105-110) */
          ss_sum[j] += ss;
          sink[j] = ss_sum[j] / (double)(ihr + 1);
          sub[j] = outq;
        } else { /* obsin_flow_node.py:43-57 */
          const double in = up[j] + lat[j];
          if (seg_outflow[j] >= 0.0) {
            ss_sum[j] += seg_outflow[j] - in;
          } else {
            seg_outflow[j] = in;
            ss_sum[j] += 0.0;
          }
          sink[j] = ss_sum[j] / (double)(ihr + 1);
          sub[j] = seg_outflow[j];
        }
        const int to = o->tosegment[j] - 1;
        if (to >= 0) up[to] += sub[j];
      }
      for (int j = 0; j < nn; ++j) up_acc[j] += up[j];
    }
    double *q = out + (int64_t)t * 7 * nn;
    for (int j = 0; j < nn; ++j) {
      double stor_change = 0.0;
      if (kind[j] == 0) { /* finalize_timestep :107-111 */
        seg_outflow[j] = seg_outflow[j] / 24.0;
        seg_inflow[j] = seg_inflow[j] / 24.0;
        stor_change = seg_inflow[j] - seg_outflow[j];
      }
      q[0 * nn + j] = o->ch_outflow_mask[j] ? seg_outflow[j] : 0.0;
      q[1 * nn + j] = up_acc[j] / 24.0;
      q[2 * nn + j] = seg_outflow[j];
      q[3 * nn + j] = stor_change;
      q[4 * nn + j] = NAN;
      q[5 * nn + j] = sink[j];
      q[6 * nn + j] = -1.0 * sink[j];
    }
  }
  free(seg_outflow); free(sub); free(ss_sum); free(sink);
  return 0;
}
