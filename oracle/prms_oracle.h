/*
 * prms_oracle.h -- CPU restatement (plain C, scalar loops) of pywatershed's
 * per-timestep PRMS/NHM process chain.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  The product (pywatershed_b200/) never links,
 * includes or imports this code.
 *
 * Parity status: PINNED against outputs of the reference itself (numpy
 * calc_method), generated in the build container by oracle/gen_golden.py and
 * committed under tests/golden/ (the reference ships no golden vectors for
 * this path).
 *
 * Registries (X-macros) of the parameters and public variables the chain
 * touches; names are the reference's own (static/metadata, the yaml files).
 */
#ifndef PRMS_ORACLE_H
#define PRMS_ORACLE_H

#include <stdint.h>

/* ---- per-HRU float64 parameters ------------------------------------- */
#define ORC_HRU_F8_PARAMS(X) \
  X(hru_slope) X(hru_aspect) X(hru_lat) X(hru_area) X(hru_in_to_cf) \
  X(radj_sppt) X(radj_wppt) X(jh_coef_hru) X(transp_tmax) \
  X(covden_sum) X(covden_win) X(srain_intcp) X(wrain_intcp) X(snow_intcp) \
  X(potet_sublim) X(emis_noppt) X(freeh2o_cap) X(rad_trncf) X(snarea_thresh) \
  X(den_init) X(den_max) X(settle_const) X(snowpack_init) \
  X(hru_percent_imperv) X(imperv_stor_max) X(dprst_frac) X(carea_max) \
  X(smidx_coef) X(smidx_exp) X(soil_moist_max) X(snowinfil_max) \
  X(dprst_depth_avg) X(dprst_et_coef) X(dprst_flow_coef) X(dprst_frac_init) \
  X(dprst_frac_open) X(dprst_seep_rate_clos) X(dprst_seep_rate_open) \
  X(sro_to_dprst_imperv) X(sro_to_dprst_perv) X(va_open_exp) X(va_clos_exp) \
  X(op_flow_thres) X(fastcoef_lin) X(fastcoef_sq) X(pref_flow_den) \
  X(pref_flow_infil_frac) X(sat_threshold) X(slowcoef_lin) X(slowcoef_sq) \
  X(soil_moist_init_frac) X(soil_rechr_init_frac) X(soil_rechr_max_frac) \
  X(soil2gw_max) X(ssr2gw_exp) X(ssr2gw_rate) X(ssstor_init_frac) \
  X(gwflow_coef) X(gwsink_coef) X(gwstor_init)

/* ---- per-HRU integer parameters (int32 on this side) ---------------- */
#define ORC_HRU_I4_PARAMS(X) \
  X(transp_beg) X(transp_end) X(cov_type) X(hru_type) X(hru_deplcrv) \
  X(melt_force) X(melt_look) X(soil_type) X(hru_segment)

/* ---- monthly [12][nhru] float64 parameters -------------------------- */
#define ORC_MON_F8_PARAMS(X) \
  X(radadj_intcp) X(radadj_slope) X(tmax_index) X(dday_slope) X(dday_intcp) \
  X(radmax) X(ppt_rad_adj) X(tmax_allsnow) X(tmax_allrain_offset) X(jh_coef) \
  X(tmax_cbh_adj) X(tmin_cbh_adj) X(snow_cbh_adj) X(rain_cbh_adj) \
  X(adjmix_rain) X(cecn_coef)

#define ORC_MON_I4_PARAMS(X) X(tstorm_mo)

/* ---- per-segment parameters ----------------------------------------- */
#define ORC_SEG_F8_PARAMS(X) \
  X(mann_n) X(seg_depth) X(seg_length) X(seg_slope) X(x_coef) \
  X(segment_flow_init)
#define ORC_SEG_I4_PARAMS(X) X(segment_type) X(tosegment)

/* ---- public per-HRU variables, in process order (name, init) -------- */
/* init values: get_init_values() of each reference process.             */
#define ORC_NAN (0.0 / 0.0)
#define ORC_HRU_VARS(X) \
  /* PRMSAtmosphere (prms_atmosphere.py:245-262) */ \
  X(tmaxf, ORC_NAN) X(tminf, ORC_NAN) X(prmx, ORC_NAN) X(hru_ppt, ORC_NAN) \
  X(hru_rain, ORC_NAN) X(hru_snow, ORC_NAN) X(swrad, ORC_NAN) \
  X(potet, ORC_NAN) X(transp_on, 0) X(tmaxc, ORC_NAN) X(tavgc, ORC_NAN) \
  X(tminc, ORC_NAN) X(pptmix, -9999) X(orad_hru, ORC_NAN) \
  /* PRMSCanopy (prms_canopy.py:131-146) */ \
  X(net_ppt, 0) X(net_rain, 0) X(net_snow, 0) X(intcp_changeover, 0) \
  X(intcp_evap, 0) X(intcp_form, 0) X(intcp_stor, 0) X(intcp_transp_on, 0) \
  X(hru_intcpevap, 0) X(hru_intcpstor, 0) X(hru_intcpstor_change, 0) \
  X(hru_intcpstor_old, 0) \
  /* PRMSSnow (prms_snow.py:212-255) */ \
  X(ai, 0) X(albedo, 0) X(frac_swe, 0) X(freeh2o, 0) \
  X(freeh2o_change, ORC_NAN) X(freeh2o_prev, ORC_NAN) X(iasw, 0) \
  X(int_alb, 1) X(iso, 1) X(lso, 0) X(lst, 0) X(mso, 1) X(newsnow, 0) \
  X(pk_def, 0) X(pk_den, 0) X(pk_depth, 0) X(pk_ice, 0) \
  X(pk_ice_change, ORC_NAN) X(pk_ice_prev, ORC_NAN) X(pk_precip, 0) \
  X(pk_temp, 0) X(pksv, 0) X(pkwater_equiv, ORC_NAN) X(pptmix_nopack, 0) \
  X(pss, ORC_NAN) X(pst, ORC_NAN) X(salb, 0) X(scrv, 0) X(slst, 0) \
  X(snow_evap, 0) X(snowcov_area, 0) X(snowcov_areasv, 0) X(snowmelt, 0) \
  X(snsv, 0) X(tcal, 0) X(through_rain, ORC_NAN) \
  /* PRMSRunoff (prms_runoff.py:193-226) */ \
  X(contrib_fraction, 0) X(infil, 0) X(infil_hru, 0) X(sroff, 0) \
  X(sroff_vol, 0) X(hru_sroffp, 0) X(hru_sroffi, 0) X(imperv_stor, 0) \
  X(imperv_evap, 0) X(hru_impervevap, 0) X(hru_impervstor, 0) \
  X(hru_impervstor_old, 0) X(hru_impervstor_change, 0) X(dprst_vol_frac, 0) \
  X(dprst_vol_clos, 0) X(dprst_vol_open, 0) X(dprst_vol_clos_frac, 0) \
  X(dprst_vol_open_frac, 0) X(dprst_area_clos, 0) X(dprst_area_open, 0) \
  X(dprst_area_clos_max, 0) X(dprst_area_open_max, 0) X(dprst_sroff_hru, 0) \
  X(dprst_seep_hru, 0) X(dprst_evap_hru, 0) X(dprst_insroff_hru, 0) \
  X(dprst_stor_hru, 0) X(dprst_stor_hru_old, 0) X(dprst_stor_hru_change, 0) \
  /* PRMSSoilzone (prms_soilzone.py:186-233) */ \
  X(cap_infil_tot, 0) X(cap_waterin, 0) X(dunnian_flow, 0) X(hru_actet, 0) \
  X(perv_actet, 0) X(perv_actet_hru, 0) X(potet_lower, 0) X(potet_rechr, 0) \
  X(pref_flow, 0) X(pref_flow_in, 0) X(pref_flow_infil, 0) \
  X(pref_flow_max, 0) X(pref_flow_stor, 0) \
  X(pref_flow_stor_change, ORC_NAN) X(pref_flow_stor_prev, ORC_NAN) \
  X(pref_flow_thrsh, 0) X(recharge, 0) X(slow_flow, 0) X(slow_stor, 0) \
  X(slow_stor_change, 0) X(slow_stor_prev, ORC_NAN) X(soil_lower, ORC_NAN) \
  X(soil_lower_change, ORC_NAN) X(soil_lower_change_hru, ORC_NAN) \
  X(soil_lower_prev, ORC_NAN) X(soil_lower_ratio, 0) \
  X(soil_lower_max, ORC_NAN) X(soil_moist, ORC_NAN) \
  X(soil_moist_tot, ORC_NAN) X(soil_rechr, ORC_NAN) \
  X(soil_rechr_change, ORC_NAN) X(soil_rechr_change_hru, ORC_NAN) \
  X(soil_rechr_prev, ORC_NAN) X(soil_to_gw, 0) X(soil_to_ssr, 0) \
  X(soil_zone_max, ORC_NAN) X(ssr_to_gw, 0) X(ssres_flow, 0) \
  X(ssres_flow_vol, ORC_NAN) X(ssres_in, 0) X(ssres_stor, ORC_NAN) \
  X(swale_actet, 0) X(unused_potet, 0) \
  /* PRMSGroundwater (prms_groundwater.py:134-143) */ \
  X(gwres_flow, ORC_NAN) X(gwres_flow_vol, ORC_NAN) X(gwres_sink, ORC_NAN) \
  X(gwres_stor, ORC_NAN) X(gwres_stor_old, ORC_NAN) \
  X(gwres_stor_change, ORC_NAN) \
  /* PRMSChannel, HRU-dimensioned (prms_channel.py:153-156) */ \
  X(channel_sroff_vol, ORC_NAN) X(channel_ssres_flow_vol, ORC_NAN) \
  X(channel_gwres_flow_vol, ORC_NAN)

/* ---- public per-segment variables (prms_channel.py:157-162) --------- */
#define ORC_SEG_VARS(X) \
  X(channel_outflow_vol, ORC_NAN) X(seg_lateral_inflow, 0) \
  X(seg_upstream_inflow, 0) X(seg_outflow, 0) X(seg_stor_change, 0)

enum {
#define X(name, init) ORC_V_##name,
  ORC_HRU_VARS(X)
#undef X
  ORC_N_HRU_VARS
};
enum {
#define X(name, init) ORC_S_##name,
  ORC_SEG_VARS(X)
#undef X
  ORC_N_SEG_VARS
};

typedef struct Orc Orc;

#ifdef __cplusplus
extern "C" {
#endif

Orc *orc_create(int nhru, int nseg, int ndeplval);
void orc_destroy(Orc *o);
/* setters copy the data; return 0 on success, -1 unknown name / bad size */
int orc_set_f8(Orc *o, const char *name, const double *v, int64_t n);
int orc_set_i4(Orc *o, const char *name, const int32_t *v, int64_t n);
/* options: "dprst_flag", "temp_units", "start_month", "start_doy",
 * "albset_rna|rnm|sna|snm" go through orc_set_f8 with n == 1 */
const char *orc_last_error(Orc *o);

/* derive everything the reference derives at construction; must be called
 * once after all parameters are set and before any run call.  */
int orc_init(Orc *o);

int orc_n_hru_vars(void);
int orc_n_seg_vars(void);
const char *orc_hru_var_name(int i);
const char *orc_seg_var_name(int i);

/* soltab tables [366][nhru] (valid after orc_init) */
const double *orc_soltab_potsw(Orc *o);
const double *orc_soltab_horad_potsw(Orc *o);
const double *orc_soltab_sunhrs(Orc *o);

/*
 * Advance the full NHM chain `ndays` days from the current state.
 *   cal       [ndays][4]  int32: doy, dowy, month, day-of-month
 *   prcp/tmax/tmin [ndays][nhru]
 *   hru_idx   [n_hru_out] variable indices to record (ORC_V_*)
 *   hru_out   [ndays][n_hru_out][nhru]   (ints/bools stored as doubles)
 *   seg_idx / seg_out analogous with [nseg]
 *   budget_viol [ndays][6] number of units violating each process budget
 *               (canopy, snow, runoff, soilzone, groundwater, channel(global))
 * Any output pointer may be NULL.
 */
int orc_run(Orc *o, int ndays, const int32_t *cal, const double *prcp,
            const double *tmax, const double *tmin, int n_hru_out,
            const int32_t *hru_idx, double *hru_out, int n_seg_out,
            const int32_t *seg_idx, double *seg_out, int32_t *budget_viol);

/* Channel alone: lateral inflows given per day [ndays][nseg] (config 5) */
int orc_run_channel(Orc *o, int ndays, const double *seg_lateral_inflow,
                    int n_seg_out, const int32_t *seg_idx, double *seg_out);
/* test aid: override the Muskingum coefficients derived at orc_init */
int orc_set_channel_coefficients(Orc *o, const int32_t *tsi, const double *ts, const double *c0,
                                 const double *c1, const double *c2);
/* FlowGraph (base/flow_graph.py) over this handle's segments as nodes: kind[j]
 * 0 PRMSChannelFlowNode, 1 PassThroughFlowNode, 2 ObsInFlowNode (obs[t][j]),
 * 3 SourceSinkFlowNode (obs[t][j] = the day's requested source / sink,
 * flow_min[j], may be NULL = 0); out [ndays][7][nnodes] in
 * FlowGraph.get_init_values order. */
int orc_run_flow_graph(Orc *o, int ndays, const double *inflows, const int32_t *kind,
                       const double *obs, const double *flow_min, double *out);

/* current value of one public variable (pointer into oracle storage) */
const double *orc_hru_var(Orc *o, int idx);
/* Overwrite the current value of one public HRU variable ([nhru]; integer
 * variables as doubles).  Every prognostic quantity of the column is a public
 * variable (the reference's state IS its public arrays, process.py:324-372), so
 * setting all of them makes the oracle continue from somebody else's day: the
 * one-step ("teacher-forced") parity check of tests/parity.py. */
int orc_set_hru_var(Orc *o, int idx, const double *v);
const double *orc_seg_var(Orc *o, int idx);

#ifdef __cplusplus
}
#endif
#endif
