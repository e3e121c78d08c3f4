// pws_registry.h -- names of every parameter, public variable and prognostic
// state of the PRMS/NHM column, in the reference's own vocabulary
// (pywatershed/static/metadata/{parameters,variables}.yaml; per-process lists:
// get_parameters()/get_init_values() of atmosphere/prms_atmosphere.py:214-262,
// hydrology/prms_canopy.py:107-146, prms_snow.py:166-255, prms_runoff.py:151-226,
// prms_soilzone.py:140-233, prms_groundwater.py:103-143, prms_channel.py:132-162).
//
// X-macro registries: one line per name, expanded into the device SoA pointer
// tables, the name lookup of the C ABI and the output-slot table.
#pragma once

// ---- per-HRU float64 parameters [nhru] --------------------------------
#define PWS_HRU_F64_PARAMS(X)                                                 \
  X(hru_slope) X(hru_aspect) X(hru_lat) X(hru_area) X(hru_in_to_cf)           \
  X(radj_sppt) X(radj_wppt) X(jh_coef_hru) X(transp_tmax)                     \
  X(covden_sum) X(covden_win) X(srain_intcp) X(wrain_intcp) X(snow_intcp)     \
  X(potet_sublim) X(emis_noppt) X(freeh2o_cap) X(rad_trncf) X(snarea_thresh)  \
  X(den_init) X(den_max) X(settle_const) X(snowpack_init)                     \
  X(hru_percent_imperv) X(imperv_stor_max) X(dprst_frac) X(carea_max)         \
  X(smidx_coef) X(smidx_exp) X(soil_moist_max) X(snowinfil_max)               \
  X(dprst_depth_avg) X(dprst_et_coef) X(dprst_flow_coef) X(dprst_frac_init)   \
  X(dprst_frac_open) X(dprst_seep_rate_clos) X(dprst_seep_rate_open)          \
  X(sro_to_dprst_imperv) X(sro_to_dprst_perv) X(va_open_exp) X(va_clos_exp)   \
  X(op_flow_thres) X(fastcoef_lin) X(fastcoef_sq) X(pref_flow_den)            \
  X(pref_flow_infil_frac) X(sat_threshold) X(slowcoef_lin) X(slowcoef_sq)     \
  X(soil_moist_init_frac) X(soil_rechr_init_frac) X(soil_rechr_max_frac)      \
  X(soil2gw_max) X(ssr2gw_exp) X(ssr2gw_rate) X(ssstor_init_frac)             \
  X(gwflow_coef) X(gwsink_coef) X(gwstor_init)

// ---- per-HRU integer parameters (int32 on the device) -----------------
#define PWS_HRU_I32_PARAMS(X)                                                 \
  X(transp_beg) X(transp_end) X(cov_type) X(hru_type) X(hru_deplcrv)          \
  X(melt_force) X(melt_look) X(soil_type) X(hru_segment)

// ---- monthly parameters [12][nhru] -------------------------------------
#define PWS_MON_F64_PARAMS(X)                                                 \
  X(radadj_intcp) X(radadj_slope) X(tmax_index) X(dday_slope) X(dday_intcp)   \
  X(radmax) X(ppt_rad_adj) X(tmax_allsnow) X(tmax_allrain_offset) X(jh_coef)  \
  X(tmax_cbh_adj) X(tmin_cbh_adj) X(snow_cbh_adj) X(rain_cbh_adj)             \
  X(adjmix_rain) X(cecn_coef)
#define PWS_MON_I32_PARAMS(X) X(tstorm_mo)

// ---- per-segment parameters [nsegment] ----------------------------------
#define PWS_SEG_F64_PARAMS(X)                                                 \
  X(mann_n) X(seg_depth) X(seg_length) X(seg_slope) X(x_coef)                 \
  X(segment_flow_init)
#define PWS_SEG_I32_PARAMS(X) X(segment_type) X(tosegment)

// ---- derived per-HRU constants, computed once on the device by k_hru_init
// (the reference derives them in basin_init/dprst_init prms_runoff.py:253-389,
// _initialize_soilzone_data prms_soilzone.py:261-536, prms_snow.py:295-325,
// prms_atmosphere.py:519) ------------------------------------------------
#define PWS_HRU_F64_DERIVED(X)                                                \
  X(hru_cossl) X(deninv) X(denmaxinv)                                         \
  X(ro_hru_perv) X(ro_hru_frac_perv) X(ro_hru_imperv) X(dprst_area_max)       \
  X(dprst_frac_clos) X(dprst_vol_open_max) X(dprst_vol_clos_max)              \
  X(dprst_vol_thres_open) X(dprst_area_open_max) X(dprst_area_clos_max)       \
  X(dprst_area_clos_init)                                                     \
  X(sz_hru_frac_perv) X(sz_soil_rechr_max) X(sz_soil_moist_max)               \
  X(sz_sat_threshold) X(sz_pref_flow_den) X(pref_flow_thrsh) X(pref_flow_max) \
  X(soil_zone_max) X(soil_lower_max) X(soil_lower_ratio_init)

// ---- public per-HRU variables: X(name, kind, init); kind F = float64,
// I = int32 (meta "int32"), B = bool.  Order = process order. -------------
#define PWS_NAN (__builtin_nan(""))
#define PWS_HRU_VARS(X)                                                       \
  /* PRMSAtmosphere */                                                        \
  X(tmaxf, F, PWS_NAN) X(tminf, F, PWS_NAN) X(prmx, F, PWS_NAN)               \
  X(hru_ppt, F, PWS_NAN) X(hru_rain, F, PWS_NAN) X(hru_snow, F, PWS_NAN)      \
  X(swrad, F, PWS_NAN) X(potet, F, PWS_NAN) X(transp_on, I, 0)                \
  X(tmaxc, F, PWS_NAN) X(tavgc, F, PWS_NAN) X(tminc, F, PWS_NAN)              \
  X(pptmix, I, -9999) X(orad_hru, F, PWS_NAN)                                 \
  /* PRMSCanopy */                                                            \
  X(net_ppt, F, 0) X(net_rain, F, 0) X(net_snow, F, 0)                        \
  X(intcp_changeover, F, 0) X(intcp_evap, F, 0) X(intcp_form, I, 0)           \
  X(intcp_stor, F, 0) X(intcp_transp_on, I, 0) X(hru_intcpevap, F, 0)         \
  X(hru_intcpstor, F, 0) X(hru_intcpstor_change, F, 0)                        \
  X(hru_intcpstor_old, F, 0)                                                  \
  /* PRMSSnow */                                                              \
  X(ai, F, 0) X(albedo, F, 0) X(frac_swe, F, 0) X(freeh2o, F, 0)              \
  X(freeh2o_change, F, PWS_NAN) X(freeh2o_prev, F, PWS_NAN) X(iasw, B, 0)     \
  X(int_alb, I, 1) X(iso, I, 1) X(lso, I, 0) X(lst, I, 0) X(mso, I, 1)        \
  X(newsnow, I, 0) X(pk_def, F, 0) X(pk_den, F, 0) X(pk_depth, F, 0)          \
  X(pk_ice, F, 0) X(pk_ice_change, F, PWS_NAN) X(pk_ice_prev, F, PWS_NAN)     \
  X(pk_precip, F, 0) X(pk_temp, F, 0) X(pksv, F, 0)                           \
  X(pkwater_equiv, F, PWS_NAN) X(pptmix_nopack, I, 0) X(pss, F, PWS_NAN)      \
  X(pst, F, PWS_NAN) X(salb, F, 0) X(scrv, F, 0) X(slst, F, 0)                \
  X(snow_evap, F, 0) X(snowcov_area, F, 0) X(snowcov_areasv, F, 0)            \
  X(snowmelt, F, 0) X(snsv, F, 0) X(tcal, F, 0) X(through_rain, F, PWS_NAN)   \
  /* PRMSRunoff */                                                            \
  X(contrib_fraction, F, 0) X(infil, F, 0) X(infil_hru, F, 0) X(sroff, F, 0)  \
  X(sroff_vol, F, 0) X(hru_sroffp, F, 0) X(hru_sroffi, F, 0)                  \
  X(imperv_stor, F, 0) X(imperv_evap, F, 0) X(hru_impervevap, F, 0)           \
  X(hru_impervstor, F, 0) X(hru_impervstor_old, F, 0)                         \
  X(hru_impervstor_change, F, 0) X(dprst_vol_frac, F, 0)                      \
  X(dprst_vol_clos, F, 0) X(dprst_vol_open, F, 0)                             \
  X(dprst_vol_clos_frac, F, 0) X(dprst_vol_open_frac, F, 0)                   \
  X(dprst_area_clos, F, 0) X(dprst_area_open, F, 0)                           \
  X(dprst_area_clos_max, F, 0) X(dprst_area_open_max, F, 0)                   \
  X(dprst_sroff_hru, F, 0) X(dprst_seep_hru, F, 0) X(dprst_evap_hru, F, 0)    \
  X(dprst_insroff_hru, F, 0) X(dprst_stor_hru, F, 0)                          \
  X(dprst_stor_hru_old, F, 0) X(dprst_stor_hru_change, F, 0)                  \
  /* PRMSSoilzone */                                                          \
  X(cap_infil_tot, F, 0) X(cap_waterin, F, 0) X(dunnian_flow, F, 0)           \
  X(hru_actet, F, 0) X(perv_actet, F, 0) X(perv_actet_hru, F, 0)              \
  X(potet_lower, F, 0) X(potet_rechr, F, 0) X(pref_flow, F, 0)                \
  X(pref_flow_in, F, 0) X(pref_flow_infil, F, 0) X(pref_flow_max, F, 0)       \
  X(pref_flow_stor, F, 0) X(pref_flow_stor_change, F, PWS_NAN)                \
  X(pref_flow_stor_prev, F, PWS_NAN) X(pref_flow_thrsh, F, 0)                 \
  X(recharge, F, 0) X(slow_flow, F, 0) X(slow_stor, F, 0)                     \
  X(slow_stor_change, F, 0) X(slow_stor_prev, F, PWS_NAN)                     \
  X(soil_lower, F, PWS_NAN) X(soil_lower_change, F, PWS_NAN)                  \
  X(soil_lower_change_hru, F, PWS_NAN) X(soil_lower_prev, F, PWS_NAN)         \
  X(soil_lower_ratio, F, 0) X(soil_lower_max, F, PWS_NAN)                     \
  X(soil_moist, F, PWS_NAN) X(soil_moist_tot, F, PWS_NAN)                     \
  X(soil_rechr, F, PWS_NAN) X(soil_rechr_change, F, PWS_NAN)                  \
  X(soil_rechr_change_hru, F, PWS_NAN) X(soil_rechr_prev, F, PWS_NAN)         \
  X(soil_to_gw, F, 0) X(soil_to_ssr, F, 0) X(soil_zone_max, F, PWS_NAN)       \
  X(ssr_to_gw, F, 0) X(ssres_flow, F, 0) X(ssres_flow_vol, F, PWS_NAN)        \
  X(ssres_in, F, 0) X(ssres_stor, F, PWS_NAN) X(swale_actet, F, 0)            \
  X(unused_potet, F, 0)                                                       \
  /* PRMSGroundwater */                                                       \
  X(gwres_flow, F, PWS_NAN) X(gwres_flow_vol, F, PWS_NAN)                     \
  X(gwres_sink, F, PWS_NAN) X(gwres_stor, F, PWS_NAN)                         \
  X(gwres_stor_old, F, PWS_NAN) X(gwres_stor_change, F, PWS_NAN)              \
  /* PRMSChannel, HRU-dimensioned */                                          \
  X(channel_sroff_vol, F, PWS_NAN) X(channel_ssres_flow_vol, F, PWS_NAN)      \
  X(channel_gwres_flow_vol, F, PWS_NAN)

// ---- the HRU variables of the reference's default output selection
// (`netcdf_output_var_names` of pywatershed/data/drb_2yr/nhm.control intersected with
// the public variables: tests/golden/nhm_default_vars.json), in registry order: the
// compile-time selection of k_hru_column<SINK_NHM> -------------------------
#define PWS_NHM_DEFAULT_HRU_VARS(X)                                           \
  X(tmaxf) X(tminf) X(prmx) X(hru_ppt) X(hru_rain) X(hru_snow) X(swrad)       \
  X(potet) X(transp_on) X(tmaxc) X(tavgc) X(tminc) X(pptmix) X(orad_hru)      \
  X(net_ppt) X(net_rain) X(net_snow) X(intcp_changeover) X(intcp_evap)        \
  X(intcp_stor) X(hru_intcpevap) X(hru_intcpstor) X(albedo) X(freeh2o)        \
  X(iso) X(newsnow) X(pk_ice) X(pkwater_equiv) X(pptmix_nopack) X(pst)        \
  X(salb) X(snow_evap) X(snowcov_area) X(snowmelt) X(tcal)                    \
  X(contrib_fraction) X(infil) X(sroff) X(hru_sroffp) X(hru_sroffi)           \
  X(hru_impervevap) X(hru_impervstor) X(dprst_sroff_hru) X(dprst_seep_hru)    \
  X(dprst_evap_hru) X(dprst_stor_hru) X(cap_infil_tot) X(hru_actet)           \
  X(perv_actet) X(potet_lower) X(potet_rechr) X(pref_flow)                    \
  X(pref_flow_infil) X(pref_flow_stor) X(recharge) X(slow_flow)               \
  X(slow_stor) X(soil_lower) X(soil_lower_ratio) X(soil_moist)                \
  X(soil_moist_tot) X(soil_rechr) X(soil_to_gw) X(soil_to_ssr) X(ssr_to_gw)   \
  X(ssres_flow) X(ssres_in) X(ssres_stor) X(gwres_flow) X(gwres_sink)         \
  X(gwres_stor)

// ---- public per-segment variables (all float64) -------------------------
#define PWS_SEG_VARS(X)                                                       \
  X(channel_outflow_vol) X(seg_lateral_inflow) X(seg_upstream_inflow)         \
  X(seg_outflow) X(seg_stor_change)

// ---- prognostic per-HRU state carried from day to day: X(name, kind).
// Lives in registers inside the column kernel, in a device SoA between
// launches; pws_get_state/pws_set_state expose it (restart-ready; the
// reference declares but does not implement restart, prms_snow.py:269-293).
#define PWS_HRU_STATE_UPPER(X)                                                \
  X(transp_on, I) X(transp_check, I) X(tmax_sum, F)                           \
  X(intcp_stor, F) X(intcp_transp_on, I) X(hru_intcpstor, F)                  \
  X(albedo, F) X(freeh2o, F) X(iasw, I) X(int_alb, I) X(iso, I) X(lso, I)     \
  X(lst, I) X(mso, I) X(pk_def, F) X(pk_depth, F) X(pk_den, F) X(pk_ice, F)   \
  X(pk_temp, F) X(pksv, F) X(pss, F) X(pst, F) X(salb, F) X(scrv, F)          \
  X(slst, F) X(snowcov_area, F) X(snowcov_areasv, F) X(snsv, F)               \
  X(pkwater_equiv, F)
#define PWS_HRU_STATE_LOWER(X)                                                \
  X(imperv_stor, F) X(hru_impervstor, F) X(dprst_vol_open, F)                 \
  X(dprst_vol_clos, F) X(dprst_area_open, F)                                  \
  X(soil_moist, F) X(soil_rechr, F) X(slow_stor, F) X(pref_flow_stor, F)      \
  X(gwres_stor, F)
// upper half of the column (atmosphere, canopy, snow) first, then the lower
// half (runoff, soilzone, groundwater): k_hru_column gives each half its own warps
#define PWS_HRU_STATE(X) PWS_HRU_STATE_UPPER(X) PWS_HRU_STATE_LOWER(X)


// ---- what the two halves of the column read every day, per HRU (generated
// from the P.<name> uses in column_upper / column_lower, hru_column.cuh);
// k_hru_column stages exactly these in shared memory, one private column per
// thread.  Monthly parameters are only read by the upper half.
#define PWS_UPPER_F64(X) \
  X(covden_sum) X(covden_win) X(den_max) X(deninv) X(denmaxinv) \
  X(emis_noppt) X(freeh2o_cap) X(hru_cossl) X(jh_coef_hru) X(potet_sublim) \
  X(rad_trncf) X(radj_sppt) X(radj_wppt) X(settle_const) X(snarea_thresh) \
  X(snow_intcp) X(srain_intcp) X(transp_tmax) X(wrain_intcp)
#define PWS_UPPER_I32(X) \
  X(cov_type) X(hru_deplcrv) X(hru_type) X(melt_force) X(melt_look) \
  X(transp_beg) X(transp_end)
#define PWS_LOWER_F64(X) \
  X(carea_max) X(dprst_area_clos_init) X(dprst_area_clos_max) \
  X(dprst_area_max) X(dprst_area_open_max) X(dprst_et_coef) \
  X(dprst_flow_coef) X(dprst_frac_clos) X(dprst_frac_open) \
  X(dprst_seep_rate_clos) X(dprst_seep_rate_open) X(dprst_vol_clos_max) \
  X(dprst_vol_open_max) X(dprst_vol_thres_open) X(fastcoef_lin) \
  X(fastcoef_sq) X(gwflow_coef) X(gwsink_coef) X(hru_area) X(hru_in_to_cf) \
  X(hru_percent_imperv) X(imperv_stor_max) X(pref_flow_infil_frac) \
  X(pref_flow_max) X(pref_flow_thrsh) X(ro_hru_frac_perv) X(ro_hru_imperv) \
  X(ro_hru_perv) X(slowcoef_lin) X(slowcoef_sq) X(smidx_coef) X(smidx_exp) \
  X(snowinfil_max) X(soil2gw_max) X(soil_lower_max) \
  X(soil_lower_ratio_init) X(soil_moist_max) X(soil_zone_max) \
  X(sro_to_dprst_imperv) X(sro_to_dprst_perv) X(ssr2gw_exp) X(ssr2gw_rate) \
  X(sz_hru_frac_perv) X(sz_pref_flow_den) X(sz_sat_threshold) \
  X(sz_soil_moist_max) X(sz_soil_rechr_max) X(va_clos_exp) X(va_open_exp)
#define PWS_LOWER_I32(X) \
  X(cov_type) X(hru_segment) X(hru_type) X(soil_type)
#define PWS_UPPER_MON_F64(X) \
  X(adjmix_rain) X(cecn_coef) X(dday_intcp) X(dday_slope) X(jh_coef) \
  X(ppt_rad_adj) X(radadj_intcp) X(radadj_slope) X(radmax) X(rain_cbh_adj) \
  X(snow_cbh_adj) X(tmax_allrain_offset) X(tmax_allsnow) X(tmax_allsnow_c) \
  X(tmax_cbh_adj) X(tmax_index) X(tmin_cbh_adj)
#define PWS_UPPER_MON_I32(X) \
  X(tstorm_mo)

// ---- inputs a process can take from the caller instead of from the process
// upstream of it (names = the reference's get_inputs(); see PWS_OV in
// hru_column.cuh).  transp_on, pptmix, pptmix_nopack are integer-valued.
#define PWS_OV_INPUTS(X) \
  X(tmaxc) X(tminc) X(tavgc) X(prmx) X(hru_ppt) X(hru_rain) X(hru_snow) \
  X(swrad) X(orad_hru) X(potet) X(transp_on) X(pptmix) X(pk_ice_prev) \
  X(freeh2o_prev) X(net_ppt) X(net_rain) X(net_snow) X(hru_intcpevap) \
  X(intcp_changeover) X(snowmelt) X(snow_evap) X(pkwater_equiv) \
  X(pptmix_nopack) X(snowcov_area) X(through_rain) X(soil_lower_prev) \
  X(soil_rechr_prev) X(dprst_evap_hru) X(dprst_seep_hru) X(hru_impervevap) \
  X(infil_hru) X(sroff) X(soil_to_gw) X(ssr_to_gw) X(sroff_vol) \
  X(ssres_flow_vol) X(gwres_flow_vol)
enum PwsOvInput {
#define X(name) OV_##name,
  PWS_OV_INPUTS(X)
#undef X
  PWS_N_OV
};

enum PwsHruVar {
#define X(name, kind, init) V_##name,
  PWS_HRU_VARS(X)
#undef X
  PWS_N_HRU_VARS
};
enum PwsSegVar {
#define X(name) S_##name,
  PWS_SEG_VARS(X)
#undef X
  PWS_N_SEG_VARS
};
enum PwsHruState {
#define X(name, kind) ST_##name,
  PWS_HRU_STATE(X)
#undef X
  PWS_N_HRU_STATE
};
