/*
 * pws_b200.h -- C ABI of libpwsb200.so, the B200 (sm_100a) implementation of
 * pywatershed's per-timestep PRMS/NHM process chain.
 *
 * The reference has no FFI for this path: it is pure Python, each process a
 * class whose calculate() runs a scalar loop over HRUs
 * (pywatershed/base/process.py:407-422 -> e.g.
 * pywatershed/hydrology/prms_snow.py:623).  The entry points below are what a
 * ctypes binding of that path needs; each cites the reference interface it
 * stands in for.  Paths are relative to /root/reference/pywatershed/.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function
 * returns 0 on success and a non-zero code on failure, after which
 * pws_last_error(h) describes it (the Python wrapper raises the exception
 * class the reference would: ValueError / RuntimeError).  A handle belongs to
 * one host thread and one GPU.  Arrays named d_* are DEVICE pointers (e.g.
 * torch.Tensor.data_ptr()), h_* are HOST pointers (pinned for real overlap).
 *
 * There is no CPU fallback: pws_create fails when no CUDA device is present.
 */
#ifndef PWS_B200_H
#define PWS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pws_handle pws_handle;

/* ---- registry (names follow static/metadata/variables.yaml) ------------ */
int pws_n_hru_vars(void);               /* 143 public [nhru] variables       */
const char *pws_hru_var_name(int idx);
int pws_hru_var_kind(int idx);          /* 0 float64, 1 int32, 2 bool        */
int pws_n_seg_vars(void);               /* 5 public [nsegment] variables     */
const char *pws_seg_var_name(int idx);
int pws_n_hru_state(void);              /* prognostic state, restart-ready   */
const char *pws_hru_state_name(int idx);
int pws_abi_version(void);

/* ---- lifetime ---------------------------------------------------------- */
/* Process.__init__(control, discretization, parameters): base/process.py:91 */
int pws_create(pws_handle **out, int device, int64_t nhru, int64_t nsegment,
               int ndeplval);
int pws_destroy(pws_handle *h);
const char *pws_last_error(pws_handle *h); /* h may be NULL: last create error */

/* ---- parameters: Process._set_params, base/process.py:219-241 -----------
 * name = the reference's parameter name; n = nhru | 12*nhru | nsegment |
 * ndeplval | 1 (1 broadcasts, as prms_snow.py:302-320 does for den_init...). */
int pws_set_param_f64(pws_handle *h, const char *name, const double *h_v, int64_t n);
int pws_set_param_i32(pws_handle *h, const char *name, const int32_t *h_v, int64_t n);
/* options: "dprst_flag" (Process._set_options, base/process.py:339-360),
 * "temp_units", "start_month", "start_doy" (control.start_month/start_doy,
 * prms_atmosphere.py:695-696), "adjust_parameters" (prms_soilzone.py:100),
 * "albset_rna|rnm|sna|snm" (scalar parameters, prms_snow.py:911-914),
 * "block_days" (days per device time block of pws_run_host),
 * "tma" (1, default: the column kernel stages each day's forcing rows in shared
 * memory with TMA bulk copies when they are 16-byte aligned; 0: per-thread loads),
 * "column_hrus" (HRUs per CTA of the column kernel: 0, default = by domain size,
 * 128, 192 (one uncapped-register CTA per SM; picked for domains of 128..192 HRUs per SM)
 * or 256),
 * "soltab_device" (1, default: build the PRMSSolarGeometry tables with the CUDA
 * kernel k_soltab; 0: on host threads with the host libm the reference's tables
 * come from -- a debugging aid for last-bit comparisons, not a product path),
 * "forcing_period" (P > 0: a parameter ensemble over one base domain, member-
 * major -- nhru = nmember * P and the forcing arrays of every pws_run_* call are
 * [ndays][P], HRU i reading column i mod P, so the members' common weather is
 * uploaded and read once (examples/05_parameter_ensemble.ipynb cells 20-26
 * re-read it per member); 0, default: forcings are [ndays][nhru]),
 * "lanes" (column-kernel launches per block of days, each over a contiguous
 * range of CTA tiles on its own stream: 0, default = by domain size, 1..8),
 * "overlap_routing" (1: the routing of block b runs beside the column kernels
 * of block b+1; 0: blocks are serialised; -1, default: by domain size -- on when
 * the column grid leaves SMs idle or lanes are in use),
 * "channel_chunk" (segments per routing chunk = per CTA of k_route_chunks,
 * default and maximum 512; outlet trees larger than a chunk are cut and linked
 * through HBM),
 * "route_hours" (hours a segment advances per step of k_route_chunks: 0,
 * default = 12 for segment graphs of at most 160 levels, 2 for more than 2,000
 * levels, else 4),
 * "route_overlap" (1: the routing launch of block b+1 may start while that of
 * block b still runs, on a second stream, each chunk of connected segments
 * continuing from the state the same chunk of the previous launch stored -- a deep
 * network then pays its pipeline fill once per run, not per block; only while all
 * chunks of a launch fit the GPU with room to spare; 0: launches are serialised;
 * -1, default: on for segment graphs deeper than 160 levels),
 * "nhm_sink" (1, default: the reference's default output selection -- the 71 HRU
 * variables of drb_2yr/nhm.control's netcdf_output_var_names, selected in registry
 * order -- runs the kernel variant that knows this selection at compile time),
 * "wide_column" / "one_cta_carveout" (1, default: grids of at most one CTA per SM
 * run the column kernel variant without a register cap, with the shared-memory
 * carve-out of one CTA). */
int pws_set_option(pws_handle *h, const char *name, double value);

/* Everything the reference derives at construction: soltab tables
 * (atmosphere/prms_solar_geometry.py:138-161), basin_init/dprst_init
 * (hydrology/prms_runoff.py:253-389), _initialize_soilzone_data
 * (prms_soilzone.py:261-536), snow/groundwater initial conditions
 * (prms_snow.py:295-431, prms_groundwater.py:145-149), the transp_on start-up
 * (prms_atmosphere.py:690-716) and _initialize_channel_data
 * (prms_channel.py:189-342).  Resets state to the initial conditions. */
int pws_init(pws_handle *h);

/* soltab_potsw (0), soltab_horad_potsw (1), soltab_sunhrs (2): [366][nhru] */
int pws_get_soltab(pws_handle *h, int which, double *h_out);

/* ---- output selection: control.options["netcdf_output_var_names"],
 * base/process.py:463-569.  Indices into the registries above; outputs of a
 * run are laid out [day][selected var][unit] in selection order, float64
 * variables in out_f64, int32/bool variables (as int32) in out_i32. */
int pws_select_outputs(pws_handle *h, int n_hru_vars, const int32_t *hru_var_idx,
                       int n_seg_vars, const int32_t *seg_var_idx);
int pws_n_selected(pws_handle *h, int *n_f64, int *n_i32, int *n_seg);

/* ---- budget accumulations: Budget.calculate, base/budget.py:262-269
 * (`self._accumulations[component][var] += value * dt` every step).  The
 * chosen float64 variables (indices into the registries) are summed per unit
 * on the device over every day the handle computes, in day order; they need
 * not be part of the drained selection (the kernels write them behind it: the
 * device-side rows per day are what pws_n_selected reports, pws_run_host only
 * drains the selected ones).  pws_get_accumulations: h_hru [n_hru_vars][nhru],
 * h_seg [n_seg_vars][nsegment] (either may be NULL).  pws_reset_state and
 * pws_clear_accumulations zero the sums; option "accumulate" 0 pauses them. */
int pws_select_accumulations(pws_handle *h, int n_hru_vars, const int32_t *hru_var_idx,
                             int n_seg_vars, const int32_t *seg_var_idx);
int pws_get_accumulations(pws_handle *h, double *h_hru, double *h_seg);
int pws_clear_accumulations(pws_handle *h);

/* ---- the time loop: Model.run / advance / calculate,
 * base/model.py:678-743.  cal = [ndays][4] int32 (doy, dowy, month,
 * day-of-month) = control.current_doy/current_dowy/current_month
 * (base/control.py:439-453, utils/time_utils.py:25-68).
 * budget_viol [ndays][6] int32 (may be NULL): number of units whose
 * Budget.calculate closure fails (base/budget.py:336-416) for canopy, snow,
 * runoff, soilzone, groundwater (unit basis) and channel (global basis). */

/* forcings and outputs already resident in HBM; asynchronous (pws_sync).
 * h_budget_viol must stay valid (and should be page-locked) until pws_sync. */
int pws_run_device(pws_handle *h, int ndays, const int32_t *h_cal,
                   const double *d_prcp, const double *d_tmax, const double *d_tmin,
                   double *d_out_f64, int32_t *d_out_i32, double *d_seg_out,
                   int32_t *h_budget_viol);

/* forcings and outputs in host memory, processed in time blocks with H2D /
 * compute / D2H overlapped on three streams; returns when the last block has
 * been drained.  Page-locked caller buffers are used directly; pageable ones
 * are staged through the library's own pinned slots (inputs and outputs
 * independently). */
int pws_run_host(pws_handle *h, int ndays, const int32_t *h_cal,
                 const double *h_prcp, const double *h_tmax, const double *h_tmin,
                 double *h_out_f64, int32_t *h_out_i32, double *h_seg_out,
                 int32_t *h_budget_viol);

/* The same with the forcings as float32 arrays [ndays][nhru]: the storage type
 * of the reference's forcing files (test_data/{drb_2yr,ucb_2yr,hru_1}/
 * {prcp,tmax,tmin}.nc are NC_FLOAT; numpy promotes them to float64 in the
 * reference's first arithmetic on them, prms_atmosphere.py:311-312,399).  Half
 * the bytes cross PCIe; the rows are widened on the device, which is exact,
 * so the results equal pws_run_host on the widened arrays bit for bit. */
int pws_run_host_f32(pws_handle *h, int ndays, const int32_t *h_cal,
                     const float *h_prcp, const float *h_tmax, const float *h_tmin,
                     double *h_out_f64, int32_t *h_out_i32, double *h_seg_out,
                     int32_t *h_budget_viol);

/* A SUBSET of the chain's processes, the others replaced by the caller's
 * arrays: what the reference does when a process (or a partial chain) is
 * driven from files, e.g. autotest/test_self_drive.py:83-115,
 * autotest/test_prms_snow.py:96-132, the two-process example of
 * base/model.py:121-163.  Input registry: index 0..2 = prcp, tmax, tmin, then
 * the inter-process inputs of Process.get_inputs() (base/process.py:160-163).
 * h_in[k] = [ndays][nhru] float64 (integer inputs as float64).  process_mask:
 * bit 0 canopy, 1 snow, 2 runoff, 3 soilzone, 4 groundwater, 5 channel --
 * budgets / errors / routing of the processes that are part of the model. */
int pws_n_inputs(void);
const char *pws_input_name(int idx);
int pws_run_host_inputs(pws_handle *h, int ndays, const int32_t *h_cal, int n_in,
                        const int32_t *in_idx, const double *const *h_in,
                        uint32_t process_mask, double *h_out_f64, int32_t *h_out_i32,
                        double *h_seg_out, int32_t *h_budget_viol);

/* PRMSChannel alone (examples/05_parameter_ensemble.ipynb cell 20):
 * d_seg_lateral_inflow [ndays][nsegment] */
int pws_run_channel_device(pws_handle *h, int ndays, const double *d_seg_lateral_inflow,
                           double *d_seg_out);

/* FlowGraph (base/flow_graph.py:131-657): the handle's segments are the NODES of
 * a flow graph (tosegment = to_graph_index + 1; the channel parameters of a node
 * that is not a Muskingum node are stand-ins).  pws_set_flow_graph, before
 * pws_init: kind[node] = 0 PRMSChannelFlowNode (prms_channel_flow_graph.py:30-160),
 * 1 PassThroughFlowNode (pass_through_flow_node.py:6-53), 2 ObsInFlowNode
 * (obsin_flow_node.py:8-90, obs_col[node] = its column of the observations),
 * 3 SourceSinkFlowNode (source_sink_flow_node.py:10-125, obs_col[node] = its column
 * of the requested sources (+) / sinks (-), pws_set_flow_node_params: its flow_min);
 * kind == NULL turns the handle back into a PRMSChannel one.
 * pws_run_flow_graph_host = FlowGraph.calculate for ndays days: inflows
 * [ndays][nnodes] (the `inflows` input), obs [ndays][n_obs] (the day's observed flow
 * of every ObsIn node, negative = pass the inflow on; the day's request of every
 * SourceSink node; NULL without such nodes),
 * out [ndays][5][nnodes]: outflows, node_upstream_inflows, node_outflows,
 * node_storage_changes, node_sink_source (node_negative_sink_source is its
 * negative, node_storages is NaN for these node kinds: flow_graph.py:629-650).
 * StarfitFlowNode (hydrology/starfit.py) is not built. */
int pws_set_flow_graph(pws_handle *h, const int32_t *kind, const int32_t *obs_col, int n_obs);
/* flow_min [nnodes] (SourceSinkFlowNodeMaker's `flow_min` parameter,
 * source_sink_flow_node.py:155-159; read for kind 3 only); after
 * pws_set_flow_graph, before pws_init; NULL = zero everywhere */
int pws_set_flow_node_params(pws_handle *h, const double *flow_min);
int pws_run_flow_graph_host(pws_handle *h, int ndays, const double *inflows, const double *obs,
                            double *out);

/* PRMSEt (hydrology/prms_et.py:22-181), the bookkeeping process that turns the
 * other processes' evaporative losses into hru_actet / unused_potet: inputs
 * h_* [ndays][nhru] as PRMSEt.get_inputs() (:101-116), parameters [nhru]
 * (:88-99), outputs [ndays][nhru] as get_variables() (:118-126).  Stateless,
 * needs no handle; on failure pws_last_error(NULL) tells why. */
int pws_et_host(int device, int ndays, int64_t nhru, const double *h_potet,
                const double *h_hru_impervevap, const double *h_hru_intcpevap,
                const double *h_snow_evap, const double *h_dprst_evap_hru,
                const double *h_perv_actet, const double *h_hru_percent_imperv,
                const double *h_dprst_frac, double *h_unused_potet,
                double *h_hru_perv_actet, double *h_hru_actet);

int pws_sync(pws_handle *h);

/* Tiled result layout of pws_run_device / pws_run_host[_f32], chosen with
 * pws_set_option(h, "tiled_outputs", 1) when every HRU variable is selected in
 * registry order: out_f64 is [day][tile][n_f64][tile_units], out_i32
 * [day][tile][n_i32][tile_units], tile = unit / tile_units, the last tile
 * padded (padded_units = tiles * tile_units >= nhru units per variable and
 * day; the padding is never written).  A variable's row inside a tile is then
 * a compile-time offset for the kernel: one store instruction per value, no
 * address arithmetic.  The reference has no counterpart (its variables are
 * separate [nhru] numpy arrays, base/process.py:324-372); the Python Engine
 * hands out per-variable strided views of this buffer. */
int pws_output_tile(pws_handle *h, int *tile_units, int64_t *padded_units);

/* ---- state: get_restart_variables (prms_snow.py:269-293 etc.) ----------
 * name = a per-HRU state of the registry above ([nhru]) or one of PRMSChannel's
 * two per-segment states ([nsegment], the caller's segment order): "outflow_ts"
 * and "seg_inflow", what _advance_variables carries from day to day
 * (prms_channel.py:384-386, 456-585).  Saving every state after a pws_run_* call
 * and restoring it into a fresh handle continues the run bit for bit. */
int pws_n_seg_state(void);
const char *pws_seg_state_name(int idx);
int pws_get_state(pws_handle *h, const char *name, double *h_out);
int pws_set_state(pws_handle *h, const char *name, const double *h_in);

/* ---- instrumentation ---------------------------------------------------- */
/* kernels launched since create, and device milliseconds spent in the hru
 * column kernels / channel kernels since the last pws_sync: the length of the
 * UNION of the kernels' (start, stop) CUDA-event intervals on their streams
 * (the lanes of a block, and routing beside the next block, run concurrently) */
int64_t pws_launch_count(pws_handle *h);
int pws_last_timing(pws_handle *h, double *ms_column, double *ms_channel);
int pws_column_kernel_info(pws_handle *h, int *regs, int *local_bytes, int *max_blocks_per_sm);
/* CUDA events on the library's compute stream (torch.cuda.Event only sees
 * torch's stream): mark which = 0 (start) / 1 (stop); elapsed after pws_sync */
int pws_timer_mark(pws_handle *h, int which);
int pws_timer_elapsed_ms(pws_handle *h, double *ms);
/* number of hru-column kernel launches and channel kernel launches in the last
 * run; blocks of days it was cut into and column launches (lanes) per block */
int pws_last_launches(pws_handle *h, int *n_column, int *n_channel);
int pws_last_blocks(pws_handle *h, int *n_blocks, int *lanes);
/* reset prognostic state to the initial conditions without rebuilding the
 * static tables (what re-instantiating the reference's processes would do) */
int pws_reset_state(pws_handle *h);

#ifdef __cplusplus
}
#endif
#endif
