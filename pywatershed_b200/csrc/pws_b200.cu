// pws_b200.cu -- host side of libpwsb200.so: the C ABI declared in
// include/pws_b200.h, device memory layout, streams and the time-block loop.
//
// Data layout in HBM (per handle):
//   parameters   one float64 slab [n_f64_rows][stride] and one int32 slab,
//                SoA: row = one parameter (monthly ones are 12 rows), stride =
//                nhru rounded up to 32 so every row starts 256-byte aligned
//   state        [PWS_N_HRU_STATE][stride] float64 (flags stored as 0/1/2)
//   soltab       3 x [366][stride]
//   block ring   pws_run_host: 2 x { forcings [T][3][nhru], out_f64
//                [T][n_f64][nhru], out_i32 [T][n_i32][nhru], seg_out }
//   channel      per-segment coefficients/state in level order, the hourly
//                outflow scratch ots[24*T][sstride], hru_lat[T][stride]
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pws_b200.h"
#include "pws_kernels.cuh"

using namespace pws;

namespace {

std::string g_create_error;

struct NameTables {
  std::vector<std::string> hru_f64, hru_i32, mon_f64, mon_i32, seg_f64, seg_i32, derived;
  std::vector<std::string> hru_vars, seg_vars, states;
  std::vector<int> hru_var_kind;
  NameTables() {
#define X(n) hru_f64.push_back(#n);
    PWS_HRU_F64_PARAMS(X)
#undef X
#define X(n) hru_i32.push_back(#n);
    PWS_HRU_I32_PARAMS(X)
#undef X
#define X(n) mon_f64.push_back(#n);
    PWS_MON_F64_PARAMS(X)
#undef X
#define X(n) mon_i32.push_back(#n);
    PWS_MON_I32_PARAMS(X)
#undef X
#define X(n) seg_f64.push_back(#n);
    PWS_SEG_F64_PARAMS(X)
#undef X
#define X(n) seg_i32.push_back(#n);
    PWS_SEG_I32_PARAMS(X)
#undef X
#define X(n) derived.push_back(#n);
    PWS_HRU_F64_DERIVED(X)
#undef X
#define KIND_F 0
#define KIND_I 1
#define KIND_B 2
#define X(n, kind, init) hru_vars.push_back(#n); hru_var_kind.push_back(KIND_##kind);
    PWS_HRU_VARS(X)
#undef X
#define X(n) seg_vars.push_back(#n);
    PWS_SEG_VARS(X)
#undef X
#define X(n, kind) states.push_back(#n);
    PWS_HRU_STATE(X)
#undef X
  }
};
const NameTables& names() {
  static NameTables t;
  return t;
}

int find(const std::vector<std::string>& v, const char* n) {
  for (size_t k = 0; k < v.size(); ++k)
    if (v[k] == n) return (int)k;
  return -1;
}

int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

}  // namespace

struct pws_handle {
  int device = 0;
  int64_t nhru = 0, nseg = 0, stride = 0, sstride = 0;
  int ndeplval = 0;
  std::string err;
  bool inited = false;
  int64_t launches = 0;
  // options
  int dprst_flag = 1, temp_units = 0, start_month = 1, start_doy = 1, adjust_parameters = 1;
  int block_days = 0;
  int soltab_device = 1;
  int tma = 1;  // option: stage forcings with TMA bulk copies when the rows are aligned
  int lanes_opt = 0;        // option "lanes": 0 = by domain size, else column launches per block
  int overlap_routing = -1;  // option "overlap_routing": routing of block b beside the column of b+1 (-1: by domain size)
  int64_t forcing_period = 0;  // option "forcing_period": forcings are [ndays][period], HRU i reads i % period
  bool route_attr_set = false;
  bool smem_attr_set[64] = {};  // (mode 5) x (H 3) x period x wide
  bool nhm_sink = true;     // option "nhm_sink": SINK_NHM for the default output selection
  bool wide_column = true;  // option "wide_column": the uncapped-register column variant on grids of <= one CTA per SM
  bool one_cta_carveout = true;  // option "one_cta_carveout": L1-heavy carve-out for grids of <= one CTA per SM
  bool tiled_outputs = false;  // option "tiled_outputs": [day][tile][variable][H] (SINK_TILED)
  int column_hrus_opt = 0;  // option "column_hrus": 0 = by domain size, else 128 | 256
  double albset[4] = {0, 0, 0, 0};
  // host copies of the parameters (set before pws_init)
  std::map<std::string, std::vector<double>> pf;
  std::map<std::string, std::vector<int32_t>> pi;
  // device slabs
  double* d_f64 = nullptr;
  int32_t* d_i32 = nullptr;
  double* d_state = nullptr;
  double* d_soltab[3] = {nullptr, nullptr, nullptr};
  double* d_snarea = nullptr;
  int* d_err = nullptr;
  HruParams P{};
  HruDerivedMut D{};
  // output selection
  short slot_hru[PWS_N_HRU_VARS];
  short slot_seg[PWS_N_SEG_VARS];
  // n_f64 / n_seg_sel count the slots the kernels write per day: the drained
  // selection first (n_f64_drain, n_seg_drain), then variables that are only
  // accumulated on the device (pws_select_accumulations)
  int n_f64 = 0, n_i32 = 0, n_seg_sel = 0, n_f64_drain = 0, n_seg_drain = 0;
  std::vector<int> sel_hru, sel_seg;   // drained selection, in the caller's order
  std::vector<int> acc_hru, acc_seg;   // accumulated variables (float64), in the caller's order
  std::vector<int> acc_hru_slot, acc_seg_slot;
  double *d_acc_hru = nullptr, *d_acc_seg = nullptr;  // [n][stride], [n][sstride]
  int accumulate_on = 1;  // option "accumulate": 0 pauses the accumulation (values are kept)
  // channel (level order)
  int nlevels = 0, nchunks = 0, n_cross_edges = 0, n_ext_rows = 0, max_dmax = 0;
  int channel_chunk = PWS_CHUNK_THREADS;  // option: segments per routing chunk (<= CTA size)
  int route_hours_opt = 0;                // option "route_hours": 0 = by network depth
  int32_t *d_tsi = nullptr, *d_up_ptr = nullptr, *d_up_idx = nullptr, *d_hru_ptr = nullptr,
          *d_hru_idx = nullptr, *d_seg_flags = nullptr, *d_orig = nullptr,
          *d_chunk_begin = nullptr, *d_chunk_dmax = nullptr, *d_delay = nullptr,
          *d_ext_row = nullptr;
  double* d_rts = nullptr;
  double *d_ts = nullptr, *d_c0 = nullptr, *d_c1 = nullptr, *d_c2 = nullptr;
  double *d_outflow_ts = nullptr, *d_seg_inflow = nullptr;
  // FlowGraph (pws_set_flow_graph): node kinds / observation columns, caller's node order
  bool flow_graph = false;
  std::vector<int32_t> fg_kind, fg_obs_col;
  std::vector<double> fg_flow_min;  // pws_set_flow_node_params: SourceSink nodes' minimum flow
  int fg_n_obs = 0;
  int32_t *d_fg_kind = nullptr, *d_fg_obs_col = nullptr;  // pos order
  double* d_fg_flow_min = nullptr;                         // pos order
  // per-block scratch (sized for scratch_days).  What a block's column kernels
  // write and its routing reads is double-buffered by block parity, so the
  // column of block b+1 runs beside the routing of block b.
  int scratch_days = 0;
  int scratch_copies = 0;  // copies of the routing scratch (2 when routing launches may overlap)
  // (the routing scratch exists twice when consecutive routing launches may overlap,
  // see route_overlap(): index = the block's slot, else 0)
  double *d_hru_lat[2] = {nullptr, nullptr}, *d_ots[2] = {nullptr, nullptr},
         *d_day_outvol[2] = {nullptr, nullptr}, *d_day_dstor[2] = {nullptr, nullptr},
         *d_day_lat[2] = {nullptr, nullptr};
  int4* d_cal[2] = {nullptr, nullptr};
  int* d_viol[2] = {nullptr, nullptr};
  int* d_progress[2] = {nullptr, nullptr};  // [sstride] days published + 1 ticket counter at the end
  int* d_chunk_done = nullptr;  // [nchunks] serial of the last routing launch whose state a chunk has written
  int route_serial = 0;         // routing launches so far
  int route_overlap_opt = -1;   // option "route_overlap": -1 by network, 0 off, 1 on where it is safe
  cudaStream_t s_route2 = nullptr;
  int64_t blocks_enqueued = 0;  // parity = scratch slot of the next block
  // host ring for pws_run_host
  int ring_days = 0;
  double* d_forc[2] = {nullptr, nullptr};
  float* d_forc32[2] = {nullptr, nullptr};  // float32 forcings as uploaded (pws_run_host_f32)
  int forc32_days = 0;
  double* d_of64[2] = {nullptr, nullptr};
  int32_t* d_oi32[2] = {nullptr, nullptr};
  double* d_oseg[2] = {nullptr, nullptr};
  size_t ring_f64_elems = 0, ring_i32_elems = 0, ring_seg_elems = 0;
  // pinned staging for pageable caller buffers (pws_run_host)
  int stage_days = 0;
  size_t stage_f64_elems = 0, stage_i32_elems = 0, stage_seg_elems = 0;
  double* hs_in[2] = {nullptr, nullptr};
  double* hs_of[2] = {nullptr, nullptr};
  int32_t* hs_oi[2] = {nullptr, nullptr};
  double* hs_os[2] = {nullptr, nullptr};
  int32_t* hs_viol[2] = {nullptr, nullptr};
  int32_t* hs_viol_run = nullptr;  // pinned, whole run: the verdicts never block the pipeline
  size_t hs_viol_run_cap = 0;
  // streams / events.  s_compute: per-block preparation (calendar, counters,
  // forcing widening); s_lane[g]: the column kernels of lane g (a contiguous
  // range of CTA tiles), one launch per block each; s_route: routing, channel
  // budget and the verdict copy of a block, after all its lanes.
  static constexpr int kMaxLanes = 8;
  cudaStream_t s_compute = nullptr, s_h2d = nullptr, s_d2h = nullptr, s_route = nullptr;
  cudaStream_t s_lane[kMaxLanes] = {};
  cudaEvent_t ev_prep[2] = {nullptr, nullptr};        // block's scratch slot is prepared
  cudaEvent_t ev_lane[2][kMaxLanes] = {};             // lane g has finished the block in slot p
  cudaEvent_t ev_block[2] = {nullptr, nullptr};       // everything of the block in slot p is done
  cudaEvent_t ev_col0 = nullptr, ev_col1 = nullptr, ev_ch1 = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr},
              ev_d2h[2] = {nullptr, nullptr};
  double ms_column = 0.0, ms_channel = 0.0;
  int n_col_launch = 0, n_ch_launch = 0, last_channel_launches = 0, last_col_blocks = 0,
      n_col_kernels = 0;
  std::vector<double> h_seg_inflow_init;
  std::vector<int32_t> h_orig;  // routing position -> original segment index
  bool timing_pending = false;
  // (start, stop) of every column / channel launch since the last pws_sync; the
  // device time of a kind is the UNION of its intervals (lanes run concurrently)
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pairs_col, ev_pairs_ch;
  cudaEvent_t ev_base = nullptr;  // time origin of the intervals
  double ms_union_all = 0.0;
};

#define PWS_FAIL(h, ...)                                \
  do {                                                  \
    char _b[512];                                       \
    snprintf(_b, sizeof _b, __VA_ARGS__);               \
    (h)->err = _b;                                      \
    return 1;                                           \
  } while (0)

#define PWS_CUDA(h, call)                                                          \
  do {                                                                             \
    cudaError_t _e = (call);                                                       \
    if (_e != cudaSuccess)                                                         \
      PWS_FAIL(h, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,    \
               __LINE__, cudaGetErrorString(_e));                                  \
  } while (0)

// --------------------------------------------------------------- registry
extern "C" int pws_abi_version(void) { return 1; }
extern "C" int pws_n_hru_vars(void) { return PWS_N_HRU_VARS; }
extern "C" const char* pws_hru_var_name(int i) {
  return (i >= 0 && i < PWS_N_HRU_VARS) ? names().hru_vars[i].c_str() : nullptr;
}
extern "C" int pws_hru_var_kind(int i) {
  return (i >= 0 && i < PWS_N_HRU_VARS) ? names().hru_var_kind[i] : -1;
}
extern "C" int pws_n_seg_vars(void) { return PWS_N_SEG_VARS; }
extern "C" const char* pws_seg_var_name(int i) {
  return (i >= 0 && i < PWS_N_SEG_VARS) ? names().seg_vars[i].c_str() : nullptr;
}
extern "C" int pws_n_hru_state(void) { return PWS_N_HRU_STATE; }
extern "C" const char* pws_hru_state_name(int i) {
  return (i >= 0 && i < PWS_N_HRU_STATE) ? names().states[i].c_str() : nullptr;
}

// --------------------------------------------------------------- lifetime
extern "C" const char* pws_last_error(pws_handle* h) {
  return h ? h->err.c_str() : g_create_error.c_str();
}

extern "C" int pws_create(pws_handle** out, int device, int64_t nhru, int64_t nseg,
                          int ndeplval) {
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("pws_create: no CUDA device (") + cudaGetErrorString(e) +
                     "); libpwsb200 has no CPU path";
    return 1;
  }
  if (device < 0 || device >= ndev || nhru <= 0 || nseg < 0 || ndeplval < 11) {
    g_create_error = "pws_create: bad device / nhru / nsegment / ndeplval";
    return 1;
  }
  if ((e = cudaSetDevice(device)) != cudaSuccess) {
    g_create_error = std::string("pws_create: cudaSetDevice: ") + cudaGetErrorString(e);
    return 1;
  }
  pws_handle* h = new pws_handle();
  h->device = device;
  h->nhru = nhru;
  h->nseg = nseg;
  h->stride = round_up(nhru, 32);
  h->sstride = round_up(std::max<int64_t>(nseg, 1), 32);
  h->ndeplval = ndeplval;
  for (int k = 0; k < PWS_N_HRU_VARS; ++k) h->slot_hru[k] = -1;
  for (int k = 0; k < PWS_N_SEG_VARS; ++k) h->slot_seg[k] = -1;
  auto ok = [&](cudaError_t ee) { return ee == cudaSuccess; };
  int prio_lo = 0, prio_hi = 0;  // routing is the short, latency-bound tail of a block: first in line
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  bool good = ok(cudaStreamCreateWithFlags(&h->s_compute, cudaStreamNonBlocking)) &&
              ok(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking)) &&
              ok(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking)) &&
              ok(cudaStreamCreateWithPriority(&h->s_route, cudaStreamNonBlocking, prio_hi)) &&
              ok(cudaStreamCreateWithPriority(&h->s_route2, cudaStreamNonBlocking, prio_hi)) &&
              ok(cudaEventCreate(&h->ev_col0)) && ok(cudaEventCreate(&h->ev_col1)) &&
              ok(cudaEventCreate(&h->ev_ch1)) && ok(cudaEventCreate(&h->ev_base));
  for (int g = 0; g < pws_handle::kMaxLanes && good; ++g)
    good = ok(cudaStreamCreateWithFlags(&h->s_lane[g], cudaStreamNonBlocking));
  for (int b = 0; b < 2 && good; ++b) {
    good = ok(cudaEventCreateWithFlags(&h->ev_h2d[b], cudaEventDisableTiming)) &&
           ok(cudaEventCreateWithFlags(&h->ev_comp[b], cudaEventDisableTiming)) &&
           ok(cudaEventCreateWithFlags(&h->ev_d2h[b], cudaEventDisableTiming)) &&
           ok(cudaEventCreateWithFlags(&h->ev_prep[b], cudaEventDisableTiming)) &&
           ok(cudaEventCreateWithFlags(&h->ev_block[b], cudaEventDisableTiming));
    for (int g = 0; g < pws_handle::kMaxLanes && good; ++g)
      good = ok(cudaEventCreateWithFlags(&h->ev_lane[b][g], cudaEventDisableTiming));
  }
  if (!good) {
    g_create_error = "pws_create: stream/event creation failed";
    delete h;
    return 1;
  }
  *out = h;
  return 0;
}

static void clear_timing(pws_handle* h);
static cudaError_t sync_all(pws_handle* h);

static void free_stage(pws_handle* h) {
  for (int b = 0; b < 2; ++b) {
    cudaFreeHost(h->hs_in[b]); h->hs_in[b] = nullptr;
    cudaFreeHost(h->hs_of[b]); h->hs_of[b] = nullptr;
    cudaFreeHost(h->hs_oi[b]); h->hs_oi[b] = nullptr;
    cudaFreeHost(h->hs_os[b]); h->hs_os[b] = nullptr;
    cudaFreeHost(h->hs_viol[b]); h->hs_viol[b] = nullptr;
  }
  h->stage_days = 0;
  h->stage_f64_elems = h->stage_i32_elems = h->stage_seg_elems = 0;
}

static void free_ring(pws_handle* h) {
  free_stage(h);
  for (int b = 0; b < 2; ++b) {
    cudaFree(h->d_forc[b]); h->d_forc[b] = nullptr;
    cudaFree(h->d_forc32[b]); h->d_forc32[b] = nullptr;
    cudaFree(h->d_of64[b]); h->d_of64[b] = nullptr;
    cudaFree(h->d_oi32[b]); h->d_oi32[b] = nullptr;
    cudaFree(h->d_oseg[b]); h->d_oseg[b] = nullptr;
  }
  h->ring_days = 0;
  h->forc32_days = 0;
  h->ring_f64_elems = h->ring_i32_elems = h->ring_seg_elems = 0;
}

static void free_scratch(pws_handle* h) {
  for (int b = 0; b < 2; ++b) {
    cudaFree(h->d_hru_lat[b]); h->d_hru_lat[b] = nullptr;
    cudaFree(h->d_cal[b]); h->d_cal[b] = nullptr;
    cudaFree(h->d_viol[b]); h->d_viol[b] = nullptr;
  }
  for (int b = 0; b < 2; ++b) {
    cudaFree(h->d_ots[b]); h->d_ots[b] = nullptr;
    cudaFree(h->d_day_outvol[b]); h->d_day_outvol[b] = nullptr;
    cudaFree(h->d_day_dstor[b]); h->d_day_dstor[b] = nullptr;
    cudaFree(h->d_day_lat[b]); h->d_day_lat[b] = nullptr;
    cudaFree(h->d_progress[b]); h->d_progress[b] = nullptr;
  }
  h->scratch_days = 0;
  h->scratch_copies = 0;
}

static void free_device(pws_handle* h) {
  cudaFree(h->d_f64); h->d_f64 = nullptr;
  cudaFree(h->d_i32); h->d_i32 = nullptr;
  cudaFree(h->d_state); h->d_state = nullptr;
  for (auto& p : h->d_soltab) { cudaFree(p); p = nullptr; }
  cudaFree(h->d_snarea); h->d_snarea = nullptr;
  cudaFree(h->d_err); h->d_err = nullptr;
  cudaFree(h->d_acc_hru); h->d_acc_hru = nullptr;
  cudaFree(h->d_acc_seg); h->d_acc_seg = nullptr;
  void* ch[] = {h->d_tsi, h->d_up_ptr, h->d_up_idx, h->d_hru_ptr, h->d_hru_idx, h->d_seg_flags,
                h->d_orig, h->d_chunk_begin, h->d_chunk_dmax, h->d_delay, h->d_ext_row, h->d_rts,
                h->d_ts, h->d_c0, h->d_c1, h->d_c2, h->d_outflow_ts, h->d_seg_inflow};
  for (void* p : ch) cudaFree(p);
  cudaFree(h->d_fg_kind); cudaFree(h->d_fg_obs_col); cudaFree(h->d_fg_flow_min);
  cudaFree(h->d_chunk_done); h->d_chunk_done = nullptr;
  h->d_fg_kind = h->d_fg_obs_col = nullptr;
  h->d_fg_flow_min = nullptr;
  h->d_tsi = h->d_up_ptr = h->d_up_idx = h->d_hru_ptr = h->d_hru_idx = h->d_seg_flags = h->d_orig =
      h->d_chunk_begin = h->d_chunk_dmax = h->d_delay = h->d_ext_row = nullptr;
  h->d_rts = nullptr;
  h->d_ts = h->d_c0 = h->d_c1 = h->d_c2 = h->d_outflow_ts = h->d_seg_inflow = nullptr;
  free_scratch(h);
  free_ring(h);
}

extern "C" int pws_destroy(pws_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  cudaFreeHost(h->hs_viol_run);
  free_device(h);
  clear_timing(h);
  cudaEventDestroy(h->ev_col0); cudaEventDestroy(h->ev_col1); cudaEventDestroy(h->ev_ch1);
  cudaEventDestroy(h->ev_base);
  for (int b = 0; b < 2; ++b) {
    cudaEventDestroy(h->ev_h2d[b]); cudaEventDestroy(h->ev_comp[b]); cudaEventDestroy(h->ev_d2h[b]);
    cudaEventDestroy(h->ev_prep[b]); cudaEventDestroy(h->ev_block[b]);
    for (auto e : h->ev_lane[b]) cudaEventDestroy(e);
  }
  cudaStreamDestroy(h->s_compute); cudaStreamDestroy(h->s_h2d); cudaStreamDestroy(h->s_d2h);
  cudaStreamDestroy(h->s_route);
  cudaStreamDestroy(h->s_route2);
  for (auto st : h->s_lane) cudaStreamDestroy(st);
  delete h;
  return 0;
}

// --------------------------------------------------------------- parameters
static int64_t expected_len(pws_handle* h, const char* name, bool is_f64) {
  const NameTables& t = names();
  if (is_f64) {
    if (find(t.hru_f64, name) >= 0) return h->nhru;
    if (find(t.mon_f64, name) >= 0) return 12 * h->nhru;
    if (find(t.seg_f64, name) >= 0) return h->nseg;
    if (!strcmp(name, "snarea_curve")) return h->ndeplval;
  } else {
    if (find(t.hru_i32, name) >= 0) return h->nhru;
    if (find(t.mon_i32, name) >= 0) return 12 * h->nhru;
    if (find(t.seg_i32, name) >= 0) return h->nseg;
  }
  return -1;
}

extern "C" int pws_set_param_f64(pws_handle* h, const char* name, const double* v, int64_t n) {
  const int64_t len = expected_len(h, name, true);
  if (len < 0) PWS_FAIL(h, "pws_set_param_f64: unknown parameter '%s'", name);
  if (n != len && n != 1)
    PWS_FAIL(h, "pws_set_param_f64: '%s' has %lld values, expected %lld (or 1)", name,
             (long long)n, (long long)len);
  std::vector<double>& dst = h->pf[name];
  dst.resize((size_t)len);
  for (int64_t k = 0; k < len; ++k) dst[(size_t)k] = n == 1 ? v[0] : v[k];
  h->inited = false;
  return 0;
}

extern "C" int pws_set_param_i32(pws_handle* h, const char* name, const int32_t* v, int64_t n) {
  const int64_t len = expected_len(h, name, false);
  if (len < 0) PWS_FAIL(h, "pws_set_param_i32: unknown parameter '%s'", name);
  if (n != len && n != 1)
    PWS_FAIL(h, "pws_set_param_i32: '%s' has %lld values, expected %lld (or 1)", name,
             (long long)n, (long long)len);
  std::vector<int32_t>& dst = h->pi[name];
  dst.resize((size_t)len);
  for (int64_t k = 0; k < len; ++k) dst[(size_t)k] = n == 1 ? v[0] : v[k];
  h->inited = false;
  return 0;
}

extern "C" int pws_set_option(pws_handle* h, const char* name, double value) {
  if (!strcmp(name, "dprst_flag")) h->dprst_flag = value != 0.0;
  else if (!strcmp(name, "temp_units")) h->temp_units = (int)value;
  else if (!strcmp(name, "start_month")) h->start_month = (int)value;
  else if (!strcmp(name, "start_doy")) h->start_doy = (int)value;
  else if (!strcmp(name, "adjust_parameters")) h->adjust_parameters = value != 0.0;
  else if (!strcmp(name, "albset_rna")) h->albset[0] = value;
  else if (!strcmp(name, "albset_rnm")) h->albset[1] = value;
  else if (!strcmp(name, "albset_sna")) h->albset[2] = value;
  else if (!strcmp(name, "albset_snm")) h->albset[3] = value;
  else if (!strcmp(name, "block_days")) { h->block_days = (int)value; return 0; }
  else if (!strcmp(name, "soltab_device")) h->soltab_device = value != 0.0;
  else if (!strcmp(name, "tma")) { h->tma = value != 0.0; return 0; }
  else if (!strcmp(name, "tiled_outputs")) { h->tiled_outputs = value != 0.0; return 0; }
  else if (!strcmp(name, "lanes")) {
    if (value < 0 || value > pws_handle::kMaxLanes)
      PWS_FAIL(h, "pws_set_option: lanes must be in 0..%d", pws_handle::kMaxLanes);
    PWS_CUDA(h, sync_all(h));  // a lane's stream orders the blocks of ITS tiles
    h->lanes_opt = (int)value;
    return 0;
  }
  else if (!strcmp(name, "nhm_sink")) { h->nhm_sink = value != 0.0; return 0; }
  else if (!strcmp(name, "wide_column")) { h->wide_column = value != 0.0; return 0; }
  else if (!strcmp(name, "one_cta_carveout")) { h->one_cta_carveout = value != 0.0; return 0; }
  else if (!strcmp(name, "overlap_routing")) { h->overlap_routing = value < 0 ? -1 : (value != 0.0); return 0; }
  else if (!strcmp(name, "accumulate")) { h->accumulate_on = value != 0.0; return 0; }
  else if (!strcmp(name, "forcing_period")) {
    if (value < 0 || (value > 0 && h->nhru % (int64_t)value != 0) || h->nhru >= ((int64_t)1 << 31))
      PWS_FAIL(h, "pws_set_option: forcing_period must divide nhru (%lld < 2^31)", (long long)h->nhru);
    PWS_CUDA(h, sync_all(h));
    h->forcing_period = (int64_t)value;
    free_ring(h);  // the forcing ring is sized by the number of forcing columns
    return 0;
  }
  else if (!strcmp(name, "column_hrus")) {
    if (value != 0 && value != 128 && value != 192 && value != PWS_COLUMN_HRUS)
      PWS_FAIL(h, "pws_set_option: column_hrus must be 0, 128, 192 or %d", PWS_COLUMN_HRUS);
    h->column_hrus_opt = (int)value;
    return 0;
  }
  else if (!strcmp(name, "route_overlap")) { h->route_overlap_opt = value < 0 ? -1 : (value != 0.0); return 0; }
  else if (!strcmp(name, "route_hours")) {
    if (value != 0 && value != PWS_ROUTE_HOURS && value != PWS_ROUTE_HOURS_SHALLOW &&
        value != PWS_ROUTE_HOURS_DEEP)
      PWS_FAIL(h, "pws_set_option: route_hours must be 0 (automatic), %d, %d or %d", PWS_ROUTE_HOURS_DEEP,
               PWS_ROUTE_HOURS, PWS_ROUTE_HOURS_SHALLOW);
    h->route_hours_opt = (int)value;
  }
  else if (!strcmp(name, "channel_chunk")) {
    if (value < 1 || value > PWS_CHUNK_THREADS)
      PWS_FAIL(h, "pws_set_option: channel_chunk must be in 1..%d", PWS_CHUNK_THREADS);
    h->channel_chunk = (int)value;
  }
  else PWS_FAIL(h, "pws_set_option: unknown option '%s'", name);
  h->inited = false;
  return 0;
}

// Hours per systolic step of k_route_chunks: long steps for shallow segment
// graphs (fewer barriers), short ones for deep graphs (the pipeline fill is one
// step per level; see PWS_ROUTE_HOURS in pws_kernels.cuh).
static int route_hours(const pws_handle* h) {
  if (h->route_hours_opt) return h->route_hours_opt;
  if (h->nlevels <= 160) return PWS_ROUTE_HOURS_SHALLOW;
  return h->nlevels > PWS_ROUTE_DEEP_LEVELS ? PWS_ROUTE_HOURS_DEEP : PWS_ROUTE_HOURS;
}

// May the routing launch of block b+1 start while that of block b still runs?
// A deep network pays its systolic fill (levels x one step, ~25 ms for 8,856
// levels) per launch; with two routing streams a chunk of launch b+1 starts as
// soon as the SAME chunk of launch b has stored its state (d_chunk_done), so the
// fill is paid once per run.  Safe only while the spinning CTAs of the younger
// launch cannot keep the older launch's CTAs off the SMs: every chunk of a launch
// must fit the GPU with room to spare (a CTA of k_route_chunks takes a whole SM).
static bool route_overlap(const pws_handle* h) {
  if (h->route_overlap_opt == 0 || h->nseg == 0) return false;
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
  if (h->nchunks > nsm - 8) return false;
  return h->route_overlap_opt == 1 || h->nlevels > 160;
}

// --------------------------------------------------------------- channel init
// prms_channel.py:189-342 on the host (tiny, once per run): Muskingum
// coefficients, topological levels (replaces networkx), CSR of upstream
// segments and of contributing HRUs, everything permuted to level order.
static int channel_build(pws_handle* h) {
  const int64_t nseg = h->nseg;
  if (nseg == 0) return 0;
  const NameTables& t = names();
  for (auto& n : t.seg_f64)
    if (!h->pf.count(n)) PWS_FAIL(h, "pws_init: missing parameter '%s'", n.c_str());
  for (auto& n : t.seg_i32)
    if (!h->pi.count(n)) PWS_FAIL(h, "pws_init: missing parameter '%s'", n.c_str());
  const std::vector<int32_t>& toseg = h->pi["tosegment"];
  const std::vector<int32_t>& segtype = h->pi["segment_type"];
  const std::vector<int32_t>& hru_segment = h->pi["hru_segment"];
  const std::vector<double>&mann_n = h->pf["mann_n"], &seg_depth = h->pf["seg_depth"],
        &seg_length = h->pf["seg_length"], &seg_slope = h->pf["seg_slope"],
        &x_coef = h->pf["x_coef"], &flow_init = h->pf["segment_flow_init"];
  // levels by Kahn: level = 1 + max(level of upstream)
  std::vector<int> indeg(nseg, 0), level(nseg, 0), order;
  order.reserve(nseg);
  for (int64_t i = 0; i < nseg; ++i) {
    const int to = toseg[i] - 1;
    if (to >= nseg) PWS_FAIL(h, "pws_init: tosegment[%lld] out of range", (long long)i);
    if (to >= 0) indeg[to]++;
  }
  for (int64_t i = 0; i < nseg; ++i)
    if (indeg[i] == 0) order.push_back((int)i);
  for (size_t head = 0; head < order.size(); ++head) {
    const int i = order[head];
    const int to = toseg[i] - 1;
    if (to >= 0) {
      level[to] = std::max(level[to], level[i] + 1);
      if (--indeg[to] == 0) order.push_back(to);
    }
  }
  if ((int64_t)order.size() != nseg) PWS_FAIL(h, "pws_init: the segment graph has a cycle");
  int nlev = 0;
  for (int64_t i = 0; i < nseg; ++i) nlev = std::max(nlev, level[i] + 1);
  std::vector<int> rank(nseg);
  for (int64_t k = 0; k < nseg; ++k) rank[order[k]] = (int)k;
  h->nlevels = nlev;
  // ---- chunks.  (1) Cut the forest into connected pieces of <= cap segments:
  // walking upstream-first, a segment keeps its upstream subtrees until they
  // no longer fit, then sheds the largest ones (each becomes a piece of its
  // own).  Outlet trees that fit stay whole.
  const int cap = h->channel_chunk;
  std::vector<std::vector<int>> kids(nseg);
  for (int k = 0; k < nseg; ++k) {
    const int i = order[k];
    if (toseg[i] > 0) kids[toseg[i] - 1].push_back(i);
  }
  std::vector<int> resid(nseg, 1);
  std::vector<char> is_root(nseg, 0);
  for (int k = 0; k < nseg; ++k) {
    const int i = order[k];
    int tot = 1;
    for (int c : kids[i]) tot += resid[c];
    if (tot > cap) {
      std::vector<int> by = kids[i];
      std::stable_sort(by.begin(), by.end(), [&](int x, int y) { return resid[x] > resid[y]; });
      for (int c : by) {
        if (tot <= cap) break;
        is_root[c] = 1;
        tot -= resid[c];
      }
    }
    resid[i] = tot;
    if (toseg[i] <= 0) is_root[i] = 1;
  }
  // (2) Number the pieces upstream-first (depth-first over the piece forest,
  // outlets in original order) and pack them next-fit, so that a chunk only
  // ever waits for chunks with a lower number.
  std::vector<int> piece_of(nseg, -1), piece_size, piece_root;
  {
    // piece forest post-order: iterative DFS from each outlet
    std::vector<int> root_order;  // piece roots, upstream pieces first
    std::vector<std::pair<int, int>> stack;  // (root segment, state)
    std::vector<std::vector<int>> piece_kids(nseg);  // piece root -> roots of upstream pieces
    // a cut child c hangs below the piece that contains its downstream segment;
    // find that piece's root by walking down until a root is met
    std::vector<int> root_of(nseg, -1);
    for (int k = (int)nseg - 1; k >= 0; --k) {  // downstream-first
      const int i = order[k];
      root_of[i] = is_root[i] ? i : root_of[toseg[i] - 1];
    }
    for (int k = 0; k < nseg; ++k) {
      const int i = order[k];
      if (is_root[i] && toseg[i] > 0) piece_kids[root_of[toseg[i] - 1]].push_back(i);
    }
    for (int64_t o = 0; o < nseg; ++o) {
      if (toseg[o] > 0) continue;
      stack.push_back({(int)o, 0});
      while (!stack.empty()) {
        auto& top = stack.back();
        const int r = top.first;
        if (top.second < (int)piece_kids[r].size()) {
          const int nxt = piece_kids[r][top.second++];
          stack.push_back({nxt, 0});
        } else {
          root_order.push_back(r);
          stack.pop_back();
        }
      }
    }
    std::vector<int> pid_of_root(nseg, -1);
    for (size_t k = 0; k < root_order.size(); ++k) {
      pid_of_root[root_order[k]] = (int)k;
      piece_root.push_back(root_order[k]);
      piece_size.push_back(resid[root_order[k]]);
    }
    for (int64_t i = 0; i < nseg; ++i) piece_of[i] = pid_of_root[root_of[i]];
  }
  std::vector<int> chunk_of_piece(piece_size.size(), 0);
  int nchunks = 0;
  {
    int fill = 0;
    for (size_t p = 0; p < piece_size.size(); ++p) {
      if (p == 0 || fill + piece_size[p] > cap) {
        nchunks++;
        fill = 0;
      }
      chunk_of_piece[p] = nchunks - 1;
      fill += piece_size[p];
    }
  }
  h->nchunks = nchunks;
  std::vector<int> chunk_of(nseg);
  for (int64_t i = 0; i < nseg; ++i) chunk_of[i] = chunk_of_piece[piece_of[i]];
  // systolic delay: hops to the root of the chunk-local tree, counted from the
  // deepest segment of the chunk, so a local upstream segment is one step ahead
  std::vector<int> hops(nseg, 0), chunk_dmax_v(nchunks, 0), delay_v(nseg, 0);
  for (int k = (int)nseg - 1; k >= 0; --k) {  // downstream-first
    const int i = order[k];
    const int to = toseg[i] - 1;
    hops[i] = (to >= 0 && chunk_of[to] == chunk_of[i]) ? hops[to] + 1 : 0;
    chunk_dmax_v[chunk_of[i]] = std::max(chunk_dmax_v[chunk_of[i]], hops[i]);
  }
  h->max_dmax = 0;
  for (int64_t i = 0; i < nseg; ++i) {
    delay_v[i] = chunk_dmax_v[chunk_of[i]] - hops[i];
    h->max_dmax = std::max(h->max_dmax, chunk_dmax_v[chunk_of[i]]);
  }
  // pos order: by (chunk, delay mod steps-per-day, delay, original index), so the lanes
  // of a warp cross their day boundaries in the same step; rank in `order`
  // kept for the accumulation order of upstream inflows
  std::vector<int> bypos(nseg);
  for (int64_t i = 0; i < nseg; ++i) bypos[i] = (int)i;
  std::stable_sort(bypos.begin(), bypos.end(), [&](int x, int y) {
    if (chunk_of[x] != chunk_of[y]) return chunk_of[x] < chunk_of[y];
    const int spd = 24 / route_hours(h);  // steps per day
    if (delay_v[x] % spd != delay_v[y] % spd) return delay_v[x] % spd < delay_v[y] % spd;
    return delay_v[x] < delay_v[y];
  });
  std::vector<int> pos_of(nseg);
  for (int64_t p = 0; p < nseg; ++p) pos_of[bypos[p]] = (int)p;
  std::vector<int32_t> chunk_begin(nchunks + 1, 0);
  for (int64_t i = 0; i < nseg; ++i) chunk_begin[chunk_of[i] + 1]++;
  for (int cidx = 0; cidx < nchunks; ++cidx) chunk_begin[cidx + 1] += chunk_begin[cidx];
  // coefficients
  std::vector<int32_t> tsi(nseg), seg_flags(nseg, 0), orig(nseg), delay(nseg), ext_row(nseg, -1);
  std::vector<double> rts(nseg);
  h->n_ext_rows = 0;
  std::vector<double> ts(nseg), c0(nseg), c1(nseg), c2(nseg), seg_inflow(nseg, 0.0),
      outflow_ts(nseg, 0.0);
  for (int64_t p = 0; p < nseg; ++p) {
    const int i = bypos[p];
    // velocity from the unadjusted slope (prms_channel.py:234-242)
    const double velocity =
        ((1.0 / mann_n[i]) * std::sqrt(seg_slope[i]) * std::pow(seg_depth[i], 2.0 / 3.0)) * 60.0 * 60.0;
    double K = 24.0;
    if (velocity > 0.0) K = seg_length[i] / velocity;
    if (segtype[i] == 2) K = 24.0;
    if (K < 0.01) K = 0.01;
    if (K > 24.0) K = 24.0;
    double tsv = 1.0;
    int tsiv = 1;
    if (K < 1.0) tsiv = -1;
    else if (K < 2.0) { tsv = 1.0; tsiv = 1; }
    else if (K < 3.0) { tsv = 2.0; tsiv = 2; }
    else if (K < 4.0) { tsv = 3.0; tsiv = 3; }
    else if (K < 6.0) { tsv = 4.0; tsiv = 4; }
    else if (K < 8.0) { tsv = 6.0; tsiv = 6; }
    else if (K < 12.0) { tsv = 8.0; tsiv = 8; }
    else if (K < 24.0) { tsv = 12.0; tsiv = 12; }
    else { tsv = 24.0; tsiv = 24; }
    // FlowGraph: a PassThroughFlowNode puts out what comes in, every hour -- the
    // recurrence of a K < 1 h segment; an ObsIn node's routed value is never used
    if (h->flow_graph && h->fg_kind[i] != NODE_MUSKINGUM) { tsv = 1.0; tsiv = -1; }
    const double x = x_coef[i];
    double d = K - (K * x) + (0.5 * tsv);
    if (std::fabs(d) < 1e-6) d = 0.0001;
    double a0 = (-(K * x) + (0.5 * tsv)) / d;
    double a1 = ((K * x) + (0.5 * tsv)) / d;
    double a2 = (K - (K * x) - (0.5 * tsv)) / d;
    if (a2 < 0.0) { a1 += a2; a2 = 0.0; }
    if (a0 < 0.0) { a1 += a0; a0 = 0.0; }
    tsi[p] = tsiv; ts[p] = tsv; c0[p] = a0; c1[p] = a1; c2[p] = a2;
    rts[p] = 1.0 / tsv;
    delay[p] = delay_v[i];
    if (toseg[i] > 0) {
      seg_flags[p] |= SEG_HAS_DOWN;
      if (chunk_of[toseg[i] - 1] == chunk_of[i]) seg_flags[p] |= SEG_DOWN_LOCAL;
      else ext_row[p] = h->n_ext_rows++;
    }
    orig[p] = i;
  }
  // initial _seg_inflow: assignment, not sum, in index order (prms_channel.py:336-340)
  for (int64_t i = 0; i < nseg; ++i) {
    const int to = toseg[i] - 1;
    // (a PRMSChannelFlowNode starts from zero: prms_channel_flow_graph.py:63-69)
    if (to >= 0 && !h->flow_graph) seg_inflow[pos_of[to]] = flow_init[i];
  }
  // upstream CSR, each list ordered by topological rank
  std::vector<std::vector<int>> ups(nseg);
  for (int k = 0; k < nseg; ++k) {
    const int i = order[k];
    const int to = toseg[i] - 1;
    if (to >= 0) ups[pos_of[to]].push_back(pos_of[i]);
  }
  std::vector<int32_t> up_ptr(nseg + 1, 0), up_idx;
  h->n_cross_edges = 0;
  for (int64_t p = 0; p < nseg; ++p) {
    up_ptr[p + 1] = up_ptr[p] + (int32_t)ups[p].size();
    bool all_local = true;
    for (int u : ups[p]) {
      up_idx.push_back(u);
      if (chunk_of[bypos[u]] != chunk_of[bypos[p]]) {
        all_local = false;
        h->n_cross_edges++;
      }
    }
    if (all_local) seg_flags[p] |= SEG_UPS_LOCAL;
  }
  // HRU CSR, ascending HRU index
  std::vector<std::vector<int>> hrus(nseg);
  for (int64_t hh = 0; hh < h->nhru; ++hh) {
    const int sg = hru_segment[hh] - 1;
    if (sg >= nseg) PWS_FAIL(h, "pws_init: hru_segment[%lld] out of range", (long long)hh);
    if (sg >= 0) hrus[pos_of[sg]].push_back((int)hh);
  }
  std::vector<int32_t> hru_ptr(nseg + 1, 0), hru_idx;
  for (int64_t p = 0; p < nseg; ++p) {
    hru_ptr[p + 1] = hru_ptr[p] + (int32_t)hrus[p].size();
    for (int u : hrus[p]) hru_idx.push_back(u);
  }
  auto up_i32 = [&](int32_t** dst, const std::vector<int32_t>& v) -> cudaError_t {
    cudaError_t e = cudaMalloc(dst, std::max<size_t>(v.size(), 1) * sizeof(int32_t));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, v.data(), v.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
  };
  auto up_f64 = [&](double** dst, const std::vector<double>& v) -> cudaError_t {
    cudaError_t e = cudaMalloc(dst, std::max<size_t>(v.size(), 1) * sizeof(double));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice);
  };
  PWS_CUDA(h, up_i32(&h->d_tsi, tsi));
  PWS_CUDA(h, up_i32(&h->d_up_ptr, up_ptr));
  PWS_CUDA(h, up_i32(&h->d_up_idx, up_idx));
  PWS_CUDA(h, up_i32(&h->d_hru_ptr, hru_ptr));
  PWS_CUDA(h, up_i32(&h->d_hru_idx, hru_idx));
  PWS_CUDA(h, up_i32(&h->d_seg_flags, seg_flags));
  PWS_CUDA(h, up_i32(&h->d_chunk_begin, chunk_begin));
  {
    // serial of the last launch that stored a chunk's state: none yet
    std::vector<int32_t> zero((size_t)std::max(h->nchunks, 1), 0);
    int32_t* p = nullptr;
    PWS_CUDA(h, up_i32(&p, zero));
    cudaFree(h->d_chunk_done);
    h->d_chunk_done = p;
    h->route_serial = 0;
  }
  {
    std::vector<int32_t> dm(chunk_dmax_v.begin(), chunk_dmax_v.end());
    PWS_CUDA(h, up_i32(&h->d_chunk_dmax, dm));
  }
  PWS_CUDA(h, up_i32(&h->d_delay, delay));
  PWS_CUDA(h, up_i32(&h->d_ext_row, ext_row));
  PWS_CUDA(h, up_f64(&h->d_rts, rts));
  PWS_CUDA(h, up_i32(&h->d_orig, orig));
  PWS_CUDA(h, up_f64(&h->d_ts, ts));
  PWS_CUDA(h, up_f64(&h->d_c0, c0));
  PWS_CUDA(h, up_f64(&h->d_c1, c1));
  PWS_CUDA(h, up_f64(&h->d_c2, c2));
  PWS_CUDA(h, up_f64(&h->d_outflow_ts, outflow_ts));
  PWS_CUDA(h, up_f64(&h->d_seg_inflow, seg_inflow));
  if (h->flow_graph) {
    std::vector<int32_t> kind(nseg), col(nseg);
    std::vector<double> fmin(nseg, 0.0);
    for (int64_t p = 0; p < nseg; ++p) {
      kind[p] = h->fg_kind[bypos[p]];
      col[p] = h->fg_obs_col[bypos[p]];
      if (!h->fg_flow_min.empty()) fmin[p] = h->fg_flow_min[bypos[p]];
    }
    PWS_CUDA(h, up_i32(&h->d_fg_kind, kind));
    PWS_CUDA(h, up_i32(&h->d_fg_obs_col, col));
    PWS_CUDA(h, up_f64(&h->d_fg_flow_min, fmin));
  }
  h->h_seg_inflow_init = seg_inflow;
  h->h_orig = orig;
  return 0;
}

// --------------------------------------------------------------- init
extern "C" int pws_init(pws_handle* h) {
  PWS_CUDA(h, cudaSetDevice(h->device));
  const NameTables& t = names();
  for (auto& n : t.hru_f64)
    if (!h->pf.count(n)) PWS_FAIL(h, "pws_init: missing parameter '%s'", n.c_str());
  for (auto& n : t.mon_f64)
    if (!h->pf.count(n)) PWS_FAIL(h, "pws_init: missing parameter '%s'", n.c_str());
  if (!h->pf.count("snarea_curve")) PWS_FAIL(h, "pws_init: missing parameter 'snarea_curve'");
  for (auto& n : t.hru_i32)
    if (!h->pi.count(n)) {
      if (n == "hru_segment" && h->nseg == 0) {
        h->pi[n].assign((size_t)h->nhru, 0);
        continue;
      }
      PWS_FAIL(h, "pws_init: missing parameter '%s'", n.c_str());
    }
  for (auto& n : t.mon_i32)
    if (!h->pi.count(n)) PWS_FAIL(h, "pws_init: missing parameter '%s'", n.c_str());
  {
    const std::vector<int32_t>& dc = h->pi["hru_deplcrv"];
    for (int64_t i = 0; i < h->nhru; ++i)
      if (dc[i] < 1 || dc[i] * 11 > h->ndeplval)
        PWS_FAIL(h, "pws_init: hru_deplcrv[%lld]=%d outside snarea_curve", (long long)i, dc[i]);
  }
  cudaDeviceSynchronize();
  free_device(h);
  const int64_t S = h->stride, n = h->nhru;
  const size_t rows_f64 = t.hru_f64.size() + 12 * t.mon_f64.size() + t.derived.size() + 12;
  const size_t rows_i32 = t.hru_i32.size() + 12 * t.mon_i32.size();
  // + one CTA's worth of slack: lanes past the last HRU of the last CTA may read behind a row
  const size_t slack = 2 * PWS_COLUMN_HRUS;  // >= the largest CTA
  PWS_CUDA(h, cudaMalloc(&h->d_f64, (rows_f64 * S + slack) * sizeof(double)));
  PWS_CUDA(h, cudaMemset(h->d_f64, 0, (rows_f64 * S + slack) * sizeof(double)));
  PWS_CUDA(h, cudaMalloc(&h->d_i32, (rows_i32 * S + slack) * sizeof(int32_t)));
  PWS_CUDA(h, cudaMemset(h->d_i32, 0, (rows_i32 * S + slack) * sizeof(int32_t)));
  PWS_CUDA(h, cudaMalloc(&h->d_state, (size_t)PWS_N_HRU_STATE * S * sizeof(double)));
  PWS_CUDA(h, cudaMemset(h->d_state, 0, (size_t)PWS_N_HRU_STATE * S * sizeof(double)));
  for (auto& p : h->d_soltab) PWS_CUDA(h, cudaMalloc(&p, (size_t)kNdoy * S * sizeof(double)));
  PWS_CUDA(h, cudaMalloc(&h->d_snarea, (size_t)h->ndeplval * sizeof(double)));
  PWS_CUDA(h, cudaMemcpy(h->d_snarea, h->pf["snarea_curve"].data(),
                         (size_t)h->ndeplval * sizeof(double), cudaMemcpyHostToDevice));
  PWS_CUDA(h, cudaMalloc(&h->d_err, sizeof(int)));
  PWS_CUDA(h, cudaMemset(h->d_err, 0, sizeof(int)));
  // lay the rows out and upload
  size_t row = 0;
  HruParams& P = h->P;
  P = HruParams{};
  P.nhru = n;
  P.stride = S;
  auto put_f64 = [&](const std::vector<double>& v, int nrows) -> double* {
    double* base = h->d_f64 + row * S;
    for (int r = 0; r < nrows; ++r)
      cudaMemcpy(base + (size_t)r * S, v.data() + (size_t)r * n, (size_t)n * sizeof(double),
                 cudaMemcpyHostToDevice);
    row += nrows;
    return base;
  };
#define X(nm) P.nm = put_f64(h->pf[#nm], 1);
  PWS_HRU_F64_PARAMS(X)
#undef X
#define X(nm) P.nm = put_f64(h->pf[#nm], 12);
  PWS_MON_F64_PARAMS(X)
#undef X
#define X(nm) h->D.nm = h->d_f64 + row * S; P.nm = h->D.nm; row += 1;
  PWS_HRU_F64_DERIVED(X)
#undef X
  h->D.tmax_allsnow_c = h->d_f64 + row * S;
  P.tmax_allsnow_c = h->D.tmax_allsnow_c;
  row += 12;
  size_t irow = 0;
  auto put_i32 = [&](const std::vector<int32_t>& v, int nrows) -> int32_t* {
    int32_t* base = h->d_i32 + irow * S;
    for (int r = 0; r < nrows; ++r)
      cudaMemcpy(base + (size_t)r * S, v.data() + (size_t)r * n, (size_t)n * sizeof(int32_t),
                 cudaMemcpyHostToDevice);
    irow += nrows;
    return base;
  };
#define X(nm) P.nm = put_i32(h->pi[#nm], 1);
  PWS_HRU_I32_PARAMS(X)
#undef X
#define X(nm) P.nm = put_i32(h->pi[#nm], 12);
  PWS_MON_I32_PARAMS(X)
#undef X
  P.snarea_curve = h->d_snarea;
  P.soltab_potsw = h->d_soltab[0];
  P.soltab_horad_potsw = h->d_soltab[1];
  P.albset_rna = h->albset[0];
  P.albset_rnm = h->albset[1];
  P.albset_sna = h->albset[2];
  P.albset_snm = h->albset[3];
  P.dprst_flag = h->dprst_flag;
  P.temp_units = h->temp_units;
  PWS_CUDA(h, cudaGetLastError());
  // PRMSSolarGeometry tables + hru_cossl.  Default: k_soltab, with correctly
  // rounded elementary functions (pws_crmath.cuh; the 366 declination constants
  // come from the same functions here).  Option "soltab_device" 0: host threads
  // and the machine's libm -- the bits an oracle or reference on THIS machine
  // produces, a debugging aid for last-bit comparisons.
  {
    double sin_d[kNdoy], cos_d[kNdoy], tan_d[kNdoy], r1[kNdoy];
    for (int j = 0; j < kNdoy; ++j) {
      double decl;
      if (h->soltab_device) {
        sol_constants<CrMath>(j, decl, r1[j]);
        sin_d[j] = CrMath::sin(decl);
        cos_d[j] = CrMath::cos(decl);
        tan_d[j] = CrMath::tan(decl);
      } else {
        sol_constants(j, decl, r1[j]);
        sin_d[j] = std::sin(decl);
        cos_d[j] = std::cos(decl);
        tan_d[j] = std::tan(decl);
      }
    }
    if (h->soltab_device) {
      PWS_CUDA(h, cudaMemcpyToSymbol(c_sol_sin_d, sin_d, sizeof sin_d));
      PWS_CUDA(h, cudaMemcpyToSymbol(c_sol_cos_d, cos_d, sizeof cos_d));
      PWS_CUDA(h, cudaMemcpyToSymbol(c_sol_tan_d, tan_d, sizeof tan_d));
      PWS_CUDA(h, cudaMemcpyToSymbol(c_sol_r1, r1, sizeof r1));
      dim3 grid((unsigned)((n + 127) / 128), kSoltabSlices);
      k_soltab<<<grid, 128, 0, h->s_compute>>>(n, S, P.hru_slope, P.hru_aspect, P.hru_lat,
                                               h->d_soltab[0], h->d_soltab[1], h->d_soltab[2],
                                               h->D.hru_cossl);
      h->launches++;
    } else {
      const std::vector<double>&slope = h->pf["hru_slope"], &aspect = h->pf["hru_aspect"],
            &lat = h->pf["hru_lat"];
      std::vector<double> tab[3], cossl((size_t)n);
      for (auto& tb : tab) tb.assign((size_t)kNdoy * S, 0.0);
      const int nthr = (int)std::max<int64_t>(
          1, std::min<int64_t>(std::thread::hardware_concurrency(), (n + 255) / 256));
      auto work = [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
          cossl[(size_t)i] = std::cos(std::atan(slope[(size_t)i]));
          const SolHru flat = sol_hru(0.0, 0.0, lat[(size_t)i]);
          const SolHru hs = sol_hru(slope[(size_t)i], aspect[(size_t)i], lat[(size_t)i]);
          for (int j = 0; j < kNdoy; ++j) {
            double solt, sunh;
            sol_day(flat, sin_d[j], cos_d[j], tan_d[j], r1[j], solt, sunh);
            tab[1][(size_t)j * S + i] = solt;
            sol_day(hs, sin_d[j], cos_d[j], tan_d[j], r1[j], solt, sunh);
            tab[0][(size_t)j * S + i] = solt;
            tab[2][(size_t)j * S + i] = sunh;
          }
        }
      };
      std::vector<std::thread> pool;
      const int64_t chunk = (n + nthr - 1) / nthr;
      for (int t = 0; t < nthr; ++t) {
        const int64_t lo = t * chunk, hi = std::min<int64_t>(n, lo + chunk);
        if (lo < hi) pool.emplace_back(work, lo, hi);
      }
      for (auto& th : pool) th.join();
      for (int k = 0; k < 3; ++k)
        PWS_CUDA(h, cudaMemcpy(h->d_soltab[k], tab[k].data(), (size_t)kNdoy * S * sizeof(double),
                               cudaMemcpyHostToDevice));
      PWS_CUDA(h, cudaMemcpy(h->D.hru_cossl, cossl.data(), (size_t)n * sizeof(double),
                             cudaMemcpyHostToDevice));
    }
  }
  k_hru_init<<<(unsigned)((n + 127) / 128), 128, 0, h->s_compute>>>(
      P, h->D, h->d_state, h->start_month, h->start_doy, h->adjust_parameters, h->d_err);
  h->launches++;
  PWS_CUDA(h, cudaGetLastError());
  PWS_CUDA(h, cudaStreamSynchronize(h->s_compute));
  int errflag = 0;
  PWS_CUDA(h, cudaMemcpy(&errflag, h->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (errflag & 2)
    PWS_FAIL(h, "pws_init: snowpack_init > 0 is not runnable in the reference (prms_snow.py:413)");
  if (errflag & 4) PWS_FAIL(h, "pws_init: dprst_frac > 0 and dprst_depth_avg == 0");
  if (errflag & 8) PWS_FAIL(h, "pws_init: values of pref_flow_infil_frac outside of [0,1]");
  if (channel_build(h)) return 1;
  h->inited = true;
  return 0;
}

extern "C" int pws_get_soltab(pws_handle* h, int which, double* out) {
  if (!h->inited) PWS_FAIL(h, "pws_get_soltab before pws_init");
  if (which < 0 || which > 2) PWS_FAIL(h, "pws_get_soltab: which must be 0..2");
  PWS_CUDA(h, cudaSetDevice(h->device));
  PWS_CUDA(h, cudaMemcpy2D(out, (size_t)h->nhru * sizeof(double), h->d_soltab[which],
                           (size_t)h->stride * sizeof(double), (size_t)h->nhru * sizeof(double),
                           kNdoy, cudaMemcpyDeviceToHost));
  return 0;
}

// --------------------------------------------------------------- outputs
// slots of one day of results: drained variables first, in selection order
// (float64 and int32 counted separately), then the accumulate-only ones
static void rebuild_slots(pws_handle* h) {
  for (auto& v : h->slot_hru) v = -1;
  for (auto& v : h->slot_seg) v = -1;
  int nf = 0, ni = 0, ns = 0;
  for (int v : h->sel_hru) h->slot_hru[v] = (short)(names().hru_var_kind[v] == 0 ? nf++ : ni++);
  for (int v : h->sel_seg) h->slot_seg[v] = (short)ns++;
  h->n_f64_drain = nf;
  h->n_seg_drain = ns;
  h->acc_hru_slot.clear();
  h->acc_seg_slot.clear();
  for (int v : h->acc_hru) {
    if (h->slot_hru[v] < 0) h->slot_hru[v] = (short)nf++;
    h->acc_hru_slot.push_back(h->slot_hru[v]);
  }
  for (int v : h->acc_seg) {
    if (h->slot_seg[v] < 0) h->slot_seg[v] = (short)ns++;
    h->acc_seg_slot.push_back(h->slot_seg[v]);
  }
  h->n_f64 = nf;
  h->n_i32 = ni;
  h->n_seg_sel = ns;
}

extern "C" int pws_select_outputs(pws_handle* h, int n_hru, const int32_t* hru_idx, int n_seg,
                                  const int32_t* seg_idx) {
  std::vector<char> seen_h(PWS_N_HRU_VARS, 0), seen_s(PWS_N_SEG_VARS, 0);
  for (int k = 0; k < n_hru; ++k) {
    const int v = hru_idx[k];
    if (v < 0 || v >= PWS_N_HRU_VARS || seen_h[v])
      PWS_FAIL(h, "pws_select_outputs: bad or repeated HRU variable index %d", v);
    seen_h[v] = 1;
  }
  for (int k = 0; k < n_seg; ++k) {
    const int v = seg_idx[k];
    if (v < 0 || v >= PWS_N_SEG_VARS || seen_s[v])
      PWS_FAIL(h, "pws_select_outputs: bad or repeated segment variable index %d", v);
    seen_s[v] = 1;
  }
  h->sel_hru.assign(hru_idx, hru_idx + n_hru);
  h->sel_seg.assign(seg_idx, seg_idx + n_seg);
  rebuild_slots(h);
  // the rings of pws_run_host are kept: they are re-allocated only when a run
  // needs more room than they have
  return 0;
}

// Budget accumulations on the device (base/budget.py:262-269: every term of a
// process's mass budget is summed over the steps, `acc += value * dt`): the
// chosen float64 variables are written by the kernels like selected outputs
// but need not be drained -- k_accumulate adds each block's days to a running
// per-unit sum, in day order, exactly the reference's sequence of additions.
extern "C" int pws_select_accumulations(pws_handle* h, int n_hru, const int32_t* hru_idx, int n_seg,
                                        const int32_t* seg_idx) {
  std::vector<char> seen_h(PWS_N_HRU_VARS, 0), seen_s(PWS_N_SEG_VARS, 0);
  for (int k = 0; k < n_hru; ++k) {
    const int v = hru_idx[k];
    if (v < 0 || v >= PWS_N_HRU_VARS || seen_h[v] || names().hru_var_kind[v] != 0)
      PWS_FAIL(h, "pws_select_accumulations: bad, repeated or non-float64 HRU variable index %d", v);
    seen_h[v] = 1;
  }
  for (int k = 0; k < n_seg; ++k) {
    const int v = seg_idx[k];
    if (v < 0 || v >= PWS_N_SEG_VARS || seen_s[v])
      PWS_FAIL(h, "pws_select_accumulations: bad or repeated segment variable index %d", v);
    seen_s[v] = 1;
  }
  PWS_CUDA(h, cudaSetDevice(h->device));
  PWS_CUDA(h, sync_all(h));
  h->acc_hru.assign(hru_idx, hru_idx + n_hru);
  h->acc_seg.assign(seg_idx, seg_idx + n_seg);
  cudaFree(h->d_acc_hru); h->d_acc_hru = nullptr;
  cudaFree(h->d_acc_seg); h->d_acc_seg = nullptr;
  if (n_hru > 0) {
    PWS_CUDA(h, cudaMalloc(&h->d_acc_hru, (size_t)n_hru * h->stride * sizeof(double)));
    PWS_CUDA(h, cudaMemset(h->d_acc_hru, 0, (size_t)n_hru * h->stride * sizeof(double)));
  }
  if (n_seg > 0 && h->nseg > 0) {
    PWS_CUDA(h, cudaMalloc(&h->d_acc_seg, (size_t)n_seg * h->sstride * sizeof(double)));
    PWS_CUDA(h, cudaMemset(h->d_acc_seg, 0, (size_t)n_seg * h->sstride * sizeof(double)));
  }
  rebuild_slots(h);
  return 0;
}

extern "C" int pws_get_accumulations(pws_handle* h, double* h_hru, double* h_seg) {
  PWS_CUDA(h, cudaSetDevice(h->device));
  PWS_CUDA(h, sync_all(h));
  if (h_hru && h->d_acc_hru)
    PWS_CUDA(h, cudaMemcpy2D(h_hru, (size_t)h->nhru * sizeof(double), h->d_acc_hru,
                             (size_t)h->stride * sizeof(double), (size_t)h->nhru * sizeof(double),
                             h->acc_hru.size(), cudaMemcpyDeviceToHost));
  if (h_seg && h->d_acc_seg)
    PWS_CUDA(h, cudaMemcpy2D(h_seg, (size_t)h->nseg * sizeof(double), h->d_acc_seg,
                             (size_t)h->sstride * sizeof(double), (size_t)h->nseg * sizeof(double),
                             h->acc_seg.size(), cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int pws_clear_accumulations(pws_handle* h) {
  PWS_CUDA(h, cudaSetDevice(h->device));
  PWS_CUDA(h, sync_all(h));
  if (h->d_acc_hru)
    PWS_CUDA(h, cudaMemset(h->d_acc_hru, 0, h->acc_hru.size() * h->stride * sizeof(double)));
  if (h->d_acc_seg)
    PWS_CUDA(h, cudaMemset(h->d_acc_seg, 0, h->acc_seg.size() * h->sstride * sizeof(double)));
  return 0;
}

// rows per day of the device-side result buffers (pws_run_device): the drained
// selection plus the accumulate-only variables behind it
extern "C" int pws_n_selected(pws_handle* h, int* n_f64, int* n_i32, int* n_seg) {
  if (n_f64) *n_f64 = h->n_f64;
  if (n_i32) *n_i32 = h->n_i32;
  if (n_seg) *n_seg = h->n_seg_sel;
  return 0;
}

// --------------------------------------------------------------- run
static int ensure_scratch(pws_handle* h, int ndays) {
  const int copies = (route_overlap(h) && h->nseg > 0) ? 2 : 1;
  if (ndays <= h->scratch_days && copies <= h->scratch_copies) return 0;
  ndays = std::max(ndays, h->scratch_days);
  cudaDeviceSynchronize();
  free_scratch(h);
  const size_t T = (size_t)ndays;
  for (int b = 0; b < 2; ++b) {
    PWS_CUDA(h, cudaMalloc(&h->d_cal[b], T * sizeof(int4)));
    PWS_CUDA(h, cudaMalloc(&h->d_viol[b], T * BUD_N * sizeof(int)));
    if (h->nseg > 0) PWS_CUDA(h, cudaMalloc(&h->d_hru_lat[b], T * h->stride * sizeof(double)));
  }
  for (int b = 0; b < copies && h->nseg > 0; ++b) {
    PWS_CUDA(h, cudaMalloc(&h->d_ots[b], T * 24 * std::max(h->n_ext_rows, 1) * sizeof(double)));
    PWS_CUDA(h, cudaMalloc(&h->d_progress[b], (size_t)(h->sstride + 32) * sizeof(int)));
    PWS_CUDA(h, cudaMalloc(&h->d_day_outvol[b], T * h->sstride * sizeof(double)));
    PWS_CUDA(h, cudaMalloc(&h->d_day_dstor[b], T * h->sstride * sizeof(double)));
    PWS_CUDA(h, cudaMalloc(&h->d_day_lat[b], T * h->sstride * sizeof(double)));
  }
  h->scratch_days = ndays;
  h->scratch_copies = copies;
  return 0;
}

// shared-memory carve-out (percent of the 228 KB maximum) that holds the
// resident CTAs of k_hru_column<., H> (1 KB per CTA is reserved by the system)
// `one_cta`: the grid has at most one CTA per SM (a small domain, a basin shard of
// a multi-GPU run): the second CTA's share stays L1, where the spilled registers
// and the global-memory parameter rows of the lone CTA then hit.
static int column_carveout_pct(int H, bool one_cta = false) {
  const size_t per_cta = column_smem_bytes(H) + 1024;
  const size_t ctas = one_cta ? 1 : std::max<size_t>(1, std::min<size_t>((228 * 1024) / per_cta, 65536 / (2 * H * 128)));
  return (int)std::min<size_t>(100, (per_cta * ctas * 100 + 228 * 1024 - 1) / (228 * 1024));
}

static cudaEvent_t new_event(pws_handle* /*h*/) {
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

// first timed launch since the last pws_sync: the time origin of the intervals
static void mark_base(pws_handle* h, cudaStream_t st) {
  if (h->ev_pairs_col.empty() && h->ev_pairs_ch.empty()) cudaEventRecord(h->ev_base, st);
}

// k_route_chunks<G> and its FlowGraph variant: shared-memory attribute, launch
template <int G>
static cudaError_t route_attrs() {
  cudaError_t e = cudaFuncSetAttribute(k_route_chunks<G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)route_smem_bytes(G));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_route_chunks<G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)route_smem_bytes(G));
  return e;
}
template <int G>
static void route_launch(bool graph, int nchunks, cudaStream_t st, const ChannelArgs& a) {
  if (graph)
    k_route_chunks<G, true><<<(unsigned)nchunks, PWS_CHUNK_THREADS, route_smem_bytes(G), st>>>(a);
  else
    k_route_chunks<G><<<(unsigned)nchunks, PWS_CHUNK_THREADS, route_smem_bytes(G), st>>>(a);
}

// enqueue the channel kernels for one block on stream `st`; `viol` = the
// block's verdict rows (or null: no channel budget)
static int enqueue_channel(pws_handle* h, cudaStream_t st, int ndays, const double* d_hru_lat,
                           const double* d_seg_lat_in, double* d_seg_out, int* viol,
                           const double* d_fg_obs = nullptr, double* d_fg_out = nullptr, int q = 0,
                           bool chained = false) {
  // q: which copy of the routing scratch; chained: the launch may start before the
  // previous one (on the other routing stream) has finished, chunk by chunk
  ChannelArgs a{};
  a.chunk_done = h->d_chunk_done;
  a.epoch = ++h->route_serial;
  a.wait_prev = chained ? 1 : 0;
  const bool graph = d_fg_out != nullptr;
  if (graph != h->flow_graph)
    PWS_FAIL(h, graph ? "FlowGraph run on a handle without pws_set_flow_graph"
                      : "this handle routes a FlowGraph (pws_set_flow_graph): use pws_run_flow_graph_host");
  a.kind = h->d_fg_kind; a.obs_col = h->d_fg_obs_col; a.obs = d_fg_obs; a.n_obs = h->fg_n_obs;
  a.flow_min = h->d_fg_flow_min;
  a.fg_out = d_fg_out;
  a.nseg = h->nseg;
  a.sstride = h->sstride;
  a.ndays = ndays;
  a.tsi = h->d_tsi; a.ts = h->d_ts; a.c0 = h->d_c0; a.c1 = h->d_c1; a.c2 = h->d_c2;
  a.up_ptr = h->d_up_ptr; a.up_idx = h->d_up_idx; a.hru_ptr = h->d_hru_ptr;
  a.hru_idx = h->d_hru_idx; a.flags = h->d_seg_flags; a.orig = h->d_orig;
  a.chunk_begin = h->d_chunk_begin; a.nchunks = h->nchunks;
  a.chunk_dmax = h->d_chunk_dmax; a.delay = h->d_delay; a.ext_row = h->d_ext_row; a.rts = h->d_rts;
  a.outflow_ts = h->d_outflow_ts; a.seg_inflow = h->d_seg_inflow;
  a.hru_lat = d_hru_lat; a.hstride = h->stride; a.seg_lat_in = d_seg_lat_in;
  a.ots = h->d_ots[q]; a.day_outvol = h->d_day_outvol[q]; a.day_dstor = h->d_day_dstor[q];
  a.day_lat = h->d_day_lat[q];
  a.seg_out = h->n_seg_sel > 0 ? d_seg_out : nullptr;
  a.n_seg_sel = h->n_seg_sel;
  memcpy(a.slot, h->slot_seg, sizeof a.slot);
  if (graph) {
    const int rows[6] = {FG_outflows, -1, FG_node_upstream_inflows, FG_node_outflows,
                         FG_node_storage_changes, FG_node_sink_source};
    for (int k = 0; k < 6; ++k) a.out_ptr[k] = rows[k] < 0 ? nullptr : d_fg_out + (int64_t)rows[k] * h->nseg;
    a.out_day_stride = (int64_t)FG_N * h->nseg;
  } else {
    const int vars[6] = {S_channel_outflow_vol, S_seg_lateral_inflow, S_seg_upstream_inflow, S_seg_outflow,
                         S_seg_stor_change, -1};
    for (int k = 0; k < 6; ++k)
      a.out_ptr[k] = (a.seg_out == nullptr || vars[k] < 0 || h->slot_seg[vars[k]] < 0)
                         ? nullptr : a.seg_out + (int64_t)h->slot_seg[vars[k]] * h->nseg;
    a.out_day_stride = (int64_t)h->n_seg_sel * h->nseg;
  }
  a.progress = h->d_progress[q];
  a.ticket = h->d_progress[q] + h->sstride;
  a.err = h->d_err;
  // progress counters (cross-chunk edges) and the chunk ticket start at 0
  PWS_CUDA(h, cudaMemsetAsync(h->d_progress[q], 0, (size_t)(h->sstride + 32) * sizeof(int), st));
  if (!h->route_attr_set) {
    if (route_attrs<PWS_ROUTE_HOURS>() != cudaSuccess || route_attrs<PWS_ROUTE_HOURS_SHALLOW>() != cudaSuccess ||
        route_attrs<PWS_ROUTE_HOURS_DEEP>() != cudaSuccess)
      PWS_CUDA(h, cudaGetLastError());
    h->route_attr_set = true;
  }
  cudaEvent_t e0 = new_event(h), e1 = new_event(h);
  mark_base(h, st);
  cudaEventRecord(e0, st);
  switch (route_hours(h)) {
    case PWS_ROUTE_HOURS_SHALLOW: route_launch<PWS_ROUTE_HOURS_SHALLOW>(graph, h->nchunks, st, a); break;
    case PWS_ROUTE_HOURS_DEEP: route_launch<PWS_ROUTE_HOURS_DEEP>(graph, h->nchunks, st, a); break;
    default: route_launch<PWS_ROUTE_HOURS>(graph, h->nchunks, st, a); break;
  }
  h->launches++;
  h->last_channel_launches += 1;
  if (viol) {
    k_channel_budget<<<ndays, 256, 0, st>>>(h->nseg, h->sstride, h->d_day_lat[q], h->d_day_outvol[q],
                                            h->d_day_dstor[q], viol);
    h->launches++;
    h->last_channel_launches += 1;
  }
  cudaEventRecord(e1, st);
  h->ev_pairs_ch.push_back({e0, e1});
  PWS_CUDA(h, cudaGetLastError());
  return 0;
}

// HRUs per CTA of k_hru_column.  Domains that fill the GPU several times: 256.  A
// smaller one is ONE wave and takes as long as its fullest SM, so the grid should (i)
// put one CTA on an SM, not two of 128 on some and one on others, and (ii) leave SMs
// for the routing CTAs of the previous block (a chunk takes a whole SM too): the
// smallest of 128, 192 (both with uncapped registers) and 256 whose grid fits beside
// the routing chunks; when none does, 128.  A quarter of the CONUS domain (27,540 HRUs, 36 chunks) per 256-day
// block: CTAs of 128 40.5 ms per pass, 192 without room for the routing 45.2, 256 with
// the routing on the 40 SMs it leaves 33.2 (profiles/r02_column_hrus_experiment.log).
static int column_hrus(pws_handle* h) {
  if (h->column_hrus_opt == 128 || h->column_hrus_opt == PWS_COLUMN_HRUS) return h->column_hrus_opt;
  if (h->column_hrus_opt == 192 && h->forcing_period == 0) return 192;
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
  const int64_t room = nsm - (h->nseg > 0 ? h->nchunks : 0);
  if ((h->nhru + 127) / 128 <= room) return 128;
  if (h->wide_column && h->forcing_period == 0 && (h->nhru + 191) / 192 <= room) return 192;
  if ((h->nhru + PWS_COLUMN_HRUS - 1) / PWS_COLUMN_HRUS <= room) return PWS_COLUMN_HRUS;
  return h->nhru <= (int64_t)nsm * PWS_COLUMN_HRUS ? 128 : PWS_COLUMN_HRUS;
}

// Column launches per block.  A CTA of the column kernel holds an SM's whole
// register file for a block of days, so a launch whose grid is not a multiple
// of the SM count leaves SMs idle in its last wave (CONUS: 431 CTAs = 2.9
// waves, an 8-GPU share of the 1024-member ensemble: 2.58).  The grid is
// therefore cut into `lanes` contiguous ranges of tiles, each on its own
// stream: a lane's next block only depends on that lane's previous block
// (HRUs are independent), so the tail of one lane's wave is filled by the head
// of the others' next block.
static int column_lanes(pws_handle* h, int ntiles, int H) {
  if (h->lanes_opt > 0) return std::min(h->lanes_opt, std::max(1, ntiles));
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
  const int slots = nsm * (H == 128 ? 2 : 1);  // CTAs resident at a time
  if (ntiles <= slots) return 1;               // one wave: nothing to smooth
  const int tail = ntiles % slots;             // CTAs of the last wave
  const int waves = (ntiles + slots - 1) / slots;
  // SM-time the last wave leaves idle, as a share of the launch: lanes cost a few
  // percent themselves (measured: 4 lanes on 2.58 waves 24.9 -> 23.9 ms, on 20.7
  // waves 194 -> 199 ms), so they are for launches that lose more than that
  if (tail == 0 || (slots - tail) * 12 < waves * slots) return 1;
  const int g = (ntiles + (2 * slots) / 3 - 1) / ((2 * slots) / 3);
  return std::max(1, std::min(g, pws_handle::kMaxLanes));
}

// Routing of block b beside the column kernels of block b+1?  A column CTA and a
// routing CTA each take a whole SM, so on a domain whose column grid fills the GPU
// the overlap only re-orders the same SM-time (measured on the CONUS domain: 130
// ms per pass against 128 serialised); it pays when the column grid leaves SMs
// idle (a basin shard of a multi-GPU run) or when lanes are on anyway.
static bool overlap_routing(pws_handle* h, int ntiles, int H, int lanes) {
  if (h->overlap_routing >= 0) return h->overlap_routing != 0;
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
  const int slots = nsm * (H == 128 ? 2 : 1);
  // (also when the routing chunks fit on the SMs the column grid leaves: column_hrus())
  return lanes > 1 || ntiles * 10 < slots * 9 || ntiles + h->nchunks <= nsm;
}

// number of forcing columns (the row length of prcp / tmax / tmin)
static int64_t forcing_columns(const pws_handle* h) {
  return h->forcing_period > 0 ? h->forcing_period : h->nhru;
}

static void fill_column_args(pws_handle* h, ColumnArgs& a, int ndays, const double* d_prcp,
                             const double* d_tmax, const double* d_tmin, int64_t fstride,
                             double* d_out_f64, int32_t* d_out_i32, int slot, bool want_viol) {
  a.P = h->P;
  a.state = h->d_state;
  a.prcp = d_prcp; a.tmax = d_tmax; a.tmin = d_tmin;
  a.fstride = fstride;
  a.fperiod = (uint32_t)h->forcing_period;
  // TMA bulk copies need 16-byte aligned rows of a multiple of 16 bytes
  // (an odd row is copied with one padding element, which must exist: fstride > nhru);
  // with a forcing period a CTA's row may wrap once: both pieces must stay aligned
  const int H = (h->column_hrus_opt == 128 || h->column_hrus_opt == PWS_COLUMN_HRUS)
                    ? h->column_hrus_opt : 0;
  (void)H;
  const int64_t ncol = forcing_columns(h);
  bool tma = h->tma && fstride % 2 == 0 &&
             ((uintptr_t)d_prcp | (uintptr_t)d_tmax | (uintptr_t)d_tmin) % 16 == 0;
  if (h->forcing_period > 0)
    tma = tma && ncol % 2 == 0 && ncol >= PWS_COLUMN_HRUS;
  else
    tma = tma && (h->nhru % 2 == 0 || fstride > h->nhru);
  a.use_tma = tma ? 1 : 0;
  a.cal = h->d_cal[slot];
  a.ndays = ndays;
  a.out_f64 = d_out_f64; a.out_i32 = d_out_i32;
  a.ostride = h->nhru;
  a.n_f64 = h->n_f64; a.n_i32 = h->n_i32;
  a.hru_lat = h->nseg > 0 ? h->d_hru_lat[slot] : nullptr;
  a.viol = want_viol ? h->d_viol[slot] : nullptr;
  a.err = h->d_err;
  a.ntiles = 0;
  a.hru_base = 0;
  for (int v = 0; v < PWS_N_HRU_VARS; ++v) {
    const int sl = h->slot_hru[v];
    const int64_t esz = names().hru_var_kind[v] == 0 ? 8 : 4;
    a.off[v] = sl < 0 ? -1 : (int64_t)sl * a.ostride * esz;
  }
}

// The arguments of a launch over the HRUs [h0, h0 + n) of the domain: every
// per-HRU pointer advanced by h0, P.nhru = the HRUs of the range that exist.
// With `offset_forcings` false (a forcing period) the forcing pointers stay and
// the kernel is told where the range starts (hru_base).
static void lane_view(ColumnArgs& a, int64_t h0, int64_t n, bool tiled, int H, bool offset_forcings) {
  HruParams& P = a.P;
#define X(nm) P.nm += h0;
  PWS_HRU_F64_PARAMS(X)
  PWS_MON_F64_PARAMS(X)
  PWS_HRU_F64_DERIVED(X)
  PWS_HRU_I32_PARAMS(X)
  PWS_MON_I32_PARAMS(X)
#undef X
  P.tmax_allsnow_c += h0;
  P.soltab_potsw += h0;
  P.soltab_horad_potsw += h0;
  P.nhru = std::min<int64_t>(P.nhru - h0, n);
  a.state += h0;
  if (offset_forcings) {
    a.prcp += h0; a.tmax += h0; a.tmin += h0;
  }
  a.hru_base = (uint32_t)h0;
  if (tiled) {  // [day][tile][variable][H]: the range's first tile
    if (a.out_f64) a.out_f64 += (h0 / H) * a.n_f64 * H;
    if (a.out_i32) a.out_i32 += (h0 / H) * a.n_i32 * H;
  } else {
    if (a.out_f64) a.out_f64 += h0;
    if (a.out_i32) a.out_i32 += h0;
  }
  if (a.hru_lat) a.hru_lat += h0;
}

// Enqueue one block of days (forcings / outputs are device pointers; what the
// caller enqueued on s_compute before -- an upload wait, k_widen_forcings --
// is ordered in front of it).  Stream structure per block, scratch slot p =
// block parity:
//   s_compute  wait(block in slot p done) -> calendar upload, counters := 0 -> ev_prep[p]
//   s_lane[g]  wait(ev_prep[p]) -> k_hru_column over the lane's tiles -> ev_lane[p][g]
//   s_route    wait(all ev_lane[p][.]) -> k_route_chunks, k_channel_budget,
//              verdict copy -> ev_block[p]
// so the column kernels of block b+1 run beside the routing of block b, and a
// lane does not wait for the other lanes' tails.  With the option
// "overlap_routing" 0 the lanes of block b+1 also wait for ev_block of block b.
// h_viol_dst (pinned or null) receives the block's [ndays][6] verdict rows.
// On return ev_block[p] marks the completion of the block; *slot_out = p.
static int enqueue_block(pws_handle* h, int ndays, const int32_t* h_cal, const double* d_prcp,
                         const double* d_tmax, const double* d_tmin, int64_t fstride,
                         double* d_out_f64, int32_t* d_out_i32, double* d_seg_out,
                         int32_t* h_viol_dst, int* slot_out) {
  // sink: nothing selected / everything in registry order / a subset
  int mode = SINK_SELECT;
  if (h->n_f64 + h->n_i32 == 0) {
    mode = SINK_NONE;
  } else if (h->n_f64 + h->n_i32 == PWS_N_HRU_VARS) {
    mode = SINK_FULL;
    for (int v = 0; v < PWS_N_HRU_VARS; ++v)
      if (h->slot_hru[v] != full_slot(v)) mode = SINK_SELECT;
  }
  if (h->tiled_outputs) {
    if (mode != SINK_FULL)
      PWS_FAIL(h, "tiled_outputs needs every HRU variable selected in registry order");
    mode = SINK_TILED;
  }
  if (mode == SINK_SELECT && h->nhm_sink) {
    // the reference's default output selection, in registry order: the kernel variant
    // that knows the selection at compile time
    bool same = true;
    for (int v = 0; v < PWS_N_HRU_VARS && same; ++v) same = h->slot_hru[v] == nhm_slot(v);
    if (same) mode = SINK_NHM;
  }
  if (ensure_scratch(h, ndays)) return 1;
  const bool want_viol = h_viol_dst != nullptr;
  const int p = (int)(h->blocks_enqueued & 1);
  const int prev = p ^ 1;
  h->blocks_enqueued++;
  if (slot_out) *slot_out = p;
  // the slot's previous block (two blocks ago) must be complete
  PWS_CUDA(h, cudaStreamWaitEvent(h->s_compute, h->ev_block[p], 0));
  PWS_CUDA(h, cudaMemcpyAsync(h->d_cal[p], h_cal, (size_t)ndays * sizeof(int4),
                              cudaMemcpyHostToDevice, h->s_compute));
  if (want_viol)
    PWS_CUDA(h, cudaMemsetAsync(h->d_viol[p], 0, (size_t)ndays * BUD_N * sizeof(int), h->s_compute));
  PWS_CUDA(h, cudaEventRecord(h->ev_prep[p], h->s_compute));
  ColumnArgs a{};
  fill_column_args(h, a, ndays, d_prcp, d_tmax, d_tmin, fstride, d_out_f64, d_out_i32, p, want_viol);
  const int H = column_hrus(h);
  const int ntiles = (int)((h->nhru + H - 1) / H);
  const int lanes = column_lanes(h, ntiles, H);
  const bool overlap = overlap_routing(h, ntiles, H, lanes);
  a.ntiles = ntiles;
#define PWS_LAUNCH_COLUMN_P(MODE, HH, PER, WIDE, GRID, ST)                                       \
  do {                                                                                           \
    const int key_ = (((MODE) * 3 + ((HH) == 128 ? 1 : (HH) == 192 ? 2 : 0)) * 2 + (PER)) * 2 + (WIDE); \
    if (!h->smem_attr_set[key_]) {                                                               \
      PWS_CUDA(h, (cudaFuncSetAttribute(k_hru_column<MODE, HH, PER, WIDE>,                       \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,             \
                                        (int)column_smem_bytes(HH))));                           \
      h->smem_attr_set[key_] = true;                                                             \
    }                                                                                            \
    /* no more shared memory than the CTAs of one SM need: the rest is L1 (a function            \
       attribute shared by every handle of the process, so it is set per launch) */              \
    PWS_CUDA(h, (cudaFuncSetAttribute(k_hru_column<MODE, HH, PER, WIDE>,                         \
                                      cudaFuncAttributePreferredSharedMemoryCarveout,            \
                                      column_carveout_pct(HH, one_cta))));                       \
    k_hru_column<MODE, HH, PER, WIDE><<<(GRID), 2 * (HH), column_smem_bytes(HH), (ST)>>>(al);    \
  } while (0)
#define PWS_LAUNCH_COLUMN(MODE, HH, GRID, ST)                                                    \
  do {                                                                                           \
    if (period) PWS_LAUNCH_COLUMN_P(MODE, HH, true, false, GRID, ST);                            \
    else PWS_LAUNCH_COLUMN_P(MODE, HH, false, false, GRID, ST);                                  \
  } while (0)
  // routing stream of this block: alternating when consecutive routing launches may
  // overlap (deep networks, route_overlap()); block b+2 then follows block b on its stream
  const bool chain_route = route_overlap(h);
  cudaStream_t sr = (chain_route && p == 1) ? h->s_route2 : h->s_route;
  const bool period = h->forcing_period > 0;
  int nsm_ = 148;
  cudaDeviceGetAttribute(&nsm_, cudaDevAttrMultiProcessorCount, h->device);
  const bool one_cta = ntiles <= nsm_ && h->one_cta_carveout;
  const bool wide = one_cta && !period && h->wide_column;
  for (int g = 0; g < lanes; ++g) {
    const int t0 = (int)((int64_t)ntiles * g / lanes), t1 = (int)((int64_t)ntiles * (g + 1) / lanes);
    if (t1 <= t0) {
      PWS_CUDA(h, cudaEventRecord(h->ev_lane[p][g], h->s_lane[g]));
      continue;
    }
    cudaStream_t st = h->s_lane[g];
    PWS_CUDA(h, cudaStreamWaitEvent(st, h->ev_prep[p], 0));
    if (!overlap) PWS_CUDA(h, cudaStreamWaitEvent(st, h->ev_block[prev], 0));
    ColumnArgs al = a;  // the lane's view: a sub-domain starting at its first HRU
    lane_view(al, (int64_t)t0 * H, (int64_t)(t1 - t0) * H, mode == SINK_TILED, H, !period);
    const unsigned grid = (unsigned)(t1 - t0);
    cudaEvent_t e0 = new_event(h), e1 = new_event(h);
    mark_base(h, st);
    cudaEventRecord(e0, st);
    if (H == 192) {  // one uncapped-register CTA of 192 HRUs per SM
      if (mode == SINK_FULL) PWS_LAUNCH_COLUMN_P(SINK_FULL, 192, false, true, grid, st);
      else if (mode == SINK_TILED) PWS_LAUNCH_COLUMN_P(SINK_TILED, 192, false, true, grid, st);
      else if (mode == SINK_NONE) PWS_LAUNCH_COLUMN_P(SINK_NONE, 192, false, true, grid, st);
      else if (mode == SINK_NHM) PWS_LAUNCH_COLUMN_P(SINK_NHM, 192, false, true, grid, st);
      else PWS_LAUNCH_COLUMN_P(SINK_SELECT, 192, false, true, grid, st);
    } else if (H == 128 && wide) {  // at most one CTA per SM: the variant without a register cap
      if (mode == SINK_FULL) PWS_LAUNCH_COLUMN_P(SINK_FULL, 128, false, true, grid, st);
      else if (mode == SINK_TILED) PWS_LAUNCH_COLUMN_P(SINK_TILED, 128, false, true, grid, st);
      else if (mode == SINK_NONE) PWS_LAUNCH_COLUMN_P(SINK_NONE, 128, false, true, grid, st);
      else if (mode == SINK_NHM) PWS_LAUNCH_COLUMN_P(SINK_NHM, 128, false, true, grid, st);
      else PWS_LAUNCH_COLUMN_P(SINK_SELECT, 128, false, true, grid, st);
    } else if (H == 128) {
      if (mode == SINK_FULL) PWS_LAUNCH_COLUMN(SINK_FULL, 128, grid, st);
      else if (mode == SINK_TILED) PWS_LAUNCH_COLUMN(SINK_TILED, 128, grid, st);
      else if (mode == SINK_NONE) PWS_LAUNCH_COLUMN(SINK_NONE, 128, grid, st);
      else if (mode == SINK_NHM) PWS_LAUNCH_COLUMN(SINK_NHM, 128, grid, st);
      else PWS_LAUNCH_COLUMN(SINK_SELECT, 128, grid, st);
    } else {
      if (mode == SINK_FULL) PWS_LAUNCH_COLUMN(SINK_FULL, PWS_COLUMN_HRUS, grid, st);
      else if (mode == SINK_TILED) PWS_LAUNCH_COLUMN(SINK_TILED, PWS_COLUMN_HRUS, grid, st);
      else if (mode == SINK_NONE) PWS_LAUNCH_COLUMN(SINK_NONE, PWS_COLUMN_HRUS, grid, st);
      else if (mode == SINK_NHM) PWS_LAUNCH_COLUMN(SINK_NHM, PWS_COLUMN_HRUS, grid, st);
      else PWS_LAUNCH_COLUMN(SINK_SELECT, PWS_COLUMN_HRUS, grid, st);
    }
    h->launches++;
    PWS_CUDA(h, cudaGetLastError());
    cudaEventRecord(e1, st);
    h->ev_pairs_col.push_back({e0, e1});
    PWS_CUDA(h, cudaEventRecord(h->ev_lane[p][g], st));
    PWS_CUDA(h, cudaStreamWaitEvent(sr, h->ev_lane[p][g], 0));
  }
#undef PWS_LAUNCH_COLUMN_P
#undef PWS_LAUNCH_COLUMN
  h->last_col_blocks++;
  // (without route_overlap the routing of block b-1 precedes this block's on s_route:
  // the segment state is handed over by stream order and the scratch is single)
  if (h->nseg > 0)
    if (enqueue_channel(h, sr, ndays, h->d_hru_lat[p], nullptr, d_seg_out,
                        want_viol ? h->d_viol[p] : nullptr, nullptr, nullptr, chain_route ? p : 0,
                        chain_route))
      return 1;
  if (want_viol)
    PWS_CUDA(h, cudaMemcpyAsync(h_viol_dst, h->d_viol[p], (size_t)ndays * BUD_N * sizeof(int),
                                cudaMemcpyDeviceToHost, sr));
  // the sums below are read-modify-write: behind the previous block's
  if (chain_route) PWS_CUDA(h, cudaStreamWaitEvent(sr, h->ev_block[prev], 0));
  if (h->accumulate_on && (!h->acc_hru.empty() || (!h->acc_seg.empty() && h->nseg > 0))) {
    if (mode == SINK_TILED) PWS_FAIL(h, "accumulations need the rows layout (tiled_outputs 0)");
    AccSlots sl{};
    if (!h->acc_hru.empty()) {
      if (h->acc_hru.size() > (size_t)AccSlots::kMax) PWS_FAIL(h, "too many accumulated variables");
      for (size_t k = 0; k < h->acc_hru.size(); ++k) sl.s[k] = h->acc_hru_slot[k];
      k_accumulate<<<dim3((unsigned)((h->nhru + 255) / 256), (unsigned)h->acc_hru.size()), 256, 0,
                     sr>>>(sl, d_out_f64, (int64_t)h->n_f64 * h->nhru, h->nhru, ndays, h->nhru,
                                   h->d_acc_hru, h->stride);
      h->launches++;
    }
    if (!h->acc_seg.empty() && h->nseg > 0) {
      for (size_t k = 0; k < h->acc_seg.size(); ++k) sl.s[k] = h->acc_seg_slot[k];
      k_accumulate<<<dim3((unsigned)((h->nseg + 255) / 256), (unsigned)h->acc_seg.size()), 256, 0,
                     sr>>>(sl, d_seg_out, (int64_t)h->n_seg_sel * h->nseg, h->nseg, ndays,
                                   h->nseg, h->d_acc_seg, h->sstride);
      h->launches++;
    }
    PWS_CUDA(h, cudaGetLastError());
  }
  PWS_CUDA(h, cudaEventRecord(h->ev_block[p], sr));
  return 0;
}

static void clear_timing(pws_handle* h) {
  for (auto& pr : h->ev_pairs_col) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  for (auto& pr : h->ev_pairs_ch) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  h->ev_pairs_col.clear();
  h->ev_pairs_ch.clear();
}

static int check_device_errors(pws_handle* h) {
  int errflag = 0;
  PWS_CUDA(h, cudaMemcpy(&errflag, h->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (errflag) {
    cudaMemset(h->d_err, 0, sizeof(int));
    if (errflag & 1)
      PWS_FAIL(h, "ERROR, in compute_interflow sos=0, please contact code developers "
                  "(prms_soilzone.py:1282-1287)");
    if (errflag & 16)
      PWS_FAIL(h, "k_route_chunks: a segment waited too long for its upstream segments");
    PWS_FAIL(h, "device error flag %d", errflag);
  }
  return 0;
}

// every stream of the handle is idle afterwards
static cudaError_t sync_all(pws_handle* h) {
  cudaError_t e = cudaStreamSynchronize(h->s_h2d);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->s_compute);
  for (int g = 0; g < pws_handle::kMaxLanes && e == cudaSuccess; ++g)
    e = cudaStreamSynchronize(h->s_lane[g]);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->s_route);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->s_route2);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->s_d2h);
  return e;
}

// length of the union of the (start, stop) intervals, in ms from ev_base
static double union_ms(pws_handle* h, const std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& ev) {
  std::vector<std::pair<float, float>> iv;
  for (auto& pr : ev) {
    float t0 = 0, t1 = 0;
    if (cudaEventElapsedTime(&t0, h->ev_base, pr.first) != cudaSuccess ||
        cudaEventElapsedTime(&t1, h->ev_base, pr.second) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    iv.push_back({t0, t1});
  }
  std::sort(iv.begin(), iv.end());
  double tot = 0.0;
  float lo = 0, hi = -1;
  for (auto& x : iv) {
    if (hi < lo || x.first > hi) {
      if (hi >= lo) tot += hi - lo;
      lo = x.first;
      hi = x.second;
    } else {
      hi = std::max(hi, x.second);
    }
  }
  if (hi >= lo) tot += hi - lo;
  return tot;
}

extern "C" int pws_sync(pws_handle* h) {
  PWS_CUDA(h, cudaSetDevice(h->device));
  PWS_CUDA(h, sync_all(h));
  if (!h->ev_pairs_col.empty() || !h->ev_pairs_ch.empty()) {
    h->n_col_launch = h->last_col_blocks;
    h->n_col_kernels = (int)h->ev_pairs_col.size();
    h->n_ch_launch = h->last_channel_launches;
    h->last_channel_launches = 0;
    h->last_col_blocks = 0;
    h->ms_column = union_ms(h, h->ev_pairs_col);
    h->ms_channel = union_ms(h, h->ev_pairs_ch);
    clear_timing(h);
  }
  return check_device_errors(h);
}

extern "C" int pws_run_device(pws_handle* h, int ndays, const int32_t* h_cal,
                              const double* d_prcp, const double* d_tmax, const double* d_tmin,
                              double* d_out_f64, int32_t* d_out_i32, double* d_seg_out,
                              int32_t* h_budget_viol) {
  if (!h->inited) PWS_FAIL(h, "pws_run_device before pws_init");
  if (ndays <= 0) return 0;
  PWS_CUDA(h, cudaSetDevice(h->device));
  return enqueue_block(h, ndays, h_cal, d_prcp, d_tmax, d_tmin, forcing_columns(h), d_out_f64,
                       d_out_i32, d_seg_out, h_budget_viol, nullptr);
}

extern "C" int pws_set_flow_graph(pws_handle* h, const int32_t* kind, const int32_t* obs_col,
                                  int n_obs) {
  if (kind == nullptr) {
    h->flow_graph = false;
    h->fg_kind.clear();
    h->fg_obs_col.clear();
    h->fg_flow_min.clear();
    h->fg_n_obs = 0;
    h->inited = false;
    return 0;
  }
  if (h->nseg == 0) PWS_FAIL(h, "pws_set_flow_graph: the handle has no segments (= nodes)");
  if (n_obs < 0) PWS_FAIL(h, "pws_set_flow_graph: n_obs < 0");
  h->fg_kind.assign(kind, kind + h->nseg);
  h->fg_obs_col.assign((size_t)h->nseg, -1);
  for (int64_t i = 0; i < h->nseg; ++i) {
    if (kind[i] < NODE_MUSKINGUM || kind[i] > NODE_SOURCESINK)
      PWS_FAIL(h, "pws_set_flow_graph: kind[%lld] = %d (0 Muskingum-Mann, 1 pass-through, 2 observed flow, "
                  "3 source / sink)", (long long)i, kind[i]);
    if (kind[i] >= NODE_OBSIN) {
      if (obs_col == nullptr || obs_col[i] < 0 || obs_col[i] >= n_obs)
        PWS_FAIL(h, "pws_set_flow_graph: node %lld is an ObsIn / SourceSink node without a column of "
                    "observations / requests", (long long)i);
      h->fg_obs_col[(size_t)i] = obs_col[i];
    }
  }
  h->fg_n_obs = n_obs;
  h->flow_graph = true;
  h->inited = false;
  return 0;
}

// SourceSinkFlowNode's flow_min (hydrology/source_sink_flow_node.py:16-34), one value per
// node (read for kind 3 only); after pws_set_flow_graph, before pws_init.  NULL = all zero.
extern "C" int pws_set_flow_node_params(pws_handle* h, const double* flow_min) {
  if (!h->flow_graph) PWS_FAIL(h, "pws_set_flow_node_params before pws_set_flow_graph");
  if (flow_min == nullptr) h->fg_flow_min.clear();
  else h->fg_flow_min.assign(flow_min, flow_min + h->nseg);
  h->inited = false;
  return 0;
}

// FlowGraph.calculate for `ndays` days (base/flow_graph.py:576-657): host arrays in,
// host arrays out, one launch of k_route_chunks<G, true>.  A compatibility path
// (the reference steps a Python object per node and hour); blocks of a year or so
// are what the Python FlowGraph class sends.
extern "C" int pws_run_flow_graph_host(pws_handle* h, int ndays, const double* inflows,
                                       const double* obs, double* out) {
  if (!h->inited) PWS_FAIL(h, "pws_run_flow_graph_host before pws_init");
  if (!h->flow_graph) PWS_FAIL(h, "pws_run_flow_graph_host: pws_set_flow_graph was not called");
  if (ndays <= 0) return 0;
  if (h->fg_n_obs > 0 && obs == nullptr) PWS_FAIL(h, "pws_run_flow_graph_host: ObsIn / SourceSink nodes need obs");
  PWS_CUDA(h, cudaSetDevice(h->device));
  PWS_CUDA(h, sync_all(h));
  if (ensure_scratch(h, ndays)) return 1;
  const size_t nn = (size_t)h->nseg, T = (size_t)ndays;
  double *d_in = nullptr, *d_obs = nullptr, *d_out = nullptr;
  auto release = [&]() { cudaFree(d_in); cudaFree(d_obs); cudaFree(d_out); };
  cudaError_t e = cudaMalloc(&d_in, T * nn * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_obs, std::max<size_t>(1, T * h->fg_n_obs) * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_out, T * FG_N * nn * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(d_in, inflows, T * nn * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && h->fg_n_obs > 0)
    e = cudaMemcpy(d_obs, obs, T * h->fg_n_obs * sizeof(double), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    release();
    PWS_CUDA(h, e);
  }
  const int rc = enqueue_channel(h, h->s_route, ndays, nullptr, d_in, nullptr, nullptr, d_obs, d_out);
  if (rc == 0) e = cudaStreamSynchronize(h->s_route);
  if (rc == 0 && e == cudaSuccess)
    e = cudaMemcpy(out, d_out, T * FG_N * nn * sizeof(double), cudaMemcpyDeviceToHost);
  release();
  if (rc) return rc;
  PWS_CUDA(h, e);
  int errflag = 0;
  PWS_CUDA(h, cudaMemcpy(&errflag, h->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (errflag & 16) PWS_FAIL(h, "pws_run_flow_graph_host: a chunk of nodes waited for its upstream chunk in vain");
  return 0;
}

extern "C" int pws_run_channel_device(pws_handle* h, int ndays, const double* d_seg_lat,
                                      double* d_seg_out) {
  if (!h->inited) PWS_FAIL(h, "pws_run_channel_device before pws_init");
  if (h->nseg == 0) PWS_FAIL(h, "pws_run_channel_device: no segments");
  if (ndays <= 0) return 0;
  PWS_CUDA(h, cudaSetDevice(h->device));
  if (ensure_scratch(h, ndays)) return 1;
  const int p = (int)(h->blocks_enqueued & 1);
  h->blocks_enqueued++;
  const bool chain_route = route_overlap(h);
  cudaStream_t sr = (chain_route && p == 1) ? h->s_route2 : h->s_route;
  // serialised: behind the previous call; overlapped: only behind the call before it
  // (stream order), the segment state being handed over chunk by chunk on the device
  if (!chain_route) PWS_CUDA(h, cudaStreamWaitEvent(sr, h->ev_block[p ^ 1], 0));
  if (enqueue_channel(h, sr, ndays, nullptr, d_seg_lat, d_seg_out, nullptr, nullptr, nullptr,
                      chain_route ? p : 0, chain_route))
    return 1;
  PWS_CUDA(h, cudaEventRecord(h->ev_block[p], sr));
  return 0;
}

// memcpy spread over a few threads (a single core moves ~10 GB/s, the PCIe link 30+)
static void par_memcpy(void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return;
  const size_t kChunk = (size_t)8 << 20;
  const int nthr = (int)std::min<size_t>(8, std::max<size_t>(1, bytes / kChunk));
  if (nthr <= 1) {
    memcpy(dst, src, bytes);
    return;
  }
  std::vector<std::thread> pool;
  const size_t per = (bytes / nthr + 63) & ~(size_t)63;
  for (int t = 0; t < nthr; ++t) {
    const size_t lo = (size_t)t * per, hi = std::min(bytes, lo + per);
    if (lo < hi) pool.emplace_back([=]() { memcpy((char*)dst + lo, (const char*)src + lo, hi - lo); });
  }
  for (auto& th : pool) th.join();
}

static bool is_pinned(const void* p) {
  if (p == nullptr) return true;
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

// Time-blocked host loop: block b's forcings go up on s_h2d while block b-1
// computes on s_compute and block b-2's outputs drain on s_d2h.  Pinned caller
// buffers are used directly; pageable ones (numpy arrays) would turn every
// cudaMemcpyAsync into a blocking copy, so they are staged through two pinned
// slots per direction, filled / emptied by the host while the GPU works.
//
// `esz` is the element size of the caller's forcing arrays: 8 (float64) or 4
// (float32: half the upload; the rows are widened on the device into the same
// float64 ring, k_widen_forcings, before the block's column kernel).
static int run_host_impl(pws_handle* h, int ndays, const int32_t* h_cal, const void* h_prcp,
                         const void* h_tmax, const void* h_tmin, size_t esz, double* h_out_f64,
                         int32_t* h_out_i32, double* h_seg_out, int32_t* h_budget_viol) {
  if (!h->inited) PWS_FAIL(h, "pws_run_host before pws_init");
  if (ndays <= 0) return 0;
  PWS_CUDA(h, cudaSetDevice(h->device));
  clear_timing(h);
  // n: units per variable and day in the result buffers (the padded tile rows
  // with tiled_outputs, see pws_output_tile); nc: forcing columns
  int64_t n = h->nhru;
  if (h->tiled_outputs) {
    const int H = column_hrus(h);
    n = (h->nhru + H - 1) / H * H;
  }
  const int64_t nc = forcing_columns(h);
  const int64_t cstride = round_up(nc, 32);  // forcing rows in the ring: 256-byte aligned
  // block length: option, else bounded by ~2 GiB per output ring slot
  int T = h->block_days;
  if (T <= 0) {
    const double bytes_per_day = (double)nc * 24.0 + (double)n * (8.0 * h->n_f64 + 4.0 * h->n_i32) +
                                 (double)h->nseg * (8.0 * h->n_seg_sel + 24 * 8.0);
    T = (int)std::max(1.0, std::min(366.0, 2.0e9 / bytes_per_day));
  }
  T = std::min(T, ndays);
  const size_t need_f64 = (size_t)T * std::max(h->n_f64, 1) * n;
  const size_t need_i32 = (size_t)T * std::max(h->n_i32, 1) * n;
  const size_t need_seg = (size_t)T * std::max(h->n_seg_sel, 1) * std::max<int64_t>(h->nseg, 1);
  if (T > h->ring_days || need_f64 > h->ring_f64_elems || need_i32 > h->ring_i32_elems ||
      need_seg > h->ring_seg_elems) {
    cudaDeviceSynchronize();
    const int days = std::max(T, h->ring_days);
    const size_t cf = std::max(need_f64, h->ring_f64_elems), ci = std::max(need_i32, h->ring_i32_elems),
                 cs = std::max(need_seg, h->ring_seg_elems);
    free_ring(h);
    h->ring_f64_elems = cf;
    h->ring_i32_elems = ci;
    h->ring_seg_elems = cs;
    for (int b = 0; b < 2; ++b) {
      PWS_CUDA(h, cudaMalloc(&h->d_forc[b], ((size_t)days * 3 * cstride + 2) * sizeof(double)));
      PWS_CUDA(h, cudaMalloc(&h->d_of64[b], h->ring_f64_elems * sizeof(double)));
      PWS_CUDA(h, cudaMalloc(&h->d_oi32[b], h->ring_i32_elems * sizeof(int32_t)));
      PWS_CUDA(h, cudaMalloc(&h->d_oseg[b], h->ring_seg_elems * sizeof(double)));
    }
    h->ring_days = days;
  }
  if (esz == 4 && h->forc32_days < h->ring_days) {
    cudaDeviceSynchronize();
    for (int b = 0; b < 2; ++b) {
      cudaFree(h->d_forc32[b]);
      h->d_forc32[b] = nullptr;
    }
    h->forc32_days = 0;
    for (int b = 0; b < 2; ++b)
      PWS_CUDA(h, cudaMalloc(&h->d_forc32[b], (size_t)h->ring_days * 3 * nc * sizeof(float)));
    h->forc32_days = h->ring_days;
  }
  if (ensure_scratch(h, T)) return 1;
  const bool want_f64 = h->n_f64_drain > 0 && h_out_f64, want_i32 = h->n_i32 > 0 && h_out_i32;
  const bool want_seg = h->n_seg_drain > 0 && h->nseg > 0 && h_seg_out;
  const int nfd = h->n_f64_drain, nsd = h->n_seg_drain;  // drained rows per day
  // Pageable caller buffers are staged through pinned slots, inputs and outputs
  // independently (a Model keeps its forcings in ordinary numpy arrays but can
  // hand in page-locked result buffers).
  const bool stage_in = !(is_pinned(h_prcp) && is_pinned(h_tmax) && is_pinned(h_tmin));
  const bool stage_out = !(is_pinned(want_f64 ? h_out_f64 : nullptr) &&
                           is_pinned(want_i32 ? h_out_i32 : nullptr) &&
                           is_pinned(want_seg ? h_seg_out : nullptr));
  if (stage_in && h->stage_days < h->ring_days) {
    for (int b = 0; b < 2; ++b) {
      cudaFreeHost(h->hs_in[b]);
      h->hs_in[b] = nullptr;
    }
    h->stage_days = 0;
    for (int b = 0; b < 2; ++b)
      PWS_CUDA(h, cudaHostAlloc(&h->hs_in[b], (size_t)h->ring_days * 3 * nc * sizeof(double), 0));
    h->stage_days = h->ring_days;
  }
  if (stage_out && (h->stage_f64_elems < h->ring_f64_elems || h->stage_i32_elems < h->ring_i32_elems ||
                    h->stage_seg_elems < h->ring_seg_elems)) {
    for (int b = 0; b < 2; ++b) {
      cudaFreeHost(h->hs_of[b]); h->hs_of[b] = nullptr;
      cudaFreeHost(h->hs_oi[b]); h->hs_oi[b] = nullptr;
      cudaFreeHost(h->hs_os[b]); h->hs_os[b] = nullptr;
    }
    h->stage_f64_elems = h->stage_i32_elems = h->stage_seg_elems = 0;
    for (int b = 0; b < 2; ++b) {
      PWS_CUDA(h, cudaHostAlloc(&h->hs_of[b], h->ring_f64_elems * sizeof(double), 0));
      PWS_CUDA(h, cudaHostAlloc(&h->hs_oi[b], h->ring_i32_elems * sizeof(int32_t), 0));
      PWS_CUDA(h, cudaHostAlloc(&h->hs_os[b], h->ring_seg_elems * sizeof(double), 0));
    }
    h->stage_f64_elems = h->ring_f64_elems;
    h->stage_i32_elems = h->ring_i32_elems;
    h->stage_seg_elems = h->ring_seg_elems;
  }
  // budget verdicts: always through a pinned buffer (a pageable destination
  // would make the copy, and with it the whole enqueue loop, synchronous)
  if (h_budget_viol && h->hs_viol_run_cap < (size_t)ndays * BUD_N) {
    cudaFreeHost(h->hs_viol_run);
    h->hs_viol_run = nullptr;
    h->hs_viol_run_cap = 0;
    PWS_CUDA(h, cudaHostAlloc(&h->hs_viol_run, (size_t)ndays * BUD_N * sizeof(int32_t), 0));
    h->hs_viol_run_cap = (size_t)ndays * BUD_N;
  }
  const int RT = h->ring_days;  // row capacity of the ring / stage slots
  const int nblocks = (ndays + T - 1) / T;
  // copy block `blk`'s drained outputs from its stage slot to the caller's arrays
  auto drain_stage = [&](int blk) -> int {
    const int b = blk & 1, d0 = blk * T, nd = std::min(T, ndays - d0);
    PWS_CUDA(h, cudaEventSynchronize(h->ev_d2h[b]));
    if (want_f64)
      par_memcpy(h_out_f64 + (size_t)d0 * nfd * n, h->hs_of[b], (size_t)nd * nfd * n * sizeof(double));
    if (want_i32)
      par_memcpy(h_out_i32 + (size_t)d0 * h->n_i32 * n, h->hs_oi[b],
                 (size_t)nd * h->n_i32 * n * sizeof(int32_t));
    if (want_seg)
      par_memcpy(h_seg_out + (size_t)d0 * nsd * h->nseg, h->hs_os[b],
                 (size_t)nd * nsd * h->nseg * sizeof(double));
    return 0;
  };
  int ring_slot[2] = {0, 0};  // scratch slot (enqueue_block) of the last block in ring slot b
  for (int blk = 0; blk < nblocks; ++blk) {
    const int b = blk & 1;
    const int d0 = blk * T;
    const int nd = std::min(T, ndays - d0);
    const char* srcs[3] = {(const char*)h_prcp + (size_t)d0 * nc * esz,
                           (const char*)h_tmax + (size_t)d0 * nc * esz,
                           (const char*)h_tmin + (size_t)d0 * nc * esz};
    // slot b: block blk-2 must have drained before its stage buffers are reused
    if (stage_out && blk >= 2 && drain_stage(blk - 2)) return 1;
    if (stage_in) {
      if (blk >= 2) PWS_CUDA(h, cudaEventSynchronize(h->ev_h2d[b]));  // hs_in[b] has been uploaded
      for (int k = 0; k < 3; ++k) {
        char* const slot = (char*)h->hs_in[b] + (size_t)k * RT * nc * esz;
        par_memcpy(slot, srcs[k], (size_t)nd * nc * esz);
        srcs[k] = slot;
      }
    }
    // the ring slot is free once the block that last used it is complete
    if (blk >= 2) PWS_CUDA(h, cudaStreamWaitEvent(h->s_h2d, h->ev_block[ring_slot[b]], 0));
    // forcing rows go into the ring with the padded row stride (16-byte aligned
    // rows for the TMA staging of k_hru_column)
    const size_t S8 = (size_t)cstride * sizeof(double), n8 = (size_t)nc * sizeof(double);
    for (int k = 0; k < 3; ++k) {
      if (esz == 8)
        PWS_CUDA(h, cudaMemcpy2DAsync(h->d_forc[b] + (size_t)k * RT * cstride, S8, srcs[k], n8, n8,
                                      (size_t)nd, cudaMemcpyHostToDevice, h->s_h2d));
      else  // float32 rows go up as they are (slot b was last read by block blk-2's widening)
        PWS_CUDA(h, cudaMemcpyAsync(h->d_forc32[b] + (size_t)k * RT * nc, srcs[k],
                                    (size_t)nd * nc * sizeof(float), cudaMemcpyHostToDevice, h->s_h2d));
    }
    PWS_CUDA(h, cudaEventRecord(h->ev_h2d[b], h->s_h2d));
    PWS_CUDA(h, cudaStreamWaitEvent(h->s_compute, h->ev_h2d[b], 0));
    if (blk >= 2) PWS_CUDA(h, cudaStreamWaitEvent(h->s_compute, h->ev_d2h[b], 0));
    if (esz == 4) {
      // (the float64 ring slot was last read by block blk-2: wait for it here,
      // enqueue_block only orders the scratch slot)
      if (blk >= 2) PWS_CUDA(h, cudaStreamWaitEvent(h->s_compute, h->ev_block[ring_slot[b]], 0));
      k_widen_forcings<<<dim3((unsigned)((nc + 255) / 256), (unsigned)nd, 3), 256, 0, h->s_compute>>>(
          h->d_forc32[b], h->d_forc[b], nc, cstride, RT, RT);
      h->launches++;
      PWS_CUDA(h, cudaGetLastError());
    }
    int p = 0;
    if (enqueue_block(h, nd, h_cal + (size_t)d0 * 4, h->d_forc[b],
                      h->d_forc[b] + (size_t)RT * cstride, h->d_forc[b] + (size_t)2 * RT * cstride,
                      cstride, h->d_of64[b], h->d_oi32[b], h->d_oseg[b],
                      h_budget_viol ? h->hs_viol_run + (size_t)d0 * BUD_N : nullptr, &p))
      return 1;
    ring_slot[b] = p;
    PWS_CUDA(h, cudaStreamWaitEvent(h->s_d2h, h->ev_block[p], 0));
    if (want_f64)  // the first nfd of the n_f64 rows of every day
      PWS_CUDA(h, cudaMemcpy2DAsync(stage_out ? h->hs_of[b] : h_out_f64 + (size_t)d0 * nfd * n,
                                    (size_t)nfd * n * sizeof(double), h->d_of64[b],
                                    (size_t)h->n_f64 * n * sizeof(double), (size_t)nfd * n * sizeof(double),
                                    (size_t)nd, cudaMemcpyDeviceToHost, h->s_d2h));
    if (want_i32)
      PWS_CUDA(h, cudaMemcpyAsync(stage_out ? h->hs_oi[b] : h_out_i32 + (size_t)d0 * h->n_i32 * n,
                                  h->d_oi32[b], (size_t)nd * h->n_i32 * n * sizeof(int32_t),
                                  cudaMemcpyDeviceToHost, h->s_d2h));
    if (want_seg)
      PWS_CUDA(h, cudaMemcpy2DAsync(stage_out ? h->hs_os[b] : h_seg_out + (size_t)d0 * nsd * h->nseg,
                                    (size_t)nsd * h->nseg * sizeof(double), h->d_oseg[b],
                                    (size_t)h->n_seg_sel * h->nseg * sizeof(double),
                                    (size_t)nsd * h->nseg * sizeof(double), (size_t)nd,
                                    cudaMemcpyDeviceToHost, h->s_d2h));
    PWS_CUDA(h, cudaEventRecord(h->ev_d2h[b], h->s_d2h));
  }
  if (stage_out)
    for (int blk = std::max(0, nblocks - 2); blk < nblocks; ++blk)
      if (drain_stage(blk)) return 1;
  const int rc = pws_sync(h);
  if (h_budget_viol)
    memcpy(h_budget_viol, h->hs_viol_run, (size_t)ndays * BUD_N * sizeof(int32_t));
  return rc;
}

extern "C" int pws_run_host(pws_handle* h, int ndays, const int32_t* h_cal, const double* h_prcp,
                            const double* h_tmax, const double* h_tmin, double* h_out_f64,
                            int32_t* h_out_i32, double* h_seg_out, int32_t* h_budget_viol) {
  return run_host_impl(h, ndays, h_cal, h_prcp, h_tmax, h_tmin, sizeof(double), h_out_f64,
                       h_out_i32, h_seg_out, h_budget_viol);
}

extern "C" int pws_run_host_f32(pws_handle* h, int ndays, const int32_t* h_cal, const float* h_prcp,
                                const float* h_tmax, const float* h_tmin, double* h_out_f64,
                                int32_t* h_out_i32, double* h_seg_out, int32_t* h_budget_viol) {
  return run_host_impl(h, ndays, h_cal, h_prcp, h_tmax, h_tmin, sizeof(float), h_out_f64,
                       h_out_i32, h_seg_out, h_budget_viol);
}

// --------------------------------------------------------------- sub-chains
static const char* const kInputNames[] = {"prcp", "tmax", "tmin",
#define X(n) #n,
                                          PWS_OV_INPUTS(X)
#undef X
};
extern "C" int pws_n_inputs(void) { return 3 + PWS_N_OV; }
extern "C" const char* pws_input_name(int i) {
  return (i >= 0 && i < 3 + PWS_N_OV) ? kInputNames[i] : nullptr;
}

// A subset of the chain's processes (process_mask, bits = BUD_* numbering; the
// atmosphere is implied by supplying prcp/tmax/tmin), everything else fed from
// the caller's arrays.  Compatibility path: synchronous per block of days.
extern "C" int pws_run_host_inputs(pws_handle* h, int ndays, const int32_t* h_cal, int n_in,
                                   const int32_t* in_idx, const double* const* h_in,
                                   uint32_t process_mask, double* h_out_f64, int32_t* h_out_i32,
                                   double* h_seg_out, int32_t* h_budget_viol) {
  if (!h->inited) PWS_FAIL(h, "pws_run_host_inputs before pws_init");
  if (h->n_f64 != h->n_f64_drain || h->n_seg_sel != h->n_seg_drain)
    PWS_FAIL(h, "pws_run_host_inputs: device accumulations are not available for sub-chains");
  if (ndays <= 0) return 0;
  PWS_CUDA(h, cudaSetDevice(h->device));
  PWS_CUDA(h, sync_all(h));
  for (int k = 0; k < n_in; ++k)
    if (in_idx[k] < 0 || in_idx[k] >= 3 + PWS_N_OV || h_in[k] == nullptr)
      PWS_FAIL(h, "pws_run_host_inputs: bad input index %d", in_idx[k]);
  clear_timing(h);
  const int64_t n = h->nhru;
  int T = h->block_days > 0 ? h->block_days : 64;
  T = std::min(T, ndays);
  if (ensure_scratch(h, T)) return 1;
  const bool channel = h->nseg > 0 && ((process_mask >> BUD_CHANNEL) & 1u);
  double *d_in = nullptr, *d_zero = nullptr, *d_of = nullptr, *d_os = nullptr;
  int32_t* d_oi = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_in); cudaFree(d_zero); cudaFree(d_of); cudaFree(d_oi); cudaFree(d_os);
  };
  const size_t blk = (size_t)T * n;
  cudaError_t e = cudaMalloc(&d_in, std::max<size_t>(n_in, 1) * blk * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_zero, blk * sizeof(double));
  if (e == cudaSuccess) e = cudaMemset(d_zero, 0, blk * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_of, std::max<size_t>(h->n_f64, 1) * blk * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&d_oi, std::max<size_t>(h->n_i32, 1) * blk * sizeof(int32_t));
  if (e == cudaSuccess)
    e = cudaMalloc(&d_os, (size_t)T * std::max(h->n_seg_sel, 1) * std::max<int64_t>(h->nseg, 1) *
                              sizeof(double));
  if (e != cudaSuccess) {
    cleanup();
    PWS_FAIL(h, "pws_run_host_inputs: %s", cudaGetErrorString(e));
  }
  int rc = 0;
  for (int d0 = 0; d0 < ndays && rc == 0; d0 += T) {
    const int nd = std::min(T, ndays - d0);
    OvArgs a{};
    const double* forc[3] = {d_zero, d_zero, d_zero};
    for (int k = 0; k < n_in; ++k) {
      double* dst = d_in + (size_t)k * blk;
      cudaMemcpyAsync(dst, h_in[k] + (size_t)d0 * n, (size_t)nd * n * sizeof(double),
                      cudaMemcpyHostToDevice, h->s_compute);
      if (in_idx[k] < 3) forc[in_idx[k]] = dst;
      else a.ov[in_idx[k] - 3] = dst;
    }
    cudaMemcpyAsync(h->d_cal[0], h_cal + (size_t)d0 * 4, (size_t)nd * sizeof(int4),
                    cudaMemcpyHostToDevice, h->s_compute);
    if (h_budget_viol)
      cudaMemsetAsync(h->d_viol[0], 0, (size_t)nd * BUD_N * sizeof(int), h->s_compute);
    fill_column_args(h, a.c, nd, forc[0], forc[1], forc[2], n, d_of, d_oi, 0, h_budget_viol != nullptr);
    a.c.fperiod = 0;  // every input of this path is [ndays][nhru]
    a.ovstride = n;
    a.proc_mask = process_mask;
    k_hru_column_ov<<<(unsigned)((n + 127) / 128), 128, 0, h->s_compute>>>(a);
    h->launches++;
    if (channel && rc == 0)
      rc = enqueue_channel(h, h->s_compute, nd, h->d_hru_lat[0], nullptr, d_os,
                           h_budget_viol ? h->d_viol[0] : nullptr);
    if (h_budget_viol)
      cudaMemcpyAsync(h_budget_viol + (size_t)d0 * BUD_N, h->d_viol[0], (size_t)nd * BUD_N * sizeof(int),
                      cudaMemcpyDeviceToHost, h->s_compute);
    if (h->n_f64 > 0 && h_out_f64)
      cudaMemcpyAsync(h_out_f64 + (size_t)d0 * h->n_f64 * n, d_of,
                      (size_t)nd * h->n_f64 * n * sizeof(double), cudaMemcpyDeviceToHost, h->s_compute);
    if (h->n_i32 > 0 && h_out_i32)
      cudaMemcpyAsync(h_out_i32 + (size_t)d0 * h->n_i32 * n, d_oi,
                      (size_t)nd * h->n_i32 * n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->s_compute);
    if (channel && h->n_seg_sel > 0 && h_seg_out)
      cudaMemcpyAsync(h_seg_out + (size_t)d0 * h->n_seg_sel * h->nseg, d_os,
                      (size_t)nd * h->n_seg_sel * h->nseg * sizeof(double), cudaMemcpyDeviceToHost,
                      h->s_compute);
    if (cudaStreamSynchronize(h->s_compute) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
      h->err = "pws_run_host_inputs: CUDA error in the block loop";
      rc = 1;
    }
  }
  cleanup();
  clear_timing(h);
  h->last_channel_launches = 0;
  if (rc) return rc;
  // budget verdicts of processes that are not part of the model are not reported
  if (h_budget_viol)
    for (int d = 0; d < ndays; ++d)
      for (int p = 0; p < BUD_N; ++p)
        if (!((process_mask >> p) & 1u)) h_budget_viol[(size_t)d * BUD_N + p] = 0;
  return check_device_errors(h);
}

// --------------------------------------------------------------- state
// PRMSChannel's prognostic state (prms_channel.py:384-386, 456-585: `outflow_ts`
// and yesterday's `seg_inflow`, carried from day to day), [nsegment] in the
// caller's segment order; the device keeps them in routing order.
static const char* const kSegStateNames[] = {"outflow_ts", "seg_inflow"};
extern "C" int pws_n_seg_state(void) { return 2; }
extern "C" const char* pws_seg_state_name(int i) { return (i >= 0 && i < 2) ? kSegStateNames[i] : nullptr; }

static double* seg_state_ptr(pws_handle* h, const char* name) {
  if (h->nseg == 0) return nullptr;
  if (!strcmp(name, "outflow_ts")) return h->d_outflow_ts;
  if (!strcmp(name, "seg_inflow")) return h->d_seg_inflow;
  return nullptr;
}

extern "C" int pws_get_state(pws_handle* h, const char* name, double* out) {
  if (!h->inited) PWS_FAIL(h, "pws_get_state before pws_init");
  PWS_CUDA(h, cudaSetDevice(h->device));
  if (double* d = seg_state_ptr(h, name)) {
    PWS_CUDA(h, sync_all(h));
    std::vector<double> tmp((size_t)h->nseg);
    PWS_CUDA(h, cudaMemcpy(tmp.data(), d, (size_t)h->nseg * sizeof(double), cudaMemcpyDeviceToHost));
    for (int64_t p = 0; p < h->nseg; ++p) out[h->h_orig[(size_t)p]] = tmp[(size_t)p];
    return 0;
  }
  const int k = find(names().states, name);
  if (k < 0) PWS_FAIL(h, "pws_get_state: unknown state '%s'", name);
  PWS_CUDA(h, sync_all(h));
  PWS_CUDA(h, cudaMemcpy(out, h->d_state + (size_t)k * h->stride, (size_t)h->nhru * sizeof(double),
                         cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int pws_set_state(pws_handle* h, const char* name, const double* in) {
  if (!h->inited) PWS_FAIL(h, "pws_set_state before pws_init");
  PWS_CUDA(h, cudaSetDevice(h->device));
  if (double* d = seg_state_ptr(h, name)) {
    PWS_CUDA(h, sync_all(h));
    std::vector<double> tmp((size_t)h->nseg);
    for (int64_t p = 0; p < h->nseg; ++p) tmp[(size_t)p] = in[h->h_orig[(size_t)p]];
    PWS_CUDA(h, cudaMemcpy(d, tmp.data(), (size_t)h->nseg * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
  }
  const int k = find(names().states, name);
  if (k < 0) PWS_FAIL(h, "pws_set_state: unknown state '%s'", name);
  PWS_CUDA(h, sync_all(h));
  PWS_CUDA(h, cudaMemcpy(h->d_state + (size_t)k * h->stride, in, (size_t)h->nhru * sizeof(double),
                         cudaMemcpyHostToDevice));
  return 0;
}

// --------------------------------------------------------------- PRMSEt
// hydrology/prms_et.py:22-181 for a block of days: six [ndays][nhru] inputs and
// two [nhru] parameters in host memory -> the three public variables.  No
// handle: PRMSEt has no state and belongs to no NHM model (the reference keeps
// it as "an intermediate class that should not exist in the future", :13-20).
extern "C" int pws_et_host(int device, int ndays, int64_t nhru, const double* h_potet,
                           const double* h_hru_impervevap, const double* h_hru_intcpevap,
                           const double* h_snow_evap, const double* h_dprst_evap_hru,
                           const double* h_perv_actet, const double* h_hru_percent_imperv,
                           const double* h_dprst_frac, double* h_unused_potet,
                           double* h_hru_perv_actet, double* h_hru_actet) {
  if (ndays <= 0 || nhru <= 0) return 0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
    g_create_error = "pws_et_host: no CUDA device; libpwsb200 has no CPU path";
    return 1;
  }
  cudaSetDevice(device);
  const size_t n = (size_t)ndays * (size_t)nhru, nb = n * sizeof(double), pb = (size_t)nhru * sizeof(double);
  double* d = nullptr;
  cudaError_t e = cudaMalloc(&d, 9 * nb + 2 * pb);
  if (e == cudaSuccess) {
    const double* src[6] = {h_potet, h_hru_impervevap, h_hru_intcpevap, h_snow_evap, h_dprst_evap_hru,
                            h_perv_actet};
    for (int k = 0; k < 6 && e == cudaSuccess; ++k) e = cudaMemcpy(d + k * n, src[k], nb, cudaMemcpyHostToDevice);
    double* dp = d + 9 * n;
    if (e == cudaSuccess) e = cudaMemcpy(dp, h_hru_percent_imperv, pb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dp + nhru, h_dprst_frac, pb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
      const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16);
      k_prms_et<<<grid, 256>>>((int64_t)n, nhru, d, d + n, d + 2 * n, d + 3 * n, d + 4 * n, d + 5 * n, dp,
                               dp + nhru, d + 6 * n, d + 7 * n, d + 8 * n);
      e = cudaGetLastError();
    }
    double* dst[3] = {h_unused_potet, h_hru_perv_actet, h_hru_actet};
    for (int k = 0; k < 3 && e == cudaSuccess; ++k)
      e = cudaMemcpy(dst[k], d + (6 + k) * n, nb, cudaMemcpyDeviceToHost);
  }
  cudaFree(d);
  if (e != cudaSuccess) {
    g_create_error = std::string("pws_et_host: ") + cudaGetErrorString(e);
    return 1;
  }
  return 0;
}

// --------------------------------------------------------------- instrumentation
extern "C" int64_t pws_launch_count(pws_handle* h) { return h->launches; }

extern "C" int pws_last_timing(pws_handle* h, double* ms_column, double* ms_channel) {
  if (ms_column) *ms_column = h->ms_column;
  if (ms_channel) *ms_channel = h->ms_channel;
  return 0;
}

extern "C" int pws_column_kernel_info(pws_handle* h, int* regs, int* local_bytes,
                                      int* max_blocks_per_sm) {
  cudaFuncAttributes at{};
  int nb = 0;
  if (column_hrus(h) == 128) {
    PWS_CUDA(h, cudaFuncGetAttributes(&at, k_hru_column<SINK_FULL, 128>));
    PWS_CUDA(h, cudaFuncSetAttribute(k_hru_column<SINK_FULL, 128>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)column_smem_bytes(128)));
    PWS_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_hru_column<SINK_FULL, 128>,
                                                              256, column_smem_bytes(128)));
  } else {
    PWS_CUDA(h, cudaFuncGetAttributes(&at, k_hru_column<SINK_FULL, PWS_COLUMN_HRUS>));
    PWS_CUDA(h, cudaFuncSetAttribute(k_hru_column<SINK_FULL, PWS_COLUMN_HRUS>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)column_smem_bytes(PWS_COLUMN_HRUS)));
    PWS_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                    &nb, k_hru_column<SINK_FULL, PWS_COLUMN_HRUS>, 2 * PWS_COLUMN_HRUS,
                    column_smem_bytes(PWS_COLUMN_HRUS)));
  }
  if (regs) *regs = at.numRegs;
  if (local_bytes) *local_bytes = (int)at.localSizeBytes;
  if (max_blocks_per_sm) *max_blocks_per_sm = nb;
  return 0;
}

extern "C" int pws_output_tile(pws_handle* h, int* tile_units, int64_t* padded_units) {
  PWS_CUDA(h, cudaSetDevice(h->device));
  const int H = column_hrus(h);
  if (tile_units) *tile_units = H;
  if (padded_units) *padded_units = (h->nhru + H - 1) / H * H;
  return 0;
}

extern "C" int pws_timer_mark(pws_handle* h, int which) {
  PWS_CUDA(h, cudaSetDevice(h->device));
  // blocks complete on s_route (and, without segments, on the lanes): join them
  for (int b = 0; b < 2; ++b) PWS_CUDA(h, cudaStreamWaitEvent(h->s_compute, h->ev_block[b], 0));
  PWS_CUDA(h, cudaEventRecord(which == 0 ? h->ev_col0 : h->ev_col1, h->s_compute));
  return 0;
}

extern "C" int pws_timer_elapsed_ms(pws_handle* h, double* ms) {
  float f = 0.f;
  PWS_CUDA(h, cudaEventSynchronize(h->ev_col1));
  PWS_CUDA(h, cudaEventElapsedTime(&f, h->ev_col0, h->ev_col1));
  *ms = f;
  return 0;
}

extern "C" int pws_last_launches(pws_handle* h, int* n_column, int* n_channel) {
  if (n_column) *n_column = h->n_col_kernels;
  if (n_channel) *n_channel = h->n_ch_launch;
  return 0;
}

extern "C" int pws_last_blocks(pws_handle* h, int* n_blocks, int* lanes) {
  if (n_blocks) *n_blocks = h->n_col_launch;
  if (lanes) *lanes = h->n_col_launch > 0 ? h->n_col_kernels / h->n_col_launch : 0;
  return 0;
}

extern "C" int pws_reset_state(pws_handle* h) {
  if (!h->inited) PWS_FAIL(h, "pws_reset_state before pws_init");
  PWS_CUDA(h, cudaSetDevice(h->device));
  PWS_CUDA(h, sync_all(h));
  k_hru_init<<<(unsigned)((h->nhru + 127) / 128), 128, 0, h->s_compute>>>(
      h->P, h->D, h->d_state, h->start_month, h->start_doy, h->adjust_parameters, h->d_err);
  h->launches++;
  if (h->nseg > 0) {
    PWS_CUDA(h, cudaMemsetAsync(h->d_outflow_ts, 0, (size_t)h->nseg * sizeof(double), h->s_compute));
    PWS_CUDA(h, cudaMemcpyAsync(h->d_seg_inflow, h->h_seg_inflow_init.data(),
                                (size_t)h->nseg * sizeof(double), cudaMemcpyHostToDevice,
                                h->s_compute));
  }
  if (h->d_acc_hru)
    PWS_CUDA(h, cudaMemsetAsync(h->d_acc_hru, 0, h->acc_hru.size() * h->stride * sizeof(double), h->s_compute));
  if (h->d_acc_seg)
    PWS_CUDA(h, cudaMemsetAsync(h->d_acc_seg, 0, h->acc_seg.size() * h->sstride * sizeof(double), h->s_compute));
  PWS_CUDA(h, cudaGetLastError());
  PWS_CUDA(h, cudaStreamSynchronize(h->s_compute));
  return 0;
}
