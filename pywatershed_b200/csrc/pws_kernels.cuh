// pws_kernels.cuh -- the CUDA kernels of libpwsb200 (sm_100a, fp64):
//   k_soltab        PRMSSolarGeometry tables, once per run
//   k_hru_init      derived parameters + initial conditions, once per run
//   k_hru_column    the fused atmosphere..groundwater column, 1 thread / HRU,
//                   state in registers across a block of days
//   k_route_chunks  Muskingum-Mann routing: one CTA per chunk of connected
//                   segments, systolic hour-by-hour schedule over a block of
//                   days, hand-over through double-buffered shared memory
//   k_channel_budget per-day global channel mass balance
#pragma once
#include <cuda_runtime.h>

#include "hru_column.cuh"

namespace pws {

constexpr int kNdoy = 366;


// sin / cos / tan of the solar declination and r1 per day of year
// (solar_constants.py:10-43), filled from the host at pws_init
__constant__ double c_sol_sin_d[kNdoy];
__constant__ double c_sol_cos_d[kNdoy];
__constant__ double c_sol_tan_d[kNdoy];
__constant__ double c_sol_r1[kNdoy];

// Device version of the soltab tables (option "soltab_device"): grid
// (ceil(nhru/128), 366), one (doy, HRU) per thread.
__global__ void __launch_bounds__(128)
k_soltab(int64_t nhru, int64_t stride, const double* __restrict__ hru_slope,
         const double* __restrict__ hru_aspect, const double* __restrict__ hru_lat,
         double* __restrict__ potsw, double* __restrict__ horad,
         double* __restrict__ sunhrs, double* __restrict__ hru_cossl) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= nhru) return;
  double solt, sunh;
  // horizontal surface (prms_solar_geometry.py:141-147), then slope/aspect
  const SolHru flat = sol_hru(0.0, 0.0, hru_lat[i]);
  sol_day(flat, c_sol_sin_d[j], c_sol_cos_d[j], c_sol_tan_d[j], c_sol_r1[j], solt, sunh);
  horad[(int64_t)j * stride + i] = solt;
  const SolHru h = sol_hru(hru_slope[i], hru_aspect[i], hru_lat[i]);
  sol_day(h, c_sol_sin_d[j], c_sol_cos_d[j], c_sol_tan_d[j], c_sol_r1[j], solt, sunh);
  potsw[(int64_t)j * stride + i] = solt;
  sunhrs[(int64_t)j * stride + i] = sunh;
  if (j == 0) hru_cossl[i] = cos(atan(hru_slope[i]));  // prms_atmosphere.py:519
}

// ---------------------------------------------------------------- init
__global__ void __launch_bounds__(128)
k_hru_init(HruParams P, HruDerivedMut D, double* __restrict__ state, int start_month,
           int start_doy, int adjust_parameters, int* __restrict__ err) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.nhru) return;
  const int e = hru_init_one(P, D, state, i, start_month, start_doy, adjust_parameters);
  if (e) atomicOr(err, e);
}

// ---------------------------------------------------------------- column
struct ColumnArgs {
  HruParams P;
  double* state;            // [PWS_N_HRU_STATE][stride]
  const double* prcp;       // [ndays][nhru] (row stride = fstride)
  const double* tmax;
  const double* tmin;
  int64_t fstride;
  const int4* cal;          // [ndays] (doy, dowy, month, dom), device
  int ndays;
  double* out_f64;          // [ndays][n_f64][ostride]
  int32_t* out_i32;         // [ndays][n_i32][ostride]
  int64_t ostride;
  int n_f64, n_i32;
  double* hru_lat;          // [ndays][stride] lateral inflow per HRU (cfs) or null
  int* viol;                // [ndays][BUD_N] or null
  int* err;
  // byte offset of each public variable's row inside one day of out_f64 /
  // out_i32 (slot * ostride * sizeof), -1 = not requested
  int64_t off[PWS_N_HRU_VARS];
};

// How k_hru_column writes the public variables:
//   SINK_SELECT  any subset, row offsets looked up in ColumnArgs::off
//   SINK_FULL    every variable, in registry order: the row of a variable is a
//                compile-time constant, one IMAD.WIDE + one STG per value
//   SINK_NONE    nothing (routing-only runs: the lateral inflow is the output)
enum { SINK_SELECT = 0, SINK_FULL = 1, SINK_NONE = 2 };

constexpr int kHruVarKind[PWS_N_HRU_VARS] = {
#define KIND_F 0
#define KIND_I 1
#define KIND_B 1
#define X(name, kind, init) KIND_##kind,
    PWS_HRU_VARS(X)
#undef X
#undef KIND_F
#undef KIND_I
#undef KIND_B
};
// slot of a variable among the variables of its storage class (float64 /
// int32) when all are selected in registry order
constexpr int full_slot(int var) {
  int n = 0;
  for (int k = 0; k < var; ++k)
    if (kHruVarKind[k] == kHruVarKind[var]) ++n;
  return n;
}

template <int var>
struct FullSlot {
  static constexpr unsigned value = (unsigned)full_slot(var);
};

template <int kMode>
struct GpuSink {
  const ColumnArgs& a;
  char* pf;
  char* pi;
  double* plat;
  unsigned stride8, stride4;  // bytes between the rows of two variables
  unsigned violmask;
  int errmask;
  template <int var>
  __device__ __forceinline__ void f(double v) {
    if constexpr (kMode == SINK_FULL) {
      constexpr unsigned slot = FullSlot<var>::value;
      __stcs(reinterpret_cast<double*>(pf + (size_t)slot * stride8), v);
    } else if constexpr (kMode == SINK_SELECT) {
      const int64_t o = a.off[var];
      if (o >= 0) __stcs(reinterpret_cast<double*>(pf + o), v);
    }
  }
  template <int var>
  __device__ __forceinline__ void i(int v) {
    if constexpr (kMode == SINK_FULL) {
      constexpr unsigned slot = FullSlot<var>::value;
      __stcs(reinterpret_cast<int*>(pi + (size_t)slot * stride4), v);
    } else if constexpr (kMode == SINK_SELECT) {
      const int64_t o = a.off[var];
      if (o >= 0) __stcs(reinterpret_cast<int*>(pi + o), v);
    }
  }
  __device__ __forceinline__ void viol(int proc, bool open) {
    if (open) violmask |= 1u << proc;
  }
  __device__ __forceinline__ void lateral(double v) {
    if (plat) *plat = v;
  }
  __device__ __forceinline__ void error(int code) { errmask |= code; }
};

#ifndef PWS_COLUMN_THREADS
#define PWS_COLUMN_THREADS 128
#endif
#ifndef PWS_COLUMN_MIN_BLOCKS
#define PWS_COLUMN_MIN_BLOCKS 2
#endif
// 1: every per-HRU parameter column_day reads is staged once per launch (the
// monthly ones once per month) in shared memory, one private column per
// thread -- no inter-thread sharing, so no barrier.  0: read through L1/L2.
#ifndef PWS_COLUMN_SMEM
#define PWS_COLUMN_SMEM 1
#endif

constexpr int kDailyF64 = 0
#define X(n) +1
    PWS_DAILY_F64(X)
#undef X
    ;
constexpr int kDailyMonF64 = 0
#define X(n) +1
    PWS_DAILY_MON_F64(X)
#undef X
    ;
constexpr int kDailyI32 = 0
#define X(n) +1
    PWS_DAILY_I32(X) PWS_DAILY_MON_I32(X)
#undef X
    ;
constexpr size_t kColumnSmemBytes =
    PWS_COLUMN_SMEM ? (size_t)PWS_COLUMN_THREADS * ((kDailyF64 + kDailyMonF64) * 8 + kDailyI32 * 4)
                    : 0;

template <int kMode>
__global__ void __launch_bounds__(PWS_COLUMN_THREADS, PWS_COLUMN_MIN_BLOCKS)
k_hru_column(const __grid_constant__ ColumnArgs a) {
  constexpr int T = PWS_COLUMN_THREADS;
  const int tid = threadIdx.x;
  const int64_t i = (int64_t)blockIdx.x * T + tid;
  const bool live = i < a.P.nhru;
  const int64_t ii = live ? i : a.P.nhru - 1;  // dead lanes shadow the last HRU, never store
  const int64_t S = a.P.stride;
  HruState s;
#define PWS_LOADST_F(name) s.name = a.state[(int64_t)ST_##name * S + ii];
#define PWS_LOADST_I(name) s.name = (int)a.state[(int64_t)ST_##name * S + ii];
#define X(name, kind) PWS_LOADST_##kind(name)
  PWS_HRU_STATE(X)
#undef X

#if PWS_COLUMN_SMEM
  // Q: the parameter view column_day sees, indexed by tid.  Staged columns
  // live in shared memory; anything else is the global array re-based so that
  // [tid] is this thread's HRU.
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* const sp = reinterpret_cast<double*>(smem_raw);
  int32_t* const spi = reinterpret_cast<int32_t*>(sp + (size_t)(kDailyF64 + kDailyMonF64) * T);
  HruParams Q = a.P;
  Q.stride = 0;  // monthly rows: only the current month is staged
  {
    int k = 0;
#define X(n) sp[k * T + tid] = __ldg(a.P.n + ii); Q.n = sp + k * T; ++k;
    PWS_DAILY_F64(X)
#undef X
    k = 0;
#define X(n) spi[k * T + tid] = __ldg(a.P.n + ii); Q.n = spi + k * T; ++k;
    PWS_DAILY_I32(X)
#undef X
  }
  auto stage_month = [&](int month) {
    const int64_t mo = (int64_t)(month - 1) * S + ii;
    int k = kDailyF64;
#define X(n) sp[k * T + tid] = __ldg(a.P.n + mo); ++k;
    PWS_DAILY_MON_F64(X)
#undef X
    k = kDailyI32 - 1;
#define X(n) spi[k * T + tid] = __ldg(a.P.n + mo);
    PWS_DAILY_MON_I32(X)
#undef X
  };
  {
    int k = kDailyF64;
#define X(n) Q.n = sp + k * T; ++k;
    PWS_DAILY_MON_F64(X)
#undef X
    k = kDailyI32 - 1;
#define X(n) Q.n = spi + k * T;
    PWS_DAILY_MON_I32(X)
#undef X
  }
  int staged_month = 0;
  const int64_t pidx = tid;
#else
  const HruParams& Q = a.P;
  const int64_t pidx = ii;
#endif

  GpuSink<kMode> out{a, nullptr, nullptr, nullptr, (unsigned)(a.ostride * 8),
                     (unsigned)(a.ostride * 4), 0u, 0};
  // forcing prefetch: day d+1 is in flight while day d computes
  DayIn nxt;
  {
    const int4 c = a.cal[0];
    nxt.doy = c.x; nxt.dowy = c.y; nxt.month = c.z; nxt.dom = c.w;
    nxt.prcp = __ldcs(a.prcp + ii);
    nxt.tmax = __ldcs(a.tmax + ii);
    nxt.tmin = __ldcs(a.tmin + ii);
    nxt.potsw = __ldg(a.P.soltab_potsw + (int64_t)(c.x - 1) * S + ii);
    nxt.horad = __ldg(a.P.soltab_horad_potsw + (int64_t)(c.x - 1) * S + ii);
  }
  const int64_t day_f64 = (int64_t)a.n_f64 * a.ostride * 8;
  const int64_t day_i32 = (int64_t)a.n_i32 * a.ostride * 4;
  out.pf = reinterpret_cast<char*>(a.out_f64 + ii);
  out.pi = reinterpret_cast<char*>(a.out_i32 + ii);
  for (int d = 0; d < a.ndays; ++d) {
    const DayIn in = nxt;
    if (d + 1 < a.ndays) {
      const int4 c = a.cal[d + 1];
      const int64_t fo = (int64_t)(d + 1) * a.fstride + ii;
      nxt.doy = c.x; nxt.dowy = c.y; nxt.month = c.z; nxt.dom = c.w;
      nxt.prcp = __ldcs(a.prcp + fo);
      nxt.tmax = __ldcs(a.tmax + fo);
      nxt.tmin = __ldcs(a.tmin + fo);
      nxt.potsw = __ldg(a.P.soltab_potsw + (int64_t)(c.x - 1) * S + ii);
      nxt.horad = __ldg(a.P.soltab_horad_potsw + (int64_t)(c.x - 1) * S + ii);
    }
#if PWS_COLUMN_SMEM
    if (in.month != staged_month) {  // warp-uniform
      stage_month(in.month);
      staged_month = in.month;
    }
#endif
    out.violmask = 0u;
    if (live) {
      out.plat = a.hru_lat ? a.hru_lat + (int64_t)d * S + i : nullptr;
      column_day(Q, pidx, s, in, out);
    }
    out.pf += day_f64;
    out.pi += day_i32;
    if (a.viol != nullptr && __any_sync(0xffffffffu, out.violmask != 0u)) {
      for (int p = 0; p < BUD_CHANNEL; ++p) {
        const unsigned m = __ballot_sync(0xffffffffu, (out.violmask >> p) & 1u);
        if (m != 0u && (threadIdx.x & 31) == 0) atomicAdd(a.viol + d * BUD_N + p, __popc(m));
      }
    }
  }
  if (live) {
#define PWS_STOREST(name) a.state[(int64_t)ST_##name * S + i] = (double)s.name;
#define X(name, kind) PWS_STOREST(name)
    PWS_HRU_STATE(X)
#undef X
    if (out.errmask) atomicOr(a.err, out.errmask);
  }
}

// ---------------------------------------------------------------- channel
// Muskingum-Mann routing (prms_channel.py:456-585), thread = segment.
//
// The hourly recurrence of a segment needs, at hour h, the hour-h outflow of
// its upstream segments: a dependency chain as deep as the network, 24 times a
// day.  The kernel runs it as a SYSTOLIC schedule inside one CTA per chunk of
// connected segments:
//   * The host cuts the segment forest into connected pieces of at most
//     `channel_chunk` segments (whole outlet trees when they fit) and packs
//     them into chunks.  In a tree every segment has ONE downstream segment,
//     so with delay(i) = dmax - (hops from i to the root of its chunk-local
//     tree) every local upstream segment of i is exactly one step ahead of i.
//   * At step t, segment i advances hour t - delay(i).  The outflow it
//     publishes goes to a double-buffered shared-memory slot; its downstream
//     segment reads it in the next step.  One __syncthreads() per step is the
//     only synchronisation; a block of T days takes 24 T + dmax steps.
//   * Edges that cross chunks (only when a tree does not fit one chunk) go
//     through a global scratch row per day and a per-segment progress counter
//     with device-scope fences.  Chunks are numbered so that a chunk only
//     waits for lower-numbered ones, and a CTA takes its chunk from an atomic
//     ticket, so whatever it waits for is already running: a plain launch
//     cannot deadlock and may share the GPU with other kernels.
// Upstream flows are summed in the reference's accumulation order
// (topological rank, prms_channel.py:566-571).
struct ChannelArgs {
  int64_t nseg, sstride;
  int ndays;
  // static, in pos order (chunk, then delay, then original index)
  const int32_t* tsi;
  const double* ts;
  const double* rts;         // 1 / ts
  const double* c0;
  const double* c1;
  const double* c2;
  const int32_t* up_ptr;     // [nseg+1] CSR of upstream segments (pos)
  const int32_t* up_idx;
  const int32_t* hru_ptr;    // [nseg+1] CSR of contributing HRUs (ascending HRU index)
  const int32_t* hru_idx;
  const int32_t* flags;      // SEG_* bits
  const int32_t* orig;       // pos -> original segment index
  const int32_t* delay;      // systolic delay inside the chunk, hours
  const int32_t* ext_row;    // row of `ots` for a segment whose downstream is in another chunk, else -1
  const int32_t* chunk_begin;  // [nchunks+1]
  const int32_t* chunk_dmax;   // [nchunks] largest delay in the chunk
  int nchunks;
  int* ticket;               // chunk ticket counter, zeroed before each launch
  // state, in pos order
  double* outflow_ts;
  double* seg_inflow;        // yesterday's daily mean inflow (prms_channel.py:384-386)
  // per-block inputs / scratch
  const double* hru_lat;     // [ndays][hstride] or null
  int64_t hstride;
  const double* seg_lat_in;  // [ndays][nseg] in ORIGINAL order (channel-only runs) or null
  double* ots;               // [n_ext][ndays][24] hourly outflow across chunk boundaries
  int* progress;             // [nseg] days published, for edges that cross chunks
  double* day_outvol;        // [ndays][sstride] channel_outflow_vol (pos order)
  double* day_dstor;         // [ndays][sstride] seg_stor_change
  double* day_lat;           // [ndays][sstride] seg_lateral_inflow
  // outputs, original segment order: [ndays][n_seg_sel][nseg]
  double* seg_out;
  int n_seg_sel;
  int* err;
  short slot[PWS_N_SEG_VARS];
};

enum {
  SEG_HAS_DOWN = 1,     // tosegment > 0
  SEG_DOWN_LOCAL = 2,   // the downstream segment is in the same chunk
  SEG_UPS_LOCAL = 4,    // every upstream segment is in the same chunk
};

#ifndef PWS_CHUNK_THREADS
#define PWS_CHUNK_THREADS 512
#endif
#ifndef PWS_ROUTE_MAX_IDLE
#define PWS_ROUTE_MAX_IDLE (1 << 24)
#endif

__global__ void __launch_bounds__(PWS_CHUNK_THREADS, 1)
k_route_chunks(const __grid_constant__ ChannelArgs a) {
  constexpr int kZero = PWS_CHUNK_THREADS;  // a slot that always holds 0.0
  __shared__ double s_out[2][PWS_CHUNK_THREADS + 1];
  __shared__ int s_chunk;
  const int tid = threadIdx.x;
  if (tid == 0) {
    s_chunk = atomicAdd(a.ticket, 1);
    s_out[0][kZero] = 0.0;
    s_out[1][kZero] = 0.0;
  }
  __syncthreads();
  const int chunk = s_chunk;
  if (chunk >= a.nchunks) return;
  const int cb = a.chunk_begin[chunk], ce = a.chunk_begin[chunk + 1];
  const int dmax = a.chunk_dmax[chunk];
  const int pos = cb + tid;
  const bool live = pos < ce;
  const int T = a.ndays;
  const int64_t row_len = (int64_t)T * 24;

  // ---- static description of this thread's segment
  int tsi = 1, flags = 0, ub = 0, ue = 0, hb = 0, he = 0, orig = 0, delay = 0, ext_row = -1;
  double ts = 1.0, rts = 1.0, c0 = 0.0, c1 = 0.0, c2 = 0.0;
  unsigned fire = 0u;  // bit h set: hour h+1 is a routing update, (h + 1) % tsi == 0 (:529)
  int u0 = kZero, u1 = kZero, u2 = kZero;  // local upstream segments (slot in the chunk)
  int h0 = -1, h1 = -1, h2 = -1;           // contributing HRUs
  if (live) {
    tsi = a.tsi[pos];
    ts = a.ts[pos]; rts = a.rts[pos]; c0 = a.c0[pos]; c1 = a.c1[pos]; c2 = a.c2[pos];
    ub = a.up_ptr[pos]; ue = a.up_ptr[pos + 1];
    hb = a.hru_ptr[pos]; he = a.hru_ptr[pos + 1];
    flags = a.flags[pos];
    orig = a.orig[pos];
    delay = a.delay[pos];
    ext_row = a.ext_row[pos];
    for (int h = 0; h < 24; ++h)
      if (tsi < 0 || ((h + 1) % tsi) == 0) fire |= 1u << h;
  }
  const int nup = ue - ub, nh = he - hb;
  const bool fast_up = (flags & SEG_UPS_LOCAL) && nup <= 3;
  if (live && fast_up) {
    if (nup > 0) u0 = a.up_idx[ub] - cb;
    if (nup > 1) u1 = a.up_idx[ub + 1] - cb;
    if (nup > 2) u2 = a.up_idx[ub + 2] - cb;
  }
  const bool fast_lat = a.seg_lat_in != nullptr || nh <= 3;
  if (live && a.seg_lat_in == nullptr && nh <= 3) {
    if (nh > 0) h0 = a.hru_idx[hb];
    if (nh > 1) h1 = a.hru_idx[hb + 1];
    if (nh > 2) h2 = a.hru_idx[hb + 2];
  }
  const bool divide = ts != 1.0;  // x / 1.0 == x exactly
  const bool muskingum = tsi > 0;
  const bool down = (flags & SEG_HAS_DOWN) != 0;

  // Lateral inflow of day k as up to three raw addends (summed a window later,
  // so the loads stay in flight): HRU contributions in ascending HRU order
  // (prms_channel.py:396-421).
  double qa = 0.0, qb = 0.0, qr = 0.0;
  auto fetch_lateral = [&](int k) {
    qa = qb = qr = 0.0;
    if (!live || k < 0 || k >= T) return;
    if (a.seg_lat_in != nullptr) {
      qa = __ldcg(a.seg_lat_in + (int64_t)k * a.nseg + orig);
    } else if (fast_lat) {
      const double* row = a.hru_lat + (int64_t)k * a.hstride;
      if (h0 >= 0) qa = __ldcg(row + h0);
      if (h1 >= 0) qb = __ldcg(row + h1);
      if (h2 >= 0) qr = __ldcg(row + h2);
    } else {
      double lat = 0.0;
      for (int j = hb; j < he; ++j)
        lat += __ldcg(a.hru_lat + (int64_t)k * a.hstride + a.hru_idx[j]);
      qa = lat;
    }
  };
  // Results of the day a segment finished during the current 24-step window;
  // written to HBM by all threads together at the next window start, so the
  // stores are off the per-step critical path.
  double st_outvol = 0.0, st_dstor = 0.0, st_lat = 0.0, st_up = 0.0, st_out = 0.0;
  int st_day = -1;
  auto flush_day = [&]() {
    if (st_day < 0) return;
    const int64_t drow = (int64_t)st_day * a.sstride + pos;
    a.day_outvol[drow] = st_outvol;
    a.day_dstor[drow] = st_dstor;
    a.day_lat[drow] = st_lat;
    if (a.seg_out != nullptr) {
      double* q = a.seg_out + (int64_t)st_day * a.n_seg_sel * a.nseg + orig;
      if (a.slot[S_channel_outflow_vol] >= 0) q[(int64_t)a.slot[S_channel_outflow_vol] * a.nseg] = st_outvol;
      if (a.slot[S_seg_lateral_inflow] >= 0) q[(int64_t)a.slot[S_seg_lateral_inflow] * a.nseg] = st_lat;
      if (a.slot[S_seg_upstream_inflow] >= 0) q[(int64_t)a.slot[S_seg_upstream_inflow] * a.nseg] = st_up;
      if (a.slot[S_seg_outflow] >= 0) q[(int64_t)a.slot[S_seg_outflow] * a.nseg] = st_out;
      if (a.slot[S_seg_stor_change] >= 0) q[(int64_t)a.slot[S_seg_stor_change] * a.nseg] = st_dstor;
    }
    st_day = -1;
  };

  // ---- prognostic state
  double outflow_ts = 0.0, seg_inflow_prev = 0.0;
  if (live) {
    outflow_ts = a.outflow_ts[pos];
    seg_inflow_prev = a.seg_inflow[pos];
  }
  // The start of day k falls into window k + delay / 24; at every window start
  // the lateral inflow for the day starting in the NEXT window is requested.
  const int dd = delay / 24;
  fetch_lateral(-dd);
  double lat_window = 0.0;  // lateral inflow of the day that starts in the current window
  double lateral = 0.0, seg_inflow0 = 0.0, seg_inflow = 0.0, seg_outflow = 0.0;
  double inflow_ts = 0.0, cur_sum = 0.0;
  int d = 0, h = 0;
  const int nsteps = 24 * T + dmax;
  int x = -delay;  // hour of the block this segment advances in step t
  for (int t = 0, tw = 0, w = 0; t < nsteps; ++t, ++x) {
    if (tw == 0) {  // window start (uniform across the CTA)
      flush_day();
      lat_window = ((0.0 + qa) + qb) + qr;
      fetch_lateral(w + 1 - dd);
    }
    if (++tw == 24) {
      tw = 0;
      ++w;
    }
    const double* rd = s_out[(t & 1) ^ 1];  // published in step t - 1
    if (live && x >= 0 && x < 24 * T) {
      if (h == 0) {
        lateral = lat_window;
        seg_inflow0 = seg_inflow_prev;  // _advance_variables, prms_channel.py:384-386
        seg_inflow = seg_inflow0 * 0.0;
        seg_outflow = seg_inflow0 * 0.0;
        inflow_ts = seg_inflow0 * 0.0;
        cur_sum = seg_inflow0 * 0.0;
        if (!(flags & SEG_UPS_LOCAL)) {
          // upstream chunks must have published this day
          for (int k = ub; k < ue; ++k) {
            const int u = a.up_idx[k];
            if (u >= cb && u < ce) continue;
            int idle = 0;
            while (*reinterpret_cast<volatile const int*>(a.progress + u) <= d) {
              if (++idle > PWS_ROUTE_MAX_IDLE) {
                atomicOr(a.err, 16);
                break;
              }
              __nanosleep(64);
            }
          }
          __threadfence();  // acquire: the rows are visible after the counter
        }
      }
      double up;
      if (fast_up) {
        up = (rd[u0] + rd[u1]) + rd[u2];  // absent ones read the 0.0 slot
      } else {
        up = 0.0;
        for (int k = ub; k < ue; ++k) {
          const int u = a.up_idx[k];
          if (u >= cb && u < ce) up += rd[u - cb];
          else up += __ldcg(a.ots + (int64_t)a.ext_row[u] * row_len + d * 24 + h);
        }
      }
      const double cur = lateral + up;
      seg_inflow += cur;
      cur_sum += up;
      const double acc = inflow_ts + cur;
      // routing update (branch-free; most segments update every hour)
      const bool upd = (fire >> h) & 1u;
      const double xq = divide ? div_by_ts(acc, ts, rts) : acc;
      const double routed = muskingum ? (xq * c0 + seg_inflow0 * c1) + outflow_ts * c2 : xq;
      outflow_ts = upd ? routed : outflow_ts;
      seg_inflow0 = upd ? xq : seg_inflow0;
      inflow_ts = upd ? 0.0 : acc;
      seg_outflow += outflow_ts;
      s_out[t & 1][tid] = outflow_ts;
      if (ext_row >= 0) __stcg(a.ots + (int64_t)ext_row * row_len + d * 24 + h, outflow_ts);
      if (h == 23) {
        seg_outflow = div_by_ts(seg_outflow, 24.0, 1.0 / 24.0);
        seg_inflow = div_by_ts(seg_inflow, 24.0, 1.0 / 24.0);
        st_up = div_by_ts(cur_sum, 24.0, 1.0 / 24.0);
        st_dstor = (seg_inflow - seg_outflow) * 86400.0;
        st_outvol = (down ? 0.0 : seg_outflow) * 86400.0;
        st_out = seg_outflow;
        st_lat = lateral;
        st_day = d;
        seg_inflow_prev = seg_inflow;
        if (ext_row >= 0) {
          __threadfence();  // release: the day's row before the counter
          *reinterpret_cast<volatile int*>(a.progress + pos) = d + 1;
        }
        ++d;
        h = 0;
      } else {
        ++h;
      }
    }
    __syncthreads();
  }
  flush_day();
  if (live) {
    a.outflow_ts[pos] = outflow_ts;
    a.seg_inflow[pos] = seg_inflow_prev;
  }
}

// Global-basis channel budget (prms_channel.py:115,165-174; budget.py:388-416):
// one CTA per day, fixed-order tree reduction so the sums are reproducible.
__global__ void __launch_bounds__(256)
k_channel_budget(int64_t nseg, int64_t sstride, const double* __restrict__ day_lat,
                 const double* __restrict__ day_outvol, const double* __restrict__ day_dstor,
                 int* __restrict__ viol) {
  __shared__ double sh[3][256];
  const int d = blockIdx.x;
  double in = 0.0, outv = 0.0, ds = 0.0;
  for (int64_t j = threadIdx.x; j < nseg; j += blockDim.x) {
    in += day_lat[(int64_t)d * sstride + j] * 86400.0;
    outv += day_outvol[(int64_t)d * sstride + j];
    ds += day_dstor[(int64_t)d * sstride + j];
  }
  sh[0][threadIdx.x] = in;
  sh[1][threadIdx.x] = outv;
  sh[2][threadIdx.x] = ds;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + w];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + w];
      sh[2][threadIdx.x] += sh[2][threadIdx.x + w];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && viol != nullptr)
    viol[d * BUD_N + BUD_CHANNEL] = budget_open(sh[0][0], sh[1][0] + sh[2][0]) ? 1 : 0;
}

}  // namespace pws
