// pws_kernels.cuh -- the CUDA kernels of libpwsb200 (sm_100a, fp64):
//   k_soltab        PRMSSolarGeometry tables, once per run
//   k_hru_init      derived parameters + initial conditions, once per run
//   k_hru_column    the fused atmosphere..groundwater column, 1 thread / HRU,
//                   state in registers across a block of days
//   k_route_wavefront  Muskingum-Mann routing, all topological levels of the
//                   segment DAG concurrently over a block of days (persistent,
//                   cooperative); k_route_level = per-level fallback
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

struct GpuSink {
  const ColumnArgs& a;
  char* pf;
  char* pi;
  double* plat;
  unsigned violmask;
  int errmask;
  __device__ __forceinline__ void f(int var, double v) {
    const int64_t o = a.off[var];
    if (o >= 0) __stcs(reinterpret_cast<double*>(pf + o), v);
  }
  __device__ __forceinline__ void i(int var, int v) {
    const int64_t o = a.off[var];
    if (o >= 0) __stcs(reinterpret_cast<int*>(pi + o), v);
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

  GpuSink out{a, nullptr, nullptr, nullptr, 0u, 0};
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
// Muskingum-Mann routing (prms_channel.py:456-585).  Segments are stored in
// topological-level order ("pos"); thread = segment.  Every segment with a
// downstream neighbour publishes its 24 hourly outflows of a day as one
// contiguous 192-byte row of the HBM scratch ots[pos][day][hour]; a consumer
// gathers whole rows of its upstream segments (fully used sectors) before it
// enters the hourly recurrence, so no load sits on the recurrence's critical
// path.  Two drivers share route_day():
//   k_route_wavefront  ONE cooperative launch per block of days; all levels
//                      run concurrently, a segment starts day d as soon as
//                      its upstream segments have published day d (progress
//                      counters in HBM, acquire/release through L2).  Critical
//                      path = nlevels + ndays segment-days, not their product.
//   k_route_level      one launch per level (fallback when the segment
//                      threads cannot all be co-resident).
struct ChannelArgs {
  int64_t nseg, sstride;
  int ndays;
  // static, in pos order
  const int32_t* tsi;
  const double* ts;
  const double* c0;
  const double* c1;
  const double* c2;
  const int32_t* up_ptr;     // [nseg+1] CSR of upstream segments (pos)
  const int32_t* up_idx;
  const int32_t* hru_ptr;    // [nseg+1] CSR of contributing HRUs (ascending HRU index)
  const int32_t* hru_idx;
  const int32_t* has_down;   // 1 if tosegment > 0
  const int32_t* orig;       // pos -> original segment index
  // state, in pos order
  double* outflow_ts;
  double* seg_inflow;        // yesterday's daily mean inflow (prms_channel.py:384-386)
  // per-block inputs / scratch
  const double* hru_lat;     // [ndays][hstride] or null
  int64_t hstride;
  const double* seg_lat_in;  // [ndays][nseg] in ORIGINAL order (channel-only runs) or null
  double* ots;               // [nseg][ndays][24] hourly outflow of segments with a downstream
  int* progress;             // [nseg] days published (wavefront)
  double* day_outvol;        // [ndays][sstride] channel_outflow_vol (pos order)
  double* day_dstor;         // [ndays][sstride] seg_stor_change
  double* day_lat;           // [ndays][sstride] seg_lateral_inflow
  // outputs, original segment order: [ndays][n_seg_sel][nseg]
  double* seg_out;
  int n_seg_sel;
  int* err;
  short slot[PWS_N_SEG_VARS];
};

struct SegStatic {
  int tsi, ub, ue, hb, he, orig;
  unsigned fire;  // bit h set: hour h+1 is a routing update, (h + 1) % tsi == 0 (:529)
  bool down;
  double ts, c0, c1, c2;
};

__device__ __forceinline__ SegStatic seg_static(const ChannelArgs& a, int pos) {
  SegStatic g;
  g.tsi = a.tsi[pos];
  g.ts = a.ts[pos]; g.c0 = a.c0[pos]; g.c1 = a.c1[pos]; g.c2 = a.c2[pos];
  g.ub = a.up_ptr[pos]; g.ue = a.up_ptr[pos + 1];
  g.hb = a.hru_ptr[pos]; g.he = a.hru_ptr[pos + 1];
  g.down = a.has_down[pos] != 0;
  g.orig = a.orig[pos];
  g.fire = 0u;
  for (int h = 0; h < 24; ++h)
    if (g.tsi < 0 || ((h + 1) % g.tsi) == 0) g.fire |= 1u << h;
  return g;
}

// One segment, one day (24 hourly steps).  Upstream hours are summed in the
// reference's accumulation order (topological rank, prms_channel.py:566-571).
__device__ __forceinline__ void route_day(const ChannelArgs& a, const SegStatic& g, int pos,
                                          int d, double& outflow_ts, double& seg_inflow_prev) {
  // lateral inflow: HRU contributions in ascending HRU order (prms_channel.py:396-421)
  double lateral = 0.0;
  if (a.seg_lat_in != nullptr) {
    lateral = a.seg_lat_in[(int64_t)d * a.nseg + g.orig];
  } else {
    for (int k = g.hb; k < g.he; ++k)
      lateral += __ldcg(a.hru_lat + (int64_t)d * a.hstride + a.hru_idx[k]);
  }
  double up[24];
#pragma unroll
  for (int h = 0; h < 24; ++h) up[h] = 0.0;
  const int64_t day_off = (int64_t)d * 24;
  const int64_t row_len = (int64_t)a.ndays * 24;
  for (int k = g.ub; k < g.ue; ++k) {
    const double2* src =
        reinterpret_cast<const double2*>(a.ots + (int64_t)a.up_idx[k] * row_len + day_off);
#pragma unroll
    for (int h = 0; h < 12; ++h) {
      const double2 v = __ldcg(src + h);
      up[2 * h] += v.x;
      up[2 * h + 1] += v.y;
    }
  }
  double seg_inflow0 = seg_inflow_prev;  // _advance_variables, prms_channel.py:384-386
  double seg_inflow = seg_inflow0 * 0.0, seg_outflow = seg_inflow0 * 0.0;
  double inflow_ts = seg_inflow0 * 0.0, cur_sum = seg_inflow0 * 0.0;
  double2* dst = reinterpret_cast<double2*>(a.ots + (int64_t)pos * row_len + day_off);
  double o_even = 0.0;
#pragma unroll
  for (int h = 0; h < 24; ++h) {
    const double cur = lateral + up[h];
    seg_inflow += cur;
    inflow_ts += cur;
    cur_sum += up[h];
    if ((g.fire >> h) & 1u) {
      if (g.ts != 1.0) inflow_ts /= g.ts;  // x / 1.0 == x exactly
      if (g.tsi > 0) outflow_ts = inflow_ts * g.c0 + seg_inflow0 * g.c1 + outflow_ts * g.c2;
      else outflow_ts = inflow_ts;
      seg_inflow0 = inflow_ts;
      inflow_ts = 0.0;
    }
    seg_outflow += outflow_ts;
    if ((h & 1) == 0) o_even = outflow_ts;
    else if (g.down) __stcg(dst + (h >> 1), make_double2(o_even, outflow_ts));
  }
  seg_outflow /= 24.0;
  seg_inflow /= 24.0;
  const double upstream = cur_sum / 24.0;
  const double dstor = (seg_inflow - seg_outflow) * 86400.0;
  const double outvol = (g.down ? 0.0 : seg_outflow) * 86400.0;
  seg_inflow_prev = seg_inflow;
  const int64_t drow = (int64_t)d * a.sstride + pos;
  a.day_outvol[drow] = outvol;
  a.day_dstor[drow] = dstor;
  a.day_lat[drow] = lateral;
  if (a.seg_out != nullptr) {
    double* q = a.seg_out + (int64_t)d * a.n_seg_sel * a.nseg + g.orig;
    if (a.slot[S_channel_outflow_vol] >= 0) q[(int64_t)a.slot[S_channel_outflow_vol] * a.nseg] = outvol;
    if (a.slot[S_seg_lateral_inflow] >= 0) q[(int64_t)a.slot[S_seg_lateral_inflow] * a.nseg] = lateral;
    if (a.slot[S_seg_upstream_inflow] >= 0) q[(int64_t)a.slot[S_seg_upstream_inflow] * a.nseg] = upstream;
    if (a.slot[S_seg_outflow] >= 0) q[(int64_t)a.slot[S_seg_outflow] * a.nseg] = seg_outflow;
    if (a.slot[S_seg_stor_change] >= 0) q[(int64_t)a.slot[S_seg_stor_change] * a.nseg] = dstor;
  }
}

__global__ void __launch_bounds__(128)
k_route_level(const __grid_constant__ ChannelArgs a, int level_begin, int level_count) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= level_count) return;
  const int pos = level_begin + t;
  const SegStatic g = seg_static(a, pos);
  double outflow_ts = a.outflow_ts[pos];
  double seg_inflow_prev = a.seg_inflow[pos];
  for (int d = 0; d < a.ndays; ++d) route_day(a, g, pos, d, outflow_ts, seg_inflow_prev);
  a.outflow_ts[pos] = outflow_ts;
  a.seg_inflow[pos] = seg_inflow_prev;
}

// Persistent wavefront over (segment, day).  Cooperative launch: every CTA is
// resident, so a spinning consumer can never keep its producer off the GPU.
// Rounds are warp-synchronous: in each round a lane whose upstream segments
// have published day d routes that day and publishes; the other lanes poll
// again next round, so producer and consumer lanes of one warp cannot block
// each other.  A bounded number of idle rounds turns a would-be hang into an
// error flag.
#ifndef PWS_WAVEFRONT_MAX_IDLE
#define PWS_WAVEFRONT_MAX_IDLE (1 << 22)
#endif
__global__ void __launch_bounds__(128, 4)
k_route_wavefront(const __grid_constant__ ChannelArgs a) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = pos < a.nseg;
  SegStatic g{};
  double outflow_ts = 0.0, seg_inflow_prev = 0.0;
  if (live) {
    g = seg_static(a, pos);
    outflow_ts = a.outflow_ts[pos];
    seg_inflow_prev = a.seg_inflow[pos];
  }
  int d = live ? 0 : a.ndays;
  int idle = 0;
  while (__any_sync(0xffffffffu, d < a.ndays)) {
    bool ready = d < a.ndays;
    if (ready) {
      for (int k = g.ub; k < g.ue; ++k) {
        const int pu = *reinterpret_cast<volatile const int*>(a.progress + a.up_idx[k]);
        ready = ready && (pu > d);
      }
    }
    if (ready) {
      __threadfence();  // acquire: the published rows are visible after the counter
      route_day(a, g, pos, d, outflow_ts, seg_inflow_prev);
      ++d;
      if (g.down) {
        __threadfence();  // release: rows before the counter
        *reinterpret_cast<volatile int*>(a.progress + pos) = d;
      }
      idle = 0;
    } else if (d < a.ndays) {
      if (++idle > PWS_WAVEFRONT_MAX_IDLE) {
        atomicOr(a.err, 16);
        break;
      }
      __nanosleep(256);
    }
  }
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
