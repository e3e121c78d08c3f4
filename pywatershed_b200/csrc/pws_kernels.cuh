// pws_kernels.cuh -- the CUDA kernels of libpwsb200 (sm_100a, fp64):
//   k_soltab        PRMSSolarGeometry tables, once per run
//   k_hru_init      derived parameters + initial conditions, once per run
//   k_hru_column    the fused atmosphere..groundwater column, 1 thread / HRU,
//                   state in registers across a block of days
//   k_route_level   Muskingum-Mann routing of one topological level of the
//                   segment DAG over a whole block of days
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
  short slot[PWS_N_HRU_VARS];  // output slot of each variable, -1 = not requested
};

struct GpuSink {
  const ColumnArgs& a;
  double* pf;
  int32_t* pi;
  double* plat;
  unsigned violmask;
  int errmask;
  __device__ __forceinline__ void f(int var, double v) {
    const int sl = a.slot[var];
    if (sl >= 0) __stcs(pf + (int64_t)sl * a.ostride, v);
  }
  __device__ __forceinline__ void i(int var, int v) {
    const int sl = a.slot[var];
    if (sl >= 0) __stcs(pi + (int64_t)sl * a.ostride, v);
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

__global__ void __launch_bounds__(PWS_COLUMN_THREADS, PWS_COLUMN_MIN_BLOCKS)
k_hru_column(const __grid_constant__ ColumnArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < a.P.nhru;
  const int64_t ii = live ? i : a.P.nhru - 1;  // dead lanes shadow the last HRU, never store
  const int64_t S = a.P.stride;
  HruState s;
#define PWS_LOADST_F(name) s.name = a.state[(int64_t)ST_##name * S + ii];
#define PWS_LOADST_I(name) s.name = (int)a.state[(int64_t)ST_##name * S + ii];
#define X(name, kind) PWS_LOADST_##kind(name)
  PWS_HRU_STATE(X)
#undef X
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
    out.violmask = 0u;
    if (live) {
      out.pf = a.out_f64 + (int64_t)d * a.n_f64 * a.ostride + i;
      out.pi = a.out_i32 + (int64_t)d * a.n_i32 * a.ostride + i;
      out.plat = a.hru_lat ? a.hru_lat + (int64_t)d * S + i : nullptr;
      column_day(a.P, i, s, in, out);
    }
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
// Muskingum-Mann routing (prms_channel.py:456-585) of the segments of ONE
// topological level for every hour of a block of days.  Segments are stored
// in level order ("pos"); thread = segment.  The hourly outflow of every
// segment with a downstream neighbour goes to the HBM scratch
// ots[hour][pos] (coalesced across the warp); the next level's launch gathers
// its upstream hours from it in the reference's accumulation order.
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
  double* ots;               // [ndays*24][sstride]
  double* day_outvol;        // [ndays][sstride] channel_outflow_vol (pos order)
  double* day_dstor;         // [ndays][sstride] seg_stor_change
  double* day_lat;           // [ndays][sstride] seg_lateral_inflow
  // outputs, original segment order: [ndays][n_seg_sel][nseg]
  double* seg_out;
  int n_seg_sel;
  short slot[PWS_N_SEG_VARS];
};

__global__ void __launch_bounds__(128)
k_route_level(const __grid_constant__ ChannelArgs a, int level_begin, int level_count) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= level_count) return;
  const int pos = level_begin + t;
  const int tsi = a.tsi[pos];
  const double ts = a.ts[pos], c0 = a.c0[pos], c1 = a.c1[pos], c2 = a.c2[pos];
  const int ub = a.up_ptr[pos], ue = a.up_ptr[pos + 1];
  const int hb = a.hru_ptr[pos], he = a.hru_ptr[pos + 1];
  const bool down = a.has_down[pos] != 0;
  const int orig = a.orig[pos];
  double outflow_ts = a.outflow_ts[pos];
  double seg_inflow_prev = a.seg_inflow[pos];
  for (int d = 0; d < a.ndays; ++d) {
    // lateral inflow: HRU contributions in ascending HRU order (prms_channel.py:396-421)
    double lateral = 0.0;
    if (a.seg_lat_in != nullptr) {
      lateral = a.seg_lat_in[(int64_t)d * a.nseg + orig];
    } else {
      for (int k = hb; k < he; ++k)
        lateral += __ldcg(a.hru_lat + (int64_t)d * a.hstride + a.hru_idx[k]);
    }
    double seg_inflow0 = seg_inflow_prev;  // _advance_variables, prms_channel.py:384-386
    double seg_inflow = seg_inflow0 * 0.0, seg_outflow = seg_inflow0 * 0.0;
    double inflow_ts = seg_inflow0 * 0.0, cur_sum = seg_inflow0 * 0.0;
    for (int h = 0; h < 24; ++h) {
      const int64_t row = ((int64_t)d * 24 + h) * a.sstride;
      double up = seg_inflow * 0.0;
      for (int k = ub; k < ue; ++k) up += __ldcg(a.ots + row + a.up_idx[k]);
      const double cur = lateral + up;
      seg_inflow += cur;
      inflow_ts += cur;
      cur_sum += up;
      const bool fire = tsi < 0 ? true : ((h + 1) % tsi) == 0;
      if (fire) {
        inflow_ts /= ts;
        if (tsi > 0) outflow_ts = inflow_ts * c0 + seg_inflow0 * c1 + outflow_ts * c2;
        else outflow_ts = inflow_ts;
        seg_inflow0 = inflow_ts;
        inflow_ts = 0.0;
      }
      seg_outflow += outflow_ts;
      if (down) a.ots[row + pos] = outflow_ts;
    }
    seg_outflow /= 24.0;
    seg_inflow /= 24.0;
    const double upstream = cur_sum / 24.0;
    const double dstor = (seg_inflow - seg_outflow) * 86400.0;
    const double outvol = (down ? 0.0 : seg_outflow) * 86400.0;
    seg_inflow_prev = seg_inflow;
    const int64_t drow = (int64_t)d * a.sstride + pos;
    a.day_outvol[drow] = outvol;
    a.day_dstor[drow] = dstor;
    a.day_lat[drow] = lateral;
    if (a.seg_out != nullptr) {
      double* o = a.seg_out + (int64_t)d * a.n_seg_sel * a.nseg + orig;
      if (a.slot[S_channel_outflow_vol] >= 0) o[(int64_t)a.slot[S_channel_outflow_vol] * a.nseg] = outvol;
      if (a.slot[S_seg_lateral_inflow] >= 0) o[(int64_t)a.slot[S_seg_lateral_inflow] * a.nseg] = lateral;
      if (a.slot[S_seg_upstream_inflow] >= 0) o[(int64_t)a.slot[S_seg_upstream_inflow] * a.nseg] = upstream;
      if (a.slot[S_seg_outflow] >= 0) o[(int64_t)a.slot[S_seg_outflow] * a.nseg] = seg_outflow;
      if (a.slot[S_seg_stor_change] >= 0) o[(int64_t)a.slot[S_seg_stor_change] * a.nseg] = dstor;
    }
  }
  a.outflow_ts[pos] = outflow_ts;
  a.seg_inflow[pos] = seg_inflow_prev;
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
