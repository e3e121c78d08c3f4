// pws_kernels.cuh -- the CUDA kernels of libpwsb200 (sm_100a, fp64):
//   k_soltab        PRMSSolarGeometry tables, once per run
//   k_hru_init      derived parameters + initial conditions, once per run
//   k_hru_column    the fused atmosphere..groundwater column over a block of
//                   days: two warp groups per CTA (upper / lower half of the
//                   column, one thread per HRU each) as a producer/consumer
//                   pipeline, state in registers, parameters in per-thread
//                   shared-memory columns, forcings staged by TMA bulk copies
//   k_hru_column_ov one thread per HRU, any subset of the processes, the rest
//                   replaced by caller-supplied inputs (compatibility path)
//   k_route_chunks  Muskingum-Mann routing: one CTA per chunk of connected
//                   segments, systolic hour-by-hour schedule over a block of
//                   days, hand-over through double-buffered shared memory
//   k_channel_budget per-day global channel mass balance
//   k_widen_forcings float32 forcing rows (as uploaded) -> the float64 forcing ring
#pragma once
#include <cuda_runtime.h>
#include <type_traits>

#include "hru_column.cuh"

// ---- build-time knobs of k_hru_column
// HRUs per CTA for large domains; the CTA has twice as many threads (an
// upper-half and a lower-half thread per HRU, see k_hru_column).  Domains that
// do not fill the GPU with such CTAs use 128 HRUs per CTA (twice as many CTAs).
#ifndef PWS_COLUMN_HRUS
#define PWS_COLUMN_HRUS 256
#endif
// 1: per warp group barriers at the process boundaries inside a day (see
// column_upper / column_lower): the warps of a group execute the same stretch
// of code at the same time and share its instruction fetches; 2 = also inside
// the large processes.  Worth 16 % when one warp ran the whole column; with
// the column split over two warp groups each warp walks half of the code and
// the barriers only cost (CONUS domain, ms per pass: 0 -> 118, 1 -> 119, 2 -> 123).
#ifndef PWS_COLUMN_SYNC
#define PWS_COLUMN_SYNC 0
#endif

// 1: the upper half's monthly parameters are staged in shared memory once per
// month; 0: read from global memory (L2) each day, which frees 35 KB that can
// hold the lower half's prognostic state (next knob).  Measured on the CONUS
// domain (ms per 10-year pass, default 118): monthly from L2 124; + lower state
// in shared memory 125 (136/120 registers), 122 (144/112), 132 (128/128).
#ifndef PWS_COLUMN_MONTHLY_SMEM
#define PWS_COLUMN_MONTHLY_SMEM 1
#endif
// The lower half is the critical path of the pipeline and spills at 120
// registers.  1: all 10 of its prognostic values live in shared memory instead
// of registers (20 KB, only fits with PWS_COLUMN_MONTHLY_SMEM 0); 2: the five
// that are touched in one short stretch of the day (impervious and depression
// storage) do, the five soil / groundwater stores stay in registers (10 KB:
// the space a second forcing slot would take).
#ifndef PWS_COLUMN_LOWER_STATE_SMEM
#define PWS_COLUMN_LOWER_STATE_SMEM 2
#endif
#ifndef PWS_LOWER_STATE_IN_SMEM
#define PWS_LOWER_STATE_IN_SMEM(X) \
  X(imperv_stor) X(hru_impervstor) X(dprst_vol_open) X(dprst_vol_clos) X(dprst_area_open)
#define PWS_LOWER_STATE_IN_REGS(X) \
  X(soil_moist) X(soil_rechr) X(slow_stor) X(pref_flow_stor) X(gwres_stor)
#endif
// Registers per thread of the two warp groups (setmaxnreg, multiples of 8;
// with 256 HRUs per CTA, upper + lower must be <= 256).  Unconstrained, the
// upper half wants 204 registers (29 prognostic values, the snow energy
// balance) and the lower half 148.  Measured on the CONUS domain, ms per
// 10-year pass, all lower state in registers: 104/152 142, 120/136 137,
// 128/128 131, 136/120 122, 144/112 130; with PWS_COLUMN_LOWER_STATE_SMEM 2
// (and the later store / forcing-slot changes): 128/128 117, 136/120 109,
// 144/112 104, 152/104 115, 160/96 117; with the tiled result layout
// (SINK_TILED): 136/120 105.6, 144/112 98.4, 152/104 105.5.
#ifndef PWS_COLUMN_REGS_UPPER
#define PWS_COLUMN_REGS_UPPER 144
#endif
#ifndef PWS_COLUMN_REGS_LOWER
#define PWS_COLUMN_REGS_LOWER 112
#endif

// Cache hint of the output stores (PTX st.global.<hint>).  The 143 values a
// thread writes per day must not displace the few spilled registers from the
// 28 KB of L1 left beside the shared-memory carve-out.  CONUS domain, column
// kernel ms per 10-year pass: cs 116.5, cg 121.5, wt 121.7, L1::no_allocate 112.5.
#ifndef PWS_COLUMN_ST
#define PWS_COLUMN_ST "L1::no_allocate"
#endif

namespace pws {

constexpr int kNdoy = 366;


// sin / cos / tan of the solar declination and r1 per day of year
// (solar_constants.py:10-43), filled from the host at pws_init
__constant__ double c_sol_sin_d[kNdoy];
__constant__ double c_sol_cos_d[kNdoy];
__constant__ double c_sol_tan_d[kNdoy];
__constant__ double c_sol_r1[kNdoy];

// PRMSSolarGeometry's tables on the device (the default; option "soltab_device"
// 0 computes them on host threads with the machine's libm instead): grid
// (ceil(nhru/128), kSoltabSlices), a thread = one HRU and every kSoltabSlices-th day
// of the year, so the per-HRU geometry (24 of the ~34 correctly rounded elementary
// functions of a table entry, CrMath, pws_crmath.cuh) is computed once per thread.
constexpr int kSoltabSlices = 6;
__global__ void __launch_bounds__(128)
k_soltab(int64_t nhru, int64_t stride, const double* __restrict__ hru_slope,
         const double* __restrict__ hru_aspect, const double* __restrict__ hru_lat,
         double* __restrict__ potsw, double* __restrict__ horad,
         double* __restrict__ sunhrs, double* __restrict__ hru_cossl) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nhru) return;
  // horizontal surface (prms_solar_geometry.py:141-147), then slope/aspect
  const SolHru flat = sol_hru<CrMath>(0.0, 0.0, hru_lat[i]);
  const SolHru h = sol_hru<CrMath>(hru_slope[i], hru_aspect[i], hru_lat[i]);
  for (int j = blockIdx.y; j < kNdoy; j += kSoltabSlices) {
    double solt, sunh;
    sol_day<CrMath>(flat, c_sol_sin_d[j], c_sol_cos_d[j], c_sol_tan_d[j], c_sol_r1[j], solt, sunh);
    horad[(int64_t)j * stride + i] = solt;
    sol_day<CrMath>(h, c_sol_sin_d[j], c_sol_cos_d[j], c_sol_tan_d[j], c_sol_r1[j], solt, sunh);
    potsw[(int64_t)j * stride + i] = solt;
    sunhrs[(int64_t)j * stride + i] = sunh;
  }
  if (blockIdx.y == 0) hru_cossl[i] = CrMath::cos(CrMath::atan(hru_slope[i]));  // prms_atmosphere.py:519
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
  // > 0: the forcing arrays are [ndays][fperiod] and HRU i reads column i mod
  // fperiod (a parameter ensemble over one base domain, member-major: every
  // member sees the same weather, examples/05_parameter_ensemble.ipynb)
  uint32_t fperiod;         // (32-bit: a 64-bit remainder is a subroutine call inside the day loop)
  int use_tma;              // forcing rows are 16-byte aligned: stage them with cp.async.bulk
  // A launch covers a contiguous range of the domain's CTA tiles (a launch per
  // lane, see enqueue_block in pws_b200.cu): every pointer above and below is
  // advanced to the range's first HRU on the host and P.nhru counts the HRUs of
  // the range, so the kernel indexes from 0 as if the range were the domain
  // (an offset added to blockIdx.x in the kernel cost 8 bytes of spills and 56
  // instructions).  ntiles = tiles of the WHOLE domain (the day stride of the
  // tiled layout); hru_base = domain index of the range's first HRU, only read
  // by the forcing-period variant.
  int ntiles;
  uint32_t hru_base;
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
//   SINK_TILED   every variable, in registry order, laid out [day][CTA tile]
//                [variable][H units]: the row of a variable is a compile-time
//                byte offset inside the tile, so a value costs one STG with an
//                immediate offset and no address arithmetic at all
//   SINK_NHM     the reference's default output selection (PWS_NHM_DEFAULT_HRU_VARS,
//                71 variables) in registry order, rows layout: the selection is
//                a compile-time fact, so the values nobody asked for are not
//                computed at all and the rows are constants as in SINK_FULL
enum { SINK_SELECT = 0, SINK_FULL = 1, SINK_NONE = 2, SINK_TILED = 3, SINK_NHM = 4 };

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
// ... and among the default output selection (-1: not part of it)
constexpr bool nhm_selected(int var) {
#define X(name) if (var == V_##name) return true;
  PWS_NHM_DEFAULT_HRU_VARS(X)
#undef X
  return false;
}
constexpr int nhm_slot(int var) {
  if (!nhm_selected(var)) return -1;
  int n = 0;
  for (int k = 0; k < var; ++k)
    if (nhm_selected(k) && kHruVarKind[k] == kHruVarKind[var]) ++n;
  return n;
}
template <int var>
struct NhmSlot {
  static constexpr int value = nhm_slot(var);
};

template <int kMode, int H>
struct GpuSink {
  static constexpr bool kReread = true;  // small integers are re-read from shared memory (hru_column.cuh)
  const ColumnArgs& a;
  char* pf;
  char* pi;
  double* plat;
  unsigned stride8, stride4;  // bytes between the rows of two variables
  unsigned violmask;
  int errmask;
  bool live;  // lanes past the last HRU compute a shadow column and store nothing
  int group_barrier;  // named barrier of this thread's warp group
  template <int var>
  __device__ __forceinline__ void f(double v) {
    if constexpr (kMode == SINK_FULL) {
      // one IMAD.WIDE (row * stride + pointer) and one STG (L1::no_allocate) per value
      constexpr unsigned slot = FullSlot<var>::value;
      unsigned long long addr;
      asm("mad.wide.u32 %0, %1, %2, %3;"
          : "=l"(addr)
          : "r"(stride8), "n"(slot), "l"((unsigned long long)pf));
      if (live) asm volatile("st.global." PWS_COLUMN_ST ".f64 [%0], %1;" ::"l"(addr), "d"(v));
    } else if constexpr (kMode == SINK_TILED) {
      constexpr unsigned byte_off = FullSlot<var>::value * (unsigned)H * 8u;
      if (live)
        asm volatile("st.global." PWS_COLUMN_ST ".f64 [%0+%1], %2;" ::"l"(pf), "n"(byte_off), "d"(v));
    } else if constexpr (kMode == SINK_NHM) {
      constexpr int slot = NhmSlot<var>::value;
      if constexpr (slot >= 0) {
        unsigned long long addr;
        asm("mad.wide.u32 %0, %1, %2, %3;"
            : "=l"(addr)
            : "r"(stride8), "n"(slot), "l"((unsigned long long)pf));
        if (live) asm volatile("st.global." PWS_COLUMN_ST ".f64 [%0], %1;" ::"l"(addr), "d"(v));
      }
    } else if constexpr (kMode == SINK_SELECT) {
      const int64_t o = a.off[var];
      if (o >= 0 && live)
        asm volatile("st.global." PWS_COLUMN_ST ".f64 [%0], %1;" ::"l"(pf + o), "d"(v));
    }
  }
  template <int var>
  __device__ __forceinline__ void i(int v) {
    if constexpr (kMode == SINK_FULL) {
      constexpr unsigned slot = FullSlot<var>::value;
      unsigned long long addr;
      asm("mad.wide.u32 %0, %1, %2, %3;"
          : "=l"(addr)
          : "r"(stride4), "n"(slot), "l"((unsigned long long)pi));
      if (live) asm volatile("st.global." PWS_COLUMN_ST ".s32 [%0], %1;" ::"l"(addr), "r"(v));
    } else if constexpr (kMode == SINK_TILED) {
      constexpr unsigned byte_off = FullSlot<var>::value * (unsigned)H * 4u;
      if (live)
        asm volatile("st.global." PWS_COLUMN_ST ".s32 [%0+%1], %2;" ::"l"(pi), "n"(byte_off), "r"(v));
    } else if constexpr (kMode == SINK_NHM) {
      constexpr int slot = NhmSlot<var>::value;
      if constexpr (slot >= 0) {
        unsigned long long addr;
        asm("mad.wide.u32 %0, %1, %2, %3;"
            : "=l"(addr)
            : "r"(stride4), "n"(slot), "l"((unsigned long long)pi));
        if (live) asm volatile("st.global." PWS_COLUMN_ST ".s32 [%0], %1;" ::"l"(addr), "r"(v));
      }
    } else if constexpr (kMode == SINK_SELECT) {
      const int64_t o = a.off[var];
      if (o >= 0 && live)
        asm volatile("st.global." PWS_COLUMN_ST ".s32 [%0], %1;" ::"l"(pi + o), "r"(v));
    }
  }
  __device__ __forceinline__ void viol(int proc, bool open) {
    if (open && live) violmask |= 1u << proc;
  }
  __device__ __forceinline__ void lateral(double v) {
    if (plat && live) asm volatile("st.global." PWS_COLUMN_ST ".f64 [%0], %1;" ::"l"(plat), "d"(v));
  }
  __device__ __forceinline__ void error(int code) {
    if (live) errmask |= code;
  }
  template <int id>
  __device__ __forceinline__ double ov(double v) const { return v; }
  template <int id>
  __device__ __forceinline__ int ovi(int v) const { return v; }
  __device__ __forceinline__ void sync() {
#if PWS_COLUMN_SYNC >= 1
    asm volatile("bar.sync %0, %1;" ::"r"(group_barrier), "n"(H) : "memory");
#endif
  }
  __device__ __forceinline__ void sync2() {
#if PWS_COLUMN_SYNC >= 2
    asm volatile("bar.sync %0, %1;" ::"r"(group_barrier), "n"(H) : "memory");
#endif
  }
};

#define PWS_COUNT(n) +1
constexpr int kUpperF64 = 0 PWS_UPPER_F64(PWS_COUNT);
constexpr int kUpperMonF64 = 0 PWS_UPPER_MON_F64(PWS_COUNT);
constexpr int kUpperI32 = 0 PWS_UPPER_I32(PWS_COUNT) PWS_UPPER_MON_I32(PWS_COUNT);
constexpr int kLowerF64 = 0 PWS_LOWER_F64(PWS_COUNT);
constexpr int kLowerI32 = 0 PWS_LOWER_I32(PWS_COUNT);
#undef PWS_COUNT
constexpr int kMidF64 = 10, kMidI32 = 2;  // ColumnMid
#ifndef PWS_COLUMN_MID_SLOTS
#define PWS_COLUMN_MID_SLOTS 1  // days the upper half may run ahead of the lower half (2: no gain)
#endif
constexpr int kMidSlots = PWS_COLUMN_MID_SLOTS;
constexpr int kLowerState = 0
#define X(name, kind) +1
    PWS_HRU_STATE_LOWER(X)
#undef X
    ;
constexpr int kUpperMonStaged = PWS_COLUMN_MONTHLY_SMEM ? kUpperMonF64 : 0;
constexpr int kLowerStateSome = 0
#define X(name) +1
    PWS_LOWER_STATE_IN_SMEM(X)
#undef X
    ;
constexpr int kLowerStateStaged =
    PWS_COLUMN_LOWER_STATE_SMEM == 1 ? kLowerState : (PWS_COLUMN_LOWER_STATE_SMEM == 2 ? kLowerStateSome : 0);
constexpr int kColF64 =
    kUpperF64 + kUpperMonStaged + kLowerF64 + kMidSlots * kMidF64 + kLowerStateStaged;
constexpr int kColI32 = kUpperI32 + kLowerI32 + kMidSlots * kMidI32;
// forcing stage of the upper half: one day x {prcp, tmax, tmin, soltab_potsw row,
// soltab_horad_potsw row} x H, filled by TMA bulk copies, + its mbarrier.  A day's
// five values are read into registers first thing in the morning, so the slot
// can take the next day's rows while this day computes.
constexpr int kForcRows = 5;
template <int H>
struct ColumnSmem {
  static constexpr size_t kParamBytes = (size_t)H * (kColF64 * 8 + kColI32 * 4);
  static constexpr size_t kForcBytes = (size_t)kForcRows * H * 8;
  // + mbarrier (16) + two calendar rows (32) + one flag byte per HRU for the lower half
  static constexpr size_t kBytes = kParamBytes + kForcBytes + 16 + 32 + (size_t)H;
  static_assert(kParamBytes % 16 == 0, "forcing stage must be 16-byte aligned");
};
constexpr size_t column_smem_bytes(int H) {
  return (size_t)H * (kColF64 * 8 + kColI32 * 4) + (size_t)kForcRows * H * 8 + 16 + 32 + (size_t)H;
}

// named barriers of k_hru_column (0 is __syncthreads)
enum { BAR_UPPER = 1, BAR_LOWER = 2, BAR_MID_FULL = 3, BAR_MID_EMPTY = 3 + PWS_COLUMN_MID_SLOTS };
static_assert(3 + 2 * PWS_COLUMN_MID_SLOTS <= 16, "named barriers");

// ---- TMA bulk copy + mbarrier (sm_90+ PTX; SASS: UBLKCP / SYNCS)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                             uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; "
        "selp.u32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}

__device__ __forceinline__ void bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// The lower half's prognostic state as references (into shared memory)
struct LowerStateRef {
#define X(name, kind) double& name;
  PWS_HRU_STATE_LOWER(X)
#undef X
};
// ... or only part of it (PWS_COLUMN_LOWER_STATE_SMEM 2)
struct LowerStateMixed {
#define X(name) double& name;
  PWS_LOWER_STATE_IN_SMEM(X)
#undef X
#define X(name) double name;
  PWS_LOWER_STATE_IN_REGS(X)
#undef X
};

// The fused column.  A CTA owns H = PWS_COLUMN_HRUS HRUs and has 2 H threads:
// warps [0, H/32) run the UPPER half of the column (atmosphere, canopy, snow)
// and warps [H/32, 2H/32) the LOWER half (runoff, soilzone, groundwater) of
// the same HRUs, one day behind, as a producer/consumer pipeline: the upper
// thread of an HRU leaves the ten values the lower half needs (ColumnMid) in
// shared memory and the two groups hand the slot over with named barriers
// (BAR_MID_FULL / BAR_MID_EMPTY).  Nothing in the upper half depends on the
// lower half, so the groups overlap completely; each thread carries only its
// half of the prognostic state in registers and stages only its half of the
// parameters, which doubles the number of resident warps for the same
// register file and shared memory, and each warp walks only its half of the
// code.
// kPeriod: the forcing arrays are [ndays][fperiod] (shared-forcing ensembles); a
// separate instantiation because the wrap-around addressing costs the plain
// kernel 16 bytes of spills and 120 instructions even when it is not taken.
// kWide: one CTA per SM (launch bounds (2 H, 1), no setmaxnreg): for grids of at
// most one CTA per SM -- small domains, the basin shard of a multi-GPU run -- where
// the register file is not what limits residency, so neither half needs to spill.
template <int kMode, int H, bool kPeriod = false, bool kWide = false>
__global__ void __launch_bounds__(2 * H, (2 * H <= 256 && !kWide) ? 512 / (2 * H) : 1)
k_hru_column(const __grid_constant__ ColumnArgs a) {
  static_assert(!kWide || H == 128 || H == 192, "the wide variant is built for 128 or 192 HRUs per CTA");
  // setmaxnreg acts on whole warpgroups (4 warps): each half must be a multiple of one
  static_assert(kWide ? H % 32 == 0 : H % 128 == 0, "HRUs per CTA must be a multiple of 128 (wide: of 32)");
  constexpr size_t kColumnParamBytes = ColumnSmem<H>::kParamBytes;
  constexpr size_t kColumnForcBytes = ColumnSmem<H>::kForcBytes;
  const int tid = threadIdx.x;
  const bool upper = tid < H;  // warpgroup-uniform
  const int r = upper ? tid : tid - H;
  const int64_t tile = (int64_t)blockIdx.x;
  const int64_t i = tile * H + r;
  const bool live = i < a.P.nhru;
  const int64_t ii = live ? i : a.P.nhru - 1;  // dead lanes shadow the last HRU, never store
  const int64_t S = a.P.stride;

  // shared memory: one private column per thread and staged quantity
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* const spU = reinterpret_cast<double*>(smem_raw);           // upper f64 (+ monthly)
  double* const spL = spU + (size_t)(kUpperF64 + kUpperMonStaged) * H;  // lower f64
  double* const spM = spL + (size_t)kLowerF64 * H;                      // ColumnMid f64
  double* const spS = spM + (size_t)kMidSlots * kMidF64 * H;            // lower state
  int32_t* const siU = reinterpret_cast<int32_t*>(spS + (size_t)kLowerStateStaged * H);
  int32_t* const siL = siU + (size_t)kUpperI32 * H;
  int32_t* const siM = siL + (size_t)kLowerI32 * H;
  double* const sF = reinterpret_cast<double*>(smem_raw + kColumnParamBytes);  // [kForcRows][H]
  const uint32_t mbar0 = smem_u32(smem_raw + kColumnParamBytes + kColumnForcBytes);
  // {doy, dowy, month, dom} of today and tomorrow, and the lower half's flag bytes
  int* const sCal = reinterpret_cast<int*>(smem_raw + kColumnParamBytes + kColumnForcBytes + 16);
  [[maybe_unused]] unsigned char* const sFlag =
      smem_raw + kColumnParamBytes + kColumnForcBytes + 16 + 32;

  GpuSink<kMode, H> out{a, nullptr, nullptr, nullptr, (unsigned)(a.ostride * 8),
                     (unsigned)(a.ostride * 4), 0u, 0, live, upper ? BAR_UPPER : BAR_LOWER};
  // bytes from one day to the next, and this thread's element of the first day
  const int64_t day_units = kMode == SINK_TILED ? (int64_t)a.ntiles * H : a.ostride;
  const int64_t day_f64 = (int64_t)a.n_f64 * day_units * 8;
  const int64_t day_i32 = (int64_t)a.n_i32 * day_units * 4;
  if constexpr (kMode == SINK_TILED) {
    out.pf = reinterpret_cast<char*>(a.out_f64 + tile * a.n_f64 * H + r);
    out.pi = reinterpret_cast<char*>(a.out_i32 + tile * a.n_i32 * H + r);
  } else {
    out.pf = reinterpret_cast<char*>(a.out_f64 + ii);
    out.pi = reinterpret_cast<char*>(a.out_i32 + ii);
  }
  auto reduce_violations = [&](int d) {
    if (a.viol != nullptr && __any_sync(0xffffffffu, out.violmask != 0u)) {
      for (int p = 0; p < BUD_CHANNEL; ++p) {
        const unsigned m = __ballot_sync(0xffffffffu, (out.violmask >> p) & 1u);
        if (m != 0u && (threadIdx.x & 31) == 0) atomicAdd(a.viol + d * BUD_N + p, __popc(m));
      }
    }
  };
  HruState s;
  // Q: the parameter view column_upper / column_lower see, indexed by r
  HruParams Q = a.P;
  Q.stride = 0;  // monthly rows: only the current month is staged

  if (upper) {
    if constexpr (!kWide) {
#if PWS_COLUMN_REGS_UPPER < PWS_COLUMN_REGS_LOWER
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PWS_COLUMN_REGS_UPPER));
#elif PWS_COLUMN_REGS_UPPER > PWS_COLUMN_REGS_LOWER
      asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PWS_COLUMN_REGS_UPPER));
#endif
    }
#define PWS_LOADST_F(name) s.name = a.state[(int64_t)ST_##name * S + ii];
#define PWS_LOADST_I(name) s.name = (int)a.state[(int64_t)ST_##name * S + ii];
#define X(name, kind) PWS_LOADST_##kind(name)
    PWS_HRU_STATE_UPPER(X)
#undef X
    {
      int k = 0;
#define X(n) spU[k * H + r] = __ldg(a.P.n + ii); Q.n = spU + k * H; ++k;
      PWS_UPPER_F64(X)
#undef X
#if PWS_COLUMN_MONTHLY_SMEM
#define X(n) Q.n = spU + k * H; ++k;
      PWS_UPPER_MON_F64(X)
#undef X
#else
      // monthly rows straight from global memory: row (month - 1) * S, element
      // hru0 + r (the slabs are padded so lanes past the last HRU stay inside)
      Q.stride = S;
#define X(n) Q.n = a.P.n + tile * H;
      PWS_UPPER_MON_F64(X)
      PWS_UPPER_MON_I32(X)
#undef X
#endif
      k = 0;
#define X(n) siU[k * H + r] = __ldg(a.P.n + ii); Q.n = siU + k * H; ++k;
      PWS_UPPER_I32(X)
#undef X
#if PWS_COLUMN_MONTHLY_SMEM
#define X(n) Q.n = siU + k * H; ++k;
      PWS_UPPER_MON_I32(X)
#undef X
#endif
    }
#if PWS_COLUMN_MONTHLY_SMEM
    auto stage_month = [&](int month) {
      const int64_t mo = (int64_t)(month - 1) * S + ii;
      int k = kUpperF64;
#define X(n) spU[k * H + r] = __ldg(a.P.n + mo); ++k;
      PWS_UPPER_MON_F64(X)
#undef X
      k = kUpperI32 - 1;
#define X(n) siU[k * H + r] = __ldg(a.P.n + mo);
      PWS_UPPER_MON_I32(X)
#undef X
    };
#endif
    // Forcings of day d+1 are in flight while day d computes.  When the rows
    // are 16-byte aligned (a.use_tma) one thread streams the five rows of the
    // CTA's HRUs into shared memory with TMA bulk copies that complete on an
    // mbarrier; otherwise every thread prefetches its own values
    // into registers.
    const bool tma = a.use_tma != 0;
    const int64_t hru0 = tile * H;
    // an odd tail is copied with the padding element behind it (see fill_column_args)
    const int64_t row_n = (a.P.nhru - hru0) < H ? (a.P.nhru - hru0) : H;
    const uint32_t row_bytes = (uint32_t)(((row_n + 1) & ~(int64_t)1) * 8);
    // forcing column of the CTA's first HRU; with a forcing period the CTA's row
    // may wrap around the end of the period once (fill_column_args only allows
    // TMA for an even period >= H, so both pieces stay 16-byte aligned)
    [[maybe_unused]] uint32_t fcol0 = 0, head_bytes = row_bytes;
    if constexpr (kPeriod) {
      fcol0 = (a.hru_base + (uint32_t)hru0) % a.fperiod;
      if (fcol0 + (uint32_t)row_n > a.fperiod) head_bytes = (a.fperiod - fcol0) * 8u;
    }
    auto tma_issue = [&](int d) {  // one thread
      const uint32_t bar = mbar0;
      const uint32_t dst = smem_u32(sF);
      const int doy = a.cal[d].x;
      const int64_t so = (int64_t)(doy - 1) * S + hru0;
      mbar_expect_tx(bar, kForcRows * row_bytes + 16u);
      tma_bulk_g2s(smem_u32(sCal + 4 * (d & 1)), a.cal + d, 16u, bar);
      if constexpr (!kPeriod) {
        const int64_t fo = (int64_t)d * a.fstride + hru0;
        tma_bulk_g2s(dst + 0 * H * 8, a.prcp + fo, row_bytes, bar);
        tma_bulk_g2s(dst + 1 * H * 8, a.tmax + fo, row_bytes, bar);
        tma_bulk_g2s(dst + 2 * H * 8, a.tmin + fo, row_bytes, bar);
      } else {
        const int64_t fo = (int64_t)d * a.fstride + fcol0;
        tma_bulk_g2s(dst + 0 * H * 8, a.prcp + fo, head_bytes, bar);
        tma_bulk_g2s(dst + 1 * H * 8, a.tmax + fo, head_bytes, bar);
        tma_bulk_g2s(dst + 2 * H * 8, a.tmin + fo, head_bytes, bar);
        if (head_bytes != row_bytes) {  // the rest of the row starts the period again
          const int64_t f0 = (int64_t)d * a.fstride;
          tma_bulk_g2s(dst + 0 * H * 8 + head_bytes, a.prcp + f0, row_bytes - head_bytes, bar);
          tma_bulk_g2s(dst + 1 * H * 8 + head_bytes, a.tmax + f0, row_bytes - head_bytes, bar);
          tma_bulk_g2s(dst + 2 * H * 8 + head_bytes, a.tmin + f0, row_bytes - head_bytes, bar);
        }
      }
      tma_bulk_g2s(dst + 3 * H * 8, a.P.soltab_potsw + so, row_bytes, bar);
      tma_bulk_g2s(dst + 4 * H * 8, a.P.soltab_horad_potsw + so, row_bytes, bar);
    };
    DayIn nxt;
    if (tma) {
      if (r == 0) {
        mbar_init(mbar0, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
      for (int k = 0; k < kForcRows; ++k) sF[k * H + r] = 0.0;  // lanes past the last HRU
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      bar_sync(BAR_UPPER, H);
      if (r == 0) tma_issue(0);
    } else {
      const int4 c = a.cal[0];
      nxt.doy = c.x; nxt.dowy = c.y; nxt.month = c.z; nxt.dom = c.w;
      int64_t fcol = ii;
      if constexpr (kPeriod) fcol = (a.hru_base + (uint32_t)ii) % a.fperiod;
      nxt.prcp = __ldcs(a.prcp + fcol);
      nxt.tmax = __ldcs(a.tmax + fcol);
      nxt.tmin = __ldcs(a.tmin + fcol);
      nxt.potsw = __ldg(a.P.soltab_potsw + (int64_t)(c.x - 1) * S + ii);
      nxt.horad = __ldg(a.P.soltab_horad_potsw + (int64_t)(c.x - 1) * S + ii);
    }
    for (int d = 0; d < a.ndays; ++d) {
      DayIn in;
      if (tma) {
        mbar_wait(mbar0, (uint32_t)(d & 1));
        in.cal = sCal + 4 * (d & 1);  // stays valid all day: tomorrow's row goes to the other entry
        in.doy = in.cal[0]; in.dowy = in.cal[1]; in.month = in.cal[2]; in.dom = in.cal[3];
        const double* f = sF + r;
        in.prcp = f[0 * H];
        in.tmax = f[1 * H];
        in.tmin = f[2 * H];
        in.potsw = f[3 * H];
        in.horad = f[4 * H];
        // every thread of the group holds its day-d values in registers: the
        // slot can take day d+1 while this day computes
        bar_sync(BAR_UPPER, H);
        if (r == 0 && d + 1 < a.ndays) tma_issue(d + 1);
      } else {
        in = nxt;
        in.cal = reinterpret_cast<const int*>(a.cal + d);
        if (d + 1 < a.ndays) {
          const int4 c = a.cal[d + 1];
          int64_t fo = (int64_t)(d + 1) * a.fstride;
          if constexpr (kPeriod) fo += (a.hru_base + (uint32_t)ii) % a.fperiod;
          else fo += ii;
          nxt.doy = c.x; nxt.dowy = c.y; nxt.month = c.z; nxt.dom = c.w;
          nxt.prcp = __ldcs(a.prcp + fo);
          nxt.tmax = __ldcs(a.tmax + fo);
          nxt.tmin = __ldcs(a.tmin + fo);
          nxt.potsw = __ldg(a.P.soltab_potsw + (int64_t)(c.x - 1) * S + ii);
          nxt.horad = __ldg(a.P.soltab_horad_potsw + (int64_t)(c.x - 1) * S + ii);
        }
      }
#if PWS_COLUMN_MONTHLY_SMEM
      // a new month starts on day-of-month 1 (and the block may start anywhere); uniform
      if (d == 0 || PWS_LDV(in.cal + 3) == 1) stage_month(in.month);
#endif
      out.violmask = 0u;
      ColumnMid mid;
      column_upper(Q, (int64_t)r, s, in, out, mid);
      // hand the day over to the lower half of this HRU (slot d mod kMidSlots)
      const int slot = d % kMidSlots;
      double* const mf = spM + (size_t)slot * kMidF64 * H + r;
      int32_t* const mi32 = siM + (size_t)slot * kMidI32 * H + r;
      if (d >= kMidSlots) bar_sync(BAR_MID_EMPTY + slot, 2 * H);  // day d - kMidSlots has been read
      mf[0 * H] = mid.potet;
      mf[1 * H] = mid.net_rain;
      mf[2 * H] = mid.net_snow;
      mf[3 * H] = mid.hru_intcpevap;
      mf[4 * H] = mid.intcp_changeover;
      mf[5 * H] = mid.snowmelt;
      mf[6 * H] = mid.snow_evap;
      mf[7 * H] = mid.pkwater_equiv;
      mf[8 * H] = mid.snowcov_area;
      mf[9 * H] = mid.through_rain;
      mi32[0 * H] = mid.transp_on;
      mi32[1 * H] = mid.pptmix_nopack;
      __threadfence_block();
      bar_arrive(BAR_MID_FULL + slot, 2 * H);
      out.pf += day_f64;
      out.pi += day_i32;
      reduce_violations(d);
    }
    if (live) {
#define PWS_STOREST(name) a.state[(int64_t)ST_##name * S + i] = (double)s.name;
#define X(name, kind) PWS_STOREST(name)
      PWS_HRU_STATE_UPPER(X)
#undef X
    }
  } else {
    if constexpr (!kWide) {
#if PWS_COLUMN_REGS_UPPER < PWS_COLUMN_REGS_LOWER
      asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PWS_COLUMN_REGS_LOWER));
#elif PWS_COLUMN_REGS_UPPER > PWS_COLUMN_REGS_LOWER
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PWS_COLUMN_REGS_LOWER));
#endif
    }
#if PWS_COLUMN_LOWER_STATE_SMEM == 2
    int ks = 0;
#define X(name) double& sm_##name = spS[(ks++) * H + r]; sm_##name = a.state[(int64_t)ST_##name * S + ii];
    PWS_LOWER_STATE_IN_SMEM(X)
#undef X
    LowerStateMixed sl{
#define X(name) sm_##name,
        PWS_LOWER_STATE_IN_SMEM(X)
#undef X
#define X(name) a.state[(int64_t)ST_##name * S + ii],
        PWS_LOWER_STATE_IN_REGS(X)
#undef X
    };
#elif PWS_COLUMN_LOWER_STATE_SMEM
    // the lower half's state lives in this thread's shared-memory column
    int ks = 0;
#define X(name, kind) double& sm_##name = spS[(ks++) * H + r]; sm_##name = a.state[(int64_t)ST_##name * S + ii];
    PWS_HRU_STATE_LOWER(X)
#undef X
    LowerStateRef sl{
#define X(name, kind) sm_##name,
        PWS_HRU_STATE_LOWER(X)
#undef X
    };
#else
#define X(name, kind) PWS_LOADST_##kind(name)
    PWS_HRU_STATE_LOWER(X)
#undef X
    HruState& sl = s;
#endif
    {
      int k = 0;
#define X(n) spL[k * H + r] = __ldg(a.P.n + ii); Q.n = spL + k * H; ++k;
      PWS_LOWER_F64(X)
#undef X
      k = 0;
#define X(n) siL[k * H + r] = __ldg(a.P.n + ii); Q.n = siL + k * H; ++k;
      PWS_LOWER_I32(X)
#undef X
    }
    for (int d = 0; d < a.ndays; ++d) {
      ColumnMid mid;
      const int slot = d % kMidSlots;
      const double* const mf = spM + (size_t)slot * kMidF64 * H + r;
      const int32_t* const mi32 = siM + (size_t)slot * kMidI32 * H + r;
      bar_sync(BAR_MID_FULL + slot, 2 * H);  // the upper group has written day d
      mid.potet = mf[0 * H];
      mid.net_rain = mf[1 * H];
      mid.net_snow = mf[2 * H];
      mid.hru_intcpevap = mf[3 * H];
      mid.intcp_changeover = mf[4 * H];
      mid.snowmelt = mf[5 * H];
      mid.snow_evap = mf[6 * H];
      mid.pkwater_equiv = mf[7 * H];
      mid.snowcov_area = mf[8 * H];
      mid.through_rain = mf[9 * H];
      mid.transp_on = mi32[0 * H];
      mid.pptmix_nopack = mi32[1 * H];
#if PWS_REREAD_MID
      sFlag[r] = (unsigned char)((mid.transp_on & 1) | ((mid.pptmix_nopack & 1) << 1));
      mid.flags = sFlag + r;
#endif
      if (d + kMidSlots < a.ndays) bar_arrive(BAR_MID_EMPTY + slot, 2 * H);
      out.violmask = 0u;
      out.plat = a.hru_lat ? a.hru_lat + (int64_t)d * S + ii : nullptr;
      column_lower(Q, (int64_t)r, sl, mid, out);
      out.pf += day_f64;
      out.pi += day_i32;
      reduce_violations(d);
    }
    if (live) {
#define X(name, kind) a.state[(int64_t)ST_##name * S + i] = (double)sl.name;
      PWS_HRU_STATE_LOWER(X)
#undef X
    }
  }
  if (live && out.errmask) atomicOr(a.err, out.errmask);
}

// ---------------------------------------------------------------- sub-chains
// Any subset of the column's processes, the others replaced by caller-supplied
// inputs (PWS_OV in hru_column.cuh): one thread per HRU runs both halves of
// the column in sequence with an overriding sink.  This is the compatibility
// path for single processes and partial chains (the reference's
// autotest/test_self_drive.py pattern); the full chain runs k_hru_column.
struct OvArgs {
  ColumnArgs c;
  const double* ov[PWS_N_OV];  // [ndays][ovstride] per overridden input, else null
  int64_t ovstride;
  unsigned proc_mask;          // bit BUD_*: the process is part of the model
};

struct OvSink {
  static constexpr bool kReread = false;
  const OvArgs& a;
  char* pf;
  char* pi;
  double* plat;
  int64_t idx;  // day * ovstride + hru
  unsigned violmask;
  int errmask;
  bool live;
  template <int var>
  __device__ __forceinline__ void f(double v) {
    const int64_t o = a.c.off[var];
    if (o >= 0 && live) __stcs(reinterpret_cast<double*>(pf + o), v);
  }
  template <int var>
  __device__ __forceinline__ void i(int v) {
    const int64_t o = a.c.off[var];
    if (o >= 0 && live) __stcs(reinterpret_cast<int*>(pi + o), v);
  }
  __device__ __forceinline__ void viol(int proc, bool open) {
    if (open && live && ((a.proc_mask >> proc) & 1u)) violmask |= 1u << proc;
  }
  __device__ __forceinline__ void lateral(double v) {
    if (plat && live) *plat = v;
  }
  __device__ __forceinline__ void error(int code) {
    if (live && ((a.proc_mask >> BUD_SOILZONE) & 1u)) errmask |= code;
  }
  template <int id>
  __device__ __forceinline__ double ov(double v) const {
    const double* p = a.ov[id];
    return p != nullptr ? __ldg(p + idx) : v;
  }
  template <int id>
  __device__ __forceinline__ int ovi(int v) const {
    const double* p = a.ov[id];
    return p != nullptr ? (int)__ldg(p + idx) : v;
  }
  __device__ __forceinline__ void sync() {}
  __device__ __forceinline__ void sync2() {}
};

__global__ void __launch_bounds__(128)
k_hru_column_ov(const __grid_constant__ OvArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < a.c.P.nhru;
  const int64_t ii = live ? i : a.c.P.nhru - 1;
  const int64_t S = a.c.P.stride;
  HruState s;
#define X(name, kind) s.name = (decltype(s.name))a.c.state[(int64_t)ST_##name * S + ii];
  PWS_HRU_STATE(X)
#undef X
  OvSink out{a, reinterpret_cast<char*>(a.c.out_f64 + ii), reinterpret_cast<char*>(a.c.out_i32 + ii),
             nullptr, 0, 0u, 0, live};
  const int64_t day_f64 = (int64_t)a.c.n_f64 * a.c.ostride * 8;
  const int64_t day_i32 = (int64_t)a.c.n_i32 * a.c.ostride * 4;
  for (int d = 0; d < a.c.ndays; ++d) {
    const int4 c = a.c.cal[d];
    DayIn in;
    in.doy = c.x; in.dowy = c.y; in.month = c.z; in.dom = c.w;
    const int64_t fo = (int64_t)d * a.c.fstride + (a.c.fperiod > 0 ? (uint32_t)ii % a.c.fperiod : (uint32_t)ii);
    in.prcp = a.c.prcp[fo];
    in.tmax = a.c.tmax[fo];
    in.tmin = a.c.tmin[fo];
    in.potsw = a.c.P.soltab_potsw[(int64_t)(c.x - 1) * S + ii];
    in.horad = a.c.P.soltab_horad_potsw[(int64_t)(c.x - 1) * S + ii];
    out.idx = (int64_t)d * a.ovstride + ii;
    out.violmask = 0u;
    out.plat = a.c.hru_lat ? a.c.hru_lat + (int64_t)d * S + ii : nullptr;
    ColumnMid mid;
    column_upper(a.c.P, ii, s, in, out, mid);
    column_lower(a.c.P, ii, s, mid, out);
    out.pf += day_f64;
    out.pi += day_i32;
    if (a.c.viol != nullptr && __any_sync(0xffffffffu, out.violmask != 0u)) {
      for (int p = 0; p < BUD_CHANNEL; ++p) {
        const unsigned m = __ballot_sync(0xffffffffu, (out.violmask >> p) & 1u);
        if (m != 0u && (threadIdx.x & 31) == 0) atomicAdd(a.c.viol + d * BUD_N + p, __popc(m));
      }
    }
  }
  if (live) {
#define X(name, kind) a.c.state[(int64_t)ST_##name * S + i] = (double)s.name;
    PWS_HRU_STATE(X)
#undef X
    if (out.errmask) atomicOr(a.c.err, out.errmask);
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
//   * At step t, segment i advances the G = PWS_ROUTE_HOURS hours
//     G (t - delay(i)) ...  The outflows it publishes go to a double-buffered
//     shared-memory slot; its downstream segment reads them in the next step.
//     One __syncthreads() per step is the only synchronisation; a block of T
//     days takes 24 T / G + dmax steps.
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
  // consecutive launches on two streams (pws_b200.cu: route_overlap): a chunk stores
  // `epoch` here after its state; with wait_prev it first waits for epoch - 1
  int* chunk_done;           // [nchunks]
  int epoch;
  int wait_prev;
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
  // the same, resolved by the host into one pointer per result of a day (row of the
  // variable in day 0, or null when it is not selected) and the day stride in
  // elements: channel_outflow_vol | outflows, seg_lateral_inflow, seg_upstream_inflow |
  // node_upstream_inflows, seg_outflow | node_outflows, seg_stor_change |
  // node_storage_changes, node_sink_source
  double* out_ptr[6];
  int64_t out_day_stride;
  int* err;
  short slot[PWS_N_SEG_VARS];
  // FlowGraph runs (k_route_chunks<G, true>; base/flow_graph.py): nodes are segments
  const int32_t* kind;       // [nseg] pos order: NODE_MUSKINGUM / NODE_PASS / NODE_OBSIN
  const int32_t* obs_col;    // [nseg] pos order: column of `obs` of an ObsIn / SourceSink node, else -1
  const double* obs;         // [ndays][n_obs] observed flows of the ObsIn nodes, requests of the SourceSink nodes
  const double* flow_min;    // [nseg] pos order: minimum flow of a SourceSink node
  int n_obs;
  double* fg_out;            // [ndays][FG_N][nseg], original node order
};

// FlowNode kinds (base/flow_graph.py:16-99): PRMSChannelFlowNode
// (prms_channel_flow_graph.py:30-160), PassThroughFlowNode (pass_through_flow_node.py,
// routed as a Muskingum node with K < 1 h: outflow = inflow every hour, the daily
// values are the same sums), ObsInFlowNode (obsin_flow_node.py), SourceSinkFlowNode
// (source_sink_flow_node.py)
enum { NODE_MUSKINGUM = 0, NODE_PASS = 1, NODE_OBSIN = 2, NODE_SOURCESINK = 3 };
// rows of fg_out
enum { FG_outflows = 0, FG_node_upstream_inflows, FG_node_outflows, FG_node_storage_changes,
       FG_node_sink_source, FG_N };

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
// Hours a segment advances per systolic step (G, divides 24): one barrier, one
// shared-memory round trip and one set of loop overheads per G hourly updates,
// but the pipeline fill is dmax steps of G hours each.  Per launch of T days a
// chunk takes about (24 T / G + dmax) (0.39 + 0.2 G) us (fit to G = 4 .. 12 on the
// CONUS domain, where G = 12 routes a 10-year pass in 24.2 ms against 28.4 for
// G = 4), so shallow networks get G = 12, deep ones G = 4 and very deep ones G = 2 (pws_init
// picks by the number of levels of the segment graph; option "route_hours").
#ifndef PWS_ROUTE_HOURS
#define PWS_ROUTE_HOURS 4    // deep networks
#endif
#ifndef PWS_ROUTE_HOURS_SHALLOW
#define PWS_ROUTE_HOURS_SHALLOW 12
#endif
// Very deep graphs (thousands of levels) in launches of a few hundred days are all
// fill: G = 2 routes the chained configs[4] network (8,856 levels) at 2.06e9
// segment-days/s in 256-day launches against 1.55e9 for G = 4 and 1.80e9 for G = 3,
// and at 3.78e9 against 4.06e9 when the ten years are one launch
// (profiles/r02_route_hours_experiment.log); by the fit above G = 2 wins while
// levels > ~6 x days per launch.
#ifndef PWS_ROUTE_HOURS_DEEP
#define PWS_ROUTE_HOURS_DEEP 2
#endif
#ifndef PWS_ROUTE_DEEP_LEVELS
#define PWS_ROUTE_DEEP_LEVELS 2000
#endif
// lateral-inflow ring: kLatDepth days x 3 addends per segment, filled by cp.async
constexpr int kLatDepth = 3;
constexpr size_t route_handover_bytes(int G) {
  return (size_t)2 * G * (PWS_CHUNK_THREADS + 1) * sizeof(double);
}
constexpr size_t route_smem_bytes(int G) {
  return route_handover_bytes(G) + (size_t)kLatDepth * 3 * PWS_CHUNK_THREADS * sizeof(double);
}

// kGraph: the segments are the nodes of a FlowGraph (base/flow_graph.py:576-657):
// node states start from zero every day (FlowNode.prepare_timestep), ObsIn nodes
// put their observation out instead of the routed flow and account the difference
// as sink / source, and the results are FlowGraph's variables in flow units
// (a.fg_out) instead of PRMSChannel's.
template <int G, bool kGraph = false>
__global__ void __launch_bounds__(PWS_CHUNK_THREADS, 1)
k_route_chunks(const __grid_constant__ ChannelArgs a) {
  constexpr int kZero = PWS_CHUNK_THREADS;  // a slot that always holds 0.0
  constexpr int kStepsPerDay = 24 / G;
  static_assert(24 % G == 0, "hours per step must divide 24");
  // double-buffered hand-over of the G hourly outflows of every segment (dynamic
  // shared memory: 2 G (threads + 1) doubles, more than the 48 KB static limit for G > 5)
  extern __shared__ __align__(16) unsigned char route_smem[];
  double(*const s_out)[G][PWS_CHUNK_THREADS + 1] =
      reinterpret_cast<double(*)[G][PWS_CHUNK_THREADS + 1]>(route_smem);
  __shared__ int s_chunk;
  const int tid = threadIdx.x;
  if (tid == 0) {
    s_chunk = atomicAdd(a.ticket, 1);
    for (int g = 0; g < G; ++g) s_out[0][g][kZero] = s_out[1][g][kZero] = 0.0;
  }
  __syncthreads();
  const int chunk = s_chunk;
  if (chunk >= a.nchunks) return;
  if (a.wait_prev) {
    // the previous launch may still be routing this chunk: its state (outflow_ts,
    // seg_inflow) is ours to continue from once it says so
    if (tid == 0) {
      int idle = 0;
      while (*reinterpret_cast<volatile const int*>(a.chunk_done + chunk) < a.epoch - 1) {
        if (++idle > PWS_ROUTE_MAX_IDLE) {
          atomicOr(a.err, 16);
          break;
        }
        __nanosleep(256);
      }
      __threadfence();
    }
    __syncthreads();
  }
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
  // up to three upstream segments sit in registers: a local one as its slot of the
  // hand-over buffer, one in another chunk as its row of `ots` (x >= 0).  A segment
  // below a chunk boundary used to walk the CSR with three dependent global loads per
  // hour, and the whole chunk waited for it at every step's barrier: that, not the
  // Muskingum recurrence, set the pace of a deep chained network.
  // (an external one is kept as -(row + 1) in the same register as a local slot).
  // Only in the G <= 6 kernels, i.e. for deep networks: in the 12-hour kernel the
  // extra instantiation of the hour loop costs registers (8 bytes of spills).
  constexpr bool kRegXup = G <= 6;
  const bool fast_up = nup <= 3 && (kRegXup || (flags & SEG_UPS_LOCAL));
  if (live && fast_up) {
    auto slot = [&](int k) {
      const int s = a.up_idx[ub + k];
      return (s >= cb && s < ce) ? s - cb : -(a.ext_row[s] + 1);
    };
    if (nup > 0) u0 = slot(0);
    if (nup > 1) u1 = slot(1);
    if (nup > 2) u2 = slot(2);
  }
  const bool xup = kRegXup && live && !(flags & SEG_UPS_LOCAL);  // an upstream segment in another chunk
  const bool fast_lat = a.seg_lat_in != nullptr || nh <= 3;
  if (live && a.seg_lat_in == nullptr && nh <= 3) {
    if (nh > 0) h0 = a.hru_idx[hb];
    if (nh > 1) h1 = a.hru_idx[hb + 1];
    if (nh > 2) h2 = a.hru_idx[hb + 2];
  }
  // No special cases in the hour step: with ts == 1 the reciprocal quotient
  // returns x exactly (q = x, remainder 0), and a pass-through segment
  // (tsi < 0, K < 1 h: outflow = inflow, prms_channel.py:529-547) is the
  // recurrence with c0 = 1, c1 = c2 = 0, exact for finite flows.
  if (tsi < 0) {
    c0 = 1.0;
    c1 = 0.0;
    c2 = 0.0;
  }
  const bool down = (flags & SEG_HAS_DOWN) != 0;
  [[maybe_unused]] int kind = NODE_MUSKINGUM, obs_col = -1;
  [[maybe_unused]] double obs_q = 0.0, ss_sum = 0.0, st_sink = 0.0, flow_min = 0.0;
  if constexpr (kGraph) {
    if (live) {
      kind = a.kind[pos];
      obs_col = a.obs_col[pos];
      flow_min = a.flow_min[pos];
    }
  }

  // Lateral inflow of day k as up to three raw addends (summed a window later,
  // so the loads stay in flight): HRU contributions in ascending HRU order
  // (prms_channel.py:396-421).
  // Lateral inflow: days are fetched in order, kLatDepth - 1 windows ahead, with
  // cp.async into a shared-memory ring (slot = window mod kLatDepth, three raw
  // addends per segment, summed when the day starts).  Asynchronous copies are
  // waited for with cp.async.wait_group, not through a register scoreboard: with
  // plain loads the compiler put the scoreboard wait at the top of the step loop,
  // which exposed one HBM latency per simulated day (13 % of the kernel in ncu's
  // source view).  The row offset of the day is a running sum.
  constexpr size_t kHandoverBytes = (size_t)2 * G * (PWS_CHUNK_THREADS + 1) * sizeof(double);
  static_assert(kHandoverBytes % 16 == 0, "lateral ring must stay 8-byte aligned");
  double(*const s_lat)[3][PWS_CHUNK_THREADS] =
      reinterpret_cast<double(*)[3][PWS_CHUNK_THREADS]>(route_smem + kHandoverBytes);
  int64_t lat_off = 0;  // row offset of the day fetch_lateral is called for
  auto cp_async8 = [](double* dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
  };
  auto fetch_lateral = [&](int k, int slot) {
    const int64_t off = lat_off;
    lat_off += a.seg_lat_in != nullptr ? a.nseg : a.hstride;
    double* const q0 = &s_lat[slot][0][tid];
    double* const q1 = &s_lat[slot][1][tid];
    double* const q2 = &s_lat[slot][2][tid];
    if (!live || k < 0 || k >= T) {
      *q0 = 0.0; *q1 = 0.0; *q2 = 0.0;
    } else if (a.seg_lat_in != nullptr) {
      cp_async8(q0, a.seg_lat_in + off + orig);
      *q1 = 0.0; *q2 = 0.0;
    } else if (fast_lat) {
      const double* row = a.hru_lat + off;
      if (h0 >= 0) cp_async8(q0, row + h0); else *q0 = 0.0;
      if (h1 >= 0) cp_async8(q1, row + h1); else *q1 = 0.0;
      if (h2 >= 0) cp_async8(q2, row + h2); else *q2 = 0.0;
    } else {
      double lat = 0.0;
      for (int j = hb; j < he; ++j)
        lat += __ldcg(a.hru_lat + off + a.hru_idx[j]);
      *q0 = lat; *q1 = 0.0; *q2 = 0.0;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // Results of the day a segment finished during the current 24-step window;
  // written to HBM by all threads together at the next window start, so the
  // stores are off the per-step critical path.
  double st_outvol = 0.0, st_dstor = 0.0, st_lat = 0.0, st_up = 0.0, st_out = 0.0;
  int st_day = -1;
  int64_t scr_off = pos;   // st_day * sstride + pos (a segment flushes its days in order)
  int64_t out_off = orig;  // st_day * out_day_stride + orig
  auto flush_day = [&]() {
    if (st_day < 0) return;
    if constexpr (!kGraph) {
      a.day_outvol[scr_off] = st_outvol;
      a.day_dstor[scr_off] = st_dstor;
      a.day_lat[scr_off] = st_lat;
      scr_off += a.sstride;
    }
    if (a.out_ptr[0] != nullptr) a.out_ptr[0][out_off] = st_outvol;
    if (a.out_ptr[1] != nullptr) a.out_ptr[1][out_off] = st_lat;
    if (a.out_ptr[2] != nullptr) a.out_ptr[2][out_off] = st_up;
    if (a.out_ptr[3] != nullptr) a.out_ptr[3][out_off] = st_out;
    if (a.out_ptr[4] != nullptr) a.out_ptr[4][out_off] = st_dstor;
    if constexpr (kGraph) {
      if (a.out_ptr[5] != nullptr) a.out_ptr[5][out_off] = st_sink;
    }
    out_off += a.out_day_stride;
    st_day = -1;
  };

  // ---- prognostic state
  double outflow_ts = 0.0, seg_inflow_prev = 0.0;
  if (live) {
    // (L2 loads: with overlapped launches the previous launch stored these while this
    // kernel was already running, and L1 is not coherent)
    outflow_ts = __ldcg(a.outflow_ts + pos);
    seg_inflow_prev = __ldcg(a.seg_inflow + pos);
  }
  // A window = the kStepsPerDay steps of one day.  The start of day k falls
  // into window k + delay / kStepsPerDay; at every window start the lateral
  // inflow for the day starting in the NEXT window is requested.
  const int dd = delay / kStepsPerDay;
  lat_off = -(int64_t)dd * (a.seg_lat_in != nullptr ? a.nseg : a.hstride);
  for (int j = 0; j < kLatDepth - 1; ++j) fetch_lateral(j - dd, j);
  double lat_window = 0.0;  // lateral inflow of the day that starts in the current window
  double lateral = 0.0, seg_inflow0 = 0.0, seg_inflow = 0.0, seg_outflow = 0.0;
  double inflow_ts = 0.0, cur_sum = 0.0;
  int d = 0, h = 0;
  const int nsteps = kStepsPerDay * T + dmax;
  for (int t = 0, tw = 0, w = 0; t < nsteps; ++t) {
    if (tw == 0) {  // window start (uniform across the CTA)
      flush_day();
      // the day starting in this window was requested kLatDepth - 1 windows ago
      asm volatile("cp.async.wait_group %0;" ::"n"(kLatDepth - 2) : "memory");
      const int sl = w % kLatDepth;
      lat_window = ((0.0 + s_lat[sl][0][tid]) + s_lat[sl][1][tid]) + s_lat[sl][2][tid];
      fetch_lateral(w + kLatDepth - 1 - dd, (w + kLatDepth - 1) % kLatDepth);
    }
    if (++tw == kStepsPerDay) {
      tw = 0;
      ++w;
    }
    const double(*rd)[PWS_CHUNK_THREADS + 1] = s_out[(t & 1) ^ 1];  // published in step t - 1
    double(*wr)[PWS_CHUNK_THREADS + 1] = s_out[t & 1];
    const int ts_local = t - delay;  // this segment's own step number
    if (live && ts_local >= 0 && ts_local < kStepsPerDay * T) {
      if (h == 0) {
        lateral = lat_window;
        seg_inflow0 = seg_inflow_prev;  // _advance_variables, prms_channel.py:384-386
        if constexpr (kGraph) {
          // PRMSChannelFlowNode.advance / prepare_timestep (prms_channel_flow_graph.py:76-82, 113-115)
          seg_inflow = 0.0;
          seg_outflow = 0.0;
          inflow_ts = 0.0;
          cur_sum = 0.0;
          ss_sum = 0.0;
          if (obs_col >= 0) obs_q = __ldcg(a.obs + (int64_t)d * a.n_obs + obs_col);  // obsin_flow_node.py:37-41
        } else {
          seg_inflow = seg_inflow0 * 0.0;
          seg_outflow = seg_inflow0 * 0.0;
          inflow_ts = seg_inflow0 * 0.0;
          cur_sum = seg_inflow0 * 0.0;
        }
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
      // The G hours of the step, as straight-line code: the per-segment cases
      // (upstream slots in registers or the general CSR walk, an outflow row for
      // another chunk or not) are decided ONCE per step and the hour loop is
      // instantiated per case, so nothing inside it branches and the compiler
      // schedules the shared-memory reads and the non-recurrent sums of all G
      // hours ahead of the outflow recurrence.
      const unsigned fbits = fire >> h;  // bit g: hour h + g is a routing update
      auto hours = [&](auto fast_c, auto ext_c, auto xup_c) {
        constexpr bool kFast = decltype(fast_c)::value, kExt = decltype(ext_c)::value;
        constexpr bool kXup = decltype(xup_c)::value;  // some upstream segment is in another chunk
        // the hand-over reads of kB hours are issued together (all G at once costs
        // 6 G registers and pushed the kernel's addresses out of the register file)
        constexpr int kB = G > 6 ? 6 : G;
#pragma unroll
        for (int g0 = 0; g0 < G; g0 += kB) {
          double upv[kB];
          if constexpr (kFast && !kXup) {
#pragma unroll
            for (int b = 0; b < kB; ++b)
              upv[b] = (rd[g0 + b][u0] + rd[g0 + b][u1]) + rd[g0 + b][u2];  // absent ones read the 0.0 slot
          } else if constexpr (kFast) {
            // the day's 24 values of an external upstream are complete (the progress
            // wait at h == 0), so the loads of a batch are independent and go out together
            const int64_t hh = (int64_t)d * 24 + h + g0;
#pragma unroll
            for (int b = 0; b < kB; ++b) {
              const double v0 = u0 < 0 ? __ldcg(a.ots + (int64_t)(-u0 - 1) * row_len + hh + b) : rd[g0 + b][u0];
              const double v1 = u1 < 0 ? __ldcg(a.ots + (int64_t)(-u1 - 1) * row_len + hh + b) : rd[g0 + b][u1];
              const double v2 = u2 < 0 ? __ldcg(a.ots + (int64_t)(-u2 - 1) * row_len + hh + b) : rd[g0 + b][u2];
              upv[b] = (v0 + v1) + v2;
            }
          } else {
#pragma unroll
            for (int b = 0; b < kB; ++b) {
              double up = 0.0;
              for (int k = ub; k < ue; ++k) {
                const int u = a.up_idx[k];
                if (u >= cb && u < ce) up += rd[g0 + b][u - cb];
                else up += __ldcg(a.ots + (int64_t)a.ext_row[u] * row_len + d * 24 + h + g0 + b);
              }
              upv[b] = up;
            }
          }
#pragma unroll
          for (int b = 0; b < kB; ++b) {
            const int g = g0 + b;
            const double up = upv[b];
            const double cur = lateral + up;
            seg_inflow += cur;
            cur_sum += up;
            const double acc = inflow_ts + cur;
            // routing update (branch-free; most segments update every hour)
            const bool upd = (fbits >> g) & 1u;
            const double xq = div_by_ts(acc, ts, rts);
            const double routed = (xq * c0 + seg_inflow0 * c1) + outflow_ts * c2;
            outflow_ts = upd ? routed : outflow_ts;
            seg_inflow0 = upd ? xq : seg_inflow0;
            inflow_ts = upd ? 0.0 : acc;
            if constexpr (kGraph) {
              if (kind == NODE_OBSIN) {  // obsin_flow_node.py:43-57
                if (obs_q >= 0.0) {
                  ss_sum += obs_q - cur;
                } else {
                  obs_q = cur;  // a missing observation: the first hour's inflow goes out all day
                  ss_sum += 0.0;
                }
                outflow_ts = obs_q;
              } else if (kind == NODE_SOURCESINK) {  // source_sink_flow_node.py:60-92
                double ss = obs_q, outq = cur;       // obs_q: the day's request, source > 0 > sink
                if (ss >= 0.0) {
                  outq = cur + ss;                   // a source is always applied
                } else if (ss < 0.0 && cur < flow_min) {
                  ss = 0.0;                          // no sink below the minimum flow
                } else if (ss < 0.0 && cur >= flow_min) {
                  if (cur + ss < flow_min) {
                    ss = flow_min - cur;             // the sink takes what the minimum flow leaves
                    outq = flow_min;
                  } else {
                    outq = cur + ss;
                  }
                }
                ss_sum += ss;
                outflow_ts = outq;
              }
            }
            seg_outflow += outflow_ts;
            wr[g][tid] = outflow_ts;
            if constexpr (kExt) {
              if (ext_row >= 0)
                __stcg(a.ots + (int64_t)ext_row * row_len + d * 24 + h + g, outflow_ts);
            }
          }
        }
      };
      if (kRegXup && fast_up && xup) hours(std::true_type{}, std::true_type{}, std::bool_constant<kRegXup>{});
      else if (fast_up && ext_row < 0) hours(std::true_type{}, std::false_type{}, std::false_type{});
      else if (fast_up) hours(std::true_type{}, std::true_type{}, std::false_type{});
      else hours(std::false_type{}, std::true_type{}, std::false_type{});
      h += G;
      if (h == 24) {
        seg_outflow = div_by_ts(seg_outflow, 24.0, 1.0 / 24.0);
        seg_inflow = div_by_ts(seg_inflow, 24.0, 1.0 / 24.0);
        st_up = div_by_ts(cur_sum, 24.0, 1.0 / 24.0);
        if constexpr (kGraph) {
          // flow units; an ObsIn node reports its observation and no storage change
          // (obsin_flow_node.py:65-90), flow_graph.py:629-650
          // a SourceSink node reports its LAST substep's outflow (:105-110)
          if (kind == NODE_OBSIN) seg_outflow = obs_q;
          if (kind == NODE_SOURCESINK) seg_outflow = outflow_ts;
          st_sink = kind >= NODE_OBSIN ? div_by_ts(ss_sum, 24.0, 1.0 / 24.0) : 0.0;
          st_dstor = kind >= NODE_OBSIN ? 0.0 : seg_inflow - seg_outflow;
          st_outvol = down ? 0.0 : seg_outflow;
        } else {
          st_dstor = (seg_inflow - seg_outflow) * 86400.0;
          st_outvol = (down ? 0.0 : seg_outflow) * 86400.0;
        }
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
      }
    }
    __syncthreads();
  }
  flush_day();
  if (live) {
    a.outflow_ts[pos] = outflow_ts;
    a.seg_inflow[pos] = seg_inflow_prev;
  }
  // release: every thread's state stores before the chunk's serial
  __threadfence();
  __syncthreads();
  if (tid == 0) *reinterpret_cast<volatile int*>(a.chunk_done + chunk) = a.epoch;
}

// Forcings uploaded as float32 (pws_run_host_f32: the storage type of the
// reference's prcp / tmax / tmin netCDF files) widened into the float64 ring
// k_hru_column reads: grid (ceil(n / 256), days, 3 variables).  float -> double
// is exact, so the column sees the numbers the reference gets from
// `ds[var][:].astype(float64)`.  src rows are n floats apart, dst rows `stride`
// doubles (the padded, TMA-aligned ring rows).
__global__ void __launch_bounds__(256)
k_widen_forcings(const float* __restrict__ src, double* __restrict__ dst, int64_t n, int64_t stride,
                 int64_t src_var_rows, int64_t dst_var_rows) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int64_t d = blockIdx.y, k = blockIdx.z;
  dst[(k * dst_var_rows + d) * stride + i] = (double)__ldg(src + (k * src_var_rows + d) * n + i);
}

// Budget accumulations (base/budget.py:262-269): `acc += value` for every day of
// a block, in day order, per unit; grid (ceil(n / 256), variables).  buf is the
// block's result buffer [ndays][rows][units]; slot = the variable's row.
struct AccSlots {
  static constexpr int kMax = 64;
  short s[kMax];
};
__global__ void __launch_bounds__(256)
k_accumulate(const AccSlots slots, const double* __restrict__ buf, int64_t day_stride, int64_t units,
             int ndays, int64_t n, double* __restrict__ acc, int64_t acc_stride) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int k = blockIdx.y;
  const double* p = buf + (int64_t)slots.s[k] * units + i;
  double a = acc[(int64_t)k * acc_stride + i];
  for (int d = 0; d < ndays; ++d) a += __ldcs(p + (int64_t)d * day_stride);
  acc[(int64_t)k * acc_stride + i] = a;
}

// PRMSEt (hydrology/prms_et.py:140-181): the bookkeeping process that sums the
// evaporative losses of the other processes.  Elementwise over [ndays][nhru], in
// the reference's order of operations (frac_perv :145, available_potet :160-181,
// hru_actet :157).
__global__ void __launch_bounds__(256)
k_prms_et(int64_t n_total, int64_t nhru, const double* __restrict__ potet,
          const double* __restrict__ hru_impervevap, const double* __restrict__ hru_intcpevap,
          const double* __restrict__ snow_evap, const double* __restrict__ dprst_evap_hru,
          const double* __restrict__ perv_actet, const double* __restrict__ hru_percent_imperv,
          const double* __restrict__ dprst_frac, double* __restrict__ unused_potet,
          double* __restrict__ hru_perv_actet, double* __restrict__ hru_actet) {
  for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < n_total; j += (int64_t)gridDim.x * 256) {
    const int64_t i = j % nhru;
    const double frac_perv = 1.0 - hru_percent_imperv[i] - dprst_frac[i];
    const double hpa = perv_actet[j] * frac_perv;
    double loss = hru_intcpevap[j];
    loss += snow_evap[j];
    loss += dprst_evap_hru[j];
    loss += hru_impervevap[j];
    loss += hpa;
    const double unused = fmax(potet[j] - loss, 0.0);
    hru_perv_actet[j] = hpa;
    unused_potet[j] = unused;
    hru_actet[j] = potet[j] - unused;
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
