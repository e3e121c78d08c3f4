// column_host.cpp -- TEST-ONLY host build of the device physics in
// pywatershed_b200/csrc/hru_column.cuh (the same source the CUDA kernels
// inline), so its logic can be checked against the oracle on a machine
// without a GPU.  Never part of libpwsb200.so; built into tests/_build/ by
// tests/test_column_host.py.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../pywatershed_b200/csrc/hru_column.cuh"

using namespace pws;

namespace {
struct Shim {
  int64_t nhru;
  int ndepl;
  std::map<std::string, std::vector<double>> pf;
  std::map<std::string, std::vector<int32_t>> pi;
  std::map<std::string, double> opt;
  std::vector<double> derived, state, soltab[3], snarea, allsnow_c;
  HruParams P{};
  HruDerivedMut D{};
};

struct HostSink {
  static constexpr bool kReread = false;
  double* out;  // [PWS_N_HRU_VARS][nhru] for the current day, at this HRU
  int64_t nhru;
  int* pv;
  int err = 0;
  template <int var> void f(double v) { out[(int64_t)var * nhru] = v; }
  template <int var> void i(int v) { out[(int64_t)var * nhru] = (double)v; }
  void viol(int proc, bool open) { if (open && pv) pv[proc]++; }
  void lateral(double) {}
  void sync() {}
  void sync2() {}
  template <int id> double ov(double v) const { return v; }
  template <int id> int ovi(int v) const { return v; }
  void error(int c) { err |= c; }
};
}  // namespace

extern "C" {
// a / b through the correctly rounded reciprocal (column_lower's PER_AREA)
void shim_div_r(const double* a, const double* b, int64_t n, double* out) {
  for (int64_t k = 0; k < n; ++k) out[k] = pws::div_r(a[k], b[k], 1.0 / b[k]);
}
// the zero-numerator guard of the runoff / soilzone divisions
void shim_div0(const double* a, const double* b, int64_t n, double* out) {
  for (int64_t k = 0; k < n; ++k) out[k] = pws::div0(a[k], b[k]);
}
// Muskingum sub-daily division, exposed for tests/test_column_host.py
void shim_div_by_ts(const double* x, int64_t n, double ts, double* out) {
  const double r = 1.0 / ts;
  for (int64_t k = 0; k < n; ++k) out[k] = pws::div_by_ts(x[k], ts, r);
}
void* shim_create(int64_t nhru, int ndepl) {
  Shim* s = new Shim();
  s->nhru = nhru;
  s->ndepl = ndepl;
  return s;
}
void shim_destroy(void* p) { delete (Shim*)p; }
void shim_set_f64(void* p, const char* name, const double* v, int64_t n) {
  ((Shim*)p)->pf[name].assign(v, v + n);
}
void shim_set_i32(void* p, const char* name, const int32_t* v, int64_t n) {
  ((Shim*)p)->pi[name].assign(v, v + n);
}
void shim_set_opt(void* p, const char* name, double v) { ((Shim*)p)->opt[name] = v; }

int shim_init(void* p) {
  Shim* s = (Shim*)p;
  const int64_t n = s->nhru;
  HruParams& P = s->P;
  P.nhru = n;
  P.stride = n;
  auto bc = [&](std::vector<double>& v, int64_t len) {
    if ((int64_t)v.size() == 1) v.assign(len, v[0]);
  };
#define X(nm) bc(s->pf[#nm], n); if ((int64_t)s->pf[#nm].size() != n) return -1; P.nm = s->pf[#nm].data();
  PWS_HRU_F64_PARAMS(X)
#undef X
#define X(nm) if ((int64_t)s->pf[#nm].size() != 12 * n) return -2; P.nm = s->pf[#nm].data();
  PWS_MON_F64_PARAMS(X)
#undef X
#define X(nm) if ((int64_t)s->pi[#nm].size() != n) return -3; P.nm = s->pi[#nm].data();
  PWS_HRU_I32_PARAMS(X)
#undef X
#define X(nm) if ((int64_t)s->pi[#nm].size() != 12 * n) return -4; P.nm = s->pi[#nm].data();
  PWS_MON_I32_PARAMS(X)
#undef X
  int nd = 0;
#define X(nm) ++nd;
  PWS_HRU_F64_DERIVED(X)
#undef X
  s->derived.assign((size_t)nd * n, 0.0);
  s->allsnow_c.assign((size_t)12 * n, 0.0);
  int k = 0;
#define X(nm) s->D.nm = s->derived.data() + (size_t)(k++) * n; P.nm = s->D.nm;
  PWS_HRU_F64_DERIVED(X)
#undef X
  s->D.tmax_allsnow_c = s->allsnow_c.data();
  P.tmax_allsnow_c = s->allsnow_c.data();
  s->snarea = s->pf["snarea_curve"];
  P.snarea_curve = s->snarea.data();
  P.albset_rna = s->opt["albset_rna"]; P.albset_rnm = s->opt["albset_rnm"];
  P.albset_sna = s->opt["albset_sna"]; P.albset_snm = s->opt["albset_snm"];
  P.dprst_flag = (int)s->opt["dprst_flag"];
  P.temp_units = (int)s->opt["temp_units"];
  for (auto& t : s->soltab) t.assign((size_t)366 * n, 0.0);
  for (int64_t i = 0; i < n; ++i) s->D.hru_cossl[i] = std::cos(std::atan(P.hru_slope[i]));
  for (int j = 0; j < 366; ++j) {
    double decl, r1;
    sol_constants(j, decl, r1);
    const double sd = std::sin(decl), cd = std::cos(decl), td = std::tan(decl);
    for (int64_t i = 0; i < n; ++i) {
      double solt, sunh;
      sol_day(sol_hru(0.0, 0.0, P.hru_lat[i]), sd, cd, td, r1, solt, sunh);
      s->soltab[1][(size_t)j * n + i] = solt;
      sol_day(sol_hru(P.hru_slope[i], P.hru_aspect[i], P.hru_lat[i]), sd, cd, td, r1, solt, sunh);
      s->soltab[0][(size_t)j * n + i] = solt;
      s->soltab[2][(size_t)j * n + i] = sunh;
    }
  }
  P.soltab_potsw = s->soltab[0].data();
  P.soltab_horad_potsw = s->soltab[1].data();
  s->state.assign((size_t)PWS_N_HRU_STATE * n, 0.0);
  int err = 0;
  for (int64_t i = 0; i < n; ++i)
    err |= hru_init_one(P, s->D, s->state.data(), i, (int)s->opt["start_month"],
                        (int)s->opt["start_doy"], 1);
  return err;
}

const double* shim_soltab(void* p, int which) { return ((Shim*)p)->soltab[which].data(); }

// out [ndays][PWS_N_HRU_VARS][nhru] float64, viol [ndays][6]
int shim_run(void* p, int ndays, const int32_t* cal, const double* prcp, const double* tmax,
             const double* tmin, double* out, int32_t* viol) {
  Shim* sh = (Shim*)p;
  const int64_t n = sh->nhru;
  int err = 0;
  for (int64_t i = 0; i < n; ++i) {  // thread = HRU, state "in registers" over the block
    HruState s;
#define LD_F(name) s.name = sh->state[(size_t)ST_##name * n + i];
#define LD_I(name) s.name = (int)sh->state[(size_t)ST_##name * n + i];
#define X(name, kind) LD_##kind(name)
    PWS_HRU_STATE(X)
#undef X
    for (int d = 0; d < ndays; ++d) {
      DayIn in;
      in.doy = cal[4 * d]; in.dowy = cal[4 * d + 1]; in.month = cal[4 * d + 2]; in.dom = cal[4 * d + 3];
      in.prcp = prcp[(size_t)d * n + i]; in.tmax = tmax[(size_t)d * n + i]; in.tmin = tmin[(size_t)d * n + i];
      in.potsw = sh->soltab[0][(size_t)(in.doy - 1) * n + i];
      in.horad = sh->soltab[1][(size_t)(in.doy - 1) * n + i];
      HostSink sink{out + (size_t)d * PWS_N_HRU_VARS * n + i, n, viol ? viol + 6 * d : nullptr};
      column_day(sh->P, i, s, in, sink);
      err |= sink.err;
    }
#define X(name, kind) sh->state[(size_t)ST_##name * n + i] = (double)s.name;
    PWS_HRU_STATE(X)
#undef X
  }
  return err;
}
const char* shim_var_name(int i) {
  static const char* nm[] = {
#define X(n, kind, init) #n,
      PWS_HRU_VARS(X)
#undef X
  };
  return nm[i];
}
int shim_n_vars() { return PWS_N_HRU_VARS; }
}
