// libmek: kernels and the C ABI (include/mek.h) of the batched MPC-EKF fast-charge control loop.
//
// Execution model: one thread per cell; a persistent grid walks the shard's cells and, for each
// cell, runs every sample instant of the charge without leaving the kernel (runMPC.m:83-112).
// The hot rows of the ROM table bank are staged once per CTA into shared memory with one TMA bulk
// copy (cp.async.bulk + mbarrier); per-cell state, covariance and trajectory logs are
// structure-of-arrays in HBM with the cell index fastest, so each warp access is coalesced.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/mek.h"
#include "mek_core.cuh"
#ifndef MEK_REGROUP_DEFAULT
#define MEK_REGROUP_DEFAULT 128  // sample instants between regroupings of a heterogeneous population (mek_config.regroup_every == 0)
#endif
#ifndef MEK_DEFER_DEFAULT
#define MEK_DEFER_DEFAULT 16     // sample instants per deferral block when mek_config.defer_block == 0
#endif

using namespace mek;

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                            \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) return fail(MEK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// ------------------------------------------------------------------------------------------
// bank (host object) : resolves tfData names to rows, packs the hot rows for shared memory
// ------------------------------------------------------------------------------------------
struct mek_bank {
  int n, nz, nT, nZ, M;
  double Tsamp;
  std::vector<double> Tpts_K, Zpts, A, C, D, params;
  std::vector<std::string> names;
  std::vector<double> xloc;
  // resolved rows (0-based)
  int r_ifdl0, r_ifdl3, r_if0, r_if3, r_tss0, r_tss3, r_the0, r_the3, r_phie_last, r_phise0, r_phi;
  std::vector<int> negtss, postss, thetae, negphise, posphise, phie, posphis;
  // hot bank blob: [Tpts nT][Zpts nZ][pad][M][mstride]
  int R, nx_negtss, nx_postss, nx_thetae, mstride, off_mdl;
  std::vector<double> blob;
  // optional tabulated cell functions (mek_bank_set_lut): LUT_NTAB tables of lut_n points over [lut_lo, lut_hi]
  int lut_n = 0;
  double lut_lo = 0.0, lut_hi = 1.0;
  std::vector<double> lut;
};

static std::vector<int> find_rows(const mek_bank& b, const char* nm, double loc = -1.0, bool use_loc = false) {
  std::vector<int> out;
  for (int i = 0; i < b.nz; ++i)
    if (b.names[i] == nm && (!use_loc || b.xloc[i] == loc)) out.push_back(i);
  return out;
}
static void append(std::vector<int>& a, const std::vector<int>& b) { a.insert(a.end(), b.begin(), b.end()); }

static int bank_finalize(mek_bank* b) {
  // structural checks the reference performs at set-up
  for (int m = 0; m < b->M; ++m)
    if (b->A[(size_t)m * 6 + 5] != 1.0) return fail(MEK_ERR_BANK, "A does not have integrator state (initKF.m:74-77)");
  for (int i = 1; i < b->nT; ++i)
    if (!(b->Tpts_K[i] > b->Tpts_K[i - 1])) return fail(MEK_ERR_BANK, "temperature set-points must be ascending");
  for (int i = 1; i < b->nZ; ++i)
    if (!(b->Zpts[i] > b->Zpts[i - 1])) return fail(MEK_ERR_BANK, "SOC set-points must be ascending");
  const double eps = 2.220446049250313e-16;
  std::vector<int> negIfdl0 = find_rows(*b, "negIfdl", 0, true), posIfdl3 = find_rows(*b, "posIfdl", 3, true);
  std::vector<int> negIf0 = find_rows(*b, "negIf", 0, true), posIf3 = find_rows(*b, "posIf", 3, true);
  std::vector<int> negT0 = find_rows(*b, "negThetass", 0, true), posT3 = find_rows(*b, "posThetass", 3, true);
  std::vector<int> negP0 = find_rows(*b, "negPhise", 0, true);
  b->negtss = find_rows(*b, "negThetass");
  b->postss = find_rows(*b, "posThetass");
  b->negphise = find_rows(*b, "negPhise");
  b->posphise = find_rows(*b, "posPhise");
  b->posphis = find_rows(*b, "posPhis");
  b->thetae = find_rows(*b, "negThetae");
  append(b->thetae, find_rows(*b, "sepThetae"));
  append(b->thetae, find_rows(*b, "posThetae"));
  b->phie = find_rows(*b, "negPhie");
  append(b->phie, find_rows(*b, "sepPhie"));
  append(b->phie, find_rows(*b, "posPhie"));
  if (negIfdl0.empty()) return fail(MEK_ERR_BANK, "Requires ifdl at negative collector (OB_step.m:140)");
  if (posIfdl3.empty()) return fail(MEK_ERR_BANK, "Requires ifdl at positive collector (OB_step.m:141)");
  if (negIf0.empty()) return fail(MEK_ERR_BANK, "Requires if at negative collector (OB_step.m:142)");
  if (posIf3.empty()) return fail(MEK_ERR_BANK, "Requires if at positive collector (OB_step.m:143)");
  if (b->thetae.empty() || b->xloc[b->thetae.front()] > 0) return fail(MEK_ERR_BANK, "Requires thetae at negative collector (OB_step.m:144)");
  if (b->xloc[b->thetae.back()] > 3 + eps || b->xloc[b->thetae.back()] < 3 - eps)
    return fail(MEK_ERR_BANK, "Requires thetae at positive collector (OB_step.m:145)");
  if (negT0.empty()) return fail(MEK_ERR_BANK, "Requires thetass at negative collector (OB_step.m:148)");
  if (posT3.empty()) return fail(MEK_ERR_BANK, "Requires thetass at positive collector (OB_step.m:149)");
  if (negP0.empty()) return fail(MEK_ERR_BANK, "Requires phise at negative collector (OB_step.m:150)");
  if (!b->phie.empty() && b->xloc[b->phie.front()] == 0) b->phie.erase(b->phie.begin());   // OB_step.m:151-155
  if (b->phie.empty() || b->xloc[b->phie.back()] > 3 + eps || b->xloc[b->phie.back()] < 3 - eps)
    return fail(MEK_ERR_BANK, "Requires phie at positive collector (OB_step.m:156)");
  if (b->negphise.size() < 2) return fail(MEK_ERR_BANK, "EKFmatsHandler.m:97 needs ind.negPhise(2)");
  b->r_ifdl0 = negIfdl0[0]; b->r_ifdl3 = posIfdl3[0]; b->r_if0 = negIf0[0]; b->r_if3 = posIf3[0];
  b->r_tss0 = negT0[0]; b->r_tss3 = posT3[0];
  b->r_the0 = b->thetae.front(); b->r_the3 = b->thetae.back();
  b->r_phie_last = b->phie.back(); b->r_phise0 = negP0[0]; b->r_phi = b->negphise[1];
  // hot rows: the ten fixed ones, then the rows that only feed the EKF's warning counters
  std::vector<int> hot = {b->r_ifdl0, b->r_ifdl3, b->r_if0, b->r_if3, b->r_tss0, b->r_tss3, b->r_the0, b->r_the3, b->r_phie_last, b->r_phi};
  b->nx_negtss = b->nx_postss = b->nx_thetae = 0;
  for (int r : b->negtss) if (r != b->r_tss0) { hot.push_back(r); b->nx_negtss++; }
  for (int r : b->postss) if (r != b->r_tss3) { hot.push_back(r); b->nx_postss++; }
  for (int r : b->thetae) if (r != b->r_the0 && r != b->r_the3) { hot.push_back(r); b->nx_thetae++; }
  b->R = (int)hot.size();
  if (b->R > HR_MAX) return fail(MEK_ERR_UNSUPPORTED, "too many Thetass/Thetae rows for the shared-memory hot bank");
  b->mstride = HB_ROWS + b->R * HRS;
  int head = b->nT + b->nZ;
  if (head & 1) head++;                                      // keep the model block 16-byte aligned
  b->off_mdl = head;
  size_t total = (size_t)head + (size_t)b->M * b->mstride;
  if (total & 1) total++;                                    // TMA bulk size must be a multiple of 16 bytes
  b->blob.assign(total, 0.0);
  for (int i = 0; i < b->nT; ++i) b->blob[i] = b->Tpts_K[i];
  for (int i = 0; i < b->nZ; ++i) b->blob[b->nT + i] = b->Zpts[i];
  for (int m = 0; m < b->M; ++m) {
    double* dst = &b->blob[(size_t)b->off_mdl + (size_t)m * b->mstride];
    for (int j = 0; j < 6; ++j) dst[j] = b->A[(size_t)m * 6 + j];
    {                                                        // a_i*a_j, upper triangle row-major (sym_idx order): the covariance time update's factors
      int q = 0;
      for (int i = 0; i < N5; ++i)
        for (int j = i; j < N5; ++j, ++q) dst[HB_AA + q] = b->A[(size_t)m * 6 + i] * b->A[(size_t)m * 6 + j];
    }
    for (int r = 0; r < b->R; ++r) {
      const double* Cr = &b->C[((size_t)m * b->nz + hot[r]) * 6];
      double* row = dst + HB_ROWS + r * HRS;                         // [C(1..5) | D | C(6) | pad]  (mek_core.cuh HRS)
      for (int j = 0; j < 5; ++j) row[j] = Cr[j];
      row[HR_D] = b->D[(size_t)m * b->nz + hot[r]];
      row[HR_C6] = Cr[5];
      row[7] = 0.0;
    }
  }
  return MEK_OK;
}

extern "C" int mek_bank_create(const mek_bank_desc* d, mek_bank** out) {
  if (!d || !out) return fail(MEK_ERR_ARG, "null argument");
  if (d->n != N5) return fail(MEK_ERR_UNSUPPORTED, "bank must have n = 5 transient states (runMPC.m:17)");
  if (d->nparams != N_PARAMS) return fail(MEK_ERR_ARG, "parameter vector has wrong length");
  if (d->nT < 1 || d->nZ < 1 || d->nz < 1 || d->nz > 64) return fail(MEK_ERR_ARG, "bad bank dimensions");
  mek_bank* b = new mek_bank();
  b->n = d->n; b->nz = d->nz; b->nT = d->nT; b->nZ = d->nZ; b->M = d->nT * d->nZ; b->Tsamp = d->Tsamp;
  for (int i = 0; i < d->nT; ++i) b->Tpts_K.push_back(d->Tpts_C[i] + 273.15);     // OB_step.m:166
  for (int i = 0; i < d->nZ; ++i) b->Zpts.push_back(d->Zpts_pct[i] / 100.0);      // OB_step.m:167
  b->A.assign(d->A, d->A + (size_t)b->M * 6);
  b->C.assign(d->C, d->C + (size_t)b->M * b->nz * 6);
  b->D.assign(d->D, d->D + (size_t)b->M * b->nz);
  b->params.assign(d->params, d->params + N_PARAMS);
  for (int i = 0; i < d->nz; ++i) { b->names.push_back(d->names[i]); b->xloc.push_back(d->xloc[i]); }
  int rc = bank_finalize(b);
  if (rc != MEK_OK) { delete b; return rc; }
  *out = b;
  return MEK_OK;
}

extern "C" int mek_bank_load(const char* path, mek_bank** out) {
  if (!path || !out) return fail(MEK_ERR_ARG, "null argument");
  FILE* f = fopen(path, "rb");
  if (!f) return fail(MEK_ERR_ARG, std::string("cannot open ") + path);
  std::vector<unsigned char> buf;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  buf.resize(sz > 0 ? (size_t)sz : 0);
  size_t got = fread(buf.data(), 1, buf.size(), f);
  fclose(f);
  if (got != buf.size() || buf.size() < 40 || memcmp(buf.data(), "MEKBANK1", 8) != 0) return fail(MEK_ERR_BANK, "not a MEKBANK1 file");
  size_t off = 8;
  int hdr[6];
  memcpy(hdr, &buf[off], 24); off += 24;
  const int n = hdr[1], nz = hdr[2], nT = hdr[3], nZ = hdr[4], npar = hdr[5];
  if (hdr[0] != 1 && hdr[0] != 2) return fail(MEK_ERR_BANK, "unsupported bank version");
  const size_t M = (size_t)nT * nZ;
  const size_t need = off + 8 + 8 * ((size_t)nT + nZ) + (size_t)nz * 24 + 8 * (2 * M + M * (n + 1) + M * nz * (n + 1) + M * nz + npar);
  if (n < 1 || nz < 1 || nT < 1 || nZ < 1 || buf.size() < need) return fail(MEK_ERR_BANK, "truncated bank file");
  double Tsamp;
  memcpy(&Tsamp, &buf[off], 8); off += 8;
  auto take = [&](size_t cnt) { std::vector<double> v(cnt); memcpy(v.data(), &buf[off], 8 * cnt); off += 8 * cnt; return v; };
  std::vector<double> Tp = take(nT), Zp = take(nZ);
  std::vector<std::string> names;
  std::vector<const char*> cnames;
  std::vector<double> xloc;
  for (int i = 0; i < nz; ++i) {
    char nm[17];
    memcpy(nm, &buf[off], 16); nm[16] = 0; off += 16;
    names.push_back(nm);
    double x;
    memcpy(&x, &buf[off], 8); off += 8;
    xloc.push_back(x);
  }
  for (auto& s : names) cnames.push_back(s.c_str());
  take(M); take(M);   // ROMmdls.T / .SOC per model (the grid vectors above are what the loop uses)
  std::vector<double> A = take(M * (n + 1)), C = take(M * nz * (n + 1)), D = take(M * nz), par = take(npar);
  mek_bank_desc d = {n, nz, nT, nZ, Tsamp, Tp.data(), Zp.data(), cnames.data(), xloc.data(), A.data(), C.data(), D.data(), par.data(), npar};
  int rc = mek_bank_create(&d, out);
  if (rc != MEK_OK || hdr[0] < 2) return rc;
  // version 2: [int32 lut_n][f64 lo][f64 hi][f64 tables[LUT_NTAB][lut_n]] behind the parameter vector (lut_n = 0: none)
  if (buf.size() < off + 20) { mek_bank_destroy(*out); *out = nullptr; return fail(MEK_ERR_BANK, "truncated bank file"); }
  int ln;
  double lo, hi;
  memcpy(&ln, &buf[off], 4); off += 4;
  memcpy(&lo, &buf[off], 8); off += 8;
  memcpy(&hi, &buf[off], 8); off += 8;
  if (ln > 0) {
    if (ln > (1 << 20) || buf.size() < off + 8 * (size_t)LUT_NTAB * (size_t)ln) { mek_bank_destroy(*out); *out = nullptr; return fail(MEK_ERR_BANK, "truncated bank file"); }
    std::vector<double> tb = take((size_t)LUT_NTAB * (size_t)ln);
    rc = mek_bank_set_lut(*out, ln, lo, hi, tb.data());
    if (rc != MEK_OK) { mek_bank_destroy(*out); *out = nullptr; }
  }
  return rc;
}

extern "C" double mek_noise(uint64_t seed, uint64_t cell, uint64_t k) { return noise_gauss(seed, cell, k); }

extern "C" int mek_bank_set_lut(mek_bank* b, int n, double theta_lo, double theta_hi, const double* tables) {
  if (!b) return fail(MEK_ERR_ARG, "null bank");
  if (n == 0) { b->lut_n = 0; b->lut.clear(); return MEK_OK; }     // back to the closed forms
  if (!tables || n < 2 || n > (1 << 20) || !(theta_hi > theta_lo)) return fail(MEK_ERR_ARG, "bad table description");
  for (size_t i = 0; i < (size_t)LUT_NTAB * (size_t)n; ++i)
    if (!(std::fabs(tables[i]) <= kDblMax)) return fail(MEK_ERR_BANK, "non-finite entry in a cell-function table");
  b->lut_n = n; b->lut_lo = theta_lo; b->lut_hi = theta_hi;
  b->lut.assign(tables, tables + (size_t)LUT_NTAB * (size_t)n);
  return MEK_OK;
}
extern "C" int mek_bank_lut_points(const mek_bank* b) { return b ? b->lut_n : 0; }

extern "C" void mek_bank_destroy(mek_bank* b) { delete b; }
extern "C" int mek_bank_dims(const mek_bank* b, int* n, int* nz, int* nT, int* nZ) {
  if (!b) return fail(MEK_ERR_ARG, "null bank");
  if (n) *n = b->n; if (nz) *nz = b->nz; if (nT) *nT = b->nT; if (nZ) *nZ = b->nZ;
  return MEK_OK;
}
extern "C" double mek_bank_capacity_Ah(const mek_bank* b) { return b ? b->params[P_Q] : NAN; }

#include "mek_kernels.cuh"
#include "mek_entry.h"
#include "mek_ekf_mb.cuh"

// ---- initialisation: initKF.m:37-133, initMPC.m:32-72, OB_step.m:63-71 and the first I = 0 call (runMPC.m:72-74) ----
__global__ void __launch_bounds__(128) k_pop_init(const __grid_constant__ Params P) {
  extern __shared__ __align__(16) double sm[];
  BankSmem bk;
  stage_bank(sm, P, bk);
  const Dims& d = P.d;
  const double* p = P.cp;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x; cell < d.B; cell += stride) {
    Cell c = cell_ref(P.st, d, cell);
    for (int f = 0; f < d.nfields; ++f) c.f(f) = 0.0;
    const long long id = cell_id(P, cell);
    const double Tc = P.in_T[id], z0 = P.in_soc0[id] / 100.0;
    const double Tk1 = Tc + 273.15;
    c.f(S_TC) = Tc;
    c.f(S_SOC0N) = soc_e(p, P_NEG, z0);                              // OB_step.m:64-71
    c.f(S_SOC0P) = soc_e(p, P_POS, z0);
    c.f(S_SOCN) = c.f(S_SOC0N);
    c.f(S_SOCP) = c.f(S_SOC0P);
    c.f(S_CDLN) = cdl_eff(p, P_NEG);                                 // OB_step.m:212-219 (run constants)
    c.f(S_CDLP) = cdl_eff(p, P_POS);
    c.f(S_ARRK0N) = arrhenius(p, Tk1, p[P_NEG + E_EA_K0]);
    c.f(S_ARRK0P) = arrhenius(p, Tk1, p[P_POS + E_EA_K0]);
    c.f(S_ARRRFN) = arrhenius(p, Tk1, -p[P_NEG + E_EA_RF]);
    c.f(S_ARRRFP) = arrhenius(p, Tk1, -p[P_POS + E_EA_RF]);
    const double* sx = P.in_sigx0 + id * 36;
    c.f(S_SIGX0) = sx[5 * 6 + 5];                                    // initKF.m:99
    c.f(S_EKFSOC0) = P.in_ekf[id] / 100.0;                           // initKF.m:133
    for (int m = 0; m < d.M; ++m) {                                  // initKF.m:70,94-95
      int q = 0;
      for (int i = 0; i < N5; ++i)
        for (int j = i; j < N5; ++j, ++q) c.f(d.f_sig + m * NSYM + q) = sx[j * 6 + i];
    }
    if (d.mb) {                                                      // initKF.m:100-101: xhat = 0 (zero fill above), SigmaX = SigmaX0
      int q = 0;
      for (int i = 0; i < N6; ++i)
        for (int j = i; j < N6; ++j, ++q) c.f(d.f_mbs + q) = sx[j * 6 + i];
    }
    for (int k = 0; k < 7; ++k) c.f(S_UMAX + k) = P.in_lim[id * 7 + k];
    c.f(S_MIN_PHISE) = INFINITY;
    c.f(S_MAX_V) = -INFINITY;
    c.f(S_T_REACH) = -1.0;
    plant_step<true>(d, bk, p, c, 0.0, -1);                          // runMPC.m:74 (Q20)
  }
}

// ---- function-level kernels (each is one line of the loop body) ----
__global__ void __launch_bounds__(128) k_ob_step(const __grid_constant__ Params P, const double* Iapp, double* V) {
  extern __shared__ __align__(16) double sm[];
  BankSmem bk;
  stage_bank(sm, P, bk);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x; cell < P.d.B; cell += stride) {
    Cell c = cell_ref(P.st, P.d, cell);
    const long long id = cell_id(P, cell);
    V[id] = plant_step<true>(P.d, bk, P.cp, c, Iapp[id], -1, nullptr,
                             plant_q_factor(P.d.plant_q_scale, P.d.plant_q_spread, P.d.noise_seed, (unsigned long long)(P.d.first_cell + id)), P.d.plant_rc_scale);
  }
}

__device__ __forceinline__ void pack_mats(const EkfOut& e, double* m) {
  int q = 0;
  m[q++] = (double)e.lost; m[q++] = (double)(e.iT + 1); m[q++] = (double)(e.iZ + 1);
  for (int j = 0; j < 6; ++j) m[q++] = e.xhat[j];
  for (int j = 0; j < 5; ++j) m[q++] = e.a[j];
  for (int j = 0; j < 5; ++j) m[q++] = e.Cv[j];
  m[q++] = e.Dv;
  for (int j = 0; j < 5; ++j) m[q++] = e.Cphi[j];
  m[q++] = e.Dphi; m[q++] = e.bv; m[q++] = e.bphi; m[q++] = e.vhat; m[q++] = e.soc;
  m[q++] = 0.0;   // reserved
}
__device__ __forceinline__ void unpack_mats(const double* m, EkfOut& e) {
  int q = 0;
  e.lost = (int)m[q++]; e.iT = (int)m[q++] - 1; e.iZ = (int)m[q++] - 1;
  for (int j = 0; j < 6; ++j) e.xhat[j] = m[q++];
  for (int j = 0; j < 5; ++j) e.a[j] = m[q++];
  for (int j = 0; j < 5; ++j) e.Cv[j] = m[q++];
  e.Dv = m[q++];
  for (int j = 0; j < 5; ++j) e.Cphi[j] = m[q++];
  e.Dphi = m[q++]; e.bv = m[q++]; e.bphi = m[q++]; e.vhat = m[q++]; e.soc = m[q++];
}

// full zk = [Z; vhat; SOC] (iterEKF.m:183) for the function-level entry: all nz rows from the full bank
__device__ void full_zk(const Params& P, const BankSmem& bk, const Cell& c, double ik, double vhat, double soc, double* zk) {
  const Dims& d = P.d;
  const double Tk = c.f(S_TC) + 273.15;
  Scratch sc{const_cast<double*>(bk.Tpts) + P.blob_doubles + threadIdx.x, (int)blockDim.x};
  get_xind(d, bk, Tk, soc, sc, -1);                                  // iterEKF.m:181
  double xk[4][N5], gam4[4];
  int mk[4];
  for (int k = 0; k < 4; ++k) {
    mk[k] = (int)sc[SC_M + k];
    gam4[k] = sc[SC_GAMMA + k];
    for (int j = 0; j < N5; ++j) xk[k][j] = c.xhat(mk[k], j);
  }
  for (int r = 0; r < d.nz; ++r) {                                   // iterEKF.m:302-313
    double yc[4];
    for (int k = 0; k < 4; ++k) {
      const double* cr = P.Cfull + ((size_t)mk[k] * d.nz + r) * 6;
      double acc = cr[0] * xk[k][0];
      for (int j = 1; j < N5; ++j) acc = fma(cr[j], xk[k][j], acc);
      yc[k] = fma(P.Dfull[(size_t)mk[k] * d.nz + r], ik, acc);
    }
    zk[r] = blend4(yc, gam4);
  }
  zk[d.nz] = vhat;
  zk[d.nz + 1] = soc;       // the per-kind corrections (iterEKF.m:331-413) follow in k_iter_ekf
}


// boundzk = 3*sqrt([diag(SigmaZ); SigmaV; SigmaSOC]) (iterEKF.m:185-205, getChatZ :523-602).  Function-level entry only:
// the control path never reads the bounds.  zk is the corrected output vector of this call.
__device__ void ekf_bounds(const Params& P, const BankSmem& bk, const RowLists& RL, const Cell& c, Scratch sc, const double* zk,
                           double* bound) {
  const Dims& d = P.d;
  const double* p = P.cp;
  const double Tk = c.f(S_TC) + 273.15;
  const double rpa = d.Ts / (3600.0 * p[P_Q]);
  const double sx0 = c.f(S_SIGX0);
  const double xSOC = c.f(S_EKFSOC0) - c.f(S_X0) * rpa;              // :576
  const double sn = soc_e(p, P_NEG, xSOC), sp = soc_e(p, P_POS, xSOC);
  // ---- getChatV(Xind, zk, Tk) (:421-519) ----
  const double Rfn = rf_e(p, P_NEG, sn, c.f(S_ARRRFN)), Rfp = rf_e(p, P_POS, sp, c.f(S_ARRRFP));
  const double k0n = k0_e(p, P_NEG, sn, c.f(S_ARRK0N)), k0p = k0_e(p, P_POS, sp, c.f(S_ARRK0P));
  const double th0 = zk[RL.r_tss0], th3 = zk[RL.r_tss3];
  const double i0n = k0n * sqrt((zk[RL.r_the0] * (1.0 - th0)) * th0);
  const double i0p = k0p * sqrt((zk[RL.r_the3] * (1.0 - th3)) * th3);
  const double Rctn = (p[P_R] * Tk) / (p[P_F] * i0n), Rctp = (p[P_R] * Tk) / (p[P_F] * i0p);
  const double dUn0 = duocp_n(p, th0), dUp3 = duocp_p(p, th3);
  const double dsn = soc_e(p, P_NEG, 1.0) - soc_e(p, P_NEG, 0.0), dsp = soc_e(p, P_POS, 1.0) - soc_e(p, P_POS, 0.0);
  const double ChatV0 = (((-dUp3) * d.Ts) * dsp) / (3600.0 * p[P_Q]) - (((-dUn0) * d.Ts) * dsn) / (3600.0 * p[P_Q]);   // :497-501
  const double res0n = ((-d.Ts) * dsn) / (3600.0 * p[P_Q]), res0p = ((-d.Ts) * dsp) / (3600.0 * p[P_Q]);              // :562-565
  const double dUn = duocp_n(p, sn), dUp = duocp_p(p, sp);           // :579-580
  double S1[NSYM], ChatV[4][N5], g[4];
  int mk[4];
  for (int k = 0; k < 4; ++k) { mk[k] = (int)sc[SC_M + k]; g[k] = sc[SC_GAMMA + k]; }
  for (int q = 0; q < NSYM; ++q) S1[q] = c.sig(mk[0], q);            // :192 (Q4): neighbour #1's SigmaX for all four terms
  auto quad = [&](const double* row) {                                // row*SigmaX*row'
    double t = 0.0;
    for (int i = 0; i < N5; ++i) {
      double acc = 0.0;
      for (int j = 0; j < N5; ++j) acc += S1[sym_idx(i, j)] * row[j];
      t += row[i] * acc;
    }
    return t;
  };
  double sigV = 0.0;
  for (int k = 0; k < 4; ++k) {
    const double* Cm = P.Cfull + (size_t)mk[k] * d.nz * 6;
    for (int j = 0; j < N5; ++j) {
      double v = Rfp * (g[k] * Cm[RL.r_ifdl3 * 6 + j]) - Rfn * (g[k] * Cm[RL.r_ifdl0 * 6 + j]);
      v = (v + Rctp * (g[k] * Cm[RL.r_if3 * 6 + j])) - Rctn * (g[k] * Cm[RL.r_if0 * 6 + j]);
      v = v + g[k] * Cm[RL.r_phie_last * 6 + j];
      v = v + (dUp3 * (g[k] * Cm[RL.r_tss3 * 6 + j]) - dUn0 * (g[k] * Cm[RL.r_tss0 * 6 + j]));
      ChatV[k][j] = v;
    }
    sigV += quad(ChatV[k]);                                          // :194
  }
  sigV += (ChatV0 * sx0) * ChatV0;                                   // :197
  for (int r = 0; r < d.nz; ++r) {
    const int kind = RL.kind[r];
    double c0 = 0.0;                                                 // ChatZ0(r) (:540-586)
    if (kind == 1) c0 = ChatV0;
    else if (kind == 2) c0 = res0n;
    else if (kind == 3) c0 = res0p;
    else if (kind == 4) c0 = dUn * res0n;
    else if (kind == 5) c0 = dUp * res0p;
    else if (kind == 6) c0 = -dUn * res0n;
    double acc = 0.0;
    for (int k = 0; k < 4; ++k) {
      const double* Cm = P.Cfull + (size_t)mk[k] * d.nz * 6;
      double row[N5];
      for (int j = 0; j < N5; ++j) {
        double v = g[k] * Cm[r * 6 + j];
        if (kind == 1) v += ChatV[k][j];                             // :546-552 posPhis carries the cell voltage
        if (kind == 6) v -= g[k] * Cm[RL.r_phise0 * 6 + j];          // :589-595 Phie is referred to Phise(0)
        row[j] = v;
      }
      acc += quad(row);                                              // :193
    }
    acc += (c0 * sx0) * c0;                                          // :196
    bound[r] = 3.0 * sqrt(acc);                                      // :205
  }
  bound[d.nz] = 3.0 * sqrt(sigV);
  const double res = -d.Ts / (3600.0 * p[P_Q]);
  bound[d.nz + 1] = 3.0 * sqrt((res * sx0) * res);                   // :198
}

__global__ void __launch_bounds__(128) k_iter_ekf(const __grid_constant__ Params P, const __grid_constant__ RowLists RL, const double* vk,
                                                   const double* ik, double* zk, double* mats, double* bounds) {
  extern __shared__ __align__(16) double sm[];
  BankSmem bk;
  stage_bank(sm, P, bk);
  const Dims& d = P.d;
  const double* p = P.cp;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x; cell < d.B; cell += stride) {
    Cell c = cell_ref(P.st, d, cell);
    EkfOut e;
    memset(&e, 0, sizeof(e));
    const long long id = cell_id(P, cell);
    ekf_step<true>(d, bk, p, c, vk[id], ik[id], e, Scratch{sm + P.blob_doubles + threadIdx.x, (int)blockDim.x}, -1);
    if (mats) pack_mats(e, mats + id * MEK_NMATS);
    if (zk) {
      double* z = zk + id * (d.nz + 2);
      if (e.lost) {
        for (int r = 0; r < d.nz + 2; ++r) { z[r] = NAN; if (bounds) bounds[id * (d.nz + 2) + r] = NAN; }
        continue;
      }
      full_zk(P, bk, c, ik[id], e.vhat, e.soc, z);
      // nonlinear corrections of getVariables on the full vector (iterEKF.m:331-413)
      const double rpa = d.Ts / (3600.0 * p[P_Q]);
      const double xSOC = c.f(S_EKFSOC0) - c.f(S_X0) * rpa;
      double SOCn = soc_e(p, P_NEG, xSOC), SOCp = soc_e(p, P_POS, xSOC);
      if (SOCn < 0.0) SOCn = 1e-6;
      if (SOCn > 1.0) SOCn = 1.0 - 1e-6;
      if (SOCp < 0.0) SOCp = 1e-6;
      if (SOCp > 0.998) SOCp = 0.998;
      for (int k = 0; k < RL.n_negtss; ++k) {
        double v = z[RL.negtss[k]] + SOCn;
        if (v < 0.0) v = 1e-6;
        if (v > 1.0) v = 1.0 - 1e-6;
        z[RL.negtss[k]] = v;
      }
      for (int k = 0; k < RL.n_postss; ++k) {
        double v = z[RL.postss[k]] + SOCp;
        if (v < 0.0) v = 1e-6;
        if (v > 0.998) v = 0.998;
        z[RL.postss[k]] = v;
      }
      const double UnAvg = uocp_n(p, SOCn), UpAvg = uocp_p(p, SOCp);
      for (int k = 0; k < RL.n_negphise; ++k) z[RL.negphise[k]] = z[RL.negphise[k]] + UnAvg;
      for (int k = 0; k < RL.n_posphise; ++k) z[RL.posphise[k]] = z[RL.posphise[k]] + UpAvg;
      const double Phise0 = z[RL.r_phise0];
      for (int k = 0; k < RL.n_phie; ++k) z[RL.phie[k]] = z[RL.phie[k]] - Phise0;
      for (int k = 0; k < RL.n_thetae; ++k) {
        double v = z[RL.thetae[k]] + 1.0;
        if (v < 0.0) v = 1e-6;
        z[RL.thetae[k]] = v;
      }
      for (int k = 0; k < RL.n_posphis; ++k) z[RL.posphis[k]] = z[RL.posphis[k]] + e.vhat;
      if (bounds) ekf_bounds(P, bk, RL, c, Scratch{sm + P.blob_doubles + threadIdx.x, (int)blockDim.x}, z, bounds + id * (d.nz + 2));
    }
  }
}

// iterEKF with the model-blend method + EKFmatsHandler (function-level entry of a population created with ekf_method = 1)
__global__ void __launch_bounds__(128) k_iter_ekf_mb(const __grid_constant__ Params P, const __grid_constant__ RowLists RL, const double* vk,
                                                      const double* ik, double* zk, double* mats, double* bounds) {
  extern __shared__ __align__(16) double sm[];
  BankSmem bk;
  stage_bank(sm, P, bk);
  const Dims& d = P.d;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x; cell < d.B; cell += stride) {
    Cell c = cell_ref(P.st, d, cell);
    EkfOut e;
    memset(&e, 0, sizeof(e));
    const long long id = cell_id(P, cell);
    ekf_step_mb(P, bk, RL, c, vk[id], ik[id], e, Scratch{sm + P.blob_doubles + threadIdx.x, (int)blockDim.x},
                zk ? zk + id * (d.nz + 2) : nullptr, bounds ? bounds + id * (d.nz + 2) : nullptr);
    if (mats) pack_mats(e, mats + id * MEK_NMATS);
  }
}

// The loop of runMPC.m:83-112 with the model-blend filter (a script whose initKF call says 'MdlB' instead of runMPC.m:16's
// 'OutB').  Plain instant-by-instant form: OB_step advances all M local models (the filter keeps ONE blended model per
// cell, so there is no per-model EKF stream and nothing to defer), then ekf_step_mb, then the runtime-shape MPC.
template <int NPM, int NCM>
__global__ void __launch_bounds__(128) k_run_population_mb(const __grid_constant__ Params P, const __grid_constant__ RowLists RL, int k0, int k1) {
  extern __shared__ __align__(16) double sm[];
  BankSmem bk;
  stage_bank(sm, P, bk);
  const Dims& d = P.d;
  const Scratch sc{sm + P.blob_doubles + threadIdx.x, (int)blockDim.x};
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x; cell < d.B; cell += stride) {
    Cell c = cell_ref(P.st, d, cell);
    const int tt = two_nearest(c.f(S_TC) + 273.15, bk.Tpts, d.nT);
    const long long id = cell_id(P, cell);
    const double qs = plant_q_factor(d.plant_q_scale, d.plant_q_spread, d.noise_seed, (unsigned long long)(d.first_cell + id));
#pragma unroll 1
    for (int k = k0; k < k1; ++k) {
      StepLog lg;
      lg.u = NAN; lg.V = NAN; lg.soc = NAN; lg.phise = NAN; lg.iT = 0; lg.iZ = 0; lg.nexec = 0;
      for (int w = 0; w < 4; ++w) lg.active[w] = 0u;
      if (!((unsigned)c.f(S_STATUS) & ST_LOST)) {                    // Q19: a lost cell idles
        const double uk = c.f(S_UK);
        const double V = plant_step<true>(d, bk, P.cp, c, uk, tt, nullptr, qs, d.plant_rc_scale);   // runMPC.m:85
        EkfOut e;
        memset(&e, 0, sizeof(e));
        double vk = V;
        if (d.meas_sigma != 0.0) vk = V + d.meas_sigma * noise_gauss(d.noise_seed, (unsigned long long)(d.first_cell + id), (unsigned long long)k);
        ekf_step_mb(P, bk, RL, c, vk, uk, e, sc, nullptr, nullptr);   // runMPC.m:88,92
        if (!e.lost) {
          MpcOut o;
          mpc_step<NPM, NCM>(d, P.cp, c, e, uk, o);                  // runMPC.m:94-103
          lg.u = o.u; lg.V = V; lg.soc = e.soc; lg.phise = o.phise;  // runMPC.m:106-109
          lg.iT = e.iT + 1; lg.iZ = e.iZ + 1; lg.nexec = o.nexec;
          for (int w = 0; w < 4; ++w) lg.active[w] = o.active[w];
        }
      }
      if (lg.nexec > 0) {
        c.f(S_SUM_NEXEC) += (double)lg.nexec;
        c.f(S_N_QP) += 1.0;
        if (lg.nexec >= d.maxHild) c.f(S_N_CAP) += 1.0;
      }
      c.f(S_MIN_PHISE) = fmin(c.f(S_MIN_PHISE), lg.phise);
      c.f(S_MAX_V) = fmax(c.f(S_MAX_V), lg.V);
      if (c.f(S_T_REACH) < 0.0 && lg.soc >= c.f(S_ZMAX) - 1e-3) c.f(S_T_REACH) = (double)k;
      c.f(S_LAST_SOC) = lg.soc; c.f(S_LAST_V) = lg.V; c.f(S_LAST_PHISE) = lg.phise;
      if (id < P.nlog && k < P.log_steps) {
        const long long o = (long long)k * P.nlog + id;
        P.log_u[o] = lg.u; P.log_V[o] = lg.V; P.log_soc[o] = lg.soc; P.log_phise[o] = lg.phise;
        P.log_iT[o] = (unsigned char)lg.iT; P.log_iZ[o] = (unsigned char)lg.iZ; P.log_nexec[o] = (unsigned short)lg.nexec;
        for (int w = 0; w < P.log_words; ++w) P.log_active[o * P.log_words + w] = lg.active[w];
      }
    }
    c.f(S_STEPS) += (double)(k1 - k0);
  }
}

template <int NPM, int NCM>
__global__ void __launch_bounds__(128) k_iter_mpc(const __grid_constant__ Params P, const double* mats, const double* uk_applied, double* uk,
                                                   double* phise, int* nexec, unsigned* active, double* cost) {
  const Dims& d = P.d;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x; cell < d.B; cell += stride) {
    Cell c = cell_ref(P.st, d, cell);
    EkfOut e;
    const long long id = cell_id(P, cell);
    unpack_mats(mats + id * MEK_NMATS, e);
    if (e.lost) {
      uk[id] = NAN; if (phise) phise[id] = NAN; if (nexec) nexec[id] = 0;
      if (active) for (int w = 0; w < 4; ++w) active[id * 4 + w] = 0u;
      if (cost) for (int w = 0; w < MEK_NCOST; ++w) cost[id * MEK_NCOST + w] = NAN;
      continue;
    }
    MpcOut o;
    double cj[MEK_NCOST];
    mpc_step<NPM, NCM>(d, P.cp, c, e, uk_applied[id], o, cost ? cj : nullptr);
    if (cost) for (int w = 0; w < MEK_NCOST; ++w) cost[id * MEK_NCOST + w] = cj[w];
    uk[id] = o.u;
    if (phise) phise[id] = o.phise;
    if (nexec) nexec[id] = o.nexec;
    if (active) for (int w = 0; w < 4; ++w) active[id * 4 + w] = o.active[w];
  }
}

// stateless hildreth.m: one thread per QP
template <int NCM, int NCONM>
__global__ void __launch_bounds__(128) k_hildreth(long long B, int Nc, int nC, const double* E, const double* F, const double* M,
                                                   const double* gamma, const double* lambda0, int maxIter, double* DU, double* lam_out,
                                                   int* nexec) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < B; q += stride) {
    double Ep[NCM * NCM], LU[NCM * NCM], Fp[NCM], Mp[NCONM * NCM], gp[NCONM], lam[NCONM], du[NCM];
    for (int i = 0; i < NCM * NCM; ++i) Ep[i] = 0.0;
    for (int i = 0; i < Nc; ++i) {
      Fp[i] = F[q * Nc + i];
      for (int j = 0; j < Nc; ++j) Ep[i * NCM + j] = E[q * Nc * Nc + i * Nc + j];
    }
    for (int i = 0; i < nC; ++i) {
      for (int j = 0; j < NCM; ++j) Mp[i * NCM + j] = (j < Nc) ? M[(q * nC + i) * Nc + j] : 0.0;
      gp[i] = gamma[q * nC + i];
      lam[i] = lambda0 ? lambda0[q * nC + i] : 0.0;                  // hildreth.m:18-22
    }
    spd_factor<NCM>(Nc, Ep, LU);
    const int k = hildreth_core<NCM, NCONM>(Nc, nC, LU, Fp, Mp, gp, maxIter, lam, du);
    for (int i = 0; i < Nc; ++i) DU[q * Nc + i] = du[i];
    for (int i = 0; i < nC; ++i) lam_out[q * nC + i] = lam[i];
    nexec[q] = k;
  }
}


// stateless constraintsMPC.m (+ the pre-QP part of iterMPC.m: F, E, unconstrained DU)
template <int NPM, int NCM>
__global__ void __launch_bounds__(128) k_constraints(long long B, const __grid_constant__ Dims d, double r, const double* x_aug,
                                                      const double* mats, const double* SOCk_1, const double* limits, double* M, double* gamma,
                                                      double* E, double* F, double* DU) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < B; q += stride) {
    EkfOut e;
    unpack_mats(mats + q * MEK_NMATS, e);
    for (int j = 0; j < 6; ++j) e.xhat[j] = x_aug[q * 7 + j];       // x_aug_k = [x_k; u_k]  constraintsMPC.m:4
    e.soc = SOCk_1[q];                                               // mpcData.SOCk_1 (runMPC.m:99)
    const double* l = limits + q * 7;
    Limits lim = {l[0], l[1], l[2], l[3], l[4], l[5], l[6]};
    MpcProblem<NPM, NCM> pr;
    mpc_setup<NPM, NCM>(d, r, lim, e, x_aug[q * 7 + 6], pr);
    for (int i = 0; i < pr.nC; ++i) {
      for (int j = 0; j < d.Nc; ++j) M[(q * d.nC + i) * d.Nc + j] = pr.Mm[i * NCM + j];
      gamma[q * d.nC + i] = pr.gam[i];
    }
    for (int i = 0; i < d.Nc; ++i) {
      if (F) F[q * d.Nc + i] = pr.F[i];
      if (DU) DU[q * d.Nc + i] = pr.DU[i];
      if (E) for (int j = 0; j < d.Nc; ++j) E[(q * d.Nc + i) * d.Nc + j] = pr.E[i * NCM + j];
    }
  }
}

// stateless predMat.m (scalar output, diagonal A, B = 1): Phi [Np][7], G [Np][Nc]
__global__ void __launch_bounds__(128) k_pred_mat(long long B, const double* a, const double* cc, const double* dd, int Np, int Nc,
                                                   double* Phi, double* G) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < B; q += stride) {
    const double* aq = a + q * 6;
    const double* cq = cc + q * 6;
    double pk[6], pw[6], h[33];
    for (int j = 0; j < 6; ++j) { pk[j] = 0.0; pw[j] = 1.0; }
    for (int k = 0; k <= Np; ++k) {                                  // predMat.m:31-36 (one extra term: Phi(Np,7))
      double acc = cq[0] * pk[0];
      for (int j = 1; j < 6; ++j) acc = acc + cq[j] * pk[j];
      h[k] = acc + dd[q];
      for (int j = 0; j < 6; ++j) pk[j] = aq[j] * pk[j] + 1.0;
    }
    for (int i = 1; i <= Np; ++i) {                                  // predMat.m:39-50
      double* ph = Phi + (q * Np + (i - 1)) * 7;
      for (int j = 0; j < 6; ++j) { pw[j] = aq[j] * pw[j]; ph[j] = cq[j] * pw[j]; }
      ph[6] = h[i];
      for (int j = 0; j < Nc; ++j) G[(q * Np + (i - 1)) * Nc + j] = (i - 1 - j >= 0) ? h[i - 1 - j] : 0.0;
    }
  }
}

// population statistics: per-block partials in shared memory, then one atomic per block
__global__ void __launch_bounds__(256) k_summary(const __grid_constant__ Params P, double* totals) {
  __shared__ double sh[MEK_NSTAT][8];
  const Dims& d = P.d;
  double acc[MEK_NSTAT];
  for (int i = 0; i < MEK_NSTAT; ++i) acc[i] = 0.0;
  acc[MEK_STAT_MIN_PHISE] = INFINITY;
  acc[MEK_STAT_MAX_V] = -INFINITY;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x; cell < d.B; cell += stride) {
    Cell c = cell_ref(P.st, d, cell);
    const unsigned st = (unsigned)c.f(S_STATUS);
    acc[MEK_STAT_CELLS] += 1.0;
    acc[MEK_STAT_CELLSTEPS] += c.f(S_STEPS);
    acc[MEK_STAT_SUM_NEXEC] += c.f(S_SUM_NEXEC);
    acc[MEK_STAT_N_QP] += c.f(S_N_QP);
    acc[MEK_STAT_N_CAP] += c.f(S_N_CAP);
    const double tr = c.f(S_T_REACH);
    if (tr >= 0.0) { acc[MEK_STAT_N_REACHED] += 1.0; acc[MEK_STAT_SUM_T_REACH] += tr; }
    acc[MEK_STAT_MIN_PHISE] = fmin(acc[MEK_STAT_MIN_PHISE], c.f(S_MIN_PHISE));
    acc[MEK_STAT_MAX_V] = fmax(acc[MEK_STAT_MAX_V], c.f(S_MAX_V));
    if (st & ST_LOST) acc[MEK_STAT_N_LOST] += 1.0;
    if (st & ST_NONFINITE) acc[MEK_STAT_N_NONFINITE] += 1.0;
    if (st & ST_Q18) acc[MEK_STAT_N_Q18] += 1.0;
    const double ls = c.f(S_LAST_SOC);
    if (ls == ls) acc[MEK_STAT_SUM_SOC] += ls;
    acc[MEK_STAT_N_JACOBI] += c.f(S_NJAC);
    acc[MEK_STAT_N_BUMP] += c.f(S_NBUMP);
    acc[MEK_STAT_N_WARN] += c.f(S_WARN);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = 0; i < MEK_NSTAT; ++i) {
    double v = acc[i];
    for (int o = 16; o > 0; o >>= 1) {
      const double w = __shfl_down_sync(0xffffffffu, v, o);
      if (i == MEK_STAT_MIN_PHISE) v = fmin(v, w); else if (i == MEK_STAT_MAX_V) v = fmax(v, w); else v += w;
    }
    if (lane == 0) sh[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < MEK_NSTAT) {
    const int i = threadIdx.x;
    double v = sh[i][0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      if (i == MEK_STAT_MIN_PHISE) v = fmin(v, sh[i][w]); else if (i == MEK_STAT_MAX_V) v = fmax(v, sh[i][w]); else v += sh[i][w];
    }
    if (i == MEK_STAT_MIN_PHISE || i == MEK_STAT_MAX_V) {
      unsigned long long* addr = reinterpret_cast<unsigned long long*>(totals + i);
      unsigned long long old = *addr, assumed;
      do {
        assumed = old;
        const double cur = __longlong_as_double((long long)assumed);
        const double nv = (i == MEK_STAT_MIN_PHISE) ? fmin(cur, v) : fmax(cur, v);
        old = atomicCAS(addr, assumed, (unsigned long long)__double_as_longlong(nv));
      } while (assumed != old);
    } else {
      atomicAdd(totals + i, v);
    }
  }
}

// ---- initial cell order and heterogeneity test on the device (mek_pop_init) ----
// order-preserving 64-bit image of a double (NaN last): radix-sortable
__device__ __forceinline__ unsigned long long orderable(double x) {
  if (x != x) return ~0ull;
  if (x == 0.0) x = 0.0;                                             // -0 and +0 compare equal
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__global__ void k_order_keys(long long B, const double* v, const int* idx_in, unsigned long long* keys, int* idx) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (long long)gridDim.x * blockDim.x) {
    const int c = idx_in ? idx_in[i] : (int)i;
    keys[i] = orderable(v[c]);
    idx[i] = c;
  }
}
__global__ void k_heterogeneous(long long B, const double* Tc, const double* lim, int* flag) {
  bool het = false;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (long long)gridDim.x * blockDim.x) {
    if (Tc[i] != Tc[0]) het = true;
    for (int q = 0; q < 7; ++q)
      if (lim[7 * i + q] != lim[q]) het = true;
  }
  if (het) atomicOr(flag, 1);
}

// log read-back: the compact on-device records (u8 / u16 / u32 x words) widened to the ABI's int32 / uint32[4] columns
__global__ void k_widen_logs(size_t n, int W, const unsigned char* iT, const unsigned char* iZ, const unsigned short* nexec,
                             const unsigned* active, int* oT, int* oZ, int* on, unsigned* oa) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (oT) oT[i] = iT[i];
    if (oZ) oZ[i] = iZ[i];
    if (on) on[i] = nexec[i];
    if (oa) {
      uint4 a = make_uint4(0u, 0u, 0u, 0u);
      a.x = active[i * W];
      if (W > 1) a.y = active[i * W + 1];
      if (W > 2) a.z = active[i * W + 2];
      if (W > 3) a.w = active[i * W + 3];
      reinterpret_cast<uint4*>(oa)[i] = a;
    }
  }
}

// AoS <-> SoA transposition for mek_read_state / mek_write_state
__global__ void k_state_gather(const __grid_constant__ Params P, double* aos, int to_aos) {
  const Dims& d = P.d;
  const long long total = d.B * (long long)d.nfields;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long f = i / d.B, cell = i % d.B, id = cell_id(P, cell);
    double* slot = P.st + (cell >> 5) * ((long long)d.nfields * TILE) + f * TILE + (cell & 31);
    if (to_aos) aos[id * d.nfields + f] = *slot;
    else *slot = aos[id * d.nfields + f];
  }
}

// ---- dynamic regrouping: cells of a warp should be in the same phase of the charge ----
// mek_pop_init orders the cells by (T, SOC0) once; in a heterogeneous population (per-cell C-rate, limits, temperature)
// the cells of a warp drift apart as the charge proceeds: different local models (a wide hot set), different Hildreth
// sweep counts, different branches.  Between two launches of the fused kernel the state block is therefore re-sorted
// by (temperature bracket, present SOC, C-rate): a key per slot, a radix sort, one gather pass over the AoSoA tiles.
// Pure scheduling -- the caller's cell indices are kept by slot2cell.
__global__ void k_regroup_keys(const __grid_constant__ Params P, unsigned* keys, int* idx) {
  const Dims& d = P.d;
  const double* Tpts = P.bank_blob;                                  // blob layout: Tpts[nT] (K), Zpts[nZ], ...
  const double th0n = P.cp[P_NEG + E_THETA0], th100n = P.cp[P_NEG + E_THETA100];
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < d.B; s += (long long)gridDim.x * blockDim.x) {
    const double* b = P.st + (s >> 5) * ((long long)d.nfields * TILE) + (s & 31);
    const int tt = two_nearest(b[S_TC * TILE] + 273.15, Tpts, d.nT), t1 = tt & 0xffff, t2 = tt >> 16;
    const unsigned bracket = (unsigned)((t1 < t2 ? t1 : t2) * d.nT + (t1 < t2 ? t2 : t1)) & 0xffu;
    double soc = (b[S_SOCN * TILE] - th0n) / (th100n - th0n);
    soc = soc > 0.0 ? (soc < 1.2 ? soc : 1.2) : 0.0;                 // also maps NaN to 0
    const unsigned bin = (unsigned)(soc * 400.0);                    // 0.25 % SOC
    const double um = fabs(b[S_UMIN * TILE]);
    const unsigned rate = (unsigned)((um < 255.0 ? um : 255.0) * 4.0) & 0x3ffu;
    const unsigned lost = ((unsigned)b[S_STATUS * TILE] & ST_LOST) ? 1u : 0u;
    // which constraint blocks hold an active multiplier (constraintsMPC.m:99 [u; du; V; eta; SOC]): cells riding the same
    // limits walk the same rows of the Hildreth sweep and need similar numbers of sweeps
    unsigned act = 0u;
    if (b[S_LAMVALID * TILE] != 0.0) {
      const double* lam = b + (long long)d.f_lambda * TILE;
      int row = 0;
      const int len[5] = {d.useCurrent ? 2 * d.Nc : 0, d.useCurrent ? 2 * d.Nc : 0, d.useVoltage ? d.Np : 0, d.useEta ? d.Np : 0, d.useZmax ? d.Np : 0};
      for (int blk = 0; blk < 5; ++blk)
        for (int i = 0; i < len[blk]; ++i, ++row)
          if (!(lam[(long long)row * TILE] <= 0.0)) act |= 1u << blk;
    }
    keys[s] = (lost << 31) | ((bracket & 0x1fu) << 26) | (act << 21) | ((bin & 0x7ffu) << 10) | rate;
    idx[s] = (int)s;
  }
}

// st_new[slot j] = st_old[slot perm[j]] for every field; one warp per destination tile
__global__ void __launch_bounds__(256) k_regroup_move(const __grid_constant__ Params P, const int* __restrict__ perm, double* __restrict__ st_new,
                                                      const int* __restrict__ s2c_old, int* __restrict__ s2c_new) {
  const Dims& d = P.d;
  const long long ntiles = d.ld / TILE;
  const int lane = threadIdx.x & 31;
  const long long tsz = (long long)d.nfields * TILE;
  for (long long tile = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; tile < ntiles; tile += ((long long)gridDim.x * blockDim.x) >> 5) {
    const long long j = tile * TILE + lane;
    const bool live = j < d.B;
    const long long src = live ? (long long)perm[j] : 0;
    const double* q = P.st + (src >> 5) * tsz + (src & 31);
    double* o = st_new + tile * tsz + lane;
    if (live) s2c_new[j] = s2c_old[src];
    int f = 0;
    for (; f + 8 <= d.nfields; f += 8) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = live ? __ldcs(q + (long long)(f + u) * TILE) : 0.0;
#pragma unroll
      for (int u = 0; u < 8; ++u) __stcs(o + (long long)(f + u) * TILE, v[u]);
    }
    for (; f < d.nfields; ++f) __stcs(o + (long long)f * TILE, live ? __ldcs(q + (long long)f * TILE) : 0.0);
  }
}

// ------------------------------------------------------------------------------------------
// population (host object)
// ------------------------------------------------------------------------------------------
struct mek_pop {
  const mek_bank* bank;
  mek_config cfg;
  int device, sms, threads, blocks;
  size_t smem_bytes, blob_bytes;
  cudaStream_t stream;
  cudaEvent_t ev0, ev1;
  Params P;
  RowLists RL;
  double *d_blob, *d_Cfull, *d_Dfull, *d_st, *d_in, *d_totals, *d_scratch;
  double* d_lut;               // tabulated cell functions of the bank (null: closed forms)
  int* d_slot2cell;           // cell order inside the state block (sorted by initial SOC at mek_pop_init)
  size_t scratch_bytes;
  void* d_logs;
  bool inited;
  int64_t launches;
  int inst;     // kernel instance: 0 = <5,2>, 1 = <20,5>, 2 = <32,8>
  int minb;     // register cap variant of the <5,2> instance (CTAs per SM)
  int cons;     // >= 0: static-shape fast instance with these constraint switches
  size_t run_smem; int run_threads;
  int wide_ws_slots;           // > 0: doubles per thread of the long-horizon workspace in global memory
  double* d_wide_ws;
  // dynamic regrouping (k_regroup_*): second state block, key / permutation buffers, the order chosen at init
  int regroup_every;           // sample instants between regroupings inside mek_run (0: off)
  bool heterogeneous;          // set by mek_pop_init: temperatures or limits differ between cells
  double* d_st2;
  int *d_slot2cell0, *d_slot2cell2, *d_idx, *d_perm;
  unsigned *d_keys, *d_keys2;
  void* d_cubtmp; size_t cubtmp_bytes;
  int64_t regroups;
  int since_regroup;           // sample instants run since the last regrouping (or the initialisation): kept across mek_run calls
  const void* fn_run;          // entry point of the fused kernel instance in use (mek_entry.h)
};

static double sigma_min_gtg(int Np, int Nc, double r) {
  // smallest singular value of G_soc'*G_soc (iterMPC.m:41-43), a run constant: G(i,j) = r*(i-j), cyclic Jacobi
  double A[8][8];
  for (int j = 0; j < Nc; ++j)
    for (int l = 0; l < Nc; ++l) {
      double acc = 0.0;
      for (int i = 1; i <= Np; ++i) {
        const double gj = (i - 1 - j >= 0) ? r * (double)(i - 1 - j) : 0.0;
        const double gl = (i - 1 - l >= 0) ? r * (double)(i - 1 - l) : 0.0;
        acc = acc + gj * gl;
      }
      A[j][l] = acc;
    }
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < Nc; ++i)
      for (int j = i + 1; j < Nc; ++j) off += A[i][j] * A[i][j];
    if (off == 0.0) break;
    for (int pI = 0; pI < Nc - 1; ++pI)
      for (int q = pI + 1; q < Nc; ++q) {
        const double apq = A[pI][q];
        if (apq == 0.0) continue;
        const double theta = (A[q][q] - A[pI][pI]) / (2.0 * apq);
        double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
        if (theta < 0.0) t = -t;
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < Nc; ++k) { const double akp = A[k][pI], akq = A[k][q]; A[k][pI] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
        for (int k = 0; k < Nc; ++k) { const double apk = A[pI][k], aqk = A[q][k]; A[pI][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
      }
  }
  double mn = fabs(A[0][0]);
  for (int i = 1; i < Nc; ++i) if (fabs(A[i][i]) < mn) mn = fabs(A[i][i]);
  return mn;
}

template <typename K>
static int prep_kernel(K kern, size_t smem) {
  if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return MEK_OK;
}

static int grid_for(const mek_pop* p, long long B, int threads) {
  long long blocks = (B + threads - 1) / threads;
  const long long cap = (long long)p->blocks;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---- kernel launch / preparation through the entry points of mek_entry.h ----
static int launch_kernel(const void* fn, int grid, int threads, size_t smem, cudaStream_t st, void** args) {
  CUDA_TRY(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3((unsigned)threads), args, smem, st));
  return MEK_OK;
}
static int prep_fn(const void* fn, int threads, size_t smem, int* occ) {
  if (!fn) return fail(MEK_ERR_UNSUPPORTED, "kernel instance not built");
  if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, fn, threads, smem));
  return MEK_OK;
}

extern "C" void mek_pop_destroy(mek_pop* p);
// fills a zero-initialised mek_pop; on any failure the caller (mek_pop_create) hands the partly built object to
// mek_pop_destroy, which releases whatever was allocated so far
static int pop_build(mek_pop* p, const mek_bank* bank, const mek_config* cfg, int64_t B, int device) {
  memset(&p->P, 0, sizeof(Params));
  p->bank = bank;
  p->cfg = *cfg;
  p->device = device;
  p->inited = false;
  p->launches = 0;
  p->d_logs = nullptr;
  p->d_slot2cell = nullptr;
  p->d_wide_ws = nullptr;
  p->wide_ws_slots = 0;
  p->regroup_every = cfg->regroup_every;      // resolved in mek_pop_init (0 = automatic)
  p->heterogeneous = false;
  p->d_st2 = nullptr; p->d_slot2cell0 = nullptr; p->d_slot2cell2 = nullptr; p->d_idx = nullptr; p->d_perm = nullptr;
  p->d_keys = nullptr; p->d_keys2 = nullptr; p->d_cubtmp = nullptr; p->cubtmp_bytes = 0; p->regroups = 0;
  p->d_scratch = nullptr;
  p->scratch_bytes = 0;
  p->d_lut = nullptr;
  int Np = cfg->Np < 1 ? 1 : cfg->Np;                               // initMPC.m:32-33 (Q23)
  int Nc = cfg->Nc < 1 ? 1 : (cfg->Nc > Np ? Np : cfg->Nc);
  if (Np == 5 && Nc == 2) p->inst = 0;
  else if (Np <= 20 && Nc <= 5) p->inst = 1;
  else if (Np <= 32 && Nc <= 8) p->inst = 2;
  else { return fail(MEK_ERR_UNSUPPORTED, "Np > 32 or Nc > 8 is outside the built kernel instances"); }
  Dims& d = p->P.d;
  d.M = bank->M; d.nT = bank->nT; d.nZ = bank->nZ; d.nz = bank->nz;
  d.R = bank->R; d.nx_negtss = bank->nx_negtss; d.nx_postss = bank->nx_postss; d.nx_thetae = bank->nx_thetae;
  d.mstride = bank->mstride;
  d.Np = Np; d.Nc = Nc; d.maxHild = cfg->maxHild > 0 ? cfg->maxHild : 100;
  d.useCurrent = cfg->constraints[0] == 1; d.useVoltage = cfg->constraints[1] == 1; d.useEta = cfg->constraints[2] == 1;
  d.useZmax = cfg->use_zmax != 0;
  d.useTerm = cfg->z_target_terminal != 0;
  d.z_tol = cfg->z_tol;
  d.nC = (d.useCurrent ? 4 * Nc : 0) + Np * (d.useVoltage + d.useEta + d.useZmax) + d.useTerm;
  if (d.nC > 128) { return fail(MEK_ERR_UNSUPPORTED, "more than 128 constraint rows"); }
  d.force_jacobi = cfg->force_jacobi;
  d.no_lean_sweep = cfg->general_sweep_test == 1 ? 1 : 0;
  d.no_row_skip = cfg->general_sweep_test == 2 ? 1 : 0;
  d.Ts = bank->Tsamp; d.ref = cfg->ref; d.SigmaV = cfg->SigmaV; d.SigmaW = cfg->SigmaW;
  d.sigmin = sigma_min_gtg(Np, Nc, -bank->Tsamp / (3600.0 * bank->params[P_Q]));
  d.zinv = bank->nZ > 1 ? (double)(bank->nZ - 1) / (bank->Zpts[bank->nZ - 1] - bank->Zpts[0]) : 0.0;
  d.B = B;
  d.ld = (B + 31) / 32 * 32;
  d.f_lambda = S_HDR; d.f_bigx = d.f_lambda + d.nC; d.f_xhat = d.f_bigx + 6 * d.M; d.f_sig = d.f_xhat + N5 * d.M;
  d.nfields = d.f_sig + NSYM * d.M;
  d.meas_sigma = cfg->meas_sigma; d.noise_seed = cfg->noise_seed; d.first_cell = cfg->first_cell;
  d.plant_q_scale = cfg->plant_q_scale > 0.0 ? cfg->plant_q_scale : 1.0; d.plant_q_spread = cfg->plant_q_spread;
  d.plant_rc_scale = cfg->plant_rc_scale > 0.0 ? cfg->plant_rc_scale : 1.0;
  d.mb = cfg->ekf_method == 1 ? 1 : 0;                               // initKF.m:44-53: 'OutB' (0) or 'MdlB' (1)
  d.f_mbx = d.f_mbs = 0;
  if (d.mb) { d.f_mbx = d.nfields; d.f_mbs = d.f_mbx + N6; d.nfields = d.f_mbs + NSYM6; }
  memcpy(p->P.cp, bank->params.data(), sizeof(double) * N_PARAMS);
  // row lists for the full-zk corrections
  RowLists& RL = p->RL;
  memset(&RL, 0, sizeof(RL));
  auto fill = [](const std::vector<int>& v, int* dst, int& n) { n = (int)(v.size() > 64 ? 64 : v.size()); for (int i = 0; i < n; ++i) dst[i] = v[i]; };
  for (const std::vector<int>* v : {&bank->negtss, &bank->postss, &bank->thetae, &bank->negphise, &bank->posphise, &bank->phie, &bank->posphis})
    if (v->size() > 64) { return fail(MEK_ERR_UNSUPPORTED, "more than 64 output rows of one kind"); }
  fill(bank->negtss, RL.negtss, RL.n_negtss); fill(bank->postss, RL.postss, RL.n_postss); fill(bank->thetae, RL.thetae, RL.n_thetae);
  fill(bank->negphise, RL.negphise, RL.n_negphise); fill(bank->posphise, RL.posphise, RL.n_posphise);
  fill(bank->phie, RL.phie, RL.n_phie); fill(bank->posphis, RL.posphis, RL.n_posphis);
  RL.r_phise0 = bank->r_phise0;
  RL.r_ifdl0 = bank->r_ifdl0; RL.r_ifdl3 = bank->r_ifdl3; RL.r_if0 = bank->r_if0; RL.r_if3 = bank->r_if3;
  RL.r_tss0 = bank->r_tss0; RL.r_tss3 = bank->r_tss3; RL.r_the0 = bank->r_the0; RL.r_the3 = bank->r_the3; RL.r_phie_last = bank->r_phie_last; RL.r_phi = bank->r_phi;
  for (int r : bank->posphis) RL.kind[r] = 1;
  for (int r : bank->negtss) RL.kind[r] = 2;
  for (int r : bank->postss) RL.kind[r] = 3;
  for (int r : bank->negphise) RL.kind[r] = 4;
  for (int r : bank->posphise) RL.kind[r] = 5;
  for (int r : bank->phie) RL.kind[r] = 6;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  p->sms = prop.multiProcessorCount;
  p->threads = 128;   // the fused kernels are compiled for a fixed CTA size (their shared-memory columns use it as a constant stride)
  p->blob_bytes = bank->blob.size() * sizeof(double);
  p->smem_bytes = p->blob_bytes + sizeof(double) * (size_t)SC_TOTAL * (size_t)p->threads;   // hot bank + per-thread scratch columns
  if (p->smem_bytes > 200 * 1024) { return fail(MEK_ERR_UNSUPPORTED, "hot bank does not fit in shared memory"); }
  CUDA_TRY(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreate(&p->ev0));
  CUDA_TRY(cudaEventCreate(&p->ev1));
  if (bank->lut_n > 0) {                                             // tabulated cell functions: tables in global memory, descriptor behind the parameters
    const size_t lb = sizeof(double) * bank->lut.size();
    CUDA_TRY(cudaMalloc(&p->d_lut, lb));
    CUDA_TRY(cudaMemcpy(p->d_lut, bank->lut.data(), lb, cudaMemcpyHostToDevice));
    p->P.cp[P_LUT_N] = (double)bank->lut_n;
    p->P.cp[P_LUT_LO] = bank->lut_lo;
    p->P.cp[P_LUT_INVH] = (double)(bank->lut_n - 1) / (bank->lut_hi - bank->lut_lo);
    p->P.cp[P_LUT_PTR] = bits_to_double((unsigned long long)(uintptr_t)p->d_lut);
  }
  CUDA_TRY(cudaMalloc(&p->d_blob, p->blob_bytes));
  CUDA_TRY(cudaMemcpy(p->d_blob, bank->blob.data(), p->blob_bytes, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(&p->d_Cfull, bank->C.size() * sizeof(double)));
  CUDA_TRY(cudaMemcpy(p->d_Cfull, bank->C.data(), bank->C.size() * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(&p->d_Dfull, bank->D.size() * sizeof(double)));
  CUDA_TRY(cudaMemcpy(p->d_Dfull, bank->D.data(), bank->D.size() * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(&p->d_st, sizeof(double) * (size_t)d.nfields * (size_t)d.ld));
  CUDA_TRY(cudaMemset(p->d_st, 0, sizeof(double) * (size_t)d.nfields * (size_t)d.ld));
  CUDA_TRY(cudaMalloc(&p->d_in, sizeof(double) * (size_t)B * (3 + 36 + 7)));
  CUDA_TRY(cudaMalloc(&p->d_totals, sizeof(double) * MEK_NSTAT));
  p->P.st = p->d_st;
  p->P.bank_blob = p->d_blob;
  p->P.blob_doubles = (int)bank->blob.size();
  p->P.off_mdl = bank->off_mdl;
  p->P.Cfull = p->d_Cfull;
  p->P.Dfull = p->d_Dfull;
  p->P.in_soc0 = p->d_in;
  p->P.in_ekf = p->d_in + B;
  p->P.in_T = p->d_in + 2 * B;
  p->P.in_sigx0 = p->d_in + 3 * B;
  p->P.in_lim = p->d_in + 39 * B;
  p->P.nlog = 0;
  p->P.log_steps = 0;
  p->P.log_words = (d.nC + 31) / 32;
  p->P.sync_steps = cfg->sync_steps >= 0 ? 1 : 0;   // default on: the barrier keeps a CTA's warps on the same code (I-cache sharing)
  // opt in to the dynamic shared memory each kernel needs and size the persistent grid
  int rc = MEK_OK, occ = 0;
  rc |= prep_kernel(k_pop_init, p->smem_bytes);
  rc |= prep_kernel(k_ob_step, p->smem_bytes);
  rc |= prep_kernel(k_iter_ekf, p->smem_bytes);
  rc |= prep_kernel(k_iter_ekf_mb, p->smem_bytes);
  rc |= prep_kernel(k_run_population_mb<5, 2>, p->smem_bytes);
  rc |= prep_kernel(k_run_population_mb<20, 5>, p->smem_bytes);
  rc |= prep_kernel(k_run_population_mb<32, 8>, p->smem_bytes);
  p->minb = 1;
  p->cons = -1;
  p->run_threads = p->threads;
  p->run_smem = p->smem_bytes;
  // tabulated cell functions and the noisy-sensor scenario are compiled into the full-featured instances only (k_fused_fast_lut.cu,
  // k_fused_generic.cu); the lean ones (MEK_LUT_OFF) serve everything else
  const bool full = bank->lut_n > 0 || cfg->meas_sigma != 0.0 || d.plant_q_scale != 1.0 || d.plant_q_spread != 0.0 || d.plant_rc_scale != 1.0;
  {
    const int mask = (d.useCurrent ? 1 : 0) | (d.useVoltage ? 2 : 0) | (d.useEta ? 4 : 0) | (d.useZmax ? 8 : 0);
    if (p->inst == 0 && (mask == 15 || mask == 11) && !cfg->generic_path && !d.useTerm) {
      p->cons = mask;
      p->run_threads = (cfg->threads_per_block == 384 && !full) ? 384 : 256;   // 256 threads x 255 registers: fewest spills, least DRAM traffic
      const int slots = (mask == 15) ? FastShape<5, 2, 15>::SLOTS : FastShape<5, 2, 11>::SLOTS;
      p->run_smem = p->blob_bytes + sizeof(double) * (size_t)(slots > sc_slots(bank->R) ? slots : sc_slots(bank->R)) * (size_t)p->run_threads;
      if (p->run_smem > 227 * 1024) p->cons = -1, p->run_threads = p->threads, p->run_smem = p->smem_bytes;
    }
    if (p->inst == 1 && Np == 20 && Nc == 5 && mask == 15 && !cfg->generic_path && !d.useTerm && !full) {   // config C4: long-horizon static-shape instance
      using WS = WideShape<20, 5, 15>;
      p->cons = mask;
      // 256 threads (default): lambda in shared memory, the read-mostly columns in L2; 64 threads: every column in shared memory
      p->run_threads = cfg->threads_per_block == 64 ? 64 : 256;
      const int slots = p->run_threads == 64 ? WS::SLOTS : WS::LSLOTS;
      p->run_smem = p->blob_bytes + sizeof(double) * (size_t)(slots > sc_slots(bank->R) ? slots : sc_slots(bank->R)) * (size_t)p->run_threads;
      p->wide_ws_slots = p->run_threads == 64 ? 0 : WS::WSLOTS;
    }
  }
  // deferral blocks (temporal blocking of the model stream): history columns behind the scratch columns
  {
    int kb = cfg->defer_block == 0 ? MEK_DEFER_DEFAULT : cfg->defer_block;
    if (kb < 2 || d.M > 64) kb = 0;                  // the hot set is a 64-bit mask
    if (kb > 64) kb = 64;
    while (kb > 0 && p->run_smem + sizeof(double) * (size_t)(kb + 1) * (size_t)p->run_threads > 227 * 1024) kb /= 2;
    if (kb < 2) kb = 0;
    p->P.defer_k = kb;
    p->P.defer_stagger = cfg->defer_stagger > 0 ? 1 : 0;
    p->P.defer_narrow = cfg->defer_narrow_test > 0 ? 1 : 0;
    p->P.hist_off = (int)(p->run_smem / sizeof(double));
    if (kb > 0) p->run_smem += sizeof(double) * (size_t)(kb + 1) * (size_t)p->run_threads;
  }
  if (p->inst == 1 && p->cons == 15) p->fn_run = mek_fused_wide_kernel(p->run_threads);
  else if (p->cons >= 0) p->fn_run = full ? mek_fused_fast_lut_kernel(p->cons, p->run_threads) : mek_fused_fast_kernel(p->cons, p->run_threads);
  else p->fn_run = mek_fused_generic_kernel(p->inst);
  rc |= prep_fn(p->fn_run, p->run_threads, p->run_smem, &occ);
  if (rc != MEK_OK) { return rc; }
  if (occ < 1) occ = 1;
  if (cfg->blocks_per_sm == 1) occ = 1;
  p->blocks = p->sms * occ;
  if (p->wide_ws_slots > 0) {
    const size_t bytes = sizeof(double) * (size_t)p->wide_ws_slots * (size_t)p->run_threads * (size_t)p->blocks;
    CUDA_TRY(cudaMalloc(&p->d_wide_ws, bytes));
    CUDA_TRY(cudaMemset(p->d_wide_ws, 0, bytes));
    p->P.wide_ws = p->d_wide_ws;
  }
  return MEK_OK;
}

extern "C" int mek_pop_create(const mek_bank* bank, const mek_config* cfg, int64_t B, int device, mek_pop** out) {
  if (!bank || !cfg || !out || B < 1) return fail(MEK_ERR_ARG, "bad argument to mek_pop_create");
  int ndev = 0;
  cudaError_t e0 = cudaGetDeviceCount(&ndev);
  if (e0 != cudaSuccess || ndev < 1)
    return fail(MEK_ERR_CUDA, "no CUDA device: libmek has no CPU path (the oracle under oracle/ is test infrastructure only)");
  if (device < 0 || device >= ndev) return fail(MEK_ERR_ARG, "bad device index");
  CUDA_TRY(cudaSetDevice(device));
  mek_pop* p = new mek_pop();                  // value-initialised: every pointer / handle starts null
  const int rc = pop_build(p, bank, cfg, B, device);
  if (rc != MEK_OK) {
    const std::string msg = g_err;             // mek_pop_destroy must not lose the reason
    mek_pop_destroy(p);
    g_err = msg;
    return rc;
  }
  *out = p;
  return MEK_OK;
}

extern "C" void mek_pop_destroy(mek_pop* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  if (p->d_slot2cell) cudaFree(p->d_slot2cell);
  if (p->d_wide_ws) cudaFree(p->d_wide_ws);
  if (p->d_st2) cudaFree(p->d_st2);
  if (p->d_slot2cell0) cudaFree(p->d_slot2cell0);
  if (p->d_slot2cell2) cudaFree(p->d_slot2cell2);
  if (p->d_idx) cudaFree(p->d_idx);
  if (p->d_perm) cudaFree(p->d_perm);
  if (p->d_keys) cudaFree(p->d_keys);
  if (p->d_keys2) cudaFree(p->d_keys2);
  if (p->d_cubtmp) cudaFree(p->d_cubtmp);
  if (p->d_lut) cudaFree(p->d_lut);
  cudaFree(p->d_blob); cudaFree(p->d_Cfull); cudaFree(p->d_Dfull); cudaFree(p->d_st); cudaFree(p->d_in); cudaFree(p->d_totals);
  if (p->d_logs) cudaFree(p->d_logs);
  if (p->d_scratch) cudaFree(p->d_scratch);
  if (p->ev0) cudaEventDestroy(p->ev0);
  if (p->ev1) cudaEventDestroy(p->ev1);
  if (p->stream) cudaStreamDestroy(p->stream);
  delete p;
}

static int ensure_scratch(mek_pop* p, size_t bytes) {
  if (bytes <= p->scratch_bytes) return MEK_OK;
  if (p->d_scratch) CUDA_TRY(cudaFree(p->d_scratch));
  p->d_scratch = nullptr;
  p->scratch_bytes = 0;
  CUDA_TRY(cudaMalloc(&p->d_scratch, bytes));
  p->scratch_bytes = bytes;
  return MEK_OK;
}

extern "C" int mek_pop_set_logging(mek_pop* p, int64_t n_log_cells, int max_steps) {
  if (!p || n_log_cells < 0 || max_steps < 0) return fail(MEK_ERR_ARG, "bad argument to mek_pop_set_logging");
  CUDA_TRY(cudaSetDevice(p->device));
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  if (p->d_logs) { CUDA_TRY(cudaFree(p->d_logs)); p->d_logs = nullptr; }
  if (n_log_cells > p->P.d.B) n_log_cells = p->P.d.B;
  p->P.nlog = n_log_cells;
  p->P.log_steps = max_steps;
  const size_t n = (size_t)n_log_cells * (size_t)max_steps;
  if (n == 0) { p->P.nlog = 0; p->P.log_steps = 0; return MEK_OK; }
  const size_t bytes = n * (4 * 8 + 1 + 1 + 2 + 4 * (size_t)p->P.log_words) + 64;
  CUDA_TRY(cudaMalloc(&p->d_logs, bytes));
  CUDA_TRY(cudaMemset(p->d_logs, 0xFF, bytes));
  char* base = reinterpret_cast<char*>(p->d_logs);
  p->P.log_u = reinterpret_cast<double*>(base); base += n * 8;
  p->P.log_V = reinterpret_cast<double*>(base); base += n * 8;
  p->P.log_soc = reinterpret_cast<double*>(base); base += n * 8;
  p->P.log_phise = reinterpret_cast<double*>(base); base += n * 8;
  p->P.log_active = reinterpret_cast<unsigned*>(base); base += n * 4 * (size_t)p->P.log_words;
  p->P.log_nexec = reinterpret_cast<unsigned short*>(base); base += n * 2;
  p->P.log_iT = reinterpret_cast<unsigned char*>(base); base += n;
  p->P.log_iZ = reinterpret_cast<unsigned char*>(base);
  return MEK_OK;
}

extern "C" int mek_pop_reset(mek_pop* p) {
  if (!p) return fail(MEK_ERR_ARG, "null pop");
  CUDA_TRY(cudaSetDevice(p->device));
  if (p->d_slot2cell0 && p->P.slot2cell)      // regrouping has permuted the slots: go back to the order chosen at init
    CUDA_TRY(cudaMemcpyAsync(p->d_slot2cell, p->d_slot2cell0, sizeof(int) * (size_t)p->P.d.B, cudaMemcpyDeviceToDevice, p->stream));
  k_pop_init<<<grid_for(p, p->P.d.B, p->threads), p->threads, p->smem_bytes, p->stream>>>(p->P);
  CUDA_TRY(cudaGetLastError());
  p->launches++;
  p->since_regroup = 0;
  p->inited = true;
  return MEK_OK;
}

extern "C" int mek_pop_init(mek_pop* p, const double* SOC0pct, const double* ekfSOC0pct, const double* Tc, const double* SigmaX0,
                            const double* limits) {
  if (!p || !SOC0pct || !Tc || !SigmaX0 || !limits) return fail(MEK_ERR_ARG, "null argument to mek_pop_init");
  CUDA_TRY(cudaSetDevice(p->device));
  const size_t B = (size_t)p->P.d.B;
  // The reference treats a temperature above 100 as Kelvin in initKF.m:38-41 and iterEKF.m:62-66 but adds 273.15 to it
  // unconditionally in OB_step.m:75 and EKFmatsHandler.m:53: with such an input its three functions run at three
  // different temperatures.  libmek takes degrees Celsius only and says so instead of reproducing that.
  for (size_t i = 0; i < B; ++i)
    if (Tc[i] > 100.0) return fail(MEK_ERR_ARG, "Tc must be in degrees Celsius (<= 100); the reference's Kelvin heuristic (initKF.m:38-41, iterEKF.m:62-66) is not reproduced");
  double* in = p->d_in;
  CUDA_TRY(cudaMemcpyAsync(in, SOC0pct, 8 * B, cudaMemcpyHostToDevice, p->stream));
  CUDA_TRY(cudaMemcpyAsync(in + B, ekfSOC0pct ? ekfSOC0pct : SOC0pct, 8 * B, cudaMemcpyHostToDevice, p->stream));
  CUDA_TRY(cudaMemcpyAsync(in + 2 * B, Tc, 8 * B, cudaMemcpyHostToDevice, p->stream));
  CUDA_TRY(cudaMemcpyAsync(in + 3 * B, SigmaX0, 8 * 36 * B, cudaMemcpyHostToDevice, p->stream));
  CUDA_TRY(cudaMemcpyAsync(in + 39 * B, limits, 8 * 7 * B, cudaMemcpyHostToDevice, p->stream));
  // Order the cells inside the state block by initial SOC: the 32 cells of a warp are then in the same phase of the
  // charge (same local models, similar Hildreth sweep counts), which removes most of the warp divergence.  Pure
  // scheduling: every per-cell input/output of the ABI keeps the caller's index (slot2cell).
  if (!p->cfg.keep_cell_order) {
    // stable radix sort by SOC0, then by temperature (LSD): the order of a host stable_sort on (T, SOC0), without the host
    if (!p->d_slot2cell) CUDA_TRY(cudaMalloc(&p->d_slot2cell, sizeof(int) * B));
    {
      const size_t kb = (sizeof(unsigned long long) * B + 255) / 256 * 256, ib = (sizeof(int) * B + 255) / 256 * 256;
      size_t tmpb = 0;
      CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmpb, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int*)nullptr,
                                               (int*)nullptr, (int)B, 0, 64, p->stream));
      tmpb = (tmpb + 255) / 256 * 256;
      int rc0 = ensure_scratch(p, 2 * kb + 2 * ib + tmpb + 256);
      if (rc0 != MEK_OK) return rc0;
      char* base = reinterpret_cast<char*>(p->d_scratch);
      unsigned long long* k1 = reinterpret_cast<unsigned long long*>(base);
      unsigned long long* k2 = reinterpret_cast<unsigned long long*>(base + kb);
      int* i1 = reinterpret_cast<int*>(base + 2 * kb);
      int* i2 = reinterpret_cast<int*>(base + 2 * kb + ib);
      void* tmp = base + 2 * kb + 2 * ib;
      int* flag = reinterpret_cast<int*>(base + 2 * kb + 2 * ib + tmpb);
      const int g = p->sms * 4;
      k_order_keys<<<g, 256, 0, p->stream>>>((long long)B, in, nullptr, k1, i1);                   // by SOC0
      CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmpb, k1, k2, i1, i2, (int)B, 0, 64, p->stream));
      k_order_keys<<<g, 256, 0, p->stream>>>((long long)B, in + 2 * B, i2, k1, i1);                // then by T (stable)
      CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmpb, k1, k2, i1, p->d_slot2cell, (int)B, 0, 64, p->stream));
      CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(int), p->stream));
      k_heterogeneous<<<g, 256, 0, p->stream>>>((long long)B, in + 2 * B, in + 39 * B, flag);
      CUDA_TRY(cudaGetLastError());
      p->launches += 5;
      int hflag = 0;
      CUDA_TRY(cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
      CUDA_TRY(cudaStreamSynchronize(p->stream));
      p->P.slot2cell = p->d_slot2cell;
      // regrouping (automatic: only where it pays, i.e. when the cells do not all share temperature and limits)
      const bool het = hflag != 0;
      p->heterogeneous = het;
      int every = p->cfg.regroup_every;
      if (every == 0) every = het ? MEK_REGROUP_DEFAULT : -1;
      p->regroup_every = (every > 0 && !p->P.d.mb) ? every : 0;   // the model-blend loop is one plain kernel: nothing to regroup for
      if (p->regroup_every > 0) {
        if (!p->d_slot2cell0) CUDA_TRY(cudaMalloc(&p->d_slot2cell0, sizeof(int) * B));
        CUDA_TRY(cudaMemcpyAsync(p->d_slot2cell0, p->d_slot2cell, sizeof(int) * B, cudaMemcpyDeviceToDevice, p->stream));
      }
    }
  } else {
    p->regroup_every = 0;
    CUDA_TRY(cudaStreamSynchronize(p->stream));    // the caller's buffers are free to reuse when mek_pop_init returns, in both branches
  }
  return mek_pop_reset(p);
}

static int launch_segment(mek_pop* p, int k0, int k1) {
  if (p->P.d.mb) {                                                   // model-blend filter: its own (runtime-shape) loop kernel
    const int g = grid_for(p, p->P.d.B, p->threads);
    if (p->inst == 0) k_run_population_mb<5, 2><<<g, p->threads, p->smem_bytes, p->stream>>>(p->P, p->RL, k0, k1);
    else if (p->inst == 1) k_run_population_mb<20, 5><<<g, p->threads, p->smem_bytes, p->stream>>>(p->P, p->RL, k0, k1);
    else k_run_population_mb<32, 8><<<g, p->threads, p->smem_bytes, p->stream>>>(p->P, p->RL, k0, k1);
    CUDA_TRY(cudaGetLastError());
    p->launches++;
    return MEK_OK;
  }
  const int g = grid_for(p, p->P.d.B, p->run_threads);
  void* args[] = {&p->P, &k0, &k1};
  const int rc = launch_kernel(p->fn_run, g, p->run_threads, p->run_smem, p->stream, args);
  if (rc != MEK_OK) return rc;
  CUDA_TRY(cudaGetLastError());
  p->launches++;
  return MEK_OK;
}


// re-sort the state block by (temperature bracket, present SOC, C-rate); see k_regroup_keys
static int regroup(mek_pop* p) {
  const size_t B = (size_t)p->P.d.B;
  const size_t st_bytes = sizeof(double) * (size_t)p->P.d.nfields * (size_t)p->P.d.ld;
  if (!p->d_st2) {
    CUDA_TRY(cudaMalloc(&p->d_st2, st_bytes));
    CUDA_TRY(cudaMalloc(&p->d_slot2cell2, sizeof(int) * B));
    CUDA_TRY(cudaMalloc(&p->d_idx, sizeof(int) * B));
    CUDA_TRY(cudaMalloc(&p->d_perm, sizeof(int) * B));
    CUDA_TRY(cudaMalloc(&p->d_keys, sizeof(unsigned) * B));
    CUDA_TRY(cudaMalloc(&p->d_keys2, sizeof(unsigned) * B));
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, p->cubtmp_bytes, p->d_keys, p->d_keys2, p->d_idx, p->d_perm, (int)B, 0, 32, p->stream));
    CUDA_TRY(cudaMalloc(&p->d_cubtmp, p->cubtmp_bytes));
  }
  k_regroup_keys<<<p->sms * 4, 256, 0, p->stream>>>(p->P, p->d_keys, p->d_idx);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(p->d_cubtmp, p->cubtmp_bytes, p->d_keys, p->d_keys2, p->d_idx, p->d_perm, (int)B, 0, 32, p->stream));
  k_regroup_move<<<p->sms * 8, 256, 0, p->stream>>>(p->P, p->d_perm, p->d_st2, p->d_slot2cell, p->d_slot2cell2);
  CUDA_TRY(cudaGetLastError());
  std::swap(p->d_st, p->d_st2);
  std::swap(p->d_slot2cell, p->d_slot2cell2);
  p->P.st = p->d_st;
  p->P.slot2cell = p->d_slot2cell;
  p->launches += 3;            // keys, sort (one or more library kernels), move
  p->regroups++;
  return MEK_OK;
}

static int launch_run(mek_pop* p, int k0, int k1) {
  const int every = (p->regroup_every > 0 && p->P.slot2cell) ? p->regroup_every : 0;
  if (every == 0) return launch_segment(p, k0, k1);
  // The instants since the last regrouping are counted across calls, so a caller that runs the charge in chunks no longer
  // than regroup_every is regrouped exactly like one that runs it in a single call.
  for (int s = k0; s < k1;) {
    if (p->since_regroup >= every) {
      const int rc = regroup(p);
      if (rc != MEK_OK) return rc;
      p->since_regroup = 0;
    }
    const int room = every - p->since_regroup;
    const int e = s + room < k1 ? s + room : k1;
    const int rc = launch_segment(p, s, e);
    if (rc != MEK_OK) return rc;
    p->since_regroup += e - s;
    s = e;
  }
  return MEK_OK;
}

extern "C" int mek_run(mek_pop* p, int k0, int k1) {
  if (!p || k1 < k0 || k0 < 0) return fail(MEK_ERR_ARG, "bad argument to mek_run");
  if (!p->inited) return fail(MEK_ERR_STATE, "mek_run before mek_pop_init");
  CUDA_TRY(cudaSetDevice(p->device));
  return launch_run(p, k0, k1);
}

extern "C" int mek_run_timed(mek_pop* p, int k0, int k1, float* ms) {
  if (!p || k1 < k0 || k0 < 0) return fail(MEK_ERR_ARG, "bad argument to mek_run_timed");
  if (!p->inited) return fail(MEK_ERR_STATE, "mek_run before mek_pop_init");
  CUDA_TRY(cudaSetDevice(p->device));
  CUDA_TRY(cudaEventRecord(p->ev0, p->stream));
  int rc = launch_run(p, k0, k1);
  if (rc != MEK_OK) return rc;
  CUDA_TRY(cudaEventRecord(p->ev1, p->stream));
  CUDA_TRY(cudaEventSynchronize(p->ev1));
  if (ms) CUDA_TRY(cudaEventElapsedTime(ms, p->ev0, p->ev1));
  return MEK_OK;
}

extern "C" int mek_sync(mek_pop* p) {
  if (!p) return fail(MEK_ERR_ARG, "null pop");
  CUDA_TRY(cudaSetDevice(p->device));
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  return MEK_OK;
}
extern "C" void* mek_pop_stream(mek_pop* p) { return p ? (void*)p->stream : nullptr; }
extern "C" int64_t mek_pop_launches(const mek_pop* p) { return p ? p->launches : 0; }
extern "C" int mek_pop_ncon(const mek_pop* p) { return p ? p->P.d.nC : 0; }
extern "C" int mek_pop_defer_block(const mek_pop* p) { return p ? p->P.defer_k : 0; }
extern "C" int64_t mek_pop_regroups(const mek_pop* p) { return p ? p->regroups : 0; }
extern "C" int mek_pop_state_len(const mek_pop* p) { return p ? p->P.d.nfields : 0; }

extern "C" int mek_read_logs(mek_pop* p, int k0, int k1, double* u, double* V, double* soc, double* phise, int32_t* iT, int32_t* iZ,
                             int32_t* nexec, uint32_t* active) {
  if (!p || k0 < 0 || k1 < k0 || k1 > p->P.log_steps) return fail(MEK_ERR_ARG, "log range outside what mek_pop_set_logging reserved");
  CUDA_TRY(cudaSetDevice(p->device));
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  const size_t nl = (size_t)p->P.nlog, n = nl * (size_t)(k1 - k0), o = nl * (size_t)k0;
  if (n == 0) return MEK_OK;
  // doubles straight from the log block; the compact integer records are widened on the device first.  All copies are
  // queued on the population's stream and waited for once: into pinned memory (mek_host_alloc) they run at link speed
  if (iT || iZ || nexec || active) {
    const size_t ib = (4 * n + 255) / 256 * 256;
    int rc = ensure_scratch(p, 3 * ib + 16 * n);
    if (rc != MEK_OK) return rc;
    char* base = reinterpret_cast<char*>(p->d_scratch);
    int* dT = reinterpret_cast<int*>(base);
    int* dZ = reinterpret_cast<int*>(base + ib);
    int* dn = reinterpret_cast<int*>(base + 2 * ib);
    unsigned* da = reinterpret_cast<unsigned*>(base + 3 * ib);
    const int W = p->P.log_words;
    k_widen_logs<<<p->sms * 8, 256, 0, p->stream>>>(n, W, p->P.log_iT + o, p->P.log_iZ + o, p->P.log_nexec + o, p->P.log_active + o * W,
                                                  iT ? dT : nullptr, iZ ? dZ : nullptr, nexec ? dn : nullptr, active ? da : nullptr);
    CUDA_TRY(cudaGetLastError());
    p->launches++;
    if (iT) CUDA_TRY(cudaMemcpyAsync(iT, dT, 4 * n, cudaMemcpyDeviceToHost, p->stream));
    if (iZ) CUDA_TRY(cudaMemcpyAsync(iZ, dZ, 4 * n, cudaMemcpyDeviceToHost, p->stream));
    if (nexec) CUDA_TRY(cudaMemcpyAsync(nexec, dn, 4 * n, cudaMemcpyDeviceToHost, p->stream));
    if (active) CUDA_TRY(cudaMemcpyAsync(active, da, 16 * n, cudaMemcpyDeviceToHost, p->stream));
  }
  if (u) CUDA_TRY(cudaMemcpyAsync(u, p->P.log_u + o, 8 * n, cudaMemcpyDeviceToHost, p->stream));
  if (V) CUDA_TRY(cudaMemcpyAsync(V, p->P.log_V + o, 8 * n, cudaMemcpyDeviceToHost, p->stream));
  if (soc) CUDA_TRY(cudaMemcpyAsync(soc, p->P.log_soc + o, 8 * n, cudaMemcpyDeviceToHost, p->stream));
  if (phise) CUDA_TRY(cudaMemcpyAsync(phise, p->P.log_phise + o, 8 * n, cudaMemcpyDeviceToHost, p->stream));
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  return MEK_OK;
}

// page-locked host memory for the caller's input / result buffers (transfers then run at link speed and asynchronously)
extern "C" int mek_host_alloc(size_t bytes, void** out) {
  if (!out || bytes == 0) return fail(MEK_ERR_ARG, "bad argument to mek_host_alloc");
  CUDA_TRY(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
  return MEK_OK;
}
extern "C" int mek_host_free(void* ptr) {
  if (ptr) CUDA_TRY(cudaFreeHost(ptr));
  return MEK_OK;
}

static int state_xfer(mek_pop* p, double* host, int to_host) {
  CUDA_TRY(cudaSetDevice(p->device));
  const size_t bytes = sizeof(double) * (size_t)p->P.d.B * (size_t)p->P.d.nfields;
  int rc = ensure_scratch(p, bytes);
  if (rc != MEK_OK) return rc;
  if (!to_host) CUDA_TRY(cudaMemcpyAsync(p->d_scratch, host, bytes, cudaMemcpyHostToDevice, p->stream));
  k_state_gather<<<p->sms * 4, 256, 0, p->stream>>>(p->P, p->d_scratch, to_host);
  CUDA_TRY(cudaGetLastError());
  p->launches++;
  if (to_host) CUDA_TRY(cudaMemcpyAsync(host, p->d_scratch, bytes, cudaMemcpyDeviceToHost, p->stream));
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  return MEK_OK;
}
extern "C" int mek_read_state(mek_pop* p, double* state) {
  if (!p || !state) return fail(MEK_ERR_ARG, "null argument");
  return state_xfer(p, state, 1);
}
extern "C" int mek_write_state(mek_pop* p, const double* state) {
  if (!p || !state) return fail(MEK_ERR_ARG, "null argument");
  int rc = state_xfer(p, const_cast<double*>(state), 0);
  if (rc == MEK_OK) p->inited = true;
  return rc;
}

// per-cell summary columns in the caller's cell order (population analytics: histograms of charge time, min phise, ...)
__global__ void k_cell_summaries(const __grid_constant__ Params P, double* out) {
  const Dims& d = P.d;
  const long long total = d.B * (long long)MEK_NCELLSUM;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long j = i / d.B, cell = i % d.B, id = cell_id(P, cell);   // consecutive threads: consecutive slots of one field
    const int f = j == 0 ? S_STATUS : (j == 1 ? S_WARN : S_SUM_NEXEC + (int)(j - 2));
    out[id * MEK_NCELLSUM + j] = P.st[(cell >> 5) * ((long long)d.nfields * TILE) + (long long)f * TILE + (cell & 31)];
  }
}

extern "C" int mek_read_cell_summaries(mek_pop* p, double* out) {
  if (!p || !out) return fail(MEK_ERR_ARG, "null argument");
  if (!p->inited) return fail(MEK_ERR_STATE, "mek_read_cell_summaries before mek_pop_init");
  CUDA_TRY(cudaSetDevice(p->device));
  const size_t bytes = sizeof(double) * (size_t)p->P.d.B * MEK_NCELLSUM;
  int rc = ensure_scratch(p, bytes);
  if (rc != MEK_OK) return rc;
  k_cell_summaries<<<p->sms * 4, 256, 0, p->stream>>>(p->P, p->d_scratch);
  CUDA_TRY(cudaGetLastError());
  p->launches++;
  CUDA_TRY(cudaMemcpyAsync(out, p->d_scratch, bytes, cudaMemcpyDeviceToHost, p->stream));
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  return MEK_OK;
}

extern "C" int mek_summary(mek_pop* p, double* totals) {
  if (!p || !totals) return fail(MEK_ERR_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(p->device));
  double init[MEK_NSTAT];
  for (int i = 0; i < MEK_NSTAT; ++i) init[i] = 0.0;
  init[MEK_STAT_MIN_PHISE] = INFINITY;
  init[MEK_STAT_MAX_V] = -INFINITY;
  CUDA_TRY(cudaMemcpyAsync(p->d_totals, init, sizeof(init), cudaMemcpyHostToDevice, p->stream));
  k_summary<<<p->sms * 2, 256, 0, p->stream>>>(p->P, p->d_totals);
  CUDA_TRY(cudaGetLastError());
  p->launches++;
  CUDA_TRY(cudaMemcpyAsync(totals, p->d_totals, sizeof(init), cudaMemcpyDeviceToHost, p->stream));
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  return MEK_OK;
}

// ---- function-level entry points ----
extern "C" int mek_pop_ob_step(mek_pop* p, const double* Iapp, double* Vcell) {
  if (!p || !Iapp || !Vcell) return fail(MEK_ERR_ARG, "null argument");
  if (!p->inited) return fail(MEK_ERR_STATE, "First call requires mek_pop_init (OB_step.m:39-42)");
  CUDA_TRY(cudaSetDevice(p->device));
  const size_t B = (size_t)p->P.d.B;
  int rc = ensure_scratch(p, 16 * B);
  if (rc != MEK_OK) return rc;
  double* dI = p->d_scratch;
  double* dV = p->d_scratch + B;
  CUDA_TRY(cudaMemcpyAsync(dI, Iapp, 8 * B, cudaMemcpyHostToDevice, p->stream));
  k_ob_step<<<grid_for(p, p->P.d.B, p->threads), p->threads, p->smem_bytes, p->stream>>>(p->P, dI, dV);
  CUDA_TRY(cudaGetLastError());
  p->launches++;
  CUDA_TRY(cudaMemcpyAsync(Vcell, dV, 8 * B, cudaMemcpyDeviceToHost, p->stream));
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  return MEK_OK;
}

extern "C" int mek_pop_iter_ekf(mek_pop* p, const double* vk, const double* ik, double* zk, double* boundzk, double* mats) {
  if (!p || !vk || !ik) return fail(MEK_ERR_ARG, "null argument");
  if (!p->inited) return fail(MEK_ERR_STATE, "mek_pop_iter_ekf before mek_pop_init");
  CUDA_TRY(cudaSetDevice(p->device));
  const size_t B = (size_t)p->P.d.B, nzk = (size_t)p->P.d.nz + 2;
  if (boundzk && !zk) return fail(MEK_ERR_ARG, "boundzk needs zk");
  int rc = ensure_scratch(p, 8 * B * (2 + 2 * nzk + MEK_NMATS));
  if (rc != MEK_OK) return rc;
  double* dv = p->d_scratch;
  double* di = dv + B;
  double* dz = di + B;
  double* db = dz + B * nzk;
  double* dm = db + B * nzk;
  CUDA_TRY(cudaMemcpyAsync(dv, vk, 8 * B, cudaMemcpyHostToDevice, p->stream));
  CUDA_TRY(cudaMemcpyAsync(di, ik, 8 * B, cudaMemcpyHostToDevice, p->stream));
  if (p->P.d.mb) k_iter_ekf_mb<<<grid_for(p, p->P.d.B, p->threads), p->threads, p->smem_bytes, p->stream>>>(p->P, p->RL, dv, di, zk ? dz : nullptr, dm, boundzk ? db : nullptr);
  else k_iter_ekf<<<grid_for(p, p->P.d.B, p->threads), p->threads, p->smem_bytes, p->stream>>>(p->P, p->RL, dv, di, zk ? dz : nullptr, dm, boundzk ? db : nullptr);
  CUDA_TRY(cudaGetLastError());
  p->launches++;
  if (zk) CUDA_TRY(cudaMemcpyAsync(zk, dz, 8 * B * nzk, cudaMemcpyDeviceToHost, p->stream));
  if (boundzk) CUDA_TRY(cudaMemcpyAsync(boundzk, db, 8 * B * nzk, cudaMemcpyDeviceToHost, p->stream));
  if (mats) CUDA_TRY(cudaMemcpyAsync(mats, dm, 8 * B * MEK_NMATS, cudaMemcpyDeviceToHost, p->stream));
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  return MEK_OK;
}

extern "C" int mek_pop_iter_mpc(mek_pop* p, const double* mats, const double* uk_applied, double* uk, double* phise, int32_t* nexec,
                                uint32_t* active) {
  return mek_pop_iter_mpc_cost(p, mats, uk_applied, uk, phise, nexec, active, nullptr);
}

extern "C" int mek_pop_iter_mpc_cost(mek_pop* p, const double* mats, const double* uk_applied, double* uk, double* phise, int32_t* nexec,
                                     uint32_t* active, double* cost) {
  if (!p || !mats || !uk_applied || !uk) return fail(MEK_ERR_ARG, "null argument");
  if (!p->inited) return fail(MEK_ERR_STATE, "mek_pop_iter_mpc before mek_pop_init");
  CUDA_TRY(cudaSetDevice(p->device));
  const size_t B = (size_t)p->P.d.B;
  int rc = ensure_scratch(p, 8 * B * (MEK_NMATS + 3 + MEK_NCOST) + 4 * B * 5);
  if (rc != MEK_OK) return rc;
  double* dm = p->d_scratch;
  double* du_in = dm + B * MEK_NMATS;
  double* du = du_in + B;
  double* dph = du + B;
  double* dc = dph + B;
  int* dn = reinterpret_cast<int*>(dc + B * MEK_NCOST);
  unsigned* da = reinterpret_cast<unsigned*>(dn + B);
  CUDA_TRY(cudaMemcpyAsync(dm, mats, 8 * B * MEK_NMATS, cudaMemcpyHostToDevice, p->stream));
  CUDA_TRY(cudaMemcpyAsync(du_in, uk_applied, 8 * B, cudaMemcpyHostToDevice, p->stream));
  const int g = grid_for(p, p->P.d.B, p->threads);
  if (p->inst == 0) k_iter_mpc<5, 2><<<g, p->threads, 0, p->stream>>>(p->P, dm, du_in, du, dph, dn, da, cost ? dc : nullptr);
  else if (p->inst == 1) k_iter_mpc<20, 5><<<g, p->threads, 0, p->stream>>>(p->P, dm, du_in, du, dph, dn, da, cost ? dc : nullptr);
  else k_iter_mpc<32, 8><<<g, p->threads, 0, p->stream>>>(p->P, dm, du_in, du, dph, dn, da, cost ? dc : nullptr);
  CUDA_TRY(cudaGetLastError());
  p->launches++;
  CUDA_TRY(cudaMemcpyAsync(uk, du, 8 * B, cudaMemcpyDeviceToHost, p->stream));
  if (phise) CUDA_TRY(cudaMemcpyAsync(phise, dph, 8 * B, cudaMemcpyDeviceToHost, p->stream));
  if (nexec) CUDA_TRY(cudaMemcpyAsync(nexec, dn, 4 * B, cudaMemcpyDeviceToHost, p->stream));
  if (active) CUDA_TRY(cudaMemcpyAsync(active, da, 16 * B, cudaMemcpyDeviceToHost, p->stream));
  if (cost) CUDA_TRY(cudaMemcpyAsync(cost, dc, 8 * B * MEK_NCOST, cudaMemcpyDeviceToHost, p->stream));
  CUDA_TRY(cudaStreamSynchronize(p->stream));
  return MEK_OK;
}

// ---- stateless entry points ----
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
};

static int use_device(int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return fail(MEK_ERR_CUDA, "no CUDA device: libmek has no CPU path (the oracle under oracle/ is test infrastructure only)");
  if (device < 0 || device >= ndev) return fail(MEK_ERR_ARG, "bad device index");
  CUDA_TRY(cudaSetDevice(device));
  return MEK_OK;
}

extern "C" int mek_hildreth(int64_t B, int Nc, int nC, const double* E, const double* F, const double* M, const double* gamma,
                            const double* lambda0, int maxIter, double* DU, double* lambda, int32_t* nexec, int device) {
  if (B < 1 || Nc < 1 || nC < 1 || !E || !F || !M || !gamma || !DU || !lambda || !nexec) return fail(MEK_ERR_ARG, "bad argument to mek_hildreth");
  if (Nc > 8 || nC > 128) return fail(MEK_ERR_UNSUPPORTED, "mek_hildreth supports Nc <= 8, nC <= 128");
  int rc = use_device(device);
  if (rc != MEK_OK) return rc;
  const size_t b = (size_t)B;
  const size_t nin = b * ((size_t)Nc * Nc + Nc + (size_t)nC * Nc + nC + (lambda0 ? nC : 0));
  const size_t nout = b * ((size_t)Nc + nC);
  DevBuf buf;
  CUDA_TRY(cudaMalloc(&buf.p, 8 * (nin + nout) + 4 * b));
  double* dE = reinterpret_cast<double*>(buf.p);
  double* dF = dE + b * Nc * Nc;
  double* dM = dF + b * Nc;
  double* dg = dM + b * nC * Nc;
  double* dl0 = dg + b * nC;
  double* dDU = dl0 + (lambda0 ? b * nC : 0);
  double* dlam = dDU + b * Nc;
  int* dn = reinterpret_cast<int*>(dlam + b * nC);
  CUDA_TRY(cudaMemcpy(dE, E, 8 * b * Nc * Nc, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dF, F, 8 * b * Nc, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dM, M, 8 * b * nC * Nc, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dg, gamma, 8 * b * nC, cudaMemcpyHostToDevice));
  if (lambda0) CUDA_TRY(cudaMemcpy(dl0, lambda0, 8 * b * nC, cudaMemcpyHostToDevice));
  const int threads = 128;
  long long blocks = (B + threads - 1) / threads;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (Nc <= 2 && nC <= 23) k_hildreth<2, 23><<<(int)blocks, threads>>>(B, Nc, nC, dE, dF, dM, dg, lambda0 ? dl0 : nullptr, maxIter, dDU, dlam, dn);
  else if (Nc <= 5 && nC <= 80) k_hildreth<5, 80><<<(int)blocks, threads>>>(B, Nc, nC, dE, dF, dM, dg, lambda0 ? dl0 : nullptr, maxIter, dDU, dlam, dn);
  else k_hildreth<8, 128><<<(int)blocks, threads>>>(B, Nc, nC, dE, dF, dM, dg, lambda0 ? dl0 : nullptr, maxIter, dDU, dlam, dn);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(DU, dDU, 8 * b * Nc, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(lambda, dlam, 8 * b * nC, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(nexec, dn, 4 * b, cudaMemcpyDeviceToHost));
  return MEK_OK;
}

extern "C" int mek_pred_mat(int64_t B, const double* a, const double* c, const double* d, int Np, int Nc, double* Phi, double* G, int device) {
  if (B < 1 || !a || !c || !d || !Phi || !G) return fail(MEK_ERR_ARG, "bad argument to mek_pred_mat");
  Np = Np < 1 ? 1 : Np;                                             // predMat.m:15-16 (Q23)
  Nc = Nc < 1 ? 1 : (Nc > Np ? Np : Nc);
  if (Np > 32) return fail(MEK_ERR_UNSUPPORTED, "mek_pred_mat supports Np <= 32");
  int rc = use_device(device);
  if (rc != MEK_OK) return rc;
  const size_t b = (size_t)B;
  DevBuf buf;
  CUDA_TRY(cudaMalloc(&buf.p, 8 * b * (6 + 6 + 1 + (size_t)Np * 7 + (size_t)Np * Nc)));
  double* da = reinterpret_cast<double*>(buf.p);
  double* dc = da + 6 * b;
  double* dd = dc + 6 * b;
  double* dPhi = dd + b;
  double* dG = dPhi + b * Np * 7;
  CUDA_TRY(cudaMemcpy(da, a, 8 * 6 * b, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dc, c, 8 * 6 * b, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dd, d, 8 * b, cudaMemcpyHostToDevice));
  const int threads = 128;
  long long blocks = (B + threads - 1) / threads;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_pred_mat<<<(int)blocks, threads>>>(B, da, dc, dd, Np, Nc, dPhi, dG);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(Phi, dPhi, 8 * b * Np * 7, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(G, dG, 8 * b * Np * Nc, cudaMemcpyDeviceToHost));
  return MEK_OK;
}


extern "C" int mek_constraints(int64_t B, int Np, int Nc, const int* constraints, int use_zmax, const double* x_aug, const double* mats,
                               const double* SOCk_1, const double* limits, double* M, double* gamma, double r, double ref, double* E,
                               double* F, double* DU, int device) {
  if (B < 1 || !constraints || !x_aug || !mats || !SOCk_1 || !limits || !M || !gamma) return fail(MEK_ERR_ARG, "bad argument to mek_constraints");
  Np = Np < 1 ? 1 : Np;
  Nc = Nc < 1 ? 1 : (Nc > Np ? Np : Nc);
  if (Np > 32 || Nc > 8) return fail(MEK_ERR_UNSUPPORTED, "mek_constraints supports Np <= 32, Nc <= 8");
  int rc = use_device(device);
  if (rc != MEK_OK) return rc;
  Dims d;
  memset(&d, 0, sizeof(d));
  d.Np = Np; d.Nc = Nc; d.Ts = 1.0; d.ref = ref;
  d.useCurrent = constraints[0] == 1; d.useVoltage = constraints[1] == 1; d.useEta = constraints[2] == 1; d.useZmax = (use_zmax & 1) != 0;
  d.useTerm = (use_zmax & 2) != 0;
  d.nC = (d.useCurrent ? 4 * Nc : 0) + Np * (d.useVoltage + d.useEta + d.useZmax) + d.useTerm;
  if (d.nC < 1 || d.nC > 128) return fail(MEK_ERR_UNSUPPORTED, "constraint row count outside 1..128");
  d.sigmin = sigma_min_gtg(Np, Nc, r);
  const size_t b = (size_t)B, nC = (size_t)d.nC;
  DevBuf buf;
  const size_t nin = b * (7 + MEK_NMATS + 1 + 7), nout = b * (nC * Nc + nC + (size_t)Nc * Nc + 2 * (size_t)Nc);
  CUDA_TRY(cudaMalloc(&buf.p, 8 * (nin + nout)));
  double* dx = reinterpret_cast<double*>(buf.p);
  double* dm = dx + 7 * b;
  double* ds = dm + MEK_NMATS * b;
  double* dl = ds + b;
  double* dM = dl + 7 * b;
  double* dg = dM + b * nC * Nc;
  double* dE = dg + b * nC;
  double* dF = dE + b * Nc * Nc;
  double* dDU = dF + b * Nc;
  CUDA_TRY(cudaMemcpy(dx, x_aug, 8 * 7 * b, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dm, mats, 8 * MEK_NMATS * b, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(ds, SOCk_1, 8 * b, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dl, limits, 8 * 7 * b, cudaMemcpyHostToDevice));
  const int threads = 128;
  long long blocks = (B + threads - 1) / threads;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (Np <= 5 && Nc <= 2) k_constraints<5, 2><<<(int)blocks, threads>>>(B, d, r, dx, dm, ds, dl, dM, dg, dE, dF, dDU);
  else if (Np <= 20 && Nc <= 5) k_constraints<20, 5><<<(int)blocks, threads>>>(B, d, r, dx, dm, ds, dl, dM, dg, dE, dF, dDU);
  else k_constraints<32, 8><<<(int)blocks, threads>>>(B, d, r, dx, dm, ds, dl, dM, dg, dE, dF, dDU);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(M, dM, 8 * b * nC * Nc, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(gamma, dg, 8 * b * nC, cudaMemcpyDeviceToHost));
  if (E) CUDA_TRY(cudaMemcpy(E, dE, 8 * b * Nc * Nc, cudaMemcpyDeviceToHost));
  if (F) CUDA_TRY(cudaMemcpy(F, dF, 8 * b * Nc, cudaMemcpyDeviceToHost));
  if (DU) CUDA_TRY(cudaMemcpy(DU, dDU, 8 * b * Nc, cudaMemcpyDeviceToHost));
  return MEK_OK;
}


// ---- FP64 FMA microbenchmark: the measured denominator of the FP64-pipe roofline ----
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters) {
  double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3, a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3,
         a7 = a0 + 7e-3;
  const double m = 0.999999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

extern "C" int mek_fp64_peak(int device, double* tflops) {
  if (!tflops) return fail(MEK_ERR_ARG, "null argument");
  int rc = use_device(device);
  if (rc != MEK_OK) return rc;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
  DevBuf buf;
  CUDA_TRY(cudaMalloc(&buf.p, sizeof(double) * (size_t)blocks * threads));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CUDA_TRY(cudaEventRecord(e0));
    k_fp64_peak<<<blocks, threads>>>(reinterpret_cast<double*>(buf.p), iters);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    const double tf = 2.0 * 8.0 * (double)iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = best;
  return MEK_OK;
}

// ---- self-test of the speculative quotient / root (mek_math.cuh) against the standard operators ----
__global__ void k_selftest_spec(long long n, const double* x, const double* y, double* qs, double* qd, double* ss, double* sd, int* fast) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double a = x[i], b = y[i];
    bool okq = false, oks = false;
    double q = a / b, s = sqrt(a);
#ifdef MEK_SPEC_ON                                                     // device pass of a build with the speculative sequences
    okq = true; oks = true;
    q = mek::div_spec(a, b, okq); s = mek::sqrt_spec(a, oks);
    if (!okq) q = a / b;
    if (!oks) s = sqrt(a);
#endif
    qs[i] = q; ss[i] = s;
    qd[i] = a / b; sd[i] = sqrt(a);
    fast[i] = (okq ? 1 : 0) | (oks ? 2 : 0);
  }
}
extern "C" int mek_selftest_spec(int device, int64_t n, const double* x, const double* y, double* q_spec, double* q_std, double* s_spec,
                                 double* s_std, int* fast) {
  if (n <= 0 || !x || !y || !q_spec || !q_std || !s_spec || !s_std || !fast) return fail(MEK_ERR_ARG, "bad argument to mek_selftest_spec");
  int rc = use_device(device);
  if (rc != MEK_OK) return rc;
  DevBuf buf;
  const size_t b = (size_t)n;
  CUDA_TRY(cudaMalloc(&buf.p, 8 * 6 * b + 4 * b));
  double* dx = reinterpret_cast<double*>(buf.p);
  double *dy = dx + b, *dqs = dy + b, *dqd = dqs + b, *dss = dqd + b, *dsd = dss + b;
  int* df = reinterpret_cast<int*>(dsd + b);
  CUDA_TRY(cudaMemcpy(dx, x, 8 * b, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dy, y, 8 * b, cudaMemcpyHostToDevice));
  k_selftest_spec<<<148 * 4, 256>>>((long long)n, dx, dy, dqs, dqd, dss, dsd, df);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(q_spec, dqs, 8 * b, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(q_std, dqd, 8 * b, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(s_spec, dss, 8 * b, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(s_std, dsd, 8 * b, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(fast, df, 4 * b, cudaMemcpyDeviceToHost));
  return MEK_OK;
}

extern "C" const char* mek_last_error(void) { return g_err.c_str(); }
extern "C" int mek_version(void) { return 100; }
