// Device functions of the fused MPC-EKF control loop: one thread per cell, state in a
// structure-of-arrays block in HBM (cell index fastest => every warp access is one coalesced
// 256-byte line), the hot rows of the ROM table bank in shared memory.
//
// Each function cites the reference lines it implements (paths under /root/reference/src).
// Arithmetic order is fixed and FMA contraction is disabled (-fmad=false): see DESIGN.md
// "numerics contract".
#pragma once
#include <cstdint>
#include <type_traits>

#include "mek_cellfn.cuh"

namespace mek {

// ---- per-cell state fields (value of field f for cell c lives at st[c/32][f][c%32]) ----
enum : int {
  S_SOCN = 0, S_SOCP, S_SOC0N, S_SOC0P, S_CDLN, S_CDLP, S_TC, S_ARRK0N, S_ARRK0P, S_ARRRFN, S_ARRRFP,
  S_X0, S_SIGX0, S_PRIORI, S_EKFSOC0, S_WARN, S_UK, S_LAMVALID, S_STATUS, S_PLANTWARN,
  S_UMAX, S_UMIN, S_DUMAX, S_DUMIN, S_VMAX, S_PHIMIN, S_ZMAX, S_NJAC, S_NBUMP, S_SPARE,
  S_SUM_NEXEC, S_N_QP, S_N_CAP, S_MIN_PHISE, S_MAX_V, S_T_REACH, S_STEPS, S_LAST_SOC, S_LAST_V, S_LAST_PHISE,
  S_HDR = 40
};
constexpr unsigned ST_LOST = 1u, ST_Q18 = 2u, ST_NONFINITE = 4u;
constexpr int N5 = 5, NSYM = 15;

// ---- hot rows of the bank kept in shared memory (subset of ROMmdls(iT,iZ).C/.D the control path reads) ----
enum : int { HR_IFDL0 = 0, HR_IFDL3, HR_IF0, HR_IF3, HR_TSS0, HR_TSS3, HR_THE0, HR_THE3, HR_PHIE, HR_PHI, HR_FIXED = 10 };
constexpr int HR_MAX = 24;
// One hot row of a local model in the shared-memory bank: 8 doubles, 16-byte aligned, [C(1..5) | D | C(6) | pad], so that the
// EKF's row (five transient states + feed-through) is three 16-byte loads and the plant's (all six states) four.
constexpr int HRS = 8, HR_D = 5, HR_C6 = 6;
// A model of the hot bank: diag A (6), the products a_i*a_j of the covariance time update (upper triangle, 15 + 1 pad), then the rows.
constexpr int HB_AA = 6, HB_ROWS = HB_AA + 16;

struct Dims {
  int M, nT, nZ, nz;
  int R;             // hot rows per model (HR_FIXED + extras)
  int nx_negtss;     // extra negThetass rows (beyond Thetass0), then extra posThetass, then extra Thetae
  int nx_postss;
  int nx_thetae;
  int mstride;       // doubles per model in the hot bank: 6 (diag A) + R*HRS
  int nC;            // constraint rows
  int Np, Nc, maxHild;
  int useCurrent, useVoltage, useEta, useZmax;
  int useTerm;       // mpcData.const.z_target_terminal (constraintsMPC.m:104-112): one more row; generic instances only
  double z_tol;      // mpcData.const.z_tol: the per-cell z_max limit carries z_max + z_tol, the terminal row wants z_max
  int force_jacobi;
  int no_lean_sweep;  // test hook: static-shape instances always take the general Hildreth sweep
  int no_row_skip;    // test hook: the lean sweep never skips the tail of an idle row
  double Ts, ref, SigmaV, SigmaW, sigmin;
  double zinv;       // (nZ-1)/(Zpts[nZ-1]-Zpts[0]): bracket guess of two_nearest_z
  long long B, ld;   // cells in this shard, cells rounded up to whole 32-cell tiles
  int f_lambda, f_bigx, f_xhat, f_sig, nfields;
  double meas_sigma;            // noisy-sensor scenario (mek_config.meas_sigma): 0 = runMPC.m:88 as written
  double plant_q_scale, plant_q_spread, plant_rc_scale;   // plant-model mismatch scenario (1, 0, 1 = the reference)
  unsigned long long noise_seed;
  long long first_cell;
  int mb, f_mbx, f_mbs;   // model-blend EKF (function-level entry only): ekfData.xhat (6) and ekfData.SigmaX (21, upper triangle) behind the model rows
};

struct BankSmem {
  const double* Tpts;  // [nT] K
  const double* Zpts;  // [nZ]
  const double* mdl;   // [M][mstride]
};

// State block layout: tiles of 32 cells, st[tile][field][lane] (AoSoA).  A warp owns one tile, so every
// field access is one coalesced 256-byte request and a static field index is an immediate offset.
constexpr int TILE = 32;
struct Cell {
  double* base;          // &st[tile][0][lane]
  const Dims* dd;        // where the arrays of a cell start (field offsets: kernel parameters, i.e. constant-bank operands)
  __device__ __forceinline__ double& f(int field) const { return base[field * TILE]; }
  // one pointer per cell instead of five: the others were spilled and re-read from the stack all over the instant
  __device__ __forceinline__ double& bigx(int m, int j) const { return base[(dd->f_bigx + m * 6 + j) * TILE]; }
  __device__ __forceinline__ double& xhat(int m, int j) const { return base[(dd->f_xhat + m * N5 + j) * TILE]; }
  __device__ __forceinline__ double& sig(int m, int q) const { return base[(dd->f_sig + m * NSYM + q) * TILE]; }
  __device__ __forceinline__ double& lam(int i) const { return base[(dd->f_lambda + i) * TILE]; }
};
__device__ __forceinline__ Cell cell_ref(double* st, const Dims& d, long long cell) {
  double* b = st + (cell >> 5) * ((long long)d.nfields * TILE) + (cell & 31);
  return Cell{b, &d};
}

// Per-thread scratch column in shared memory ([slot][thread]: conflict-free).  Small arrays that are
// indexed by a loop variable live here instead of in local memory, so the code can stay rolled
// (the kernel must fit the instruction cache) without paying L1/L2 round trips.
struct Scratch {
  double* p;
  int s;
  __device__ __forceinline__ double& operator[](int i) const { return p[i * s]; }
};
// SC_X: what getVariables leaves for EKFmatsHandler (negEta0, posEta3, Uocp_neg(thetass0), Uocp_pos(thetass3), SOC-clamp flag)
// SC_Z comes last so that a launch only has to provide SC_Z + R slots (sc_slots) for the bank at hand
// SC_S1: SigmaX of neighbour #1 before the measurement update (every gain uses it, Q1), parked here instead of in 30 registers
enum : int { SC_GAMMA = 0, SC_M = 4, SC_S1 = 8, SC_X = SC_S1 + NSYM, SC_Z = SC_X + 5, SC_TOTAL = SC_Z + HR_MAX };
__host__ __device__ constexpr int sc_slots(int R) { return SC_Z + R; }

__device__ __forceinline__ int sym_idx(int i, int j) {
  // upper triangle, row-major: (0,0..4)=0..4, (1,1..4)=5..8, (2,2..4)=9..11, (3,3..4)=12,13, (4,4)=14
  const int a = i <= j ? i : j, b = i <= j ? j : i;
  return a * 5 - (a * (a - 1)) / 2 + (b - a);
}

// first two entries of a stable ascending sort of |v - pts| (OB_step.m:251,261; iterEKF.m:221,234); returns i1 | i2 << 16
static __device__ __noinline__ int two_nearest(double v, const double* pts, int n) {
  int a = -1, b = -1;
  double b1 = INFINITY, b2 = INFINITY;
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    const double d = fabs(v - pts[i]);
    if (d < b1) { b2 = b1; b = a; b1 = d; a = i; }
    else if (d < b2) { b2 = d; b = i; }
  }
  if (a < 0) { a = 0; b = (n > 1) ? 1 : 0; }       // all distances NaN/Inf: the sort keeps the original order
  if (b < 0) b = a;
  return a | (b << 16);
}

// Same result as two_nearest(v, Zpts, nZ) without the scan: guess the bracket [lo, lo+1] from the grid's mean spacing, then
// PROVE the guess with the sort's own rule -- the nearer of the two (ties: lower index) comes first, and both outer
// neighbours sort after the farther one (left: strictly farther, it has the lower index; right: farther or tied).
// On an ascending grid |v - pts[i]| cannot decrease away from the bracket (rounding is monotonic), so the neighbours
// stand for everything beyond them.  Any doubt (NaN, guess off by one, v beyond a neighbour) falls back to the scan.
__device__ __forceinline__ int two_nearest_z(const Dims& d, double v, const double* pts) {
  const int n = d.nZ;
  if (n >= 2) {
    int lo = (int)((v - pts[0]) * d.zinv);                           // NaN / out of range -> clamped, then rejected below if wrong
    lo = lo < 0 ? 0 : (lo > n - 2 ? n - 2 : lo);
    const int hi = lo + 1;
    const double dl = fabs(v - pts[lo]), dh = fabs(v - pts[hi]);
    const bool lo_first = dl <= dh;
    const double db = lo_first ? dh : dl;
    bool ok = (dl == dl) && (dh == dh);
    if (lo > 0) ok = ok && (v >= pts[lo - 1]) && (fabs(v - pts[lo - 1]) > db);
    if (hi < n - 1) ok = ok && (v <= pts[hi + 1]) && (fabs(v - pts[hi + 1]) >= db);
    if (ok) return lo_first ? (lo | (hi << 16)) : (hi | (lo << 16));
  }
  return two_nearest(v, pts, n);
}

// getXind (iterEKF.m:219-255): ordering by value with swap, alphas not clamped (Q14).
// Writes gamma[4] and the model column m[4] = iT*nZ+iZ of the four neighbours into the scratch column.
#ifndef MEK_XIND_INLINE
#define MEK_XIND_INLINE 1   // inlined: the call sequences (registers saved around them) cost more than the code they saved
#endif
#if MEK_XIND_INLINE
static __device__ __forceinline__ void get_xind(
#else
static __device__ __noinline__ void get_xind(
#endif
    const Dims& d, const BankSmem& bk, double Tk, double SOC, Scratch sc, int tt_cached) {
  const int zz = two_nearest_z(d, SOC, bk.Zpts), z1 = zz & 0xffff, z2 = zz >> 16;
  int iZup = z1, iZlo = z1;
  double Zup = bk.Zpts[z1], Zlo = bk.Zpts[z1];
  if (d.nZ > 1) {
    Zup = bk.Zpts[z1]; iZup = z1; Zlo = bk.Zpts[z2]; iZlo = z2;
    if (Zup < Zlo) { Zup = bk.Zpts[z2]; iZup = z2; Zlo = bk.Zpts[z1]; iZlo = z1; }
  }
  const int tt = (tt_cached >= 0) ? tt_cached : two_nearest(Tk, bk.Tpts, d.nT), t1 = tt & 0xffff, t2 = tt >> 16;   // T is constant per cell
  int iTup = t1, iTlo = t1;
  double Tup = bk.Tpts[t1], Tlo = bk.Tpts[t1];
  if (d.nT > 1) {
    Tup = bk.Tpts[t1]; iTup = t1; Tlo = bk.Tpts[t2]; iTlo = t2;
    if (Tup < Tlo) { Tup = bk.Tpts[t2]; iTup = t2; Tlo = bk.Tpts[t1]; iTlo = t1; }
  }
  double aZ = 0.0, aT = 0.0;
  if (d.nZ > 1) aZ = (SOC - Zlo) / (Zup - Zlo);
  if (d.nT > 1) aT = (Tk - Tlo) / (Tup - Tlo);
  sc[SC_GAMMA + 0] = (1.0 - aT) * (1.0 - aZ);
  sc[SC_GAMMA + 1] = (1.0 - aT) * aZ;
  sc[SC_GAMMA + 2] = aT * (1.0 - aZ);
  sc[SC_GAMMA + 3] = aT * aZ;
  sc[SC_M + 0] = (double)(iTlo * d.nZ + iZlo); sc[SC_M + 1] = (double)(iTlo * d.nZ + iZup);
  sc[SC_M + 2] = (double)(iTup * d.nZ + iZlo); sc[SC_M + 3] = (double)(iTup * d.nZ + iZup);
}

// ------------------------------------------------------------------------------------------
// One pass over all M local models: plant state advance bigX = bigA.*bigX + Iapp (OB_step.m:278)
// and EKF time update xhat = A.*xhat + priorI, SigmaX = diag(A)*SigmaX*diag(A) + SigmaW
// (iterEKF.m:73-84).  This is the HBM stream of the kernel: 26 doubles per model in and out.
// Loads of two models are issued back to back before any store (the compiler may not move a load
// across a store to the same array), with evict-first hints so the stream does not flush the
// header fields out of L1.
// ------------------------------------------------------------------------------------------
template <bool PLANT, bool EKF>
struct ModelRegs {
  double bx[PLANT ? 6 : 1], x[EKF ? N5 : 1], s[EKF ? NSYM : 1];
};

// KEEP=false: evict-first accesses (.cs) for rows that are touched once and not again soon (the cold stream);
// KEEP=true: L2-cached accesses (.cg) for the models a cell advances every instant, which stay resident in L2.
template <bool KEEP> __device__ __forceinline__ double ld_state(const double* q) { return KEEP ? __ldcg(q) : __ldcs(q); }
// model rows and multipliers read or written between the streams: L2 only (.cg) as well, so that L1 is left to the header
// fields and the register spills (a row is touched once or twice per instant; 8 warps x 8 models x 26 rows of 256 B = 425 KB
// would only evict them)
#ifndef ROW_VIA_L1
__device__ __forceinline__ double ld_row(const double* q) { return __ldcg(q); }
__device__ __forceinline__ void st_row(double* q, double v) { __stcg(q, v); }
#else
__device__ __forceinline__ double ld_row(const double* q) { return *q; }
__device__ __forceinline__ void st_row(double* q, double v) { *q = v; }
#endif
template <bool KEEP> __device__ __forceinline__ void st_state(double* q, double v) { if (KEEP) __stcg(q, v); else __stcs(q, v); }

template <bool PLANT, bool EKF, bool KEEP = false>
__device__ __forceinline__ void model_load(const Cell& c, int m, ModelRegs<PLANT, EKF>& r) {
  if (PLANT) {
    const double* q = &c.bigx(m, 0);
#pragma unroll
    for (int j = 0; j < 6; ++j) r.bx[j] = ld_state<KEEP>(q + j * TILE);
  }
  if (EKF) {
    const double* q = &c.xhat(m, 0);
#pragma unroll
    for (int j = 0; j < N5; ++j) r.x[j] = ld_state<KEEP>(q + j * TILE);
    const double* w = &c.sig(m, 0);
#pragma unroll
    for (int k = 0; k < NSYM; ++k) r.s[k] = ld_state<KEEP>(w + k * TILE);
  }
}
// aa[q] = a_i*a_j (q = upper-triangle index): SigmaX(i,j) <- (a_i*a_j)*SigmaX(i,j) + SigmaW, the entries of
// diag(A)*SigmaX*diag(A) + SigmaW (iterEKF.m:81, Q12) with the two factors of each entry multiplied first -- one FMA per
// entry instead of a multiplication and an FMA (numerics contract; oracle/mek_oracle.c does the same)
template <bool PLANT, bool EKF>
__device__ __forceinline__ void model_update(const double* a, const double* aa, double Iapp, double prior, double Sw, ModelRegs<PLANT, EKF>& r) {
  if (PLANT) {
#pragma unroll
    for (int j = 0; j < 6; ++j) r.bx[j] = fma(a[j], r.bx[j], Iapp);               // OB_step.m:278
  }
  if (EKF) {
#pragma unroll
    for (int j = 0; j < N5; ++j) r.x[j] = fma(a[j], r.x[j], prior);              // iterEKF.m:77
#pragma unroll
    for (int q = 0; q < NSYM; ++q) r.s[q] = fma(aa[q], r.s[q], Sw);             // iterEKF.m:81 (Q12)
  }
}
template <bool PLANT, bool EKF, bool KEEP = false>
__device__ __forceinline__ void model_store(const Cell& c, int m, const ModelRegs<PLANT, EKF>& r) {
  if (PLANT) {
    double* q = &c.bigx(m, 0);
#pragma unroll
    for (int j = 0; j < 6; ++j) st_state<KEEP>(q + j * TILE, r.bx[j]);
  }
  if (EKF) {
    double* q = &c.xhat(m, 0);
#pragma unroll
    for (int j = 0; j < N5; ++j) st_state<KEEP>(q + j * TILE, r.x[j]);
    double* w = &c.sig(m, 0);
#pragma unroll
    for (int k = 0; k < NSYM; ++k) st_state<KEEP>(w + k * TILE, r.s[k]);
  }
}
// L2 prefetch of the state rows of models [m, m+cnt) of this warp's tile: in the AoSoA layout each of the three
// arrays is one contiguous run of 256-byte rows, so a single lane issues three bulk prefetches per request.
__device__ __forceinline__ void prefetch_models(const Dims& d, const Cell& c, int m, int cnt, bool plant, bool ekf) {
  if ((threadIdx.x & 31) != 0 || m >= d.M) return;
  if (m + cnt > d.M) cnt = d.M - m;
  if (plant) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(&c.bigx(m, 0)), "r"(cnt * 6 * TILE * 8) : "memory");
  if (ekf) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(&c.xhat(m, 0)), "r"(cnt * N5 * TILE * 8) : "memory");
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(&c.sig(m, 0)), "r"(cnt * NSYM * TILE * 8) : "memory");
  }
}

template <bool PLANT, bool EKF, int BATCH = 2>   // BATCH models' loads in flight per thread before the first store
__device__ __forceinline__ void stream_models(const Dims& d, const BankSmem& bk, const Cell& c, double Iapp, double prior) {
  const double Sw = d.SigmaW;
  constexpr int PD = 2 * BATCH + 2;                                   // prefetch distance in models
  int m = 0;
#pragma unroll 1
  for (; m + BATCH <= d.M; m += BATCH) {
    prefetch_models(d, c, m + PD, BATCH, PLANT, EKF);
    ModelRegs<PLANT, EKF> r[BATCH];
#pragma unroll
    for (int b = 0; b < BATCH; ++b) model_load<PLANT, EKF>(c, m + b, r[b]);
#pragma unroll
    for (int b = 0; b < BATCH; ++b) model_update<PLANT, EKF>(bk.mdl + (m + b) * d.mstride, bk.mdl + (m + b) * d.mstride + HB_AA, Iapp, prior, Sw, r[b]);
#pragma unroll
    for (int b = 0; b < BATCH; ++b) model_store<PLANT, EKF>(c, m + b, r[b]);
  }
#pragma unroll 1
  for (; m < d.M; ++m) {
    ModelRegs<PLANT, EKF> r0;
    model_load<PLANT, EKF>(c, m, r0);
    model_update<PLANT, EKF>(bk.mdl + m * d.mstride, bk.mdl + m * d.mstride + HB_AA, Iapp, prior, Sw, r0);
    model_store<PLANT, EKF>(c, m, r0);
  }
}

// ------------------------------------------------------------------------------------------
// Deferred advance of the models a cell is not reading (temporal blocking of the HBM stream).
//
// OB_step.m:278 and iterEKF.m:73-84 advance ALL M local models every sample instant, but one instant only READS the
// plant's four set-point neighbours and the EKF's four, and only the EKF's four receive a measurement update.  For every
// other model the instant is the data-independent recurrence x <- a.*x + u_k, S <- a.*S.*a' + SigmaW, driven by the cell's
// current history alone.  The fused kernel therefore works in blocks of KB instants: the models around the cell's
// present SOC ("hot", a warp-uniform bit mask) are advanced every instant as the reference does; the others are brought
// up to date once per block by replaying the block's recorded currents through the same recurrence, in registers —
// one HBM round trip per block instead of one per instant.  A model outside the hot set that a cell turns out to need
// (plant or EKF neighbour) is replayed on the spot and joins that cell's own set (`mine`).  Every model sees the same
// operations in the same order as in the instant-by-instant form, so results are bit-identical.
// History column in shared memory: h[0] = priorI at the start of the block, h[1+j] = u applied at block instant j;
// the EKF time update of instant j uses h[j] (iterEKF.m:210: priorI = previous ik), the plant advance h[j+1].
// ------------------------------------------------------------------------------------------
struct Defer {
  unsigned long long hot;    // warp-uniform
  unsigned long long mine;   // this cell's on-demand additions inside the current block
  int n;                     // instants recorded in the current block = advances applied to this cell's hot models
  double* h;                 // this thread's history column
  int hs;                    // column stride
};

static __device__ __noinline__ void defer_catch_up(const Dims& d, const BankSmem& bk, const Cell& c, int m, int n, const double* h, int hs) {
  ModelRegs<true, true> r;
  model_load<true, true>(c, m, r);
  double a[6], aa[NSYM];
  const double* am = bk.mdl + m * d.mstride;
#pragma unroll
  for (int j = 0; j < 6; ++j) a[j] = am[j];
#pragma unroll
  for (int q = 0; q < NSYM; ++q) aa[q] = am[HB_AA + q];
  const double Sw = d.SigmaW;
#pragma unroll 1
  for (int j = 0; j < n; ++j) model_update<true, true>(a, aa, h[(j + 1) * hs], h[j * hs], Sw, r);
  model_store<true, true>(c, m, r);
}

__device__ __forceinline__ void defer_ensure(const Dims& d, const BankSmem& bk, const Cell& c, Defer* df, int m) {
  if (!df) return;
  if (((df->hot | df->mine) >> m) & 1ull) return;
  if (df->n > 0) defer_catch_up(d, bk, c, m, df->n, df->h, df->hs);
  df->mine |= 1ull << m;
  c.f(S_SPARE) += 1.0;                                               // diagnostic: on-demand catch-ups of this cell
}

// models this cell is expected to read during the coming block: both temperature rows x the SOC set-points around the
// plant's and the EKF's present SOC, widened by the largest SOC change the current limits allow over `span` instants
static __device__ __noinline__ unsigned long long defer_hot_set(const Dims& d, const BankSmem& bk, const double* __restrict__ p, const Cell& c,
                                                         int tt, int span, int narrow) {
  const double th0n = p[P_NEG + E_THETA0], th100n = p[P_NEG + E_THETA100];
  const double rpa = d.Ts / (3600.0 * p[P_Q]);
  const double zp = (c.f(S_SOCN) - th0n) / (th100n - th0n);
  const double ze = c.f(S_EKFSOC0) - (c.f(S_X0) + c.f(S_PRIORI)) * rpa;
  double margin = 1.25 * (double)span * rpa * fmax(fabs(c.f(S_UMAX)), fabs(c.f(S_UMIN))) + 1e-4;
  double lo = fmin(zp, ze), hi = fmax(zp, ze);
  if (narrow) { lo = ze; hi = ze; margin = 0.0; }                    // test hook: everything else goes through defer_ensure
  const int a = two_nearest(lo - margin, bk.Zpts, d.nZ), b = two_nearest(hi + margin, bk.Zpts, d.nZ);
  int zl = (a & 0xffff) < (a >> 16) ? (a & 0xffff) : (a >> 16);
  int zh = (b & 0xffff) > (b >> 16) ? (b & 0xffff) : (b >> 16);
  if (zl > zh) { const int t = zl; zl = zh; zh = t; }
  const int t1 = tt & 0xffff, t2 = tt >> 16;
  // A cell on a temperature set-point gives one of its two temperature rows a weight of exactly zero (alphaT = 0 or 1, get_xind):
  // the plant and the filter never read that row's states, except the covariance of neighbour #1 = (Tlower, Zlower), which
  // every gain uses (iterEKF.m:117, Q1).  So of the zero-weight row only the models that can be neighbour #1 during the block
  // stay hot -- all of it is cold when neighbour #1 lies in the other row -- and anything else it turns out to need is
  // replayed on demand (defer_ensure).
  int zrow = -1, zrow_hi = zh;
  if (!narrow && d.nT > 1 && t1 != t2) {
    const double Tk = c.f(S_TC) + 273.15;
    const int iup = bk.Tpts[t1] < bk.Tpts[t2] ? t2 : t1, ilo = (iup == t1) ? t2 : t1;
    const double aT = (Tk - bk.Tpts[ilo]) / (bk.Tpts[iup] - bk.Tpts[ilo]);
    if (aT == 1.0) { zrow = ilo; zrow_hi = zh > zl ? zh - 1 : zl; }     // neighbour #1 is in the zero-weight row: its Zlower candidates stay
    else if (aT == 0.0) { zrow = iup; zrow_hi = zl - 1; }              // neighbour #1 is in the other row: nothing of this one is read
  }
  unsigned long long mask = 0ull;
  for (int z = zl; z <= zh; ++z) {
    if (t1 != zrow || z <= zrow_hi) mask |= 1ull << (t1 * d.nZ + z);
    if (t2 != zrow || z <= zrow_hi) mask |= 1ull << (t2 * d.nZ + z);
  }
  return mask;
}

// one instant's advance (plant with u_k, EKF time update with priorI) of the hot models; records u_k in the history
__device__ __forceinline__ void defer_stream(const Dims& d, const BankSmem& bk, const Cell& c, Defer& df, double Iapp, double prior) {
  const double Sw = d.SigmaW;
  df.h[(df.n + 1) * df.hs] = Iapp;
  unsigned long long w = df.hot;
#pragma unroll 1
  while (w) {                                                        // warp-uniform: coalesced rows, two models' loads in flight
    const int m0 = __ffsll((long long)w) - 1;
    w &= w - 1;
    if (w) {
      const int m1 = __ffsll((long long)w) - 1;
      w &= w - 1;
      ModelRegs<true, true> r0, r1;
      model_load<true, true, true>(c, m0, r0);
      model_load<true, true, true>(c, m1, r1);
      model_update<true, true>(bk.mdl + m0 * d.mstride, bk.mdl + m0 * d.mstride + HB_AA, Iapp, prior, Sw, r0);
      model_update<true, true>(bk.mdl + m1 * d.mstride, bk.mdl + m1 * d.mstride + HB_AA, Iapp, prior, Sw, r1);
      model_store<true, true, true>(c, m0, r0);
      model_store<true, true, true>(c, m1, r1);
    } else {
      ModelRegs<true, true> r0;
      model_load<true, true, true>(c, m0, r0);
      model_update<true, true>(bk.mdl + m0 * d.mstride, bk.mdl + m0 * d.mstride + HB_AA, Iapp, prior, Sw, r0);
      model_store<true, true, true>(c, m0, r0);
    }
  }
  unsigned long long x = df.mine & ~df.hot;                          // this cell's own additions (rare)
#pragma unroll 1
  while (x) {
    const int m0 = __ffsll((long long)x) - 1;
    x &= x - 1;
    ModelRegs<true, true> r0;
    model_load<true, true, true>(c, m0, r0);
    model_update<true, true>(bk.mdl + m0 * d.mstride, bk.mdl + m0 * d.mstride + HB_AA, Iapp, prior, Sw, r0);
    model_store<true, true, true>(c, m0, r0);
  }
  df.n += 1;
}

// end of a block: replay the block's history through every model outside the hot set (one HBM round trip per block)
__device__ __forceinline__ void defer_flush(const Dims& d, const BankSmem& bk, const Cell& c, Defer& df) {
#pragma unroll 1
  for (int m = 0; m < d.M; ++m) {
    if ((df.hot >> m) & 1ull) continue;                              // warp-uniform
    prefetch_models(d, c, m + 2, 1, true, true);
    if (df.n > 0 && !((df.mine >> m) & 1ull)) defer_catch_up(d, bk, c, m, df.n, df.h, df.hs);
  }
  df.n = 0;
  df.mine = 0ull;
}

// ------------------------------------------------------------------------------------------
// plant: OB_step.m simStep (:188-357).  Returns Vcell.
// ------------------------------------------------------------------------------------------
template <bool ADVANCE>   // ADVANCE=false: the caller runs stream_models<true,...> right after (fused pass)
__device__ __forceinline__ double plant_step(const Dims& d, const BankSmem& bk, const double* __restrict__ p, const Cell& c,
                                             double Iapp, int tt_cached, Defer* df = nullptr, double qs = 1.0, double rcs = 1.0) {
  // qs, rcs: plant-model mismatch scenario (mek_config.plant_q_scale / plant_rc_scale): the PLANT's capacity and collector
  // resistance are the bank's times these factors; 1.0 (the default, a compile-time constant in the lean instances) is the reference
  const double T = c.f(S_TC) + 273.15;                               // OB_step.m:75
  const double th0n = p[P_NEG + E_THETA0], th100n = p[P_NEG + E_THETA100];
  const double th0p = p[P_POS + E_THETA0], th100p = p[P_POS + E_THETA100];
  const double Q = p[P_Q] * qs;
  double SOCn = c.f(S_SOCN), SOCp = c.f(S_SOCP);
  const double negSOC = SOCn, posSOC = SOCp;                         // :226-227 (Q16)
  const double cellSOC = (SOCn - th0n) / (th100n - th0n);            // :228
  const double2 dU = docp2(p, SOCn, SOCp);                           // :231-232
  const double dUn = dU.x, dUp = dU.y;
  const double dQn = fabs(th100n - th0n), dQp = fabs(th100p - th0p);
  const double2 res0 = div2(-dQn, 3600.0 * Q - (c.f(S_CDLN) * dQn) * dUn, dQp, 3600.0 * Q - (c.f(S_CDLP) * dQp) * dUp);   // :235-236
  const double res0n = res0.x, res0p = res0.y;
  SOCn = SOCn + (res0n * Iapp) * d.Ts;                               // :239-240
  SOCp = SOCp + (res0p * Iapp) * d.Ts;
  int nw = 0;
  if (SOCn < 0.0) { SOCn = 0.0; ++nw; }                              // :241-244
  if (SOCn > 1.0) { SOCn = 1.0; ++nw; }
  if (SOCp < 0.0) { SOCp = 0.0; ++nw; }
  if (SOCp > 1.0) { SOCp = 1.0; ++nw; }
  if (nw) c.f(S_PLANTWARN) += (double)nw;
  const int zz = two_nearest_z(d, cellSOC, bk.Zpts), z1 = zz & 0xffff, z2 = zz >> 16;   // :251-259 (ordering by index, Q15)
  const int iZup = z1 > z2 ? z1 : z2, iZlo = z1 < z2 ? z1 : z2;
  const int tt = (tt_cached >= 0) ? tt_cached : two_nearest(T, bk.Tpts, d.nT), t1 = tt & 0xffff, t2 = tt >> 16;   // :261-269
  const int iTup = t1 > t2 ? t1 : t2, iTlo = t1 < t2 ? t1 : t2;
  const double Zup = bk.Zpts[iZup], Zlo = bk.Zpts[iZlo], Tup = bk.Tpts[iTup], Tlo = bk.Tpts[iTlo];
  double aZ = 0.0, aT = 0.0;                                         // :281-283
  if (Zup != Zlo) aZ = (cellSOC - Zlo) / (Zup - Zlo);
  if (Tup != Tlo) aT = (T - Tlo) / (Tup - Tlo);
  // four corner outputs (:272-275) blended as (1-aT)*((1-aZ)*y1 + aZ*y2) + aT*((1-aZ)*y3 + aZ*y4) (:285):
  // one rolled pass per temperature set-point computes the inner SOC blend of its two corners and adds its share to yk:
  // first (1-aT)*lower, then + aT*upper (the two roundings of the reference's expression); a row whose weight is exactly
  // zero is left out and the other one taken as it is
  double yk[9];
  const bool whole = (aT == 0.0) || (aT == 1.0);
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    if (h ? (aT == 0.0) : (aT == 1.0)) continue;                     // that row's blend weight is exactly zero
    const double wgt = h ? aT : (1.0 - aT);
    const int iT = h ? iTup : iTlo;
    const int ma = iT * d.nZ + iZlo, mb = iT * d.nZ + iZup;
    const double* ra = bk.mdl + ma * d.mstride + HB_ROWS;
    const double* rb = bk.mdl + mb * d.mstride + HB_ROWS;
    defer_ensure(d, bk, c, df, ma);
    defer_ensure(d, bk, c, df, mb);
    double xa[6], xb[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) { xa[j] = ld_row(&c.bigx(ma, j)); xb[j] = ld_row(&c.bigx(mb, j)); }
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const double2* pa = reinterpret_cast<const double2*>(ra + r * HRS);
      const double2* pb = reinterpret_cast<const double2*>(rb + r * HRS);
      const double2 a01 = pa[0], a23 = pa[1], a4d = pa[2], a6p = pa[3], b01 = pb[0], b23 = pb[1], b4d = pb[2], b6p = pb[3];
      const double ca[7] = {a01.x, a01.y, a23.x, a23.y, a4d.x, a6p.x, a4d.y};   // C(1..6), D
      const double cb[7] = {b01.x, b01.y, b23.x, b23.y, b4d.x, b6p.x, b4d.y};
      double ya = ca[0] * xa[0], yb = cb[0] * xb[0];
#pragma unroll
      for (int j = 1; j < 6; ++j) { ya = fma(ca[j], xa[j], ya); yb = fma(cb[j], xb[j], yb); }
      ya = fma(ca[6], Iapp, ya);
      yb = fma(cb[6], Iapp, yb);
      const double inner = (1.0 - aZ) * ya + aZ * yb;
      yk[r] = whole ? inner : (h == 0 ? wgt * inner : yk[r] + wgt * inner);   // :285
    }
  }
  if (ADVANCE) stream_models<true, false>(d, bk, c, Iapp, 0.0);         // :278 advance ALL models
  double tss0 = yk[HR_TSS0] + c.f(S_SOC0N), tss3 = yk[HR_TSS3] + c.f(S_SOC0P);   // :304-305 (Q6)
  tss0 = fmin(fmax(tss0, 1e-6), 1.0 - 1e-6);                         // :309-310
  tss3 = fmin(fmax(tss3, 1e-6), 1.0 - 1e-6);
  const double the0 = fmax(yk[HR_THE0] + 1.0, 1e-6), the3 = fmax(yk[HR_THE3] + 1.0, 1e-6);   // :325-326
  const double k0n = k0_e(p, P_NEG, negSOC, c.f(S_ARRK0N)), k0p = k0_e(p, P_POS, posSOC, c.f(S_ARRK0P));
  const double2 ri0 = sqrt2((the0 * (1.0 - tss0)) * tss0, (the3 * (1.0 - tss3)) * tss3);       // :331-332
  const double i0n = k0n * ri0.x, i0p = k0p * ri0.y;
  const double pref = ((2.0 * p[P_R]) * T) / p[P_F];
  const double2 ash = asinh2_q(yk[HR_IF0], 2.0 * i0n, yk[HR_IF3], 2.0 * i0p);
  const double negEta0 = pref * ash.x;                               // :333-334
  const double posEta3 = pref * ash.y;
  const double2 U03 = ocp2(p, tss0, tss3);                           // :337-340
  const double Un0 = U03.x, Up3 = U03.y;
  const double Rfn = rf_e(p, P_NEG, negSOC, c.f(S_ARRRFN)), Rfp = rf_e(p, P_POS, posSOC, c.f(S_ARRRFP));
  double V = ((((posEta3 - negEta0) + yk[HR_PHIE]) + Up3) - Un0) + (Rfp * yk[HR_IFDL3] - Rfn * yk[HR_IFDL0]);   // :341-342
  V = V - (p[P_RC] * rcs) * Iapp;                                    // :344
  c.f(S_SOCN) = SOCn;
  c.f(S_SOCP) = SOCp;
  return V;
}

// ------------------------------------------------------------------------------------------
// EKF pieces
// ------------------------------------------------------------------------------------------
// Blend of the four neighbours' outputs with the bilinear weights (iterEKF.m:302-313, "[...]*Xind.gamma").  A neighbour whose
// weight is EXACTLY zero contributes nothing -- the rule of this build and of oracle/mek_oracle.c: the
// reference adds its +-0 product, which changes nothing but the sign of an exact zero (or everything, if that neighbour's
// output is non-finite).  A cell whose temperature sits on a set-point of the bank (runMPC's own 25 degC) has two such
// neighbours at every instant.
__device__ __forceinline__ double blend4(const double (&yc)[4], const double (&g)[4]) {
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) acc = (g[k] != 0.0) ? fma(yc[k], g[k], acc) : acc;
  return acc;
}

// Rows of getVariables' blended output (iterEKF.m:302-313, hot rows only), left in sc[SC_Z + r].  Two bodies: all four
// neighbours, and the two neighbours k0, k0+1 of ONE temperature row for a cell whose temperature sits on a set-point (the other
// row's weights are exactly zero: blend4 leaves it out, so it is not even loaded).  k0 is a run-time value, one body serves both rows.
template <int NK>
__device__ __forceinline__ void gv_rows_n(const Dims& d, const BankSmem& bk, const Cell& c, double ik, Scratch sc, int k0) {
  double xk[NK][N5], g[NK];
  const double* mm[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int m = (int)sc[SC_M + k0 + k];
    g[k] = sc[SC_GAMMA + k0 + k];
    mm[k] = bk.mdl + m * d.mstride + HB_ROWS;
#pragma unroll
    for (int j = 0; j < N5; ++j) xk[k][j] = ld_row(&c.xhat(m, j));   // :267-270
  }
#pragma unroll 1
  for (int r = 0; r < d.R; ++r) {                                    // :302-313  Z = [C1*x1+D1*i, ...]*gamma
    if (r == HR_PHI) continue;
    double z = 0.0;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const double2* pr = reinterpret_cast<const double2*>(mm[k] + r * HRS);
      const double2 c01 = pr[0], c23 = pr[1], c4d = pr[2];
      const double cr[N5] = {c01.x, c01.y, c23.x, c23.y, c4d.x};
      double acc = cr[0] * xk[k][0];
#pragma unroll
      for (int j = 1; j < N5; ++j) acc = fma(cr[j], xk[k][j], acc);
      const double yc = fma(c4d.y, ik, acc);
      z = (g[k] != 0.0) ? fma(yc, g[k], z) : z;                      // blend4, term by term
    }
    sc[SC_Z + r] = z;
  }
}

// getVariables (iterEKF.m:259-417) restricted to the rows that feed Vcell, the warnings and
// EKFmatsHandler.  Reads the neighbour set from the scratch column (get_xind), leaves the corrected
// outputs in sc[SC_Z + HR_*] and returns Vcell.  Warnings go straight to the cell's counters.
#ifndef MEK_GV_INLINE
#define MEK_GV_INLINE 1
#endif
#if MEK_GV_INLINE
static __device__ __forceinline__ double get_variables(
#else
static __device__ __noinline__ double get_variables(
#endif
    const Dims& d, const BankSmem& bk, const double* __restrict__ p, const Cell& c, double ik,
                                             double T, Scratch sc) {
#ifdef __CUDA_ARCH__
  __builtin_assume(__isShared(sc.p));              // address-space hint: LDS/STS for the scratch column, not generic LD/ST
#endif
  const double rpa = d.Ts / (3600.0 * p[P_Q]);
  const double xSOC = c.f(S_EKFSOC0) - c.f(S_X0) * rpa;              // :277
  double SOCn = soc_e(p, P_NEG, xSOC), SOCp = soc_e(p, P_POS, xSOC); // :282-283
  int nwarn = 0;
  unsigned status = 0u;
  if (SOCn < 0.0) { ++nwarn; SOCn = 1e-6; }                          // :284-295 (Q17)
  if (SOCn > 1.0) { ++nwarn; SOCn = 1.0 - 1e-6; }
  if (SOCp < 0.0) { ++nwarn; SOCp = 1e-6; }
  if (SOCp > 0.998) { ++nwarn; SOCp = 0.998; }
  const bool soc_clamped = nwarn != 0;
  {
    // a temperature on a set-point of the bank zeroes the weights of one temperature row (both of its neighbours)
    const bool lo0 = sc[SC_GAMMA + 0] == 0.0 && sc[SC_GAMMA + 1] == 0.0, hi0 = sc[SC_GAMMA + 2] == 0.0 && sc[SC_GAMMA + 3] == 0.0;
    if (lo0 != hi0) gv_rows_n<2>(d, bk, c, ik, sc, lo0 ? 2 : 0);
    else gv_rows_n<4>(d, bk, c, ik, sc, 0);
  }
  // negThetass rows (:331-343): Thetass0 plus the extra rows, which only matter for the warnings
  double tss0, tss3, the0, the3;
  {
    double v = sc[SC_Z + HR_TSS0] + SOCn;
    bool lo = v < 0.0;
    if (lo) v = 1e-6;
    bool hi = v > 1.0;
    if (hi) v = 1.0 - 1e-6;
#pragma unroll 1
    for (int e = 0; e < d.nx_negtss; ++e) {
      double w = sc[SC_Z + HR_FIXED + e] + SOCn;
      if (w < 0.0) { lo = true; w = 1e-6; }
      if (w > 1.0) hi = true;
    }
    nwarn += (int)lo + (int)hi;
    tss0 = v;
  }
  {                                                                  // posThetass rows (:345-357)
    double v = sc[SC_Z + HR_TSS3] + SOCp;
    bool lo = v < 0.0;
    if (lo) v = 1e-6;
    bool hi = v > 0.998;
    if (hi) v = 0.998;
#pragma unroll 1
    for (int e = 0; e < d.nx_postss; ++e) {
      double w = sc[SC_Z + HR_FIXED + d.nx_negtss + e] + SOCp;
      if (w < 0.0) { lo = true; w = 1e-6; }
      if (w > 0.998) hi = true;
    }
    nwarn += (int)lo + (int)hi;
    tss3 = v;
  }
  {                                                                  // Thetae rows (:383-389, Q18)
    double v0 = sc[SC_Z + HR_THE0] + 1.0;
    double v3 = sc[SC_Z + HR_THE3] + 1.0;
    bool lo = false;
    if (v0 < 0.0) { lo = true; v0 = 1e-6; }
    if (v3 < 0.0) { lo = true; v3 = 1e-6; }
#pragma unroll 1
    for (int e = 0; e < d.nx_thetae; ++e) {
      const double w = sc[SC_Z + HR_FIXED + d.nx_negtss + d.nx_postss + e] + 1.0;
      if (w < 0.0) lo = true;
    }
    if (lo) { ++nwarn; status |= ST_Q18; }
    the0 = v0;
    the3 = v3;
  }
  sc[SC_Z + HR_TSS0] = tss0; sc[SC_Z + HR_TSS3] = tss3; sc[SC_Z + HR_THE0] = the0; sc[SC_Z + HR_THE3] = the3;
  const double If0 = sc[SC_Z + HR_IF0], If3 = sc[SC_Z + HR_IF3];     // :324-325
  const double k0n = k0_e(p, P_NEG, SOCn, c.f(S_ARRK0N)), k0p = k0_e(p, P_POS, SOCp, c.f(S_ARRK0P));   // :392-393
  const double2 ri0 = sqrt2((the0 * (1.0 - tss0)) * tss0, (the3 * (1.0 - tss3)) * tss3);       // :395-398
  const double i0n = k0n * ri0.x, i0p = k0p * ri0.y;
  const double pref = ((2.0 * p[P_R]) * T) / p[P_F];
  const double2 ash = asinh2_q(If0, 2.0 * i0n, If3, 2.0 * i0p);
  const double negEta0 = pref * ash.x;                               // :400-401
  const double posEta3 = pref * ash.y;
  const double2 U03 = ocp2(p, tss0, tss3);                           // :404-407
  const double Un0 = U03.x, Up3 = U03.y;
  const double Rfn = rf_e(p, P_NEG, SOCn, c.f(S_ARRRFN)), Rfp = rf_e(p, P_POS, SOCp, c.f(S_ARRRFP));
  if (nwarn) c.f(S_WARN) += (double)nwarn;                           // shortWarn -> warnCount (:749)
  if (status) c.f(S_STATUS) = (double)((unsigned)c.f(S_STATUS) | status);
  sc[SC_X + 0] = negEta0; sc[SC_X + 1] = posEta3; sc[SC_X + 2] = Un0; sc[SC_X + 3] = Up3;
  sc[SC_X + 4] = soc_clamped ? 1.0 : 0.0;
  return ((((posEta3 - negEta0) + sc[SC_Z + HR_PHIE]) + Up3) - Un0) + (Rfp * sc[SC_Z + HR_IFDL3] - Rfn * sc[SC_Z + HR_IFDL0]);   // :409-410 (Q5)
}

// PSD repair (iterEKF.m:143-145) on the symmetric 15-entry storage.  If an LDL^T sweep finds every
// pivot > 0 the matrix is positive definite, V*S*V' equals SigmaX and line 145 is the identity;
// otherwise a cyclic Jacobi eigen-decomposition gives HH = V|w|V'.  Returns 1 if the eigen path ran.
static __device__ __noinline__ int psd_repair_jacobi(double* S) {
  double A[5][5], V[5][5];
  for (int i = 0; i < 5; ++i)
    for (int j = 0; j < 5; ++j) { A[i][j] = S[sym_idx(i, j)]; V[i][j] = (i == j) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int i = 0; i < 5; ++i) {
      dg = dg + A[i][i] * A[i][i];
      for (int j = i + 1; j < 5; ++j) off = off + A[i][j] * A[i][j];
    }
    if (!(off > 1e-34 * dg)) break;
    for (int pI = 0; pI < 4; ++pI)
      for (int q = pI + 1; q < 5; ++q) {
        const double apq = A[pI][q];
        if (apq == 0.0) continue;
        const double theta = (A[q][q] - A[pI][pI]) / (2.0 * apq);
        double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
        if (theta < 0.0) t = -t;
        const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
        for (int k = 0; k < 5; ++k) {
          const double akp = A[k][pI], akq = A[k][q];
          A[k][pI] = cs * akp - sn * akq;
          A[k][q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < 5; ++k) {
          const double apk = A[pI][k], aqk = A[q][k];
          A[pI][k] = cs * apk - sn * aqk;
          A[q][k] = sn * apk + cs * aqk;
        }
        for (int k = 0; k < 5; ++k) {
          const double vkp = V[k][pI], vkq = V[k][q];
          V[k][pI] = cs * vkp - sn * vkq;
          V[k][q] = sn * vkp + cs * vkq;
        }
      }
  }
  for (int i = 0; i < 5; ++i)
    for (int j = i; j < 5; ++j) {
      double hh = 0.0;
      for (int k = 0; k < 5; ++k) hh = hh + (V[i][k] * fabs(A[k][k])) * V[j][k];
      const double s = S[sym_idx(i, j)];
      S[sym_idx(i, j)] = (((s + s) + hh) + hh) / 4.0;                // :145
    }
  return 1;
}

__device__ __forceinline__ int psd_repair(double* S, int force) {
  if (!force) {
    double L[5][5], dd[5];
    bool pd = true;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      if (pd) {
        double dj = S[sym_idx(j, j)];
#pragma unroll
        for (int k = 0; k < j; ++k) dj = fma(-(L[j][k] * L[j][k]), dd[k], dj);
        if (!(dj > 0.0)) {
          pd = false;
        } else {
          dd[j] = dj;
          const double idj = rcp_d(dj);
#pragma unroll
          for (int i = j + 1; i < 5; ++i) {
            double v = S[sym_idx(i, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) v = fma(-(L[i][k] * L[j][k]), dd[k], v);
            L[i][j] = v * idj;
          }
        }
      }
    }
    if (pd) return 0;
  }
  // rare path, out of line: it works on a copy so that S itself never has its address taken (an array whose address
  // escapes lives in local memory on the common path too)
  double T[NSYM];
#pragma unroll
  for (int q = 0; q < NSYM; ++q) T[q] = S[q];
  const int r = psd_repair_jacobi(T);
#pragma unroll
  for (int q = 0; q < NSYM; ++q) S[q] = T[q];
  return r;
}

struct EkfOut {          // what EKFmatsHandler hands to the MPC (EKFmatsHandler.m:101-114) + zk(end), vhat
  double vhat, soc;
  int iT, iZ, lost;
  double Cv[N5], Dv, Cphi[N5], Dphi, bv, bphi, a[N5], xhat[6];
};

// iterEKF (OB mode, iterEKF.m:55-210) followed by EKFmatsHandler (EKFmatsHandler.m:26-114)
template <bool TIME_UPDATE>   // TIME_UPDATE=false: the caller already ran stream_models<...,true> with priorI
__device__ __forceinline__ void ekf_step(const Dims& d, const BankSmem& bk, const double* __restrict__ p, const Cell& c, double vk,
                                         double ik, EkfOut& o, Scratch sc, int tt_cached, Defer* df = nullptr) {
#ifdef __CUDA_ARCH__
  __builtin_assume(__isShared(sc.p));              // address-space hint (see get_variables)
#endif
  o.lost = 0;
  if (c.f(S_WARN) > 10.0) {                                          // iterEKF.m:55-59 (Q19)
    c.f(S_STATUS) = (double)((unsigned)c.f(S_STATUS) | ST_LOST);
    o.lost = 1;
    return;
  }
  const double Tk = c.f(S_TC) + 273.15;                              // :65
  const double prior = c.f(S_PRIORI), Sw = d.SigmaW;
  if (TIME_UPDATE) stream_models<false, true>(d, bk, c, 0.0, prior); // :73-84 time update of ALL models
  double x0 = c.f(S_X0) + prior;                                     // :85
  double sx0 = c.f(S_SIGX0) + Sw;                                    // :86
  c.f(S_X0) = x0;
  const double rpa = d.Ts / (3600.0 * p[P_Q]);
  double SOC = c.f(S_EKFSOC0) - x0 * rpa;                            // :89
  get_xind(d, bk, Tk, SOC, sc, tt_cached);                           // :107
  if (df) {                                                          // what this instant reads: the weighted neighbours, and neighbour #1's SigmaX (Q1)
#pragma unroll 1
    for (int k = 0; k < 4; ++k)
      if (k == 0 || sc[SC_GAMMA + k] != 0.0) defer_ensure(d, bk, c, df, (int)sc[SC_M + k]);
  }
  double vhat = get_variables(d, bk, p, c, ik, Tk, sc);              // :108
  // SigmaX of neighbour #1 *before* any update: every gain uses it (:117, Q1).  Loaded here so that the L2 round trip
  // overlaps the scalar block below.
  {
    const int m0 = (int)sc[SC_M + 0];
    double S1[NSYM];
#pragma unroll
    for (int q = 0; q < NSYM; ++q) S1[q] = ld_row(&c.sig(m0, q));
#pragma unroll
    for (int q = 0; q < NSYM; ++q) sc[SC_S1 + q] = S1[q];
  }
  // ---- getChatV scalars (:421-519) ----
  const double tss0 = sc[SC_Z + HR_TSS0], tss3 = sc[SC_Z + HR_TSS3];
  double Rfn, Rfp, Rctn, Rctp, dUn0, dUp3, C0;
  {
    const double xSOC = c.f(S_EKFSOC0) - x0 * rpa;                   // :438
    const double sn = soc_e(p, P_NEG, xSOC), sp = soc_e(p, P_POS, xSOC);
    Rfn = rf_e(p, P_NEG, sn, c.f(S_ARRRFN)); Rfp = rf_e(p, P_POS, sp, c.f(S_ARRRFP));
    const double k0n = k0_e(p, P_NEG, sn, c.f(S_ARRK0N)), k0p = k0_e(p, P_POS, sp, c.f(S_ARRK0P));
    const double2 ri0 = sqrt2((sc[SC_Z + HR_THE0] * (1.0 - tss0)) * tss0, (sc[SC_Z + HR_THE3] * (1.0 - tss3)) * tss3);
    const double i0n = k0n * ri0.x, i0p = k0p * ri0.y;
    { const double2 rct = div2(p[P_R] * Tk, p[P_F] * i0n, p[P_R] * Tk, p[P_F] * i0p); Rctn = rct.x; Rctp = rct.y; }   // :467-468
    { const double2 dU = docp2(p, tss0, tss3); dUn0 = dU.x; dUp3 = dU.y; }   // :495-496
    const double dsn = soc_e(p, P_NEG, 1.0) - soc_e(p, P_NEG, 0.0), dsp = soc_e(p, P_POS, 1.0) - soc_e(p, P_POS, 0.0);
    const double2 res0 = div2(((-dUn0) * d.Ts) * dsn, 3600.0 * p[P_Q], ((-dUp3) * d.Ts) * dsp, 3600.0 * p[P_Q]);   // :497-500
    const double res0n = res0.x, res0p = res0.y;
    C0 = res0p - res0n;                                              // :501
  }
  const double svt0 = (C0 * sx0) * C0 + d.SigmaV;                    // :123
  const double L0 = (sx0 * C0) / svt0;                               // :124
  const double resid = vk - vhat;                                    // :133
  int njac = 0, nbump = 0;
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {                                      // :116-121 gain, :136-154 update; sequential (Q2)
    const int mo = (int)sc[SC_M + k];
    const double g = sc[SC_GAMMA + k];
    if (g == 0.0) {
      // zero weight: Chat = 0, so the gain is 0, svt = SigmaV, xhat and SigmaX stay as they are (:138,:142 add +-0) and the
      // repair (:143-145) of an unchanged covariance is the identity up to rounding -- skipped, here and in the oracle.
      // What is left of the update is the Q-bump (:148-151).
      if (resid * resid > 9.0 * (0.0 + d.SigmaV)) {
        defer_ensure(d, bk, c, df, mo);
#pragma unroll
        for (int q2 = 0; q2 < NSYM; ++q2) st_row(&c.sig(mo, q2), ld_row(&c.sig(mo, q2)) * 2.0);
        ++nbump;
      }
      continue;
    }
    const double* mm = bk.mdl + mo * d.mstride + HB_ROWS;
    double Chat[N5], v[N5], L[N5], S[NSYM], xs[N5];
    // this neighbour's state, requested before the gain is computed (the loads overlap it); after the previous
    // neighbour's stores in program order, so a duplicate neighbour sees its own update (Q2)
#pragma unroll
    for (int j = 0; j < N5; ++j) xs[j] = ld_row(&c.xhat(mo, j));
#pragma unroll
    for (int q = 0; q < NSYM; ++q) S[q] = ld_row(&c.sig(mo, q));
#pragma unroll
    for (int j = 0; j < N5; ++j) {                                   // getChatV for this neighbour
      double t = Rfp * (g * mm[HR_IFDL3 * HRS + j]) - Rfn * (g * mm[HR_IFDL0 * HRS + j]);    // :444
      t = (t + Rctp * (g * mm[HR_IF3 * HRS + j])) - Rctn * (g * mm[HR_IF0 * HRS + j]);       // :471
      t = t + g * mm[HR_PHIE * HRS + j];                                                    // :485
      t = t + (dUp3 * (g * mm[HR_TSS3 * HRS + j]) - dUn0 * (g * mm[HR_TSS0 * HRS + j]));      // :504
      Chat[j] = t;
    }
    double S1[NSYM];
#pragma unroll
    for (int q = 0; q < NSYM; ++q) S1[q] = sc[SC_S1 + q];
#pragma unroll
    for (int i = 0; i < N5; ++i) {
      double acc = S1[sym_idx(i, 0)] * Chat[0];
#pragma unroll
      for (int j = 1; j < N5; ++j) acc = fma(S1[sym_idx(i, j)], Chat[j], acc);
      v[i] = acc;
    }
    double s = Chat[0] * v[0];
#pragma unroll
    for (int i = 1; i < N5; ++i) s = fma(Chat[i], v[i], s);
    const double svt = s + d.SigmaV;                                 // :118
    const double isv = rcp_d(svt);
#pragma unroll
    for (int i = 0; i < N5; ++i) L[i] = v[i] * isv;                  // :120
#pragma unroll
    for (int j = 0; j < N5; ++j) st_row(&c.xhat(mo, j), fma(L[j], resid, xs[j]));   // :138
    int q = 0;
#pragma unroll
    for (int i = 0; i < N5; ++i)
#pragma unroll
      for (int j = i; j < N5; ++j, ++q) S[q] = fma(-(L[i] * svt), L[j], S[q]);   // :142
    njac += psd_repair(S, d.force_jacobi);                           // :143-145
    if (resid * resid > 9.0 * svt) {                                 // :148-151 Q-bump
#pragma unroll
      for (int q2 = 0; q2 < NSYM; ++q2) S[q2] = S[q2] * 2.0;
      ++nbump;
    }
#pragma unroll
    for (int q2 = 0; q2 < NSYM; ++q2) st_row(&c.sig(mo, q2), S[q2]);
  }
  x0 = x0 + L0 * resid;                                              // :155
  sx0 = sx0 - (L0 * svt0) * L0;                                      // :156
  c.f(S_X0) = x0;
  c.f(S_SIGX0) = sx0;
  SOC = c.f(S_EKFSOC0) - x0 * rpa;                                   // :159
  get_xind(d, bk, Tk, SOC, sc, tt_cached);                           // :181
  if (df) {                                                          // the estimate may have moved to other set-points
#pragma unroll 1
    for (int k = 0; k < 4; ++k)
      if (sc[SC_GAMMA + k] != 0.0) defer_ensure(d, bk, c, df, (int)sc[SC_M + k]);
  }
  vhat = get_variables(d, bk, p, c, ik, Tk, sc);                     // :182
  const double Zsoc = c.f(S_EKFSOC0) - x0 * rpa;                     // :416
  c.f(S_PRIORI) = ik;                                                // :210
  if (njac) c.f(S_NJAC) += (double)njac;
  if (nbump) c.f(S_NBUMP) += (double)nbump;
  o.vhat = vhat;
  o.soc = Zsoc;
  // ---- EKFmatsHandler ----
  int imax = 0;                                                      // EKFmatsHandler.m:26 first max, NaN ignored
  {
    bool have = false;
    double best = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double g = sc[SC_GAMMA + k];
      if (g != g) continue;
      if (!have || g > best) { best = g; imax = k; have = true; }
    }
  }
  const int mp = (int)sc[SC_M + imax];
  o.iT = mp / d.nZ;                                                  // the model column is iT*nZ + iZ (get_xind)
  o.iZ = mp - o.iT * d.nZ;
  const double* am = bk.mdl + mp * d.mstride;
#pragma unroll
  for (int j = 0; j < N5; ++j) { o.xhat[j] = ld_row(&c.xhat(mp, j)); o.a[j] = am[j]; }   // :33-35
  o.xhat[5] = 0.0;                                                   // :33 (Q3)
  {
    const double ztss0 = sc[SC_Z + HR_TSS0], ztss3 = sc[SC_Z + HR_TSS3];
    const double TK = c.f(S_TC) + 273.15;                            // :53
    const double sn = soc_e(p, P_NEG, Zsoc), sp = soc_e(p, P_POS, Zsoc);   // :56-58
    const double Rfn2 = rf_e(p, P_NEG, sn, c.f(S_ARRRFN)), Rfp2 = rf_e(p, P_POS, sp, c.f(S_ARRRFP));
    const double* mm = am + HB_ROWS;
#pragma unroll
    for (int j = 0; j < N5; ++j)                                     // :75-76
      o.Cv[j] = (Rfp2 * mm[HR_IFDL3 * HRS + j] - Rfn2 * mm[HR_IFDL0 * HRS + j]) + mm[HR_PHIE * HRS + j];
    o.Dv = (Rfp2 * mm[HR_IFDL3 * HRS + HR_D] - Rfn2 * mm[HR_IFDL0 * HRS + HR_D]) + mm[HR_PHIE * HRS + HR_D];   // :79-80
    // :57-69,84-89 evaluate, at zk's SOC and the corrected thetass / thetae, exactly what the getVariables call above (:182)
    // just evaluated -- unless that call clamped its electrode SOCs (Q17), in which case its k0 arguments differ
    double Upos, Uneg, negEta0, posEta3;
    if (sc[SC_X + 4] == 0.0) {
      negEta0 = sc[SC_X + 0]; posEta3 = sc[SC_X + 1]; Uneg = sc[SC_X + 2]; Upos = sc[SC_X + 3];
    } else {
      const double k0n = k0_e(p, P_NEG, sn, c.f(S_ARRK0N)), k0p = k0_e(p, P_POS, sp, c.f(S_ARRK0P));
      const double2 ri0 = sqrt2((sc[SC_Z + HR_THE0] * (1.0 - ztss0)) * ztss0, (sc[SC_Z + HR_THE3] * (1.0 - ztss3)) * ztss3);   // :63-66
      const double i0n = k0n * ri0.x, i0p = k0p * ri0.y;
      { const double2 U = ocp2(p, ztss0, ztss3); Uneg = U.x; Upos = U.y; }   // :84-85
      const double pref = ((2.0 * p[P_R]) * TK) / p[P_F];
      const double2 ash = asinh2_q(sc[SC_Z + HR_IF0], 2.0 * i0n, sc[SC_Z + HR_IF3], 2.0 * i0p);
      negEta0 = pref * ash.x;                                        // :88-89
      posEta3 = pref * ash.y;
    }
    o.bv = ((Upos - Uneg) + (posEta3 - negEta0)) + 0.0;              // :92
    o.bphi = uocp_n(p, sn);                                          // :96 (Q21)
#pragma unroll
    for (int j = 0; j < N5; ++j) o.Cphi[j] = mm[HR_PHI * HRS + j];     // :97
    o.Dphi = mm[HR_PHI * HRS + HR_D];                                     // :98
  }
}

// ------------------------------------------------------------------------------------------
// predMat.m for one scalar output with diagonal A and B = 1 (EKFmatsHandler.m:35-36):
//   h[k] = Cbar*Abar^k*Bbar (k = 0..Np; h[i] doubles as Phi(i,7)),  phix[i-1] = Phi(i,1:6)*x
// ------------------------------------------------------------------------------------------
template <int NPM>
__device__ __forceinline__ void pred_row(const double* a, const double* c5, double c6, double dd, const double* x6, int Np,
                                         double* h, double* phix) {
  double pk[6], pw[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) { pk[j] = 0.0; pw[j] = 1.0; }
#pragma unroll
  for (int k = 0; k <= NPM; ++k) {
    if (k <= Np) {
      double acc = c5[0] * pk[0];
#pragma unroll
      for (int j = 1; j < N5; ++j) acc = fma(c5[j], pk[j], acc);
      acc = fma(c6, pk[5], acc);
      h[k] = acc + dd;                                               // predMat.m:34
#pragma unroll
      for (int j = 0; j < N5; ++j) pk[j] = fma(a[j], pk[j], 1.0);    // predMat.m:35
      pk[5] = pk[5] + 1.0;
    }
  }
#pragma unroll
  for (int i = 1; i <= NPM; ++i) {
    if (i <= Np) {
#pragma unroll
      for (int j = 0; j < N5; ++j) pw[j] = a[j] * pw[j];             // predMat.m:41
      double acc = (c5[0] * pw[0]) * x6[0];
#pragma unroll
      for (int j = 1; j < N5; ++j) acc = fma(c5[j] * pw[j], x6[j], acc);
      acc = fma(c6 * pw[5], x6[5], acc);
      phix[i - 1] = acc;                                             // predMat.m:43 (first six columns)
    }
  }
}

// E x = b for the SPD Nc x Nc Hessian: Gaussian elimination without pivoting.  The factor keeps the
// RECIPROCAL of each pivot on the diagonal so the many solves against it need no division.
template <int NCM>
__device__ __forceinline__ void spd_factor(int Nc, const double* E, double* LU) {
#pragma unroll
  for (int i = 0; i < NCM * NCM; ++i) LU[i] = E[i];
#pragma unroll
  for (int k = 0; k < NCM; ++k)
    if (k < Nc) {
      const double ip = rcp_d(LU[k * NCM + k]);
      LU[k * NCM + k] = ip;
#pragma unroll
      for (int i = k + 1; i < NCM; ++i)
        if (i < Nc) {
          const double l = LU[i * NCM + k] * ip;
          LU[i * NCM + k] = l;
#pragma unroll
          for (int j = k + 1; j < NCM; ++j)
            if (j < Nc) LU[i * NCM + j] = fma(-l, LU[k * NCM + j], LU[i * NCM + j]);
        }
    }
}
template <int NCM>
__device__ __forceinline__ void spd_solve(int Nc, const double* LU, const double* b, double* x) {
  double y[NCM];
#pragma unroll
  for (int i = 0; i < NCM; ++i)
    if (i < Nc) {
      double acc = b[i];
#pragma unroll
      for (int k = 0; k < i; ++k) acc = fma(-LU[i * NCM + k], y[k], acc);
      y[i] = acc;
    }
#pragma unroll
  for (int i = NCM - 1; i >= 0; --i)
    if (i < Nc) {
      double acc = y[i];
#pragma unroll
      for (int k = i + 1; k < NCM; ++k)
        if (k < Nc) acc = fma(-LU[i * NCM + k], x[k], acc);
      x[i] = acc * LU[i * NCM + i];
    }
}

// Hildreth dual coordinate ascent (hildreth.m:16-46) in rank-Nc form.
// H(i,:)*lambda = M(i,:)*v with v = (E\M')*lambda kept incrementally.  A row with H_ii == 0 is an
// all-zero M row (Q7): it never touches v, and while it holds +Inf every row product is NaN,
// exactly what 0*Inf yields in the dense form ("poison").  E, M row-major with leading dim NCM.
template <int NCM, int NCONM>
__device__ __forceinline__ int hildreth_core(int Nc, int nC, const double* LU, const double* F, const double* Mm, const double* gam,
                                             int maxIter, double* lam /* in: warm start, out */, double* DU) {
  // small duals are fully unrolled so that lambda and all row indices are static
  constexpr int UNR = (NCONM <= 32) ? NCONM : 4;
  const int nrows = (NCONM <= 32) ? NCONM : nC;
  double P[NCONM * NCM], K[NCONM], Hd[NCONM], rH[NCONM], v[NCM], x[NCM];
  unsigned long long zrow[(NCONM + 63) / 64];                        // rows with H_ii == 0
#pragma unroll
  for (int w = 0; w < (NCONM + 63) / 64; ++w) zrow[w] = 0ull;
  spd_solve<NCM>(Nc, LU, F, x);                                      // E\F
  double Ei[NCM * NCM];                                              // inv(E), column by column: E\M' = inv(E)*M'
#pragma unroll
  for (int l = 0; l < NCM; ++l)
    if (l < Nc) {
      double el[NCM], col[NCM];
#pragma unroll
      for (int j = 0; j < NCM; ++j) el[j] = (j == l) ? 1.0 : 0.0;
      spd_solve<NCM>(Nc, LU, el, col);
#pragma unroll
      for (int j = 0; j < NCM; ++j)
        if (j < Nc) Ei[j * NCM + l] = col[j];
    }
#pragma unroll
  for (int j = 0; j < NCM; ++j) v[j] = 0.0;
  int npoison = 0;
#pragma unroll UNR
  for (int i = 0; i < nrows; ++i)
    if (i < nC) {
#pragma unroll
      for (int j = 0; j < NCM; ++j)
        if (j < Nc) {                                                // column i of E\M'
          double acc = Ei[j * NCM] * Mm[i * NCM];
#pragma unroll
          for (int l = 1; l < NCM; ++l)
            if (l < Nc) acc = fma(Ei[j * NCM + l], Mm[i * NCM + l], acc);
          P[i * NCM + j] = acc;
        }
      double hd = Mm[i * NCM] * P[i * NCM], k = Mm[i * NCM] * x[0];
#pragma unroll
      for (int j = 1; j < NCM; ++j)
        if (j < Nc) { hd = fma(Mm[i * NCM + j], P[i * NCM + j], hd); k = fma(Mm[i * NCM + j], x[j], k); }
      Hd[i] = hd;                                                    // H(i,i)  hildreth.m:28
      rH[i] = rcp_d(hd);                                             // the sweep multiplies by 1/H(i,i) (+Inf on a zero row)
      K[i] = k + gam[i];                                             // K(i)    hildreth.m:29
      if (hd == 0.0) {
        zrow[i >> 6] |= 1ull << (i & 63);
        if (!is_finite(lam[i])) ++npoison;
      } else {
#pragma unroll
        for (int j = 0; j < NCM; ++j)
          if (j < Nc) v[j] = fma(P[i * NCM + j], lam[i], v[j]);
      }
    }
  int k = 1;
  for (; k <= maxIter; ++k) {                                        // hildreth.m:32
    bool notconv = false;
#pragma unroll UNR
    for (int i = 0; i < nrows; ++i)                                  // hildreth.m:34-37
      if (i < nC) {
        const double li = lam[i];
        double w;
        if (npoison > 0) {
          w = NAN;
        } else {
          double hl = Mm[i * NCM] * v[0];
#pragma unroll
          for (int j = 1; j < NCM; ++j)
            if (j < Nc) hl = fma(Mm[i * NCM + j], v[j], hl);
          w = fma(-(K[i] + hl), rH[i], li);                           // == -(K + H(i,:)*lam - H_ii*lam_i)/H_ii
        }
        const double nl = (w > 0.0) ? w : 0.0;                       // max(0,w); max(0,NaN) = 0
        const double dl = nl - li;
        if ((zrow[i >> 6] >> (i & 63)) & 1ull) {
          npoison += (int)(!is_finite(nl)) - (int)(!is_finite(li));
        } else {
#pragma unroll
          for (int j = 0; j < NCM; ++j)
            if (j < Nc) v[j] = fma(P[i * NCM + j], dl, v[j]);
        }
        lam[i] = nl;
        if (!(fabs(dl) < 1e-6)) notconv = true;                      // hildreth.m:39 norm(.,Inf) < 1e-6, NaN-propagating
      }
    if (!notconv) break;
  }
  if (k > maxIter) k = maxIter;
  double rhs[NCM];
#pragma unroll
  for (int j = 0; j < NCM; ++j) rhs[j] = 0.0;
  bool first = true;
#pragma unroll UNR
  for (int i = 0; i < nrows; ++i)
    if (i < nC) {                                                    // hildreth.m:46  M'*lambda, rows in order
#pragma unroll
      for (int j = 0; j < NCM; ++j)
        if (j < Nc) {
          rhs[j] = first ? Mm[i * NCM + j] * lam[i] : fma(Mm[i * NCM + j], lam[i], rhs[j]);
        }
      first = false;
    }
#pragma unroll
  for (int j = 0; j < NCM; ++j)
    if (j < Nc) rhs[j] = F[j] + rhs[j];
  spd_solve<NCM>(Nc, LU, rhs, x);
#pragma unroll
  for (int j = 0; j < NCM; ++j)
    if (j < Nc) DU[j] = -x[j];
  return k;
}

// Everything iterMPC needs before the QP: F, E (factored), the unconstrained DU and the stacked
// constraints M, gamma in the reference's row order.
template <int NPM, int NCM>
struct MpcProblem {
  static constexpr int NCONM = 4 * NCM + 3 * NPM + 1;   // + the optional terminal row
  double F[NCM], E[NCM * NCM], LU[NCM * NCM], DU[NCM];
  double Mm[NCONM * NCM], gam[NCONM];
  int nC;
  double hs[NPM + 1], err[NPM], Ru;   // impulse response of the SOC row (G_soc entries), Ref - Phi_soc*dx, control weight
};

struct Limits {
  double u_max, u_min, du_max, du_min, v_max, phise_min, z_max;
};

// iterMPC.m:17-48 and constraintsMPC.m:18-101
template <int NPM, int NCM>
__device__ __forceinline__ void mpc_setup(const Dims& d, double r /* Csoc(end), EKFmatsHandler.m:43 */, const Limits& lim, const EkfOut& e,
                                          double uk_1, MpcProblem<NPM, NCM>& q) {
  const int Np = d.Np, Nc = d.Nc;
  const double zero5[N5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  double pxs[NPM];
  double (&hs)[NPM + 1] = q.hs;
  double (&err)[NPM] = q.err;
  pred_row<NPM>(e.a, zero5, r, 0.0, e.xhat, Np, hs, pxs);            // iterMPC.m:33
  double x[NCM];
#pragma unroll
  for (int i = 1; i <= NPM; ++i)
    if (i <= Np) err[i - 1] = d.ref - (pxs[i - 1] + hs[i] * uk_1);   // Ref - Phi_soc*dx
#pragma unroll
  for (int j = 0; j < NCM; ++j)
    if (j < Nc) {                                                    // F = -2*G'*(...)  iterMPC.m:38
      double acc = 0.0;
      bool first = true;
#pragma unroll
      for (int i = 1; i <= NPM; ++i)
        if (i <= Np) {
          const double g = (i - 1 - j >= 0) ? hs[i - 1 - j] : 0.0;
          const double t = g * err[i - 1];
          acc = first ? t : acc + t;
          first = false;
        }
      q.F[j] = -2.0 * acc;
    }
  double nf = q.F[0] * q.F[0];
#pragma unroll
  for (int j = 1; j < NCM; ++j)
    if (j < Nc) nf = nf + q.F[j] * q.F[j];
  const double Ru = sqrt(nf) / ((2.0 * lim.du_max) * sqrt((double)Nc)) - d.sigmin;   // iterMPC.m:43 (Q10)
  q.Ru = Ru;
#pragma unroll
  for (int j = 0; j < NCM; ++j)
#pragma unroll
    for (int l = 0; l < NCM; ++l) {
      double acc = 0.0;
      if (j < Nc && l < Nc) {
        bool first = true;
#pragma unroll
        for (int i = 1; i <= NPM; ++i)
          if (i <= Np) {
            const double gj = (i - 1 - j >= 0) ? hs[i - 1 - j] : 0.0, gl = (i - 1 - l >= 0) ? hs[i - 1 - l] : 0.0;
            const double t = gj * gl;
            acc = first ? t : acc + t;
            first = false;
          }
        acc = 2.0 * (acc + ((j == l) ? Ru : 0.0));                   // iterMPC.m:47
      }
      q.E[j * NCM + l] = acc;
    }
  spd_factor<NCM>(Nc, q.E, q.LU);
  spd_solve<NCM>(Nc, q.LU, q.F, x);
#pragma unroll
  for (int j = 0; j < NCM; ++j) q.DU[j] = (j < Nc) ? -x[j] : 0.0;    // iterMPC.m:48
  // ---- constraintsMPC.m: rows in the reference's order ----
  double* Mm = q.Mm;
  double* gam = q.gam;
  int nC = 0;
  if (d.useCurrent) {                                                // constraintsMPC.m:23-42
    for (int s = 0; s < 2; ++s)
      for (int i = 0; i < Nc; ++i, ++nC) {
#pragma unroll
        for (int j = 0; j < NCM; ++j) Mm[nC * NCM + j] = (j <= i && j < Nc) ? (s ? -1.0 : 1.0) : (s ? -0.0 : 0.0);
        gam[nC] = s ? -(lim.u_min - uk_1) : (lim.u_max - uk_1);
      }
    for (int s = 0; s < 2; ++s)
      for (int i = 0; i < Nc; ++i, ++nC) {
#pragma unroll
        for (int j = 0; j < NCM; ++j) Mm[nC * NCM + j] = (j == i) ? (s ? -1.0 : 1.0) : (s ? -0.0 : 0.0);
        gam[nC] = s ? -lim.du_min : lim.du_max;
      }
  }
  if (d.useVoltage) {                                                // constraintsMPC.m:45-63
    double hv[NPM + 1], pxv[NPM];
    pred_row<NPM>(e.a, e.Cv, 0.0, e.Dv, e.xhat, Np, hv, pxv);
    for (int i = 1; i <= Np; ++i, ++nC) {
#pragma unroll
      for (int j = 0; j < NCM; ++j) Mm[nC * NCM + j] = (i - 1 - j >= 0 && j < Nc) ? hv[i - 1 - j] : 0.0;
      gam[nC] = lim.v_max - ((pxv[i - 1] + hv[i] * uk_1) + e.bv);
    }
  }
  if (d.useEta) {                                                    // constraintsMPC.m:66-85
    double he[NPM + 1], pxe[NPM];
    pred_row<NPM>(e.a, e.Cphi, 0.0, e.Dphi, e.xhat, Np, he, pxe);
    for (int i = 1; i <= Np; ++i, ++nC) {
#pragma unroll
      for (int j = 0; j < NCM; ++j) Mm[nC * NCM + j] = (i - 1 - j >= 0 && j < Nc) ? -he[i - 1 - j] : -0.0;
      gam[nC] = -lim.phise_min + ((pxe[i - 1] + he[i] * uk_1) + e.bphi);
    }
  }
  if (d.useZmax) {                                                   // constraintsMPC.m:89-101
    for (int i = 1; i <= Np; ++i, ++nC) {
#pragma unroll
      for (int j = 0; j < NCM; ++j) Mm[nC * NCM + j] = (i - 1 - j >= 0 && j < Nc) ? hs[i - 1 - j] : 0.0;
      gam[nC] = lim.z_max - ((pxs[i - 1] + hs[i] * uk_1) + e.soc);
    }
  }
  if (d.useTerm) {                                                   // constraintsMPC.m:104-112: -G_s(end,:)*DU <= -z_max + Phi_s(end,:)*x_aug
#pragma unroll
    for (int j = 0; j < NCM; ++j) Mm[nC * NCM + j] = (Np - 1 - j >= 0 && j < Nc) ? -hs[Np - 1 - j >= 0 ? Np - 1 - j : 0] : -0.0;
    gam[nC] = -(lim.z_max - d.z_tol) + (pxs[Np - 1] + hs[Np] * uk_1);
    ++nC;
  }
  q.nC = nC;
}

struct MpcOut {
  double u, phise;
  int nexec;
  unsigned active[4];
};

// mpcData.cost (iterMPC.m:50-51,82-95): J = (Ref - Phi*dx - G*DU)'*(Ref - Phi*dx - G*DU) + DU'*Ru*DU
template <int NPM, int NCM>
__device__ __forceinline__ double mpc_cost_J(int Np, int Nc, const MpcProblem<NPM, NCM>& q) {
  double J = 0.0, n2 = 0.0;
  for (int i = 1; i <= Np; ++i) {
    double y = q.err[i - 1];
    for (int j = 0; j < Nc; ++j)
      if (i - 1 - j >= 0) y = y - q.hs[i - 1 - j] * q.DU[j];
    J = J + y * y;
  }
  for (int j = 0; j < Nc; ++j) n2 = n2 + q.DU[j] * q.DU[j];
  return J + q.Ru * n2;
}

// iterMPC.m:17-95 with constraintsMPC.m and hildreth.m inlined; the dual warm start lives in the cell state
template <int NPM, int NCM>
__device__ __forceinline__ void mpc_step(const Dims& d, const double* __restrict__ p, const Cell& c, const EkfOut& e, double uk_applied,
                                         MpcOut& o, double* cost = nullptr /* J_uncon, J_final, norm_DU, viol */) {
  constexpr int NCONM = MpcProblem<NPM, NCM>::NCONM;
  const int Nc = d.Nc;
  const double uk_1 = c.f(S_UK);                                     // iterMPC.m:19
  {                                                                  // runMPC.m:94-95 (Q24)
    double acc = e.Cphi[0] * e.xhat[0];
#pragma unroll
    for (int j = 1; j < N5; ++j) acc = fma(e.Cphi[j], e.xhat[j], acc);
    acc = acc + 0.0 * e.xhat[5];
    o.phise = (acc + uk_applied * e.Dphi) + e.bphi;
  }
  Limits lim = {c.f(S_UMAX), c.f(S_UMIN), c.f(S_DUMAX), c.f(S_DUMIN), c.f(S_VMAX), c.f(S_PHIMIN), c.f(S_ZMAX)};
  MpcProblem<NPM, NCM> q;
  const double r = -d.Ts / (3600.0 * p[P_Q]);                        // EKFmatsHandler.m:43
  mpc_setup<NPM, NCM>(d, r, lim, e, uk_1, q);
  const int nC = q.nC;
  if (cost) cost[0] = mpc_cost_J<NPM, NCM>(d.Np, Nc, q);             // iterMPC.m:50-51
  bool viol = false;                                                 // iterMPC.m:66 (Q8)
  for (int i = 0; i < nC; ++i) {
    double acc = q.Mm[i * NCM] * q.DU[0];
#pragma unroll
    for (int j = 1; j < NCM; ++j)
      if (j < Nc) acc = fma(q.Mm[i * NCM + j], q.DU[j], acc);
    if (acc - q.gam[i] > 0.0) viol = true;
  }
  o.nexec = 0;
#pragma unroll
  for (int w = 0; w < 4; ++w) o.active[w] = 0u;
  if (viol) {
    constexpr int UNR = (NCONM <= 32) ? NCONM : 4;
    const int nrows = (NCONM <= 32) ? NCONM : nC;
    double lam[NCONM];
    const bool warm = c.f(S_LAMVALID) != 0.0;                        // Q11: [] -> zeros on first use
#pragma unroll UNR
    for (int i = 0; i < nrows; ++i)
      if (i < nC) lam[i] = warm ? c.lam(i) : 0.0;
    o.nexec = hildreth_core<NCM, NCONM>(Nc, nC, q.LU, q.F, q.Mm, q.gam, d.maxHild, lam, q.DU);   // iterMPC.m:67
#pragma unroll UNR
    for (int i = 0; i < nrows; ++i)
      if (i < nC) {
        c.lam(i) = lam[i];                                // iterMPC.m:68
        if (lam[i] > 0.0) o.active[i >> 5] |= (1u << (i & 31));
      }
    c.f(S_LAMVALID) = 1.0;
  }
  if (cost) {                                                        // iterMPC.m:82-95
    int nviol = 0;
    double n2 = 0.0;
    for (int i = 0; i < nC; ++i) {
      double acc = 0.0;
      for (int j = 0; j < Nc; ++j) acc = acc + q.Mm[i * NCM + j] * q.DU[j];
      if (acc - q.gam[i] > 1e-9) ++nviol;
    }
    for (int j = 0; j < Nc; ++j) n2 = n2 + q.DU[j] * q.DU[j];
    cost[1] = mpc_cost_J<NPM, NCM>(d.Np, Nc, q);
    cost[2] = sqrt(n2);
    cost[3] = (double)nviol;
  }
  const double uk = q.DU[0] + uk_1;                                  // iterMPC.m:75-76
  c.f(S_UK) = uk;                                                    // iterMPC.m:78
  o.u = uk;
  if (!is_finite(uk)) c.f(S_STATUS) = (double)((unsigned)c.f(S_STATUS) | ST_NONFINITE);
}

// ------------------------------------------------------------------------------------------
// Static-shape fast path of iterMPC: Np == NP, Nc == NC and the constraint switches CONS
// (bit0 current, bit1 voltage, bit2 eta, bit3 z_max) are compile-time, so every row index is static:
// M, gamma and lambda live in registers and the per-row sweep data (K_i, 1/H_ii, column i of E\M')
// in this thread's column of shared memory ([slot][thread] layout: conflict-free, no local memory).
// Same arithmetic, in the same order, as mpc_step / hildreth_core.
// ------------------------------------------------------------------------------------------
// Shared-memory accesses as asm volatile: re-executed every sweep (never cached in a register across sweeps) without the
// ld.volatile / st.volatile semantics a C++ volatile access would carry.
// K_i and 1/H_ii of one row as a 16-byte pair: one LDS.128, one exposed shared-memory latency per row instead of two
__device__ __forceinline__ void lds_pair(unsigned addr, double& a, double& b) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr) : "memory");
}
__device__ __forceinline__ void sts_pair(unsigned addr, double a, double b) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
template <int NP, int NC, int CONS>
struct FastShape {
  static constexpr bool UC = (CONS & 1) != 0, UV = (CONS & 2) != 0, UE = (CONS & 4) != 0, UZ = (CONS & 8) != 0;
  static constexpr int O_V = UC ? 4 * NC : 0;
  static constexpr int O_E = O_V + (UV ? NP : 0);
  static constexpr int O_Z = O_E + (UE ? NP : 0);
  static constexpr int nC = O_Z + (UZ ? NP : 0);
  static constexpr int SLOTS = nC * 2;                               // doubles of shared memory per thread: K_i, 1/H_ii
};

// Tells the compiler that inv(E) may have changed: column i of E\M' = inv(E)*M(i,:)' is then RECOMPUTED in every sweep
// (four FP64 operations per row) instead of being computed once for all rows and kept -- 44 values that do not fit the
// register file and would be spilled and re-read from local memory in every sweep.
template <int N>
__device__ __forceinline__ void opaque(double (&A)[N][N]) {
#ifdef __CUDA_ARCH__
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) asm volatile("" : "+d"(A[i][j]));
#endif
}

template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

template <int NP, int NC, int CONS>
__device__ __forceinline__ void mpc_step_fast(const Dims& d, const double* __restrict__ p, const Cell& c, const EkfOut& e, double uk_applied,
                                              MpcOut& o, double* smt, int sms) {
  using FS = FastShape<NP, NC, CONS>;
  constexpr int nC = FS::nC;
  static_assert(nC <= 32, "fast path keeps the zero-row mask in one word");
  const double uk_1 = c.f(S_UK);                                     // iterMPC.m:19
  {                                                                  // runMPC.m:94-95 (Q24)
    double acc = e.Cphi[0] * e.xhat[0];
#pragma unroll
    for (int j = 1; j < N5; ++j) acc = fma(e.Cphi[j], e.xhat[j], acc);
    acc = acc + 0.0 * e.xhat[5];
    o.phise = (acc + uk_applied * e.Dphi) + e.bphi;
  }
  const double r = -d.Ts / (3600.0 * p[P_Q]);                        // EKFmatsHandler.m:43
  // The stacked constraint matrix M (constraintsMPC.m:99) is never stored: the current rows are +-1 / +-0 patterns
  // (:23-42) and every row of the voltage / eta / SOC blocks is a window [h(i-1) ... h(i-Nc)] of that output's impulse
  // response (:47-53,68-75,92-99), so three tables of Np entries stand for all of it -- 15 registers instead of 46 for M
  // plus 23 for gamma, which waits in the K slots of shared memory until the QP set-up turns it into K_i.
  double hs[NP + 1], hv[NP + 1], nhe[NP + 1];                        // h_soc, h_v, -h_eta
  double F[NC], LU[NC * NC], DU[NC], x[NC];
  // (K_i, 1/H_ii) as 16-byte pairs, re-read from shared memory every sweep (keeps them out of registers).  Row i lives
  // in scratch slots 2i and 2i+1, but inside the slots a WARP's 64 doubles are dealt out as 32 pairs: lanes
  // 0-15 take slot 2i, lanes 16-31 slot 2i+1, each the two columns 2(l mod 16), +1 of its own warp.  A warp therefore
  // still touches only its own columns (other warps may be in another phase of the instant and use theirs as scratch),
  // every quarter-warp reads 128 contiguous bytes, and the region starts 16-byte aligned (the bank blob is padded).
  const unsigned lane = threadIdx.x & 31u, wbase = threadIdx.x & ~31u;
  const unsigned kra = (unsigned)__cvta_generic_to_shared(smt - threadIdx.x) + ((lane >> 4) * (unsigned)sms + wbase + 2u * (lane & 15u)) * 8u;
  const unsigned krs = (unsigned)sms * 16u;
  // row i of M from the tables (i is a compile-time constant: every entry is a register or an immediate)
  auto mrow = [&](auto ic, double (&Mi)[NC]) {
    constexpr int i = decltype(ic)::value;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      if constexpr (FS::UC && i < 2 * NC) {                          // [Cu; -Cu], Cu lower-triangular ones (:23-32)
        constexpr int s = i / NC, q = i % NC;
        Mi[j] = (j <= q) ? (s ? -1.0 : 1.0) : (s ? -0.0 : 0.0);
      } else if constexpr (FS::UC && i < 4 * NC) {                   // [I; -I] (:33-42)
        constexpr int s = (i - 2 * NC) / NC, q = (i - 2 * NC) % NC;
        Mi[j] = (j == q) ? (s ? -1.0 : 1.0) : (s ? -0.0 : 0.0);
      } else if constexpr (FS::UV && i >= FS::O_V && i < FS::O_E) {
        constexpr int q = i - FS::O_V;
        Mi[j] = (q - j >= 0) ? hv[q - j >= 0 ? q - j : 0] : 0.0;
      } else if constexpr (FS::UE && i >= FS::O_E && i < FS::O_Z) {
        constexpr int q = i - FS::O_E;
        Mi[j] = (q - j >= 0) ? nhe[q - j >= 0 ? q - j : 0] : -0.0;
      } else {
        constexpr int q = i - FS::O_Z;
        Mi[j] = (q - j >= 0) ? hs[q - j >= 0 ? q - j : 0] : 0.0;
      }
    }
  };
  bool viol = false;                                                 // iterMPC.m:66 (Q8)
  {
    const double zero5[N5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    double pxs[NP], err[NP], E[NC * NC];
    pred_row<NP>(e.a, zero5, r, 0.0, e.xhat, NP, hs, pxs);           // iterMPC.m:33
#pragma unroll
    for (int i = 1; i <= NP; ++i) err[i - 1] = d.ref - (pxs[i - 1] + hs[i] * uk_1);
#pragma unroll
    for (int j = 0; j < NC; ++j) {                                   // iterMPC.m:38
      double acc = 0.0;
#pragma unroll
      for (int i = 1; i <= NP; ++i) {
        const double g = (i - 1 - j >= 0) ? hs[i - 1 - j] : 0.0;
        const double t = g * err[i - 1];
        acc = (i == 1) ? t : acc + t;
      }
      F[j] = -2.0 * acc;
    }
    double nf = F[0] * F[0];
#pragma unroll
    for (int j = 1; j < NC; ++j) nf = nf + F[j] * F[j];
    const double Ru = sqrt(nf) / ((2.0 * c.f(S_DUMAX)) * sqrt((double)NC)) - d.sigmin;   // iterMPC.m:43 (Q10)
#pragma unroll
    for (int j = 0; j < NC; ++j)
#pragma unroll
      for (int l = 0; l < NC; ++l) {
        double acc = 0.0;
#pragma unroll
        for (int i = 1; i <= NP; ++i) {
          const double gj = (i - 1 - j >= 0) ? hs[i - 1 - j] : 0.0, gl = (i - 1 - l >= 0) ? hs[i - 1 - l] : 0.0;
          const double t = gj * gl;
          acc = (i == 1) ? t : acc + t;
        }
        E[j * NC + l] = 2.0 * (acc + ((j == l) ? Ru : 0.0));         // iterMPC.m:47
      }
    spd_factor<NC>(NC, E, LU);
    spd_solve<NC>(NC, LU, F, x);
#pragma unroll
    for (int j = 0; j < NC; ++j) DU[j] = -x[j];                      // iterMPC.m:48
    // ---- constraintsMPC.m: gamma row by row; each row is tested against the unconstrained DU at once (iterMPC.m:66) and
    //      its gamma parked in the row's K slot ----
    auto park = [&](auto ic, double gam) {
      constexpr int i = decltype(ic)::value;
      double Mi[NC];
      mrow(ic, Mi);
      double acc = Mi[0] * DU[0];
#pragma unroll
      for (int j = 1; j < NC; ++j) acc = fma(Mi[j], DU[j], acc);
      if (acc - gam > 0.0) viol = true;
      asm volatile("st.shared.f64 [%0], %1;" ::"r"(kra + (unsigned)i * krs), "d"(gam) : "memory");
    };
    if constexpr (FS::UC) {                                          // constraintsMPC.m:23-42
      const double umax = c.f(S_UMAX), umin = c.f(S_UMIN), dumax = c.f(S_DUMAX), dumin = c.f(S_DUMIN);
      static_for<0, 4 * NC>([&](auto ic) {
        constexpr int blk = decltype(ic)::value / NC;
        park(ic, blk == 0 ? (umax - uk_1) : (blk == 1 ? -(umin - uk_1) : (blk == 2 ? dumax : -dumin)));
      });
    }
    if constexpr (FS::UV) {                                          // constraintsMPC.m:45-63
      double pxv[NP];
      pred_row<NP>(e.a, e.Cv, 0.0, e.Dv, e.xhat, NP, hv, pxv);
      const double vmax = c.f(S_VMAX);
      static_for<0, NP>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        park(std::integral_constant<int, FS::O_V + q>{}, vmax - ((pxv[q] + hv[q + 1] * uk_1) + e.bv));
      });
    }
    if constexpr (FS::UE) {                                          // constraintsMPC.m:66-85
      double he[NP + 1], pxe[NP];
      pred_row<NP>(e.a, e.Cphi, 0.0, e.Dphi, e.xhat, NP, he, pxe);
      const double pmin = c.f(S_PHIMIN);
#pragma unroll
      for (int i = 0; i <= NP; ++i) nhe[i] = -he[i];
      static_for<0, NP>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        park(std::integral_constant<int, FS::O_E + q>{}, -pmin + ((pxe[q] + he[q + 1] * uk_1) + e.bphi));
      });
    }
    if constexpr (FS::UZ) {                                          // constraintsMPC.m:89-101
      const double zmax = c.f(S_ZMAX);
      static_for<0, NP>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        park(std::integral_constant<int, FS::O_Z + q>{}, zmax - ((pxs[q] + hs[q + 1] * uk_1) + e.soc));
      });
    }
  }
  o.nexec = 0;
#pragma unroll
  for (int w = 0; w < 4; ++w) o.active[w] = 0u;
  if (viol) {                                                        // hildreth.m:16-46 (see hildreth_core)
    double lam[nC], v[NC];
    const bool warm = c.f(S_LAMVALID) != 0.0;                        // Q11
#pragma unroll
    for (int i = 0; i < nC; ++i) lam[i] = warm ? ld_row(&c.lam(i)) : 0.0;
    spd_solve<NC>(NC, LU, F, x);                                     // E\F
    double Ei[NC][NC];                                               // inv(E): column i of E\M' is recomputed as inv(E)*M(i,:)'
#pragma unroll
    for (int l = 0; l < NC; ++l) {
      double el[NC], col[NC];
#pragma unroll
      for (int j = 0; j < NC; ++j) el[j] = (j == l) ? 1.0 : 0.0;
      spd_solve<NC>(NC, LU, el, col);
#pragma unroll
      for (int j = 0; j < NC; ++j) Ei[j][l] = col[j];
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) v[j] = 0.0;
    unsigned zmask = 0u;
    double pchk = 0.0;                                               // stays (+-)0 while every entry of E\M' is finite: Inf*0 = NaN sticks
    bool hpos = true;                                                // H_ii > 0 on every row that is not all-zero
#ifdef MEK_SPEC_ON
    bool rcp_ok = true;
#endif
    static_for<0, nC>([&](auto ic) {
      constexpr int i = decltype(ic)::value;
      double Mi[NC], P[NC];
      mrow(ic, Mi);
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        double acc = Ei[j][0] * Mi[0];
#pragma unroll
        for (int l = 1; l < NC; ++l) acc = fma(Ei[j][l], Mi[l], acc);
        P[j] = acc;
        pchk = fma(acc, 0.0, pchk);
      }
      double hd = Mi[0] * P[0], k = Mi[0] * x[0];
#pragma unroll
      for (int j = 1; j < NC; ++j) { hd = fma(Mi[j], P[j], hd); k = fma(Mi[j], x[j], k); }
      double gam;
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(gam) : "r"(kra + (unsigned)i * krs) : "memory");
#ifdef MEK_SPEC_ON
      // 1/H(i,i) by the speculative quotient (mek_math.cuh): the 23 reciprocals are independent and share one basic block
      // instead of one slow-path branch each; the structurally all-zero row keeps the standard operator (1/0 = +Inf)
      double rh;
      if constexpr (FS::UZ && i == FS::O_Z) rh = rcp_d(hd); else rh = div_spec(1.0, hd, rcp_ok);
      sts_pair(kra + (unsigned)i * krs, -(k + gam), rh);             // -K(i), 1/H(i,i): the idle test below compares against -K
#else
      sts_pair(kra + (unsigned)i * krs, -(k + gam), rcp_d(hd));      // -K(i), 1/H(i,i): the idle test below compares against -K
#endif
      if (hd == 0.0) {
        zmask |= 1u << i;
      } else {
        hpos = hpos && (hd > 0.0);
#pragma unroll
        for (int j = 0; j < NC; ++j) v[j] = fma(P[j], lam[i], v[j]);
      }
    });
#ifdef MEK_SPEC_ON
    if (!rcp_ok) {                                                   // some H(i,i) outside the fast sequence's range: the standard reciprocals
      static_for<0, nC>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        double Mi[NC];
        mrow(ic, Mi);
        double hd = 0.0;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          double acc = Ei[j][0] * Mi[0];
#pragma unroll
          for (int l = 1; l < NC; ++l) acc = fma(Ei[j][l], Mi[l], acc);
          hd = (j == 0) ? Mi[0] * acc : fma(Mi[j], acc, hd);
        }
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(kra + (unsigned)i * krs + 8u), "d"(rcp_d(hd)) : "memory");
      });
    }
#endif
    // rows whose lambda is non-finite (only all-zero rows can get there): while any is set every row product is NaN
    unsigned pmask = 0u;
    constexpr unsigned ZEXP = FS::UZ ? (1u << FS::O_Z) : 0u;         // the one structurally all-zero row: first z_max row (Q7)
    if (FS::UZ && (zmask & ZEXP) && !is_finite(lam[FS::O_Z])) pmask |= ZEXP;
    if (zmask & ~ZEXP) {                                             // any other all-zero row (degenerate linearisation)
#pragma unroll
      for (int i = 0; i < nC; ++i)
        if (((zmask >> i) & 1u) && !is_finite(lam[i])) pmask |= 1u << i;
    }
    const bool pfin = pchk == 0.0;
    int k = 1;
    // Lean sweep: the only all-zero row is the first z_max row (h_soc(0) = 0 always, quirk Q7), it holds a finite lambda and
    // its K (= gamma_i) is not negative.  Then that row's update is max(0, -K*Inf) = 0 on every sweep (K = 0: NaN -> 0), it
    // never touches v, no lambda ever becomes non-finite ("poison"), and every other row is the plain update -- the masks
    // of the general loop below are all constant and drop out.  Bit-identical to the general loop by construction
    // (tests/test_gpu_parity.py::test_lean_sweep_equals_general_sweep forces the general loop for comparison).
    bool lean = (zmask == ZEXP) && (pmask == 0u) && !d.no_lean_sweep;
    if (FS::UZ) {
      double kz, hz;
      lds_pair(kra + (unsigned)FS::O_Z * krs, kz, hz);                 // kz = -K of that row
      lean = lean && !(kz > 0.0);
    }
    if (lean) {
      const bool skip_ok = pfin && hpos && !d.no_row_skip;
      bool first_nc = false;
      if (FS::UZ) { first_nc = !(lam[FS::O_Z] < 1e-6); lam[FS::O_Z] = 0.0; }
      for (; k <= d.maxHild; ++k) {                                  // hildreth.m:32
        bool notconv = (k == 1) && first_nc;
        opaque(Ei);
        static_for<0, nC>([&](auto ic) {                             // hildreth.m:34-37
          constexpr int i = decltype(ic)::value;
          if constexpr (!(FS::UZ && i == FS::O_Z)) {
            double Mi[NC];
            mrow(ic, Mi);
            double nKi, rHi;
            lds_pair(kra + (unsigned)i * krs, nKi, rHi);
            const double li = lam[i];
            double hl = Mi[0] * v[0];
#pragma unroll
            for (int j = 1; j < NC; ++j) hl = fma(Mi[j], v[j], hl);
            // An inactive row that stays inactive: lambda_i = 0 and t = K_i + H(i,:)*lambda >= 0 give w = -t/H_ii <= 0 (H_ii > 0),
            // max(0,w) = 0, lambda_i and v exactly as they were (v + P*0 = v for finite P; v is never -0).  Most rows of most
            // sweeps are like that; the cell skips the rest of the row (a plain per-cell branch: a warp vote on the predicate cost
            // three more instructions per row than the divergence it avoided).  The test reads
            // H(i,:)*lambda >= -K_i -- the sign of the rounded sum t is the sign of the exact one -- which takes the addition off
            // the dependent chain of an idle row (if K_i and H(i,:)*lambda are infinities of opposite sign, t is NaN where this
            // test says idle: the update of such a row with lambda_i = 0 is max(0, NaN) = 0, a no-op as well).
            const bool idle = skip_ok && (li == 0.0) && (hl >= nKi);
            if (!idle) {
              const double t = hl - nKi;                             // = K_i + H(i,:)*lambda, bit for bit
              const double w = fma(-t, rHi, li);
              const double nl = pos_or_zero(w);                      // max(0,w); max(0,NaN) = 0
              const double dl = nl - li;
#pragma unroll
              for (int j = 0; j < NC; ++j) {
                double acc = Ei[j][0] * Mi[0];
#pragma unroll
                for (int l = 1; l < NC; ++l) acc = fma(Ei[j][l], Mi[l], acc);
                v[j] = fma(acc, dl, v[j]);
              }
              lam[i] = nl;
              notconv = notconv || !(fabs(dl) < 1e-6);               // hildreth.m:39
            }
          }
        });
        if (!notconv) break;
      }
    } else
    for (; k <= d.maxHild; ++k) {                                    // hildreth.m:32
      bool notconv = false;
      opaque(Ei);
      static_for<0, nC>([&](auto ic) {                               // hildreth.m:34-37, branch-free
        constexpr int i = decltype(ic)::value;
        double Mi[NC];
        mrow(ic, Mi);
        const double li = lam[i];
        double hl = Mi[0] * v[0];
#pragma unroll
        for (int j = 1; j < NC; ++j) hl = fma(Mi[j], v[j], hl);
        double nKg, rHg;
        lds_pair(kra + (unsigned)i * krs, nKg, rHg);
        double w = fma(-(hl - nKg), rHg, li);                        // hl - (-K) = K + hl, bit for bit
        w = (pmask != 0u) ? NAN : w;
        const double nl = (w > 0.0) ? w : 0.0;                       // max(0,w); max(0,NaN) = 0
        const double dl = nl - li;
        const bool zr = ((zmask >> i) & 1u) != 0u;
        const double pd = zr ? 0.0 : dl;                             // an all-zero row never touches v
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          double acc = Ei[j][0] * Mi[0];
#pragma unroll
          for (int l = 1; l < NC; ++l) acc = fma(Ei[j][l], Mi[l], acc);
          v[j] = fma(acc, pd, v[j]);
        }
        pmask = (zr && !(nl <= kDblMax)) ? (pmask | (1u << i)) : (pmask & ~(1u << i));
        lam[i] = nl;
        if (!(fabs(dl) < 1e-6)) notconv = true;                      // hildreth.m:39 norm(.,Inf) < 1e-6, NaN-propagating
      });
      if (!notconv) break;
    }
    if (k > d.maxHild) k = d.maxHild;
    double rhs[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) rhs[j] = 0.0;
    static_for<0, nC>([&](auto ic) {                                 // hildreth.m:46  M'*lambda, rows in order
      constexpr int i = decltype(ic)::value;
      double Mi[NC];
      mrow(ic, Mi);
#pragma unroll
      for (int j = 0; j < NC; ++j) rhs[j] = (i == 0) ? Mi[j] * lam[i] : fma(Mi[j], lam[i], rhs[j]);
    });
#pragma unroll
    for (int j = 0; j < NC; ++j) rhs[j] = F[j] + rhs[j];
    spd_solve<NC>(NC, LU, rhs, x);
#pragma unroll
    for (int j = 0; j < NC; ++j) DU[j] = -x[j];
    o.nexec = k;                                                     // iterMPC.m:67-69
#pragma unroll
    for (int i = 0; i < nC; ++i) {
      st_row(&c.lam(i), lam[i]);
      if (lam[i] > 0.0) o.active[i >> 5] |= (1u << (i & 31));
    }
    c.f(S_LAMVALID) = 1.0;
  }
  const double uk = DU[0] + uk_1;                                    // iterMPC.m:75-76
  c.f(S_UK) = uk;
  o.u = uk;
  if (!is_finite(uk)) c.f(S_STATUS) = (double)((unsigned)c.f(S_STATUS) | ST_NONFINITE);
}

// ------------------------------------------------------------------------------------------
// Static-shape path of iterMPC for long horizons (config C4: Np = 20, Nc = 5, 80 constraint rows).
// The dual is too large for registers; per thread it keeps in shared-memory columns only
//   - the three impulse-response tables h_v, -h_eta, h_soc (Np entries each): every row of the voltage / eta / SOC
//     blocks of M is a window [h(i-1), h(i-2), ..., h(i-Nc)] of one of them (constraintsMPC.m:47-53,68-75,92-99), and
//     the current rows are +-1 / +-0 patterns (:23-42), so M itself is never stored;
//   - K_i, 1/H_ii and lambda_i per row.
// Column i of E\M' is recomputed as inv(E)*M(i,:)' where it is needed -- and it is needed only when lambda_i moves:
// a row whose lambda stays put adds P*0 to v, which leaves v unchanged bit for bit as long as P is finite (v is
// never -0: it starts at +0 and only ever receives sums).  Rows are walked in the reference's order by rolled loops
// with a sliding window over the tables.  Same arithmetic, in the same order, as mpc_setup / hildreth_core.
// ------------------------------------------------------------------------------------------
#ifndef WIDE_UNROLL
#define WIDE_UNROLL 5            // Toeplitz rows per loop trip in the long-horizon sweep
#endif
template <int NP, int NC, int CONS>
struct WideShape {
  static constexpr bool UC = (CONS & 1) != 0, UV = (CONS & 2) != 0, UE = (CONS & 4) != 0, UZ = (CONS & 8) != 0;
  static constexpr int nC = (UC ? 4 * NC : 0) + NP * ((UV ? 1 : 0) + (UE ? 1 : 0) + (UZ ? 1 : 0));
  // read-mostly columns (written once per QP, read once per row and sweep): tables, K, 1/H
  static constexpr int T_V = 0, T_E = NP, T_S = 2 * NP;              // table slots (Np entries each)
  static constexpr int W_K = 3 * NP, W_H = W_K + nC;
  static constexpr int W_LU = W_H + nC, W_F = W_LU + NC * NC;        // factor of E and F parked here during the sweeps
  static constexpr int WSLOTS = W_F + NC;
  static constexpr int LSLOTS = nC;                                  // lambda column (read and written every row)
  static constexpr int SLOTS = WSLOTS + LSLOTS;                      // all of it in shared memory (small CTAs)
};

// calls f(row, M[NC]) for every constraint row in the reference's order (constraintsMPC.m:99: [current; V; eta; SOC]).
// The current rows are compile-time patterns (fully unrolled); the Toeplitz blocks slide a window over their table,
// whose entries are loaded four rows ahead (the column may live in L2), UNR rows per loop trip.
template <int NP, int NC, int CONS, int UNR, typename F>
__device__ __forceinline__ void wide_rows(const double* __restrict__ tab, int ts, F&& f) {
  using WS = WideShape<NP, NC, CONS>;
  int row = 0;
  if (WS::UC) {
#pragma unroll
    for (int blk = 0; blk < 4; ++blk) {                              // [prefix +; prefix -; identity +; identity -]
      const bool neg = (blk & 1) != 0, ident = blk >= 2;
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        double M[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          const bool on = ident ? (j == i) : (j <= i);
          M[j] = on ? (neg ? -1.0 : 1.0) : (neg ? -0.0 : 0.0);
        }
        f(row, M);
        ++row;
      }
    }
  }
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    const bool use = b == 0 ? WS::UV : (b == 1 ? WS::UE : WS::UZ);
    if (!use) continue;
    const double* t = tab + (b == 0 ? WS::T_V : (b == 1 ? WS::T_E : WS::T_S)) * ts;
    double M[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) M[j] = (b == 1) ? -0.0 : 0.0;
    static_assert(NP >= 4, "table prefetch depth");
    double n0 = t[0], n1 = t[ts], n2 = t[2 * ts], n3 = t[3 * ts];
#pragma unroll UNR
    for (int i = 0; i < NP; ++i) {
#pragma unroll
      for (int j = NC - 1; j > 0; --j) M[j] = M[j - 1];
      M[0] = n0;
      n0 = n1; n1 = n2; n2 = n3;
      n3 = t[(i + 4 < NP ? i + 4 : NP - 1) * ts];
      f(row, M);
      ++row;
    }
  }
}

template <int NP, int NC, int CONS>
__device__ __forceinline__ void mpc_step_wide(const Dims& d, const double* __restrict__ p, const Cell& c, const EkfOut& e, double uk_applied,
                                              MpcOut& o, double* lc, int ls, double* wc, int wst) {
  // lc/ls: this thread's lambda column (shared memory); wc/wst: its read-mostly column (shared memory or the L2-resident workspace)
  using WS = WideShape<NP, NC, CONS>;
  constexpr int nC = WS::nC;
  static_assert(nC <= 96, "row flags are kept in three words");
  const double uk_1 = c.f(S_UK);                                     // iterMPC.m:19
  {                                                                  // runMPC.m:94-95 (Q24)
    double acc = e.Cphi[0] * e.xhat[0];
#pragma unroll
    for (int j = 1; j < N5; ++j) acc = fma(e.Cphi[j], e.xhat[j], acc);
    acc = acc + 0.0 * e.xhat[5];
    o.phise = (acc + uk_applied * e.Dphi) + e.bphi;
  }
  const double r = -d.Ts / (3600.0 * p[P_Q]);                        // EKFmatsHandler.m:43
  double F[NC], LU[NC * NC], DU[NC], x[NC];
  {
    const double zero5[N5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    double hs[NP + 1], pxs[NP];
    pred_row<NP>(e.a, zero5, r, 0.0, e.xhat, NP, hs, pxs);           // iterMPC.m:33
    {
      double err[NP];
#pragma unroll
      for (int i = 1; i <= NP; ++i) err[i - 1] = d.ref - (pxs[i - 1] + hs[i] * uk_1);
#pragma unroll
      for (int j = 0; j < NC; ++j) {                                 // iterMPC.m:38
        double acc = 0.0;
#pragma unroll
        for (int i = 1; i <= NP; ++i) {
          if (i - 1 - j >= 0) {
            const double t = hs[i - 1 - j] * err[i - 1];
            acc = (i == 1) ? t : acc + t;
          } else {
            const double t = 0.0 * err[i - 1];
            acc = (i == 1) ? t : acc + t;
          }
        }
        F[j] = -2.0 * acc;
      }
    }
    double nf = F[0] * F[0];
#pragma unroll
    for (int j = 1; j < NC; ++j) nf = nf + F[j] * F[j];
    const double Ru = sqrt(nf) / ((2.0 * c.f(S_DUMAX)) * sqrt((double)NC)) - d.sigmin;   // iterMPC.m:43 (Q10)
    double E[NC * NC];
#pragma unroll
    for (int j = 0; j < NC; ++j)
#pragma unroll
      for (int l = 0; l < NC; ++l) {
        double acc = 0.0;
#pragma unroll
        for (int i = 1; i <= NP; ++i) {
          const double gj = (i - 1 - j >= 0) ? hs[i - 1 - j] : 0.0, gl = (i - 1 - l >= 0) ? hs[i - 1 - l] : 0.0;
          const double t = gj * gl;
          acc = (i == 1) ? t : acc + t;
        }
        E[j * NC + l] = 2.0 * (acc + ((j == l) ? Ru : 0.0));         // iterMPC.m:47
      }
    spd_factor<NC>(NC, E, LU);
    spd_solve<NC>(NC, LU, F, x);
#pragma unroll
    for (int j = 0; j < NC; ++j) DU[j] = -x[j];                      // iterMPC.m:48
    // tables and gamma (kept in the K slots until the QP set-up turns them into K_i)
    int row = 0;
    if (WS::UC) {                                                    // constraintsMPC.m:23-42
      const double umax = c.f(S_UMAX), umin = c.f(S_UMIN), dumax = c.f(S_DUMAX), dumin = c.f(S_DUMIN);
#pragma unroll
      for (int q = 0; q < 4 * NC; ++q) {
        const int blk = q / NC;
        wc[(WS::W_K + row) * wst] = blk == 0 ? (umax - uk_1) : (blk == 1 ? -(umin - uk_1) : (blk == 2 ? dumax : -dumin));
        ++row;
      }
    }
    if (WS::UV) {                                                    // constraintsMPC.m:45-63
      double hv[NP + 1], pxv[NP];
      pred_row<NP>(e.a, e.Cv, 0.0, e.Dv, e.xhat, NP, hv, pxv);
      const double vmax = c.f(S_VMAX);
#pragma unroll
      for (int i = 1; i <= NP; ++i) {
        wc[(WS::T_V + i - 1) * wst] = hv[i - 1];
        wc[(WS::W_K + row) * wst] = vmax - ((pxv[i - 1] + hv[i] * uk_1) + e.bv);
        ++row;
      }
    }
    if (WS::UE) {                                                    // constraintsMPC.m:66-85
      double he[NP + 1], pxe[NP];
      pred_row<NP>(e.a, e.Cphi, 0.0, e.Dphi, e.xhat, NP, he, pxe);
      const double pmin = c.f(S_PHIMIN);
#pragma unroll
      for (int i = 1; i <= NP; ++i) {
        wc[(WS::T_E + i - 1) * wst] = -he[i - 1];
        wc[(WS::W_K + row) * wst] = -pmin + ((pxe[i - 1] + he[i] * uk_1) + e.bphi);
        ++row;
      }
    }
    if (WS::UZ) {                                                    // constraintsMPC.m:89-101
      const double zmax = c.f(S_ZMAX);
#pragma unroll
      for (int i = 1; i <= NP; ++i) {
        wc[(WS::T_S + i - 1) * wst] = hs[i - 1];
        wc[(WS::W_K + row) * wst] = zmax - ((pxs[i - 1] + hs[i] * uk_1) + e.soc);
        ++row;
      }
    }
  }
  bool viol = false;                                                 // iterMPC.m:66 (Q8)
  wide_rows<NP, NC, CONS, 1>(wc, wst, [&](int i, const double (&M)[NC]) {
    double acc = M[0] * DU[0];
#pragma unroll
    for (int j = 1; j < NC; ++j) acc = fma(M[j], DU[j], acc);
    if (acc - wc[(WS::W_K + i) * wst] > 0.0) viol = true;
  });
  o.nexec = 0;
#pragma unroll
  for (int w = 0; w < 4; ++w) o.active[w] = 0u;
  if (viol) {                                                        // hildreth.m:16-46 (see hildreth_core)
    const bool warm = c.f(S_LAMVALID) != 0.0;                        // Q11
#pragma unroll 4
    for (int i = 0; i < nC; ++i) lc[i * ls] = warm ? c.lam(i) : 0.0;
    spd_solve<NC>(NC, LU, F, x);                                     // E\F
    double Ei[NC][NC];                                               // inv(E)
#pragma unroll
    for (int l = 0; l < NC; ++l) {
      double el[NC], col[NC];
#pragma unroll
      for (int j = 0; j < NC; ++j) el[j] = (j == l) ? 1.0 : 0.0;
      spd_solve<NC>(NC, LU, el, col);
#pragma unroll
      for (int j = 0; j < NC; ++j) Ei[j][l] = col[j];
    }
    double v[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) v[j] = 0.0;
    unsigned z0 = 0u, z1 = 0u, z2 = 0u;                              // rows with H_ii == 0 (all-zero M rows, Q7)
    int npoison = 0;
    double pchk = 0.0;                                               // stays (+-)0 while every entry of E\M' is finite
    bool hpos = true;                                                // H_ii > 0 on every row that is not all-zero
    wide_rows<NP, NC, CONS, 1>(wc, wst, [&](int i, const double (&M)[NC]) {
      double P[NC];
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        double acc = Ei[j][0] * M[0];
#pragma unroll
        for (int l = 1; l < NC; ++l) acc = fma(Ei[j][l], M[l], acc);
        P[j] = acc;
        pchk = fma(acc, 0.0, pchk);
      }
      double hd = M[0] * P[0], k = M[0] * x[0];
#pragma unroll
      for (int j = 1; j < NC; ++j) { hd = fma(M[j], P[j], hd); k = fma(M[j], x[j], k); }
      const double li = lc[i * ls];
      wc[(WS::W_K + i) * wst] = -(k + wc[(WS::W_K + i) * wst]);  // -K(i), K(i) = M(i,:)*(E\F) + gamma(i): the block test compares against -K
      wc[(WS::W_H + i) * wst] = rcp_d(hd);                         // 1/H(i,i)
      if (hd == 0.0) {
        const unsigned bit = 1u << (i & 31);
        if (i < 32) z0 |= bit; else if (i < 64) z1 |= bit; else z2 |= bit;
        if (!is_finite(li)) ++npoison;
      } else {
        hpos = hpos && (hd > 0.0);
#pragma unroll
        for (int j = 0; j < NC; ++j) v[j] = fma(P[j], li, v[j]);
      }
    });
    const bool pfin = pchk == 0.0;
    // the factor of E and F are not needed until the sweeps are over: park them (registers are what the sweep is short of)
#pragma unroll
    for (int q = 0; q < NC * NC; ++q) wc[(WS::W_LU + q) * wst] = LU[q];
#pragma unroll
    for (int q = 0; q < NC; ++q) wc[(WS::W_F + q) * wst] = F[q];
    int k = 1;
    // The 80 rows are walked in blocks of Nc rows; every block is a Toeplitz window [h(q) ... h(q-Nc+1)] over one table
    // (the current rows too: prefix rows are the table [s s s ...], identity rows [s 0 0 ...]).  A block is first tested
    // as a whole from the present v -- Nc independent short chains, their loads issued together one block ahead --
    // and left alone by a cell in which every row of it is idle (a per-cell branch, no warp vote): lambda_i = 0 and K_i + H(i,:)lambda >= 0
    // give w <= 0 (or NaN on the all-zero row: 0*Inf), max(0,w) = 0, lambda and v exactly as they were, so the
    // sequential pass of hildreth.m:34-37 would have seen that same v in every row of the block and changed nothing.
    // Otherwise the block is walked row by row exactly as written there.
    static_assert(NP % NC == 0, "blocked sweep: the horizon must be a multiple of the block length");
    constexpr int BL = NC, BPT = NP / BL, NCUR = WS::UC ? 4 : 0, NBLK = nC / BL;
    constexpr int TS0 = WS::UV ? WS::T_V : (WS::UE ? WS::T_E : WS::T_S);
    constexpr int TS1 = (WS::UV && WS::UE) ? WS::T_E : WS::T_S;
    constexpr int TS2 = WS::T_S;
    const bool skip_base = pfin && hpos;
    double win[2 * NC - 1];                                          // win[k] = table[q0 - (Nc-1) + k]; row r reads M(j) = win[Nc-1 + r - j]
#pragma unroll
    for (int q = 0; q < 2 * NC - 1; ++q) win[q] = 0.0;
    double Kn[BL], Tn[BL];                                           // K and the new table entries of the coming block (in flight)
    auto fetch = [&](int blk, double (&Kd)[BL], double (&Td)[BL]) {
#pragma unroll
      for (int r = 0; r < BL; ++r) Kd[r] = wc[(WS::W_K + blk * BL + r) * wst];
      if (blk >= NCUR) {
        const int tb = (blk - NCUR) / BPT, bb = (blk - NCUR) % BPT;
        const double* t = wc + (tb == 0 ? TS0 : (tb == 1 ? TS1 : TS2)) * wst;
#pragma unroll
        for (int r = 0; r < BL; ++r) Td[r] = t[(bb * BL + r) * wst];
      }
    };
    fetch(0, Kn, Tn);
    for (; k <= d.maxHild; ++k) {                                    // hildreth.m:32
      bool notconv = false;
#pragma unroll 1
      for (int blk = 0; blk < NBLK; ++blk) {
        double Kc[BL], Tc[BL];
#pragma unroll
        for (int r = 0; r < BL; ++r) { Kc[r] = Kn[r]; Tc[r] = Tn[r]; }
        fetch(blk + 1 == NBLK ? 0 : blk + 1, Kn, Tn);
        if (blk < NCUR) {                                            // constraintsMPC.m:23-42: [prefix +; prefix -; identity +; identity -]
          const bool neg = (blk & 1) != 0, ident = blk >= 2;
          const double one = neg ? -1.0 : 1.0, zer = neg ? -0.0 : 0.0;
#pragma unroll
          for (int q = 0; q < 2 * NC - 1; ++q) win[q] = (q < NC - 1) ? zer : ((q == NC - 1 || !ident) ? one : zer);
        } else {
          const int tb = (blk - NCUR) / BPT, bb = (blk - NCUR) % BPT;
          const bool neg = (tb == 0 ? TS0 : (tb == 1 ? TS1 : TS2)) == WS::T_E;
#pragma unroll
          for (int q = 0; q < NC - 1; ++q) win[q] = (bb == 0) ? (neg ? -0.0 : 0.0) : win[BL + q];
#pragma unroll
          for (int r = 0; r < BL; ++r) win[NC - 1 + r] = Tc[r];
        }
        const int i0 = blk * BL;
        double li[BL];
        bool idle = skip_base && (npoison == 0);
#pragma unroll
        for (int r = 0; r < BL; ++r) {
          li[r] = lc[(i0 + r) * ls];
          double hl = win[NC - 1 + r] * v[0];
#pragma unroll
          for (int j = 1; j < NC; ++j) hl = fma(win[NC - 1 + r - j], v[j], hl);
          idle = idle && (li[r] == 0.0) && (hl >= Kc[r]);             // Kc = -K: K + H(i,:)lambda >= 0 without the addition on the chain
        }
        if (idle) continue;                                          // per cell: see the lean sweep of mpc_step_fast
        double rH[BL];
#pragma unroll
        for (int r = 0; r < BL; ++r) rH[r] = wc[(WS::W_H + i0 + r) * wst];
#pragma unroll
        for (int r = 0; r < BL; ++r) {                               // hildreth.m:34-37
          const int i = i0 + r;
          double w;
          if (npoison > 0) {
            w = NAN;
          } else {
            double hl = win[NC - 1 + r] * v[0];
#pragma unroll
            for (int j = 1; j < NC; ++j) hl = fma(win[NC - 1 + r - j], v[j], hl);
            w = fma(-(hl - Kc[r]), rH[r], li[r]);                    // hl - (-K) = K + hl, bit for bit
          }
          const double nl = (w > 0.0) ? w : 0.0;                     // max(0,w); max(0,NaN) = 0
          const double dl = nl - li[r];
          const unsigned zw = i < 32 ? z0 : (i < 64 ? z1 : z2);
          if ((zw >> (i & 31)) & 1u) {
            npoison += (int)(!is_finite(nl)) - (int)(!is_finite(li[r]));
          } else if (dl != 0.0 || !pfin) {                           // P*0 leaves v as it is
#pragma unroll
            for (int j = 0; j < NC; ++j) {
              double acc = Ei[j][0] * win[NC - 1 + r];
#pragma unroll
              for (int l = 1; l < NC; ++l) acc = fma(Ei[j][l], win[NC - 1 + r - l], acc);
              v[j] = fma(acc, dl, v[j]);
            }
          }
          lc[i * ls] = nl;
          if (!(fabs(dl) < 1e-6)) notconv = true;                    // hildreth.m:39 norm(.,Inf) < 1e-6, NaN-propagating
        }
      }
      if (!notconv) break;
    }
    if (k > d.maxHild) k = d.maxHild;
    {
      const volatile double* wv = wc;
#pragma unroll
      for (int q = 0; q < NC * NC; ++q) LU[q] = wv[(WS::W_LU + q) * wst];
#pragma unroll
      for (int q = 0; q < NC; ++q) F[q] = wv[(WS::W_F + q) * wst];
    }
    double rhs[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) rhs[j] = 0.0;
    unsigned a0 = 0u, a1 = 0u, a2 = 0u;
    wide_rows<NP, NC, CONS, 1>(wc, wst, [&](int i, const double (&M)[NC]) {     // hildreth.m:46  M'*lambda, rows in order
      const double li = lc[i * ls];
#pragma unroll
      for (int j = 0; j < NC; ++j) rhs[j] = (i == 0) ? M[j] * li : fma(M[j], li, rhs[j]);
      c.lam(i) = li;                                                 // iterMPC.m:68
      if (li > 0.0) {
        const unsigned bit = 1u << (i & 31);
        if (i < 32) a0 |= bit; else if (i < 64) a1 |= bit; else a2 |= bit;
      }
    });
#pragma unroll
    for (int j = 0; j < NC; ++j) rhs[j] = F[j] + rhs[j];
    spd_solve<NC>(NC, LU, rhs, x);
#pragma unroll
    for (int j = 0; j < NC; ++j) DU[j] = -x[j];
    o.nexec = k;                                                     // iterMPC.m:67-69
    o.active[0] = a0; o.active[1] = a1; o.active[2] = a2;
    c.f(S_LAMVALID) = 1.0;
  }
  const double uk = DU[0] + uk_1;                                    // iterMPC.m:75-76
  c.f(S_UK) = uk;
  o.u = uk;
  if (!is_finite(uk)) c.f(S_STATUS) = (double)((unsigned)c.f(S_STATUS) | ST_NONFINITE);
}


}  // namespace mek
