// Kernel parameter block, the TMA staging of the hot ROM bank and the fused persistent kernel k_run_population
// (one thread per cell, every sample instant of runMPC.m:83-112 without leaving the kernel).
#pragma once
#include <cuda_runtime.h>

#include "mek_core.cuh"

#ifndef MEK_STREAM_BATCH
#define MEK_STREAM_BATCH 2   // models whose loads are in flight per thread in the HBM stream
#endif

using namespace mek;

// ------------------------------------------------------------------------------------------
// kernel parameters
// ------------------------------------------------------------------------------------------
struct Params {
  Dims d;
  double cp[N_PARAMS_X];      // cellData closed-form parameters + the descriptor of the optional tables (mek_cellfn.cuh)
  double* st;                 // SoA state block [nfields][ld]
  const double* bank_blob;    // hot bank in global memory (source of the TMA bulk copy)
  int blob_doubles, off_mdl;
  // full bank (function-level zk output only)
  const double* Cfull;        // [M][nz][6]
  const double* Dfull;        // [M][nz]
  // inputs staged on the device (mek_pop_init / mek_pop_reset)
  const double* in_soc0; const double* in_ekf; const double* in_T; const double* in_sigx0; const double* in_lim;
  // trajectory logs [step][nlog]
  long long nlog; int log_steps;
  double* log_u; double* log_V; double* log_soc; double* log_phise;
  unsigned char* log_iT; unsigned char* log_iZ; unsigned short* log_nexec; unsigned* log_active; int log_words;
  int sync_steps;             // 1: CTA-wide barrier before every sample instant
  int defer_k;                // > 0: sample instants per deferral block (models outside a cell's hot set are advanced once per block); 0: off
  int defer_stagger;          // 1: the warps of a CTA start their blocks at different instants (spreads the block-end HBM bursts)
  double* wide_ws;            // long-horizon instance: per-CTA workspace [WSLOTS][threads] for tables, K, 1/H (null: shared memory)
  int defer_narrow;           // test hook: hot set = the EKF estimate's neighbours only, no margin
  int hist_off;               // doubles from the start of dynamic shared memory to the history columns [defer_k + 1][threads]
  const int* slot2cell;       // state slot -> caller's cell index (cells are ordered by initial SOC inside the block), or null
};

__device__ __forceinline__ long long cell_id(const Params& P, long long slot) { return P.slot2cell ? (long long)P.slot2cell[slot] : slot; }

// ---- TMA bulk copy of the hot bank into shared memory ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void stage_bank(double* sm, const Params& P, BankSmem& bk) {
  __shared__ __align__(8) unsigned long long mbar;
  const unsigned bytes = (unsigned)P.blob_doubles * 8u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bytes) : "memory");
    unsigned done = 0;
    while (done < bytes) {                                   // chunks keep each bulk request modest
      const unsigned chunk = (bytes - done) < 32768u ? (bytes - done) : 32768u;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(reinterpret_cast<char*>(sm) + done)),
                   "l"(reinterpret_cast<const char*>(P.bank_blob) + done), "r"(chunk), "r"(smem_u32(&mbar))
                   : "memory");
      done += chunk;
    }
  }
  unsigned ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(&mbar)), "r"(0u)
        : "memory");
  }
  bk.Tpts = sm;
  bk.Zpts = sm + P.d.nT;
  bk.mdl = sm + P.off_mdl;
}

// ---- one pass of the loop body runMPC.m:83-112 for one cell ----
struct StepLog {
  double u, V, soc, phise;
  int iT, iZ, nexec;
  unsigned active[4];
};

template <int NPM, int NCM, int CONS, int BATCH>   // CONS >= 0: static-shape fast MPC path (constraint switches compile-time), -1: generic
__device__ __forceinline__ void cell_step(const Params& P, const BankSmem& bk, const Cell& c, StepLog& lg, double* smt, const int sms, int tt,
                                          Defer* df, long long id, int k, double qs, double rcs) {
  const Dims& d = P.d;
  lg.u = NAN; lg.V = NAN; lg.soc = NAN; lg.phise = NAN; lg.iT = 0; lg.iZ = 0; lg.nexec = 0;
#pragma unroll
  for (int w = 0; w < 4; ++w) lg.active[w] = 0u;
  if ((unsigned)c.f(S_STATUS) & ST_LOST) return;                     // Q19: a lost cell idles
  const double uk = c.f(S_UK);
  if (!df) prefetch_models(d, c, 0, 6, true, true);                  // first rows of this instant's HBM stream -> L2 while the plant outputs are computed
  const double V = plant_step<false>(d, bk, P.cp, c, uk, tt, df, qs, rcs);   // runMPC.m:85 (outputs; the state advance is fused below)
  // OB_step.m:278 with u_k and iterEKF.m:73-84 with priorI = u_{k-1}: one HBM pass over all M local models, or (deferral
  // on) over the cell's hot models only -- the rest catch up at the end of the block
  if (df) defer_stream(d, bk, c, *df, uk, c.f(S_PRIORI));
  else stream_models<true, true, BATCH>(d, bk, c, uk, c.f(S_PRIORI));
  EkfOut e;
  double vk = V;                                                     // runMPC.m:88 feeds the filter the plant voltage itself
#ifndef MEK_LUT_OFF                                                  // scenarios live in the full-featured instances only
  if (d.meas_sigma != 0.0) vk = V + d.meas_sigma * noise_gauss(d.noise_seed, (unsigned long long)(d.first_cell + id), (unsigned long long)k);
#endif
  ekf_step<false>(d, bk, P.cp, c, vk, uk, e, Scratch{smt, sms}, tt, df);  // runMPC.m:88,92
  if (e.lost) return;
  MpcOut o;
  if constexpr (CONS >= 0 && NPM == 5 && NCM == 2) mpc_step_fast<NPM, NCM, CONS>(d, P.cp, c, e, uk, o, smt, sms);   // runMPC.m:94-103
  else if constexpr (CONS >= 0) {
    using WS = WideShape<NPM, NCM, CONS>;
    // lambda column in shared memory; tables, K, 1/H in this CTA's slice of the L2-resident workspace, or (small CTAs) in shared memory too
    double* lc = P.wide_ws ? smt : smt + WS::WSLOTS * sms;
    double* wc = P.wide_ws ? P.wide_ws + (long long)blockIdx.x * WS::WSLOTS * sms + threadIdx.x : smt;
    mpc_step_wide<NPM, NCM, CONS>(d, P.cp, c, e, uk, o, lc, sms, wc, sms);
  }
  else mpc_step<NPM, NCM>(d, P.cp, c, e, uk, o);
  lg.u = o.u; lg.V = V; lg.soc = e.soc; lg.phise = o.phise;          // runMPC.m:106-109
  lg.iT = e.iT + 1; lg.iZ = e.iZ + 1; lg.nexec = o.nexec;
#pragma unroll
  for (int w = 0; w < 4; ++w) lg.active[w] = o.active[w];
}

__device__ __forceinline__ unsigned long long warp_or64(unsigned long long v) {
  const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)v), hi = __reduce_or_sync(0xffffffffu, (unsigned)(v >> 32));
  return ((unsigned long long)hi << 32) | lo;
}

// ---- k_run_population: the fused persistent kernel ----
// KTAG only makes the instances of different translation units different symbols: a unit compiled with MEK_LUT_OFF (closed-form
// cell functions only, tag 0) and one compiled with the tabulated cell functions (tag 1) both instantiate <5,2,15,256,1,...>
#ifdef MEK_LUT_OFF
#define MEK_KTAG 0
#else
#define MEK_KTAG 1
#endif
template <int NPM, int NCM, int CONS, int THREADS, int MINB, int KTAG = MEK_KTAG>   // THREADS/MINB: launch bounds (register cap)
__global__ void __launch_bounds__(THREADS, MINB) k_run_population(const __grid_constant__ Params P, int k0, int k1) {
  extern __shared__ __align__(16) double sm[];
  BankSmem bk;
  stage_bank(sm, P, bk);
  double* smt = sm + P.blob_doubles + threadIdx.x;   // this thread's column of the per-thread shared arrays
  constexpr int sms = THREADS;                       // launched with exactly THREADS threads per CTA
  const Dims& d = P.d;
  const long long stride = (long long)gridDim.x * THREADS;
  const long long first = (long long)blockIdx.x * THREADS;
  const int KB = P.defer_k;                          // 0: every model is advanced every instant
  Defer df{0ull, 0ull, 0, sm + P.hist_off + threadIdx.x, THREADS};
  // block boundaries of this warp: instants with (k - k0 + off) % KB == 0 (and k0 itself)
  const int off = (KB > 0 && P.defer_stagger) ? (int)(((threadIdx.x >> 5) * KB) / (THREADS / 32)) : 0;
  // every thread of the CTA walks the same number of rounds so that the optional per-instant barrier is legal
  for (long long base = first; base < d.B; base += stride) {
    const long long cell = base + threadIdx.x;
    const bool active = cell < d.B;
    Cell c = cell_ref(P.st, d, active ? cell : 0);
    // the cell temperature is constant over a run: its two nearest set-points are looked up once (same rule as every step)
    const int tt = active ? two_nearest(c.f(S_TC) + 273.15, bk.Tpts, d.nT) : 0;
    const long long id = active ? cell_id(P, cell) : 0;   // the caller's index of this cell (logs are addressed by it)
#ifndef MEK_LUT_OFF                                        // scenarios live in the full-featured instances only
    const double qs = plant_q_factor(d.plant_q_scale, d.plant_q_spread, d.noise_seed, (unsigned long long)(d.first_cell + id)), rcs = d.plant_rc_scale;
#else
    constexpr double qs = 1.0, rcs = 1.0;
#endif
#pragma unroll 1
    for (int k = k0; k < k1; ++k) {
      if (P.sync_steps) __syncthreads();             // keeps the CTA's warps on the same stretch of code (instruction-cache sharing)
      if (KB > 0 && (k == k0 || (k - k0 + off) % KB == 0)) {           // warp-uniform: start of a deferral block
        if (k != k0 && active) defer_flush(d, bk, c, df);
        const bool run = active && !((unsigned)c.f(S_STATUS) & ST_LOST);
        int left = KB - (k - k0 + off) % KB;
        df.hot = warp_or64(run ? defer_hot_set(d, bk, P.cp, c, tt, left, P.defer_narrow) : 0ull);
        df.mine = 0ull;
        df.n = 0;
        if (run) df.h[0] = c.f(S_PRIORI);
      }
      if (!active) continue;
      StepLog lg;
      cell_step<NPM, NCM, CONS, MEK_STREAM_BATCH>(P, bk, c, lg, smt, sms, tt, KB > 0 ? &df : nullptr, id, k, qs, rcs);
      // per-cell summary, kept in the state header (not in registers: they are the scarce resource of this kernel)
      // L2-only accesses: ten header rows per warp that nothing else reads stay out of L1 (which holds the hot header fields)
      {
        double* h = c.base;
        if (lg.nexec > 0) {
          __stcg(h + S_SUM_NEXEC * TILE, __ldcg(h + S_SUM_NEXEC * TILE) + (double)lg.nexec);
          __stcg(h + S_N_QP * TILE, __ldcg(h + S_N_QP * TILE) + 1.0);
          if (lg.nexec >= d.maxHild) __stcg(h + S_N_CAP * TILE, __ldcg(h + S_N_CAP * TILE) + 1.0);
        }
        __stcg(h + S_MIN_PHISE * TILE, fmin(__ldcg(h + S_MIN_PHISE * TILE), lg.phise));
        __stcg(h + S_MAX_V * TILE, fmax(__ldcg(h + S_MAX_V * TILE), lg.V));
        if (__ldcg(h + S_T_REACH * TILE) < 0.0 && lg.soc >= c.f(S_ZMAX) - 1e-3) __stcg(h + S_T_REACH * TILE, (double)k);
        __stcg(h + S_LAST_SOC * TILE, lg.soc); __stcg(h + S_LAST_V * TILE, lg.V); __stcg(h + S_LAST_PHISE * TILE, lg.phise);
      }
      if (id < P.nlog && k < P.log_steps) {
        const long long o = (long long)k * P.nlog + id;
        P.log_u[o] = lg.u; P.log_V[o] = lg.V; P.log_soc[o] = lg.soc; P.log_phise[o] = lg.phise;
        P.log_iT[o] = (unsigned char)lg.iT; P.log_iZ[o] = (unsigned char)lg.iZ; P.log_nexec[o] = (unsigned short)lg.nexec;
#pragma unroll
        for (int w = 0; w < 4; ++w)                                  // static indices: lg stays in registers
          if (w < P.log_words) P.log_active[o * P.log_words + w] = lg.active[w];
      }
    }
    if (KB > 0 && active && k1 > k0) defer_flush(d, bk, c, df);      // the state block leaves the kernel fully advanced
    if (active) c.f(S_STEPS) += (double)(k1 - k0);
  }
}

