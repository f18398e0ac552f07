// Phase-split form of the control loop for large populations.
//
// k_run_population keeps a cell in one thread for the whole charge; its register allocation is the maximum over the
// plant, the EKF and the QP (255 registers, two warps per scheduler), and the loop is bound by FP64 dependency latency.
// For a population that fills the GPU many times over, one sample instant of runMPC.m:83-112 is instead run as two
// launches over ALL cells, each with its own register and shared-memory budget:
//
//   k_split_est  OB_step (:85) + advance of the hot local models + iterEKF (:88) + EKFmatsHandler (:92)
//                -> hands the linearisation (EKFmatsHandler.m:101-114), V_k and zk(end) to the next launch through
//                   a per-cell record in HBM/L2 ("handoff", 28 doubles)
//   k_split_mpc  phise_store (:94-95) + iterMPC (:103) with constraintsMPC / hildreth, the stores (:106-111) and the
//                per-cell summary
//
// and every `defer_k` instants k_split_block replays the block's currents through the models outside the hot set (the
// deferral blocks of mek_core.cuh, with the history column in HBM instead of shared memory) and chooses the next
// block's hot set.  Same device functions, same operations in the same order as the fused kernel: bit-identical results.
#pragma once
#include "mek_kernels.cuh"

// per-cell transient record (separate buffer xt[tile][field][lane], same tiling as the state block): deferral bookkeeping
// of the current block and the EKF -> MPC handoff of the current instant.  Nothing in it outlives a mek_run call.
enum : int {
  X_MINE = 0,   // bit mask (raw 64 bits) of the models this cell caught up on demand inside the block
  X_N,          // instants this cell has recorded in the block (a lost cell stops counting)
  X_TT,         // the cell's two nearest temperature set-points (constant over a run), packed like two_nearest
  X_PACK,       // handoff: lost | iT << 8 | iZ << 16, exact in a double
  X_V, X_SOC,
  X_XHAT,                   // 5
  X_A = X_XHAT + N5,        // 5
  X_CV = X_A + N5,          // 5
  X_DV = X_CV + N5,
  X_CPHI,                   // 5
  X_DPHI = X_CPHI + N5, X_BV, X_BPHI,
  X_HIST                    // defer_k + 1 entries: h[0] = priorI at the start of the block, h[1 + j] = u applied at block instant j
};
__host__ __device__ constexpr int split_xfields(int kb) { return X_HIST + kb + 1; }

struct SplitArgs {
  double* xt;                       // [ntiles][xfields][32]
  unsigned long long* tile_hot;     // [ntiles] warp-uniform hot set of the current block
  int xfields;
};

__device__ __forceinline__ double* xt_ref(const SplitArgs& S, long long cell) {
  return S.xt + (cell >> 5) * ((long long)S.xfields * TILE) + (cell & 31);
}

// start of a deferral block (and of a launch segment): hot set of the coming `span` instants, history origin
__device__ __forceinline__ void split_begin_block(const Params& P, const SplitArgs& S, const BankSmem& bk, long long cell, bool active,
                                                  const Cell& c, double* X, int span) {
  const Dims& d = P.d;
  int tt = 0;
  bool run = false;
  if (active) {
    tt = two_nearest(c.f(S_TC) + 273.15, bk.Tpts, d.nT);
    run = !((unsigned)c.f(S_STATUS) & ST_LOST);
  }
  const unsigned long long hot = warp_or64(run ? defer_hot_set(d, bk, P.cp, c, tt, span, P.defer_narrow) : 0ull);
  if ((threadIdx.x & 31) == 0) S.tile_hot[cell >> 5] = hot;
  if (active) {
    X[X_MINE * TILE] = 0.0;                       // all-zero bits
    X[X_N * TILE] = 0.0;
    X[X_TT * TILE] = (double)tt;
    X[X_HIST * TILE] = c.f(S_PRIORI);
  }
}

// ---- launch 1 of an instant: plant, hot-model advance, EKF, linearisation ----
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_split_est(const __grid_constant__ Params P, const __grid_constant__ SplitArgs S) {
  extern __shared__ __align__(16) double sm[];
  BankSmem bk;
  stage_bank(sm, P, bk);
  const Dims& d = P.d;
  double* smt = sm + P.blob_doubles + threadIdx.x;
  const long long stride = (long long)gridDim.x * THREADS;
  for (long long cell = (long long)blockIdx.x * THREADS + threadIdx.x; cell < d.ld; cell += stride) {
    if (cell >= d.B) continue;
    Cell c = cell_ref(P.st, d, cell);
    double* X = xt_ref(S, cell);
    if ((unsigned)c.f(S_STATUS) & ST_LOST) { X[X_PACK * TILE] = 1.0; continue; }   // Q19: a lost cell idles
    const int tt = (int)X[X_TT * TILE];
    Defer df{S.tile_hot[cell >> 5], double_to_bits(X[X_MINE * TILE]), (int)X[X_N * TILE], X + X_HIST * TILE, TILE};
    const double uk = c.f(S_UK);
    const double V = plant_step<false>(d, bk, P.cp, c, uk, tt, &df);                  // runMPC.m:85
    defer_stream(d, bk, c, df, uk, c.f(S_PRIORI));                                    // OB_step.m:278, iterEKF.m:73-84 (hot models)
    EkfOut e;
    ekf_step<false>(d, bk, P.cp, c, V, uk, e, Scratch{smt, THREADS}, tt, &df);        // runMPC.m:88,92
    X[X_MINE * TILE] = bits_to_double(df.mine);
    X[X_N * TILE] = (double)df.n;
    if (e.lost) { X[X_PACK * TILE] = 1.0; continue; }
    X[X_PACK * TILE] = (double)((e.iT << 8) | (e.iZ << 16));
    X[X_V * TILE] = V;
    X[X_SOC * TILE] = e.soc;
#pragma unroll
    for (int j = 0; j < N5; ++j) {
      X[(X_XHAT + j) * TILE] = e.xhat[j];
      X[(X_A + j) * TILE] = e.a[j];
      X[(X_CV + j) * TILE] = e.Cv[j];
      X[(X_CPHI + j) * TILE] = e.Cphi[j];
    }
    X[X_DV * TILE] = e.Dv; X[X_DPHI * TILE] = e.Dphi; X[X_BV * TILE] = e.bv; X[X_BPHI * TILE] = e.bphi;
  }
}

// ---- launch 2 of an instant: MPC, stores, summary ----
template <int NPM, int NCM, int CONS, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_split_mpc(const __grid_constant__ Params P, const __grid_constant__ SplitArgs S, int k) {
  extern __shared__ __align__(16) double sm[];
  const Dims& d = P.d;
  double* smt = sm + threadIdx.x;
  const long long stride = (long long)gridDim.x * THREADS;
  for (long long cell = (long long)blockIdx.x * THREADS + threadIdx.x; cell < d.B; cell += stride) {
    Cell c = cell_ref(P.st, d, cell);
    const double* X = xt_ref(S, cell);
    StepLog lg;
    lg.u = NAN; lg.V = NAN; lg.soc = NAN; lg.phise = NAN; lg.iT = 0; lg.iZ = 0; lg.nexec = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) lg.active[w] = 0u;
    const int pack = (int)X[X_PACK * TILE];
    if (!(pack & 1)) {
      EkfOut e;
      e.lost = 0; e.vhat = 0.0;
      e.iT = (pack >> 8) & 0xff; e.iZ = (pack >> 16) & 0xff;
      e.soc = X[X_SOC * TILE];
#pragma unroll
      for (int j = 0; j < N5; ++j) {
        e.xhat[j] = X[(X_XHAT + j) * TILE];
        e.a[j] = X[(X_A + j) * TILE];
        e.Cv[j] = X[(X_CV + j) * TILE];
        e.Cphi[j] = X[(X_CPHI + j) * TILE];
      }
      e.xhat[5] = 0.0;                                               // EKFmatsHandler.m:33 (Q3)
      e.Dv = X[X_DV * TILE]; e.Dphi = X[X_DPHI * TILE]; e.bv = X[X_BV * TILE]; e.bphi = X[X_BPHI * TILE];
      const double uk = c.f(S_UK);
      MpcOut o;
      mpc_step_fast<NPM, NCM, CONS>(d, P.cp, c, e, uk, o, smt, THREADS);            // runMPC.m:94-103
      lg.u = o.u; lg.V = X[X_V * TILE]; lg.soc = e.soc; lg.phise = o.phise;           // runMPC.m:106-109
      lg.iT = e.iT + 1; lg.iZ = e.iZ + 1; lg.nexec = o.nexec;
#pragma unroll
      for (int w = 0; w < 4; ++w) lg.active[w] = o.active[w];
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
    c.f(S_STEPS) += 1.0;
    const long long id = cell_id(P, cell);
    if (id < P.nlog && k < P.log_steps) {
      const long long o = (long long)k * P.nlog + id;
      P.log_u[o] = lg.u; P.log_V[o] = lg.V; P.log_soc[o] = lg.soc; P.log_phise[o] = lg.phise;
      P.log_iT[o] = (unsigned char)lg.iT; P.log_iZ[o] = (unsigned char)lg.iZ; P.log_nexec[o] = (unsigned short)lg.nexec;
      for (int w = 0; w < P.log_words; ++w) P.log_active[o * P.log_words + w] = lg.active[w];
    }
  }
}

// ---- block boundary: flush the finished block (flush != 0), open the next one (span > 0) ----
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_split_block(const __grid_constant__ Params P, const __grid_constant__ SplitArgs S, int flush, int span) {
  extern __shared__ __align__(16) double sm[];
  BankSmem bk;
  stage_bank(sm, P, bk);
  const Dims& d = P.d;
  double* hcol = sm + P.blob_doubles + threadIdx.x;                    // history column staged in shared memory for the replay
  const long long stride = (long long)gridDim.x * THREADS;
  for (long long cell = (long long)blockIdx.x * THREADS + threadIdx.x; cell < d.ld; cell += stride) {
    const bool active = cell < d.B;
    Cell c = cell_ref(P.st, d, active ? cell : 0);
    double* X = xt_ref(S, cell);
    if (flush && active) {
      const int n = (int)X[X_N * TILE];
      for (int j = 0; j <= n; ++j) hcol[j * THREADS] = X[(X_HIST + j) * TILE];
      Defer df{S.tile_hot[cell >> 5], double_to_bits(X[X_MINE * TILE]), n, hcol, THREADS};
      defer_flush(d, bk, c, df);
    }
    if (span > 0) split_begin_block(P, S, bk, cell, active, c, X, span);
  }
}
