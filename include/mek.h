/*
 * mek.h -- C ABI of libmek, the batched B200 implementation of MPC-EKF4FastCharge's per-timestep
 * control loop.  Plain pointers and sizes only; every array argument is a caller-owned HOST
 * buffer unless stated otherwise.  All functions return 0 (MEK_OK) or a negative mek_status;
 * mek_last_error() gives the message.  A trailing batch dimension B replaces the reference's
 * single cell, so B = 1 reproduces the MATLAB shapes.
 *
 * Reference interfaces replaced (paths under /root/reference/src):
 *   mek_bank_create / mek_bank_load  <- load(ROMfile,'ROM')             MPC-EKF4FastCharge/runMPC.m:4-5
 *   mek_pop_init                     <- initKF(...)                     UTILITY/initKF.m:30
 *                                       initMPC(...)                    MPC-EKF4FastCharge/initMPC.m:1
 *                                       OB_step(0,TC,[],ROM,initCfg)    MPC-EKF4FastCharge/runMPC.m:72-74, OB_step.m:39-72
 *   mek_run                          <- the k-loop of runMPC            MPC-EKF4FastCharge/runMPC.m:83-112
 *   mek_pop_ob_step                  <- OB_step(Iapp,Tc,cellState,ROM)  MPC-EKF4FastCharge/OB_step.m:1
 *   mek_pop_iter_ekf                 <- iterEKF(vk,ik,Tk,ekfData)       UTILITY/iterEKF.m:30
 *                                       EKFmatsHandler(ekfData,Xind,zk,Tk)  MPC-EKF4FastCharge/EKFmatsHandler.m:1
 *   mek_pop_iter_mpc                 <- iterMPC(xk,cellState,mpcData)   MPC-EKF4FastCharge/iterMPC.m:1
 *   mek_pred_mat                     <- predMat(A,B,C,D,Np,Nc)          MPC-EKF4FastCharge/predMat.m:1
 *   mek_constraints                  <- constraintsMPC(x_aug,cellState,mpcData)  MPC-EKF4FastCharge/constraintsMPC.m:1
 *   mek_hildreth                     <- hildreth(E,F,M,gamma,lambda0,maxIter)    MPC-EKF4FastCharge/hildreth.m:1
 *   mek_read_logs                    <- u_store, voltage_store, SOC_store, phise_store  runMPC.m:55-60,106-111
 */
#ifndef MEK_H
#define MEK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  MEK_OK = 0,
  MEK_ERR_ARG = -1,        /* bad argument / shape (MATLAB error() at initKF.m:67, OB_step.m:41)             */
  MEK_ERR_BANK = -2,       /* structural problem in the ROM bank (OB_step.m:140-157, iterEKF.m:692-733,      */
                           /*   initKF.m:74-87)                                                              */
  MEK_ERR_CUDA = -3,       /* CUDA runtime failure                                                           */
  MEK_ERR_UNSUPPORTED = -4,/* configuration outside the built kernel instances                               */
  MEK_ERR_STATE = -5       /* call order (e.g. run before init)                                              */
} mek_status;

/* per-cell status word bits (replaces NaN-latching, iterEKF.m:55-59) */
#define MEK_ST_LOST 1u       /* EKF warnCount > 10: outputs NaN from that step on, the cell idles          */
#define MEK_ST_Q18 2u        /* Thetae < 0 branch reached (iterEKF.m:384-389; the reference would throw)   */
#define MEK_ST_NONFINITE 4u  /* the current command became NaN/Inf                                         */

typedef struct mek_bank mek_bank;
typedef struct mek_pop mek_pop;

/* ROM struct surface.  Matrices are row-major with the model index
 * m = iT*nZ + iZ slowest; A is the diagonal of ROMmdls(iT,iZ).A including the integrator (== 1). */
typedef struct {
  int n;                    /* transient states (5)                                   */
  int nz;                   /* ROM outputs                                            */
  int nT, nZ;               /* set-point grid                                         */
  double Tsamp;             /* xraData.Tsamp                                          */
  const double *Tpts_C;     /* [nT] xraData.T, deg C, ascending                       */
  const double *Zpts_pct;   /* [nZ] xraData.SOC, percent, ascending                   */
  const char *const *names; /* [nz] tfData.names                                      */
  const double *xloc;       /* [nz] tfData.xLoc                                       */
  const double *A;          /* [M][n+1]                                               */
  const double *C;          /* [M][nz][n+1] (last column: integrator residue)         */
  const double *D;          /* [M][nz]                                                */
  const double *params;     /* cellData closed-form parameters (mek/bank.py P_*)      */
  int nparams;
} mek_bank_desc;

typedef struct {
  int Np, Nc;               /* runMPC.m:28-29 (clamped like initMPC.m:32-33)          */
  int maxHild;              /* initMPC.m:47 (100)                                     */
  int constraints[3];       /* [current voltage eta], runMPC.m:33                     */
  int use_zmax;             /* SOC upper-bound rows, constraintsMPC.m:89 (1 in runMPC)*/
  double ref;               /* mpcData.ref = targetSOC (95, percent! Q9)              */
  double SigmaV, SigmaW;    /* runMPC.m:18-19                                         */
  int force_jacobi;         /* 1: always run the eigen-decomposition PSD repair       */
  int threads_per_block;    /* fast instances: 0/256 = 256 threads x 255 regs (default), 384 = 384 x 168 */
  int blocks_per_sm;        /* 0 = from the occupancy calculator, 1 = one CTA per SM  */
  int generic_path;         /* 1: never use the static-shape fast instances           */
  int sync_steps;           /* CTA-wide barrier before every sample instant: 0/1 = on (default), -1 = off */
  int keep_cell_order;      /* 1: do not reorder cells by (T, SOC0) inside the state block             */
  int defer_block;          /* sample instants per deferral block of mek_run: local models outside a cell's present
                               set-point neighbourhood are advanced once per block by replaying the block's currents
                               (bit-identical to advancing them every instant).  0 = default (16), -1 = off        */
  int defer_stagger;        /* 1: warps of a CTA start their deferral blocks at different instants                */
  int regroup_every;        /* mek_run re-sorts the state block by (temperature bracket, present SOC, C-rate) every this
                               many sample instants, so that the cells of a warp stay in the same phase of the charge.
                               Scheduling only (results and the caller's cell indices are unchanged).  0 = automatic
                               (128 when temperatures or limits differ between cells, else off), -1 = off           */
  int defer_narrow_test;    /* test hook, 1: shrink the per-block model set to the EKF estimate's own neighbours so
                               that the plant's models take the on-demand replay path (results are unchanged)    */
  int general_sweep_test;   /* test hook, 1: the static-shape instances always run the general (masked) Hildreth sweep
                               instead of the lean one that applies when the only all-zero row is the first z_max row;
                               2: the lean sweep never skips the tail of a row that is idle in every cell of the warp
                               (results are unchanged either way)                                                  */
  int z_target_terminal;    /* mpcData.const.z_target_terminal (constraintsMPC.m:104-112; never set by runMPC): one more row,
                               -G_soc(end,:)*DU <= -z_max + Phi_soc(end,:)*x_aug.  Generic kernel instances only           */
  double z_tol;             /* mpcData.const.z_tol (runMPC.m:44, 0 there).  limits[6] carries z_max + z_tol; the terminal row
                               uses limits[6] - z_tol for z_max                                                            */
  int ekf_method;           /* initKF's `blend` argument (initKF.m:44-53): 0 = 'OutB' output blend (what runMPC.m:16 selects: the
                               optimised fused kernels), 1 = 'MdlB' model blend: one blended model per cell (iterEKF.m MB
                               branches).  mek_run then runs the plain loop kernel k_run_population_mb; mek_pop_iter_ekf serves
                               the function-level call                                                                      */
  /* Scenario beyond the reference's loop (runMPC.m:88 feeds iterEKF the noiseless plant voltage): measurement noise.  The filter
     sees V + meas_sigma * g(noise_seed, first_cell + cell, k) with g the counter-based sequence below; logs keep the plant's V. */
  double meas_sigma;        /* standard deviation [V]; 0 (default) = the reference's loop                                    */
  uint64_t noise_seed;      /* key of the sequence                                                                           */
  int64_t first_cell;       /* global index of this population's cell 0, so that any shard reproduces any cell's sequence    */
  /* Plant-model mismatch scenario (in the reference: OB_step called with another ROM struct than initKF / initMPC).  The PLANT's
     capacity is Q * plant_q_scale * (1 + plant_q_spread * (2u - 1)), u a per-cell uniform keyed by (noise_seed, first_cell + cell),
     and its collector resistance Rc * plant_rc_scale; the filter and the MPC keep the bank's values.  0 means 1 for the scales.  */
  double plant_q_scale, plant_q_spread, plant_rc_scale;
} mek_config;

/* g(seed, cell, k): x = seed ^ cell*0x9E3779B97F4A7C15 ^ (k+1)*0xD1B54A32D192ED03; twelve splitmix64 steps of x, each output's top
 * 53 bits as a uniform on [0,1), summed in order, minus 6 (mean 0, variance 1).  Exported for callers who want the sequence. */
double mek_noise(uint64_t seed, uint64_t cell, uint64_t k);

/* ---- bank ---- */
int mek_bank_create(const mek_bank_desc *desc, mek_bank **out);
int mek_bank_load(const char *path, mek_bank **out);           /* MEKBANK1 file (mek/bank.py; version 2 may carry tables) */
void mek_bank_destroy(mek_bank *bank);
int mek_bank_dims(const mek_bank *bank, int *n, int *nz, int *nT, int *nZ);
double mek_bank_capacity_Ah(const mek_bank *bank);               /* cellData.function.const.Q()  */
/* Tabulated cell functions (optional).  ROM.cellData.function.{neg,pos}.{Uocp,dUocp,k0,Rf} are MATLAB function handles inside
 * the ROM file (call sites: OB_step.m:231-232,329-330,337-340; iterEKF.m:362-363,392-393,404-407,441-442,463-464,495-496;
 * EKFmatsHandler.m:60-61,68-69,84-85,96).  Without tables the library evaluates the closed forms of the parameter vector;
 * with them -- eight tables of n points on a uniform stoichiometry grid [theta_lo, theta_hi], order Uocp_neg, dUocp_neg,
 * Uocp_pos, dUocp_pos, k0_neg, k0_pos, Rf_neg, Rf_pos (k0, Rf at the reference temperature: the Arrhenius factors of the
 * parameter vector still apply), sampled from the handles by matlab/mek_sample_celldata.m -- it interpolates linearly
 * (interp1 'linear', argument clamped to the grid).  Call before mek_pop_create; populations on a tabulated bank run the
 * runtime-shape kernel instances.  n = 0 removes the tables. */
int mek_bank_set_lut(mek_bank *bank, int n, double theta_lo, double theta_hi, const double *tables /* [8][n] */);
int mek_bank_lut_points(const mek_bank *bank);

/* ---- population (one shard on one device) ---- */
int mek_pop_create(const mek_bank *bank, const mek_config *cfg, int64_t B, int device, mek_pop **out);
void mek_pop_destroy(mek_pop *pop);
/* logs are kept for the first n_log_cells cells of the shard, for up to max_steps sample instants */
int mek_pop_set_logging(mek_pop *pop, int64_t n_log_cells, int max_steps);
/* initKF + initMPC + first OB_step call.  SOC0pct[B] (plant), ekfSOC0pct[B] or NULL (= SOC0pct),
 * Tc[B] deg C, SigmaX0[36*B] (6x6 column-major per cell, symmetric), limits[7*B] =
 * u_max,u_min,du_max,du_min,v_max,phise_min,z_max+z_tol per cell (runMPC.m:36-44, initMPC.m:66-67).
 * Tc is degrees Celsius ONLY: the reference's "T > 100 means Kelvin" heuristic (initKF.m:38-41, iterEKF.m:62-66; absent
 * from OB_step.m:75 and EKFmatsHandler.m:53) is not reproduced, any Tc > 100 is rejected with MEK_ERR_ARG.
 * Synchronous with respect to its arguments: every input buffer may be reused or freed as soon as the call returns. */
int mek_pop_init(mek_pop *pop, const double *SOC0pct, const double *ekfSOC0pct, const double *Tc, const double *SigmaX0,
                 const double *limits);
int mek_pop_reset(mek_pop *pop);    /* re-run the initialisation from the inputs already on the device */
/* loop body of runMPC for sample instants k in [k0,k1) (0-based), asynchronously on the pop's stream */
int mek_run(mek_pop *pop, int k0, int k1);
int mek_run_timed(mek_pop *pop, int k0, int k1, float *ms);     /* mek_run + sync, CUDA-event time of the kernel */
int mek_sync(mek_pop *pop);
void *mek_pop_stream(mek_pop *pop);                              /* cudaStream_t                  */
int64_t mek_pop_launches(const mek_pop *pop);                    /* kernels launched so far       */
int mek_pop_ncon(const mek_pop *pop);                            /* constraint rows nC            */
int mek_pop_defer_block(const mek_pop *pop);                     /* deferral block length in use (0 = off) */
int64_t mek_pop_regroups(const mek_pop *pop);                    /* regroupings of the state block so far (scheduling only) */
int mek_pop_state_len(const mek_pop *pop);                       /* doubles per cell in read/write_state */
/* logs for instants [k0,k1) x first n_log_cells cells, [step][cell] row-major.  Any pointer may be NULL.
 * u = command computed at k (applied at k+1, runMPC.m:106), V = voltage_store, soc = zk(end),
 * phise = phise_store (runMPC.m:94-95); iT,iZ = MPC.iT/iZ (1-based); nexec = mpcData.nexec;
 * active[step][cell][4] = bit i set <=> lambda_i > 0 on exit of hildreth. */
int mek_read_logs(mek_pop *pop, int k0, int k1, double *u, double *V, double *soc, double *phise, int32_t *iT, int32_t *iZ,
                  int32_t *nexec, uint32_t *active);
/* page-locked host memory for the caller's input and result buffers: copies into it run at link speed (a pageable buffer
 * is staged by the driver at a fraction of that).  The reference has no counterpart (MATLAB owns its arrays). */
int mek_host_alloc(size_t bytes, void **out);
int mek_host_free(void *ptr);
/* full per-cell state, cell-major: [B][state_len] (layout: DESIGN.md "state block") */
int mek_read_state(mek_pop *pop, double *state);
int mek_write_state(mek_pop *pop, const double *state);
/* population statistics reduced on the device: totals[MEK_NSTAT] */
#define MEK_NSTAT 16
enum {
  MEK_STAT_CELLS = 0, MEK_STAT_CELLSTEPS, MEK_STAT_SUM_NEXEC, MEK_STAT_N_QP, MEK_STAT_N_CAP, MEK_STAT_N_REACHED,
  MEK_STAT_SUM_T_REACH, MEK_STAT_MIN_PHISE, MEK_STAT_MAX_V, MEK_STAT_N_LOST, MEK_STAT_N_NONFINITE, MEK_STAT_SUM_SOC,
  MEK_STAT_N_JACOBI, MEK_STAT_N_BUMP, MEK_STAT_N_WARN, MEK_STAT_N_Q18
};
int mek_summary(mek_pop *pop, double *totals);
/* per-cell summary columns, cell-major [B][MEK_NCELLSUM], caller's cell order (population analytics; the reference plots
 * one cell's traces in plot4Figures.m, a population needs distributions):
 *   status bits, EKF warnCount (iterEKF.m:741-751), sum nexec, #QP runs, #QPs at the sweep cap, min phise_store, max voltage_store,
 *   first instant with SOC >= z_max - 1e-3 (-1: not reached), instants run, last SOC estimate, last voltage, last phise */
#define MEK_NCELLSUM 12
int mek_read_cell_summaries(mek_pop *pop, double *out);

/* ---- function-level batched entry points on a population's state (one call = one line of the loop body) ---- */
int mek_pop_ob_step(mek_pop *pop, const double *Iapp, double *Vcell);
/* zk[(nz+2)*B] cell-major or NULL; boundzk[(nz+2)*B] = 3-sigma bounds (iterEKF.m:185-205) or NULL (needs zk);
 * mats[31*B] cell-major or NULL =
 *   lost, iT, iZ, xhat(6), diag A(5), Cv(5), Dv, Cphi(5), Dphi, bv, bphi, vhat, SOC, reserved(0) */
#define MEK_NMATS 31
int mek_pop_iter_ekf(mek_pop *pop, const double *vk, const double *ik, double *zk, double *boundzk, double *mats);
/* consumes mats (from mek_pop_iter_ekf) and the applied current; returns the new command and diagnostics */
int mek_pop_iter_mpc(mek_pop *pop, const double *mats, const double *uk_applied, double *uk, double *phise, int32_t *nexec,
                     uint32_t *active);
/* the same call with iterMPC's cost log (iterMPC.m:50-51,82-95): cost[MEK_NCOST*B] cell-major =
 *   mpcData.cost.J_uncon, .J_final, .norm_DU, .viol  (t = (k-1)*Ts and nexec are the caller's / returned above) */
#define MEK_NCOST 4
int mek_pop_iter_mpc_cost(mek_pop *pop, const double *mats, const double *uk_applied, double *uk, double *phise, int32_t *nexec,
                          uint32_t *active, double *cost);

/* ---- stateless batched entry points ---- */
/* hildreth.m: E[Nc*Nc*B], F[Nc*B], M[nC*Nc*B] (row-major per problem), gamma[nC*B], lambda0[nC*B] or NULL */
int mek_hildreth(int64_t B, int Nc, int nC, const double *E, const double *F, const double *M, const double *gamma,
                 const double *lambda0, int maxIter, double *DU, double *lambda, int32_t *nexec, int device);
/* predMat.m for one scalar output, diagonal A (a[6*B] incl. the integrator entry) and B = ones:
 * c[6*B], d[B] -> Phi[Np*7*B] (row-major Np x 7 per problem), G[Np*Nc*B] */
int mek_pred_mat(int64_t B, const double *a, const double *c, const double *d, int Np, int Nc, double *Phi, double *G, int device);

/* constraintsMPC.m plus the pre-QP part of iterMPC.m.  x_aug[7*B] = [x_k(6); u_k-1]; mats[31*B] as produced by
 * mek_pop_iter_ekf (diag A, Cv, Dv, Cphi, Dphi, bv, bphi are read); SOCk_1[B]; limits[7*B]; r = -Ts/(3600*Q)
 * (EKFmatsHandler.m:43); ref = mpcData.ref.  Outputs M[nC*Nc*B], gamma[nC*B]; optional E[Nc*Nc*B], F[Nc*B],
 * DU[Nc*B] (iterMPC.m:38-48). */
/* use_zmax: bit 0 = the SOC upper-bound rows, bit 1 = the optional terminal row (with z_tol = 0) */
int mek_constraints(int64_t B, int Np, int Nc, const int *constraints, int use_zmax, const double *x_aug, const double *mats,
                    const double *SOCk_1, const double *limits, double *M, double *gamma, double r, double ref, double *E,
                    double *F, double *DU, int device);

/* measured FP64 FMA throughput of the device (2 flops per DFMA), the FP64-pipe roofline denominator */
int mek_fp64_peak(int device, double *tflops);

/* Self-test of the speculative quotient / root sequences the kernels use beside the standard operators (csrc/mek_math.cuh,
 * DESIGN.md section 8): for n host operand pairs, q_spec[i] / s_spec[i] = what a kernel computes for x[i] / y[i] and
 * sqrt(x[i]) (fast sequence, or the standard operator when the sequence rejects the operands), q_std / s_std = the standard
 * operators alone, fast[i] = bit 0: the quotient took the fast sequence, bit 1: the root did.  The two must agree bit for
 * bit on every operand (tests/test_gpu_functions.py). */
int mek_selftest_spec(int device, int64_t n, const double *x, const double *y, double *q_spec, double *q_std, double *s_spec,
                      double *s_std, int *fast);

const char *mek_last_error(void);
int mek_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MEK_H */
