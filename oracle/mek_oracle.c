/*
 * TEST INFRASTRUCTURE (oracle) -- not part of the product path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Plain-C restatement of the reference's per-timestep control loop with the structure of the
 * problem exploited (diagonal A, B = 1, symmetric covariance, only the output rows the control
 * path consumes, rank-Nc Hildreth dual).  PARITY UNPINNED: the reference has no tests/golden
 * vectors and cannot be executed here (no MATLAB/Octave, ROM .mat missing); this file is pinned
 * against oracle/literal.py (the line-by-line NumPy restatement) to 1e-9 relative with identical
 * discrete outcomes, see tests/test_oracle_cross.py.
 *
 * Reference lines followed (under /root/reference/src):
 *   plant_step      MPC-EKF4FastCharge/OB_step.m:188-357 (simStep), :63-71 (init)
 *   ekf_step        UTILITY/iterEKF.m:55-210 OB branches, helpers :219-255 (getXind),
 *                   :259-417 (getVariables), :421-519 (getChatV); UTILITY/initKF.m:37-133
 *   ekf_step_mb     the same file's MB (model-blend) branches, incl. getChatZ :523-602 for the 3-sigma bounds
 *   pick_linearise  MPC-EKF4FastCharge/EKFmatsHandler.m:26-114
 *   pred_row        MPC-EKF4FastCharge/predMat.m:15-50
 *   mpc_step        MPC-EKF4FastCharge/iterMPC.m:17-95, constraintsMPC.m:18-101, hildreth.m:16-46,
 *                   initMPC.m:32-72
 *   cell loop       MPC-EKF4FastCharge/runMPC.m:72-111
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).  No FMA
 * contraction, no fast-math: every +,-,*,/ and sqrt is a correctly rounded IEEE-754 binary64 op
 * in source order, so a second implementation with the same order reproduces it bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* The op-counting build (oracle/opcount.cpp) marks the blocks of BASELINE.md section 3 with OPC_BLOCK(id):
   0 plant, 1 EKF time update, 2 EKF measurement path, 3 PSD check/repair, 4 EKFmatsHandler, 5 iterMPC fixed part,
   6 constraintsMPC, 7 QP set-up + recovery, 8 Hildreth sweeps, 9 everything else.  A no-op in the normal build. */
#ifndef OPC_BLOCK
#define OPC_BLOCK(id) ((void)0)
#endif

#define N5 5
#define NSYM 15
#define NP_MAX 32
#define NC_MAX 8
#define NCON_MAX (4 * NC_MAX + 3 * NP_MAX + 1)   /* + the optional terminal row (constraintsMPC.m:104-112) */
#define MAXROWS 64   /* nz <= 64: no group of output rows can overflow */

/* parameter vector layout (restated; equals mek/bank.py) */
enum { P_F = 0, P_R, P_Q, P_RC, P_TREF, P_NEG = 5, P_POS = 16, P_UN = 27, P_UP = 33, P_UP_COEF = 36, N_PARAMS = 44 };
enum { E_THETA0 = 0, E_THETA100, E_K0REF, E_K0SLOPE, E_EA_K0, E_RFREF, E_RFSLOPE, E_EA_RF, E_WDL, E_CDL, E_NDL };
/* optional tables of cellData.function.{Uocp,dUocp,k0,Rf} (include/mek.h: mek_bank_set_lut): the caller appends n (0 = closed
   forms), lo, 1/step and the table base address (bit pattern of a pointer) to the parameter vector */
enum { P_LUT_N = N_PARAMS, P_LUT_LO, P_LUT_INVH, P_LUT_PTR, N_PARAMS_X };
enum { LUT_UN = 0, LUT_DUN, LUT_UP, LUT_DUP, LUT_K0N, LUT_K0P, LUT_RFN, LUT_RFP, LUT_NTAB };

/* status bits */
#define ST_LOST 1u        /* warnCount > 10 (iterEKF.m:55-59)              */
#define ST_Q18 2u         /* Thetae < 0 branch reached (iterEKF.m:384-389) */
#define ST_NONFINITE 4u   /* command became NaN/Inf                        */

typedef struct {
  int n, nz, nT, nZ;
  double Ts;
  const double *Tpts_K; /* [nT] ascending */
  const double *Zpts;   /* [nZ] ascending, fraction */
  const double *A;      /* [M][6] diag                 */
  const double *C;      /* [M][nz][6]                  */
  const double *D;      /* [M][nz]                     */
  const double *params; /* [N_PARAMS]                  */
  /* resolved 0-based rows (iterEKF.m:616-690 / OB_step.m:93-137) */
  int r_ifdl0, r_ifdl3, r_if0, r_if3, r_tss0, r_tss3, r_phie_last, r_phise0, r_phi;
  int n_negtss, n_postss, n_thetae, n_negphise, n_posphise, n_phie, n_posphis;
  int negtss[MAXROWS], postss[MAXROWS], thetae[MAXROWS], negphise[MAXROWS], posphise[MAXROWS], phie[MAXROWS], posphis[MAXROWS];
} mko_bank;

typedef struct {
  int Np, Nc, maxHild;
  int useCurrent, useVoltage, useEta, useZmax;
  double ref, SigmaV, SigmaW;
  int force_jacobi; /* 1: always run the eigen-decomposition PSD repair (validation of the PD fast path) */
  int z_target_terminal; /* mpcData.const.z_target_terminal (constraintsMPC.m:104): one more row, SOC(Np) >= z_max */
  double z_tol;          /* mpcData.const.z_tol (runMPC.m:44); the z_max limit of the state carries z_max + z_tol */
  int ekf_method;        /* initKF's blend (initKF.m:44-53): 0 'OutB' (runMPC.m:16), 1 'MdlB': mko_cell_step then runs ekf_step_mb */
  double meas_sigma;     /* scenario beyond runMPC.m:88 (which feeds the filter the noiseless plant voltage): standard deviation [V] of
                            measurement noise added to vk; 0 = the reference's loop */
  uint64_t noise_seed;   /* key of the counter-based noise sequence (seed, cell, instant) */
  int64_t first_cell;    /* global index of cell 0 of this run: any shard reproduces any cell's sequence */
  /* plant-model mismatch scenario (the reference: OB_step given another ROM struct than the filter): the PLANT's capacity is
     Q*q_scale*(1 + q_spread*(2u-1)), u per cell, and its collector resistance Rc*rc_scale; 0 means 1 for the scales */
  double plant_q_scale, plant_q_spread, plant_rc_scale;
} mko_config;

/* per-cell state, flat doubles (also the host-side AoS layout of mek_read_state) */
enum {
  S_SOCN = 0, S_SOCP, S_SOC0N, S_SOC0P, S_CDLN, S_CDLP, S_TC, S_ARRK0N, S_ARRK0P, S_ARRRFN, S_ARRRFP,
  S_X0, S_SIGX0, S_PRIORI, S_EKFSOC0, S_WARN, S_UK, S_LAMVALID, S_STATUS, S_PLANTWARN,
  S_UMAX, S_UMIN, S_DUMAX, S_DUMIN, S_VMAX, S_PHIMIN, S_ZMAX, S_NJAC, S_NBUMP, S_SPARE,
  S_LAMBDA = 30
};
static int st_off_bigx(void) { return S_LAMBDA + NCON_MAX; }
static int st_off_xhat(int M) { return st_off_bigx() + 6 * M; }
static int st_off_sig(int M) { return st_off_xhat(M) + N5 * M; }
/* model-blend EKF (initKF.m:100-101): ekfData.xhat (n+1) and ekfData.SigmaX ((n+1)x(n+1), upper triangle row-major) */
#define N6 6
#define NSYM6 21
static int st_off_mbx(int M) { return st_off_sig(M) + NSYM * M; }
static int st_off_mbs(int M) { return st_off_mbx(M) + N6; }
int mko_state_len(int M) { return st_off_mbs(M) + NSYM6; }
int mko_ncon_max(void) { return NCON_MAX; }

/* ------------------------------------------------------------------------------------------ */
/* deterministic elementary functions built from +,-,*,/,sqrt (same recipe as the product's)  */
/* ------------------------------------------------------------------------------------------ */
static const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
static const double INV_LN2 = 1.44269504088896338700e+00, LN2 = 6.93147180559945286227e-01;

static double atanh_series(double s) { /* 2*atanh(s) for |s| <= 0.1716 */
  double z = s * s;
  double p = 1.0 / 23.0;
  p = fma(p, z, 1.0 / 21.0);
  p = fma(p, z, 1.0 / 19.0);
  p = fma(p, z, 1.0 / 17.0);
  p = fma(p, z, 1.0 / 15.0);
  p = fma(p, z, 1.0 / 13.0);
  p = fma(p, z, 1.0 / 11.0);
  p = fma(p, z, 1.0 / 9.0);
  p = fma(p, z, 1.0 / 7.0);
  p = fma(p, z, 1.0 / 5.0);
  p = fma(p, z, 1.0 / 3.0);
  p = fma(p, z, 1.0);
  return (2.0 * s) * p;
}

static double mko_log(double x) {
  if (!(x > 0.0)) return (x == 0.0) ? -INFINITY : NAN;
  if (x > 1.79769313486231570815e308) return x;
  uint64_t b;
  memcpy(&b, &x, 8);
  int e = (int)(b >> 52) - 1023;
  b = (b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
  double m;
  memcpy(&m, &b, 8);
  if (m > 1.41421356237309514547) { m = m * 0.5; e += 1; }
  double s = (m - 1.0) / (m + 1.0);
  double r = atanh_series(s);
  double de = (double)e;
  return de * LN2_HI + (r + de * LN2_LO);
}

static double mko_exp(double x) {
  if (x != x) return x;
  if (x > 709.0) return INFINITY;
  if (x < -708.0) return 0.0;
  double kf = floor(x * INV_LN2 + 0.5);
  double r = (x - kf * LN2_HI) - kf * LN2_LO;
  double p = 1.0 / 6227020800.0;
  p = fma(p, r, 1.0 / 479001600.0);
  p = fma(p, r, 1.0 / 39916800.0);
  p = fma(p, r, 1.0 / 3628800.0);
  p = fma(p, r, 1.0 / 362880.0);
  p = fma(p, r, 1.0 / 40320.0);
  p = fma(p, r, 1.0 / 5040.0);
  p = fma(p, r, 1.0 / 720.0);
  p = fma(p, r, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  uint64_t b = (uint64_t)((int64_t)kf + 1023) << 52;
  double sc;
  memcpy(&sc, &b, 8);
  return p * sc;
}

static double mko_asinh(double x) {
  double ax = fabs(x);
  double r;
  if (ax != ax) return x;
  if (ax > 1.0e150) {
    r = mko_log(ax) + LN2;
  } else {
    double t = ax + (ax * ax) / (1.0 + sqrt(ax * ax + 1.0));
    if (t < 0.4) {
      r = atanh_series(t / (2.0 + t));
    } else {
      r = mko_log(1.0 + t);
    }
  }
  return (x < 0.0) ? -r : r;
}

double mko_test_log(double x) { return mko_log(x); }
double mko_test_exp(double x) { return mko_exp(x); }
double mko_test_asinh(double x) { return mko_asinh(x); }

/* ------------------------------------------------------------------------------------------ */
/* cellData.function.* closed forms                                                           */
/* ------------------------------------------------------------------------------------------ */
static double soc_e(const double *p, int base, double z) {
  return p[base + E_THETA0] + z * (p[base + E_THETA100] - p[base + E_THETA0]);
}
/* interp1(grid, table, th, 'linear') on a uniform grid, th clamped to the grid */
static double lut_eval(const double *p, int tab, double th) {
  if (th != th) return th;
  const int n = (int)p[P_LUT_N];
  uint64_t bits;
  memcpy(&bits, &p[P_LUT_PTR], 8);
  const double *y = (const double *)(uintptr_t)bits + (size_t)tab * (size_t)n;
  double t = (th - p[P_LUT_LO]) * p[P_LUT_INVH];
  if (!(t > 0.0)) t = 0.0;
  if (t > (double)(n - 1)) t = (double)(n - 1);
  int i = (int)t;
  if (i > n - 2) i = n - 2;
  const double f = t - (double)i;
  return fma(f, y[i + 1] - y[i], y[i]);
}
#define IF_LUT(p, expr) if ((p)[P_LUT_N] != 0.0) return (expr)
static double uocp_n(const double *p, double th) {
  IF_LUT(p, lut_eval(p, LUT_UN, th));
  const double *c = p + P_UN;
  double x = (th - c[4]) / c[5];
  double t = th + c[2];
  return (c[0] + c[1] / (t * t)) + c[3] * (1.0 - x / sqrt(1.0 + x * x));
}
static double duocp_n(const double *p, double th) {
  IF_LUT(p, lut_eval(p, LUT_DUN, th));
  const double *c = p + P_UN;
  double x = (th - c[4]) / c[5];
  double q = 1.0 + x * x;
  double t = th + c[2];
  return (-2.0 * c[1]) / ((t * t) * t) - (c[3] / c[5]) / (q * sqrt(q));
}
static double uocp_p(const double *p, double th) {
  IF_LUT(p, lut_eval(p, LUT_UP, th));
  double x = (th - p[P_UP]) / p[P_UP + 1];
  int nc = (int)p[P_UP + 2];
  double acc = p[P_UP_COEF + nc - 1];
  for (int i = nc - 2; i >= 0; --i) acc = fma(acc, x, p[P_UP_COEF + i]);
  return acc;
}
static double duocp_p(const double *p, double th) {
  IF_LUT(p, lut_eval(p, LUT_DUP, th));
  double x = (th - p[P_UP]) / p[P_UP + 1];
  int nc = (int)p[P_UP + 2];
  double acc = (double)(nc - 1) * p[P_UP_COEF + nc - 1];
  for (int i = nc - 2; i >= 1; --i) acc = fma(acc, x, (double)i * p[P_UP_COEF + i]);
  return acc / p[P_UP + 1];
}
static double k0_e(const double *p, int base, double th, double arr) {
  IF_LUT(p, lut_eval(p, base == P_NEG ? LUT_K0N : LUT_K0P, th) * arr);
  return (p[base + E_K0REF] * (1.0 + p[base + E_K0SLOPE] * (th - 0.5))) * arr;
}
static double rf_e(const double *p, int base, double th, double arr) {
  IF_LUT(p, lut_eval(p, base == P_NEG ? LUT_RFN : LUT_RFP, th) * arr);
  return (p[base + E_RFREF] * (1.0 + p[base + E_RFSLOPE] * th)) * arr;
}
static double arrhenius(const double *p, double T, double Ea) {
  return mko_exp((Ea / p[P_R]) * (1.0 / p[P_TREF] - 1.0 / T));
}

/* ------------------------------------------------------------------------------------------ */
/* index helpers                                                                              */
/* ------------------------------------------------------------------------------------------ */
static void two_nearest(double v, const double *pts, int n, int *i1, int *i2) {
  /* first two entries of a stable ascending sort of |v - pts| (OB_step.m:251, iterEKF.m:221), one pass */
  int a = -1, b = -1;
  double b1 = INFINITY, b2 = INFINITY;
  for (int i = 0; i < n; ++i) {
    const double d = fabs(v - pts[i]);
    if (d < b1) { b2 = b1; b = a; b1 = d; a = i; }
    else if (d < b2) { b2 = d; b = i; }
  }
  if (a < 0) { a = 0; b = (n > 1) ? 1 : 0; }   /* all distances NaN: sort keeps the original order */
  if (b < 0) b = a;
  *i1 = a;
  *i2 = b;
}

static int sym_idx(int i, int j) { /* upper triangle, row-major, i <= j */
  static const int base[5] = {0, 5, 9, 12, 14};
  return (i <= j) ? base[i] + (j - i) : base[j] + (i - j);
}

/* ------------------------------------------------------------------------------------------ */
/* plant: OB_step.m simStep                                                                   */
/* ------------------------------------------------------------------------------------------ */
/* Zero-weight rule, on by default -- it is what the CUDA kernels do: a neighbour model whose bilinear
   weight is EXACTLY zero (a temperature on a set-point of the bank: two of the four neighbours, at every instant) is left
   out of every blend, and its measurement update -- zero gain, x and SigmaX unchanged, the PSD repair of an unchanged
   covariance -- is skipped; the Q-bump is kept.  mko_set_zero_weight_skip(0) restores the term-by-term restatement of the
   reference (same trajectories; tests/test_oracle_units.py asserts it, and the op-count table is taken in that mode). */
static int zw_skip = 1;
void mko_set_zero_weight_skip(int on) { zw_skip = on != 0; }

static double plant_step(const mko_bank *bk, double *st, double Iapp, double qs, double rcs) {
  OPC_BLOCK(0);
  const double *p = bk->params;
  const int M = bk->nT * bk->nZ, nz = bk->nz;
  double *bigX = st + st_off_bigx();
  const double T = st[S_TC] + 273.15;                              /* OB_step.m:75 */
  const double th0n = p[P_NEG + E_THETA0], th100n = p[P_NEG + E_THETA100];
  const double th0p = p[P_POS + E_THETA0], th100p = p[P_POS + E_THETA100];
  const double Q = p[P_Q] * qs;                                    /* qs, rcs: plant-model mismatch factors (1 = the reference) */
  double SOCn = st[S_SOCN], SOCp = st[S_SOCP];
  const double negSOC = SOCn, posSOC = SOCp;                       /* :226-227 (Q16: pre-update) */
  const double cellSOC = (SOCn - th0n) / (th100n - th0n);          /* :228 */
  const double dUn = duocp_n(p, SOCn), dUp = duocp_p(p, SOCp);     /* :231-232 */
  const double dQn = fabs(th100n - th0n), dQp = fabs(th100p - th0p);
  const double res0n = -dQn / (3600.0 * Q - (st[S_CDLN] * dQn) * dUn); /* :235 */
  const double res0p = dQp / (3600.0 * Q - (st[S_CDLP] * dQp) * dUp);  /* :236 */
  SOCn = SOCn + (res0n * Iapp) * bk->Ts;                           /* :239 */
  SOCp = SOCp + (res0p * Iapp) * bk->Ts;
  if (SOCn < 0.0) { SOCn = 0.0; st[S_PLANTWARN] += 1.0; }          /* :241-244 */
  if (SOCn > 1.0) { SOCn = 1.0; st[S_PLANTWARN] += 1.0; }
  if (SOCp < 0.0) { SOCp = 0.0; st[S_PLANTWARN] += 1.0; }
  if (SOCp > 1.0) { SOCp = 1.0; st[S_PLANTWARN] += 1.0; }
  int z1, z2, t1, t2;
  two_nearest(cellSOC, bk->Zpts, bk->nZ, &z1, &z2);                /* :251-259 ordering by index */
  int iZup = z1 > z2 ? z1 : z2, iZlo = z1 < z2 ? z1 : z2;
  two_nearest(T, bk->Tpts_K, bk->nT, &t1, &t2);                    /* :261-269 */
  int iTup = t1 > t2 ? t1 : t2, iTlo = t1 < t2 ? t1 : t2;
  const double Zup = bk->Zpts[iZup], Zlo = bk->Zpts[iZlo], Tup = bk->Tpts_K[iTup], Tlo = bk->Tpts_K[iTlo];
  const int rows[9] = {bk->r_ifdl0, bk->r_ifdl3, bk->r_if0, bk->r_if3, bk->r_tss0, bk->r_tss3,
                       bk->thetae[0], bk->thetae[bk->n_thetae - 1], bk->r_phie_last};
  const int mc[4] = {iTlo * bk->nZ + iZlo, iTlo * bk->nZ + iZup, iTup * bk->nZ + iZlo, iTup * bk->nZ + iZup};
  double y[4][9];
  for (int c = 0; c < 4; ++c) {                                    /* :272-275 */
    const double *Cm = bk->C + (size_t)mc[c] * nz * 6, *Dm = bk->D + (size_t)mc[c] * nz;
    const double *x = bigX + 6 * mc[c];
    for (int r = 0; r < 9; ++r) {
      const double *cr = Cm + rows[r] * 6;
      double acc = cr[0] * x[0];
      for (int j = 1; j < 6; ++j) acc = fma(cr[j], x[j], acc);   /* no Phise row is consumed, so :178-181 is moot here */
      y[c][r] = fma(Dm[rows[r]], Iapp, acc);
    }
  }
  for (int m = 0; m < M; ++m)                                      /* :278 */
    for (int j = 0; j < 6; ++j) bigX[6 * m + j] = fma(bk->A[6 * m + j], bigX[6 * m + j], Iapp);
  double aZ = 0.0, aT = 0.0;                                       /* :281-283 */
  if (Zup != Zlo) aZ = (cellSOC - Zlo) / (Zup - Zlo);
  if (Tup != Tlo) aT = (T - Tlo) / (Tup - Tlo);
  double yk[9];
  for (int r = 0; r < 9; ++r) {                                    /* :285; a temperature row with weight exactly zero is left out */
    const double lo = (1.0 - aZ) * y[0][r] + aZ * y[1][r], hi = (1.0 - aZ) * y[2][r] + aZ * y[3][r];
    yk[r] = (zw_skip && aT == 0.0) ? lo : ((zw_skip && aT == 1.0) ? hi : (1.0 - aT) * lo + aT * hi);
  }
  const double negIfdl0 = yk[0], posIfdl3 = yk[1], negIf0 = yk[2], posIf3 = yk[3];
  double tss0 = yk[4] + st[S_SOC0N], tss3 = yk[5] + st[S_SOC0P];   /* :304-305 (Q6) */
  tss0 = fmin(fmax(tss0, 1e-6), 1.0 - 1e-6);                       /* :309-310 */
  tss3 = fmin(fmax(tss3, 1e-6), 1.0 - 1e-6);
  const double the0 = fmax(yk[6] + 1.0, 1e-6), the3 = fmax(yk[7] + 1.0, 1e-6); /* :325-326 */
  const double k0n = k0_e(p, P_NEG, negSOC, st[S_ARRK0N]), k0p = k0_e(p, P_POS, posSOC, st[S_ARRK0P]);
  const double i0n = k0n * sqrt((the0 * (1.0 - tss0)) * tss0);     /* :331-332 */
  const double i0p = k0p * sqrt((the3 * (1.0 - tss3)) * tss3);
  const double pref = ((2.0 * p[P_R]) * T) / p[P_F];
  const double negEta0 = pref * mko_asinh(negIf0 / (2.0 * i0n));   /* :333-334 */
  const double posEta3 = pref * mko_asinh(posIf3 / (2.0 * i0p));
  const double Uocpn0 = uocp_n(p, tss0), Uocpp3 = uocp_p(p, tss3); /* :337-340 */
  const double Rfn = rf_e(p, P_NEG, negSOC, st[S_ARRRFN]), Rfp = rf_e(p, P_POS, posSOC, st[S_ARRRFP]);
  double V = ((((posEta3 - negEta0) + yk[8]) + Uocpp3) - Uocpn0) + (Rfp * posIfdl3 - Rfn * negIfdl0); /* :341-342 */
  V = V - (p[P_RC] * rcs) * Iapp;                                  /* :344 */
  st[S_SOCN] = SOCn;
  st[S_SOCP] = SOCp;
  return V;
}

/* ------------------------------------------------------------------------------------------ */
/* EKF: iterEKF.m (OB)                                                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  double gamma[4];
  int m[4];       /* model column index (iT*nZ+iZ) of the 4 neighbours */
  int iT[4], iZ[4];
} xind_t;

static void get_xind(const mko_bank *bk, double Tk, double SOC, xind_t *X) { /* iterEKF.m:219-255 */
  int z1, z2, t1, t2;
  two_nearest(SOC, bk->Zpts, bk->nZ, &z1, &z2);
  int iZup = z1, iZlo = z1;
  double Zup = bk->Zpts[z1], Zlo = bk->Zpts[z1];
  if (bk->nZ > 1) {
    Zup = bk->Zpts[z1]; iZup = z1; Zlo = bk->Zpts[z2]; iZlo = z2;
    if (Zup < Zlo) { Zup = bk->Zpts[z2]; iZup = z2; Zlo = bk->Zpts[z1]; iZlo = z1; }
  }
  two_nearest(Tk, bk->Tpts_K, bk->nT, &t1, &t2);
  int iTup = t1, iTlo = t1;
  double Tup = bk->Tpts_K[t1], Tlo = bk->Tpts_K[t1];
  if (bk->nT > 1) {
    Tup = bk->Tpts_K[t1]; iTup = t1; Tlo = bk->Tpts_K[t2]; iTlo = t2;
    if (Tup < Tlo) { Tup = bk->Tpts_K[t2]; iTup = t2; Tlo = bk->Tpts_K[t1]; iTlo = t1; }
  }
  double aZ = 0.0, aT = 0.0;
  if (bk->nZ > 1) aZ = (SOC - Zlo) / (Zup - Zlo);                  /* :247 (Q14: not clamped) */
  if (bk->nT > 1) aT = (Tk - Tlo) / (Tup - Tlo);
  X->gamma[0] = (1.0 - aT) * (1.0 - aZ);
  X->gamma[1] = (1.0 - aT) * aZ;
  X->gamma[2] = aT * (1.0 - aZ);
  X->gamma[3] = aT * aZ;
  X->iT[0] = iTlo; X->iT[1] = iTlo; X->iT[2] = iTup; X->iT[3] = iTup;
  X->iZ[0] = iZlo; X->iZ[1] = iZup; X->iZ[2] = iZlo; X->iZ[3] = iZup;
  for (int k = 0; k < 4; ++k) X->m[k] = X->iT[k] * bk->nZ + X->iZ[k];
}

/* getVariables (iterEKF.m:259-417): Z holds all nz corrected outputs; returns Vcell */
/* xmb: NULL for the output-blend filter; the model-blend filter's ekfData.xhat (n+1) otherwise (:266-275, :310-312) */
static double get_variables(const mko_bank *bk, double *st, double ik, const xind_t *X, double T, double *Z,
                            double *SOCn_out, double *SOCp_out, const double *xmb) {
  const double *p = bk->params;
  const int M = bk->nT * bk->nZ, nz = bk->nz;
  const double *xhat = st + st_off_xhat(M);
  const double rpa = bk->Ts / (3600.0 * p[P_Q]);
  const double xSOC = st[S_EKFSOC0] - (xmb ? xmb[5] : st[S_X0]) * rpa; /* :277 */
  double SOCn = soc_e(p, P_NEG, xSOC), SOCp = soc_e(p, P_POS, xSOC); /* :282-283 */
  if (SOCn < 0.0) { st[S_WARN] += 1.0; SOCn = 1e-6; }              /* :284-295 */
  if (SOCn > 1.0) { st[S_WARN] += 1.0; SOCn = 1.0 - 1e-6; }
  if (SOCp < 0.0) { st[S_WARN] += 1.0; SOCp = 1e-6; }
  if (SOCp > 0.998) { st[S_WARN] += 1.0; SOCp = 0.998; }
  for (int r = 0; r < nz; ++r) {                                   /* :302-313 */
    double yc[4];
    for (int c = 0; c < 4; ++c) {
      const double *cr = bk->C + ((size_t)X->m[c] * nz + r) * 6;
      const double *x = xmb ? xmb : xhat + N5 * X->m[c];
      double acc = cr[0] * x[0];
      for (int j = 1; j < N5; ++j) acc = fma(cr[j], x[j], acc);
      yc[c] = fma(bk->D[(size_t)X->m[c] * nz + r], ik, acc);
    }
    /* "[...]*Xind.gamma" (:313).  Structure exploited: a neighbour whose weight is EXACTLY zero adds
       nothing -- the reference adds its +-0 product, which changes nothing but the sign of an exact zero.  A cell whose
       temperature sits on a set-point of the bank (runMPC's 25 degC) has two such neighbours at every instant. */
    if (zw_skip) {
      double acc = 0.0;
      for (int c = 0; c < 4; ++c) if (X->gamma[c] != 0.0) acc = fma(yc[c], X->gamma[c], acc);
      Z[r] = acc;
    } else {
      Z[r] = fma(yc[3], X->gamma[3], fma(yc[2], X->gamma[2], fma(yc[1], X->gamma[1], yc[0] * X->gamma[0])));
    }
  }
  const double If0 = Z[bk->r_if0], If3 = Z[bk->r_if3];             /* :324-325 */
  int any;
  for (int k = 0; k < bk->n_negtss; ++k) Z[bk->negtss[k]] = Z[bk->negtss[k]] + SOCn; /* :331 */
  any = 0;
  for (int k = 0; k < bk->n_negtss; ++k) if (Z[bk->negtss[k]] < 0.0) { any = 1; Z[bk->negtss[k]] = 1e-6; }
  if (any) st[S_WARN] += 1.0;
  any = 0;
  for (int k = 0; k < bk->n_negtss; ++k) if (Z[bk->negtss[k]] > 1.0) { any = 1; Z[bk->negtss[k]] = 1.0 - 1e-6; }
  if (any) st[S_WARN] += 1.0;
  for (int k = 0; k < bk->n_postss; ++k) Z[bk->postss[k]] = Z[bk->postss[k]] + SOCp; /* :345 */
  any = 0;
  for (int k = 0; k < bk->n_postss; ++k) if (Z[bk->postss[k]] < 0.0) { any = 1; Z[bk->postss[k]] = 1e-6; }
  if (any) st[S_WARN] += 1.0;
  any = 0;
  for (int k = 0; k < bk->n_postss; ++k) if (Z[bk->postss[k]] > 0.998) { any = 1; Z[bk->postss[k]] = 0.998; }
  if (any) st[S_WARN] += 1.0;
  const double UnAvg = uocp_n(p, SOCn), UpAvg = uocp_p(p, SOCp);   /* :362-367 */
  for (int k = 0; k < bk->n_negphise; ++k) Z[bk->negphise[k]] = Z[bk->negphise[k]] + UnAvg;
  for (int k = 0; k < bk->n_posphise; ++k) Z[bk->posphise[k]] = Z[bk->posphise[k]] + UpAvg;
  const double PhieTilde3 = Z[bk->r_phie_last];                    /* :371 */
  /* Phie rows get "- Phise0" (:372-379); they feed only zk outputs, handled by the caller if wanted */
  for (int k = 0; k < bk->n_thetae; ++k) Z[bk->thetae[k]] = Z[bk->thetae[k]] + 1.0; /* :383 */
  any = 0;
  for (int k = 0; k < bk->n_thetae; ++k) if (Z[bk->thetae[k]] < 0.0) { any = 1; Z[bk->thetae[k]] = 1e-6; }
  if (any) { st[S_WARN] += 1.0; st[S_STATUS] = (double)((unsigned)st[S_STATUS] | ST_Q18); } /* Q18 */
  const double th0 = Z[bk->r_tss0], th3 = Z[bk->r_tss3];
  const double k0n = k0_e(p, P_NEG, SOCn, st[S_ARRK0N]), k0p = k0_e(p, P_POS, SOCp, st[S_ARRK0P]); /* :392-393 */
  const double i0n = k0n * sqrt((Z[bk->thetae[0]] * (1.0 - th0)) * th0);
  const double i0p = k0p * sqrt((Z[bk->thetae[bk->n_thetae - 1]] * (1.0 - th3)) * th3);
  const double pref = ((2.0 * p[P_R]) * T) / p[P_F];
  const double negEta0 = pref * mko_asinh(If0 / (2.0 * i0n));      /* :400-401 */
  const double posEta3 = pref * mko_asinh(If3 / (2.0 * i0p));
  const double Un0 = uocp_n(p, th0), Up3 = uocp_p(p, th3);         /* :404-407 */
  const double Rfn = rf_e(p, P_NEG, SOCn, st[S_ARRRFN]), Rfp = rf_e(p, P_POS, SOCp, st[S_ARRRFP]);
  const double V = ((((posEta3 - negEta0) + PhieTilde3) + Up3) - Un0) + (Rfp * Z[bk->r_ifdl3] - Rfn * Z[bk->r_ifdl0]); /* :409-410 (Q5) */
  *SOCn_out = SOCn;
  *SOCp_out = SOCp;
  return V;
}

/* PSD repair (iterEKF.m:143-145) on symmetric storage: returns 1 if the eigen path ran */
static int psd_repair(double *S, int force) {
  if (!force) {
    /* LDL^T: if every pivot is > 0 the matrix is PD, V*S*V' == SigmaX and :145 is the identity */
    double L[5][5], d[5];
    int pd = 1;
    for (int j = 0; j < 5 && pd; ++j) {
      double dj = S[sym_idx(j, j)];
      for (int k = 0; k < j; ++k) dj = fma(-(L[j][k] * L[j][k]), d[k], dj);
      if (!(dj > 0.0)) { pd = 0; break; }
      d[j] = dj;
      const double idj = 1.0 / dj;
      for (int i = j + 1; i < 5; ++i) {
        double v = S[sym_idx(i, j)];
        for (int k = 0; k < j; ++k) v = fma(-(L[i][k] * L[j][k]), d[k], v);
        L[i][j] = v * idj;
      }
    }
    if (pd) return 0;
  }
  /* cyclic Jacobi: A = V diag(w) V', then HH = V |w| V' */
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
        double apq = A[pI][q];
        if (apq == 0.0) continue;
        double theta = (A[q][q] - A[pI][pI]) / (2.0 * apq);
        double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
        if (theta < 0.0) t = -t;
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 5; ++k) { /* columns p,q of A */
          double akp = A[k][pI], akq = A[k][q];
          A[k][pI] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 5; ++k) { /* rows p,q of A */
          double apk = A[pI][k], aqk = A[q][k];
          A[pI][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 5; ++k) {
          double vkp = V[k][pI], vkq = V[k][q];
          V[k][pI] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < 5; ++i)
    for (int j = i; j < 5; ++j) {
      double hh = 0.0;
      for (int k = 0; k < 5; ++k) hh = hh + (V[i][k] * fabs(A[k][k])) * V[j][k];
      double s = S[sym_idx(i, j)];
      S[sym_idx(i, j)] = (((s + s) + hh) + hh) / 4.0;             /* :145 */
    }
  return 1;
}

typedef struct {
  double zneed[8]; /* If0, If3, tss0, tss3, thetae_first, thetae_last, (spare) */
  double vhat, soc;
  int iT, iZ;      /* picked model (0-based) */
  int mpick;
  double Cv[N5], Dv, Cphi[N5], Dphi, bv, bphi, a[N5], xhat[6];
  int lost;
} ekf_out;

/* EKFmatsHandler.m:26-114 on the corrected outputs Z of the last getVariables call; x_int = ekfData.xhat(end) (:33, Q3) */
static void ekf_mats(const mko_bank *bk, double *st, const xind_t *Xp, const double *Z, double Zsoc, double x_int, ekf_out *o) {
  const double *p = bk->params;
  const int M = bk->nT * bk->nZ, nz = bk->nz;
  const double *xhat = st + st_off_xhat(M);
  const xind_t X = *Xp;
  OPC_BLOCK(4);
  int imax = 0;                                                    /* :26 first max, NaN ignored */
  {
    int have = 0;
    double best = 0.0;
    for (int c = 0; c < 4; ++c) {
      double g = X.gamma[c];
      if (g != g) continue;
      if (!have || g > best) { best = g; imax = c; have = 1; }
    }
  }
  const int mp = X.m[imax];
  o->iT = X.iT[imax];
  o->iZ = X.iZ[imax];
  o->mpick = mp;
  for (int j = 0; j < N5; ++j) { o->xhat[j] = xhat[N5 * mp + j]; o->a[j] = bk->A[6 * mp + j]; }
  o->xhat[5] = x_int;                                              /* :33 (Q3) ekfData.xhat(end): 0 in OB mode */
  {
    const double TK = st[S_TC] + 273.15;                           /* :53 */
    const double SOCavg = Zsoc;                                    /* :56 */
    const double sn = soc_e(p, P_NEG, SOCavg), sp = soc_e(p, P_POS, SOCavg);
    const double k0n = k0_e(p, P_NEG, sn, st[S_ARRK0N]), k0p = k0_e(p, P_POS, sp, st[S_ARRK0P]);
    const double th0 = Z[bk->r_tss0], th3 = Z[bk->r_tss3];
    const double i0n = k0n * sqrt((Z[bk->thetae[0]] * (1.0 - th0)) * th0);   /* :63-66 */
    const double i0p = k0p * sqrt((Z[bk->thetae[bk->n_thetae - 1]] * (1.0 - th3)) * th3);
    const double Rfn = rf_e(p, P_NEG, sn, st[S_ARRRFN]), Rfp = rf_e(p, P_POS, sp, st[S_ARRRFP]);
    const double *Cm = bk->C + (size_t)mp * nz * 6, *Dm = bk->D + (size_t)mp * nz;
    for (int j = 0; j < N5; ++j)                                   /* :75-76 */
      o->Cv[j] = (Rfp * Cm[bk->r_ifdl3 * 6 + j] - Rfn * Cm[bk->r_ifdl0 * 6 + j]) + Cm[bk->r_phie_last * 6 + j];
    o->Dv = (Rfp * Dm[bk->r_ifdl3] - Rfn * Dm[bk->r_ifdl0]) + Dm[bk->r_phie_last]; /* :79-80 */
    const double Upos = uocp_p(p, th3), Uneg = uocp_n(p, th0);     /* :84-85 */
    const double pref = ((2.0 * p[P_R]) * TK) / p[P_F];
    const double negEta0 = pref * mko_asinh(Z[bk->r_if0] / (2.0 * i0n)); /* :88-89 */
    const double posEta3 = pref * mko_asinh(Z[bk->r_if3] / (2.0 * i0p));
    o->bv = ((Upos - Uneg) + (posEta3 - negEta0)) + 0.0;           /* :92 */
    o->bphi = uocp_n(p, sn);                                       /* :96 */
    for (int j = 0; j < N5; ++j) o->Cphi[j] = Cm[bk->r_phi * 6 + j]; /* :97 */
    o->Dphi = Dm[bk->r_phi];                                       /* :98 */
  }
}

/* one EKF iteration + EKFmatsHandler; Zfull (nz+2) optional */
static void ekf_step(const mko_bank *bk, const mko_config *cf, double *st, double vk, double ik, ekf_out *o, double *zk_out) {
  const double *p = bk->params;
  const int M = bk->nT * bk->nZ, nz = bk->nz;
  double *xhat = st + st_off_xhat(M), *Sig = st + st_off_sig(M);
  double Z[64 + 2];
  OPC_BLOCK(1);
  o->lost = 0;
  if (st[S_WARN] > 10.0) {                                         /* iterEKF.m:55-59 (Q19) */
    st[S_STATUS] = (double)((unsigned)st[S_STATUS] | ST_LOST);
    o->lost = 1;
    if (zk_out) for (int r = 0; r < nz + 2; ++r) zk_out[r] = NAN;
    return;
  }
  const double Tk = st[S_TC] + 273.15;                             /* :65 */
  const double prior = st[S_PRIORI], Sw = cf->SigmaW;
  for (int m = 0; m < M; ++m) {                                    /* :73-84 */
    const double *a = bk->A + 6 * m;
    double *x = xhat + N5 * m, *S = Sig + NSYM * m;
    for (int j = 0; j < N5; ++j) x[j] = fma(a[j], x[j], prior);    /* :77 */
    int q = 0;                                                     /* :81 (Q12): entry (i,j) of diag(A)*SigmaX*diag(A) + SigmaW with the two factors */
    for (int i = 0; i < N5; ++i)                                   /* multiplied first, (a_i*a_j)*SigmaX(i,j) (numerics contract) */
      for (int j = i; j < N5; ++j, ++q) S[q] = fma(a[i] * a[j], S[q], Sw);
  }
  OPC_BLOCK(2);
  st[S_X0] = st[S_X0] + prior;                                     /* :85 */
  st[S_SIGX0] = st[S_SIGX0] + Sw;                                  /* :86 */
  const double rpa = bk->Ts / (3600.0 * p[P_Q]);
  double SOC = st[S_EKFSOC0] - st[S_X0] * rpa;                     /* :89 */
  xind_t X;
  get_xind(bk, Tk, SOC, &X);                                       /* :107 */
  double SOCn, SOCp;
  double vhat = get_variables(bk, st, ik, &X, Tk, Z, &SOCn, &SOCp, NULL); /* :108 */
  /* ---- getChatV (:421-519); SOCn/SOCp re-derived there from the same x0 (unclamped, :438-440) ---- */
  double Chat[4][N5], C0;
  {
    const double xSOC = st[S_EKFSOC0] - st[S_X0] * rpa;
    const double sn = soc_e(p, P_NEG, xSOC), sp = soc_e(p, P_POS, xSOC);
    const double Rfn = rf_e(p, P_NEG, sn, st[S_ARRRFN]), Rfp = rf_e(p, P_POS, sp, st[S_ARRRFP]);
    const double k0n = k0_e(p, P_NEG, sn, st[S_ARRK0N]), k0p = k0_e(p, P_POS, sp, st[S_ARRK0P]);
    const double th0 = Z[bk->r_tss0], th3 = Z[bk->r_tss3];
    const double i0n = k0n * sqrt((Z[bk->thetae[0]] * (1.0 - th0)) * th0);
    const double i0p = k0p * sqrt((Z[bk->thetae[bk->n_thetae - 1]] * (1.0 - th3)) * th3);
    const double Rctn = (p[P_R] * Tk) / (p[P_F] * i0n), Rctp = (p[P_R] * Tk) / (p[P_F] * i0p); /* :467-468 */
    const double dUn0 = duocp_n(p, th0), dUp3 = duocp_p(p, th3);   /* :495-496 */
    const double dsn = soc_e(p, P_NEG, 1.0) - soc_e(p, P_NEG, 0.0), dsp = soc_e(p, P_POS, 1.0) - soc_e(p, P_POS, 0.0);
    const double res0n = (((-dUn0) * bk->Ts) * dsn) / (3600.0 * p[P_Q]); /* :497-500 */
    const double res0p = (((-dUp3) * bk->Ts) * dsp) / (3600.0 * p[P_Q]);
    C0 = res0p - res0n;                                            /* :501 */
    for (int c = 0; c < 4; ++c) {
      const double *Cm = bk->C + (size_t)X.m[c] * nz * 6;
      const double g = X.gamma[c];
      for (int j = 0; j < N5; ++j) {
        double v = Rfp * (g * Cm[bk->r_ifdl3 * 6 + j]) - Rfn * (g * Cm[bk->r_ifdl0 * 6 + j]);       /* :444 */
        v = (v + Rctp * (g * Cm[bk->r_if3 * 6 + j])) - Rctn * (g * Cm[bk->r_if0 * 6 + j]);          /* :471 */
        v = v + g * Cm[bk->r_phie_last * 6 + j];                                                     /* :485 */
        v = v + (dUp3 * (g * Cm[bk->r_tss3 * 6 + j]) - dUn0 * (g * Cm[bk->r_tss0 * 6 + j]));        /* :504 */
        Chat[c][j] = v;
      }
    }
  }
  /* ---- gains (:116-124), all from neighbour #1's SigmaX (Q1) ---- */
  double L[4][N5], svt[4];
  {
    const double *S1 = Sig + NSYM * X.m[0];
    for (int c = 0; c < 4; ++c) {
      double v[N5];
      for (int i = 0; i < N5; ++i) {
        double acc = S1[sym_idx(i, 0)] * Chat[c][0];
        for (int j = 1; j < N5; ++j) acc = fma(S1[sym_idx(i, j)], Chat[c][j], acc);
        v[i] = acc;
      }
      double s = Chat[c][0] * v[0];
      for (int i = 1; i < N5; ++i) s = fma(Chat[c][i], v[i], s);
      svt[c] = s + cf->SigmaV;
      const double isv = 1.0 / svt[c];
      for (int i = 0; i < N5; ++i) L[c][i] = v[i] * isv;
    }
  }
  const double svt0 = (C0 * st[S_SIGX0]) * C0 + cf->SigmaV;        /* :123 */
  const double L0 = (st[S_SIGX0] * C0) / svt0;                     /* :124 */
  const double resid = vk - vhat;                                  /* :133 */
  for (int c = 0; c < 4; ++c) {                                    /* :136-154 (Q2: sequential, duplicates re-processed) */
    double *x = xhat + N5 * X.m[c], *S = Sig + NSYM * X.m[c];
    if (zw_skip && X.gamma[c] == 0.0) {
      /* zero weight: Chat = 0, the gain is 0, svt = SigmaV, xhat and SigmaX stay as they are (:138,:142 add +-0) and the repair
         (:143-145) of an unchanged covariance is the identity up to rounding: skipped.  The Q-bump (:148-151) is what is left. */
      if (resid * resid > 9.0 * (0.0 + cf->SigmaV)) {
        for (int k = 0; k < NSYM; ++k) S[k] = S[k] * 2.0;
        st[S_NBUMP] += 1.0;
      }
      continue;
    }
    for (int j = 0; j < N5; ++j) x[j] = fma(L[c][j], resid, x[j]);    /* :138 */
    int q = 0;
    for (int i = 0; i < N5; ++i)
      for (int j = i; j < N5; ++j, ++q) S[q] = fma(-(L[c][i] * svt[c]), L[c][j], S[q]); /* :142 */
    OPC_BLOCK(3);
    const int jac = psd_repair(S, cf->force_jacobi);               /* :143-145 */
    OPC_BLOCK(2);
    if (jac) st[S_NJAC] += 1.0;
    if (resid * resid > 9.0 * svt[c]) {                            /* :148-151 */
      for (int k = 0; k < NSYM; ++k) S[k] = S[k] * 2.0;
      st[S_NBUMP] += 1.0;
    }
  }
  st[S_X0] = st[S_X0] + L0 * resid;                                /* :155 */
  st[S_SIGX0] = st[S_SIGX0] - (L0 * svt0) * L0;                    /* :156 */
  SOC = st[S_EKFSOC0] - st[S_X0] * rpa;                            /* :159 */
  get_xind(bk, Tk, SOC, &X);                                       /* :181 */
  vhat = get_variables(bk, st, ik, &X, Tk, Z, &SOCn, &SOCp, NULL); /* :182 */
  const double Zsoc = st[S_EKFSOC0] - st[S_X0] * rpa;              /* :416 */
  st[S_PRIORI] = ik;                                               /* :210 */
  if (zk_out) {
    /* full zk = [Z; vhat; SOC] (:183) incl. the output-only corrections get_variables skips:
       Phie rows minus Phise0 (:372-379; no row has loc 0 after set-up) and posPhis + Vcell (:413) */
    for (int r = 0; r < nz; ++r) zk_out[r] = Z[r];
    const double Phise0 = Z[bk->r_phise0];
    for (int k = 0; k < bk->n_phie; ++k) zk_out[bk->phie[k]] = Z[bk->phie[k]] - Phise0;
    for (int k = 0; k < bk->n_posphis; ++k) zk_out[bk->posphis[k]] = Z[bk->posphis[k]] + vhat;
    zk_out[nz] = vhat;
    zk_out[nz + 1] = Zsoc;
  }
  o->vhat = vhat;
  o->soc = Zsoc;
  ekf_mats(bk, st, &X, Z, Zsoc, 0.0, o);                           /* EKFmatsHandler.m:26-114 */
}

/* ------------------------------------------------------------------------------------------ */
/* EKF: iterEKF.m, model-blend method ('MB': initKF.m:46-47,100-101; iterEKF.m:90-102,125-128,  */
/* 160-176,199-203 and the MB branches of getVariables / getChatV / getChatZ)                   */
/* ------------------------------------------------------------------------------------------ */
static int sym6(int i, int j) { /* upper triangle of a 6x6, row-major */
  static const int base[6] = {0, 6, 11, 15, 18, 20};
  return (i <= j) ? base[i] + (j - i) : base[j] + (i - j);
}

/* PSD repair (iterEKF.m:164-166) of the 6x6 covariance: same two paths as psd_repair */
static int psd_repair6(double *S, int force) {
  if (!force) {
    double L[6][6], d[6];
    int pd = 1;
    for (int j = 0; j < 6 && pd; ++j) {
      double dj = S[sym6(j, j)];
      for (int k = 0; k < j; ++k) dj = fma(-(L[j][k] * L[j][k]), d[k], dj);
      if (!(dj > 0.0)) { pd = 0; break; }
      d[j] = dj;
      const double idj = 1.0 / dj;
      for (int i = j + 1; i < 6; ++i) {
        double v = S[sym6(i, j)];
        for (int k = 0; k < j; ++k) v = fma(-(L[i][k] * L[j][k]), d[k], v);
        L[i][j] = v * idj;
      }
    }
    if (pd) return 0;
  }
  double A[6][6], V[6][6];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) { A[i][j] = S[sym6(i, j)]; V[i][j] = (i == j) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int i = 0; i < 6; ++i) {
      dg = dg + A[i][i] * A[i][i];
      for (int j = i + 1; j < 6; ++j) off = off + A[i][j] * A[i][j];
    }
    if (!(off > 1e-34 * dg)) break;
    for (int pI = 0; pI < 5; ++pI)
      for (int q = pI + 1; q < 6; ++q) {
        double apq = A[pI][q];
        if (apq == 0.0) continue;
        double theta = (A[q][q] - A[pI][pI]) / (2.0 * apq);
        double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
        if (theta < 0.0) t = -t;
        double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < 6; ++k) {
          double akp = A[k][pI], akq = A[k][q];
          A[k][pI] = c * akp - sn * akq;
          A[k][q] = sn * akp + c * akq;
        }
        for (int k = 0; k < 6; ++k) {
          double apk = A[pI][k], aqk = A[q][k];
          A[pI][k] = c * apk - sn * aqk;
          A[q][k] = sn * apk + c * aqk;
        }
        for (int k = 0; k < 6; ++k) {
          double vkp = V[k][pI], vkq = V[k][q];
          V[k][pI] = c * vkp - sn * vkq;
          V[k][q] = sn * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j) {
      double hh = 0.0;
      for (int k = 0; k < 6; ++k) hh = hh + (V[i][k] * fabs(A[k][k])) * V[j][k];
      double sv = S[sym6(i, j)];
      S[sym6(i, j)] = (((sv + sv) + hh) + hh) / 4.0;               /* :166 */
    }
  return 1;
}

/* one iteration of the model-blend filter + EKFmatsHandler; zk_out (nz+2) and bound_out (nz+2) optional */
static void ekf_step_mb(const mko_bank *bk, const mko_config *cf, double *st, double vk, double ik, ekf_out *o, double *zk_out,
                        double *bound_out) {
  const double *p = bk->params;
  const int M = bk->nT * bk->nZ, nz = bk->nz;
  double *x = st + st_off_mbx(M), *S = st + st_off_mbs(M);
  double Z[64 + 2];
  o->lost = 0;
  if (st[S_WARN] > 10.0) {                                         /* iterEKF.m:55-59 (Q19) */
    st[S_STATUS] = (double)((unsigned)st[S_STATUS] | ST_LOST);
    o->lost = 1;
    if (zk_out) for (int r = 0; r < nz + 2; ++r) zk_out[r] = NAN;
    if (bound_out) for (int r = 0; r < nz + 2; ++r) bound_out[r] = NAN;
    return;
  }
  const double Tk = st[S_TC] + 273.15;                             /* :65 */
  const double prior = st[S_PRIORI], Sw = cf->SigmaW;
  const double rpa = bk->Ts / (3600.0 * p[P_Q]);
  double SOC = st[S_EKFSOC0] - x[5] * rpa;                         /* :92 */
  xind_t X;
  get_xind(bk, Tk, SOC, &X);                                       /* :93 */
  double amb[N6];
  for (int j = 0; j < N5; ++j) {                                   /* :94-98 [A1,A2,A3,A4]*gamma */
    const double a0 = bk->A[6 * X.m[0] + j], a1 = bk->A[6 * X.m[1] + j], a2 = bk->A[6 * X.m[2] + j], a3 = bk->A[6 * X.m[3] + j];
    amb[j] = fma(a3, X.gamma[3], fma(a2, X.gamma[2], fma(a1, X.gamma[1], a0 * X.gamma[0])));
  }
  amb[5] = 1.0;
  for (int j = 0; j < N6; ++j) x[j] = fma(amb[j], x[j], prior);    /* :99 */
  for (int i = 0; i < N6; ++i)
    for (int j = i; j < N6; ++j) S[sym6(i, j)] = fma(amb[i] * S[sym6(i, j)], amb[j], Sw);   /* :101 */
  SOC = st[S_EKFSOC0] - x[5] * rpa;                                /* :102 */
  get_xind(bk, Tk, SOC, &X);                                       /* :107 */
  double SOCn, SOCp;
  double vhat = get_variables(bk, st, ik, &X, Tk, Z, &SOCn, &SOCp, x);   /* :108 */
  /* gamma-weighted row sums sum([C1(r,:); C2(r,:); C3(r,:); C4(r,:)]) with Ck = gamma(k)*C (:423-426) */
#define MB_RSUM(r, j)                                                                                                  \
  (((X.gamma[0] * bk->C[((size_t)X.m[0] * nz + (r)) * 6 + (j)] + X.gamma[1] * bk->C[((size_t)X.m[1] * nz + (r)) * 6 + (j)]) + \
    X.gamma[2] * bk->C[((size_t)X.m[2] * nz + (r)) * 6 + (j)]) +                                                       \
   X.gamma[3] * bk->C[((size_t)X.m[3] * nz + (r)) * 6 + (j)])
  /* ---- getChatV, MB branches (:448-460,475-479,489-491,512-518) ---- */
  double Chat[N6];
  {
    const double xSOC = st[S_EKFSOC0] - x[5] * rpa;                /* :449-450 */
    const double sn = soc_e(p, P_NEG, xSOC), sp = soc_e(p, P_POS, xSOC);
    const double Rfn = rf_e(p, P_NEG, sn, st[S_ARRRFN]), Rfp = rf_e(p, P_POS, sp, st[S_ARRRFP]);
    const double k0n = k0_e(p, P_NEG, sn, st[S_ARRK0N]), k0p = k0_e(p, P_POS, sp, st[S_ARRK0P]);
    const double th0 = Z[bk->r_tss0], th3 = Z[bk->r_tss3];
    const double i0n = k0n * sqrt((Z[bk->thetae[0]] * (1.0 - th0)) * th0);
    const double i0p = k0p * sqrt((Z[bk->thetae[bk->n_thetae - 1]] * (1.0 - th3)) * th3);
    const double Rctn = (p[P_R] * Tk) / (p[P_F] * i0n), Rctp = (p[P_R] * Tk) / (p[P_F] * i0p);
    const double dUn0 = duocp_n(p, th0), dUp3 = duocp_p(p, th3);
    const double dsn = soc_e(p, P_NEG, 1.0) - soc_e(p, P_NEG, 0.0), dsp = soc_e(p, P_POS, 1.0) - soc_e(p, P_POS, 0.0);
    const double res0n = (((-dUn0) * bk->Ts) * dsn) / (3600.0 * p[P_Q]);
    const double res0p = (((-dUp3) * bk->Ts) * dsp) / (3600.0 * p[P_Q]);
    for (int j = 0; j < N5; ++j) {
      double v = Rfp * MB_RSUM(bk->r_ifdl3, j) - Rfn * MB_RSUM(bk->r_ifdl0, j);       /* :456-459 */
      v = (v + Rctp * MB_RSUM(bk->r_if3, j)) - Rctn * MB_RSUM(bk->r_if0, j);           /* :476-479 */
      v = v + MB_RSUM(bk->r_phie_last, j);                                             /* :490-491 */
      v = (v + dUp3 * MB_RSUM(bk->r_tss3, j)) - dUn0 * MB_RSUM(bk->r_tss0, j);         /* :513-516 */
      Chat[j] = v;
    }
    Chat[5] = res0p - res0n;                                       /* :501, :517 */
  }
  /* ---- gain (:126-127) ---- */
  double v6[N6], L[N6];
  for (int i = 0; i < N6; ++i) {
    double acc = S[sym6(i, 0)] * Chat[0];
    for (int j = 1; j < N6; ++j) acc = fma(S[sym6(i, j)], Chat[j], acc);
    v6[i] = acc;
  }
  double sq = Chat[0] * v6[0];
  for (int i = 1; i < N6; ++i) sq = fma(Chat[i], v6[i], sq);
  const double svt = sq + cf->SigmaV;
  const double isv = 1.0 / svt;
  for (int i = 0; i < N6; ++i) L[i] = v6[i] * isv;
  const double resid = vk - vhat;                                  /* :133 */
  for (int j = 0; j < N6; ++j) x[j] = fma(L[j], resid, x[j]);      /* :161 */
  for (int i = 0; i < N6; ++i)
    for (int j = i; j < N6; ++j) S[sym6(i, j)] = fma(-(L[i] * svt), L[j], S[sym6(i, j)]);   /* :163 */
  if (psd_repair6(S, cf->force_jacobi)) st[S_NJAC] += 1.0;         /* :164-166 */
  if (resid * resid > 9.0 * svt) {                                 /* :169-172 */
    for (int k = 0; k < NSYM6; ++k) S[k] = S[k] * 2.0;
    st[S_NBUMP] += 1.0;
  }
  SOC = st[S_EKFSOC0] - x[5] * rpa;                                /* :176 */
  get_xind(bk, Tk, SOC, &X);                                       /* :181 */
  vhat = get_variables(bk, st, ik, &X, Tk, Z, &SOCn, &SOCp, x);    /* :182 */
  const double Zsoc = st[S_EKFSOC0] - x[5] * rpa;                  /* :416 */
  st[S_PRIORI] = ik;                                               /* :210 */
  double zk[64 + 2];
  for (int r = 0; r < nz; ++r) zk[r] = Z[r];
  {
    const double Phise0 = Z[bk->r_phise0];
    for (int k = 0; k < bk->n_phie; ++k) zk[bk->phie[k]] = Z[bk->phie[k]] - Phise0;
    for (int k = 0; k < bk->n_posphis; ++k) zk[bk->posphis[k]] = Z[bk->posphis[k]] + vhat;
    zk[nz] = vhat;
    zk[nz + 1] = Zsoc;
  }
  if (zk_out) for (int r = 0; r < nz + 2; ++r) zk_out[r] = zk[r];
  if (bound_out) {                                                 /* :199-205 with getChatZ's MB branches (:538,554-558,596-600) */
    const double xSOC = st[S_EKFSOC0] - x[5] * rpa;
    const double sn = soc_e(p, P_NEG, xSOC), sp = soc_e(p, P_POS, xSOC);
    const double Rfn = rf_e(p, P_NEG, sn, st[S_ARRRFN]), Rfp = rf_e(p, P_POS, sp, st[S_ARRRFP]);
    const double k0n = k0_e(p, P_NEG, sn, st[S_ARRK0N]), k0p = k0_e(p, P_POS, sp, st[S_ARRK0P]);
    const double th0 = zk[bk->r_tss0], th3 = zk[bk->r_tss3];
    const double i0n = k0n * sqrt((zk[bk->thetae[0]] * (1.0 - th0)) * th0);
    const double i0p = k0p * sqrt((zk[bk->thetae[bk->n_thetae - 1]] * (1.0 - th3)) * th3);
    const double Rctn = (p[P_R] * Tk) / (p[P_F] * i0n), Rctp = (p[P_R] * Tk) / (p[P_F] * i0p);
    const double dUn0 = duocp_n(p, th0), dUp3 = duocp_p(p, th3);
    const double dsn = soc_e(p, P_NEG, 1.0) - soc_e(p, P_NEG, 0.0), dsp = soc_e(p, P_POS, 1.0) - soc_e(p, P_POS, 0.0);
    double CV[N6];
    for (int j = 0; j < N5; ++j) {
      double v = Rfp * MB_RSUM(bk->r_ifdl3, j) - Rfn * MB_RSUM(bk->r_ifdl0, j);
      v = (v + Rctp * MB_RSUM(bk->r_if3, j)) - Rctn * MB_RSUM(bk->r_if0, j);
      v = v + MB_RSUM(bk->r_phie_last, j);
      v = (v + dUp3 * MB_RSUM(bk->r_tss3, j)) - dUn0 * MB_RSUM(bk->r_tss0, j);
      CV[j] = v;
    }
    CV[5] = (((-dUp3) * bk->Ts) * dsp) / (3600.0 * p[P_Q]) - (((-dUn0) * bk->Ts) * dsn) / (3600.0 * p[P_Q]);
    const double res0n = ((-bk->Ts) * dsn) / (3600.0 * p[P_Q]), res0p = ((-bk->Ts) * dsp) / (3600.0 * p[P_Q]);   /* :562-565 */
    const double dUn = duocp_n(p, sn), dUp = duocp_p(p, sp);       /* :579-580 */
    unsigned char kind[64];
    memset(kind, 0, sizeof kind);
    for (int k = 0; k < bk->n_posphis; ++k) kind[bk->posphis[k]] = 1;
    for (int k = 0; k < bk->n_negtss; ++k) kind[bk->negtss[k]] = 2;
    for (int k = 0; k < bk->n_postss; ++k) kind[bk->postss[k]] = 3;
    for (int k = 0; k < bk->n_negphise; ++k) kind[bk->negphise[k]] = 4;
    for (int k = 0; k < bk->n_posphise; ++k) kind[bk->posphise[k]] = 5;
    for (int k = 0; k < bk->n_phie; ++k) kind[bk->phie[k]] = 6;
    for (int r = 0; r <= nz; ++r) {
      double row[N6];
      if (r < nz) {
        double c0 = 0.0;
        if (kind[r] == 1) c0 = CV[5];
        else if (kind[r] == 2) c0 = res0n;
        else if (kind[r] == 3) c0 = res0p;
        else if (kind[r] == 4) c0 = dUn * res0n;
        else if (kind[r] == 5) c0 = dUp * res0p;
        else if (kind[r] == 6) c0 = -dUn * res0n;
        for (int j = 0; j < N5; ++j) {
          double v = MB_RSUM(r, j);
          if (kind[r] == 1) v = v + CV[j];                         /* :555-557 */
          if (kind[r] == 6) v = v - MB_RSUM(bk->r_phise0, j);      /* :597-599 */
          row[j] = v;
        }
        row[5] = c0;
      } else {
        for (int j = 0; j < N6; ++j) row[j] = CV[j];
      }
      double t = 0.0;
      for (int i = 0; i < N6; ++i) {
        double acc = 0.0;
        for (int j = 0; j < N6; ++j) acc += S[sym6(i, j)] * row[j];
        t += row[i] * acc;
      }
      bound_out[r] = 3.0 * sqrt(t);                                /* :205 */
    }
    const double res = -bk->Ts / (3600.0 * p[P_Q]);
    bound_out[nz + 1] = 3.0 * sqrt((res * S[sym6(5, 5)]) * res);   /* :203 */
  }
#undef MB_RSUM
  o->vhat = vhat;
  o->soc = Zsoc;
  ekf_mats(bk, st, &X, Z, Zsoc, x[5], o);                          /* EKFmatsHandler.m:26-114; M(iT,iZ).xhat stays 0 in MB mode */
}

/* ------------------------------------------------------------------------------------------ */
/* predMat.m for one scalar output row with diagonal A, B = 1                                 */
/*   h[k]   = Cbar*Abar^k*Bbar, k = 0..Np          (predMat.m:31-36; h[i] is also Phi(i,7))   */
/*   phix[i-1] = Phi(i,1:6)*x                         (predMat.m:40-43)                       */
/* ------------------------------------------------------------------------------------------ */
static void pred_row(const double *a, const double *c5, double c6, double d, const double *x6, int Np,
                     double *h, double *phix) {
  double pk[6], pw[6];
  for (int j = 0; j < 6; ++j) { pk[j] = 0.0; pw[j] = 1.0; }
  for (int k = 0; k <= Np; ++k) {
    double acc = c5[0] * pk[0];
    for (int j = 1; j < N5; ++j) acc = fma(c5[j], pk[j], acc);
    acc = fma(c6, pk[5], acc);
    h[k] = acc + d;                                                /* Cbar*X, X = [p;1] */
    for (int j = 0; j < N5; ++j) pk[j] = fma(a[j], pk[j], 1.0);    /* X = Abar*X */
    pk[5] = pk[5] + 1.0;
  }
  for (int i = 1; i <= Np; ++i) {
    for (int j = 0; j < N5; ++j) pw[j] = a[j] * pw[j];             /* Ap = Abar*Ap */
    double acc = (c5[0] * pw[0]) * x6[0];
    for (int j = 1; j < N5; ++j) acc = fma(c5[j] * pw[j], x6[j], acc);
    acc = fma(c6 * pw[5], x6[5], acc);
    phix[i - 1] = acc;
  }
}

/* solve E x = b, E SPD Nc x Nc, Gaussian elimination without pivoting (factor held in LU) */
static void spd_factor(int Nc, const double *E, double *LU) {
  /* LU holds the unit-lower multipliers below the diagonal, U on and above it, and the RECIPROCAL of
     each pivot on the diagonal, so that the many solves against one factor need no division */
  for (int i = 0; i < Nc * Nc; ++i) LU[i] = E[i];
  for (int k = 0; k < Nc; ++k) {
    const double ip = 1.0 / LU[k * Nc + k];
    LU[k * Nc + k] = ip;
    for (int i = k + 1; i < Nc; ++i) {
      double l = LU[i * Nc + k] * ip;
      LU[i * Nc + k] = l;
      for (int j = k + 1; j < Nc; ++j) LU[i * Nc + j] = fma(-l, LU[k * Nc + j], LU[i * Nc + j]);
    }
  }
}
static void spd_solve(int Nc, const double *LU, const double *b, double *x) {
  double y[NC_MAX];
  for (int i = 0; i < Nc; ++i) {
    double acc = b[i];
    for (int k = 0; k < i; ++k) acc = fma(-LU[i * Nc + k], y[k], acc);
    y[i] = acc;
  }
  for (int i = Nc - 1; i >= 0; --i) {
    double acc = y[i];
    for (int k = i + 1; k < Nc; ++k) acc = fma(-LU[i * Nc + k], x[k], acc);
    x[i] = acc * LU[i * Nc + i];
  }
}

/* Hildreth (hildreth.m:16-46) in rank-Nc form: H(i,:)*lambda = M(i,:)*v, v = (E\M')*lambda.
   Rows with H_ii == 0 (all-zero M row, Q7) never touch v; while one of them holds +Inf every
   row product is NaN exactly as 0*Inf makes it in the dense form ("poison").              */
int mko_hildreth(int Nc, int nC, const double *E, const double *F, const double *Mm, const double *gam,
                 const double *lam0, int maxIter, double *DU, double *lam, int *nexec) {
  double LU[NC_MAX * NC_MAX], Ei[NC_MAX][NC_MAX], P[NCON_MAX][NC_MAX], K[NCON_MAX], Hd[NCON_MAX], rH[NCON_MAX], v[NC_MAX], x[NC_MAX];
  if (Nc > NC_MAX || nC > NCON_MAX) return -1;
  OPC_BLOCK(7);
  spd_factor(Nc, E, LU);
  spd_solve(Nc, LU, F, x);                                         /* E\F */
  for (int l = 0; l < Nc; ++l) {                                   /* inv(E), column by column: E\M' = inv(E)*M' */
    double el[NC_MAX], col[NC_MAX];
    for (int j = 0; j < Nc; ++j) el[j] = (j == l) ? 1.0 : 0.0;
    spd_solve(Nc, LU, el, col);
    for (int j = 0; j < Nc; ++j) Ei[j][l] = col[j];
  }
  for (int i = 0; i < nC; ++i) {
    for (int j = 0; j < Nc; ++j) {                                 /* column i of E\M' */
      double acc = Ei[j][0] * Mm[i * Nc];
      for (int l = 1; l < Nc; ++l) acc = fma(Ei[j][l], Mm[i * Nc + l], acc);
      P[i][j] = acc;
    }
    double hd = Mm[i * Nc] * P[i][0], k = Mm[i * Nc] * x[0];
    for (int j = 1; j < Nc; ++j) { hd = fma(Mm[i * Nc + j], P[i][j], hd); k = fma(Mm[i * Nc + j], x[j], k); }
    Hd[i] = hd;                                                    /* H(i,i)  (:28) */
    rH[i] = 1.0 / hd;                                              /* the sweep multiplies by 1/H(i,i) (+Inf on a zero row) */
    K[i] = k + gam[i];                                             /* K(i)    (:29) */
    lam[i] = lam0 ? lam0[i] : 0.0;                                 /* :18-22 */
  }
  int npoison = 0;
  for (int j = 0; j < Nc; ++j) v[j] = 0.0;
  for (int i = 0; i < nC; ++i) {
    if (Hd[i] == 0.0) { if (!(fabs(lam[i]) <= 1.79769313486231570815e308)) npoison++; continue; }
    for (int j = 0; j < Nc; ++j) v[j] = fma(P[i][j], lam[i], v[j]);
  }
  int k = 0;
  OPC_BLOCK(8);
  for (k = 1; k <= maxIter; ++k) {                                 /* :32 */
    int notconv = 0;
    for (int i = 0; i < nC; ++i) {                                 /* :34-37 */
      const double li = lam[i];
      double w;
      if (npoison > 0) {
        w = NAN;
      } else {
        double hl = Mm[i * Nc] * v[0];
        for (int j = 1; j < Nc; ++j) hl = fma(Mm[i * Nc + j], v[j], hl);
        w = fma(-(K[i] + hl), rH[i], li);                       /* == -(K + H(i,:)*lam - H_ii*lam_i)/H_ii */
      }
      const double nl = (w > 0.0) ? w : 0.0;                       /* max(0,w); max(0,NaN) = 0 */
      const double dl = nl - li;
      if (Hd[i] == 0.0) {
        const int was = !(fabs(li) <= 1.79769313486231570815e308), is = !(fabs(nl) <= 1.79769313486231570815e308);
        npoison += is - was;
      } else {
        for (int j = 0; j < Nc; ++j) v[j] = fma(P[i][j], dl, v[j]);
      }
      lam[i] = nl;
      if (!(fabs(dl) < 1e-6)) notconv = 1;                         /* :39: norm(.,Inf) < 1e-6, NaN-propagating */
    }
    if (!notconv) break;
  }
  OPC_BLOCK(7);
  if (k > maxIter) k = maxIter;
  *nexec = k;                                                      /* :43 */
  double rhs[NC_MAX];
  for (int j = 0; j < Nc; ++j) {                                   /* :46 */
    double acc = Mm[j] * lam[0];
    for (int i = 1; i < nC; ++i) acc = fma(Mm[i * Nc + j], lam[i], acc);
    rhs[j] = F[j] + acc;
  }
  spd_solve(Nc, LU, rhs, x);
  for (int j = 0; j < Nc; ++j) DU[j] = -x[j];
  return 0;
}

/* smallest singular value of G_soc'*G_soc (iterMPC.m:41-43): run constant (depends on r, Np, Nc) */
double mko_sigma_min_gtg(int Np, int Nc, double r) {
  double A[NC_MAX][NC_MAX];
  for (int j = 0; j < Nc; ++j)
    for (int l = 0; l < Nc; ++l) {
      double acc = 0.0;
      for (int i = 1; i <= Np; ++i) {
        double gj = (i - 1 - j >= 0) ? r * (double)(i - 1 - j) : 0.0;
        double gl = (i - 1 - l >= 0) ? r * (double)(i - 1 - l) : 0.0;
        acc = acc + gj * gl;
      }
      A[j][l] = acc;
    }
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < Nc; ++i) for (int j = i + 1; j < Nc; ++j) off += A[i][j] * A[i][j];
    if (off == 0.0) break;
    for (int pI = 0; pI < Nc - 1; ++pI)
      for (int q = pI + 1; q < Nc; ++q) {
        double apq = A[pI][q];
        if (apq == 0.0) continue;
        double theta = (A[q][q] - A[pI][pI]) / (2.0 * apq);
        double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
        if (theta < 0.0) t = -t;
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < Nc; ++k) { double akp = A[k][pI], akq = A[k][q]; A[k][pI] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
        for (int k = 0; k < Nc; ++k) { double apk = A[pI][k], aqk = A[q][k]; A[pI][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
      }
  }
  double mn = fabs(A[0][0]);
  for (int i = 1; i < Nc; ++i) if (fabs(A[i][i]) < mn) mn = fabs(A[i][i]);
  return mn;
}

typedef struct {
  double u, phise;
  int nexec;
  uint32_t active[4];
} mpc_out;

/* iterMPC.m + constraintsMPC.m + hildreth.m for one cell */
static void mpc_step(const mko_bank *bk, const mko_config *cf, double sigmin, double *st, const ekf_out *e, double uk_applied,
                     mpc_out *o) {
  const int Np = cf->Np, Nc = cf->Nc;
  const double *p = bk->params;
  OPC_BLOCK(5);
  const double uk_1 = st[S_UK];                                    /* iterMPC.m:19 */
  const double r = -bk->Ts / (3600.0 * p[P_Q]);                    /* EKFmatsHandler.m:43 */
  const double zero5[N5] = {0, 0, 0, 0, 0};
  double hs[NP_MAX + 1], hv[NP_MAX + 1], he[NP_MAX + 1], pxs[NP_MAX], pxv[NP_MAX], pxe[NP_MAX];
  /* runMPC.m:94-95 (Q24): logged phise uses the applied uk */
  {
    double acc = e->Cphi[0] * e->xhat[0];
    for (int j = 1; j < N5; ++j) acc = fma(e->Cphi[j], e->xhat[j], acc);
    acc = acc + 0.0 * e->xhat[5];
    o->phise = (acc + uk_applied * e->Dphi) + e->bphi;
  }
  pred_row(e->a, zero5, r, 0.0, e->xhat, Np, hs, pxs);             /* iterMPC.m:33 */
  /* F = -2*G'*(Ref - Phi*dx)  (:38) */
  double err[NP_MAX], F[NC_MAX], GtG[NC_MAX * NC_MAX], E[NC_MAX * NC_MAX];
  for (int i = 1; i <= Np; ++i) err[i - 1] = cf->ref - (pxs[i - 1] + hs[i] * uk_1);
  for (int j = 0; j < Nc; ++j) {
    double acc = 0.0;
    int first = 1;
    for (int i = 1; i <= Np; ++i) {
      double g = (i - 1 - j >= 0) ? hs[i - 1 - j] : 0.0;          /* G(i,j+1) = H{i-j} */
      double t = g * err[i - 1];
      acc = first ? t : acc + t;
      first = 0;
    }
    F[j] = -2.0 * acc;
  }
  for (int j = 0; j < Nc; ++j)
    for (int l = 0; l < Nc; ++l) {
      double acc = 0.0;
      int first = 1;
      for (int i = 1; i <= Np; ++i) {
        double gj = (i - 1 - j >= 0) ? hs[i - 1 - j] : 0.0, gl = (i - 1 - l >= 0) ? hs[i - 1 - l] : 0.0;
        double t = gj * gl;
        acc = first ? t : acc + t;
        first = 0;
      }
      GtG[j * Nc + l] = acc;
    }
  double nf = F[0] * F[0];
  for (int j = 1; j < Nc; ++j) nf = nf + F[j] * F[j];
  const double Ru = sqrt(nf) / ((2.0 * st[S_DUMAX]) * sqrt((double)Nc)) - sigmin; /* :43 (Q10) */
  for (int j = 0; j < Nc; ++j)
    for (int l = 0; l < Nc; ++l) E[j * Nc + l] = 2.0 * (GtG[j * Nc + l] + ((j == l) ? Ru : 0.0)); /* :47 */
  double LU[NC_MAX * NC_MAX], DU[NC_MAX], x[NC_MAX];
  spd_factor(Nc, E, LU);
  spd_solve(Nc, LU, F, x);
  for (int j = 0; j < Nc; ++j) DU[j] = -x[j];                      /* :48 */
  /* ---- constraintsMPC.m ---- */
  OPC_BLOCK(6);
  double Mm[NCON_MAX * NC_MAX], gam[NCON_MAX];
  int nC = 0;
  if (cf->useCurrent) {                                            /* :23-42 */
    for (int s = 0; s < 2; ++s)
      for (int i = 0; i < Nc; ++i, ++nC) {
        for (int j = 0; j < Nc; ++j) Mm[nC * Nc + j] = (j <= i) ? (s ? -1.0 : 1.0) : (s ? -0.0 : 0.0);
        gam[nC] = s ? -(st[S_UMIN] - uk_1) : (st[S_UMAX] - uk_1);
      }
    for (int s = 0; s < 2; ++s)
      for (int i = 0; i < Nc; ++i, ++nC) {
        for (int j = 0; j < Nc; ++j) Mm[nC * Nc + j] = (j == i) ? (s ? -1.0 : 1.0) : (s ? -0.0 : 0.0);
        gam[nC] = s ? -st[S_DUMIN] : st[S_DUMAX];
      }
  }
  if (cf->useVoltage) {                                            /* :45-63 */
    pred_row(e->a, e->Cv, 0.0, e->Dv, e->xhat, Np, hv, pxv);
    for (int i = 1; i <= Np; ++i, ++nC) {
      for (int j = 0; j < Nc; ++j) Mm[nC * Nc + j] = (i - 1 - j >= 0) ? hv[i - 1 - j] : 0.0;
      gam[nC] = st[S_VMAX] - ((pxv[i - 1] + hv[i] * uk_1) + e->bv);
    }
  }
  if (cf->useEta) {                                                /* :66-85 */
    pred_row(e->a, e->Cphi, 0.0, e->Dphi, e->xhat, Np, he, pxe);
    for (int i = 1; i <= Np; ++i, ++nC) {
      for (int j = 0; j < Nc; ++j) Mm[nC * Nc + j] = (i - 1 - j >= 0) ? -he[i - 1 - j] : -0.0;
      gam[nC] = -st[S_PHIMIN] + ((pxe[i - 1] + he[i] * uk_1) + e->bphi);
    }
  }
  if (cf->useZmax) {                                               /* :89-101 */
    for (int i = 1; i <= Np; ++i, ++nC) {
      for (int j = 0; j < Nc; ++j) Mm[nC * Nc + j] = (i - 1 - j >= 0) ? hs[i - 1 - j] : 0.0;
      gam[nC] = st[S_ZMAX] - ((pxs[i - 1] + hs[i] * uk_1) + e->soc);
    }
  }
  if (cf->z_target_terminal) {                                     /* :104-112: -G_s(end,:)*DU <= -z_max + Phi_s(end,:)*x_aug (no SOC_k baseline) */
    for (int j = 0; j < Nc; ++j) Mm[nC * Nc + j] = (Np - 1 - j >= 0) ? -hs[Np - 1 - j] : -0.0;
    gam[nC] = -(st[S_ZMAX] - cf->z_tol) + (pxs[Np - 1] + hs[Np] * uk_1);
    ++nC;
  }
  OPC_BLOCK(5);
  int viol = 0;                                                    /* iterMPC.m:66 (Q8) */
  for (int i = 0; i < nC; ++i) {
    double acc = Mm[i * Nc] * DU[0];
    for (int j = 1; j < Nc; ++j) acc = fma(Mm[i * Nc + j], DU[j], acc);
    if (acc - gam[i] > 0.0) viol = 1;
  }
  o->nexec = 0;
  for (int w = 0; w < 4; ++w) o->active[w] = 0;
  if (viol) {
    double lam[NCON_MAX];
    int nexec = 0;
    mko_hildreth(Nc, nC, E, F, Mm, gam, (st[S_LAMVALID] != 0.0) ? st + S_LAMBDA : NULL, cf->maxHild, DU, lam, &nexec); /* :67 */
    for (int i = 0; i < nC; ++i) {
      st[S_LAMBDA + i] = lam[i];                                   /* :68 (Q11) */
      if (lam[i] > 0.0) o->active[i >> 5] |= (1u << (i & 31));
    }
    OPC_BLOCK(5);
    st[S_LAMVALID] = 1.0;
    o->nexec = nexec;
  }
  const double uk = DU[0] + uk_1;                                  /* :75-76 */
  st[S_UK] = uk;                                                   /* :78 */
  o->u = uk;
  if (!(fabs(uk) <= 1.79769313486231570815e308)) st[S_STATUS] = (double)((unsigned)st[S_STATUS] | ST_NONFINITE);
}

/* Measurement-noise sample for (seed, cell, instant): twelve splitmix64 outputs as uniforms on [0,1) with 53 bits, summed in
   order, minus 6 (Irwin-Hall: mean 0, variance 1).  Integer arithmetic and IEEE additions only, so every implementation gives
   the same bits. */
double mko_noise(uint64_t seed, uint64_t cell, uint64_t k) {
  uint64_t x = seed ^ (cell * 0x9E3779B97F4A7C15ull) ^ ((k + 1ull) * 0xD1B54A32D192ED03ull);
  double s = 0.0;
  for (int i = 0; i < 12; ++i) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    s = s + (double)(z >> 11) * 0x1.0p-53;
  }
  return s - 6.0;
}

/* per-cell uniform on [0,1) (one splitmix64 step of a key of (seed, cell)) and the plant's capacity factor of a cell */
double mko_cell_uniform(uint64_t seed, uint64_t cell) {
  uint64_t z = (seed ^ (cell * 0x9E3779B97F4A7C15ull) ^ 0xA5A5A5A5A5A5A5A5ull) + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * 0x1.0p-53;
}
static double mko_plant_q_factor(const mko_config *cf, int64_t cell) {
  const double scale = cf->plant_q_scale > 0.0 ? cf->plant_q_scale : 1.0;
  if (cf->plant_q_spread == 0.0) return scale;
  return scale * (1.0 + cf->plant_q_spread * (2.0 * mko_cell_uniform(cf->noise_seed, (uint64_t)(cf->first_cell + cell)) - 1.0));
}

/* ------------------------------------------------------------------------------------------ */
/* cell init: initKF.m:37-133 + initMPC.m:32-72 + OB_step.m:63-71 and the first I=0 call       */
/* (runMPC.m:72-74, Q20)                                                                      */
/* ------------------------------------------------------------------------------------------ */
void mko_cell_init(const mko_bank *bk, const mko_config *cf, double *st, double SOC0pct, double ekfSOC0pct, double Tc,
                   const double *SigmaX0 /* 6x6 column-major */, const double *limits /* u_max,u_min,du_max,du_min,v_max,phise_min,z_max(+tol) */) {
  const double *p = bk->params;
  const int M = bk->nT * bk->nZ;
  (void)cf;
  memset(st, 0, sizeof(double) * (size_t)mko_state_len(M));
  const double Tk1 = Tc + 273.15;
  st[S_TC] = Tc;
  st[S_SOC0N] = soc_e(p, P_NEG, SOC0pct / 100.0);                  /* OB_step.m:64-71 */
  st[S_SOC0P] = soc_e(p, P_POS, SOC0pct / 100.0);
  st[S_SOCN] = st[S_SOC0N];
  st[S_SOCP] = st[S_SOC0P];
  /* OB_step.m:212-219: Cdleff = Cdl^(2-nDL) * wDL^(nDL-1), a run constant */
  st[S_CDLN] = mko_exp((2.0 - p[P_NEG + E_NDL]) * mko_log(p[P_NEG + E_CDL])) * mko_exp((p[P_NEG + E_NDL] - 1.0) * mko_log(p[P_NEG + E_WDL]));
  st[S_CDLP] = mko_exp((2.0 - p[P_POS + E_NDL]) * mko_log(p[P_POS + E_CDL])) * mko_exp((p[P_POS + E_NDL] - 1.0) * mko_log(p[P_POS + E_WDL]));
  st[S_ARRK0N] = arrhenius(p, Tk1, p[P_NEG + E_EA_K0]);
  st[S_ARRK0P] = arrhenius(p, Tk1, p[P_POS + E_EA_K0]);
  st[S_ARRRFN] = arrhenius(p, Tk1, -p[P_NEG + E_EA_RF]);
  st[S_ARRRFP] = arrhenius(p, Tk1, -p[P_POS + E_EA_RF]);
  st[S_X0] = 0.0;                                                  /* initKF.m:98-99 */
  st[S_SIGX0] = SigmaX0[5 * 6 + 5];
  st[S_PRIORI] = 0.0;                                              /* :121 */
  st[S_EKFSOC0] = ekfSOC0pct / 100.0;                              /* :133 */
  double *Sig = st + st_off_sig(M);
  for (int m = 0; m < M; ++m) {                                    /* :70,94-95 */
    int q = 0;
    for (int i = 0; i < N5; ++i)
      for (int j = i; j < N5; ++j, ++q) Sig[NSYM * m + q] = SigmaX0[j * 6 + i];
  }
  {                                                                /* initKF.m:100-101 (used by the MB method only) */
    double *xm = st + st_off_mbx(M), *Sm = st + st_off_mbs(M);
    for (int j = 0; j < N6; ++j) xm[j] = 0.0;
    for (int i = 0; i < N6; ++i)
      for (int j = i; j < N6; ++j) Sm[sym6(i, j)] = SigmaX0[j * 6 + i];
  }
  st[S_UK] = 0.0;                                                  /* initMPC.m:51, runMPC.m:72 */
  for (int k = 0; k < 7; ++k) st[S_UMAX + k] = limits[k];
  plant_step(bk, st, 0.0, 1.0, 1.0);                               /* runMPC.m:74 (Q20); I = 0: the mismatch factors do not enter */
}

/* one pass of the loop body runMPC.m:83-112 */
void mko_cell_step(const mko_bank *bk, const mko_config *cf, double sigmin, double *st, double *log_u, double *log_V, double *log_soc,
                   double *log_phise, int *log_iT, int *log_iZ, int *log_nexec, uint32_t *log_active, int64_t cell, int k) {
  const double uk = st[S_UK];
  if ((unsigned)st[S_STATUS] & ST_LOST) {                          /* Q19: a lost cell idles, outputs NaN */
    *log_u = NAN; *log_V = NAN; *log_soc = NAN; *log_phise = NAN; *log_iT = 0; *log_iZ = 0; *log_nexec = 0;
    for (int w = 0; w < 4; ++w) log_active[w] = 0;
    return;
  }
  const double V = plant_step(bk, st, uk, mko_plant_q_factor(cf, cell), cf->plant_rc_scale > 0.0 ? cf->plant_rc_scale : 1.0);   /* :85 */
  ekf_out e;
  double vk = V;                                                   /* :88 feeds the plant voltage itself */
  if (cf->meas_sigma != 0.0) vk = V + cf->meas_sigma * mko_noise(cf->noise_seed, (uint64_t)(cf->first_cell + cell), (uint64_t)k);
  if (cf->ekf_method == 1) ekf_step_mb(bk, cf, st, vk, uk, &e, NULL, NULL);
  else ekf_step(bk, cf, st, vk, uk, &e, NULL);                     /* :88,92 */
  if (e.lost) {
    *log_u = NAN; *log_V = NAN; *log_soc = NAN; *log_phise = NAN; *log_iT = 0; *log_iZ = 0; *log_nexec = 0;
    for (int w = 0; w < 4; ++w) log_active[w] = 0;
    return;
  }
  mpc_out o;
  mpc_step(bk, cf, sigmin, st, &e, uk, &o);                        /* :94-103 */
  OPC_BLOCK(9);
  *log_u = o.u; *log_V = V; *log_soc = e.soc; *log_phise = o.phise; /* :106-109 */
  *log_iT = e.iT + 1; *log_iZ = e.iZ + 1; *log_nexec = o.nexec;
  for (int w = 0; w < 4; ++w) log_active[w] = o.active[w];
}

/* run B cells for steps [k0,k1); state [B][state_len] AoS; logs [step][cell] (may be NULL) */
int mko_run(const mko_bank *bk, const mko_config *cf, int B, double *state, int k0, int k1, double *u, double *V, double *soc,
            double *phise, int32_t *iT, int32_t *iZ, int32_t *nexec, uint32_t *active /* [step][cell][4] */, int nthreads) {
  const int M = bk->nT * bk->nZ, SL = mko_state_len(M);
  const double r = -bk->Ts / (3600.0 * bk->params[P_Q]);
  const double sigmin = mko_sigma_min_gtg(cf->Np, cf->Nc, r);
  if (cf->Np > NP_MAX || cf->Nc > NC_MAX || bk->nz > 64) return -1;
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int c = 0; c < B; ++c) {
    double *st = state + (size_t)c * SL;
    for (int k = k0; k < k1; ++k) {
      double lu, lV, ls, lp;
      int liT, liZ, ln;
      uint32_t la[4];
      mko_cell_step(bk, cf, sigmin, st, &lu, &lV, &ls, &lp, &liT, &liZ, &ln, la, (int64_t)c, k);
      if (u) {
        size_t o = (size_t)(k - k0) * B + c;
        u[o] = lu; V[o] = lV; soc[o] = ls; phise[o] = lp; iT[o] = liT; iZ[o] = liZ; nexec[o] = ln;
        for (int w = 0; w < 4; ++w) active[o * 4 + w] = la[w];
      }
    }
  }
  return 0;
}

/* function-level entries used by tests */
int mko_test_psd(double *S15, int force) { return psd_repair(S15, force); }
void mko_test_two_nearest(double v, const double *pts, int n, int *i1, int *i2) { two_nearest(v, pts, n, i1, i2); }

double mko_ob_step(const mko_bank *bk, double *st, double Iapp) { return plant_step(bk, st, Iapp, 1.0, 1.0); }
/* the same for cell `cell` of a run with the plant-model mismatch scenario of cf */
double mko_ob_step_s(const mko_bank *bk, const mko_config *cf, double *st, double Iapp, int64_t cell) {
  return plant_step(bk, st, Iapp, mko_plant_q_factor(cf, cell), cf->plant_rc_scale > 0.0 ? cf->plant_rc_scale : 1.0);
}

/* iterEKF + EKFmatsHandler; mats = [lost,iT,iZ, xhat(6), a(5), Cv(5), Dv, Cphi(5), Dphi, bv, bphi, vhat, soc] (31) */
void mko_iter_ekf(const mko_bank *bk, const mko_config *cf, double *st, double vk, double ik, double *zk, double *mats) {
  ekf_out e;
  memset(&e, 0, sizeof e);
  ekf_step(bk, cf, st, vk, ik, &e, zk);
  int q = 0;
  mats[q++] = e.lost; mats[q++] = e.iT + 1; mats[q++] = e.iZ + 1;
  for (int j = 0; j < 6; ++j) mats[q++] = e.xhat[j];
  for (int j = 0; j < 5; ++j) mats[q++] = e.a[j];
  for (int j = 0; j < 5; ++j) mats[q++] = e.Cv[j];
  mats[q++] = e.Dv;
  for (int j = 0; j < 5; ++j) mats[q++] = e.Cphi[j];
  mats[q++] = e.Dphi; mats[q++] = e.bv; mats[q++] = e.bphi; mats[q++] = e.vhat; mats[q++] = e.soc;
  mats[q++] = 0.0; /* reserved */
}

/* the same call with the model-blend filter ('MdlB'); bounds (nz+2) optional */
void mko_iter_ekf_mb(const mko_bank *bk, const mko_config *cf, double *st, double vk, double ik, double *zk, double *bounds, double *mats) {
  ekf_out e;
  memset(&e, 0, sizeof e);
  ekf_step_mb(bk, cf, st, vk, ik, &e, zk, bounds);
  int q = 0;
  mats[q++] = e.lost; mats[q++] = e.iT + 1; mats[q++] = e.iZ + 1;
  for (int j = 0; j < 6; ++j) mats[q++] = e.xhat[j];
  for (int j = 0; j < 5; ++j) mats[q++] = e.a[j];
  for (int j = 0; j < 5; ++j) mats[q++] = e.Cv[j];
  mats[q++] = e.Dv;
  for (int j = 0; j < 5; ++j) mats[q++] = e.Cphi[j];
  mats[q++] = e.Dphi; mats[q++] = e.bv; mats[q++] = e.bphi; mats[q++] = e.vhat; mats[q++] = e.soc;
  mats[q++] = 0.0;
}

/* predMat for a scalar output (returns h[0..Np], Phi*x part) */
void mko_pred_row(const double *a5, const double *c6, double d, const double *x6, int Np, double *h, double *phix) {
  pred_row(a5, c6, c6[5], d, x6, Np, h, phix);
}
