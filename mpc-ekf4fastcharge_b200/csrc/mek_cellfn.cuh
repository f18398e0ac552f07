// Closed forms behind ROM.cellData.function.* for the synthetic NMC30-like bank.
//
// In the reference these are MATLAB function handles stored in ROM_NMC30_HRA.mat (missing); call
// sites: OB_step.m:205-217,231-232,313-314,329-330,337-340; iterEKF.m:282-283,362-363,392-393,
// 404-407,439-442,463-464,495-500; EKFmatsHandler.m:57-69,84-85,96.  Here they are a parameter
// vector (layout = mek/bank.py P_*) evaluated in a fixed operation order.
#pragma once
#include "mek_math.cuh"

namespace mek {

enum : int { P_F = 0, P_R, P_Q, P_RC, P_TREF, P_NEG = 5, P_POS = 16, P_UN = 27, P_UP = 33, P_UP_COEF = 36, N_PARAMS = 44 };
enum : int { E_THETA0 = 0, E_THETA100, E_K0REF, E_K0SLOPE, E_EA_K0, E_RFREF, E_RFSLOPE, E_EA_RF, E_WDL, E_CDL, E_NDL };
constexpr int UP_MAXCOEF = 8;

// ---- tabulated cell functions (optional) ----
// The reference's cellData.function.* are MATLAB function handles inside ROM_NMC30_HRA.mat; a user who owns that file samples
// them on a uniform stoichiometry grid (matlab/mek_sample_celldata.m) and hands the tables to mek_bank_set_lut.  Eight tables
// of n points over [lo, hi]: Uocp, dUocp (both electrodes), k0 and Rf at the reference temperature (the Arrhenius factor
// of the parameter vector still applies).  Evaluation = MATLAB's interp1(...,'linear') with the argument clamped to the grid.
// The library appends four entries to the parameter vector it hands to the kernels: n (0 = closed forms), lo, 1/step and the
// table base address (bit pattern of a pointer).  Compiled out of the static-shape fused instances (MEK_LUT_OFF): a
// population on a tabulated bank runs the runtime-shape instances.
enum : int { P_LUT_N = N_PARAMS, P_LUT_LO, P_LUT_INVH, P_LUT_PTR, N_PARAMS_X };
enum : int { LUT_UN = 0, LUT_DUN, LUT_UP, LUT_DUP, LUT_K0N, LUT_K0P, LUT_RFN, LUT_RFP, LUT_NTAB };
#ifndef MEK_LUT_OFF
MEK_HDN double lut_eval(const double* p, int tab, double th) {
  if (th != th) return th;
  const int n = (int)p[P_LUT_N];
  const double* y = reinterpret_cast<const double*>((uintptr_t)double_to_bits(p[P_LUT_PTR])) + (size_t)tab * (size_t)n;
  double t = (th - p[P_LUT_LO]) * p[P_LUT_INVH];
  if (!(t > 0.0)) t = 0.0;                              // clamped to the grid
  if (t > (double)(n - 1)) t = (double)(n - 1);
  int i = (int)t;
  if (i > n - 2) i = n - 2;
  const double f = t - (double)i;
  return fma(f, y[i + 1] - y[i], y[i]);
}
#define MEK_IF_LUT(p, expr) if ((p)[P_LUT_N] != 0.0) return (expr)
#else
#define MEK_IF_LUT(p, expr)
#endif

// function.{neg,pos}.soc(z,T)
MEK_HD double soc_e(const double* p, int base, double z) {
  return p[base + E_THETA0] + z * (p[base + E_THETA100] - p[base + E_THETA0]);
}
// function.neg.Uocp / dUocp
// (inlined, like the positive-electrode pair below: the call sequence -- registers saved around it -- cost more than the body)
MEK_HD double uocp_n(const double* p, double th) {
  MEK_IF_LUT(p, lut_eval(p, LUT_UN, th));
  const double* c = p + P_UN;
  const double x = (th - c[4]) / c[5];
  const double t = th + c[2];
  return (c[0] + c[1] / (t * t)) + c[3] * (1.0 - x / sqrt(1.0 + x * x));
}
MEK_HD double duocp_n(const double* p, double th) {
  MEK_IF_LUT(p, lut_eval(p, LUT_DUN, th));
  const double* c = p + P_UN;
  const double x = (th - c[4]) / c[5];
  const double q = 1.0 + x * x;
  const double t = th + c[2];
  return (-2.0 * c[1]) / ((t * t) * t) - (c[3] / c[5]) / (q * sqrt(q));
}
// function.pos.Uocp / dUocp: Horner polynomial in x = (th - centre)/width
// (inlined: with the static-degree Horner form the body is shorter than a call)
MEK_HD double uocp_p(const double* p, double th) {
  MEK_IF_LUT(p, lut_eval(p, LUT_UP, th));
  const double x = (th - p[P_UP]) / p[P_UP + 1];
  const int nc = (int)p[P_UP + 2];
  if (nc == 7) {                                        // the canonical bank's degree: the same Horner steps with static coefficient
    double acc = p[P_UP_COEF + 6];                      // indices (constant-bank operands, no indexed loads, no loop)
#pragma unroll
    for (int i = 5; i >= 0; --i) acc = fma(acc, x, p[P_UP_COEF + i]);
    return acc;
  }
  double acc = p[P_UP_COEF + nc - 1];
  for (int i = nc - 2; i >= 0; --i) acc = fma(acc, x, p[P_UP_COEF + i]);
  return acc;
}
MEK_HD double duocp_p(const double* p, double th) {
  MEK_IF_LUT(p, lut_eval(p, LUT_DUP, th));
  const double x = (th - p[P_UP]) / p[P_UP + 1];
  const int nc = (int)p[P_UP + 2];
  if (nc == 7) {                                        // as in uocp_p
    double acc = 6.0 * p[P_UP_COEF + 6];
#pragma unroll
    for (int i = 5; i >= 1; --i) acc = fma(acc, x, (double)i * p[P_UP_COEF + i]);
    return acc / p[P_UP + 1];
  }
  double acc = (double)(nc - 1) * p[P_UP_COEF + nc - 1];
  for (int i = nc - 2; i >= 1; --i) acc = fma(acc, x, (double)i * p[P_UP_COEF + i]);
  return acc / p[P_UP + 1];
}
// ---- the two electrodes' OCP (and dOCP) side by side ----
// Same operations in the same order as uocp_n / uocp_p (duocp_n / duocp_p); on the device the quotients and the root are the
// speculative fast sequences of mek_math.cuh, so that the negative electrode's two chains and the positive electrode's Horner
// chain share one basic block, and anything they do not accept (or a tabulated bank, or a non-canonical polynomial degree)
// takes the standard evaluation.
static __device__ __noinline__ double2 ocp2_std(const double* p, double thn, double thp) { return make_double2(uocp_n(p, thn), uocp_p(p, thp)); }
static __device__ __noinline__ double2 docp2_std(const double* p, double thn, double thp) { return make_double2(duocp_n(p, thn), duocp_p(p, thp)); }
#ifdef MEK_SPEC_ON
__device__ __forceinline__ bool ocp_spec_ok(const double* p) {
#ifndef MEK_LUT_OFF
  if (p[P_LUT_N] != 0.0) return false;
#endif
  return (int)p[P_UP + 2] == 7;
}
__device__ __forceinline__ double2 ocp2(const double* p, double thn, double thp) {
  if (!ocp_spec_ok(p)) return ocp2_std(p, thn, thp);
  bool ok = true;
  const double* c = p + P_UN;
  const double x = div_spec(thn - c[4], c[5], ok);
  const double t = thn + c[2];
  const double a = div_spec(c[1], t * t, ok);
  const double b = div_spec(x, sqrt_spec(1.0 + x * x, ok), ok);
  const double un = (c[0] + a) + c[3] * (1.0 - b);
  const double xp = div_spec(thp - p[P_UP], p[P_UP + 1], ok);
  double acc = p[P_UP_COEF + 6];
#pragma unroll
  for (int i = 5; i >= 0; --i) acc = fma(acc, xp, p[P_UP_COEF + i]);
  if (!ok) return ocp2_std(p, thn, thp);
  return make_double2(un, acc);
}
__device__ __forceinline__ double2 docp2(const double* p, double thn, double thp) {
  if (!ocp_spec_ok(p)) return docp2_std(p, thn, thp);
  bool ok = true;
  const double* c = p + P_UN;
  const double x = div_spec(thn - c[4], c[5], ok);
  const double q = 1.0 + x * x;
  const double t = thn + c[2];
  const double a = div_spec(-2.0 * c[1], (t * t) * t, ok);
  const double b = div_spec(div_spec(c[3], c[5], ok), q * sqrt_spec(q, ok), ok);
  const double xp = div_spec(thp - p[P_UP], p[P_UP + 1], ok);
  double acc = 6.0 * p[P_UP_COEF + 6];
#pragma unroll
  for (int i = 5; i >= 1; --i) acc = fma(acc, xp, (double)i * p[P_UP_COEF + i]);
  const double dp = div_spec(acc, p[P_UP + 1], ok);
  if (!ok) return docp2_std(p, thn, thp);
  return make_double2(a - b, dp);
}
#else
MEK_HD double2 ocp2(const double* p, double thn, double thp) { return make_double2(uocp_n(p, thn), uocp_p(p, thp)); }
MEK_HD double2 docp2(const double* p, double thn, double thp) { return make_double2(duocp_n(p, thn), duocp_p(p, thp)); }
#endif

// function.{neg,pos}.k0(theta,T), .Rf(theta,T): Arrhenius factor hoisted (T is constant per cell)
MEK_HD double k0_e(const double* p, int base, double th, double arr) {
  MEK_IF_LUT(p, lut_eval(p, base == P_NEG ? LUT_K0N : LUT_K0P, th) * arr);
  return (p[base + E_K0REF] * (1.0 + p[base + E_K0SLOPE] * (th - 0.5))) * arr;
}
MEK_HD double rf_e(const double* p, int base, double th, double arr) {
  MEK_IF_LUT(p, lut_eval(p, base == P_NEG ? LUT_RFN : LUT_RFP, th) * arr);
  return (p[base + E_RFREF] * (1.0 + p[base + E_RFSLOPE] * th)) * arr;
}
MEK_HD double arrhenius(const double* p, double T, double Ea) {
  return exp_d((Ea / p[P_R]) * (1.0 / p[P_TREF] - 1.0 / T));
}
// OB_step.m:218-219  Cdleff = Cdl^(2-nDL) * wDL^(nDL-1)
MEK_HD double cdl_eff(const double* p, int base) {
  return exp_d((2.0 - p[base + E_NDL]) * log_d(p[base + E_CDL])) * exp_d((p[base + E_NDL] - 1.0) * log_d(p[base + E_WDL]));
}

}  // namespace mek
