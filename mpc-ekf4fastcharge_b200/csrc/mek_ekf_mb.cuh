// Model-blend EKF (iterEKF.m with ekfData.method = 'MB': initKF.m:46-47,100-101; iterEKF.m:90-102,125-128,160-176,
// 199-203 and the MB branches of getVariables / getChatV / getChatZ) + EKFmatsHandler, one thread per cell.
//
// runMPC.m:16 selects the output-blend filter, so the fused kernel never runs this; it backs the function-level entry
// mek_pop_iter_ekf of a population created with mek_config.ekf_method = 1 (initKF(..., 'MdlB', ...)).  One blended model per
// cell: xhat (n+1, integrator last) and the full (n+1)x(n+1) covariance (upper triangle, 21 doubles) live behind the
// per-model rows of the state block.  Works on the full bank in global memory (all nz output rows), operations in the
// order of oracle/mek_oracle.c::ekf_step_mb.
#pragma once
#include "mek_kernels.cuh"

constexpr int N6 = 6, NSYM6 = 21;

struct RowLists {            // output-row kinds for the full-zk corrections and the 3-sigma bounds
  int n_negtss, n_postss, n_thetae, n_negphise, n_posphise, n_phie, n_posphis, r_phise0;
  int negtss[64], postss[64], thetae[64], negphise[64], posphise[64], phie[64], posphis[64];   // nz <= 64 (mek_bank_create): no group can overflow
  int r_ifdl0, r_ifdl3, r_if0, r_if3, r_tss0, r_tss3, r_the0, r_the3, r_phie_last, r_phi;   // rows in the full output numbering
  unsigned char kind[64];    // per output row: 0 plain, 1 posPhis, 2 negThetass, 3 posThetass, 4 negPhise, 5 posPhise, 6 Phie
};

__device__ __forceinline__ int sym6(int i, int j) {
  const int a = i <= j ? i : j, b = i <= j ? j : i;
  return a * 6 - (a * (a - 1)) / 2 + (b - a);
}

// PSD repair (iterEKF.m:164-166) of the 6x6 covariance: LDL^T positive-definiteness check, else cyclic Jacobi V|w|V'
__device__ __noinline__ int psd_repair6(double* S, int force) {
  if (!force) {
    double L[6][6], dd[6];
    bool pd = true;
    for (int j = 0; j < 6 && pd; ++j) {
      double dj = S[sym6(j, j)];
      for (int k = 0; k < j; ++k) dj = fma(-(L[j][k] * L[j][k]), dd[k], dj);
      if (!(dj > 0.0)) { pd = false; break; }
      dd[j] = dj;
      const double idj = rcp_d(dj);
      for (int i = j + 1; i < 6; ++i) {
        double v = S[sym6(i, j)];
        for (int k = 0; k < j; ++k) v = fma(-(L[i][k] * L[j][k]), dd[k], v);
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
        const double apq = A[pI][q];
        if (apq == 0.0) continue;
        const double theta = (A[q][q] - A[pI][pI]) / (2.0 * apq);
        double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
        if (theta < 0.0) t = -t;
        const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
        for (int k = 0; k < 6; ++k) {
          const double akp = A[k][pI], akq = A[k][q];
          A[k][pI] = cs * akp - sn * akq;
          A[k][q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < 6; ++k) {
          const double apk = A[pI][k], aqk = A[q][k];
          A[pI][k] = cs * apk - sn * aqk;
          A[q][k] = sn * apk + cs * aqk;
        }
        for (int k = 0; k < 6; ++k) {
          const double vkp = V[k][pI], vkq = V[k][q];
          V[k][pI] = cs * vkp - sn * vkq;
          V[k][q] = sn * vkp + cs * vkq;
        }
      }
  }
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j) {
      double hh = 0.0;
      for (int k = 0; k < 6; ++k) hh = hh + (V[i][k] * fabs(A[k][k])) * V[j][k];
      const double sv = S[sym6(i, j)];
      S[sym6(i, j)] = (((sv + sv) + hh) + hh) / 4.0;                 // :166
    }
  return 1;
}

// getVariables (iterEKF.m:259-417) on the full output vector with the MB state (every neighbour reads xhat(1:n), :310-312).
// Z[nz] gets the corrected outputs the voltage equation and EKFmatsHandler read (Phie rows and posPhis left to the caller,
// as in the oracle); returns Vcell.
__device__ double mb_get_variables(const Params& P, const RowLists& RL, const Cell& c, const double* x, double ik, double T, Scratch sc,
                                   double* Z) {
  const Dims& d = P.d;
  const double* p = P.cp;
  const double rpa = d.Ts / (3600.0 * p[P_Q]);
  const double xSOC = c.f(S_EKFSOC0) - x[5] * rpa;                    // :277
  double SOCn = soc_e(p, P_NEG, xSOC), SOCp = soc_e(p, P_POS, xSOC);  // :282-283
  double warn = 0.0;
  if (SOCn < 0.0) { warn += 1.0; SOCn = 1e-6; }                       // :284-295
  if (SOCn > 1.0) { warn += 1.0; SOCn = 1.0 - 1e-6; }
  if (SOCp < 0.0) { warn += 1.0; SOCp = 1e-6; }
  if (SOCp > 0.998) { warn += 1.0; SOCp = 0.998; }
  int mk[4];
  double g[4];
  for (int k = 0; k < 4; ++k) { mk[k] = (int)sc[SC_M + k]; g[k] = sc[SC_GAMMA + k]; }
  for (int r = 0; r < d.nz; ++r) {                                    // :302-313
    double yc[4];
    for (int k = 0; k < 4; ++k) {
      const double* cr = P.Cfull + ((size_t)mk[k] * d.nz + r) * 6;
      double acc = cr[0] * x[0];
      for (int j = 1; j < N5; ++j) acc = fma(cr[j], x[j], acc);
      yc[k] = fma(P.Dfull[(size_t)mk[k] * d.nz + r], ik, acc);
    }
    Z[r] = blend4(yc, g);
  }
  const double If0 = Z[RL.r_if0], If3 = Z[RL.r_if3];                  // :324-325
  bool any;
  for (int k = 0; k < RL.n_negtss; ++k) Z[RL.negtss[k]] = Z[RL.negtss[k]] + SOCn;   // :331
  any = false;
  for (int k = 0; k < RL.n_negtss; ++k) if (Z[RL.negtss[k]] < 0.0) { any = true; Z[RL.negtss[k]] = 1e-6; }
  if (any) warn += 1.0;
  any = false;
  for (int k = 0; k < RL.n_negtss; ++k) if (Z[RL.negtss[k]] > 1.0) { any = true; Z[RL.negtss[k]] = 1.0 - 1e-6; }
  if (any) warn += 1.0;
  for (int k = 0; k < RL.n_postss; ++k) Z[RL.postss[k]] = Z[RL.postss[k]] + SOCp;   // :345
  any = false;
  for (int k = 0; k < RL.n_postss; ++k) if (Z[RL.postss[k]] < 0.0) { any = true; Z[RL.postss[k]] = 1e-6; }
  if (any) warn += 1.0;
  any = false;
  for (int k = 0; k < RL.n_postss; ++k) if (Z[RL.postss[k]] > 0.998) { any = true; Z[RL.postss[k]] = 0.998; }
  if (any) warn += 1.0;
  const double UnAvg = uocp_n(p, SOCn), UpAvg = uocp_p(p, SOCp);      // :362-367
  for (int k = 0; k < RL.n_negphise; ++k) Z[RL.negphise[k]] = Z[RL.negphise[k]] + UnAvg;
  for (int k = 0; k < RL.n_posphise; ++k) Z[RL.posphise[k]] = Z[RL.posphise[k]] + UpAvg;
  const double PhieTilde3 = Z[RL.r_phie_last];                        // :371
  for (int k = 0; k < RL.n_thetae; ++k) Z[RL.thetae[k]] = Z[RL.thetae[k]] + 1.0;    // :383
  any = false;
  for (int k = 0; k < RL.n_thetae; ++k) if (Z[RL.thetae[k]] < 0.0) { any = true; Z[RL.thetae[k]] = 1e-6; }
  if (any) { warn += 1.0; c.f(S_STATUS) = (double)((unsigned)c.f(S_STATUS) | ST_Q18); }   // Q18
  if (warn != 0.0) c.f(S_WARN) += warn;
  const double th0 = Z[RL.r_tss0], th3 = Z[RL.r_tss3];
  const double k0n = k0_e(p, P_NEG, SOCn, c.f(S_ARRK0N)), k0p = k0_e(p, P_POS, SOCp, c.f(S_ARRK0P));   // :392-393
  const double i0n = k0n * sqrt((Z[RL.r_the0] * (1.0 - th0)) * th0);
  const double i0p = k0p * sqrt((Z[RL.r_the3] * (1.0 - th3)) * th3);
  const double pref = ((2.0 * p[P_R]) * T) / p[P_F];
  const double negEta0 = pref * asinh_d(If0 / (2.0 * i0n));           // :400-401
  const double posEta3 = pref * asinh_d(If3 / (2.0 * i0p));
  const double Un0 = uocp_n(p, th0), Up3 = uocp_p(p, th3);            // :404-407
  const double Rfn = rf_e(p, P_NEG, SOCn, c.f(S_ARRRFN)), Rfp = rf_e(p, P_POS, SOCp, c.f(S_ARRRFP));
  return ((((posEta3 - negEta0) + PhieTilde3) + Up3) - Un0) + (Rfp * Z[RL.r_ifdl3] - Rfn * Z[RL.r_ifdl0]);   // :409-410 (Q5)
}

// one iteration of the model-blend filter + EKFmatsHandler; zk / bound (nz+2 each) optional
__device__ void ekf_step_mb(const Params& P, const BankSmem& bk, const RowLists& RL, const Cell& c, double vk, double ik, EkfOut& o,
                            Scratch sc, double* zk_out, double* bound_out) {
  const Dims& d = P.d;
  const double* p = P.cp;
  const int nz = d.nz;
  o.lost = 0;
  if (c.f(S_WARN) > 10.0) {                                           // iterEKF.m:55-59 (Q19)
    c.f(S_STATUS) = (double)((unsigned)c.f(S_STATUS) | ST_LOST);
    o.lost = 1;
    if (zk_out) for (int r = 0; r < nz + 2; ++r) zk_out[r] = NAN;
    if (bound_out) for (int r = 0; r < nz + 2; ++r) bound_out[r] = NAN;
    return;
  }
  double x[N6], S[NSYM6], Z[66];
  for (int j = 0; j < N6; ++j) x[j] = c.f(d.f_mbx + j);
  for (int q = 0; q < NSYM6; ++q) S[q] = c.f(d.f_mbs + q);
  const double Tk = c.f(S_TC) + 273.15;                               // :65
  const double prior = c.f(S_PRIORI), Sw = d.SigmaW;
  const double rpa = d.Ts / (3600.0 * p[P_Q]);
  double SOC = c.f(S_EKFSOC0) - x[5] * rpa;                           // :92
  get_xind(d, bk, Tk, SOC, sc, -1);                                   // :93
  int mk[4];
  double g[4];
  for (int k = 0; k < 4; ++k) { mk[k] = (int)sc[SC_M + k]; g[k] = sc[SC_GAMMA + k]; }
  double amb[N6];
  for (int j = 0; j < N5; ++j)                                        // :94-98 [A1,A2,A3,A4]*gamma
    amb[j] = fma(bk.mdl[mk[3] * d.mstride + j], g[3], fma(bk.mdl[mk[2] * d.mstride + j], g[2],
             fma(bk.mdl[mk[1] * d.mstride + j], g[1], bk.mdl[mk[0] * d.mstride + j] * g[0])));
  amb[5] = 1.0;
  for (int j = 0; j < N6; ++j) x[j] = fma(amb[j], x[j], prior);       // :99
  for (int i = 0; i < N6; ++i)
    for (int j = i; j < N6; ++j) S[sym6(i, j)] = fma(amb[i] * S[sym6(i, j)], amb[j], Sw);   // :101
  SOC = c.f(S_EKFSOC0) - x[5] * rpa;                                  // :102
  get_xind(d, bk, Tk, SOC, sc, -1);                                   // :107
  double vhat = mb_get_variables(P, RL, c, x, ik, Tk, sc, Z);         // :108
  for (int k = 0; k < 4; ++k) { mk[k] = (int)sc[SC_M + k]; g[k] = sc[SC_GAMMA + k]; }
  // gamma-weighted row sums sum([C1(r,:); C2(r,:); C3(r,:); C4(r,:)]) with Ck = gamma(k)*C (:423-426)
  auto rsum = [&](int r, int j) {
    return ((g[0] * P.Cfull[((size_t)mk[0] * nz + r) * 6 + j] + g[1] * P.Cfull[((size_t)mk[1] * nz + r) * 6 + j]) +
            g[2] * P.Cfull[((size_t)mk[2] * nz + r) * 6 + j]) +
           g[3] * P.Cfull[((size_t)mk[3] * nz + r) * 6 + j];
  };
  double Chat[N6];
  {                                                                   // getChatV, MB branches (:448-460,475-479,489-491,512-518)
    const double xSOC = c.f(S_EKFSOC0) - x[5] * rpa;
    const double sn = soc_e(p, P_NEG, xSOC), sp = soc_e(p, P_POS, xSOC);
    const double Rfn = rf_e(p, P_NEG, sn, c.f(S_ARRRFN)), Rfp = rf_e(p, P_POS, sp, c.f(S_ARRRFP));
    const double k0n = k0_e(p, P_NEG, sn, c.f(S_ARRK0N)), k0p = k0_e(p, P_POS, sp, c.f(S_ARRK0P));
    const double th0 = Z[RL.r_tss0], th3 = Z[RL.r_tss3];
    const double i0n = k0n * sqrt((Z[RL.r_the0] * (1.0 - th0)) * th0);
    const double i0p = k0p * sqrt((Z[RL.r_the3] * (1.0 - th3)) * th3);
    const double Rctn = (p[P_R] * Tk) / (p[P_F] * i0n), Rctp = (p[P_R] * Tk) / (p[P_F] * i0p);
    const double dUn0 = duocp_n(p, th0), dUp3 = duocp_p(p, th3);
    const double dsn = soc_e(p, P_NEG, 1.0) - soc_e(p, P_NEG, 0.0), dsp = soc_e(p, P_POS, 1.0) - soc_e(p, P_POS, 0.0);
    const double res0n = (((-dUn0) * d.Ts) * dsn) / (3600.0 * p[P_Q]);
    const double res0p = (((-dUp3) * d.Ts) * dsp) / (3600.0 * p[P_Q]);
    for (int j = 0; j < N5; ++j) {
      double v = Rfp * rsum(RL.r_ifdl3, j) - Rfn * rsum(RL.r_ifdl0, j);        // :456-459
      v = (v + Rctp * rsum(RL.r_if3, j)) - Rctn * rsum(RL.r_if0, j);            // :476-479
      v = v + rsum(RL.r_phie_last, j);                                          // :490-491
      v = (v + dUp3 * rsum(RL.r_tss3, j)) - dUn0 * rsum(RL.r_tss0, j);          // :513-516
      Chat[j] = v;
    }
    Chat[5] = res0p - res0n;                                          // :501, :517
  }
  double v6[N6], L[N6];                                               // gain (:126-127)
  for (int i = 0; i < N6; ++i) {
    double acc = S[sym6(i, 0)] * Chat[0];
    for (int j = 1; j < N6; ++j) acc = fma(S[sym6(i, j)], Chat[j], acc);
    v6[i] = acc;
  }
  double sq = Chat[0] * v6[0];
  for (int i = 1; i < N6; ++i) sq = fma(Chat[i], v6[i], sq);
  const double svt = sq + d.SigmaV;
  const double isv = rcp_d(svt);
  for (int i = 0; i < N6; ++i) L[i] = v6[i] * isv;
  const double resid = vk - vhat;                                     // :133
  for (int j = 0; j < N6; ++j) x[j] = fma(L[j], resid, x[j]);         // :161
  for (int i = 0; i < N6; ++i)
    for (int j = i; j < N6; ++j) S[sym6(i, j)] = fma(-(L[i] * svt), L[j], S[sym6(i, j)]);   // :163
  if (psd_repair6(S, d.force_jacobi)) c.f(S_NJAC) += 1.0;             // :164-166
  if (resid * resid > 9.0 * svt) {                                    // :169-172
    for (int k = 0; k < NSYM6; ++k) S[k] = S[k] * 2.0;
    c.f(S_NBUMP) += 1.0;
  }
  for (int j = 0; j < N6; ++j) c.f(d.f_mbx + j) = x[j];
  for (int q = 0; q < NSYM6; ++q) c.f(d.f_mbs + q) = S[q];
  SOC = c.f(S_EKFSOC0) - x[5] * rpa;                                  // :176
  get_xind(d, bk, Tk, SOC, sc, -1);                                   // :181
  vhat = mb_get_variables(P, RL, c, x, ik, Tk, sc, Z);                // :182
  for (int k = 0; k < 4; ++k) { mk[k] = (int)sc[SC_M + k]; g[k] = sc[SC_GAMMA + k]; }
  const double Zsoc = c.f(S_EKFSOC0) - x[5] * rpa;                    // :416
  c.f(S_PRIORI) = ik;                                                 // :210
  double zk[66];
  for (int r = 0; r < nz; ++r) zk[r] = Z[r];
  {
    const double Phise0 = Z[RL.r_phise0];
    for (int k = 0; k < RL.n_phie; ++k) zk[RL.phie[k]] = Z[RL.phie[k]] - Phise0;
    for (int k = 0; k < RL.n_posphis; ++k) zk[RL.posphis[k]] = Z[RL.posphis[k]] + vhat;
    zk[nz] = vhat;
    zk[nz + 1] = Zsoc;
  }
  if (zk_out) for (int r = 0; r < nz + 2; ++r) zk_out[r] = zk[r];
  if (bound_out) {                                                    // :199-205 with getChatZ's MB branches (:538,554-558,596-600)
    const double xSOC = c.f(S_EKFSOC0) - x[5] * rpa;
    const double sn = soc_e(p, P_NEG, xSOC), sp = soc_e(p, P_POS, xSOC);
    const double Rfn = rf_e(p, P_NEG, sn, c.f(S_ARRRFN)), Rfp = rf_e(p, P_POS, sp, c.f(S_ARRRFP));
    const double k0n = k0_e(p, P_NEG, sn, c.f(S_ARRK0N)), k0p = k0_e(p, P_POS, sp, c.f(S_ARRK0P));
    const double th0 = zk[RL.r_tss0], th3 = zk[RL.r_tss3];
    const double i0n = k0n * sqrt((zk[RL.r_the0] * (1.0 - th0)) * th0);
    const double i0p = k0p * sqrt((zk[RL.r_the3] * (1.0 - th3)) * th3);
    const double Rctn = (p[P_R] * Tk) / (p[P_F] * i0n), Rctp = (p[P_R] * Tk) / (p[P_F] * i0p);
    const double dUn0 = duocp_n(p, th0), dUp3 = duocp_p(p, th3);
    const double dsn = soc_e(p, P_NEG, 1.0) - soc_e(p, P_NEG, 0.0), dsp = soc_e(p, P_POS, 1.0) - soc_e(p, P_POS, 0.0);
    double CV[N6];
    for (int j = 0; j < N5; ++j) {
      double v = Rfp * rsum(RL.r_ifdl3, j) - Rfn * rsum(RL.r_ifdl0, j);
      v = (v + Rctp * rsum(RL.r_if3, j)) - Rctn * rsum(RL.r_if0, j);
      v = v + rsum(RL.r_phie_last, j);
      v = (v + dUp3 * rsum(RL.r_tss3, j)) - dUn0 * rsum(RL.r_tss0, j);
      CV[j] = v;
    }
    CV[5] = (((-dUp3) * d.Ts) * dsp) / (3600.0 * p[P_Q]) - (((-dUn0) * d.Ts) * dsn) / (3600.0 * p[P_Q]);
    const double res0n = ((-d.Ts) * dsn) / (3600.0 * p[P_Q]), res0p = ((-d.Ts) * dsp) / (3600.0 * p[P_Q]);   // :562-565
    const double dUn = duocp_n(p, sn), dUp = duocp_p(p, sp);          // :579-580
    for (int r = 0; r <= nz; ++r) {
      double row[N6];
      if (r < nz) {
        const int kind = RL.kind[r];
        double c0 = 0.0;
        if (kind == 1) c0 = CV[5];
        else if (kind == 2) c0 = res0n;
        else if (kind == 3) c0 = res0p;
        else if (kind == 4) c0 = dUn * res0n;
        else if (kind == 5) c0 = dUp * res0p;
        else if (kind == 6) c0 = -dUn * res0n;
        for (int j = 0; j < N5; ++j) {
          double v = rsum(r, j);
          if (kind == 1) v = v + CV[j];                               // :555-557
          if (kind == 6) v = v - rsum(RL.r_phise0, j);                // :597-599
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
      bound_out[r] = 3.0 * sqrt(t);                                   // :205
    }
    const double res = -d.Ts / (3600.0 * p[P_Q]);
    bound_out[nz + 1] = 3.0 * sqrt((res * S[sym6(5, 5)]) * res);      // :203
  }
  o.vhat = vhat;
  o.soc = Zsoc;
  // ---- EKFmatsHandler (EKFmatsHandler.m:26-114) on the full bank; M(iT,iZ).xhat is never written by the MB filter ----
  int imax = 0;                                                       // :26 first max, NaN ignored
  {
    bool have = false;
    double best = 0.0;
    for (int k = 0; k < 4; ++k) {
      const double gk = g[k];
      if (gk != gk) continue;
      if (!have || gk > best) { best = gk; imax = k; have = true; }
    }
  }
  const int mp = mk[imax];
  o.iT = mp / d.nZ;
  o.iZ = mp - o.iT * d.nZ;
  for (int j = 0; j < N5; ++j) { o.xhat[j] = c.xhat(mp, j); o.a[j] = bk.mdl[mp * d.mstride + j]; }   // :33-35
  o.xhat[5] = x[5];                                                   // :33 (Q3) ekfData.xhat(end): the MB integrator state
  {
    const double TK = c.f(S_TC) + 273.15;                             // :53
    const double sn = soc_e(p, P_NEG, Zsoc), sp = soc_e(p, P_POS, Zsoc);   // :56-58
    const double k0n = k0_e(p, P_NEG, sn, c.f(S_ARRK0N)), k0p = k0_e(p, P_POS, sp, c.f(S_ARRK0P));
    const double th0 = Z[RL.r_tss0], th3 = Z[RL.r_tss3];
    const double i0n = k0n * sqrt((Z[RL.r_the0] * (1.0 - th0)) * th0);     // :63-66
    const double i0p = k0p * sqrt((Z[RL.r_the3] * (1.0 - th3)) * th3);
    const double Rfn = rf_e(p, P_NEG, sn, c.f(S_ARRRFN)), Rfp = rf_e(p, P_POS, sp, c.f(S_ARRRFP));
    const double* Cm = P.Cfull + (size_t)mp * nz * 6;
    const double* Dm = P.Dfull + (size_t)mp * nz;
    for (int j = 0; j < N5; ++j)                                      // :75-76
      o.Cv[j] = (Rfp * Cm[RL.r_ifdl3 * 6 + j] - Rfn * Cm[RL.r_ifdl0 * 6 + j]) + Cm[RL.r_phie_last * 6 + j];
    o.Dv = (Rfp * Dm[RL.r_ifdl3] - Rfn * Dm[RL.r_ifdl0]) + Dm[RL.r_phie_last];   // :79-80
    const double Upos = uocp_p(p, th3), Uneg = uocp_n(p, th0);        // :84-85
    const double pref = ((2.0 * p[P_R]) * TK) / p[P_F];
    const double negEta0 = pref * asinh_d(Z[RL.r_if0] / (2.0 * i0n)); // :88-89
    const double posEta3 = pref * asinh_d(Z[RL.r_if3] / (2.0 * i0p));
    o.bv = ((Upos - Uneg) + (posEta3 - negEta0)) + 0.0;               // :92
    o.bphi = uocp_n(p, sn);                                           // :96
    for (int j = 0; j < N5; ++j) o.Cphi[j] = Cm[RL.r_phi * 6 + j];    // :97
    o.Dphi = Dm[RL.r_phi];                                            // :98
  }
}
