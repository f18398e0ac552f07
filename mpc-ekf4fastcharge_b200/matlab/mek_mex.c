/*
 * mek_mex.c -- the single MEX gateway between MATLAB and libmek (include/mek.h).
 *
 *   out = mek_mex(cmd, args...)
 *
 * The .m shims next to this file (initKF.m, iterEKF.m, EKFmatsHandler.m, initMPC.m, iterMPC.m,
 * OB_step.m, hildreth.m, predMat.m, runMPC.m) keep the reference's argument shapes and call this
 * gateway; handles travel inside the MATLAB structs as uint64 scalars.  Batch size B is the number
 * of columns of the per-cell inputs, so B = 1 is the reference's single cell.
 *
 * Build (needs MATLAB, not available in the development container):
 *   mex -I<repo>/include mek_mex.c -L<repo>/mpc-ekf4fastcharge_b200/mek -lmek
 */
#ifdef MEK_MEX_STUB
#include "mex_stub.h"
#else
#include "mex.h"
#endif
#include <string.h>

#include "mek.h"

static void chk(int rc) {
  if (rc != MEK_OK) mexErrMsgIdAndTxt("mek:error", "%s", mek_last_error());
}
static void *get_handle(const mxArray *a) { return (void *)(uintptr_t)(*(uint64_t *)mxGetData(a)); }
static mxArray *put_handle(void *h) {
  mxArray *a = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
  *(uint64_t *)mxGetData(a) = (uint64_t)(uintptr_t)h;
  return a;
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
  char *cmd;
  (void)nlhs;
  if (nrhs < 1) mexErrMsgIdAndTxt("mek:usage", "mek_mex(cmd, ...)");
  cmd = mxArrayToString(prhs[0]);
  if (!strcmp(cmd, "bank_load")) {                       /* h = mek_mex('bank_load', path)        <- load(ROMfile,'ROM') */
    mek_bank *b = NULL;
    char *path = mxArrayToString(prhs[1]);
    chk(mek_bank_load(path, &b));
    mxFree(path);
    mexLock();
    plhs[0] = put_handle(b);
  } else if (!strcmp(cmd, "bank_set_lut")) {             /* mek_mex('bank_set_lut', bank, theta_lo, theta_hi, tables(8,n))  <- mek_sample_celldata.m */
    const mwSize n = mxGetN(prhs[4]);                    /* column-major 8 x n -> row-major [8][n] */
    double *t = (double *)mxMalloc(sizeof(double) * 8 * (n > 0 ? n : 1));
    const double *src = mxGetPr(prhs[4]);
    mwSize i, k;
    if (mxGetM(prhs[4]) != 8) mexErrMsgIdAndTxt("mek:usage", "tables must be 8 x n");
    for (k = 0; k < 8; ++k)
      for (i = 0; i < n; ++i) t[k * n + i] = src[i * 8 + k];
    chk(mek_bank_set_lut((mek_bank *)get_handle(prhs[1]), (int)n, mxGetScalar(prhs[2]), mxGetScalar(prhs[3]), t));
    mxFree(t);
  } else if (!strcmp(cmd, "pop_create")) {               /* h = mek_mex('pop_create', bank, cfgvec, B, device) */
    const double *c = mxGetPr(prhs[2]);                  /* [Np Nc maxHild c1 c2 c3 use_zmax ref SigmaV SigmaW (ekf_method (z_target_terminal z_tol (meas_sigma noise_seed (plant_q_scale plant_q_spread plant_rc_scale))))] */
    mek_config cfg;
    mek_pop *p = NULL;
    memset(&cfg, 0, sizeof cfg);
    cfg.Np = (int)c[0]; cfg.Nc = (int)c[1]; cfg.maxHild = (int)c[2];
    cfg.constraints[0] = (int)c[3]; cfg.constraints[1] = (int)c[4]; cfg.constraints[2] = (int)c[5];
    cfg.use_zmax = (int)c[6]; cfg.ref = c[7]; cfg.SigmaV = c[8]; cfg.SigmaW = c[9];
    if (mxGetNumberOfElements(prhs[2]) > 10) cfg.ekf_method = (int)c[10];   /* initKF's blend: 0 'OutB', 1 'MdlB' */
    if (mxGetNumberOfElements(prhs[2]) > 12) {                              /* mpcData.const.z_target_terminal, .z_tol (constraintsMPC.m:104-112) */
      cfg.z_target_terminal = c[11] != 0.0; cfg.z_tol = c[12];
    }
    if (mxGetNumberOfElements(prhs[2]) > 14) {                              /* noisy-sensor scenario (include/mek.h): sigma [V], seed */
      cfg.meas_sigma = c[13]; cfg.noise_seed = (uint64_t)c[14];
    }
    if (mxGetNumberOfElements(prhs[2]) > 17) {                              /* plant-model mismatch scenario: q scale, q spread, Rc scale */
      cfg.plant_q_scale = c[15]; cfg.plant_q_spread = c[16]; cfg.plant_rc_scale = c[17];
    }
    chk(mek_pop_create((mek_bank *)get_handle(prhs[1]), &cfg, (int64_t)mxGetScalar(prhs[3]), (int)mxGetScalar(prhs[4]), &p));
    plhs[0] = put_handle(p);
  } else if (!strcmp(cmd, "pop_init")) {                 /* mek_mex('pop_init', pop, SOC0, ekfSOC0, Tc, SigmaX0(6,6,B), limits(7,B)) */
    chk(mek_pop_init((mek_pop *)get_handle(prhs[1]), mxGetPr(prhs[2]), mxIsEmpty(prhs[3]) ? NULL : mxGetPr(prhs[3]), mxGetPr(prhs[4]),
                     mxGetPr(prhs[5]), mxGetPr(prhs[6])));
  } else if (!strcmp(cmd, "set_logging")) {              /* mek_mex('set_logging', pop, ncells, nsteps) */
    chk(mek_pop_set_logging((mek_pop *)get_handle(prhs[1]), (int64_t)mxGetScalar(prhs[2]), (int)mxGetScalar(prhs[3])));
  } else if (!strcmp(cmd, "run")) {                      /* mek_mex('run', pop, k0, k1)           <- runMPC.m:83-112 */
    mek_pop *p = (mek_pop *)get_handle(prhs[1]);
    chk(mek_run(p, (int)mxGetScalar(prhs[2]), (int)mxGetScalar(prhs[3])));
    chk(mek_sync(p));
  } else if (!strcmp(cmd, "read_logs")) {                /* [u,V,soc,phise,nexec] = mek_mex('read_logs', pop, k0, k1, ncells) */
    mek_pop *p = (mek_pop *)get_handle(prhs[1]);
    int k0 = (int)mxGetScalar(prhs[2]), k1 = (int)mxGetScalar(prhs[3]);
    mwSize nl = (mwSize)mxGetScalar(prhs[4]), n = (mwSize)(k1 - k0);
    int i;
    for (i = 0; i < 4; ++i) plhs[i] = mxCreateDoubleMatrix(nl, n, mxREAL);   /* column k = sample instant k (cells down the rows) */
    plhs[4] = mxCreateNumericMatrix(nl, n, mxINT32_CLASS, mxREAL);
    chk(mek_read_logs(p, k0, k1, mxGetPr(plhs[0]), mxGetPr(plhs[1]), mxGetPr(plhs[2]), mxGetPr(plhs[3]), NULL, NULL,
                      (int32_t *)mxGetData(plhs[4]), NULL));
  } else if (!strcmp(cmd, "ob_step")) {                  /* V = mek_mex('ob_step', pop, Iapp)     <- OB_step.m:1 */
    mek_pop *p = (mek_pop *)get_handle(prhs[1]);
    plhs[0] = mxCreateDoubleMatrix(1, mxGetNumberOfElements(prhs[2]), mxREAL);
    chk(mek_pop_ob_step(p, mxGetPr(prhs[2]), mxGetPr(plhs[0])));
  } else if (!strcmp(cmd, "iter_ekf")) {                 /* [zk,boundzk,mats] = mek_mex('iter_ekf', pop, vk, ik, nz)  <- iterEKF.m:30 + EKFmatsHandler.m:1 */
    mek_pop *p = (mek_pop *)get_handle(prhs[1]);
    mwSize B = mxGetNumberOfElements(prhs[2]), nz = (mwSize)mxGetScalar(prhs[4]);
    plhs[0] = mxCreateDoubleMatrix(nz + 2, B, mxREAL);
    plhs[1] = mxCreateDoubleMatrix(nz + 2, B, mxREAL);
    plhs[2] = mxCreateDoubleMatrix(MEK_NMATS, B, mxREAL);
    chk(mek_pop_iter_ekf(p, mxGetPr(prhs[2]), mxGetPr(prhs[3]), mxGetPr(plhs[0]), mxGetPr(plhs[1]), mxGetPr(plhs[2])));
  } else if (!strcmp(cmd, "iter_mpc")) {                 /* [uk,phise,nexec,cost] = mek_mex('iter_mpc', pop, mats, uk_applied)  <- iterMPC.m:1 */
    mek_pop *p = (mek_pop *)get_handle(prhs[1]);          /* cost(:,b) = [J_uncon; J_final; norm_DU; viol]  (iterMPC.m:88-95) */
    mwSize B = mxGetNumberOfElements(prhs[3]);
    plhs[0] = mxCreateDoubleMatrix(1, B, mxREAL);
    plhs[1] = mxCreateDoubleMatrix(1, B, mxREAL);
    plhs[2] = mxCreateNumericMatrix(1, B, mxINT32_CLASS, mxREAL);
    plhs[3] = mxCreateDoubleMatrix(MEK_NCOST, B, mxREAL);
    chk(mek_pop_iter_mpc_cost(p, mxGetPr(prhs[2]), mxGetPr(prhs[3]), mxGetPr(plhs[0]), mxGetPr(plhs[1]), (int32_t *)mxGetData(plhs[2]), NULL,
                              mxGetPr(plhs[3])));
  } else if (!strcmp(cmd, "cell_summaries")) {           /* S = mek_mex('cell_summaries', pop, B): 12 x B per-cell outcomes */
    mek_pop *p = (mek_pop *)get_handle(prhs[1]);
    plhs[0] = mxCreateDoubleMatrix(MEK_NCELLSUM, (mwSize)mxGetScalar(prhs[2]), mxREAL);
    chk(mek_read_cell_summaries(p, mxGetPr(plhs[0])));
  } else if (!strcmp(cmd, "hildreth")) {                 /* [DU,lambda,nexec] = mek_mex('hildreth', Et, F, Mt, gamma, lambda0, maxIter) */
    /* Et = E.' and Mt = M.' (column-major MATLAB -> row-major ABI), one problem */
    int Nc = (int)mxGetM(prhs[1]), nC = (int)mxGetN(prhs[3]);
    plhs[0] = mxCreateDoubleMatrix(Nc, 1, mxREAL);
    plhs[1] = mxCreateDoubleMatrix(nC, 1, mxREAL);
    plhs[2] = mxCreateNumericMatrix(1, 1, mxINT32_CLASS, mxREAL);
    chk(mek_hildreth(1, Nc, nC, mxGetPr(prhs[1]), mxGetPr(prhs[2]), mxGetPr(prhs[3]), mxGetPr(prhs[4]),
                     mxIsEmpty(prhs[5]) ? NULL : mxGetPr(prhs[5]), (int)mxGetScalar(prhs[6]), mxGetPr(plhs[0]), mxGetPr(plhs[1]),
                     (int32_t *)mxGetData(plhs[2]), 0));
  } else if (!strcmp(cmd, "pred_mat")) {                 /* [PhiT,GT] = mek_mex('pred_mat', a(6,1), c(6,1), d, Np, Nc)  <- predMat.m:1 */
    int Np = (int)mxGetScalar(prhs[4]), Nc = (int)mxGetScalar(prhs[5]);
    double d = mxGetScalar(prhs[3]);
    if (Np < 1) Np = 1;
    if (Nc < 1) Nc = 1;
    if (Nc > Np) Nc = Np;
    plhs[0] = mxCreateDoubleMatrix(7, Np, mxREAL);       /* Phi.' */
    plhs[1] = mxCreateDoubleMatrix(Nc, Np, mxREAL);      /* G.'   */
    chk(mek_pred_mat(1, mxGetPr(prhs[1]), mxGetPr(prhs[2]), &d, Np, Nc, mxGetPr(plhs[0]), mxGetPr(plhs[1]), 0));
  } else if (!strcmp(cmd, "pop_destroy")) {
    mek_pop_destroy((mek_pop *)get_handle(prhs[1]));
  } else if (!strcmp(cmd, "bank_destroy")) {
    mek_bank_destroy((mek_bank *)get_handle(prhs[1]));
  } else {
    mexErrMsgIdAndTxt("mek:usage", "unknown command %s", cmd);
  }
  mxFree(cmd);
}
