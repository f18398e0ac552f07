/* Minimal stand-in for MATLAB's mex.h / matrix.h, used ONLY to syntax-check mek_mex.c in an
 * environment without MATLAB (tests/test_host_abi.py).  A real build uses MATLAB's own mex.h:
 *     mex -I<repo>/include mek_mex.c -L<repo>/mpc-ekf4fastcharge_b200/mek -lmek              */
#ifndef MEX_STUB_H
#define MEX_STUB_H
#include <stddef.h>
#include <stdint.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef enum { mxREAL = 0 } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6, mxINT32_CLASS = 12, mxUINT32_CLASS = 13, mxUINT64_CLASS = 15 } mxClassID;
double *mxGetPr(const mxArray *);
void *mxGetData(const mxArray *);
double mxGetScalar(const mxArray *);
mwSize mxGetM(const mxArray *);
mwSize mxGetN(const mxArray *);
mwSize mxGetNumberOfElements(const mxArray *);
int mxIsEmpty(const mxArray *);
int mxIsDouble(const mxArray *);
mxArray *mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray *mxCreateDoubleScalar(double);
mxArray *mxCreateNumericMatrix(mwSize, mwSize, mxClassID, mxComplexity);
char *mxArrayToString(const mxArray *);
void *mxMalloc(size_t);
void mxFree(void *);
void mexErrMsgIdAndTxt(const char *, const char *, ...);
void mexLock(void);
#endif
