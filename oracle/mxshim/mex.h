/* oracle/mxshim/mex.h -- TEST INFRASTRUCTURE, not product code.
 *
 * A minimal stand-in for MATLAB's mex.h, just large enough to compile the
 * reference's native E-step (/root/reference/src/vbhem/vbhem_hmm_bwd_fwd_mex.c)
 * unmodified and in place, so that it can be driven from ctypes as the parity
 * oracle and as the CPU baseline.  Only the 19 mx / mex symbols that file uses
 * are declared (SURVEY.md section 8c lists them).
 */
#ifndef VBHEM_ORACLE_MEX_SHIM_H
#define VBHEM_ORACLE_MEX_SHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef size_t mwSize;
typedef size_t mwIndex;

typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxUNKNOWN_CLASS = 0, mxCELL_CLASS = 1, mxSTRUCT_CLASS = 2, mxDOUBLE_CLASS = 6 } mxClassID;

typedef struct mxArray_tag mxArray;

#define mxAssert(cond, msg) ((void)0)

/* queries */
int     mxIsCell(const mxArray *a);
int     mxIsDouble(const mxArray *a);
int     mxIsComplex(const mxArray *a);
size_t  mxGetM(const mxArray *a);
size_t  mxGetN(const mxArray *a);
mwSize  mxGetNumberOfDimensions(const mxArray *a);
const mwSize *mxGetDimensions(const mxArray *a);
size_t  mxGetNumberOfElements(const mxArray *a);
double *mxGetPr(const mxArray *a);
double  mxGetScalar(const mxArray *a);
mxArray *mxGetCell(const mxArray *a, mwIndex i);
mxArray *mxGetField(const mxArray *a, mwIndex i, const char *name);

/* construction */
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity c);
mxArray *mxCreateCellMatrix(mwSize m, mwSize n);
void     mxSetCell(mxArray *a, mwIndex i, mxArray *v);

/* memory */
void *mxMalloc(size_t n);
void *mxCalloc(size_t n, size_t sz);
void  mxFree(void *p);

/* error: never returns (longjmps back into the shim driver) */
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...);
int  mexPrintf(const char *fmt, ...);

/* the gateway every MEX file exports */
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);

#ifdef __cplusplus
}
#endif
#endif
