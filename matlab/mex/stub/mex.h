/* Stub of MATLAB's mex.h -- ONLY for the compile check of matlab/mex/mexDotSocpGPU.cpp in this repository
 * (MATLAB is not available offline).  It declares exactly the subset of the MEX C API the gateway uses, with the
 * documented MathWorks signatures (R2018a+ "interleaved complex" API not needed: real double arrays only).
 * A real build uses MATLAB's own header:  mex -R2018a mexDotSocpGPU.cpp -I../../include -L../../dotsocp_b200 -ldotsocp
 */
#ifndef DOTSOCP_STUB_MEX_H
#define DOTSOCP_STUB_MEX_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef enum { mxREAL = 0, mxCOMPLEX } mxComplexity;
double *mxGetPr(const mxArray *pa);
double mxGetScalar(const mxArray *pa);
size_t mxGetM(const mxArray *pa);
size_t mxGetN(const mxArray *pa);
size_t mxGetNumberOfElements(const mxArray *pa);
bool mxIsDouble(const mxArray *pa);
bool mxIsComplex(const mxArray *pa);
bool mxIsSparse(const mxArray *pa);
bool mxIsEmpty(const mxArray *pa);
bool mxIsStruct(const mxArray *pa);
bool mxIsChar(const mxArray *pa);
mxArray *mxGetField(const mxArray *pa, mwIndex index, const char *fieldname);
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity flag);
mxArray *mxCreateDoubleScalar(double value);
mxArray *mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char **fieldnames);
void mxSetField(mxArray *pa, mwIndex index, const char *fieldname, mxArray *value);
char *mxArrayToString(const mxArray *pa);
void mxFree(void *ptr);
double mxGetNaN(void);
void mexErrMsgIdAndTxt(const char *identifier, const char *fmt, ...);
void mexWarnMsgIdAndTxt(const char *identifier, const char *fmt, ...);
int mexPrintf(const char *fmt, ...);
void mexLock(void);
int mexAtExit(void (*fn)(void));
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
#ifdef __cplusplus
}
#endif
#endif
