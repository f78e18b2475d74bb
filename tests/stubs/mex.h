/* Minimal stand-in for MATLAB's mex.h: only the declarations fmpc_mex.c uses, so the shim can be
 * compile-checked (gcc -fsyntax-only) in an image without MATLAB.  Not a MATLAB API implementation. */
#ifndef STUB_MEX_H
#define STUB_MEX_H
#include <stddef.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6, mxINT32_CLASS = 12, mxUINT64_CLASS = 15 } mxClassID;
double *mxGetPr(const mxArray *);
void *mxGetData(const mxArray *);
double mxGetScalar(const mxArray *);
int mxIsEmpty(const mxArray *);
int mxIsUint64(const mxArray *);
size_t mxGetM(const mxArray *);
size_t mxGetN(const mxArray *);
size_t mxGetNumberOfElements(const mxArray *);
mwSize mxGetNumberOfDimensions(const mxArray *);
const mwSize *mxGetDimensions(const mxArray *);
mxArray *mxGetField(const mxArray *, mwSize, const char *);
int mxGetString(const mxArray *, char *, mwSize);
mxArray *mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray *mxCreateDoubleScalar(double);
mxArray *mxCreateNumericMatrix(mwSize, mwSize, mxClassID, mxComplexity);
mxArray *mxCreateNumericArray(mwSize, const mwSize *, mxClassID, mxComplexity);
void mxDestroyArray(mxArray *);
void mexErrMsgIdAndTxt(const char *, const char *, ...);
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
#endif
