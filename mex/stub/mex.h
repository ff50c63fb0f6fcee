/* Stand-in for the MATLAB / GNU Octave <mex.h> in an image that has neither (no mkoctfile, no mex.h): it
 * declares exactly the API subset mex/ekfslam_mex.c uses.  mex/stub/mexrt.c IMPLEMENTS that subset (a tiny
 * mxArray runtime: full double matrices, char rows, struct arrays), so the gateway can be compiled, linked
 * against libekfslam.so and driven from a test (tests/test_mex_gateway.py) exactly as Octave would drive it. */
#ifndef EKFSLAM_MEX_STUB_H
#define EKFSLAM_MEX_STUB_H
#include <stddef.h>
#include <stdint.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
#ifdef __cplusplus
extern "C" {
#endif
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...);
void mexLock(void);
int mexAtExit(void (*fn)(void));
mxArray* mxGetField(const mxArray* s, mwIndex i, const char* name);
void mxSetField(mxArray* s, mwIndex i, const char* name, mxArray* v);
int mxGetFieldNumber(const mxArray* s, const char* name);
int mxGetNumberOfFields(const mxArray* s);
const char* mxGetFieldNameByNumber(const mxArray* s, int k);
mxArray* mxGetFieldByNumber(const mxArray* s, mwIndex i, int k);
void mxSetFieldByNumber(mxArray* s, mwIndex i, int k, mxArray* v);
int mxAddField(mxArray* s, const char* name);
double* mxGetPr(const mxArray* a);
double mxGetScalar(const mxArray* a);
size_t mxGetM(const mxArray* a);
size_t mxGetN(const mxArray* a);
size_t mxGetNumberOfElements(const mxArray* a);
int mxIsEmpty(const mxArray* a);
int mxIsStruct(const mxArray* a);
int mxIsChar(const mxArray* a);
int mxIsDouble(const mxArray* a);
int mxIsSparse(const mxArray* a);
mwIndex* mxGetIr(const mxArray* a);
mwIndex* mxGetJc(const mxArray* a);
char* mxArrayToString(const mxArray* a);
void mxFree(void* p);
void* mxCalloc(size_t n, size_t sz);
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray* mxCreateDoubleScalar(double v);
mxArray* mxCreateString(const char* s);
mxArray* mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char** names);
mxArray* mxDuplicateArray(const mxArray* a);
void mxDestroyArray(mxArray* a);
#ifdef __cplusplus
}
#endif
#endif
