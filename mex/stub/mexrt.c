/* mexrt.c - a minimal mxArray runtime implementing the API subset of mex/stub/mex.h (TEST INFRASTRUCTURE:
 * it lets mex/ekfslam_mex.c be linked and driven without MATLAB / Octave).  Full double matrices
 * (column-major), char row vectors and 1xN struct arrays; sparse arrays cannot be created (mxIsSparse is
 * always 0).  mexErrMsgIdAndTxt does not return: mexrt_call() wraps mexFunction in setjmp/longjmp and hands
 * the message back, like the interpreter turns it into an error(). */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mex.h"

enum { CLS_DOUBLE = 0, CLS_CHAR = 1, CLS_STRUCT = 2 };
struct mxArray_tag {
    int cls;
    size_t m, n;
    double* pr;         /* doubles, or character codes for CLS_CHAR */
    int nfields;
    char** names;
    mxArray** fields;   /* [numel][nfields] */
};

static jmp_buf g_jmp;
static int g_armed = 0;
static char g_err[1024];
static char g_errid[128];
static void (*g_atexit)(void) = NULL;
static int g_locked = 0;

void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    snprintf(g_errid, sizeof g_errid, "%s", id ? id : "");
    if (g_armed) longjmp(g_jmp, 1);
    fprintf(stderr, "mexErrMsgIdAndTxt outside mexrt_call: %s: %s\n", g_errid, g_err);
    abort();
}
void mexLock(void) { g_locked++; }
int mexAtExit(void (*fn)(void)) { g_atexit = fn; return 0; }

/* test entry points ------------------------------------------------------------------------ */
int mexrt_call(int nlhs, mxArray** plhs, int nrhs, const mxArray** prhs) {
    int rc = 0;
    g_err[0] = 0; g_errid[0] = 0;
    g_armed = 1;
    if (setjmp(g_jmp) == 0) mexFunction(nlhs, plhs, nrhs, prhs);
    else rc = 1;
    g_armed = 0;
    return rc;
}
const char* mexrt_last_error(void) { return g_err; }
const char* mexrt_last_error_id(void) { return g_errid; }
int mexrt_lock_count(void) { return g_locked; }
void mexrt_run_atexit(void) { if (g_atexit) g_atexit(); }

/* mx API --------------------------------------------------------------------------------------- */
static mxArray* new_array(int cls, size_t m, size_t n) {
    mxArray* a = (mxArray*)calloc(1, sizeof(mxArray));
    a->cls = cls; a->m = m; a->n = n;
    return a;
}
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) {
    mxArray* a = new_array(CLS_DOUBLE, m, n);
    (void)c;
    a->pr = (double*)calloc((m * n) != 0 ? m * n : 1, sizeof(double));
    return a;
}
mxArray* mxCreateDoubleScalar(double v) {
    mxArray* a = mxCreateDoubleMatrix(1, 1, mxREAL);
    a->pr[0] = v;
    return a;
}
mxArray* mxCreateString(const char* s) {
    const size_t n = strlen(s);
    mxArray* a = new_array(CLS_CHAR, n ? 1 : 0, n);
    size_t i;
    a->pr = (double*)calloc(n ? n : 1, sizeof(double));
    for (i = 0; i < n; ++i) a->pr[i] = (double)(unsigned char)s[i];
    return a;
}
mxArray* mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char** names) {
    mxArray* a = new_array(CLS_STRUCT, m, n);
    int k;
    a->nfields = nfields;
    a->names = (char**)calloc(nfields ? nfields : 1, sizeof(char*));
    for (k = 0; k < nfields; ++k) a->names[k] = strdup(names[k]);
    a->fields = (mxArray**)calloc(((m * n) != 0 ? m * n : 1) * (nfields ? nfields : 1), sizeof(mxArray*));
    return a;
}
size_t mxGetM(const mxArray* a) { return a->m; }
size_t mxGetN(const mxArray* a) { return a->n; }
size_t mxGetNumberOfElements(const mxArray* a) { return a->m * a->n; }
int mxIsEmpty(const mxArray* a) { return a->m * a->n == 0; }
int mxIsStruct(const mxArray* a) { return a->cls == CLS_STRUCT; }
int mxIsChar(const mxArray* a) { return a->cls == CLS_CHAR; }
int mxIsDouble(const mxArray* a) { return a->cls == CLS_DOUBLE; }
int mxIsSparse(const mxArray* a) { (void)a; return 0; }
mwIndex* mxGetIr(const mxArray* a) { (void)a; return NULL; }
mwIndex* mxGetJc(const mxArray* a) { (void)a; return NULL; }
double* mxGetPr(const mxArray* a) { return a->pr; }
double mxGetScalar(const mxArray* a) { return a->pr[0]; }
char* mxArrayToString(const mxArray* a) {
    const size_t n = a->m * a->n;
    char* s = (char*)calloc(n + 1, 1);
    size_t i;
    for (i = 0; i < n; ++i) s[i] = (char)(int)a->pr[i];
    return s;
}
void mxFree(void* p) { free(p); }
void* mxCalloc(size_t n, size_t sz) { return calloc(n ? n : 1, sz ? sz : 1); }
int mxGetNumberOfFields(const mxArray* s) { return s->cls == CLS_STRUCT ? s->nfields : 0; }
const char* mxGetFieldNameByNumber(const mxArray* s, int k) { return (k >= 0 && k < s->nfields) ? s->names[k] : NULL; }
int mxGetFieldNumber(const mxArray* s, const char* name) {
    int k;
    if (s->cls != CLS_STRUCT) return -1;
    for (k = 0; k < s->nfields; ++k)
        if (!strcmp(s->names[k], name)) return k;
    return -1;
}
mxArray* mxGetFieldByNumber(const mxArray* s, mwIndex i, int k) {
    if (s->cls != CLS_STRUCT || k < 0 || k >= s->nfields || i >= s->m * s->n) return NULL;
    return s->fields[i * s->nfields + k];
}
mxArray* mxGetField(const mxArray* s, mwIndex i, const char* name) {
    return mxGetFieldByNumber(s, i, mxGetFieldNumber(s, name));
}
void mxSetFieldByNumber(mxArray* s, mwIndex i, int k, mxArray* v) {
    if (s->cls != CLS_STRUCT || k < 0 || k >= s->nfields || i >= s->m * s->n) return;
    s->fields[i * s->nfields + k] = v;
}
void mxSetField(mxArray* s, mwIndex i, const char* name, mxArray* v) {
    mxSetFieldByNumber(s, i, mxGetFieldNumber(s, name), v);
}
int mxAddField(mxArray* s, const char* name) {
    const size_t ne = (s->m * s->n) != 0 ? s->m * s->n : 1;
    const int nf = s->nfields;
    mxArray** nfld = (mxArray**)calloc(ne * (nf + 1), sizeof(mxArray*));
    size_t i;
    int k;
    for (i = 0; i < ne; ++i)
        for (k = 0; k < nf; ++k) nfld[i * (nf + 1) + k] = s->fields[i * nf + k];
    free(s->fields);
    s->fields = nfld;
    s->names = (char**)realloc(s->names, sizeof(char*) * (nf + 1));
    s->names[nf] = strdup(name);
    s->nfields = nf + 1;
    return nf;
}
mxArray* mxDuplicateArray(const mxArray* a) {
    mxArray* d;
    size_t i, ne;
    int k;
    if (!a) return NULL;
    ne = a->m * a->n;
    if (a->cls != CLS_STRUCT) {
        d = new_array(a->cls, a->m, a->n);
        d->pr = (double*)calloc(ne ? ne : 1, sizeof(double));
        memcpy(d->pr, a->pr, ne * sizeof(double));
        return d;
    }
    d = mxCreateStructMatrix(a->m, a->n, a->nfields, (const char**)a->names);
    for (i = 0; i < ne; ++i)
        for (k = 0; k < a->nfields; ++k) d->fields[i * a->nfields + k] = mxDuplicateArray(a->fields[i * a->nfields + k]);
    return d;
}
void mxDestroyArray(mxArray* a) {
    size_t i, ne;
    int k;
    if (!a) return;
    ne = a->m * a->n;
    if (a->cls == CLS_STRUCT) {
        for (i = 0; i < ne; ++i)
            for (k = 0; k < a->nfields; ++k) mxDestroyArray(a->fields[i * a->nfields + k]);
        for (k = 0; k < a->nfields; ++k) free(a->names[k]);
        free(a->names);
        free(a->fields);
    }
    free(a->pr);
    free(a);
}
