/*
 * ekfslam_mex.c — MATLAB / GNU Octave MEX gateway over the C ABI of libekfslam.so
 * (include/ekfslam.h), so that the reference driver matlab_code/mono_slam.m:50-82 can swap the GPU
 * path in with an addpath of mex/shims (same-named .m files that forward here).
 *
 *   [filter, features_info] = ekfslam_mex('ekf_prediction', filter, features_info)
 *   features_info = ekfslam_mex('search_IC_matches', filter, features_info, cam, cand)   % cand = [2xN zc; 1xN has]
 *   features_info = ekfslam_mex('ransac_hypotheses', filter, features_info, cam, u)      % u = uniform stream
 *   filter        = ekfslam_mex('ekf_update_li_inliers', filter, features_info)
 *   features_info = ekfslam_mex('rescue_hi_inliers', filter, features_info, cam)
 *   filter        = ekfslam_mex('ekf_update_hi_inliers', filter, features_info)
 *   [filter, features_info] = ekfslam_mex('map_management', filter, features_info, cam, det, min_n, step)  % det = [u; v; descriptor] 3xK
 *   [filter, features_info] = ekfslam_mex('inversedepth_2_cartesian', filter, features_info)
 *   [X, P]                  = ekfslam_mex('delete_a_feature', X, P, featToDelete, features_info)
 *   [X, P, newFeature]      = ekfslam_mex('add_features_inverse_depth', uvd, X, P, cam, std_pxl, initial_rho, std_rho)
 *
 * Struct layouts: `filter` mc/ekf_filter.m:37-59, `features_info` mc/add_feature_to_info_vector.m:7-32,
 * `cam` mc/initialize_cam.m:12-25.  Value semantics like the reference (inputs are never modified;
 * outputs are modified copies).  A persistent single-filter context is kept per (N, n) and released
 * at exit (mexLock / mexAtExit).  Errors surface through mexErrMsgIdAndTxt with the library message.
 *
 * Build (where Octave or MATLAB exists):  mkoctfile --mex -Iinclude mex/ekfslam_mex.c -Lekf-slam_b200 -lekfslam
 *                                         mex -Iinclude mex/ekfslam_mex.c -Lekf-slam_b200 -lekfslam
 * Neither tool is installed in the authoring image.  The gateway is therefore compiled against mex/stub/mex.h and
 * LINKED with mex/stub/mexrt.c (a minimal mxArray runtime) and libekfslam.so, and tests/test_mex_gateway.py drives
 * mexFunction through that runtime the way Octave would (golden frame, full step, map management, error paths).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "mex.h"
#include "ekfslam.h"

static ekfslam_ctx* g_ctx = NULL;
static int g_N = 0, g_n = 0;

static void cleanup(void) {
    if (g_ctx) { ekfslam_destroy(g_ctx); g_ctx = NULL; }
}

static void ck(int rc, const char* what) {
    if (rc != 0) mexErrMsgIdAndTxt("ekfslam:lib", "%s failed (%d): %s", what, rc, ekfslam_last_error());
}

static double field_scalar(const mxArray* s, const char* name, double dflt) {
    const mxArray* f = mxGetField(s, 0, name);
    return (f && !mxIsEmpty(f)) ? mxGetScalar(f) : dflt;
}

static double field_scalar_at(const mxArray* s, mwIndex i, const char* name, double dflt) {
    const mxArray* f = mxIsStruct(s) ? mxGetField(s, i, name) : NULL;
    return (f && !mxIsEmpty(f)) ? mxGetScalar(f) : dflt;
}

static int feature_type(const mxArray* fi, mwIndex i) {
    const mxArray* t = mxGetField(fi, i, "type");
    char* str;
    int ty;
    if (!t || !mxIsChar(t)) mexErrMsgIdAndTxt("ekfslam:arg", "features_info(%d).type missing", (int)i + 1);
    str = mxArrayToString(t);
    ty = (str[0] == 'c') ? EKFSLAM_FEAT_CARTESIAN : EKFSLAM_FEAT_INVERSEDEPTH;  /* strncmp(.,1) as mc/calculate_Hi_inverse_depth.m:10 */
    mxFree(str);
    return ty;
}

/* (re)creates the context for N features / state size n and loads camera + noise parameters */
static void ensure_ctx(int N, int n, const mxArray* filter, const mxArray* cam) {
    if (N < 1) N = 1;
    if (!g_ctx || g_N != N || g_n != n) {
        cleanup();
        ck(ekfslam_create(&g_ctx, 0, 1, N, n), "ekfslam_create");
        g_N = N; g_n = n;
        mexLock();
        mexAtExit(cleanup);
    }
    if (filter) {
        ekfslam_params p;
        ekfslam_default_params(&p);
        p.std_a = field_scalar(filter, "std_a", p.std_a);
        p.std_alpha = field_scalar(filter, "std_alpha", p.std_alpha);
        p.std_z = field_scalar(filter, "std_z", p.std_z);
        ck(ekfslam_set_params(g_ctx, &p), "ekfslam_set_params");
    }
    if (cam) {
        ekfslam_camera c;
        ekfslam_default_camera(&c);
        c.k1 = field_scalar(cam, "k1", c.k1); c.k2 = field_scalar(cam, "k2", c.k2);
        c.Cx = field_scalar(cam, "Cx", c.Cx); c.Cy = field_scalar(cam, "Cy", c.Cy);
        c.f = field_scalar(cam, "f", c.f); c.dx = field_scalar(cam, "dx", c.dx); c.dy = field_scalar(cam, "dy", c.dy);
        c.nRows = (int32_t)field_scalar(cam, "nRows", c.nRows); c.nCols = (int32_t)field_scalar(cam, "nCols", c.nCols);
        ck(ekfslam_set_camera(g_ctx, &c), "ekfslam_set_camera");
    }
}

typedef struct {
    int N, n;
    uint8_t *type, *flags;
    int32_t* off;
    double *h, *Hc, *S, *z;
} feats_t;

static void feats_alloc(feats_t* f, int N) {
    const int M = N > 0 ? N : 1;
    f->N = N;
    f->type = (uint8_t*)mxCalloc(M, 1); f->flags = (uint8_t*)mxCalloc(M, 1); f->off = (int32_t*)mxCalloc(M, 4);
    f->h = (double*)mxCalloc(2 * M, 8); f->Hc = (double*)mxCalloc(26 * M, 8); f->S = (double*)mxCalloc(4 * M, 8);
    f->z = (double*)mxCalloc(2 * M, 8);
}

static double dense_or_sparse(const mxArray* H, int r, int c) {
    /* features_info(i).H is sparse in the reference (mc/calculate_derivatives.m:12,16) */
    if (mxIsSparse(H)) {
        const mwIndex *ir = mxGetIr(H), *jc = mxGetJc(H);
        const double* pr = mxGetPr(H);
        mwIndex k;
        for (k = jc[c]; k < jc[c + 1]; ++k)
            if ((int)ir[k] == r) return pr[k];
        return 0.0;
    }
    return mxGetPr(H)[(size_t)c * mxGetM(H) + r];
}

/* features_info struct array -> flat arrays */
static void feats_read(const mxArray* fi, feats_t* f) {
    const int N = mxIsStruct(fi) ? (int)mxGetNumberOfElements(fi) : 0;   /* features_info = [] (mono_slam.m:35) */
    int i, pos = 13, r, c;
    feats_alloc(f, N);
    for (i = 0; i < N; ++i) {
        const mxArray *h = mxGetField(fi, i, "h"), *H = mxGetField(fi, i, "H"), *S = mxGetField(fi, i, "S"),
                      *z = mxGetField(fi, i, "z");
        const int w = (f->type[i] = (uint8_t)feature_type(fi, i)) == EKFSLAM_FEAT_INVERSEDEPTH ? 6 : 3;
        f->off[i] = pos;
        if (h && !mxIsEmpty(h)) { f->h[2 * i] = mxGetPr(h)[0]; f->h[2 * i + 1] = mxGetPr(h)[1]; f->flags[i] |= EKFSLAM_F_HAS_H; }
        if (z && !mxIsEmpty(z)) { f->z[2 * i] = mxGetPr(z)[0]; f->z[2 * i + 1] = mxGetPr(z)[1]; f->flags[i] |= EKFSLAM_F_HAS_Z; }
        if (S && !mxIsEmpty(S)) for (c = 0; c < 2; ++c) for (r = 0; r < 2; ++r) f->S[4 * i + 2 * r + c] = mxGetPr(S)[2 * c + r];
        if (H && !mxIsEmpty(H))
            for (r = 0; r < 2; ++r) {
                for (c = 0; c < 7; ++c) f->Hc[26 * i + 13 * r + c] = dense_or_sparse(H, r, c);
                for (c = 0; c < w; ++c) f->Hc[26 * i + 13 * r + 7 + c] = dense_or_sparse(H, r, pos + c);
            }
        {
            const mxArray* a = mxGetField(fi, i, "individually_compatible");
            const mxArray* b = mxGetField(fi, i, "low_innovation_inlier");
            const mxArray* d = mxGetField(fi, i, "high_innovation_inlier");
            if (a && !mxIsEmpty(a) && mxGetScalar(a) != 0) f->flags[i] |= EKFSLAM_F_IC;
            if (b && !mxIsEmpty(b) && mxGetScalar(b) != 0) f->flags[i] |= EKFSLAM_F_LI;
            if (d && !mxIsEmpty(d) && mxGetScalar(d) != 0) f->flags[i] |= EKFSLAM_F_HI;
        }
        pos += w;
    }
    f->n = pos;
}

static void feats_upload(const feats_t* f) {
    int32_t nf = f->N;
    ck(ekfslam_upload_feature_types(g_ctx, 0, 1, f->type, &nf), "ekfslam_upload_feature_types");
    ck(ekfslam_upload_features(g_ctx, 0, 1, f->h, f->Hc, f->S, f->z, f->flags), "ekfslam_upload_features");
}

static void set_or_add(mxArray* s, mwIndex i, const char* name, mxArray* v) {
    if (mxGetFieldNumber(s, name) < 0) mxAddField(s, name);
    mxSetField(s, i, name, v);
}

/* device -> a modified copy of features_info; what: bit0 h, bit1 H, bit2 S, bit3 z, bit4 flags */
static mxArray* feats_download(const mxArray* fi_in, const feats_t* f, int what) {
    mxArray* fi = mxDuplicateArray(fi_in);
    const int N = f->N;
    int i, r, c;
    feats_t d;
    feats_alloc(&d, N);
    ck(ekfslam_download_features(g_ctx, 0, 1, d.h, d.Hc, d.S, d.z, d.flags, NULL, NULL), "ekfslam_download_features");
    for (i = 0; i < N; ++i) {
        const int fl = d.flags[i], w = f->type[i] == EKFSLAM_FEAT_INVERSEDEPTH ? 6 : 3;
        if (what & 1) {
            mxArray* h = (fl & EKFSLAM_F_HAS_H) ? mxCreateDoubleMatrix(1, 2, mxREAL) : mxCreateDoubleMatrix(0, 0, mxREAL);
            if (fl & EKFSLAM_F_HAS_H) { mxGetPr(h)[0] = d.h[2 * i]; mxGetPr(h)[1] = d.h[2 * i + 1]; }
            set_or_add(fi, i, "h", h);
        }
        if (what & 2) {
            mxArray* H = (fl & EKFSLAM_F_HAS_H) ? mxCreateDoubleMatrix(2, f->n, mxREAL) : mxCreateDoubleMatrix(0, 0, mxREAL);
            if (fl & EKFSLAM_F_HAS_H)
                for (r = 0; r < 2; ++r) {
                    for (c = 0; c < 7; ++c) mxGetPr(H)[2 * c + r] = d.Hc[26 * i + 13 * r + c];
                    for (c = 0; c < w; ++c) mxGetPr(H)[2 * (f->off[i] + c) + r] = d.Hc[26 * i + 13 * r + 7 + c];
                }
            set_or_add(fi, i, "H", H);
        }
        if (what & 4) {
            mxArray* S = (fl & EKFSLAM_F_HAS_H) ? mxCreateDoubleMatrix(2, 2, mxREAL) : mxCreateDoubleMatrix(0, 0, mxREAL);
            if (fl & EKFSLAM_F_HAS_H) for (c = 0; c < 2; ++c) for (r = 0; r < 2; ++r) mxGetPr(S)[2 * c + r] = d.S[4 * i + 2 * r + c];
            set_or_add(fi, i, "S", S);
        }
        if (what & 8) {
            mxArray* z = (fl & EKFSLAM_F_HAS_Z) ? mxCreateDoubleMatrix(2, 1, mxREAL) : mxCreateDoubleMatrix(0, 0, mxREAL);
            if (fl & EKFSLAM_F_HAS_Z) { mxGetPr(z)[0] = d.z[2 * i]; mxGetPr(z)[1] = d.z[2 * i + 1]; }
            set_or_add(fi, i, "z", z);
        }
        if (what & 16) {
            set_or_add(fi, i, "individually_compatible", mxCreateDoubleScalar((fl & EKFSLAM_F_IC) ? 1 : 0));
            set_or_add(fi, i, "low_innovation_inlier", mxCreateDoubleScalar((fl & EKFSLAM_F_LI) ? 1 : 0));
            set_or_add(fi, i, "high_innovation_inlier", mxCreateDoubleScalar((fl & EKFSLAM_F_HI) ? 1 : 0));
        }
    }
    return fi;
}

static void state_upload(const mxArray* filter, const char* xname, const char* pname, int which, int n) {
    const mxArray *x = mxGetField(filter, 0, xname), *P = pname ? mxGetField(filter, 0, pname) : NULL;
    int32_t ns = n;
    if (!x || (int)mxGetNumberOfElements(x) != n)
        mexErrMsgIdAndTxt("ekfslam:arg", "filter.%s must have %d elements (13 + 6*N_id + 3*N_c)", xname, n);
    if (P && (mxIsSparse(P) || (int)mxGetM(P) != n || (int)mxGetN(P) != n))
        mexErrMsgIdAndTxt("ekfslam:arg", "filter.%s must be a full %d x %d matrix", pname, n, n);
    ck(ekfslam_upload_state(g_ctx, 0, 1, which, mxGetPr(x), P ? mxGetPr(P) : NULL, &ns), "ekfslam_upload_state");
}

static mxArray* state_download(const mxArray* filter_in, const char* xname, const char* pname, int which, int n) {
    mxArray* filter = mxDuplicateArray(filter_in);
    mxArray *x = mxCreateDoubleMatrix(n, 1, mxREAL), *P = mxCreateDoubleMatrix(n, n, mxREAL);
    ck(ekfslam_download_state(g_ctx, 0, 1, which, mxGetPr(x), mxGetPr(P), NULL), "ekfslam_download_state");
    set_or_add(filter, 0, xname, x);
    set_or_add(filter, 0, pname, P);
    return filter;
}

/* ---- map management (mc/map_management.m and the functions it calls) --------------------------------------- */
/* state of size n into a context whose n_max may be larger (zero padded) */
static void state_upload_padded(const double* x, const double* P, int n, int nmax) {
    double* xs = (double*)mxCalloc(nmax, 8);
    double* Ps = (double*)mxCalloc((size_t)nmax * nmax, 8);
    int32_t ns = n;
    int r, c;
    for (r = 0; r < n; ++r) xs[r] = x[r];
    for (c = 0; c < n; ++c) for (r = 0; r < n; ++r) Ps[(size_t)c * nmax + r] = P[(size_t)c * n + r];
    ck(ekfslam_upload_state(g_ctx, 0, 1, 0, xs, Ps, &ns), "ekfslam_upload_state");
    mxFree(xs); mxFree(Ps);
}

/* (x_k_k, p_k_k) of the context, trimmed to its current state size; returns n */
static int state_download_trimmed(int nmax, mxArray** xo, mxArray** Po) {
    double* xs = (double*)mxCalloc(nmax, 8);
    double* Ps = (double*)mxCalloc((size_t)nmax * nmax, 8);
    int32_t ns = 0;
    int r, c, n;
    ck(ekfslam_download_state(g_ctx, 0, 1, 0, xs, Ps, &ns), "ekfslam_download_state");
    n = ns;
    *xo = mxCreateDoubleMatrix(n, 1, mxREAL);
    *Po = mxCreateDoubleMatrix(n, n, mxREAL);
    for (r = 0; r < n; ++r) mxGetPr(*xo)[r] = xs[r];
    for (c = 0; c < n; ++c) for (r = 0; r < n; ++r) mxGetPr(*Po)[(size_t)c * n + r] = Ps[(size_t)c * nmax + r];
    mxFree(xs); mxFree(Ps);
    return n;
}

static const char* FI_FIELDS[] = {   /* mc/add_feature_to_info_vector.m:7-32 */
    "patch_when_initialized", "feature_when_initialized", "patch_when_matching", "r_wc_when_initialized",
    "R_wc_when_initialized", "uv_when_initialized", "half_patch_size_when_initialized", "half_patch_size_when_matching",
    "times_predicted", "times_measured", "init_frame", "init_measurement", "type", "yi", "individually_compatible",
    "low_innovation_inlier", "high_innovation_inlier", "z", "h", "H", "S", "state_size", "measurement_size", "R"};
#define FI_NFIELDS ((int)(sizeof(FI_FIELDS) / sizeof(FI_FIELDS[0])))

static mxArray* vec(const double* v, int m, int n) {
    mxArray* a = mxCreateDoubleMatrix(m, n, mxREAL);
    int i;
    for (i = 0; i < m * n; ++i) mxGetPr(a)[i] = v[i];
    return a;
}

/* [filter, features_info] = map_management(filter, features_info, cam, det, min_n, step) */
static void cmd_map_management(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[], feats_t* f, int only_convert) {
    const mxArray* filter = prhs[1];
    const mxArray* fi_in = prhs[2];
    const int N = f->N;
    const int K = (!only_convert && nrhs > 4 && !mxIsEmpty(prhs[4])) ? (int)mxGetN(prhs[4]) : 0;
    const int Kc = K > 50 ? 50 : K;
    const int Ncap = N + Kc > 0 ? N + Kc : 1, nmax = 13 + 6 * Ncap;
    const mxArray *x = mxGetField(filter, 0, "x_k_k"), *P = mxGetField(filter, 0, "p_k_k");
    int32_t *counters, *tags, *dtag, nd = K, nf = 0, conv = -1;
    uint8_t* types;
    double *duv, xv7[7];
    mxArray *xo, *Po, *fo, *fi_out;
    int i, k, n, pos;
    if (!only_convert && nrhs < 7) mexErrMsgIdAndTxt("ekfslam:arg", "map_management needs cam, det, min_number_of_features_in_image, step");
    if (K && (int)mxGetM(prhs[4]) != 3) mexErrMsgIdAndTxt("ekfslam:arg", "det must be 3 x K = [u; v; descriptor]");
    if (!x || !P || (int)mxGetNumberOfElements(x) != f->n || (int)mxGetM(P) != f->n)
        mexErrMsgIdAndTxt("ekfslam:arg", "filter.x_k_k / p_k_k do not match features_info (n = %d)", f->n);
    ensure_ctx(Ncap, nmax, filter, only_convert ? NULL : prhs[3]);
    /* layout, per-frame fields, bookkeeping */
    types = (uint8_t*)mxCalloc(Ncap, 1);
    for (i = 0; i < N; ++i) types[i] = f->type[i];
    nf = N;
    ck(ekfslam_upload_feature_types(g_ctx, 0, 1, types, &nf), "ekfslam_upload_feature_types");
    {
        feats_t g;
        feats_alloc(&g, Ncap);
        for (i = 0; i < N; ++i) g.flags[i] = f->flags[i];
        ck(ekfslam_upload_features(g_ctx, 0, 1, g.h, g.Hc, g.S, g.z, g.flags), "ekfslam_upload_features");
    }
    counters = (int32_t*)mxCalloc(2 * Ncap, 4);
    tags = (int32_t*)mxCalloc(Ncap, 4);
    for (i = 0; i < Ncap; ++i) tags[i] = -1;
    for (i = 0; i < N; ++i) {
        counters[2 * i] = (int32_t)field_scalar_at(fi_in, i, "times_predicted", 0);
        counters[2 * i + 1] = (int32_t)field_scalar_at(fi_in, i, "times_measured", 0);
        tags[i] = i;
    }
    ck(ekfslam_upload_feature_meta(g_ctx, 0, 1, counters, tags), "ekfslam_upload_feature_meta");
    state_upload_padded(mxGetPr(x), mxGetPr(P), f->n, nmax);
    if (only_convert) {
        ck(ekfslam_inversedepth_2_cartesian(g_ctx, 0.1, -1, &conv), "ekfslam_inversedepth_2_cartesian");
    } else {
        const int Ku = K > 0 ? K : 1;
        duv = (double*)mxCalloc(2 * Ku, 8);
        dtag = (int32_t*)mxCalloc(Ku, 4);
        for (k = 0; k < K; ++k) { duv[2 * k] = mxGetPr(prhs[4])[3 * k]; duv[2 * k + 1] = mxGetPr(prhs[4])[3 * k + 1]; dtag[k] = N + k; }
        ck(ekfslam_upload_detections(g_ctx, 0, 1, Ku, duv, dtag, &nd), "ekfslam_upload_detections");
        ck(ekfslam_map_management(g_ctx, (int)mxGetScalar(prhs[5])), "ekfslam_map_management");
    }
    /* results */
    n = state_download_trimmed(nmax, &xo, &Po);
    ck(ekfslam_download_feature_types(g_ctx, 0, 1, types, &nf), "ekfslam_download_feature_types");
    ck(ekfslam_download_feature_tags(g_ctx, 0, 1, tags), "ekfslam_download_feature_tags");
    ck(ekfslam_download_features(g_ctx, 0, 1, NULL, NULL, NULL, NULL, NULL, NULL, counters), "ekfslam_download_features");
    fo = mxDuplicateArray(filter);
    set_or_add(fo, 0, "x_k_k", xo);
    set_or_add(fo, 0, "p_k_k", Po);
    plhs[0] = fo;
    if (nlhs < 2) return;
    for (k = 0; k < 7; ++k) xv7[k] = mxGetPr(xo)[k];
    if (mxIsStruct(fi_in)) {
        const int nfl = mxGetNumberOfFields(fi_in);
        const char** names = (const char**)mxCalloc(nfl, sizeof(char*));
        for (k = 0; k < nfl; ++k) names[k] = mxGetFieldNameByNumber(fi_in, k);
        fi_out = mxCreateStructMatrix(1, nf, nfl, names);
    } else {
        fi_out = mxCreateStructMatrix(1, nf, FI_NFIELDS, FI_FIELDS);
    }
    pos = 13;
    for (i = 0; i < nf; ++i) {
        const int w = types[i] == EKFSLAM_FEAT_INVERSEDEPTH ? 6 : 3;
        if (tags[i] >= 0 && tags[i] < N) {                       /* surviving feature: copy every field */
            const int nfl = mxGetNumberOfFields(fi_in);
            for (k = 0; k < nfl; ++k) mxSetFieldByNumber(fi_out, i, k, mxDuplicateArray(mxGetFieldByNumber(fi_in, tags[i], k)));
        } else {                                                 /* new feature, mc/add_feature_to_info_vector.m:7-32 */
            const int j = tags[i] - N;
            const double* d = mxGetPr(prhs[4]) + 3 * j;
            const double eye2[4] = {1, 0, 0, 1};
            double R[9], r, qx, qy, qz;
            r = xv7[3]; qx = xv7[4]; qy = xv7[5]; qz = xv7[6];   /* q2r.m, column-major */
            R[0] = r*r+qx*qx-qy*qy-qz*qz; R[3] = 2*(qx*qy-r*qz);       R[6] = 2*(qz*qx+r*qy);
            R[1] = 2*(qx*qy+r*qz);       R[4] = r*r-qx*qx+qy*qy-qz*qz; R[7] = 2*(qy*qz-r*qx);
            R[2] = 2*(qz*qx-r*qy);       R[5] = 2*(qy*qz+r*qx);       R[8] = r*r-qx*qx-qy*qy+qz*qz;
            set_or_add(fi_out, i, "feature_when_initialized", mxCreateDoubleScalar(d[2]));
            set_or_add(fi_out, i, "r_wc_when_initialized", vec(xv7, 3, 1));
            set_or_add(fi_out, i, "R_wc_when_initialized", vec(R, 3, 3));
            set_or_add(fi_out, i, "uv_when_initialized", vec(d, 1, 2));
            set_or_add(fi_out, i, "half_patch_size_when_initialized", mxCreateDoubleScalar(20));
            set_or_add(fi_out, i, "half_patch_size_when_matching", mxCreateDoubleScalar(6));
            set_or_add(fi_out, i, "init_frame", mxCreateDoubleScalar(mxGetScalar(prhs[6])));
            set_or_add(fi_out, i, "init_measurement", vec(d, 2, 1));
            set_or_add(fi_out, i, "yi", vec(mxGetPr(xo) + pos, 6, 1));
            set_or_add(fi_out, i, "state_size", mxCreateDoubleScalar(6));
            set_or_add(fi_out, i, "measurement_size", mxCreateDoubleScalar(2));
            set_or_add(fi_out, i, "R", vec(eye2, 2, 2));
        }
        set_or_add(fi_out, i, "type", mxCreateString(types[i] == EKFSLAM_FEAT_INVERSEDEPTH ? "inversedepth" : "cartesian"));
        if (!only_convert) {                                     /* mc/update_features_info.m:4-18 ran on the device */
            set_or_add(fi_out, i, "times_predicted", mxCreateDoubleScalar(counters[2 * i]));
            set_or_add(fi_out, i, "times_measured", mxCreateDoubleScalar(counters[2 * i + 1]));
            set_or_add(fi_out, i, "individually_compatible", mxCreateDoubleScalar(0));
            set_or_add(fi_out, i, "low_innovation_inlier", mxCreateDoubleScalar(0));
            set_or_add(fi_out, i, "high_innovation_inlier", mxCreateDoubleScalar(0));
            set_or_add(fi_out, i, "h", mxCreateDoubleMatrix(0, 0, mxREAL));
            set_or_add(fi_out, i, "z", mxCreateDoubleMatrix(0, 0, mxREAL));
            set_or_add(fi_out, i, "H", mxCreateDoubleMatrix(0, 0, mxREAL));
            set_or_add(fi_out, i, "S", mxCreateDoubleMatrix(0, 0, mxREAL));
        }
        pos += w;
    }
    (void)n;
    plhs[1] = fi_out;
}

/* [X, P] = delete_a_feature(X, P, featToDelete, features_info)   (featToDelete 1-based, as in MATLAB) */
static void cmd_delete_a_feature(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    feats_t f;
    uint8_t* del;
    int32_t nf;
    int feat;
    mxArray *xo, *Po;
    if (nrhs < 5) mexErrMsgIdAndTxt("ekfslam:arg", "usage: ekfslam_mex('delete_a_feature', X, P, featToDelete, features_info)");
    feats_read(prhs[4], &f);
    feat = (int)mxGetScalar(prhs[3]);
    if (feat < 1 || feat > f.N) mexErrMsgIdAndTxt("ekfslam:arg", "featToDelete out of range");
    if ((int)mxGetNumberOfElements(prhs[1]) != f.n) mexErrMsgIdAndTxt("ekfslam:arg", "X does not match features_info");
    ensure_ctx(f.N, f.n, NULL, NULL);
    nf = f.N;
    ck(ekfslam_upload_feature_types(g_ctx, 0, 1, f.type, &nf), "ekfslam_upload_feature_types");
    state_upload_padded(mxGetPr(prhs[1]), mxGetPr(prhs[2]), f.n, f.n);
    del = (uint8_t*)mxCalloc(f.N, 1);
    del[feat - 1] = 1;
    ck(ekfslam_delete_features(g_ctx, 0, 1, del), "ekfslam_delete_features");
    state_download_trimmed(f.n, &xo, &Po);
    plhs[0] = xo;
    if (nlhs > 1) plhs[1] = Po;
}

/* [X_RES, P_RES, newFeature] = add_features_inverse_depth(uvd, X, P, cam, std_pxl, initial_rho, std_rho) */
static void cmd_add_features(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    int n, nNew, nold, Ncap, nmax, j;
    uint8_t* types;
    int32_t nf;
    mxArray *xo, *Po;
    if (nrhs < 8) mexErrMsgIdAndTxt("ekfslam:arg", "usage: ekfslam_mex('add_features_inverse_depth', uvd, X, P, cam, std_pxl, initial_rho, std_rho)");
    n = (int)mxGetNumberOfElements(prhs[2]);
    nNew = mxIsEmpty(prhs[1]) ? 0 : (int)mxGetN(prhs[1]);
    if ((n - 13) % 3 || n < 13) mexErrMsgIdAndTxt("ekfslam:arg", "state size %d is not 13 + 6*N_id + 3*N_c", n);
    nold = (n - 13) / 3;                 /* only the SIZE matters for an append: describe the map as Cartesian triples */
    Ncap = nold + nNew > 0 ? nold + nNew : 1;
    nmax = n + 6 * nNew;
    ensure_ctx(Ncap, nmax, NULL, prhs[4]);
    types = (uint8_t*)mxCalloc(Ncap, 1);
    for (j = 0; j < nold; ++j) types[j] = EKFSLAM_FEAT_CARTESIAN;
    nf = nold;
    ck(ekfslam_upload_feature_types(g_ctx, 0, 1, types, &nf), "ekfslam_upload_feature_types");
    state_upload_padded(mxGetPr(prhs[2]), mxGetPr(prhs[3]), n, nmax);
    for (j = 0; j < nNew; ++j)
        ck(ekfslam_add_features(g_ctx, 0, 1, mxGetPr(prhs[1]) + 2 * j, NULL, mxGetScalar(prhs[5]), mxGetScalar(prhs[6]), mxGetScalar(prhs[7])),
           "ekfslam_add_features");
    n = state_download_trimmed(nmax, &xo, &Po);
    plhs[0] = xo;
    if (nlhs > 1) plhs[1] = Po;
    if (nlhs > 2) plhs[2] = nNew ? vec(mxGetPr(xo) + n - 6, 6, 1) : mxCreateDoubleMatrix(0, 0, mxREAL);
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    char* cmd;
    feats_t f;
    if (nrhs < 1 || !mxIsChar(prhs[0])) mexErrMsgIdAndTxt("ekfslam:arg", "usage: ekfslam_mex(cmd, ...)");
    cmd = mxArrayToString(prhs[0]);
    if (!strcmp(cmd, "delete_a_feature")) { cmd_delete_a_feature(nlhs, plhs, nrhs, prhs); mxFree(cmd); return; }
    if (!strcmp(cmd, "add_features_inverse_depth")) { cmd_add_features(nlhs, plhs, nrhs, prhs); mxFree(cmd); return; }
    if (nrhs < 3 || !mxIsStruct(prhs[1]))
        mexErrMsgIdAndTxt("ekfslam:arg", "usage: ekfslam_mex(cmd, filter, features_info, ...)");
    feats_read(prhs[2], &f);
    if (!strcmp(cmd, "map_management")) { cmd_map_management(nlhs, plhs, nrhs, prhs, &f, 0); mxFree(cmd); return; }
    if (!strcmp(cmd, "inversedepth_2_cartesian")) { cmd_map_management(nlhs, plhs, nrhs, prhs, &f, 1); mxFree(cmd); return; }

    if (!strcmp(cmd, "ekf_prediction")) {                                   /* mc/ekf_prediction.m */
        ensure_ctx(f.N, f.n, prhs[1], NULL);
        feats_upload(&f);
        state_upload(prhs[1], "x_k_k", "p_k_k", 0, f.n);
        ck(ekfslam_predict(g_ctx), "ekfslam_predict");
        plhs[0] = state_download(prhs[1], "x_k_km1", "p_k_km1", 1, f.n);
        if (nlhs > 1) plhs[1] = mxDuplicateArray(prhs[2]);
    } else if (!strcmp(cmd, "search_IC_matches")) {                         /* mc/search_IC_matches.m */
        if (nrhs < 4) mexErrMsgIdAndTxt("ekfslam:arg", "search_IC_matches needs cam");
        ensure_ctx(f.N, f.n, prhs[1], prhs[3]);
        feats_upload(&f);
        state_upload(prhs[1], "x_k_km1", "p_k_km1", 1, f.n);
        ck(ekfslam_measure(g_ctx, 1), "ekfslam_measure");
        if (nrhs > 4 && !mxIsEmpty(prhs[4])) {                              /* cand: 3 x N = [u; v; has] */
            const double* cd = mxGetPr(prhs[4]);
            double* zc = (double*)mxCalloc(2 * (f.N > 0 ? f.N : 1), 8);
            uint8_t* has = (uint8_t*)mxCalloc(f.N > 0 ? f.N : 1, 1);
            int i;
            if ((int)mxGetM(prhs[4]) != 3 || (int)mxGetN(prhs[4]) != f.N) mexErrMsgIdAndTxt("ekfslam:arg", "cand must be 3 x N");
            for (i = 0; i < f.N; ++i) { zc[2 * i] = cd[3 * i]; zc[2 * i + 1] = cd[3 * i + 1]; has[i] = cd[3 * i + 2] != 0; }
            ck(ekfslam_upload_candidates(g_ctx, 0, 1, zc, has), "ekfslam_upload_candidates");
            ck(ekfslam_gate(g_ctx), "ekfslam_gate");
        }
        plhs[0] = feats_download(prhs[2], &f, 31);
    } else if (!strcmp(cmd, "ransac_hypotheses")) {                         /* mc/ransac_hypotheses.m */
        if (nrhs < 5) mexErrMsgIdAndTxt("ekfslam:arg", "ransac_hypotheses needs cam and the uniform stream u");
        ensure_ctx(f.N, f.n, prhs[1], prhs[3]);
        feats_upload(&f);
        state_upload(prhs[1], "x_k_km1", "p_k_km1", 1, f.n);
        ck(ekfslam_upload_uniforms(g_ctx, 0, 1, mxGetPr(prhs[4]), (int)mxGetNumberOfElements(prhs[4])), "ekfslam_upload_uniforms");
        ck(ekfslam_ransac(g_ctx), "ekfslam_ransac");
        plhs[0] = feats_download(prhs[2], &f, 16);
    } else if (!strcmp(cmd, "ekf_update_li_inliers") || !strcmp(cmd, "ekf_update_hi_inliers")) {  /* mc/update.m */
        const int li = cmd[11] == 'l';
        ensure_ctx(f.N, f.n, prhs[1], NULL);
        feats_upload(&f);
        state_upload(prhs[1], li ? "x_k_km1" : "x_k_k", li ? "p_k_km1" : "p_k_k", li ? 1 : 0, f.n);
        ck(ekfslam_hp(g_ctx, li ? EKFSLAM_F_LI : EKFSLAM_F_HI, 0), "ekfslam_hp");
        ck(ekfslam_update_masked(g_ctx, li ? EKFSLAM_F_LI : EKFSLAM_F_HI, li), "ekfslam_update_masked");
        plhs[0] = state_download(prhs[1], "x_k_k", "p_k_k", 0, f.n);
    } else if (!strcmp(cmd, "rescue_hi_inliers")) {                         /* mc/rescue_hi_inliers.m */
        if (nrhs < 4) mexErrMsgIdAndTxt("ekfslam:arg", "rescue_hi_inliers needs cam");
        ensure_ctx(f.N, f.n, prhs[1], prhs[3]);
        feats_upload(&f);
        state_upload(prhs[1], "x_k_k", "p_k_k", 0, f.n);
        ck(ekfslam_rescue(g_ctx), "ekfslam_rescue");
        plhs[0] = feats_download(prhs[2], &f, 1 | 2 | 16);
    } else {
        mexErrMsgIdAndTxt("ekfslam:arg", "unknown command '%s'", cmd);
    }
    mxFree(cmd);
}
