/*
 * ekf_oracle.c — plain-C restatement of the reference filter step (TEST INFRASTRUCTURE ONLY).
 *
 * Second, independent CPU statement of the MATLAB hot path of diwakar-vsingh/EKF-SLAM
 * (matlab_code/*.m, cited as mc/<file>:<lines>), written "as the reference writes it":
 * K = P H' inv(S) with an explicit general inverse, P - K S K', 0.5P + 0.5P'.  It is used
 *   (1) by tests/ to cross-check the numpy oracle (oracle/ekf_oracle.py) and the CUDA path at
 *       sizes where the numpy loops are slow, and
 *   (2) by bench.py as the multi-threaded CPU baseline (`cpu_baseline.kind = "port"`,
 *       `--impl reference`): OpenMP over independent filters, all host cores.
 * The product package never links or loads this file.  Parity status: pinned by the reference's
 * features_information.mat for h/H/S and, since round 2, by the reference's OWN EXECUTION for
 * RANSAC / update / rescue / the Cartesian model (tests/golden/ref_*.npz, produced by running the
 * unmodified matlab_code/*.m through oracle/mref; tests/test_oracle_ref.py::test_c_oracle_matches_reference_execution).
 *
 * Build: make -C oracle   ->  oracle/libekf_oracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define F_HAS_H 1
#define F_HAS_Z 2
#define F_IC 4
#define F_LI 8
#define F_HI 16

typedef struct {
    double k1, k2, Cx, Cy, f, dx, dy;
    int nRows, nCols;
} cam_t;

typedef struct {
    double std_a, std_alpha, std_z, delta_t, chi2, p_free;
    int max_hyp, fixed_hyp;
} prm_t;

/* mc/initialize_cam.m:3-25 */
static void default_cam(cam_t* c) {
    const double d = 0.0112;
    c->k1 = 6.333e-2; c->k2 = 1.390e-2; c->nRows = 240; c->nCols = 320;
    c->Cx = 1.7945 / d; c->Cy = 1.4433 / d; c->f = 2.1735; c->dx = d; c->dy = d;
}

/* mc/q2r.m:8-10 */
static void q2r(const double* q, double R[9]) {
    const double r = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = r * r + x * x - y * y - z * z; R[1] = 2 * (x * y - r * z); R[2] = 2 * (z * x + r * y);
    R[3] = 2 * (x * y + r * z); R[4] = r * r - x * x + y * y - z * z; R[5] = 2 * (y * z - r * x);
    R[6] = 2 * (z * x - r * y); R[7] = 2 * (y * z + r * x); R[8] = r * r - x * x - y * y + z * z;
}

/* inv() of a 3x3 (the reference calls MATLAB inv(q2r(q)) — mc/calculate_Hi_inverse_depth.m:46,109,144) */
static void inv3(const double A[9], double B[9]) {
    const double c0 = A[4] * A[8] - A[5] * A[7], c1 = A[5] * A[6] - A[3] * A[8], c2 = A[3] * A[7] - A[4] * A[6];
    const double det = A[0] * c0 + A[1] * c1 + A[2] * c2;
    B[0] = c0 / det; B[1] = (A[2] * A[7] - A[1] * A[8]) / det; B[2] = (A[1] * A[5] - A[2] * A[4]) / det;
    B[3] = c1 / det; B[4] = (A[0] * A[8] - A[2] * A[6]) / det; B[5] = (A[2] * A[3] - A[0] * A[5]) / det;
    B[6] = c2 / det; B[7] = (A[1] * A[6] - A[0] * A[7]) / det; B[8] = (A[0] * A[4] - A[1] * A[3]) / det;
}

/* mc/distort_fm.m:22-38 */
static void distort_fm(const cam_t* c, double uu, double vu, double* ud, double* vd) {
    const double xu = (uu - c->Cx) * c->dx, yu = (vu - c->Cy) * c->dy;
    const double ru = sqrt(xu * xu + yu * yu);
    double rd = ru / (1 + c->k1 * ru * ru + c->k2 * ru * ru * ru * ru);
    for (int k = 0; k < 10; ++k) {
        const double f = rd + c->k1 * rd * rd * rd + c->k2 * rd * rd * rd * rd * rd - ru;
        const double fp = 1 + 3 * c->k1 * rd * rd + 5 * c->k2 * rd * rd * rd * rd;
        rd = rd - f / fp;
    }
    const double D = 1 + c->k1 * rd * rd + c->k2 * rd * rd * rd * rd;
    *ud = (xu / D) / c->dx + c->Cx;
    *vd = (yu / D) / c->dy + c->Cy;
}

/* mc/jacob_undistor_fm.m:20-34 */
static void jacob_undistor_fm(const cam_t* c, double ud, double vd, double J[4]) {
    const double xd = (ud - c->Cx) * c->dx, yd = (vd - c->Cy) * c->dy;
    const double rd2 = xd * xd + yd * yd, rd4 = rd2 * rd2;
    J[0] = (1 + c->k1 * rd2 + c->k2 * rd4) + (ud - c->Cx) * (c->k1 + 2 * c->k2 * rd2) * (2 * (ud - c->Cx) * c->dx * c->dx);
    J[3] = (1 + c->k1 * rd2 + c->k2 * rd4) + (vd - c->Cy) * (c->k1 + 2 * c->k2 * rd2) * (2 * (vd - c->Cy) * c->dy * c->dy);
    J[1] = (ud - c->Cx) * (c->k1 + 2 * c->k2 * rd2) * (2 * (vd - c->Cy) * c->dy * c->dy);
    J[2] = (vd - c->Cy) * (c->k1 + 2 * c->k2 * rd2) * (2 * (ud - c->Cx) * c->dx * c->dx);
}

/* ray of a feature before rotation: (y - r) rho + m(theta, phi)  or  y - r */
static void ray(const double* xv, const double* y, int type, double d[3]) {
    if (type == 1) {
        const double cphi = cos(y[4]);
        d[0] = (y[0] - xv[0]) * y[5] + cphi * sin(y[3]);
        d[1] = (y[1] - xv[1]) * y[5] - sin(y[4]);
        d[2] = (y[2] - xv[2]) * y[5] + cphi * cos(y[3]);
    } else {
        d[0] = y[0] - xv[0]; d[1] = y[1] - xv[1]; d[2] = y[2] - xv[2];
    }
}

/* mc/hi_inverse_depth.m:7-57 / mc/hi_cartesian.m:7-49.  Returns 1 if visible. */
static int hi(const cam_t* c, const double* xv, const double* y, int type, double h[2]) {
    double R[9], M[9], d[3], hrl[3];
    q2r(xv + 3, R);
    ray(xv, y, type, d);
    if (type == 1) {  /* r_cw = r_wc' */
        for (int i = 0; i < 3; ++i) hrl[i] = R[0 + i] * d[0] + R[3 + i] * d[1] + R[6 + i] * d[2];
    } else {          /* r_cw = inv(r_wc) */
        inv3(R, M);
        for (int i = 0; i < 3; ++i) hrl[i] = M[3 * i] * d[0] + M[3 * i + 1] * d[1] + M[3 * i + 2] * d[2];
    }
    const double ax = atan2(hrl[0], hrl[2]) * 180 / M_PI, ay = atan2(hrl[1], hrl[2]) * 180 / M_PI;
    if (ax < -60 || ax > 60 || ay < -60 || ay > 60) return 0;
    const double uu = c->Cx + (hrl[0] / hrl[2]) * c->f * (1 / c->dx);  /* mc/hu.m:12-13 */
    const double vu = c->Cy + (hrl[1] / hrl[2]) * c->f * (1 / c->dy);
    double ud, vd;
    distort_fm(c, uu, vu, &ud, &vd);
    if (ud > 0 && ud < c->nCols && vd > 0 && vd < c->nRows) { h[0] = ud; h[1] = vd; return 1; }
    return 0;
}

/* mc/calculate_Hi_inverse_depth.m / mc/calculate_Hi_cartesian.m -> compact 2x13 */
static void calc_H(const cam_t* c, const double* xv, const double* y, int type, const double* zi, double Hc[26]) {
    double R[9], Ri[9], d[3], hc[3], Ju[4], A[4];
    q2r(xv + 3, R);
    inv3(R, Ri);
    ray(xv, y, type, d);
    for (int i = 0; i < 3; ++i) hc[i] = Ri[3 * i] * d[0] + Ri[3 * i + 1] * d[1] + Ri[3 * i + 2] * d[2];
    const double f = c->f, ku = 1 / c->dx, kv = 1 / c->dy;
    const double a[6] = {f * ku / hc[2], 0, -hc[0] * f * ku / (hc[2] * hc[2]), 0, f * kv / hc[2], -hc[1] * f * kv / (hc[2] * hc[2])};
    jacob_undistor_fm(c, zi[0], zi[1], Ju);
    const double det = Ju[0] * Ju[3] - Ju[1] * Ju[2];
    A[0] = Ju[3] / det; A[1] = -Ju[1] / det; A[2] = -Ju[2] / det; A[3] = Ju[0] / det;
    double D[6];  /* dh_dhrl = dhd_dhu * dhu_dhrl */
    for (int r = 0; r < 2; ++r)
        for (int k = 0; k < 3; ++k) D[3 * r + k] = A[2 * r] * a[k] + A[2 * r + 1] * a[3 + k];
    const double rho = (type == 1) ? y[5] : 1.0;
    /* dhrl_drw = -inv(R)*rho */
    for (int r = 0; r < 2; ++r)
        for (int k = 0; k < 3; ++k) {
            double s = 0;
            for (int m = 0; m < 3; ++m) s += D[3 * r + m] * (-(Ri[3 * m + k]) * rho);
            Hc[13 * r + k] = s;
        }
    /* dhrl_dqwr = dRq_times_a_by_dq(qconj(q), d) * diag(1,-1,-1,-1)   (mc/dRq_times_a_by_dq.m) */
    {
        const double q0 = xv[3], qx = -xv[4], qy = -xv[5], qz = -xv[6];
        const double dR[4][9] = {{2 * q0, -2 * qz, 2 * qy, 2 * qz, 2 * q0, -2 * qx, -2 * qy, 2 * qx, 2 * q0},
                                 {2 * qx, 2 * qy, 2 * qz, 2 * qy, -2 * qx, -2 * q0, 2 * qz, 2 * q0, -2 * qx},
                                 {-2 * qy, 2 * qx, 2 * q0, 2 * qx, 2 * qy, 2 * qz, -2 * q0, 2 * qz, -2 * qy},
                                 {-2 * qz, -2 * q0, 2 * qx, 2 * q0, -2 * qz, 2 * qy, 2 * qx, 2 * qy, 2 * qz}};
        for (int k = 0; k < 4; ++k) {
            double t[3];
            for (int i = 0; i < 3; ++i) t[i] = dR[k][3 * i] * d[0] + dR[k][3 * i + 1] * d[1] + dR[k][3 * i + 2] * d[2];
            const double sg = (k == 0) ? 1.0 : -1.0;
            for (int r = 0; r < 2; ++r) Hc[13 * r + 3 + k] = (D[3 * r] * t[0] + D[3 * r + 1] * t[1] + D[3 * r + 2] * t[2]) * sg;
        }
    }
    for (int r = 0; r < 2; ++r)
        for (int k = 7; k < 13; ++k) Hc[13 * r + k] = 0;
    if (type == 1) {
        const double th = y[3], ph = y[4];
        const double dmt[3] = {cos(ph) * cos(th), 0, -cos(ph) * sin(th)};
        const double dmp[3] = {-sin(ph) * sin(th), -cos(ph), -sin(ph) * cos(th)};
        const double w[3] = {y[0] - xv[0], y[1] - xv[1], y[2] - xv[2]};
        double a6[18];  /* 3x6: [rho*Ri, Ri*dmt, Ri*dmp, Ri*w] */
        for (int i = 0; i < 3; ++i) {
            for (int k = 0; k < 3; ++k) a6[6 * i + k] = rho * Ri[3 * i + k];
            a6[6 * i + 3] = Ri[3 * i] * dmt[0] + Ri[3 * i + 1] * dmt[1] + Ri[3 * i + 2] * dmt[2];
            a6[6 * i + 4] = Ri[3 * i] * dmp[0] + Ri[3 * i + 1] * dmp[1] + Ri[3 * i + 2] * dmp[2];
            a6[6 * i + 5] = Ri[3 * i] * w[0] + Ri[3 * i + 1] * w[1] + Ri[3 * i + 2] * w[2];
        }
        for (int r = 0; r < 2; ++r)
            for (int k = 0; k < 6; ++k)
                Hc[13 * r + 7 + k] = D[3 * r] * a6[k] + D[3 * r + 1] * a6[6 + k] + D[3 * r + 2] * a6[12 + k];
    } else {
        for (int r = 0; r < 2; ++r)
            for (int k = 0; k < 3; ++k)
                Hc[13 * r + 7 + k] = D[3 * r] * Ri[k] + D[3 * r + 1] * Ri[3 + k] + D[3 * r + 2] * Ri[6 + k];
    }
}

/* general inverse (Gauss-Jordan, partial pivoting) — stands in for MATLAB inv(S) (mc/update.m:9) */
static int inv_general(int k, double* A /* k x k, destroyed */, double* Ai) {
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) Ai[i * k + j] = (i == j);
    for (int c = 0; c < k; ++c) {
        int p = c;
        for (int r = c + 1; r < k; ++r)
            if (fabs(A[r * k + c]) > fabs(A[p * k + c])) p = r;
        if (A[p * k + c] == 0.0) return -1;
        if (p != c)
            for (int j = 0; j < k; ++j) {
                double t = A[c * k + j]; A[c * k + j] = A[p * k + j]; A[p * k + j] = t;
                t = Ai[c * k + j]; Ai[c * k + j] = Ai[p * k + j]; Ai[p * k + j] = t;
            }
        const double piv = 1.0 / A[c * k + c];
        for (int j = 0; j < k; ++j) { A[c * k + j] *= piv; Ai[c * k + j] *= piv; }
        for (int r = 0; r < k; ++r) {
            if (r == c) continue;
            const double f = A[r * k + c];
            if (f == 0.0) continue;
            for (int j = 0; j < k; ++j) { A[r * k + j] -= f * A[c * k + j]; Ai[r * k + j] -= f * Ai[c * k + j]; }
        }
    }
    return 0;
}

/* per-filter workspace */
typedef struct {
    int N, nmax;
    double *h, *Hc, *S, *z, *PHt, *Sk, *Ski, *K, *KS, *xi;
    int *off, *sel;
    uint8_t* fl;
} ws_t;

static ws_t* ws_new(int N, int nmax) {
    ws_t* w = (ws_t*)calloc(1, sizeof(ws_t));
    const int k = 2 * N;
    w->N = N; w->nmax = nmax;
    w->h = (double*)calloc(2 * N, 8); w->Hc = (double*)calloc(26 * N, 8); w->S = (double*)calloc(4 * N, 8);
    w->z = (double*)calloc(2 * N, 8);
    w->PHt = (double*)calloc((size_t)nmax * k, 8); w->Sk = (double*)calloc((size_t)k * k, 8);
    w->Ski = (double*)calloc((size_t)k * k, 8); w->K = (double*)calloc((size_t)nmax * k, 8);
    w->KS = (double*)calloc((size_t)nmax * k, 8); w->xi = (double*)calloc(nmax, 8);
    w->off = (int*)calloc(N, 4); w->sel = (int*)calloc(N, 4); w->fl = (uint8_t*)calloc(N, 1);
    return w;
}
static void ws_free(ws_t* w) {
    free(w->h); free(w->Hc); free(w->S); free(w->z); free(w->PHt); free(w->Sk); free(w->Ski); free(w->K);
    free(w->KS); free(w->xi); free(w->off); free(w->sel); free(w->fl); free(w);
}

/* mc/predict_state_and_covariance.m:3-27 (fv, dfv_by_dxv, func_Q), in place on x, P (row-major n x n) */
static void predict(const prm_t* p, int n, double* x, double* P) {
    const double dt = p->delta_t;
    const double q0 = x[3], qx = x[4], qy = x[5], qz = x[6], wx = x[10], wy = x[11], wz = x[12];
    const double ax = wx * dt, ay = wy * dt, az = wz * dt;
    const double theta = sqrt(ax * ax + ay * ay + az * az);
    double p0 = 1, px = 0, py = 0, pz = 0;
    if (!(theta < 2.220446049250313e-16)) {  /* mc/v2q.m:11-15 */
        p0 = cos(theta / 2); px = sin(theta / 2) * (ax / theta); py = sin(theta / 2) * (ay / theta); pz = sin(theta / 2) * (az / theta);
    }
    double F[13][13], Q[13][13], G[13][6];
    memset(F, 0, sizeof(F)); memset(Q, 0, sizeof(Q)); memset(G, 0, sizeof(G));
    for (int i = 0; i < 13; ++i) F[i][i] = 1;
    const double Fqq[4][4] = {{p0, -px, -py, -pz}, {px, p0, pz, -py}, {py, -pz, p0, px}, {pz, py, -px, p0}};  /* dq3_by_dq2 */
    const double om = sqrt(wx * wx + wy * wy + wz * wz), w[3] = {wx, wy, wz};
    double dq[4][3];  /* mc/dqomegadt_by_domega.m */
    for (int a = 0; a < 3; ++a) {
        dq[0][a] = (-dt / 2.0) * (w[a] / om) * sin(om * dt / 2.0);
        for (int c = 0; c < 3; ++c)
            dq[1 + a][c] = (a == c) ? (dt / 2.0) * w[a] * w[a] / (om * om) * cos(om * dt / 2.0) +
                                          (1.0 / om) * (1.0 - w[a] * w[a] / (om * om)) * sin(om * dt / 2.0)
                                    : (w[a] * w[c] / (om * om)) * ((dt / 2.0) * cos(om * dt / 2.0) - (1.0 / om) * sin(om * dt / 2.0));
    }
    const double L[4][4] = {{q0, -qx, -qy, -qz}, {qx, q0, -qz, qy}, {qy, qz, q0, -qx}, {qz, -qy, qx, q0}};  /* dq3_by_dq1 */
    double M[4][3];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 3; ++c) { double s = 0; for (int k = 0; k < 4; ++k) s += L[r][k] * dq[k][c]; M[r][c] = s; }
    for (int r = 0; r < 4; ++r) {
        for (int c = 0; c < 4; ++c) F[3 + r][3 + c] = Fqq[r][c];
        for (int c = 0; c < 3; ++c) { F[3 + r][10 + c] = M[r][c]; G[3 + r][3 + c] = M[r][c]; }
    }
    for (int r = 0; r < 3; ++r) { F[r][7 + r] = dt; G[r][r] = dt; G[7 + r][r] = 1; G[10 + r][3 + r] = 1; }
    const double Pn[6] = {(p->std_a * dt) * (p->std_a * dt), (p->std_a * dt) * (p->std_a * dt), (p->std_a * dt) * (p->std_a * dt),
                          (p->std_alpha * dt) * (p->std_alpha * dt), (p->std_alpha * dt) * (p->std_alpha * dt),
                          (p->std_alpha * dt) * (p->std_alpha * dt)};
    for (int r = 0; r < 13; ++r)
        for (int c = 0; c < 13; ++c) { double s = 0; for (int k = 0; k < 6; ++k) s += G[r][k] * Pn[k] * G[c][k]; Q[r][c] = s; }
    /* state */
    double xn[7];
    xn[0] = x[0] + x[7] * dt; xn[1] = x[1] + x[8] * dt; xn[2] = x[2] + x[9] * dt;
    xn[3] = q0 * p0 - (qx * px + qy * py + qz * pz);
    xn[4] = (q0 * px + p0 * qx) + (qy * pz - qz * py);
    xn[5] = (q0 * py + p0 * qy) + (qz * px - qx * pz);
    xn[6] = (q0 * pz + p0 * qz) + (qx * py - qy * px);
    for (int i = 0; i < 7; ++i) x[i] = xn[i];
    /* covariance: [F Pxx F' + Q, F Pxy; Pyx F', Pyy] */
    double T[13][13], Pxx[13][13];
    for (int r = 0; r < 13; ++r) for (int c = 0; c < 13; ++c) Pxx[r][c] = P[(size_t)r * n + c];
    for (int r = 0; r < 13; ++r) for (int c = 0; c < 13; ++c) { double s = 0; for (int k = 0; k < 13; ++k) s += F[r][k] * Pxx[k][c]; T[r][c] = s; }
    for (int r = 0; r < 13; ++r) for (int c = 0; c < 13; ++c) { double s = 0; for (int k = 0; k < 13; ++k) s += T[r][k] * F[c][k]; P[(size_t)r * n + c] = s + Q[r][c]; }
    for (int j = 13; j < n; ++j) {
        double col[13], row[13];
        for (int r = 0; r < 13; ++r) { col[r] = P[(size_t)r * n + j]; row[r] = P[(size_t)j * n + r]; }
        for (int r = 0; r < 13; ++r) {
            double s = 0, t = 0;
            for (int k = 0; k < 13; ++k) { s += F[r][k] * col[k]; t += row[k] * F[r][k]; }
            P[(size_t)r * n + j] = s; P[(size_t)j * n + r] = t;
        }
    }
}

/* mc/predict_camera_measurements.m + mc/calculate_derivatives.m at state x */
static void measure(const cam_t* c, ws_t* w, int nf, const uint8_t* type, const double* x) {
    for (int i = 0; i < nf; ++i) {
        double hh[2];
        if (hi(c, x, x + w->off[i], type[i], hh)) { w->h[2 * i] = hh[0]; w->h[2 * i + 1] = hh[1]; w->fl[i] |= F_HAS_H; }
        if (w->fl[i] & F_HAS_H) calc_H(c, x, x + w->off[i], type[i], w->h + 2 * i, w->Hc + 26 * i);
    }
}

/* S_i = H_i P H_i' (+R) with sparse H_i (13 columns) */
static void innov(const ws_t* w, int n, const double* P, int i, int type, int addR, double S[4]) {
    int cols[13];
    const int wd = (type == 1) ? 6 : 3;
    for (int k = 0; k < 7; ++k) cols[k] = k;
    for (int k = 0; k < wd; ++k) cols[7 + k] = w->off[i] + k;
    const double* H = w->Hc + 26 * i;
    for (int r = 0; r < 2; ++r)
        for (int s = 0; s < 2; ++s) {
            double acc = 0;
            for (int a = 0; a < 7 + wd; ++a) {
                double t = 0;
                for (int b = 0; b < 7 + wd; ++b) t += P[(size_t)cols[a] * n + cols[b]] * H[13 * s + b];
                acc += H[13 * r + a] * t;
            }
            S[2 * r + s] = acc + ((addR && r == s) ? 1.0 : 0.0);
        }
}

/* mc/compute_hypothesis_support_fast.m:9-45, 57-84 */
static int support(const cam_t* c, const ws_t* w, int nf, const uint8_t* type, const double* xi, double thr, uint8_t* inl) {
    double R[9];
    q2r(xi + 3, R);
    int cnt = 0;
    const double fku = c->f * (1 / c->dx), fkv = c->f * (1 / c->dy);
    for (int i = 0; i < nf; ++i) {
        inl[i] = 0;
        if (!(w->fl[i] & F_HAS_Z)) continue;
        double d[3], hc[3];
        ray(xi, xi + w->off[i], type[i], d);
        for (int k = 0; k < 3; ++k) hc[k] = R[0 + k] * d[0] + R[3 + k] * d[1] + R[6 + k] * d[2];
        const double uu = fku * (hc[0] / hc[2]) + c->Cx, vu = fkv * (hc[1] / hc[2]) + c->Cy;
        double ud, vd;
        distort_fm(c, uu, vu, &ud, &vd);
        const double n0 = w->z[2 * i] - ud, n1 = w->z[2 * i + 1] - vd;
        if (sqrt(n0 * n0 + n1 * n1) < thr) { inl[i] = 1; ++cnt; }
    }
    return cnt;
}

/* mc/ransac_hypotheses.m:3-47 */
static int ransac(const cam_t* c, const prm_t* p, ws_t* w, int n, int nf, const uint8_t* type, const double* x,
                  const double* P, const double* u, int n_u, int* status) {
    int nic = 0;
    int* ic = w->sel;
    for (int i = 0; i < nf; ++i) if (w->fl[i] & F_IC) ic[nic++] = i;
    if (nic == 0) return 0;
    uint8_t* inl = (uint8_t*)malloc(nf);
    int n_hyp = p->max_hyp, best = 0, iters = 0;
    const int n_loop = p->fixed_hyp > 0 ? p->fixed_hyp : p->max_hyp;
    for (int i = 1; i <= n_loop; ++i) {
        if (i > n_u) { *status |= 1; break; }
        int r = (int)floor(u[i - 1] * nic);
        if (r > nic - 1) r = nic - 1;
        const int pos = ic[r];
        iters = i;
        /* K = P Hi' inv(S); xi = x + K (zi - hi) */
        const int wd = (type[pos] == 1) ? 6 : 3;
        const double* H = w->Hc + 26 * pos;
        double S[4];
        innov(w, n, P, pos, type[pos], 1, S);
        const double det = S[0] * S[3] - S[1] * S[2];
        const double Si[4] = {S[3] / det, -S[1] / det, -S[2] / det, S[0] / det};
        const double nu0 = w->z[2 * pos] - w->h[2 * pos], nu1 = w->z[2 * pos + 1] - w->h[2 * pos + 1];
        for (int j = 0; j < n; ++j) {
            double ph0 = 0, ph1 = 0;
            for (int a = 0; a < 7; ++a) { ph0 += P[(size_t)j * n + a] * H[a]; ph1 += P[(size_t)j * n + a] * H[13 + a]; }
            for (int a = 0; a < wd; ++a) { ph0 += P[(size_t)j * n + w->off[pos] + a] * H[7 + a]; ph1 += P[(size_t)j * n + w->off[pos] + a] * H[20 + a]; }
            const double k0 = ph0 * Si[0] + ph1 * Si[2], k1 = ph0 * Si[1] + ph1 * Si[3];
            w->xi[j] = x[j] + (k0 * nu0 + k1 * nu1);
        }
        const int sup = support(c, w, nf, type, w->xi, p->std_z, inl);
        if (sup > best) {
            best = sup;
            for (int f = 0; f < nf; ++f)
                if (w->fl[f] & F_HAS_Z) w->fl[f] = inl[f] ? (w->fl[f] | F_LI) : (w->fl[f] & ~F_LI);
            if (p->fixed_hyp <= 0) {
                const double eps = 1 - ((double)sup / (double)nic);
                const double den = 1 - (1 - eps);
                double v = (den <= 0.0) ? 0.0 : ceil(log(1 - p->p_free) / log(den));
                n_hyp = (v > 2e9) ? 2000000000 : (int)v;
                if (n_hyp == 0) break;
            }
        }
        if (p->fixed_hyp <= 0 && i > n_hyp) break;
    }
    free(inl);
    return iters;
}

/* mc/update.m:3-32 for the features with flag `mask`, as the reference writes it */
static int update(ws_t* w, int n, int nf, const uint8_t* type, double* x, double* P, int mask) {
    int ns = 0;
    for (int i = 0; i < nf; ++i) if (w->fl[i] & mask) w->sel[ns++] = i;
    if (ns == 0) return 0;
    const int k = 2 * ns;
    double* PHt = w->PHt;  /* n x k */
    for (int j = 0; j < n; ++j)
        for (int a = 0; a < ns; ++a) {
            const int f = w->sel[a], wd = (type[f] == 1) ? 6 : 3;
            const double* H = w->Hc + 26 * f;
            double s0 = 0, s1 = 0;
            for (int m = 0; m < 7; ++m) { s0 += P[(size_t)j * n + m] * H[m]; s1 += P[(size_t)j * n + m] * H[13 + m]; }
            for (int m = 0; m < wd; ++m) { s0 += P[(size_t)j * n + w->off[f] + m] * H[7 + m]; s1 += P[(size_t)j * n + w->off[f] + m] * H[20 + m]; }
            PHt[(size_t)j * k + 2 * a] = s0; PHt[(size_t)j * k + 2 * a + 1] = s1;
        }
    double* S = w->Sk;  /* S = H PHt + I */
    for (int a = 0; a < ns; ++a) {
        const int f = w->sel[a], wd = (type[f] == 1) ? 6 : 3;
        const double* H = w->Hc + 26 * f;
        for (int r = 0; r < 2; ++r)
            for (int c = 0; c < k; ++c) {
                double s = 0;
                for (int m = 0; m < 7; ++m) s += H[13 * r + m] * PHt[(size_t)m * k + c];
                for (int m = 0; m < wd; ++m) s += H[13 * r + 7 + m] * PHt[(size_t)(w->off[f] + m) * k + c];
                S[(size_t)(2 * a + r) * k + c] = s + ((2 * a + r == c) ? 1.0 : 0.0);
            }
    }
    double* Si = w->Ski;
    memcpy(w->KS, S, (size_t)k * k * 8);  /* scratch copy destroyed by the inversion */
    if (inv_general(k, w->KS, Si)) return -1;
    double* K = w->K;  /* K = PHt Si */
    for (int j = 0; j < n; ++j) {
        double* kr = K + (size_t)j * k;
        for (int c = 0; c < k; ++c) kr[c] = 0;
        for (int m = 0; m < k; ++m) {
            const double pv = PHt[(size_t)j * k + m];
            const double* sr = Si + (size_t)m * k;
            for (int c = 0; c < k; ++c) kr[c] += pv * sr[c];
        }
    }
    for (int j = 0; j < n; ++j) {  /* x += K (z - h) */
        double s = 0;
        for (int a = 0; a < ns; ++a) {
            const int f = w->sel[a];
            s += K[(size_t)j * k + 2 * a] * (w->z[2 * f] - w->h[2 * f]) + K[(size_t)j * k + 2 * a + 1] * (w->z[2 * f + 1] - w->h[2 * f + 1]);
        }
        x[j] += s;
    }
    double* KS = w->KS;  /* KS = K S */
    for (int j = 0; j < n; ++j) {
        double* kr = KS + (size_t)j * k;
        for (int c = 0; c < k; ++c) kr[c] = 0;
        for (int m = 0; m < k; ++m) {
            const double kv = K[(size_t)j * k + m];
            const double* sr = S + (size_t)m * k;
            for (int c = 0; c < k; ++c) kr[c] += kv * sr[c];
        }
    }
    /* P -= KS K'  (K transposed once so that the innermost loop is unit-stride and vectorises
       without re-association; every P[i][j] still accumulates its k products in order m = 0..k-1) */
    double* Kt = w->PHt;  /* PHt is dead from here on: reuse as K' (k x n) */
    for (int j = 0; j < n; ++j)
        for (int m = 0; m < k; ++m) Kt[(size_t)m * n + j] = K[(size_t)j * k + m];
    for (int i = 0; i < n; ++i) {
        const double* a = KS + (size_t)i * k;
        double* pr = P + (size_t)i * n;
        for (int m = 0; m < k; ++m) {
            const double av = a[m];
            const double* kt = Kt + (size_t)m * n;
            for (int j = 0; j < n; ++j) pr[j] -= av * kt[j];
        }
    }
    for (int i = 0; i < n; ++i)  /* 0.5 P + 0.5 P' */
        for (int j = i + 1; j < n; ++j) {
            const double v = 0.5 * P[(size_t)i * n + j] + 0.5 * P[(size_t)j * n + i];
            P[(size_t)i * n + j] = v; P[(size_t)j * n + i] = v;
        }
    /* mc/normJac.m + mc/update.m:18-24 */
    const double r = x[3], qx = x[4], qy = x[5], qz = x[6];
    const double sc = pow(r * r + qx * qx + qy * qy + qz * qz, -1.5);
    const double J[4][4] = {{sc * (qx * qx + qy * qy + qz * qz), sc * (-r * qx), sc * (-r * qy), sc * (-r * qz)},
                            {sc * (-qx * r), sc * (r * r + qy * qy + qz * qz), sc * (-qx * qy), sc * (-qx * qz)},
                            {sc * (-qy * r), sc * (-qy * qx), sc * (r * r + qx * qx + qz * qz), sc * (-qy * qz)},
                            {sc * (-qz * r), sc * (-qz * qx), sc * (-qz * qy), sc * (r * r + qx * qx + qy * qy)}};
    for (int j = 0; j < n; ++j) {  /* rows 3..6 <- J * rows ; done on the pre-image, then columns */
        double v[4];
        for (int a = 0; a < 4; ++a) v[a] = P[(size_t)(3 + a) * n + j];
        for (int a = 0; a < 4; ++a) P[(size_t)(3 + a) * n + j] = J[a][0] * v[0] + J[a][1] * v[1] + J[a][2] * v[2] + J[a][3] * v[3];
    }
    for (int i = 0; i < n; ++i) {
        double v[4];
        for (int a = 0; a < 4; ++a) v[a] = P[(size_t)i * n + 3 + a];
        for (int a = 0; a < 4; ++a) P[(size_t)i * n + 3 + a] = v[0] * J[a][0] + v[1] * J[a][1] + v[2] * J[a][2] + v[3] * J[a][3];
    }
    const double nrm = sqrt(r * r + qx * qx + qy * qy + qz * qz);
    x[3] = r / nrm; x[4] = qx / nrm; x[5] = qy / nrm; x[6] = qz / nrm;
    return ns;
}

/*
 * One reference filter step (mc/mono_slam.m:56-74 without takeImage, matcher = gating rule of
 * mc/matching.m:16,38) for B independent filters, OpenMP-parallel over filters.
 *   x [B][nmax], P [B][nmax][nmax] (row-major, in/out), type [B][N], nfeat [B], zc [B][N][2],
 *   has [B][N], u [B][n_u]; outputs flags [B][N] (HAS_H|HAS_Z|IC|LI|HI), stats [B][4] =
 *   {ransac iterations, n_li, n_hi, status}.  par[8] = std_a, std_alpha, std_z, delta_t, chi2,
 *   p_free, max_hyp, fixed_hyp.  Returns 0.
 */
int ekf_oracle_step_batch(int B, int N, int nmax, double* x, double* P, const uint8_t* type, const int32_t* nfeat,
                          const double* zc, const uint8_t* has, const double* u, int n_u, const double* par,
                          uint8_t* flags, int32_t* stats, int nthreads) {
    cam_t cam;
    default_cam(&cam);
    prm_t prm = {par[0], par[1], par[2], par[3], par[4], par[5], (int)par[6], (int)par[7]};
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        ws_t* w = ws_new(N, nmax);
        double* Pc = (double*)malloc((size_t)nmax * nmax * 8);
#pragma omp for schedule(dynamic, 1)
        for (int b = 0; b < B; ++b) {
            const uint8_t* ty = type + (size_t)b * N;
            const int nf = nfeat[b];
            int n = 13;
            for (int i = 0; i < nf; ++i) { w->off[i] = n; n += (ty[i] == 1) ? 6 : 3; w->fl[i] = 0; }
            double* xb = x + (size_t)b * nmax;
            double* Pb = P + (size_t)b * nmax * nmax;
            /* compact to n x n */
            for (int i = 0; i < n; ++i) memcpy(Pc + (size_t)i * n, Pb + (size_t)i * nmax, (size_t)n * 8);
            int status = 0;
            predict(&prm, n, xb, Pc);
            measure(&cam, w, nf, ty, xb);
            for (int i = 0; i < nf; ++i) {
                if (!(w->fl[i] & F_HAS_H)) continue;
                innov(w, n, Pc, i, ty[i], 1, w->S + 4 * i);
                if (!has[(size_t)b * N + i]) continue;
                const double* S = w->S + 4 * i;
                const double tr = S[0] + S[3], disc = sqrt((S[0] - S[3]) * (S[0] - S[3]) + 4 * S[1] * S[2]);
                const double n0 = zc[2 * ((size_t)b * N + i)] - w->h[2 * i], n1 = zc[2 * ((size_t)b * N + i) + 1] - w->h[2 * i + 1];
                const double det = S[0] * S[3] - S[1] * S[2];
                const double d2 = (n0 * (S[3] * n0 - S[1] * n1) + n1 * (-S[2] * n0 + S[0] * n1)) / det;
                if (0.5 * (tr + disc) < 100.0 && d2 < prm.chi2) {
                    w->z[2 * i] = zc[2 * ((size_t)b * N + i)]; w->z[2 * i + 1] = zc[2 * ((size_t)b * N + i) + 1];
                    w->fl[i] |= F_HAS_Z | F_IC;
                }
            }
            const int iters = ransac(&cam, &prm, w, n, nf, ty, xb, Pc, u + (size_t)b * n_u, n_u, &status);
            const int nli = update(w, n, nf, ty, xb, Pc, F_LI);
            measure(&cam, w, nf, ty, xb);  /* mc/rescue_hi_inliers.m:6-7 */
            for (int i = 0; i < nf; ++i) {
                if ((w->fl[i] & F_IC) && !(w->fl[i] & F_LI)) {
                    double S[4];
                    innov(w, n, Pc, i, ty[i], 0, S);
                    const double n0 = w->z[2 * i] - w->h[2 * i], n1 = w->z[2 * i + 1] - w->h[2 * i + 1];
                    const double det = S[0] * S[3] - S[1] * S[2];
                    const double d2 = (n0 * (S[3] * n0 - S[1] * n1) + n1 * (-S[2] * n0 + S[0] * n1)) / det;
                    if (d2 < prm.chi2) w->fl[i] |= F_HI; else w->fl[i] &= ~F_HI;
                }
            }
            const int nhi = update(w, n, nf, ty, xb, Pc, F_HI);
            for (int i = 0; i < n; ++i) memcpy(Pb + (size_t)i * nmax, Pc + (size_t)i * n, (size_t)n * 8);
            for (int i = 0; i < nf; ++i) flags[(size_t)b * N + i] = w->fl[i];
            for (int i = nf; i < N; ++i) flags[(size_t)b * N + i] = 0;
            stats[4 * b] = iters; stats[4 * b + 1] = nli < 0 ? 0 : nli; stats[4 * b + 2] = nhi < 0 ? 0 : nhi;
            stats[4 * b + 3] = status | ((nli < 0 || nhi < 0) ? 2 : 0);
        }
        free(Pc);
        ws_free(w);
    }
    return 0;
}

int ekf_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
