// EKF update of mc/update.m:3-32 for the features selected by a flag mask
// (mc/ekf_update_li_inliers.m:8-21, mc/ekf_update_hi_inliers.m:8-21), batched over filters.
//
// With G = H P (rows of the selected features, from k_hp) and S = G_sel H_sel' + I = L L':
//     K (z-h) = G_sel' inv(S) nu = W' y,      W = inv(L) G_sel,  y = inv(L) nu
//     K S K'  = G_sel' inv(S) G_sel = W' W
// so  x+ = x + W' y  and  P+ = P - W' W  (exactly symmetric, so the reference's
// 0.5 P + 0.5 P' (:14) is the identity here), followed by the quaternion normalisation
// Jacobian (:18-24) fused into the epilogue of the covariance downdate.
//
// Kernels: k_upd_S (select + stack S, nu) -> k_chol (blocked Cholesky, inv(L), y) ->
//          k_w (W = inv(L) G_sel, triangular GEMM; also x+, q normalisation, normJac) ->
//          k_downdate (P -= W'W on 64x64 tiles of the lower triangle, mirrored).
#include <cstdlib>
#include <cstring>
#include "model.cuh"
#include "tc_common.cuh"

int g_ekfslam_debug = 0;  // analysis knob, see k_downdate_ws
extern "C" void ekfslam_debug_flag(int f) { g_ekfslam_debug = f; }

#define NB 16

// ---------------------------------------------------------------------------------------
// select + S + nu.  One block per filter.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_upd_S(DevView v, int mask, int which_prior, int iter_nu) {
    const int b = blockIdx.x;
    const int N = v.N, ld = v.ld, kmax = v.kmax;
    const int nf = v.nfeat[b];
    const int tid = threadIdx.x;
    __shared__ int s_k;
    int* __restrict__ sel = v.sel + (size_t)b * N;
    if (tid == 0) {
        int cnt = 0;
        for (int i = 0; i < nf; ++i)
            if (v.ftype[b * N + i] != EKFSLAM_FEAT_NONE && (v.flags[(size_t)b * N + i] & mask)) sel[cnt++] = i;
        v.ksel[b] = cnt;
        s_k = cnt;
        if (mask & EKFSLAM_F_LI) v.stats[b].n_li = cnt;
        if (mask & EKFSLAM_F_HI) v.stats[b].n_hi = cnt;
    }
    __syncthreads();
    const int ns = s_k;
    const int k = 2 * ns;
    // innovation nu = z - h; for an iterated update (x holds the current iterate x_j, xp the prior x^-)
    // nu_j = z - h(x_j) - H_j (x^- - x_j).  Computed before x is reset to the prior below.
    {
        double* __restrict__ yv0 = v.yv + (size_t)b * kmax;
        const double* __restrict__ xj = v.x + (size_t)b * ld;
        const double* __restrict__ x0 = v.xp + (size_t)b * ld;
        for (int a = tid; a < k; a += blockDim.x) {
            const size_t t = (size_t)b * N + sel[a >> 1];
            double nu = v.z[2 * t + (a & 1)] - v.h[2 * t + (a & 1)];
            if (iter_nu) {
                const double* __restrict__ H = v.Hc + t * EKF_HSTRIDE + (a & 1) * EKF_HC;
                const int off = v.foff[t];
                const int w = (v.ftype[t] == EKFSLAM_FEAT_INVERSEDEPTH) ? 6 : 3;
                double s = 0.0;
                for (int c = 0; c < 7; ++c) s += H[c] * (x0[c] - xj[c]);
                for (int c = 0; c < w; ++c) s += H[7 + c] * (x0[off + c] - xj[off + c]);
                nu -= s;
            }
            yv0[a] = nu;
        }
    }
    __syncthreads();
    if (which_prior == 1) {  // x_k_k starts from x_k_km1 (also the pass-through of mc/update.m:28)
        const int n = v.nstate[b];
        for (int j = tid; j < n; j += blockDim.x) v.x[(size_t)b * ld + j] = v.xp[(size_t)b * ld + j];
    }
    if (k == 0) return;
    const double* __restrict__ G = v.G + (size_t)b * kmax * ld;
    double* __restrict__ S = v.Sb + (size_t)b * kmax * kmax;
    // lower triangle, feature-pair granularity: (fa >= fb) -> 2x2 block
    const int npair = ns * (ns + 1) / 2;
    for (int e = tid; e < npair; e += blockDim.x) {
        // unrank e -> (fa, fb), fa >= fb
        int fa = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
        while ((fa + 1) * (fa + 2) / 2 <= e) ++fa;
        while (fa * (fa + 1) / 2 > e) --fa;
        const int fb = e - fa * (fa + 1) / 2;
        const int ia = sel[fa], ib = sel[fb];
        const size_t tb = (size_t)b * N + ib;
        const double* __restrict__ H = v.Hc + tb * EKF_HSTRIDE;
        const int off = v.foff[tb];
        const int w = (v.ftype[tb] == EKFSLAM_FEAT_INVERSEDEPTH) ? 6 : 3;
        const double* __restrict__ g0 = G + (size_t)(2 * ia) * ld;
        const double* __restrict__ g1 = g0 + ld;
        double s00 = 0, s01 = 0, s10 = 0, s11 = 0;
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            const double a0 = g0[c], a1 = g1[c], h0 = H[c], h1 = H[EKF_HC + c];
            s00 += a0 * h0; s01 += a0 * h1; s10 += a1 * h0; s11 += a1 * h1;
        }
        for (int c = 0; c < w; ++c) {
            const double a0 = g0[off + c], a1 = g1[off + c], h0 = H[7 + c], h1 = H[EKF_HC + 7 + c];
            s00 += a0 * h0; s01 += a0 * h1; s10 += a1 * h0; s11 += a1 * h1;
        }
        if (fa == fb) { s00 += 1.0; s11 += 1.0; }  // R = eye(length(z)), mc/ekf_update_li_inliers.m:18
        S[(size_t)(2 * fa) * kmax + 2 * fb] = s00;
        S[(size_t)(2 * fa) * kmax + 2 * fb + 1] = s01;      // (for fa == fb this upper entry is never read)
        S[(size_t)(2 * fa + 1) * kmax + 2 * fb] = s10;
        S[(size_t)(2 * fa + 1) * kmax + 2 * fb + 1] = s11;
    }
}

// ---------------------------------------------------------------------------------------
// Blocked right-looking Cholesky S = L L' (lower, in place), X = inv(L), y <- X nu.
// One block per filter, panels of NB columns.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_chol(DevView v) {
    extern __shared__ double sm[];
    const int b = blockIdx.x;
    const int k = 2 * v.ksel[b];
    if (k == 0) return;
    const int kmax = v.kmax;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* __restrict__ S = v.Sb + (size_t)b * kmax * kmax;
    double* __restrict__ X = v.Li + (size_t)b * kmax * kmax;
    double* D = sm;                       // [NB][NB+1] diagonal block factor
    double* Di = D + NB * (NB + 1);       // [NB][NB+1] its inverse
    double* Pn = Di + NB * (NB + 1);      // [kmax][NB+1] panel below the diagonal block / row panel
    __shared__ int s_bad;
    if (tid == 0) s_bad = 0;

    for (int j0 = 0; j0 < k; j0 += NB) {
        const int nb = min(NB, k - j0);
        for (int e = tid; e < NB * NB; e += blockDim.x) {
            const int r = e / NB, c = e - r * NB;
            // rows >= nb are padded with the identity so that the register factorisation below is benign
            D[r * (NB + 1) + c] = (r < nb && c <= r) ? S[(size_t)(j0 + r) * kmax + j0 + c] : ((r >= nb && r == c) ? 1.0 : 0.0);
            Di[r * (NB + 1) + c] = 0.0;
        }
        __syncthreads();
        if (warp == 0) {
            // Cholesky of the NB x NB diagonal block in registers: lane i holds row i (lanes 16-31 mirror
            // lanes 0-15 so that every shuffle is full-warp); column c is broadcast by shuffles.
            const unsigned full_mask = 0xffffffffu;
            const int i = lane & (NB - 1);
            double a[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) a[c] = D[i * (NB + 1) + c];
            bool bad = false;
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                const double piv = __shfl_sync(full_mask, a[c], c);
                bad = bad || !(piv > 0.0);
                const double d = sqrt(piv);
                const double lic = (i == c) ? d : a[c] / d;  // rows above the diagonal carry don't-care values
                a[c] = lic;
#pragma unroll
                for (int j = c + 1; j < NB; ++j) {
                    const double ljc = __shfl_sync(full_mask, lic, j);
                    a[j] -= lic * ljc;
                }
            }
            if (bad && lane == 0) s_bad = 1;
            if (lane < NB) {
#pragma unroll
                for (int c = 0; c < NB; ++c) D[i * (NB + 1) + c] = (c <= i) ? a[c] : 0.0;
            }
            __syncwarp();
            // inverse of the triangular block: lane c solves column c by forward substitution; every lane
            // reads the same L entry at the same time (shared-memory broadcast)
            {
                const int c = i;
                double x[NB];
#pragma unroll
                for (int ii = 0; ii < NB; ++ii) {
                    double sacc = 0.0;
#pragma unroll
                    for (int t = 0; t < ii; ++t) sacc += D[ii * (NB + 1) + t] * x[t];
                    const double lii = D[ii * (NB + 1) + ii];
                    x[ii] = (ii == c) ? 1.0 / lii : ((ii > c) ? -sacc / lii : 0.0);
                }
                if (lane < NB) {
#pragma unroll
                    for (int ii = 0; ii < NB; ++ii) Di[ii * (NB + 1) + c] = x[ii];
                }
            }
        }
        __syncthreads();
        for (int e = tid; e < nb * nb; e += blockDim.x) {
            const int r = e / nb, c = e - r * nb;
            if (c <= r) {
                S[(size_t)(j0 + r) * kmax + j0 + c] = D[r * (NB + 1) + c];
                X[(size_t)(j0 + r) * kmax + j0 + c] = Di[r * (NB + 1) + c];
            }
        }
        // panel: L[i][j0+c] = sum_{t<=c} S[i][j0+t] * Di[c][t]
        const int i1 = j0 + nb;
        for (int i = i1 + tid; i < k; i += blockDim.x) {
            double row[NB];
#pragma unroll
            for (int t = 0; t < NB; ++t) row[t] = (t < nb) ? S[(size_t)i * kmax + j0 + t] : 0.0;
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                double s = 0.0;
#pragma unroll
                for (int t = 0; t <= c; ++t) s += row[t] * Di[c * (NB + 1) + t];
                if (c < nb) {
                    S[(size_t)i * kmax + j0 + c] = s;
                    Pn[(i - i1) * (NB + 1) + c] = s;
                }
            }
        }
        __syncthreads();
        // trailing update of the lower triangle: S[i][c] -= sum_t Pn[i][t] Pn[c][t]
        const int m = k - i1;
        if (m > 0) {
            const int mt = (m + 3) / 4;  // 4x4 micro tiles
            const int ntile = mt * (mt + 1) / 2;
            for (int e = tid; e < ntile; e += blockDim.x) {
                int ti = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
                while ((ti + 1) * (ti + 2) / 2 <= e) ++ti;
                while (ti * (ti + 1) / 2 > e) --ti;
                const int tj = e - ti * (ti + 1) / 2;
                double acc[4][4];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[a][c] = 0.0;
                for (int t = 0; t < nb; ++t) {
                    double ra[4], rc[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const int ia = ti * 4 + a, ic = tj * 4 + a;
                        ra[a] = (ia < m) ? Pn[ia * (NB + 1) + t] : 0.0;
                        rc[a] = (ic < m) ? Pn[ic * (NB + 1) + t] : 0.0;
                    }
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[a][c] += ra[a] * rc[c];
                }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int ia = ti * 4 + a, ic = tj * 4 + c;
                        if (ia < m && ic <= ia) S[(size_t)(i1 + ia) * kmax + i1 + ic] -= acc[a][c];
                    }
            }
        }
        __syncthreads();
    }

    // X = inv(L), block row by block row:  X[I][0:I0] = -Di_I * ( L[I][0:I0] * X[0:I0][0:I0] )
    for (int I0 = NB; I0 < k; I0 += NB) {
        const int nb = min(NB, k - I0);
        // stage the row panel L[I0..I0+nb)[0..I0) (transposed: Pn[t][r]) and Di of this block
        for (int e = tid; e < nb * I0; e += blockDim.x) {
            const int r = e / I0, t = e - r * I0;
            Pn[t * (NB + 1) + r] = S[(size_t)(I0 + r) * kmax + t];
        }
        for (int e = tid; e < NB * NB; e += blockDim.x) {
            const int r = e / NB, c = e - r * NB;
            Di[r * (NB + 1) + c] = (r < nb && c <= r) ? X[(size_t)(I0 + r) * kmax + I0 + c] : 0.0;
        }
        __syncthreads();
        for (int c = tid; c < I0; c += blockDim.x) {
            double y[NB];
#pragma unroll
            for (int r = 0; r < NB; ++r) y[r] = 0.0;
            for (int t = c; t < I0; ++t) {
                const double xv = X[(size_t)t * kmax + c];
#pragma unroll
                for (int r = 0; r < NB; ++r) y[r] += Pn[t * (NB + 1) + r] * xv;
            }
#pragma unroll
            for (int r = 0; r < NB; ++r) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j <= r; ++j) s += Di[r * (NB + 1) + j] * y[j];
                if (r < nb) X[(size_t)(I0 + r) * kmax + c] = -s;
            }
        }
        __syncthreads();
    }
    // explicit zeros above the diagonal of X: k_w streams X panels with cp.async and does not mask
    for (int e = tid; e < k * k; e += blockDim.x) {
        const int r = e / k, c = e - r * k;
        if (c > r) X[(size_t)r * kmax + c] = 0.0;
    }
    // y = X nu (in place through shared memory)
    double* nu = Pn;
    double* __restrict__ yv = v.yv + (size_t)b * kmax;
    for (int a = tid; a < k; a += blockDim.x) nu[a] = yv[a];
    __syncthreads();
    for (int a = tid; a < k; a += blockDim.x) {
        double s = 0.0;
        for (int t = 0; t <= a; ++t) s += X[(size_t)a * kmax + t] * nu[t];
        yv[a] = s;
    }
    __syncthreads();
    // cv = X' y = inv(S) nu  (the state update is x+ = x + G_sel' cv, accumulated inside k_w)
    for (int a = tid; a < k; a += blockDim.x) nu[a] = yv[a];
    __syncthreads();
    double* __restrict__ cv = v.cv + (size_t)b * kmax;
    for (int t = tid; t < k; t += blockDim.x) {
        double s = 0.0;
        for (int a = t; a < k; ++a) s += X[(size_t)a * kmax + t] * nu[a];
        cv[t] = s;
    }
    if (tid == 0 && s_bad) atomicOr(&v.stats[b].status, 2);
}

// ---------------------------------------------------------------------------------------
// W = X * G_sel  (X = inv(L) lower triangular k x k with an explicitly zeroed upper triangle,
// G_sel = the selected rows of G, k x n).  grid = (column tiles, row tiles, B).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 3) k_w(DevView v, int finalize) {
    extern __shared__ __align__(16) double dsm[];
    const int b = blockIdx.z;
    const int k = 2 * v.ksel[b];
    const int a0 = blockIdx.y * TM;
    if (a0 >= k) return;
    const int n = v.nstate[b];
    const int c0 = blockIdx.x * TM;
    if (c0 >= n) return;
    const int ld = v.ld, kmax = v.kmax;
    const double* __restrict__ X = v.Li + (size_t)b * kmax * kmax;
    const double* __restrict__ G = v.G + (size_t)b * kmax * ld;
    double* __restrict__ W = v.W + (size_t)b * kmax * ld;
    const int* __restrict__ sel = v.sel + (size_t)b * v.N;

    double* As = dsm;                           // [NSTAGE][64][APAD]   As[i][t] = X[a0+i][t0+t]
    double* Bs = dsm + NSTAGE * TM * APAD;      // [NSTAGE][TK][TPAD]   Bs[t][j] = G[row(t0+t)][c0+j]
    double* cs = Bs + NSTAGE * TK * TPAD;       // [kmax]               inv(S) nu
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wr = warp >> 2, wc = warp & 3, g = lane >> 2, q = lane & 3;
    // The last row tile streams every selected row of G for its 64 columns, so it also accumulates
    // the state update  x+ = x + G_sel' inv(S) nu  (mc/update.m:12) for those columns.
    const bool xrole = (a0 + TM >= k);
    if (xrole)
        for (int t = tid; t < k; t += blockDim.x) cs[t] = v.cv[(size_t)b * kmax + t];
    double xacc = 0.0;

    const int tend = min(k, a0 + TM);  // X[a][t] = 0 for t > a
    const int nk = (tend + TK - 1) / TK;
    auto load_stage = [&](int st, int t0) {
        double* as = As + st * TM * APAD;
        double* bs = Bs + st * TK * TPAD;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int ch = tid + 256 * j;          // 512 chunks of 2 doubles
            {   // A: 64 rows x 8 chunks
                const int r = ch >> 3, cc = (ch & 7) * 2;
                const bool ok = (a0 + r < k) && (t0 + cc < k);
                cp_async16(as + r * APAD + cc, ok ? X + (size_t)(a0 + r) * kmax + t0 + cc : X, ok ? 16 : 0);
            }
            {   // B: 16 rows x 32 chunks
                const int r = ch >> 5, cc = (ch & 31) * 2;
                const int tt = t0 + r;
                const bool ok = (tt < k) && (c0 + cc < ld);
                const double* src = G;
                if (ok) src = G + (size_t)(2 * sel[tt >> 1] + (tt & 1)) * ld + c0 + cc;
                cp_async16(bs + r * TPAD + cc, src, ok ? 16 : 0);
            }
        }
    };
    double acc[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int tmax_w = min(k, a0 + wr * 32 + 32);

#pragma unroll
    for (int st = 0; st < NSTAGE - 1; ++st) {
        if (st < nk) load_stage(st, st * TK);
        cp_async_commit();
    }
    for (int it = 0; it < nk; ++it) {
        cp_async_wait<NSTAGE - 2>();
        __syncthreads();
        if (it + NSTAGE - 1 < nk) load_stage((it + NSTAGE - 1) % NSTAGE, (it + NSTAGE - 1) * TK);
        cp_async_commit();
        const double* as = As + (it % NSTAGE) * TM * APAD;
        const double* bs = Bs + (it % NSTAGE) * TK * TPAD;
        if (xrole && tid < TM) {
            const int tmax = min(TK, k - it * TK);
            for (int t = 0; t < tmax; ++t) xacc += bs[t * TPAD + tid] * cs[it * TK + t];
        }
#pragma unroll
        for (int k4 = 0; k4 < TK / 4; ++k4) {
            const int tb = it * TK + k4 * 4;
            if (tb >= tmax_w) break;  // X is lower triangular: rows of this warp need t <= their own index
            double af[4], bf[2];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) af[mt] = as[(wr * 32 + mt * 8 + g) * APAD + k4 * 4 + q];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) bf[nt] = bs[(k4 * 4 + q) * TPAD + wc * 16 + nt * 8 + g];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
                // 8-row tile mt: rows r0..r0+7 exist if r0 < k and need t <= r0+7
                const int r0 = a0 + wr * 32 + mt * 8;
                if (r0 < k && tb <= r0 + 7) {
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) dmma(acc[mt][nt], af[mt], bf[nt]);
                }
            }
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const int a = a0 + wr * 32 + mt * 8 + g;
        if (a >= k) continue;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const int c = c0 + wc * 16 + nt * 8 + 2 * q;
            if (c < ld) {  // ld is even: c+1 < ld as well; the padding columns [n, ld) are kept zero
                double2 o;
                o.x = (c < n) ? acc[mt][nt][0] : 0.0;
                o.y = (c + 1 < n) ? acc[mt][nt][1] : 0.0;
                *reinterpret_cast<double2*>(W + (size_t)a * ld + c) = o;
            }
        }
    }
    if (xrole) {
        double* __restrict__ x = v.x + (size_t)b * ld;
        if (tid < TM && c0 + tid < n) x[c0 + tid] += xacc;
        if (c0 == 0 && finalize) {
            // this block owns state entries 0..63: normJac(q+) (mc/normJac.m) from the un-normalised
            // quaternion, then q+ <- q+/|q+|  (mc/update.m:18,24)
            __syncthreads();
            if (tid == 0) {
                const double r = x[3], qx = x[4], qy = x[5], qz = x[6];
                const double nn = r * r + qx * qx + qy * qy + qz * qz;
                const double sc = 1.0 / (nn * sqrt(nn));  // (.)^(-3/2)
                double* J = v.jn + (size_t)b * 16;
                J[0] = sc * (qx * qx + qy * qy + qz * qz); J[1] = sc * (-r * qx); J[2] = sc * (-r * qy); J[3] = sc * (-r * qz);
                J[4] = sc * (-qx * r); J[5] = sc * (r * r + qy * qy + qz * qz); J[6] = sc * (-qx * qy); J[7] = sc * (-qx * qz);
                J[8] = sc * (-qy * r); J[9] = sc * (-qy * qx); J[10] = sc * (r * r + qx * qx + qz * qz); J[11] = sc * (-qy * qz);
                J[12] = sc * (-qz * r); J[13] = sc * (-qz * qx); J[14] = sc * (-qz * qy); J[15] = sc * (r * r + qx * qx + qy * qy);
                const double nrm = sqrt(nn);
                x[3] = r / nrm; x[4] = qx / nrm; x[5] = qy / nrm; x[6] = qz / nrm;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// P <- Jn (P - W' W) Jn'  on 64x64 tiles of the lower triangle, each tile also stored transposed.
// grid = (lower-triangle tile index, B).  Jn = blkdiag(I3, normJac(q+), I): it only touches
// columns 3-6 (tiles of tile-column 0) and rows 3-6 (tile (0,0)).  The P tile is prefetched into
// the accumulator layout before the K loop so its HBM latency hides behind the DMMA work.
// ---------------------------------------------------------------------------------------
#ifndef DD_MINBLOCKS
#define DD_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(256, DD_MINBLOCKS) k_downdate(DevView v, const double* __restrict__ jn_all) {
    extern __shared__ __align__(16) double dsm[];
    const int b = blockIdx.y;
    const int k = 2 * v.ksel[b];
    if (k == 0) return;
    const int n = v.nstate[b];
    // unrank tile index -> (ti >= tj)
    const int e = blockIdx.x;
    int ti = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
    while ((ti + 1) * (ti + 2) / 2 <= e) ++ti;
    while (ti * (ti + 1) / 2 > e) --ti;
    const int tj = e - ti * (ti + 1) / 2;
    const int i0 = ti * TM, j0 = tj * TM;
    if (i0 >= n) return;
    const int ld = v.ld, kmax = v.kmax;
    const double* __restrict__ W = v.W + (size_t)b * kmax * ld;
    double* __restrict__ P = v.P + (size_t)b * v.nmax * ld;

    const bool diag = (ti == tj);
    double* As = dsm;                            // [NSTAGE][TK][TPAD]  As[t][i] = W[t0+t][i0+i]
    double* Bs = dsm + NSTAGE * TK * TPAD;       // [NSTAGE][TK][TPAD]  Bs[t][j] = W[t0+t][j0+j]
    double (*Ct)[TM + 1] = reinterpret_cast<double (*)[TM + 1]>(dsm);  // C tile, aliases the ring afterwards
    __shared__ double Jn[16];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wr = warp >> 2, wc = warp & 3, g = lane >> 2, q = lane & 3;
    if (tid < 16) Jn[tid] = jn_all[(size_t)b * 16 + tid];

    // prefetch this thread's part of the P tile (accumulator layout)
    double pf[4][2][2];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const int gi = i0 + wr * 32 + mt * 8 + g;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const int gj = j0 + wc * 16 + nt * 8 + 2 * q;
            double2 val = make_double2(0.0, 0.0);
            if (gi < n && gj < ld) val = *reinterpret_cast<const double2*>(P + (size_t)gi * ld + gj);
            pf[mt][nt][0] = val.x; pf[mt][nt][1] = val.y;
        }
    }
    const int nk = (k + TK - 1) / TK;
    // per-thread copy slots of a [TK][64] panel: rows r and r+8, 16-byte chunk cc (hoisted out of the loop)
    const int lr = tid >> 5, lcc = (tid & 31) * 2;
    const bool cola = (i0 + lcc < ld), colb = (j0 + lcc < ld);
    const double* __restrict__ wa = W + (size_t)lr * ld + i0 + lcc;
    const double* __restrict__ wb = W + (size_t)lr * ld + j0 + lcc;
    const int soff = lr * TPAD + lcc;
    auto load_stage = [&](int st, int t0) {
        double* as = As + st * TK * TPAD + soff;
        double* bs = Bs + st * TK * TPAD + soff;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int tt = t0 + lr + 8 * j;
            const bool oka = (tt < k) && cola;
            cp_async16(as + 8 * j * TPAD, oka ? wa + (size_t)(t0 + 8 * j) * ld : W, oka ? 16 : 0);
            if (!diag) {
                const bool okb = (tt < k) && colb;
                cp_async16(bs + 8 * j * TPAD, okb ? wb + (size_t)(t0 + 8 * j) * ld : W, okb ? 16 : 0);
            }
        }
    };
    double acc[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    // a warp whose 32x16 sub-tile lies strictly above the diagonal (diagonal tiles) or entirely
    // outside the n x n matrix has nothing to compute
    // ... and inside an active warp every 8x8 DMMA tile that is outside n x n, or strictly above the
    // diagonal of a diagonal tile, is skipped (bit mt*2+nt of onmask)
    unsigned onmask = 0;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const int r0 = i0 + wr * 32 + mt * 8, c0 = j0 + wc * 16 + nt * 8;
            const bool on = (r0 < n) && (c0 < n) && !(diag && c0 > r0 + 7);
            onmask |= (on ? 1u : 0u) << (mt * 2 + nt);
        }
    const bool warp_active = onmask != 0;

#pragma unroll
    for (int st = 0; st < NSTAGE - 1; ++st) {
        if (st < nk) load_stage(st, st * TK);
        cp_async_commit();
    }
    for (int it = 0; it < nk; ++it) {
        cp_async_wait<NSTAGE - 2>();
        __syncthreads();
        if (it + NSTAGE - 1 < nk) load_stage((it + NSTAGE - 1) % NSTAGE, (it + NSTAGE - 1) * TK);
        cp_async_commit();
        if (!warp_active) continue;
        const double* as = As + (it % NSTAGE) * TK * TPAD;
        const double* bs = diag ? as : Bs + (it % NSTAGE) * TK * TPAD;
        const int k4n = min(TK / 4, (k - it * TK + 3) >> 2);  // the tail panel stops at k (rounded to 4)
#pragma unroll
        for (int k4 = 0; k4 < TK / 4; ++k4) {
            if (k4 >= k4n) break;
            double af[4], bf[2];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) af[mt] = as[(k4 * 4 + q) * TPAD + wr * 32 + mt * 8 + g];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) bf[nt] = bs[(k4 * 4 + q) * TPAD + wc * 16 + nt * 8 + g];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
                    if (onmask & (1u << (mt * 2 + nt))) dmma(acc[mt][nt], af[mt], bf[nt]);
        }
    }
    cp_async_wait<0>();
    if (tj != 0) {
        // common case (no quaternion rows/columns in the tile): store P - W'W and its mirror image
        // straight from the accumulator fragments.  Per warp store, lanes with equal q cover 64 B runs.
        if (!warp_active) return;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const int gi = i0 + wr * 32 + mt * 8 + g;
            if (gi >= n) continue;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int gj = j0 + wc * 16 + nt * 8 + 2 * q;
                const double c0 = pf[mt][nt][0] - acc[mt][nt][0];
                const double c1 = pf[mt][nt][1] - acc[mt][nt][1];
                if (!diag) {
                    if (gj + 1 < n) {
                        *reinterpret_cast<double2*>(P + (size_t)gi * ld + gj) = make_double2(c0, c1);
                        P[(size_t)gj * ld + gi] = c0;
                        P[(size_t)(gj + 1) * ld + gi] = c1;
                    } else if (gj < n) {
                        P[(size_t)gi * ld + gj] = c0;
                        P[(size_t)gj * ld + gi] = c0;
                    }
                } else {
                    // diagonal tile: the lower triangle is authoritative, the upper one its mirror
                    if (gj <= gi && gj < n) {
                        P[(size_t)gi * ld + gj] = c0;
                        if (gj < gi) P[(size_t)gj * ld + gi] = c0;
                    }
                    if (gj + 1 <= gi && gj + 1 < n) {
                        P[(size_t)gi * ld + gj + 1] = c1;
                        if (gj + 1 < gi) P[(size_t)(gj + 1) * ld + gi] = c1;
                    }
                }
            }
        }
        return;
    }
    // tile column 0 holds state columns 3..6 (the quaternion): go through shared memory for the
    // normalisation Jacobian products
    __syncthreads();  // every warp is done with the ring: reuse it as the C tile
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const int r = wr * 32 + mt * 8 + g;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const int cc = wc * 16 + nt * 8 + 2 * q;
            const bool rin = (i0 + r < n);
            Ct[r][cc] = (rin && j0 + cc < n) ? pf[mt][nt][0] - acc[mt][nt][0] : 0.0;
            Ct[r][cc + 1] = (rin && j0 + cc + 1 < n) ? pf[mt][nt][1] - acc[mt][nt][1] : 0.0;
        }
    }
    __syncthreads();
    if (diag) {
        // tile (0,0): warps above the diagonal skipped their part -> mirror the lower triangle first
        for (int qd = tid; qd < TM * TM; qd += blockDim.x) {
            const int r = qd / TM, c = qd - r * TM;
            if (c > r) Ct[r][c] = Ct[c][r];
        }
        __syncthreads();
    }
    // columns 3..6 <- [c3 c4 c5 c6] * Jn'   (every row of the tile)
    if (tid < TM) {
        const int r = tid;
        const double c3 = Ct[r][3], c4 = Ct[r][4], c5 = Ct[r][5], c6 = Ct[r][6];
#pragma unroll
        for (int a = 0; a < 4; ++a)
            Ct[r][3 + a] = c3 * Jn[a * 4 + 0] + c4 * Jn[a * 4 + 1] + c5 * Jn[a * 4 + 2] + c6 * Jn[a * 4 + 3];
    }
    __syncthreads();
    if (diag) {
        // rows 3..6 <- Jn * [r3; r4; r5; r6]   (every column of tile (0,0))
        if (tid < TM) {
            const int c = tid;
            const double r3 = Ct[3][c], r4 = Ct[4][c], r5 = Ct[5][c], r6 = Ct[6][c];
#pragma unroll
            for (int a = 0; a < 4; ++a)
                Ct[3 + a][c] = Jn[a * 4 + 0] * r3 + Jn[a * 4 + 1] * r4 + Jn[a * 4 + 2] * r5 + Jn[a * 4 + 3] * r6;
        }
        __syncthreads();
        // the two one-sided products are symmetric only to rounding: lower -> upper once more
        for (int qd = tid; qd < TM * TM; qd += blockDim.x) {
            const int r = qd / TM, c = qd - r * TM;
            if (c > r) Ct[r][c] = Ct[c][r];
        }
        __syncthreads();
    }
    // store the tile and its mirror image
    for (int qd = tid; qd < TM * TM; qd += blockDim.x) {
        const int r = qd / TM, c = qd - r * TM;
        const int gi = i0 + r, gj = j0 + c;
        if (gi < n && gj < n) P[(size_t)gi * ld + gj] = Ct[r][c];
    }
    if (!diag) {
        for (int qd = tid; qd < TM * TM; qd += blockDim.x) {
            const int c = qd / TM, r = qd - c * TM;  // consecutive threads walk along r -> contiguous in P'
            const int gi = i0 + r, gj = j0 + c;
            if (gi < n && gj < n) P[(size_t)gj * ld + gi] = Ct[r][c];
        }
    }
}

// ---------------------------------------------------------------------------------------
// Persistent form of the downdate: a CTA walks a strided list of (filter, tile) pairs and keeps
// the cp.async ring running ACROSS tile boundaries, so the W panels of the next tile are already
// in flight while the current tile finishes and its P tile / stores overlap the DMMA work of the
// neighbours.  Same arithmetic, same summation order per output element as k_downdate.
// grid = (CTAs, 1); tile t -> filter t / T, lower-triangle tile t % T (T = nt(nt+1)/2).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) k_downdate_p(DevView v, const double* __restrict__ jn_all, int T,
                                                       long long total, int M) {
    extern __shared__ __align__(16) double dsm[];
    double* As = dsm;                            // [NSTAGE][TK][TPAD]
    double* Bs = dsm + NSTAGE * TK * TPAD;       // [NSTAGE][TK][TPAD]
    double* strip = Bs + NSTAGE * TK * TPAD;     // [64][9]: columns 0..7 of a tile-column-0 tile
    int2* meta = reinterpret_cast<int2*>(strip + TM * 9);          // [M]
    unsigned* lut = reinterpret_cast<unsigned*>(meta + M);          // [T]
    const int ld = v.ld, kmax = v.kmax;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wr = warp >> 2, wc = warp & 3, g = lane >> 2, q = lane & 3;
    const long long G = gridDim.x;
    for (int m = tid; m < M; m += blockDim.x) {
        const long long t = blockIdx.x + (long long)m * G;
        int2 kn = make_int2(0, 0);
        if (t < total) { const int b = (int)(t / T); kn = make_int2(2 * v.ksel[b], v.nstate[b]); }
        meta[m] = kn;
    }
    for (int e = tid; e < T; e += blockDim.x) {
        int ti = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
        while ((ti + 1) * (ti + 2) / 2 <= e) ++ti;
        while (ti * (ti + 1) / 2 > e) --ti;
        lut[e] = ((unsigned)ti << 16) | (unsigned)(e - ti * (ti + 1) / 2);
    }
    __syncthreads();
    // number of tiles this CTA really owns
    const int Mreal = (int)((total - blockIdx.x + G - 1) / G) < M ? (int)((total - blockIdx.x + G - 1) / G) : M;

    // loader cursor
    int lm = 0;
    DTile L = decode_tile(meta, lut, lm, Mreal, blockIdx.x + (long long)lm * G, T);
    while (lm < Mreal && L.nk == 0) { ++lm; L = decode_tile(meta, lut, lm, Mreal, blockIdx.x + (long long)lm * G, T); }
    int ls = 0;          // stage within the loader tile
    unsigned slot_l = 0; // ring slot counters (monotone)
    auto issue = [&]() {
        if (lm < Mreal) {
            const double* __restrict__ W = v.W + (size_t)L.b * kmax * ld;
            double* as = As + (slot_l % NSTAGE) * TK * TPAD;
            double* bs = Bs + (slot_l % NSTAGE) * TK * TPAD;
            const int t0 = ls * TK;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int ch = tid + 256 * j;
                const int r = ch >> 5, cc = (ch & 31) * 2;
                const int tt = t0 + r;
                const bool oka = (tt < L.k) && (L.i0 + cc < ld);
                cp_async16(as + r * TPAD + cc, oka ? W + (size_t)tt * ld + L.i0 + cc : W, oka ? 16 : 0);
                if (!L.diag) {
                    const bool okb = (tt < L.k) && (L.j0 + cc < ld);
                    cp_async16(bs + r * TPAD + cc, okb ? W + (size_t)tt * ld + L.j0 + cc : W, okb ? 16 : 0);
                }
            }
            if (++ls == L.nk) {
                ls = 0;
                do { ++lm; L = decode_tile(meta, lut, lm, Mreal, blockIdx.x + (long long)lm * G, T); } while (lm < Mreal && L.nk == 0);
            }
        }
        ++slot_l;
        cp_async_commit();
    };
#pragma unroll
    for (int st = 0; st < NSTAGE - 1; ++st) issue();

    unsigned slot_c = 0;
    for (int cm = 0; cm < Mreal; ++cm) {
        const DTile C = decode_tile(meta, lut, cm, Mreal, blockIdx.x + (long long)cm * G, T);
        if (C.nk == 0) continue;
        const int n = C.n, k = C.k, i0 = C.i0, j0 = C.j0;
        const bool diag = C.diag;
        double* __restrict__ P = v.P + (size_t)C.b * v.nmax * ld;
        // P tile -> accumulator layout (consumed after the K loop)
        double pf[4][2][2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const int gi = i0 + wr * 32 + mt * 8 + g;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int gj = j0 + wc * 16 + nt * 8 + 2 * q;
                double2 val = make_double2(0.0, 0.0);
                if (gi < n && gj < ld) val = *reinterpret_cast<const double2*>(P + (size_t)gi * ld + gj);
                pf[mt][nt][0] = val.x; pf[mt][nt][1] = val.y;
            }
        }
        unsigned onmask = 0;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int r0 = i0 + wr * 32 + mt * 8, c0 = j0 + wc * 16 + nt * 8;
                const bool on = (r0 < n) && (c0 < n) && !(diag && c0 > r0 + 7);
                onmask |= (on ? 1u : 0u) << (mt * 2 + nt);
            }
        double acc[4][2][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int it = 0; it < C.nk; ++it) {
            cp_async_wait<NSTAGE - 2>();
            __syncthreads();
            issue();
            const double* as = As + (slot_c % NSTAGE) * TK * TPAD;
            const double* bs = diag ? as : Bs + (slot_c % NSTAGE) * TK * TPAD;
            ++slot_c;
            if (onmask == 0) continue;
            const int k4n = min(TK / 4, (k - it * TK + 3) >> 2);
#pragma unroll
            for (int k4 = 0; k4 < TK / 4; ++k4) {
                if (k4 >= k4n) break;
                double af[4], bf[2];
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) af[mt] = as[(k4 * 4 + q) * TPAD + wr * 32 + mt * 8 + g];
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) bf[nt] = bs[(k4 * 4 + q) * TPAD + wc * 16 + nt * 8 + g];
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt)
                        if (onmask & (1u << (mt * 2 + nt))) dmma(acc[mt][nt], af[mt], bf[nt]);
            }
        }
        // ---- epilogue: C = P - W'W, stored with its mirror image straight from the fragments.
        // In tile column 0 the 8 leading columns (they hold the quaternion, state entries 3..6) are
        // routed through the shared strip for the normalisation Jacobian instead.
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const int gi = i0 + wr * 32 + mt * 8 + g;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                if (!(onmask & (1u << (mt * 2 + nt)))) continue;
                const int gj = j0 + wc * 16 + nt * 8 + 2 * q;
                const double c0 = pf[mt][nt][0] - acc[mt][nt][0];
                const double c1 = pf[mt][nt][1] - acc[mt][nt][1];
                if (C.col0 && wc == 0 && nt == 0) {
                    strip[(wr * 32 + mt * 8 + g) * 9 + 2 * q] = c0;
                    strip[(wr * 32 + mt * 8 + g) * 9 + 2 * q + 1] = c1;
                    continue;
                }
                if (gi >= n) continue;
                if (!diag) {
                    if (gj + 1 < n) {
                        *reinterpret_cast<double2*>(P + (size_t)gi * ld + gj) = make_double2(c0, c1);
                        P[(size_t)gj * ld + gi] = c0;
                        P[(size_t)(gj + 1) * ld + gi] = c1;
                    } else if (gj < n) {
                        P[(size_t)gi * ld + gj] = c0;
                        P[(size_t)gj * ld + gi] = c0;
                    }
                } else {
                    if (gj <= gi && gj < n) {
                        P[(size_t)gi * ld + gj] = c0;
                        if (gj < gi) P[(size_t)gj * ld + gi] = c0;
                    }
                    if (gj + 1 <= gi && gj + 1 < n) {
                        P[(size_t)gi * ld + gj + 1] = c1;
                        if (gj + 1 < gi) P[(size_t)(gj + 1) * ld + gi] = c1;
                    }
                }
            }
        }
        if (C.col0) {
            __syncthreads();
            const double* __restrict__ Jn = jn_all + (size_t)C.b * 16;
            if (diag) {
                // tile (0,0): the 8x8 corner holds the 4x4 quaternion block.  Complete the corner from its
                // lower triangle, apply Jn on both sides, and re-symmetrise (one thread; 8x8 is tiny).
                if (tid == 0) {
                    double Cn[8][8];
                    for (int r = 0; r < 8; ++r)
                        for (int c = 0; c < 8; ++c) Cn[r][c] = (c <= r) ? strip[r * 9 + c] : strip[c * 9 + r];
                    for (int r = 0; r < 8; ++r) {  // columns 3..6 <- row * Jn'
                        const double c3 = Cn[r][3], c4 = Cn[r][4], c5 = Cn[r][5], c6 = Cn[r][6];
                        for (int a = 0; a < 4; ++a)
                            Cn[r][3 + a] = c3 * Jn[a * 4 + 0] + c4 * Jn[a * 4 + 1] + c5 * Jn[a * 4 + 2] + c6 * Jn[a * 4 + 3];
                    }
                    for (int c = 0; c < 8; ++c) {  // rows 3..6 <- Jn * column
                        const double r3 = Cn[3][c], r4 = Cn[4][c], r5 = Cn[5][c], r6 = Cn[6][c];
                        for (int a = 0; a < 4; ++a)
                            Cn[3 + a][c] = Jn[a * 4 + 0] * r3 + Jn[a * 4 + 1] * r4 + Jn[a * 4 + 2] * r5 + Jn[a * 4 + 3] * r6;
                    }
                    for (int r = 0; r < 8; ++r)
                        for (int c = 0; c <= r; ++c) {
                            if (r < n && c < n) { P[(size_t)r * ld + c] = Cn[r][c]; P[(size_t)c * ld + r] = Cn[r][c]; }
                        }
                }
                // rows 8..63 of the strip: columns 3..6 <- row * Jn'
                if (tid >= 8 && tid < TM) {
                    const int r = tid;
                    if (i0 + r < n) {
                        double o[8];
                        for (int c = 0; c < 8; ++c) o[c] = strip[r * 9 + c];
                        const double c3 = o[3], c4 = o[4], c5 = o[5], c6 = o[6];
                        for (int a = 0; a < 4; ++a)
                            o[3 + a] = c3 * Jn[a * 4 + 0] + c4 * Jn[a * 4 + 1] + c5 * Jn[a * 4 + 2] + c6 * Jn[a * 4 + 3];
                        for (int c = 0; c < 8; ++c) { P[(size_t)(i0 + r) * ld + c] = o[c]; P[(size_t)c * ld + i0 + r] = o[c]; }
                    }
                }
            } else {
                if (tid < TM) {
                    const int r = tid;
                    if (i0 + r < n) {
                        double o[8];
                        for (int c = 0; c < 8; ++c) o[c] = strip[r * 9 + c];
                        const double c3 = o[3], c4 = o[4], c5 = o[5], c6 = o[6];
                        for (int a = 0; a < 4; ++a)
                            o[3 + a] = c3 * Jn[a * 4 + 0] + c4 * Jn[a * 4 + 1] + c5 * Jn[a * 4 + 2] + c6 * Jn[a * 4 + 3];
                        for (int c = 0; c < 8; ++c) { P[(size_t)(i0 + r) * ld + c] = o[c]; P[(size_t)c * ld + i0 + r] = o[c]; }
                    }
                }
            }
            __syncthreads();  // the strip is reused by the next tile-column-0 tile of this CTA
        }
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------
// Warp-specialised persistent downdate (the default).  One producer warp streams the W panels with
// bulk asynchronous copies (cp.async.bulk, SASS UBLKCP) that complete on mbarriers; eight consumer
// warps only wait, load fragments and issue DMMAs — no per-thread address arithmetic, no block
// barrier in the K loop, and the ring keeps running across tile boundaries.
//   full[s]  : 32 producer-lane arrivals + the bytes of the stage (complete_tx)
//   empty[s] : 8 consumer-warp arrivals
// ---------------------------------------------------------------------------------------
#define WS_STAGES 4
#define WS_CONSUMERS 8
#define WS_THREADS ((WS_CONSUMERS + 1) * 32)

__global__ void __launch_bounds__(WS_THREADS, 2) k_downdate_ws(DevView v, const double* __restrict__ jn_all, int T,
                                                               long long total, int M) {
    extern __shared__ __align__(16) double dsm[];
    double* As = dsm;                                   // [WS_STAGES][TK][TPAD]
    double* Bs = dsm + WS_STAGES * TK * TPAD;           // [WS_STAGES][TK][TPAD]
    double* strip = Bs + WS_STAGES * TK * TPAD;         // [64][9]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(strip + TM * 9);   // [WS_STAGES]
    unsigned long long* empty = full + WS_STAGES;                                        // [WS_STAGES]
    int2* meta = reinterpret_cast<int2*>(empty + WS_STAGES);                             // [M]
    unsigned* lut = reinterpret_cast<unsigned*>(meta + M);                               // [T]
    const int ld = v.ld, kmax = v.kmax;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long G = gridDim.x;
    for (int m = tid; m < M; m += blockDim.x) {
        const long long t = blockIdx.x + (long long)m * G;
        int2 kn = make_int2(0, 0);
        if (t < total) { const int b = (int)(t / T); kn = make_int2(2 * v.ksel[b], v.nstate[b]); }
        meta[m] = kn;
    }
    for (int e = tid; e < T; e += blockDim.x) {
        int ti = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
        while ((ti + 1) * (ti + 2) / 2 <= e) ++ti;
        while (ti * (ti + 1) / 2 > e) --ti;
        lut[e] = ((unsigned)ti << 16) | (unsigned)(e - ti * (ti + 1) / 2);
    }
    if (tid == 0) {
        for (int s2 = 0; s2 < WS_STAGES; ++s2) { mbar_init(full + s2, 32); mbar_init(empty + s2, WS_CONSUMERS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int own = (int)((total - blockIdx.x + G - 1) / G);
    const int Mreal = own < M ? own : M;

    if (warp == WS_CONSUMERS) {
        // ================= producer warp =================
        unsigned cnt = 0;
        const int pr = lane & 15;        // row of the [TK][64] panel this lane copies
        const bool isB = lane >= 16;     // lanes 0-15: A panel (rows of W at i0), 16-31: B panel (at j0)
        for (int m = 0; m < Mreal; ++m) {
            const DTile L = decode_tile(meta, lut, m, Mreal, blockIdx.x + (long long)m * G, T);
            if (L.nk == 0) continue;
            const double* __restrict__ W = v.W + (size_t)L.b * kmax * ld;
            const int c0 = isB ? L.j0 : L.i0;
            const unsigned rowbytes = (unsigned)(min(TM, ld - c0) * 8);
            const unsigned bytesA = (unsigned)(min(TM, ld - L.i0) * 8), bytesB = (unsigned)(min(TM, ld - L.j0) * 8);
            for (int st = 0; st < L.nk; ++st, ++cnt) {
                const unsigned slot = cnt % WS_STAGES, ph = (cnt / WS_STAGES) & 1u;
                mbar_wait(empty + slot, ph ^ 1u);
                const int t0 = st * TK;
                const int nvalid = min(TK, L.k - t0);
                double* dst = (isB ? Bs : As) + slot * TK * TPAD + pr * TPAD;
                if (lane == 0) {
                    mbar_arrive_expect_tx(full + slot, (unsigned)nvalid * (bytesA + (L.diag ? 0u : bytesB)));
                }
                const bool mine = !(isB && L.diag);
                if (mine && pr < nvalid) {
                    bulk_g2s(dst, W + (size_t)(t0 + pr) * ld + c0, rowbytes, full + slot);
                } else if (mine && pr < ((nvalid + 3) & ~3)) {
                    // rows between k and the next multiple of 4 are read by the last k4 step: zero them
                    for (int c = 0; c < TM; ++c) dst[c] = 0.0;
                }
                if (lane != 0) mbar_arrive(full + slot);
            }
        }
        return;
    }

    // ================= consumer warps =================
    const int wr = warp >> 2, wc = warp & 3, g = lane >> 2, q = lane & 3;
    unsigned cnt = 0;
    for (int cm = 0; cm < Mreal; ++cm) {
        const DTile C = decode_tile(meta, lut, cm, Mreal, blockIdx.x + (long long)cm * G, T);
        if (C.nk == 0) continue;
        const int n = C.n, k = C.k, i0 = C.i0, j0 = C.j0;
        const bool diag = C.diag;
        double* __restrict__ P = v.P + (size_t)C.b * v.nmax * ld;
        double pf[4][2][2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const int gi = i0 + wr * 32 + mt * 8 + g;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int gj = j0 + wc * 16 + nt * 8 + 2 * q;
                double2 val = make_double2(0.0, 0.0);
                if (gi < n && gj < ld) val = *reinterpret_cast<const double2*>(P + (size_t)gi * ld + gj);
                pf[mt][nt][0] = val.x; pf[mt][nt][1] = val.y;
            }
        }
        unsigned onmask = 0;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int r0 = i0 + wr * 32 + mt * 8, c0 = j0 + wc * 16 + nt * 8;
                const bool on = (r0 < n) && (c0 < n) && !(diag && c0 > r0 + 7);
                onmask |= (on ? 1u : 0u) << (mt * 2 + nt);
            }
        double acc[4][2][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int it = 0; it < C.nk; ++it, ++cnt) {
            const unsigned slot = cnt % WS_STAGES, ph = (cnt / WS_STAGES) & 1u;
            mbar_wait(full + slot, ph);
            const double* as = As + slot * TK * TPAD;
            const double* bs = diag ? as : Bs + slot * TK * TPAD;
            if (onmask != 0) {
                const int k4n = min(TK / 4, (k - it * TK + 3) >> 2);
#pragma unroll
                for (int k4 = 0; k4 < TK / 4; ++k4) {
                    if (k4 >= k4n) break;
                    double af[4], bf[2];
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) af[mt] = as[(k4 * 4 + q) * TPAD + wr * 32 + mt * 8 + g];
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) bf[nt] = bs[(k4 * 4 + q) * TPAD + wc * 16 + nt * 8 + g];
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt)
                            if (onmask & (1u << (mt * 2 + nt))) dmma(acc[mt][nt], af[mt], bf[nt]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + slot);
        }
        // ---- epilogue (same as k_downdate_p)
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const int gi = i0 + wr * 32 + mt * 8 + g;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                if (!(onmask & (1u << (mt * 2 + nt)))) continue;
                const int gj = j0 + wc * 16 + nt * 8 + 2 * q;
                const double c0 = pf[mt][nt][0] - acc[mt][nt][0];
                const double c1 = pf[mt][nt][1] - acc[mt][nt][1];
                if (C.col0 && wc == 0 && nt == 0) {
                    strip[(wr * 32 + mt * 8 + g) * 9 + 2 * q] = c0;
                    strip[(wr * 32 + mt * 8 + g) * 9 + 2 * q + 1] = c1;
                    continue;
                }
                if (gi >= n) continue;
                if (!diag) {
                    if (gj + 1 < n) {
                        *reinterpret_cast<double2*>(P + (size_t)gi * ld + gj) = make_double2(c0, c1);
                        P[(size_t)gj * ld + gi] = c0;
                        P[(size_t)(gj + 1) * ld + gi] = c1;
                    } else if (gj < n) {
                        P[(size_t)gi * ld + gj] = c0;
                        P[(size_t)gj * ld + gi] = c0;
                    }
                } else {
                    if (gj <= gi && gj < n) {
                        P[(size_t)gi * ld + gj] = c0;
                        if (gj < gi) P[(size_t)gj * ld + gi] = c0;
                    }
                    if (gj + 1 <= gi && gj + 1 < n) {
                        P[(size_t)gi * ld + gj + 1] = c1;
                        if (gj + 1 < gi) P[(size_t)(gj + 1) * ld + gi] = c1;
                    }
                }
            }
        }
        if (C.col0) {
            asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 consumer warps only
            const double* __restrict__ Jn = jn_all + (size_t)C.b * 16;
            if (diag && tid < 8) {
                // tile (0,0): the 8x8 corner (strip rows 0..7) holds the 4x4 quaternion block.  Threads 0..7
                // of warp 0 work in place in shared memory: complete the corner from its lower triangle,
                // apply Jn from the right (thread = row) and from the left (thread = column), store the
                // lower triangle and its mirror.
                const int r = tid;
                for (int c = r + 1; c < 8; ++c) strip[r * 9 + c] = strip[c * 9 + r];
                __syncwarp(0xffu);
                {
                    const double c3 = strip[r * 9 + 3], c4 = strip[r * 9 + 4], c5 = strip[r * 9 + 5], c6 = strip[r * 9 + 6];
                    for (int a = 0; a < 4; ++a)
                        strip[r * 9 + 3 + a] = c3 * Jn[a * 4 + 0] + c4 * Jn[a * 4 + 1] + c5 * Jn[a * 4 + 2] + c6 * Jn[a * 4 + 3];
                }
                __syncwarp(0xffu);
                {
                    const int c = tid;
                    const double r3 = strip[3 * 9 + c], r4 = strip[4 * 9 + c], r5 = strip[5 * 9 + c], r6 = strip[6 * 9 + c];
                    for (int a = 0; a < 4; ++a)
                        strip[(3 + a) * 9 + c] = Jn[a * 4 + 0] * r3 + Jn[a * 4 + 1] * r4 + Jn[a * 4 + 2] * r5 + Jn[a * 4 + 3] * r6;
                }
                __syncwarp(0xffu);
                for (int c = 0; c <= r; ++c)
                    if (r < n && c < n) { const double val = strip[r * 9 + c]; P[(size_t)r * ld + c] = val; P[(size_t)c * ld + r] = val; }
            }
            if (tid < TM && !(diag && tid < 8)) {
                const int r = tid;
                if (i0 + r < n) {
                    double o[8];
                    for (int c = 0; c < 8; ++c) o[c] = strip[r * 9 + c];
                    const double c3 = o[3], c4 = o[4], c5 = o[5], c6 = o[6];
                    for (int a = 0; a < 4; ++a)
                        o[3 + a] = c3 * Jn[a * 4 + 0] + c4 * Jn[a * 4 + 1] + c5 * Jn[a * 4 + 2] + c6 * Jn[a * 4 + 3];
                    for (int c = 0; c < 8; ++c) { P[(size_t)(i0 + r) * ld + c] = o[c]; P[(size_t)c * ld + i0 + r] = o[c]; }
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
    }
}

void launch_update(ekfslam_ctx* c, int mask, int which_prior, int flags) {
    // flags: 1 = iterated-update innovation (see k_upd_S), 2 = not the final iteration (no quaternion
    // normalisation, no covariance downdate)
    DevView& v = c->v;
    cudaStream_t st = c->stream;
    { KScope ks(c, KT_UPD_S); k_upd_S<<<v.B, 256, 0, st>>>(v, mask, which_prior, flags & 1); }
    const size_t chol_sm = sizeof(double) * (2 * NB * (NB + 1) + (size_t)v.kmax * (NB + 1));
    static size_t chol_cfg = 0;
    if (chol_sm > 48 * 1024 && chol_sm > chol_cfg) {
        cudaFuncSetAttribute(k_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_sm);
        chol_cfg = chol_sm;
    }
    { KScope ks(c, KT_CHOL); k_chol<<<v.B, 128, chol_sm, st>>>(v); }
    dim3 gw((v.nmax + TM - 1) / TM, (v.kmax + TM - 1) / TM, v.B);
    const size_t w_sm = sizeof(double) * (NSTAGE * (TM * APAD + TK * TPAD) + v.kmax);
    const size_t dd_sm = sizeof(double) * (size_t)max(2 * NSTAGE * TK * TPAD, TM * (TM + 1));
    static size_t attr_done = 0;
    if (attr_done < w_sm) {
        cudaFuncSetAttribute(k_w, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w_sm);
        cudaFuncSetAttribute(k_downdate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dd_sm);
        attr_done = w_sm;
    }
    { KScope ks(c, KT_W); k_w<<<gw, 256, w_sm, st>>>(v, (flags & 2) ? 0 : 1); }
    if (flags & 2) return;
    const int nt = (v.nmax + TM - 1) / TM;
    const int T = nt * (nt + 1) / 2;
    static int mode = -1, sms = 0;
    if (mode < 0) {
        const char* e = getenv("EKFSLAM_DOWNDATE");
        // default: warp-specialised 64x64 ("ws"); "ws128" = experimental 128x128 (slower: one CTA per SM
        // serialises K loop and epilogue); "persistent" / "tile" = cp.async 64x64
        mode = (e && !strcmp(e, "tile")) ? 0 : (e && !strcmp(e, "persistent")) ? 1 : (e && !strcmp(e, "ws128")) ? 3 : 2;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    }
    KScope ks(c, (mask & EKFSLAM_F_HI) ? KT_DOWNDATE_HI : KT_DOWNDATE);
    if (mode == 3 && launch_downdate128(c, sms)) return;
    if (mode == 0) {
        dim3 gd(T, v.B);
        k_downdate<<<gd, 256, dd_sm, st>>>(v, v.jn);
    } else if (mode >= 2) {
        const long long total = (long long)T * v.B;
        const long long ctas = total < (long long)sms * 2 ? total : (long long)sms * 2;
        const int M = (int)((total + ctas - 1) / ctas);
        const size_t ws_sm = sizeof(double) * (2 * WS_STAGES * TK * TPAD + TM * 9) + sizeof(unsigned long long) * 2 * WS_STAGES +
                             sizeof(int2) * M + sizeof(unsigned) * T;
        static size_t ws_cfg = 0;
        if (ws_sm > ws_cfg) {
            cudaFuncSetAttribute(k_downdate_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ws_sm);
            ws_cfg = ws_sm;
        }
        k_downdate_ws<<<(unsigned)ctas, WS_THREADS, ws_sm, st>>>(v, v.jn, T, total, M);
    } else {
        const long long total = (long long)T * v.B;
        const long long ctas = total < (long long)sms * 2 ? total : (long long)sms * 2;
        const int M = (int)((total + ctas - 1) / ctas);
        const size_t ddp_sm = sizeof(double) * (2 * NSTAGE * TK * TPAD + TM * 9) + sizeof(int2) * M + sizeof(unsigned) * T;
        static size_t ddp_cfg = 0;
        if (ddp_sm > ddp_cfg) {
            cudaFuncSetAttribute(k_downdate_p, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ddp_sm);
            ddp_cfg = ddp_sm;
        }
        k_downdate_p<<<(unsigned)ctas, 256, ddp_sm, st>>>(v, v.jn, T, total, M);
    }
}
