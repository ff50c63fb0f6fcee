// Per-frame bookkeeping, state/covariance prediction, measurement prediction + analytic
// Jacobians, G = H*P row products and per-feature innovation covariances.
#include <cstdlib>
#include "model.cuh"
#include "tc_common.cuh"

// ---------------------------------------------------------------------------------------
// mc/update_features_info.m:4-18 — one thread per (filter, feature)
// ---------------------------------------------------------------------------------------
__global__ void k_begin_frame(DevView v) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= v.B * v.N) return;
    const int b = t / v.N, i = t - b * v.N;
    if (i >= v.nfeat[b]) return;
    const uint8_t f = v.flags[t];
    if (f & EKFSLAM_F_HAS_H) v.counters[2 * t] += 1;
    if (f & (EKFSLAM_F_LI | EKFSLAM_F_HI)) v.counters[2 * t + 1] += 1;
    v.flags[t] = f & EKFSLAM_F_CAND;  // a staged candidate survives the reset, everything else clears
}

void launch_begin_frame(ekfslam_ctx* c) {
    const int tot = c->v.B * c->v.N;
    KScope ks(c, KT_BEGIN_FRAME);
    k_begin_frame<<<(tot + 255) / 256, 256, 0, c->stream>>>(c->v);
}

// ---------------------------------------------------------------------------------------
// mc/predict_state_and_covariance.m:3-27.  One block per filter.  The covariance is updated in
// place: F differs from the identity only in rows 0-6 (r and q), so only rows/columns 0-6 of P
// change:  P[0:13,j] <- F P[0:13,j]  and  Pxx <- F Pxx F' + Q.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 3) k_predict(DevView v, ekfslam_params prm) {
    const int b = blockIdx.x;
    const int n = v.nstate[b];
    const int ld = v.ld;
    double* __restrict__ P = v.P + (size_t)b * v.nmax * ld;
    const double* __restrict__ x = v.x + (size_t)b * ld;
    double* __restrict__ xp = v.xp + (size_t)b * ld;

    __shared__ double F[13][13];
    __shared__ double Q[13][13];
    __shared__ double Pxx[13][13];
    __shared__ double T[13][13];
    __shared__ double Fqq[4][4], Fqw[4][3];

    const int tid = threadIdx.x;
    // The first cross-covariance column of every thread is requested BEFORE thread 0 builds F and Q, and inside the sweep
    // the next column is requested before the current one is transformed (ncu, round 2: 60 % of the kernel's samples were
    // long-scoreboard stalls on one DRAM round trip per column and thread, one after the other).
    // The columns are read from rows 0..12 (P is exactly symmetric in memory, so P[r][j] == P[j][r]): consecutive threads read
    // consecutive addresses.  Reading the 13 entries from row j instead made every load instruction touch 32 different lines
    // (ncu, round 2: 0.29 ms for 0.64 GB of DRAM traffic - the kernel was bound by L1 wavefronts, not by DRAM).
    double c[13];
    const int lane = tid & 31, warp = tid >> 5;
    __shared__ double tr[8][32][9];   // per-warp transpose of the 7 changed entries (+ entry 7) for the row-wise mirror stores
    int j = 13 + tid;
    if (j < n) {
#pragma unroll
        for (int r = 0; r < 13; ++r) c[r] = P[(size_t)r * ld + j];
    }
    for (int e = tid; e < 169; e += blockDim.x) {
        const int r = e / 13, cc = e - r * 13;
        Pxx[r][cc] = P[(size_t)r * ld + cc];
        F[r][cc] = (r == cc) ? 1.0 : 0.0;
        Q[r][cc] = 0.0;
    }
    __syncthreads();
    if (tid == 0) {
        const double dt = prm.delta_t;
        const double q0 = x[3], qx = x[4], qy = x[5], qz = x[6];
        const double wx = x[10], wy = x[11], wz = x[12];
        // mc/v2q.m:10-16 with quaternions(v_n, theta) = [cos(theta/2), sin(theta/2) v_n]
        const double ax = wx * dt, ay = wy * dt, az = wz * dt;
        const double theta = sqrt(ax * ax + ay * ay + az * az);
        double p0 = 1.0, px = 0.0, py = 0.0, pz = 0.0;
        if (!(theta < 2.220446049250313e-16)) {
            const double sh = sin(theta / 2.0);
            p0 = cos(theta / 2.0);
            px = sh * (ax / theta); py = sh * (ay / theta); pz = sh * (az / theta);
        }
        // mc/fv.m:43-46, mc/qprod.m:8
        xp[0] = x[0] + x[7] * dt; xp[1] = x[1] + x[8] * dt; xp[2] = x[2] + x[9] * dt;
        xp[3] = q0 * p0 - (qx * px + qy * py + qz * pz);
        xp[4] = (q0 * px + p0 * qx) + (qy * pz - qz * py);
        xp[5] = (q0 * py + p0 * qy) + (qz * px - qx * pz);
        xp[6] = (q0 * pz + p0 * qz) + (qx * py - qy * px);
        for (int k = 7; k < 13; ++k) xp[k] = x[k];
        // mc/dfv_by_dxv.m:9 — dq3_by_dq2(qwt)
        Fqq[0][0] = p0; Fqq[0][1] = -px; Fqq[0][2] = -py; Fqq[0][3] = -pz;
        Fqq[1][0] = px; Fqq[1][1] = p0;  Fqq[1][2] = pz;  Fqq[1][3] = -py;
        Fqq[2][0] = py; Fqq[2][1] = -pz; Fqq[2][2] = p0;  Fqq[2][3] = px;
        Fqq[3][0] = pz; Fqq[3][1] = py;  Fqq[3][2] = -px; Fqq[3][3] = p0;
        // mc/dqomegadt_by_domega.m:6-48   (giving this half of the serial work to a second warp was measured twice: no change)
        const double om = sqrt(wx * wx + wy * wy + wz * wz);
        const double sn = sin(om * dt / 2.0), cs = cos(om * dt / 2.0);
        const double w[3] = {wx, wy, wz};
        double dq[4][3];
        // The reference divides by |omega| (mc/dqomegadt_by_domega.m:33,37-48) and only avoids 0/0 because it seeds
        // omega with 1e-15 (mc/initialize_x_and_p.m:6).  An EXACTLY zero angular velocity (caller-supplied state) gets
        // the analytic limit instead of NaNs: dq0/domega = 0, dq_i/domega_i = dt/2, cross terms 0.
        const bool om_zero = !(om > 0.0);
        for (int a = 0; a < 3; ++a) {
            if (om_zero) {
                dq[0][a] = 0.0;
                for (int c2 = 0; c2 < 3; ++c2) dq[1 + a][c2] = (a == c2) ? dt / 2.0 : 0.0;
                continue;
            }
            dq[0][a] = (-dt / 2.0) * (w[a] / om) * sn;
            for (int c2 = 0; c2 < 3; ++c2) {
                if (a == c2)
                    dq[1 + a][c2] = (dt / 2.0) * w[a] * w[a] / (om * om) * cs +
                                    (1.0 / om) * (1.0 - w[a] * w[a] / (om * om)) * sn;
                else
                    dq[1 + a][c2] = (w[a] * w[c2] / (om * om)) * ((dt / 2.0) * cs - (1.0 / om) * sn);
            }
        }
        // dq3_by_dq1(qOld) (missing in the reference; left-multiplication matrix of qOld)
        const double L[4][4] = {{q0, -qx, -qy, -qz}, {qx, q0, -qz, qy}, {qy, qz, q0, -qx}, {qz, -qy, qx, q0}};
        for (int r = 0; r < 4; ++r)
            for (int c2 = 0; c2 < 3; ++c2) {
                double s = 0.0;
                for (int k = 0; k < 4; ++k) s += L[r][k] * dq[k][c2];
                Fqw[r][c2] = s;
            }
        for (int r = 0; r < 4; ++r) {
            for (int c2 = 0; c2 < 4; ++c2) F[3 + r][3 + c2] = Fqq[r][c2];
            for (int c2 = 0; c2 < 3; ++c2) F[3 + r][10 + c2] = Fqw[r][c2];
        }
        for (int r = 0; r < 3; ++r) F[r][7 + r] = dt;
        // mc/func_Q.m:15-28: Q = G Pn G', G = [dt*I 0; 0 M; I 0; 0 I], Pn = diag(la*I3, aa*I3)
        const double la = (prm.std_a * dt) * (prm.std_a * dt);
        const double aa = (prm.std_alpha * dt) * (prm.std_alpha * dt);
        for (int r = 0; r < 3; ++r) {
            Q[r][r] = dt * la * dt;
            Q[r][7 + r] = dt * la;
            Q[7 + r][r] = la * dt;
            Q[7 + r][7 + r] = la;
            Q[10 + r][10 + r] = aa;
        }
        for (int r = 0; r < 4; ++r) {
            for (int c2 = 0; c2 < 4; ++c2) {
                double s = 0.0;
                for (int k = 0; k < 3; ++k) s += Fqw[r][k] * aa * Fqw[c2][k];
                Q[3 + r][3 + c2] = s;
            }
            for (int c2 = 0; c2 < 3; ++c2) {
                Q[3 + r][10 + c2] = Fqw[r][c2] * aa;
                Q[10 + c2][3 + r] = aa * Fqw[r][c2];
            }
        }
    }
    // features are static: x_k_km1(14:end) = x_k_k(14:end)
    for (int jj = 13 + tid; jj < n; jj += blockDim.x) xp[jj] = x[jj];
    __syncthreads();

    // cross-covariance panel
    while (j - lane < n) {   // warp-uniform: the mirror stores below are a warp-wide exchange
        const int jn = j + blockDim.x;
        double cn[13];
        if (jn < n) {
#pragma unroll
            for (int r = 0; r < 13; ++r) cn[r] = P[(size_t)r * ld + jn];
        }
        double o[7];
        const double dt = prm.delta_t;
        o[0] = c[0] + dt * c[7]; o[1] = c[1] + dt * c[8]; o[2] = c[2] + dt * c[9];
#pragma unroll
        for (int r = 0; r < 4; ++r)
            o[3 + r] = Fqq[r][0] * c[3] + Fqq[r][1] * c[4] + Fqq[r][2] * c[5] + Fqq[r][3] * c[6] +
                       Fqw[r][0] * c[10] + Fqw[r][1] * c[11] + Fqw[r][2] * c[12];
        if (j < n) {
#pragma unroll
            for (int r = 0; r < 7; ++r) P[(size_t)r * ld + j] = o[r];
        }
        // mirror image: entries 0..7 of rows j (entry 7 is unchanged), eight rows of 64 bytes per warp instruction
#pragma unroll
        for (int r = 0; r < 7; ++r) tr[warp][lane][r] = o[r];
        tr[warp][lane][7] = c[7];
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int row = it * 8 + (lane >> 2), part = (lane & 3) * 2;
            const int jr = j - lane + row;
            if (jr < n) {
                const double2 val = make_double2(tr[warp][row][part], tr[warp][row][part + 1]);
                *reinterpret_cast<double2*>(P + (size_t)jr * ld + part) = val;
            }
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 13; ++r) c[r] = cn[r];
        j = jn;
    }
    // camera block: T = F Pxx ; Pxx' = T F' + Q, stored symmetric
    for (int e = tid; e < 169; e += blockDim.x) {
        const int r = e / 13, cc = e - r * 13;
        double s = 0.0;
        for (int k = 0; k < 13; ++k) s += F[r][k] * Pxx[k][cc];
        T[r][cc] = s;
    }
    __syncthreads();
    for (int e = tid; e < 169; e += blockDim.x) {
        const int r = e / 13, cc = e - r * 13;
        if (cc > r) continue;
        double s = 0.0;
        for (int k = 0; k < 13; ++k) s += T[r][k] * F[cc][k];
        s += Q[r][cc];
        P[(size_t)r * ld + cc] = s;
        P[(size_t)cc * ld + r] = s;
    }
}

void launch_predict(ekfslam_ctx* c) {
    KScope ks(c, KT_PREDICT);
    k_predict<<<c->v.B, 256, 0, c->stream>>>(c->v, c->prm);
}

// ---------------------------------------------------------------------------------------
// mc/predict_camera_measurements.m:8-28 + mc/calculate_derivatives.m:6-28.
// One thread per (filter, feature).  h is only overwritten when the feature is visible at this
// state (a stale h from earlier in the frame survives, and H is then linearised at that stale
// pixel — mc/predict_camera_measurements.m:14-16, mc/calculate_Hi_inverse_depth.m:3).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 6) k_features(DevView v, DevCam cam, int which, int parts) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= v.B * v.N) return;
    const int b = t / v.N, i = t - b * v.N;
    if (i >= v.nfeat[b]) return;
    const int type = v.ftype[t];
    if (type == EKFSLAM_FEAT_NONE) return;
    const double* __restrict__ x = (which ? v.xp : v.x) + (size_t)b * v.ld;
    double xv[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) xv[k] = x[k];
    double R[9];
    q2r_dev(xv + 3, R);
    const int off = v.foff[t];
    double y[6];
    const int w = (type == EKFSLAM_FEAT_INVERSEDEPTH) ? 6 : 3;
    if (w == 6) {   // 48 contiguous bytes per thread: 128-bit loads where the offset allows (every load instruction of a warp costs 32 L1 wavefronts)
        const double* __restrict__ px = x + off;
        if (off & 1) {
            const double2 a = *reinterpret_cast<const double2*>(px + 1), c2 = *reinterpret_cast<const double2*>(px + 3);
            y[0] = px[0]; y[1] = a.x; y[2] = a.y; y[3] = c2.x; y[4] = c2.y; y[5] = px[5];
        } else {
            const double2 a = *reinterpret_cast<const double2*>(px), c2 = *reinterpret_cast<const double2*>(px + 2),
                          e = *reinterpret_cast<const double2*>(px + 4);
            y[0] = a.x; y[1] = a.y; y[2] = c2.x; y[3] = c2.y; y[4] = e.x; y[5] = e.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) y[k] = (k < 3) ? x[off + k] : 0.0;
    }

    uint8_t f = v.flags[t];
    double hu = 0.0, hv = 0.0;
    if ((parts & 1) && predict_h_dev(cam, xv, R, y, type, hu, hv)) {
        *reinterpret_cast<double2*>(v.h + 2 * (size_t)t) = make_double2(hu, hv);
        f |= EKFSLAM_F_HAS_H;
        v.flags[t] = f;
    } else if (f & EKFSLAM_F_HAS_H) {
        const double2 h2 = *reinterpret_cast<const double2*>(v.h + 2 * (size_t)t);
        hu = h2.x; hv = h2.y;
    }
    if ((parts & 2) && (f & EKFSLAM_F_HAS_H)) {
        double Hc[EKF_HSTRIDE];
        jacobian_dev(cam, xv, R, y, type, hu, hv, Hc);
        double2* dst = reinterpret_cast<double2*>(v.Hc + (size_t)t * EKF_HSTRIDE);   // 208 bytes per feature: 13 128-bit stores
#pragma unroll
        for (int k = 0; k < EKF_HSTRIDE / 2; ++k) dst[k] = make_double2(Hc[2 * k], Hc[2 * k + 1]);
    }
}

void launch_features(ekfslam_ctx* c, int which, int parts) {
    const int tot = c->v.B * c->v.N;
    KScope ks(c, KT_FEATURES);
    k_features<<<(tot + 127) / 128, 128, 0, c->stream>>>(c->v, c->cam, which, parts);
}

// ---------------------------------------------------------------------------------------
// G = H * P for the selected features: rows 2i, 2i+1 of G are H_i (2 x n, 13 non-zero columns)
// times P.  P is symmetric, so row a of G is a combination of 13 ROWS of P, read coalesced.
// grid = (column chunks of 256, B); each thread owns two adjacent columns.
// A feature is selected iff (flags & need) == need && (flags & forbid) == 0.
// This single pass over P feeds S_i (mc/search_IC_matches.m:8), every 1-point RANSAC gain
// K = P H_i' inv(S_i) (mc/ransac_hypotheses.m:24-25) and P H' of the update (mc/update.m:8-9).
// ---------------------------------------------------------------------------------------
#define HP_CHUNK 64    // features per compaction round (128 = one round at N = 100 was measured: k_hp 1.41 -> 1.75 ms)
__global__ void __launch_bounds__(128) k_hp(DevView v, int need, int forbid, int fch) {
    const int b = blockIdx.y;
    const int n = v.nstate[b];
    const int ld = v.ld;
    const int c0 = (blockIdx.x * 128 + threadIdx.x) * 2;
    if (blockIdx.x * 256 >= n) return;
    const int nf = v.nfeat[b];
    const double* __restrict__ P = v.P + (size_t)b * v.nmax * ld;
    double* __restrict__ G = v.G + (size_t)b * v.kmax * ld;

    __shared__ double sH[HP_CHUNK][EKF_HSTRIDE];
    __shared__ int sOff[HP_CHUNK];
    __shared__ int sIdx[HP_CHUNK];
    __shared__ int sW[HP_CHUNK];
    __shared__ int sCnt;

    const bool active = c0 < ld && c0 < n;
    double2 pc[7];
    if (active) {
#pragma unroll
        for (int r = 0; r < 7; ++r) pc[r] = *reinterpret_cast<const double2*>(P + (size_t)r * ld + c0);
    }
    // blockIdx.z splits the feature chunks (fch <= HP_CHUNK features each) among CTAs: few filters would otherwise leave
    // most SMs idle (one filter of N = 100: 6 CTAs streaming 3 MB)
    for (int f0 = blockIdx.z * fch; f0 < nf; f0 += gridDim.z * fch) {
        __syncthreads();
        if (threadIdx.x < 32) {
            // warp 0 compacts the selected features of this chunk (ballot + prefix popcount)
            int base = 0;
#pragma unroll
            for (int h2 = 0; h2 < HP_CHUNK / 32; ++h2) {
                const int j = h2 * 32 + threadIdx.x;
                const int i = f0 + j;
                bool selq = false;
                int ty = 0, of = 0;
                if (j < fch && i < nf) {
                    const int t = b * v.N + i;
                    const uint8_t fl = v.flags[t];
                    ty = v.ftype[t];
                    of = v.foff[t];
                    selq = (ty != EKFSLAM_FEAT_NONE) && ((fl & need) == need) && ((fl & forbid) == 0);
                }
                const unsigned m = __ballot_sync(0xffffffffu, selq);
                if (selq) {
                    const int slot = base + __popc(m & ((1u << threadIdx.x) - 1u));
                    sIdx[slot] = i; sOff[slot] = of; sW[slot] = (ty == EKFSLAM_FEAT_INVERSEDEPTH) ? 6 : 3;
                }
                base += __popc(m);
            }
            if (threadIdx.x == 0) sCnt = base;
        }
        __syncthreads();
        const int cnt = sCnt;
        for (int e = threadIdx.x; e < cnt * EKF_HSTRIDE; e += blockDim.x) {
            const int s = e / EKF_HSTRIDE, k = e - s * EKF_HSTRIDE;
            sH[s][k] = v.Hc[((size_t)b * v.N + sIdx[s]) * EKF_HSTRIDE + k];
        }
        __syncthreads();
        if (!active) continue;
        for (int s = 0; s < cnt; ++s) {
            const double* Hs = sH[s];
            double2 g0 = make_double2(0.0, 0.0), g1 = make_double2(0.0, 0.0);
#pragma unroll
            for (int r = 0; r < 7; ++r) {
                g0.x += Hs[r] * pc[r].x; g0.y += Hs[r] * pc[r].y;
                g1.x += Hs[EKF_HC + r] * pc[r].x; g1.y += Hs[EKF_HC + r] * pc[r].y;
            }
            const int off = sOff[s];
            const int w = sW[s];
            const double* Pr = P + (size_t)off * ld + c0;
            if (w == 6) {
                double2 pf[6];
#pragma unroll
                for (int r = 0; r < 6; ++r) pf[r] = *reinterpret_cast<const double2*>(Pr + (size_t)r * ld);
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    g0.x += Hs[7 + r] * pf[r].x; g0.y += Hs[7 + r] * pf[r].y;
                    g1.x += Hs[EKF_HC + 7 + r] * pf[r].x; g1.y += Hs[EKF_HC + 7 + r] * pf[r].y;
                }
            } else {
                double2 pf[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) pf[r] = *reinterpret_cast<const double2*>(Pr + (size_t)r * ld);
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    g0.x += Hs[7 + r] * pf[r].x; g0.y += Hs[7 + r] * pf[r].y;
                    g1.x += Hs[EKF_HC + 7 + r] * pf[r].x; g1.y += Hs[EKF_HC + 7 + r] * pf[r].y;
                }
            }
            const int i = sIdx[s];
            *reinterpret_cast<double2*>(G + (size_t)(2 * i) * ld + c0) = g0;
            *reinterpret_cast<double2*>(G + (size_t)(2 * i + 1) * ld + c0) = g1;
        }
    }
}

void launch_hp(ekfslam_ctx* c, int need, int forbid, int slot) {
    const int colchunks = (c->v.nmax + 255) / 256;
    // feature-chunk groups (blockIdx.z): only when the (column chunk, filter) grid cannot fill the GPU; the chunk shrinks
    // from 64 features down to 8 until there are ~4 CTAs per SM
    const long long base = (long long)colchunks * c->v.B;
    const long long want = 4LL * c->sm_count;
    int fch = HP_CHUNK;
    while (fch > 8 && base * ((c->v.N + fch - 1) / fch) < want) fch >>= 1;
    int fz = (c->v.N + fch - 1) / fch;
    if (base >= want) fz = 1;
    else if (base * fz > want) fz = (int)((want + base - 1) / base);
    dim3 grid(colchunks, c->v.B, fz);
    KScope ks(c, slot);
    k_hp<<<grid, 128, 0, c->stream>>>(c->v, need, forbid, fch);
}

// ---------------------------------------------------------------------------------------
// The match gates.  One thread per feature.
//   mode 1: synthetic matcher gate (mc/matching.m:16,38) on candidates  -> z, HAS_Z, IC  (S_i from k_innov_gather)
//   mode 2: explicit matches (what mc/matching.m:52-53 would have written) -> z, HAS_Z, IC
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_innov(DevView v, ekfslam_params prm, int mode) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= v.B * v.N) return;
    const int b = t / v.N, i = t - b * v.N;
    if (i >= v.nfeat[b]) return;
    if (v.ftype[t] == EKFSLAM_FEAT_NONE) return;
    const uint8_t f = v.flags[t];
    if (mode == 2) {
        const uint8_t m = v.mflags[t] & (EKFSLAM_F_HAS_Z | EKFSLAM_F_IC);
        if (m) {
            v.z[2 * t] = v.zc[2 * t]; v.z[2 * t + 1] = v.zc[2 * t + 1];
            v.flags[t] = f | m;
        }
        return;
    }
    if (!(f & EKFSLAM_F_HAS_H) || !(v.mflags[t] & EKFSLAM_F_CAND)) return;
    const double s00 = v.S[4 * t], s01 = v.S[4 * t + 1], s10 = v.S[4 * t + 2], s11 = v.S[4 * t + 3];
    const double zu = v.zc[2 * t], zv = v.zc[2 * t + 1];
    const double n0 = zu - v.h[2 * t], n1 = zv - v.h[2 * t + 1];
    const double det = s00 * s11 - s01 * s10;
    // nu' inv(S) nu with inv(S) = [s11 -s01; -s10 s00]/det
    const double d2 = (n0 * (s11 * n0 - s01 * n1) + n1 * (-s10 * n0 + s00 * n1)) / det;
    // all(eig(S) < 100): the larger eigenvalue of the 2x2
    const double tr = s00 + s11;
    const double disc = sqrt((s00 - s11) * (s00 - s11) + 4.0 * s01 * s10);
    const double lmax = 0.5 * (tr + disc);
    if (lmax < 100.0 && d2 < prm.chi2_gate) {
        v.z[2 * t] = zu; v.z[2 * t + 1] = zv;
        v.flags[t] = f | EKFSLAM_F_HAS_Z | EKFSLAM_F_IC;
    }
}

// ---------------------------------------------------------------------------------------
// The 2x2 innovation covariance of a feature WITHOUT the G rows: S_i only needs the 13x13 (10x10) block P[c,c] of the
// columns H_i touches - 6 short row segments per feature (columns 0..6 and the feature's own block of rows
// off..off+5, read from the authoritative lower triangle) plus the 7x7 camera block shared by all features of a
// filter.  One thread per feature; ~25 sectors of DRAM traffic per feature instead of 2 x n doubles of G.
//   mode 0: S_i = H_i P H_i' + R_i for every predicted feature (mc/search_IC_matches.m:6-10), stored;
//   mode 3: rescue gate (mc/rescue_hi_inliers.m:11-20): S_i = H_i p_k_k H_i' (no R) for IC && !LI,
//           nu' inv(S_i) nu < chi2 -> HI.  S is not stored (it is a local in the reference).  Only the features that
//           pass then need full rows H p_k_k (k_hp on the HI rows).
// ---------------------------------------------------------------------------------------
#ifndef IG_MINB
#define IG_MINB 2   // the whole gather lives in registers (89 doubles): ~200 registers without spills
#endif
__global__ void __launch_bounds__(128, IG_MINB) k_innov_gather(DevView v, ekfslam_params prm, int mode) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= v.B * v.N) return;
    const int b = t / v.N, i = t - b * v.N;
    if (i >= v.nfeat[b]) return;
    const int type = v.ftype[t];
    if (type == EKFSLAM_FEAT_NONE) return;
    uint8_t f = v.flags[t];
    if (!(f & EKFSLAM_F_HAS_H)) return;
    if (mode == 3 && !((f & EKFSLAM_F_IC) && !(f & EKFSLAM_F_LI))) return;
    const int ld = v.ld;
    const double* __restrict__ P = v.P + (size_t)b * v.nmax * ld;
    const int off = v.foff[t];
    const int w = (type == EKFSLAM_FEAT_INVERSEDEPTH) ? 6 : 3;
    // Every element of the gather is requested ONCE, up front, with 128-bit loads where the alignment allows.  A thread's
    // rows are its own, so each load instruction of a warp touches 32 different lines = 32 L1 wavefronts whatever its width:
    // the kernel ran at exactly that bound (ncu, round 2: 146 scattered 64-bit loads per feature - the 6x7 cross block was
    // fetched twice, once per orientation, the 6x6 block as a full square - 0.198 ms = 4672 wavefronts per warp).
    double H[EKF_HSTRIDE];
    {
        const double2* __restrict__ Hp2 = reinterpret_cast<const double2*>(v.Hc + (size_t)t * EKF_HSTRIDE);   // 208 bytes per feature
#pragma unroll
        for (int k = 0; k < EKF_HSTRIDE / 2; ++k) { const double2 h2 = Hp2[k]; H[2 * k] = h2.x; H[2 * k + 1] = h2.y; }
    }
    double A[6][7];    // A[r][j] = P[off + r][j], j < 7           (rows start on 256-byte boundaries)
    double Lb[6][6];   // Lb[r][c] = P[off + r][off + c], c <= r  (the authoritative lower triangle of the feature's own block)
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) Lb[r][cc] = 0.0;
    if (w == 6) {
        const bool odd = off & 1;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const double* __restrict__ prow = P + (size_t)(off + r) * ld;
            const double2 a01 = *reinterpret_cast<const double2*>(prow), a23 = *reinterpret_cast<const double2*>(prow + 2),
                          a45 = *reinterpret_cast<const double2*>(prow + 4);
            A[r][0] = a01.x; A[r][1] = a01.y; A[r][2] = a23.x; A[r][3] = a23.y; A[r][4] = a45.x; A[r][5] = a45.y; A[r][6] = prow[6];
            const double* __restrict__ pb = prow + off;   // entries beyond c = r that ride along in a pair are valid memory, unused
            if (odd) {
                Lb[r][0] = pb[0];
                if (r >= 1) { const double2 d = *reinterpret_cast<const double2*>(pb + 1); Lb[r][1] = d.x; if (r >= 2) Lb[r][2] = d.y; }
                if (r >= 3) { const double2 d = *reinterpret_cast<const double2*>(pb + 3); Lb[r][3] = d.x; if (r >= 4) Lb[r][4] = d.y; }
                if (r >= 5) Lb[r][5] = pb[5];
            } else {
                { const double2 d = *reinterpret_cast<const double2*>(pb); Lb[r][0] = d.x; if (r >= 1) Lb[r][1] = d.y; }
                if (r >= 2) { const double2 d = *reinterpret_cast<const double2*>(pb + 2); Lb[r][2] = d.x; if (r >= 3) Lb[r][3] = d.y; }
                if (r >= 4) { const double2 d = *reinterpret_cast<const double2*>(pb + 4); Lb[r][4] = d.x; if (r >= 5) Lb[r][5] = d.y; }
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double* __restrict__ prow = P + (size_t)(off + r) * ld;
#pragma unroll
            for (int j = 0; j < 7; ++j) A[r][j] = prow[j];
#pragma unroll
            for (int cc = 0; cc <= r; ++cc) Lb[r][cc] = prow[off + cc];
        }
#pragma unroll
        for (int r = 3; r < 6; ++r)
#pragma unroll
            for (int j = 0; j < 7; ++j) A[r][j] = 0.0;
    }
    double s00 = 0.0, s01 = 0.0, s10 = 0.0, s11 = 0.0;
    // column j of the gather: t_a = sum_r H[a][r] P[c_r][c_j], then S[a][a'] += t_a H[a'][j]   (same order of operations as
    // the element-by-element version)
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        double t0 = 0.0, t1 = 0.0;
#pragma unroll
        for (int r = 0; r < 7; ++r) {
            const double pv = (r >= j) ? P[(size_t)r * ld + j] : P[(size_t)j * ld + r];   // camera block: the same addresses for all features of a filter
            t0 += H[r] * pv; t1 += H[EKF_HC + r] * pv;
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            if (r < w) { t0 += H[7 + r] * A[r][j]; t1 += H[EKF_HC + 7 + r] * A[r][j]; }
        }
        s00 += t0 * H[j]; s01 += t0 * H[EKF_HC + j]; s10 += t1 * H[j]; s11 += t1 * H[EKF_HC + j];
    }
#pragma unroll
    for (int jj = 0; jj < 6; ++jj) {
        if (jj < w) {
            double t0 = 0.0, t1 = 0.0;
#pragma unroll
            for (int r = 0; r < 7; ++r) { const double pv = A[jj][r]; t0 += H[r] * pv; t1 += H[EKF_HC + r] * pv; }
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                if (r < w) {
                    const double pv = (r <= jj) ? Lb[jj][r] : Lb[r][jj];
                    t0 += H[7 + r] * pv; t1 += H[EKF_HC + 7 + r] * pv;
                }
            }
            s00 += t0 * H[7 + jj]; s01 += t0 * H[EKF_HC + 7 + jj]; s10 += t1 * H[7 + jj]; s11 += t1 * H[EKF_HC + 7 + jj];
        }
    }
    if (mode == 0) {
        s00 += 1.0; s11 += 1.0;   // + R_i = eye(2), mc/add_feature_to_info_vector.m:32
        v.S[4 * t] = s00; v.S[4 * t + 1] = s01; v.S[4 * t + 2] = s10; v.S[4 * t + 3] = s11;
        return;
    }
    const double n0 = v.z[2 * t] - v.h[2 * t], n1 = v.z[2 * t + 1] - v.h[2 * t + 1];
    const double det = s00 * s11 - s01 * s10;
    const double d2 = (n0 * (s11 * n0 - s01 * n1) + n1 * (-s10 * n0 + s00 * n1)) / det;
    if (d2 < prm.chi2_gate) f |= EKFSLAM_F_HI; else f &= ~EKFSLAM_F_HI;
    v.flags[t] = f;
}

void launch_innov_gather(ekfslam_ctx* c, int mode) {
    const int tot = c->v.B * c->v.N;
    KScope ks(c, KT_INNOV);
    k_innov_gather<<<(tot + 127) / 128, 128, 0, c->stream>>>(c->v, c->prm, mode);
}

void launch_innov(ekfslam_ctx* c, int mode) {
    const int tot = c->v.B * c->v.N;
    KScope ks(c, KT_INNOV);
    k_innov<<<(tot + 127) / 128, 128, 0, c->stream>>>(c->v, c->prm, mode);
}

// ---------------------------------------------------------------------------------------
// Mirror the lower triangle of freshly uploaded covariances into the upper one.  The kernels
// read P by rows and write tiles together with their transposes, which assumes an exactly
// symmetric P; a host matrix such as J P J' is only symmetric to rounding.
// ---------------------------------------------------------------------------------------
__global__ void k_symmetrize(DevView v, int b0) {
    const int b = b0 + blockIdx.y;
    const int n = v.nstate[b];
    const int ld = v.ld;
    double* __restrict__ P = v.P + (size_t)b * v.nmax * ld;
    __shared__ double tile[32][33];
    const int ti = blockIdx.x / ((v.nmax + 31) / 32), tj = blockIdx.x % ((v.nmax + 31) / 32);
    if (tj > ti) return;
    const int i0 = ti * 32, j0 = tj * 32;
    if (i0 >= n) return;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int gi = i0 + r, gj = j0 + threadIdx.x;
        tile[r][threadIdx.x] = (gi < n && gj < n) ? P[(size_t)gi * ld + gj] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int gj = j0 + r, gi = i0 + threadIdx.x;  // writes P[gj][gi] = tile[gi-i0][gj-j0]
        if (gi < n && gj < n && gi > gj) P[(size_t)gj * ld + gi] = tile[threadIdx.x][r];
    }
}

void launch_symmetrize(ekfslam_ctx* c, int b0, int nb) {
    const int nt = (c->v.nmax + 31) / 32;
    dim3 grid(nt * nt, nb), block(32, 8);
    KScope ks(c, KT_SYMMETRIZE);
    k_symmetrize<<<grid, block, 0, c->stream>>>(c->v, b0);
}

// ---------------------------------------------------------------------------------------
// Map management, SURVEY §8f rank 1: append one inverse-depth feature per filter from its distorted
// pixel.  mc/add_features_inverse_depth.m:18-22 -> mc/hinv.m:3-26 (state) and
// mc/add_a_feature_covariance_inverse_depth.m:3-64 (covariance augmentation):
//   P <- [P, P(:,1:13) J'; J P(1:13,:), J Pxv J' + D],  J = dy_dxv (6x13), D = dy_dhd Padd dy_dhd'.
// J is [I3 0 0; 0 dtheta_dq 0; 0 dphi_dq 0; 0 0 0], so a new row is a combination of P rows 0..6.
// One block per filter; operates on (x_k_k, p_k_k).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_add_feature(DevView v, DevCam cam, int b0, const double* __restrict__ uvd,
                                                     int uvd_stride, const uint8_t* __restrict__ add,
                                                     const int32_t* __restrict__ quota, int jq,
                                                     const int32_t* __restrict__ tag_src, int tag_stride,
                                                     double std_pxl, double rho0, double std_rho) {
    const int lb = blockIdx.x;
    const int b = b0 + lb;
    if (add && !add[lb]) return;
    if (quota && jq >= quota[b]) return;
    const int n = v.nstate[b], nf = v.nfeat[b];
    if (n + 6 > v.nmax || nf >= v.N) {
        if (threadIdx.x == 0) atomicOr(&v.stats[b].status, 4);  // no room: reported, feature not added
        return;
    }
    const int ld = v.ld;
    double* __restrict__ P = v.P + (size_t)b * v.nmax * ld;
    double* __restrict__ x = v.x + (size_t)b * ld;
    __shared__ double dth_dq[4], dph_dq[4], D[6][6], newf[6];
    if (threadIdx.x == 0) {
        const double fku = cam.f / cam.dx, fkv = cam.f / cam.dy;  // cam.K(1,1), cam.K(2,2)
        const double ud = uvd[(size_t)lb * uvd_stride], vd = uvd[(size_t)lb * uvd_stride + 1];
        // mc/undistort_fm.m:18-27
        const double xd = (ud - cam.Cx) * cam.dx, yd = (vd - cam.Cy) * cam.dy;
        const double rd = sqrt(xd * xd + yd * yd);
        const double Dd = 1.0 + cam.k1 * rd * rd + cam.k2 * rd * rd * rd * rd;
        const double uu = xd * Dd / cam.dx + cam.Cx, vu = yd * Dd / cam.dy + cam.Cy;
        const double hc[3] = {-(cam.Cx - uu) / fku, -(cam.Cy - vu) / fkv, 1.0};
        double R[9];
        q2r_dev(x + 3, R);
        const double nx = R[0] * hc[0] + R[1] * hc[1] + R[2] * hc[2];
        const double ny = R[3] * hc[0] + R[4] * hc[1] + R[5] * hc[2];
        const double nz = R[6] * hc[0] + R[7] * hc[1] + R[8] * hc[2];
        newf[0] = x[0]; newf[1] = x[1]; newf[2] = x[2];
        newf[3] = atan2(nx, nz);
        newf[4] = atan2(-ny, sqrt(nx * nx + nz * nz));
        newf[5] = rho0;
        // derivatives, mc/add_a_feature_covariance_inverse_depth.m:28-49
        const double xz2 = nx * nx + nz * nz, n2 = xz2 + ny * ny, sxz = sqrt(xz2);
        const double dth[3] = {nz / xz2, 0.0, -nx / xz2};
        const double dph[3] = {(nx * ny) / (n2 * sxz), -sxz / n2, (nz * ny) / (n2 * sxz)};
        const double q0 = x[3], qx = x[4], qy = x[5], qz = x[6];
        // dgw_dqwr = dRq_times_a_by_dq(q, hc)  (3x4)
        double T[12];
        T[0] = 2 * q0 * hc[0] - 2 * qz * hc[1] + 2 * qy * hc[2];
        T[4] = 2 * qz * hc[0] + 2 * q0 * hc[1] - 2 * qx * hc[2];
        T[8] = -2 * qy * hc[0] + 2 * qx * hc[1] + 2 * q0 * hc[2];
        T[1] = 2 * qx * hc[0] + 2 * qy * hc[1] + 2 * qz * hc[2];
        T[5] = 2 * qy * hc[0] - 2 * qx * hc[1] - 2 * q0 * hc[2];
        T[9] = 2 * qz * hc[0] + 2 * q0 * hc[1] - 2 * qx * hc[2];
        T[2] = -2 * qy * hc[0] + 2 * qx * hc[1] + 2 * q0 * hc[2];
        T[6] = 2 * qx * hc[0] + 2 * qy * hc[1] + 2 * qz * hc[2];
        T[10] = -2 * q0 * hc[0] + 2 * qz * hc[1] - 2 * qy * hc[2];
        T[3] = -2 * qz * hc[0] - 2 * q0 * hc[1] + 2 * qx * hc[2];
        T[7] = 2 * q0 * hc[0] - 2 * qz * hc[1] + 2 * qy * hc[2];
        T[11] = 2 * qx * hc[0] + 2 * qy * hc[1] + 2 * qz * hc[2];
        for (int k = 0; k < 4; ++k) {
            dth_dq[k] = dth[0] * T[k] + dth[1] * T[4 + k] + dth[2] * T[8 + k];
            dph_dq[k] = dph[0] * T[k] + dph[1] * T[4 + k] + dph[2] * T[8 + k];
        }
        // dyprima_dhd (5x2) = dyprima_dgw * R * dgc_dhu * dhu_dhd ; rows 0..2 are zero
        const double du = ud - cam.Cx, dv = vd - cam.Cy;
        const double rd2 = xd * xd + yd * yd, rd4 = rd2 * rd2;
        const double gg = 1.0 + cam.k1 * rd2 + cam.k2 * rd4, ee = cam.k1 + 2.0 * cam.k2 * rd2;
        const double Ju[4] = {gg + du * ee * (2.0 * du * cam.dx * cam.dx), du * ee * (2.0 * dv * cam.dy * cam.dy),
                              dv * ee * (2.0 * du * cam.dx * cam.dx), gg + dv * ee * (2.0 * dv * cam.dy * cam.dy)};
        double RG[6];  // R * dgc_dhu = first two columns of R scaled by 1/fku, 1/fkv   (3x2)
        for (int r = 0; r < 3; ++r) { RG[2 * r] = R[3 * r] / fku; RG[2 * r + 1] = R[3 * r + 1] / fkv; }
        double A[4];   // rows theta, phi of dyprima_dgw * RG  (2x2)
        A[0] = dth[0] * RG[0] + dth[1] * RG[2] + dth[2] * RG[4]; A[1] = dth[0] * RG[1] + dth[1] * RG[3] + dth[2] * RG[5];
        A[2] = dph[0] * RG[0] + dph[1] * RG[2] + dph[2] * RG[4]; A[3] = dph[0] * RG[1] + dph[1] * RG[3] + dph[2] * RG[5];
        double Bm[4];  // * dhu_dhd
        Bm[0] = A[0] * Ju[0] + A[1] * Ju[2]; Bm[1] = A[0] * Ju[1] + A[1] * Ju[3];
        Bm[2] = A[2] * Ju[0] + A[3] * Ju[2]; Bm[3] = A[2] * Ju[1] + A[3] * Ju[3];
        // D = dy_dhd * diag(std_pxl^2, std_pxl^2, std_rho^2) * dy_dhd'
        for (int r = 0; r < 6; ++r) for (int cc = 0; cc < 6; ++cc) D[r][cc] = 0.0;
        const double sp = std_pxl * std_pxl;
        D[3][3] = (Bm[0] * Bm[0] + Bm[1] * Bm[1]) * sp; D[3][4] = (Bm[0] * Bm[2] + Bm[1] * Bm[3]) * sp;
        D[4][3] = D[3][4];                              D[4][4] = (Bm[2] * Bm[2] + Bm[3] * Bm[3]) * sp;
        D[5][5] = std_rho * std_rho;
    }
    __syncthreads();
    // new rows n..n+5 against the existing columns j < n (and their mirror images)
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const double p0 = P[(size_t)0 * ld + j], p1 = P[(size_t)1 * ld + j], p2 = P[(size_t)2 * ld + j];
        const double p3 = P[(size_t)3 * ld + j], p4 = P[(size_t)4 * ld + j], p5 = P[(size_t)5 * ld + j], p6 = P[(size_t)6 * ld + j];
        double o[6];
        o[0] = p0; o[1] = p1; o[2] = p2;
        o[3] = dth_dq[0] * p3 + dth_dq[1] * p4 + dth_dq[2] * p5 + dth_dq[3] * p6;
        o[4] = dph_dq[0] * p3 + dph_dq[1] * p4 + dph_dq[2] * p5 + dph_dq[3] * p6;
        o[5] = 0.0;
#pragma unroll
        for (int r = 0; r < 6; ++r) { P[(size_t)(n + r) * ld + j] = o[r]; P[(size_t)j * ld + n + r] = o[r]; }
    }
    __syncthreads();
    // corner J Pxv J' + D: (J Pxv J')[r][c] = sum_j (J P)[r][j] J[c][j], with (J P)[r][j] = P[n+r][j] just written
    if (threadIdx.x < 36) {
        const int r = threadIdx.x / 6, cc = threadIdx.x % 6;
        if (cc <= r) {
            const double* row = P + (size_t)(n + r) * ld;
            double s;
            if (cc < 3) s = row[cc];
            else if (cc == 3) s = row[3] * dth_dq[0] + row[4] * dth_dq[1] + row[5] * dth_dq[2] + row[6] * dth_dq[3];
            else if (cc == 4) s = row[3] * dph_dq[0] + row[4] * dph_dq[1] + row[5] * dph_dq[2] + row[6] * dph_dq[3];
            else s = 0.0;
            s += D[r][cc];
            P[(size_t)(n + r) * ld + n + cc] = s;
            P[(size_t)(n + cc) * ld + n + r] = s;
        }
    }
    if (threadIdx.x < 6) x[n + threadIdx.x] = newf[threadIdx.x];
    if (threadIdx.x == 0) {
        const size_t t = (size_t)b * v.N + nf;
        v.ftype[t] = EKFSLAM_FEAT_INVERSEDEPTH;
        v.foff[t] = n;
        v.flags[t] = 0; v.mflags[t] = 0;
        v.counters[2 * t] = 0; v.counters[2 * t + 1] = 0;
        v.tag[t] = tag_src ? tag_src[(size_t)lb * tag_stride] : -1;
        v.nstate[b] = n + 6;
        v.nfeat[b] = nf + 1;
    }
}

void launch_add_features(ekfslam_ctx* c, int b0, int nb, const double* d_uvd, int uvd_stride, const uint8_t* d_add,
                         const int32_t* d_quota, int j, const int32_t* d_tag, int tag_stride, double std_pxl,
                         double rho0, double std_rho) {
    KScope ks(c, KT_ADD_FEATURES);
    k_add_feature<<<nb, 128, 0, c->stream>>>(c->v, c->cam, b0, d_uvd, uvd_stride, d_add, d_quota, j, d_tag, tag_stride,
                                             std_pxl, rho0, std_rho);
}

// ---------------------------------------------------------------------------------------
// mc/initialize_x_and_p.m broadcast: every filter in [b0, b0+nb) becomes the 13-state camera-only
// filter (xv, Pxv) with an empty map (the caller zeroed x and P beforehand).
// ---------------------------------------------------------------------------------------
__global__ void k_reset_filters(DevView v, int b0, const double* __restrict__ xv, const double* __restrict__ Pxv) {
    const int b = b0 + blockIdx.x;
    const int t = threadIdx.x;
    if (t < 169) v.P[(size_t)b * v.nmax * v.ld + (size_t)(t / 13) * v.ld + (t % 13)] = Pxv[t];
    if (t < 13) v.x[(size_t)b * v.ld + t] = xv[t];
    if (t == 0) { v.nstate[b] = 13; v.nfeat[b] = 0; }
}

void launch_reset_filters(ekfslam_ctx* c, int b0, int nb, const double* d_xv, const double* d_Pxv) {
    KScope ks(c, KT_ADD_FEATURES);
    k_reset_filters<<<nb, 192, 0, c->stream>>>(c->v, b0, d_xv, d_Pxv);
}

// ---------------------------------------------------------------------------------------
// Map management, SURVEY §8f rank 2: inverse-depth -> Cartesian conversion and feature deletion.
// Both shrink the state: rows/columns [start, start+cnt) of P (and entries of x) are removed and
// everything behind them moves up.  Block-wide, in place, row by row in increasing order; inside a row
// every thread first reads its source element, then all write (the source can be the same row).
// ---------------------------------------------------------------------------------------
__device__ void compact_remove(double* __restrict__ P, double* __restrict__ x, int n, int ld, int start, int cnt) {
    const int nn = n - cnt;
    for (int r = 0; r < nn; ++r) {
        const int sr = (r < start) ? r : r + cnt;
        // each thread handles columns c = tid, tid+blockDim, ... ; read all first
        double vals[8];
        int q = 0;
        for (int c = threadIdx.x; c < nn && q < 8; c += blockDim.x, ++q) vals[q] = P[(size_t)sr * ld + ((c < start) ? c : c + cnt)];
        __syncthreads();
        q = 0;
        for (int c = threadIdx.x; c < nn && q < 8; c += blockDim.x, ++q) P[(size_t)r * ld + c] = vals[q];
        __syncthreads();
    }
    // clear the vacated tail so that the padding invariants (zeros beyond n) keep holding
    for (int r = 0; r < n; ++r)
        for (int c = threadIdx.x; c < n; c += blockDim.x)
            if (r >= nn || c >= nn) P[(size_t)r * ld + c] = 0.0;
    __syncthreads();
    double xv[8];
    int q = 0;
    for (int c = threadIdx.x; c < nn && q < 8; c += blockDim.x, ++q) xv[q] = x[(c < start) ? c : c + cnt];
    __syncthreads();
    q = 0;
    for (int c = threadIdx.x; c < nn && q < 8; c += blockDim.x, ++q) x[c] = xv[q];
    for (int c = nn + threadIdx.x; c < n; c += blockDim.x) x[c] = 0.0;
    __syncthreads();
}

// mc/inversedepth_2_cartesian.m:3-52: per filter, the FIRST inverse-depth feature whose linearity index
// 4*std_d*cos(alpha)/d is below the threshold is converted (at most one per call, :49).
// force_index >= 0 converts that feature unconditionally (the transformation of :35-48 alone).
// Requires blockDim.x * 8 >= n.  conv_out[b] = converted feature index or -1.
__global__ void __launch_bounds__(512) k_id2cart(DevView v, double threshold, int force_index, int32_t* conv_out) {
    const int b = blockIdx.x;
    const int n = v.nstate[b], nf = v.nfeat[b], ld = v.ld, N = v.N;
    double* __restrict__ P = v.P + (size_t)b * v.nmax * ld;
    double* __restrict__ x = v.x + (size_t)b * ld;
    double* __restrict__ R3 = v.G + (size_t)b * v.kmax * ld;  // scratch: 3 rows of length ld
    __shared__ int s_sel, s_pos;
    __shared__ double J[3][6], pnew[3];
    if (threadIdx.x == 0) {
        int sel = -1;
        for (int i = 0; i < nf && sel < 0; ++i) {
            const size_t t = (size_t)b * N + i;
            if (v.ftype[t] != EKFSLAM_FEAT_INVERSEDEPTH) continue;
            if (force_index >= 0) { if (i == force_index) sel = i; continue; }
            const int pos = v.foff[t];
            const double rho = x[pos + 5], th = x[pos + 3], ph = x[pos + 4];
            const double std_rho = sqrt(P[(size_t)(pos + 5) * ld + pos + 5]);
            const double std_d = std_rho / (rho * rho);
            const double m0 = cos(ph) * sin(th), m1 = -sin(ph), m2 = cos(ph) * cos(th);
            const double p0 = x[pos] + (1.0 / rho) * m0, p1 = x[pos + 1] + (1.0 / rho) * m1, p2 = x[pos + 2] + (1.0 / rho) * m2;
            const double a0 = p0 - x[pos], a1 = p1 - x[pos + 1], a2 = p2 - x[pos + 2];   // p - x_c1
            const double c0 = p0 - x[0], c1 = p1 - x[1], c2 = p2 - x[2];                 // p - x_c2
            const double na = sqrt(a0 * a0 + a1 * a1 + a2 * a2), nc = sqrt(c0 * c0 + c1 * c1 + c2 * c2);
            const double cos_alpha = (a0 * c0 + a1 * c1 + a2 * c2) / (na * nc);
            if (4.0 * std_d * cos_alpha / nc < threshold) sel = i;
        }
        s_sel = sel;
        if (sel >= 0) {
            const int pos = v.foff[(size_t)b * N + sel];
            s_pos = pos;
            const double rho = x[pos + 5], th = x[pos + 3], ph = x[pos + 4];
            const double mi[3] = {cos(ph) * sin(th), -sin(ph), cos(ph) * cos(th)};
            const double dmt[3] = {cos(ph) * cos(th), 0.0, -cos(ph) * sin(th)};
            const double dmp[3] = {-sin(ph) * sin(th), -cos(ph), -sin(ph) * cos(th)};
            for (int r = 0; r < 3; ++r) {
                pnew[r] = x[pos + r] + (1.0 / rho) * mi[r];
                for (int c = 0; c < 3; ++c) J[r][c] = (r == c) ? 1.0 : 0.0;
                J[r][3] = (1.0 / rho) * dmt[r];
                J[r][4] = (1.0 / rho) * dmp[r];
                J[r][5] = -mi[r] / (rho * rho);
            }
        }
        if (conv_out) conv_out[b] = sel;
    }
    __syncthreads();
    const int sel = s_sel;
    if (sel < 0) return;
    const int pos = s_pos;
    // R3 = J * P[pos..pos+5, :]   (3 x n)
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        double pc[6];
#pragma unroll
        for (int m = 0; m < 6; ++m) pc[m] = P[(size_t)(pos + m) * ld + c];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            double sacc = 0.0;
#pragma unroll
            for (int m = 0; m < 6; ++m) sacc += J[r][m] * pc[m];
            R3[(size_t)r * ld + c] = sacc;
        }
    }
    __syncthreads();
    // block = R3[:, pos..pos+5] * J'
    __shared__ double blk[3][3];
    if (threadIdx.x < 9) {
        const int r = threadIdx.x / 3, sc = threadIdx.x % 3;
        double sacc = 0.0;
        for (int m = 0; m < 6; ++m) sacc += R3[(size_t)r * ld + pos + m] * J[sc][m];
        blk[r][sc] = sacc;
    }
    __syncthreads();
    // write the 3 new rows / columns at pos..pos+2 (columns pos+3..pos+5 are removed next)
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        if (c >= pos && c < pos + 6) continue;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double val = R3[(size_t)r * ld + c];
            P[(size_t)(pos + r) * ld + c] = val;
            P[(size_t)c * ld + pos + r] = val;
        }
    }
    if (threadIdx.x < 9) {
        const int r = threadIdx.x / 3, sc = threadIdx.x % 3;
        P[(size_t)(pos + r) * ld + pos + sc] = (sc <= r) ? blk[r][sc] : blk[sc][r];  // lower triangle authoritative
    }
    if (threadIdx.x < 3) x[pos + threadIdx.x] = pnew[threadIdx.x];
    __syncthreads();
    compact_remove(P, x, n, ld, pos + 3, 3);
    if (threadIdx.x == 0) {
        v.ftype[(size_t)b * N + sel] = EKFSLAM_FEAT_CARTESIAN;
        for (int i = sel + 1; i < nf; ++i) v.foff[(size_t)b * N + i] -= 3;
        v.nstate[b] = n - 3;
    }
}

// mc/delete_a_feature.m:4-25 for every feature with del[b][i] != 0 (highest index first so that the
// offsets of the remaining candidates stay valid); the per-feature arrays are compacted as well.
__global__ void __launch_bounds__(512) k_delete_features(DevView v, int b0, const uint8_t* __restrict__ del) {
    const int lb = blockIdx.x, b = b0 + lb;
    const int N = v.N, ld = v.ld;
    double* __restrict__ P = v.P + (size_t)b * v.nmax * ld;
    double* __restrict__ x = v.x + (size_t)b * ld;
    __shared__ int s_n, s_nf;
    if (threadIdx.x == 0) { s_n = v.nstate[b]; s_nf = v.nfeat[b]; }
    __syncthreads();
    for (int i = s_nf - 1; i >= 0; --i) {
        if (!del[(size_t)lb * N + i]) continue;    // uniform over the block
        const size_t t = (size_t)b * N + i;
        const int pos = v.foff[t];
        const int w = (v.ftype[t] == EKFSLAM_FEAT_INVERSEDEPTH) ? 6 : 3;
        const int n = s_n, nf = s_nf;
        __syncthreads();
        compact_remove(P, x, n, ld, pos, w);
        // features_info(i) = []: shift the per-feature fields of the features behind it
        if (threadIdx.x == 0) {
            for (int j = i; j < nf - 1; ++j) {
                const size_t d = (size_t)b * N + j, sidx = d + 1;
                v.ftype[d] = v.ftype[sidx]; v.foff[d] = v.foff[sidx] - w; v.flags[d] = v.flags[sidx]; v.mflags[d] = v.mflags[sidx]; v.tag[d] = v.tag[sidx];
                for (int q = 0; q < 2; ++q) { v.h[2 * d + q] = v.h[2 * sidx + q]; v.z[2 * d + q] = v.z[2 * sidx + q]; v.zc[2 * d + q] = v.zc[2 * sidx + q]; v.counters[2 * d + q] = v.counters[2 * sidx + q]; }
                for (int q = 0; q < 4; ++q) v.S[4 * d + q] = v.S[4 * sidx + q];
                for (int q = 0; q < EKF_HSTRIDE; ++q) v.Hc[EKF_HSTRIDE * d + q] = v.Hc[EKF_HSTRIDE * sidx + q];
            }
            const size_t last = (size_t)b * N + nf - 1;
            v.ftype[last] = EKFSLAM_FEAT_NONE; v.flags[last] = 0; v.mflags[last] = 0; v.foff[last] = 0; v.tag[last] = -1;
            s_n = n - w; s_nf = nf - 1;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { v.nstate[b] = s_n; v.nfeat[b] = s_nf; }
}

void launch_id2cart(ekfslam_ctx* c, double threshold, int force_index, int32_t* d_conv) {
    KScope ks(c, KT_ADD_FEATURES);
    k_id2cart<<<c->v.B, 512, 0, c->stream>>>(c->v, threshold, force_index, d_conv);
}

void launch_delete_features(ekfslam_ctx* c, int b0, int nb, const uint8_t* d_del) {
    KScope ks(c, KT_ADD_FEATURES);
    k_delete_features<<<nb, 512, 0, c->stream>>>(c->v, b0, d_del);
}
