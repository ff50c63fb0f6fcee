// Covariance downdate  P <- Jn (P - W'W) Jn'  (mc/update.m:13-22) on 128x128 tiles.
//
// Why 128x128: an in-situ ablation of the 64x64 kernel (tools/dbg_downdate.py) showed the downdate
// is bound by the bytes moved between L2 and the SMs (~6 TB/s), not by the fp64 tensor pipe: with
// 64x64 tiles every tile re-reads two k x 64 panels of W, which is half of all L2 traffic.  A 128x128
// tile halves the W bytes per output element.
//
// Structure (one persistent CTA per SM, 288 threads):
//   * producer warp: streams the W panels (2 x [16][128] per stage) and the 128x128 P tile with bulk
//     asynchronous copies (cp.async.bulk -> SASS UBLKCP) that complete on mbarriers;
//   * 16 consumer warps in a 4x4 grid, 32x32 accumulators each (4x4 DMMA m8n8k4 tiles = 64 registers),
//     fragments read conflict-free from the k-major panels (row stride 132 doubles);
//   * 8x8 DMMA tiles outside n x n or strictly above the diagonal are skipped (16-bit mask);
//   * epilogue: P tile from shared memory, C = P - acc stored with its mirror image straight from the
//     accumulator fragments; the 8 leading columns of tile column 0 (quaternion) go through a strip
//     for the normalisation Jacobian.
// Same summation order per output element as the 64x64 kernels (k ascending in steps of 4).
#include "model.cuh"
#include "tc_common.cuh"

#define T2 128               // tile edge
#define T2PAD 132            // panel row stride (doubles): % 16 == 4 -> conflict-free fragment loads
#define WS2_STAGES 2
#define WS2_CONSUMERS 16
#define WS2_THREADS ((WS2_CONSUMERS + 1) * 32)

extern int g_ekfslam_debug;

__global__ void __launch_bounds__(WS2_THREADS, 1) k_downdate_ws128(DevView v, const double* __restrict__ jn_all, int T,
                                                                   long long total, int M, int dbg) {
    extern __shared__ __align__(16) double dsm[];
    double* As = dsm;                                    // [WS2_STAGES][TK][T2PAD]
    double* Bs = As + WS2_STAGES * TK * T2PAD;           // [WS2_STAGES][TK][T2PAD]
    double* Pt = Bs + WS2_STAGES * TK * T2PAD;           // [128][T2PAD]
    double* strip = Pt + T2 * T2PAD;                     // [128][9]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(strip + T2 * 9);
    unsigned long long* empty = full + WS2_STAGES;
    unsigned long long* pfull = empty + WS2_STAGES;
    unsigned long long* pempty = pfull + 1;
    int2* meta = reinterpret_cast<int2*>(pempty + 1);    // [M]
    unsigned* lut = reinterpret_cast<unsigned*>(meta + M);  // [T]
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
        for (int s2 = 0; s2 < WS2_STAGES; ++s2) { mbar_init(full + s2, 32); mbar_init(empty + s2, WS2_CONSUMERS); }
        mbar_init(pfull, 32);
        mbar_init(pempty, WS2_CONSUMERS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int own = (int)((total - blockIdx.x + G - 1) / G);
    const int Mreal = own < M ? own : M;

    if (warp == WS2_CONSUMERS) {
        // ================= producer warp =================
        unsigned cnt = 0, pcnt = 0;
        const int pr = lane & 15;
        const bool isB = lane >= 16;
        for (int m = 0; m < Mreal; ++m) {
            const DTile L = decode_tile(meta, lut, m, Mreal, blockIdx.x + (long long)m * G, T, T2);
            if (L.nk == 0) continue;
            const double* __restrict__ W = v.W + (size_t)L.b * kmax * ld;
            const double* __restrict__ Pg = v.P + (size_t)L.b * v.nmax * ld;
            const int c0 = isB ? L.j0 : L.i0;
            const unsigned rowbytes = (unsigned)(min(T2, ld - c0) * 8);
            const unsigned bytesA = (unsigned)(min(T2, ld - L.i0) * 8), bytesB = (unsigned)(min(T2, ld - L.j0) * 8);
            // the single P buffer is free once the consumers finished the previous tile's epilogue, which is
            // guaranteed when the ring lets the producer issue stage WS2_STAGES of this tile
            const int p_at = min(L.nk - 1, WS2_STAGES);
            const unsigned pbytes = bytesB;
            const int prow_n = min(T2, L.n - L.i0);
            for (int st = 0; st < L.nk; ++st, ++cnt) {
                const unsigned slot = cnt % WS2_STAGES, ph = (cnt / WS2_STAGES) & 1u;
                mbar_wait(empty + slot, ph ^ 1u);
                const int t0 = st * TK;
                const int nvalid = min(TK, L.k - t0);
                double* dst = (isB ? Bs : As) + slot * TK * T2PAD + pr * T2PAD;
                if (lane == 0)
                    mbar_arrive_expect_tx(full + slot, (dbg & 1) ? 0u : (unsigned)nvalid * (bytesA + (L.diag ? 0u : bytesB)));
                const bool mine = !(isB && L.diag) && !(dbg & 1);
                if (mine && pr < nvalid) {
                    bulk_g2s(dst, W + (size_t)(t0 + pr) * ld + c0, rowbytes, full + slot);
                } else if (mine && pr < ((nvalid + 3) & ~3)) {
                    for (int c = 0; c < T2; ++c) dst[c] = 0.0;  // rows between k and the next multiple of 4
                }
                if (lane != 0) mbar_arrive(full + slot);
                if (st == p_at) {
                    mbar_wait(pempty, (pcnt & 1u) ^ 1u);
                    if (lane == 0) mbar_arrive_expect_tx(pfull, (dbg & 1) ? 0u : (unsigned)prow_n * pbytes);
#pragma unroll
                    for (int h2 = 0; h2 < T2 / 32; ++h2) {
                        const int r = lane + 32 * h2;
                        if (r < prow_n && !(dbg & 1)) bulk_g2s(Pt + r * T2PAD, Pg + (size_t)(L.i0 + r) * ld + L.j0, pbytes, pfull);
                    }
                    if (lane != 0) mbar_arrive(pfull);
                    ++pcnt;
                }
            }
        }
        return;
    }

    // ================= consumer warps: 4 (rows) x 4 (cols), 32 x 32 each =================
    const int wr = warp >> 2, wc = warp & 3, g = lane >> 2, q = lane & 3;
    unsigned cnt = 0, pcnt = 0;
    for (int cm = 0; cm < Mreal; ++cm) {
        const DTile C = decode_tile(meta, lut, cm, Mreal, blockIdx.x + (long long)cm * G, T, T2);
        if (C.nk == 0) continue;
        const int n = C.n, k = C.k, i0 = C.i0, j0 = C.j0;
        const bool diag = C.diag;
        double* __restrict__ P = v.P + (size_t)C.b * v.nmax * ld;
        unsigned onmask = 0;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int r0 = i0 + wr * 32 + mt * 8, c0 = j0 + wc * 32 + nt * 8;
                const bool on = (r0 < n) && (c0 < n) && !(diag && c0 > r0 + 7);
                onmask |= (on ? 1u : 0u) << (mt * 4 + nt);
            }
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        for (int it = 0; it < C.nk; ++it, ++cnt) {
            const unsigned slot = cnt % WS2_STAGES, ph = (cnt / WS2_STAGES) & 1u;
            mbar_wait(full + slot, ph);
            const double* as = As + slot * TK * T2PAD;
            const double* bs = diag ? as : Bs + slot * TK * T2PAD;
            if (onmask != 0 && !(dbg & 2)) {
                const int k4n = min(TK / 4, (k - it * TK + 3) >> 2);
#pragma unroll
                for (int k4 = 0; k4 < TK / 4; ++k4) {
                    if (k4 >= k4n) break;
                    double af[4], bf[4];
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) af[mt] = as[(k4 * 4 + q) * T2PAD + wr * 32 + mt * 8 + g];
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) bf[nt] = bs[(k4 * 4 + q) * T2PAD + wc * 32 + nt * 8 + g];
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt)
                            if (onmask & (1u << (mt * 4 + nt))) dmma(acc[mt][nt], af[mt], bf[nt]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + slot);
        }
        // ---- epilogue
        mbar_wait(pfull, pcnt & 1u);
        ++pcnt;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const int lr = wr * 32 + mt * 8 + g;
            const int gi = i0 + lr;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                if (!(onmask & (1u << (mt * 4 + nt))) || (dbg & 4)) continue;
                const int lc = wc * 32 + nt * 8 + 2 * q;
                const int gj = j0 + lc;
                double2 pv = make_double2(0.0, 0.0);
                if (gi < n && gj < ld) pv = *reinterpret_cast<const double2*>(Pt + lr * T2PAD + lc);
                const double c0 = pv.x - acc[mt][nt][0];
                const double c1 = pv.y - acc[mt][nt][1];
                if (C.col0 && wc == 0 && nt == 0) {
                    strip[lr * 9 + 2 * q] = c0;
                    strip[lr * 9 + 2 * q + 1] = c1;
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
        __syncwarp();
        if (lane == 0) mbar_arrive(pempty);
        if (C.col0) {
            asm volatile("bar.sync 1, 512;" ::: "memory");  // the 16 consumer warps only
            const double* __restrict__ Jn = jn_all + (size_t)C.b * 16;
            if (diag && tid < 8) {
                // tile (0,0): 8x8 corner with the 4x4 quaternion block, in place in shared memory
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
            if (tid < T2 && !(diag && tid < 8)) {
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
            asm volatile("bar.sync 1, 512;" ::: "memory");
        }
    }
}

// returns false if the shape does not fit this kernel (the caller falls back to the 64x64 kernels)
bool launch_downdate128(ekfslam_ctx* c, int sms) {
    DevView& v = c->v;
    const int nt = (v.nmax + T2 - 1) / T2;
    const int T = nt * (nt + 1) / 2;
    const long long total = (long long)T * v.B;
    const long long ctas = total < (long long)sms ? total : (long long)sms;
    const int M = (int)((total + ctas - 1) / ctas);
    const size_t sm = sizeof(double) * (2 * WS2_STAGES * TK * T2PAD + T2 * T2PAD + T2 * 9) +
                      sizeof(unsigned long long) * (2 * WS2_STAGES + 2) + sizeof(int2) * M + sizeof(unsigned) * T;
    if (sm > 227 * 1024) return false;
    static size_t cfg = 0;
    if (sm > cfg) {
        if (cudaFuncSetAttribute(k_downdate_ws128, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        cfg = sm;
    }
    k_downdate_ws128<<<(unsigned)ctas, WS2_THREADS, sm, c->stream>>>(v, v.jn, T, total, M, g_ekfslam_debug);
    return true;
}
