// Covariance downdate  P <- J P J' - W'W  (mc/update.m:13-22) on 64x64 tiles of the lower triangle.
//
// W holds ktot rows per filter: the rows of one update, or of a deferred li update followed by the hi
// update ([W_li; W_hi], one pass over P per frame instead of two).  Every row already carries the
// normalisation Jacobians (W <- W J', k_wfix), so J = blkdiag(I3, Jn, I) only has to be applied to the P
// tile itself: columns 3..6 of tile column 0 and rows 3..6 of tile (0,0).  0.5P + 0.5P' (:14) is the identity
// because the lower triangle is authoritative and every tile is stored together with its mirror image.
//
// Two kernels, same arithmetic and summation order (k ascending, DMMA m8n8k4):
//   k_downdate_ws   (default) persistent, warp-specialised: a producer warp streams the W panels with bulk
//                   asynchronous copies (cp.async.bulk -> SASS UBLKCP) completing on a 4-stage mbarrier ring
//                   that keeps running across tile boundaries; 8 consumer warps issue the DMMAs.
//   k_downdate_tile (EKFSLAM_DOWNDATE=tile, and the fallback for shapes whose tile list does not fit the
//                   persistent kernel's shared memory) one CTA per tile, cp.async (LDGSTS) ring.
// Measured (tools/dbg_downdate.py ablation): the kernel is bound by the bytes moved between
// L2 and the SMs (W panel re-reads + P tile + mirrored stores), not by the fp64 tensor pipe.
#include <cstdlib>
#include <cstring>
#include "model.cuh"
#include "tc_common.cuh"

#define WS_STAGES 4
#define WS_CONSUMERS 8
#define WS_THREADS ((WS_CONSUMERS + 1) * 32)

// ---- shared epilogue pieces ---------------------------------------------------------------------------
// bit mt*2+nt of the mask: the warp's 8x8 DMMA tile (mt, nt) has work (inside n x n, not strictly above the
// diagonal of a diagonal tile)
__device__ __forceinline__ unsigned tile_mask(int i0, int j0, int wr, int wc, int n, bool diag) {
    unsigned m = 0;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const int r0 = i0 + wr * 32 + mt * 8, c0 = j0 + wc * 16 + nt * 8;
            const bool on = (r0 < n) && (c0 < n) && !(diag && c0 > r0 + 7);
            m |= (on ? 1u : 0u) << (mt * 2 + nt);
        }
    return m;
}

// P tile -> accumulator layout (lane g = lane>>2, q = lane&3 holds rows .. + g, columns .. + 2q, 2q+1)
__device__ __forceinline__ void load_p_frags(const double* __restrict__ P, int ld, int n, int i0, int j0, int wr, int wc,
                                             int g, int q, double (&pf)[4][2][2]) {
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
}

// J on the P tile of tile column 0: the warps holding columns 0..7 (wc == 0, nt == 0) exchange them through
// a [64][9] strip; 64 threads apply [c3 c4 c5 c6] <- [c3 c4 c5 c6] Jn' per row and, for tile (0,0), rows 3..6 <-
// Jn * rows of the 8x8 corner.  `sync` is the barrier over the participating warps.
template <typename Sync>
__device__ __forceinline__ void apply_j_col0(double* strip, const double* __restrict__ Jn, bool diag, int wr, int wc, int g,
                                             int q, int tid, double (&pf)[4][2][2], Sync sync) {
    if (wc == 0) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            strip[(wr * 32 + mt * 8 + g) * 9 + 2 * q] = pf[mt][0][0];
            strip[(wr * 32 + mt * 8 + g) * 9 + 2 * q + 1] = pf[mt][0][1];
        }
    }
    sync();
    if (tid < TM) {
        const int r = tid;
        const double c3 = strip[r * 9 + 3], c4 = strip[r * 9 + 4], c5 = strip[r * 9 + 5], c6 = strip[r * 9 + 6];
#pragma unroll
        for (int a = 0; a < 4; ++a)
            strip[r * 9 + 3 + a] = c3 * Jn[a * 4 + 0] + c4 * Jn[a * 4 + 1] + c5 * Jn[a * 4 + 2] + c6 * Jn[a * 4 + 3];
    }
    sync();
    if (diag) {
        if (tid < 8) {
            const int c = tid;
            const double r3 = strip[3 * 9 + c], r4 = strip[4 * 9 + c], r5 = strip[5 * 9 + c], r6 = strip[6 * 9 + c];
#pragma unroll
            for (int a = 0; a < 4; ++a)
                strip[(3 + a) * 9 + c] = Jn[a * 4 + 0] * r3 + Jn[a * 4 + 1] * r4 + Jn[a * 4 + 2] * r5 + Jn[a * 4 + 3] * r6;
        }
        sync();
    }
    if (wc == 0) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            pf[mt][0][0] = strip[(wr * 32 + mt * 8 + g) * 9 + 2 * q];
            pf[mt][0][1] = strip[(wr * 32 + mt * 8 + g) * 9 + 2 * q + 1];
        }
    }
    sync();  // the strip is free again
}

// C = P - acc, stored with its mirror image straight from the accumulator fragments (lanes with equal q cover
// 64-byte runs).  Diagonal tiles: the lower triangle is authoritative, the upper one its mirror.
__device__ __forceinline__ void store_tile(double* __restrict__ P, int ld, int n, int i0, int j0, int wr, int wc, int g, int q,
                                           bool diag, unsigned onmask, const double (&pf)[4][2][2],
                                           const double (&acc)[4][2][2]) {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const int gi = i0 + wr * 32 + mt * 8 + g;
        if (gi >= n) continue;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            if (!(onmask & (1u << (mt * 2 + nt)))) continue;
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
}

// ---- one CTA per tile, cp.async ring ------------------------------------------------------------------
__global__ void __launch_bounds__(256, 3) k_downdate_tile(DevView v) {
    extern __shared__ __align__(16) double dsm[];
    const int b = blockIdx.y;
    const int k = v.ktot[b];
    if (k == 0) return;
    const int n = v.nstate[b];
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
    double* As = dsm;                            // [NSTAGE][TK][TPAD]
    double* Bs = dsm + NSTAGE * TK * TPAD;       // [NSTAGE][TK][TPAD]
    double* strip = Bs + NSTAGE * TK * TPAD;     // [64][9]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wr = warp >> 2, wc = warp & 3, g = lane >> 2, q = lane & 3;

    double pf[4][2][2];
    load_p_frags(P, ld, n, i0, j0, wr, wc, g, q, pf);
    const int nk = (k + TK - 1) / TK;
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
    const unsigned onmask = tile_mask(i0, j0, wr, wc, n, diag);

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
        if (onmask == 0) continue;
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
    if (tj == 0)
        apply_j_col0(strip, v.jn + (size_t)b * 16, diag, wr, wc, g, q, tid, pf, [] { __syncthreads(); });
    store_tile(P, ld, n, i0, j0, wr, wc, g, q, diag, onmask, pf, acc);
}

// ---- persistent, warp-specialised ---------------------------------------------------------------------
__global__ void __launch_bounds__(WS_THREADS, 2) k_downdate_ws(DevView v, int T, long long total, int M) {
    extern __shared__ __align__(16) double dsm[];
    double* As = dsm;                                   // [WS_STAGES][TK][TPAD]
    double* Bs = dsm + WS_STAGES * TK * TPAD;           // [WS_STAGES][TK][TPAD]
    double* strip = Bs + WS_STAGES * TK * TPAD;         // [64][9]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(strip + TM * 9);   // [WS_STAGES]
    unsigned long long* empty = full + WS_STAGES;                                        // [WS_STAGES]
    int2* meta = reinterpret_cast<int2*>(empty + WS_STAGES);                             // [M]  {ktot, n}
    unsigned* lut = reinterpret_cast<unsigned*>(meta + M);                               // [T]  ti<<16|tj
    const int ld = v.ld, kmax = v.kmax;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long G = gridDim.x;
    for (int m = tid; m < M; m += blockDim.x) {
        const long long t = blockIdx.x + (long long)m * G;
        int2 kn = make_int2(0, 0);
        if (t < total) { const int b = (int)(t / T); kn = make_int2(v.ktot[b], v.nstate[b]); }
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
                if (lane == 0) mbar_arrive_expect_tx(full + slot, (unsigned)nvalid * (bytesA + (L.diag ? 0u : bytesB)));
                const bool mine = !(isB && L.diag);
                if (mine && pr < nvalid) {
                    bulk_g2s(dst, W + (size_t)(t0 + pr) * ld + c0, rowbytes, full + slot);
                } else if (mine && pr < ((nvalid + 3) & ~3)) {
                    for (int c = 0; c < TM; ++c) dst[c] = 0.0;  // rows between k and the next multiple of 4
                }
                if (lane != 0) mbar_arrive(full + slot);
            }
        }
        return;
    }

    // ================= consumer warps: 2 (rows) x 4 (cols), 32 x 16 each =================
    const int wr = warp >> 2, wc = warp & 3, g = lane >> 2, q = lane & 3;
    unsigned cnt = 0;
    for (int cm = 0; cm < Mreal; ++cm) {
        const DTile C = decode_tile(meta, lut, cm, Mreal, blockIdx.x + (long long)cm * G, T);
        if (C.nk == 0) continue;
        const int n = C.n, k = C.k, i0 = C.i0, j0 = C.j0;
        const bool diag = C.diag;
        double* __restrict__ P = v.P + (size_t)C.b * v.nmax * ld;
        double pf[4][2][2];
        load_p_frags(P, ld, n, i0, j0, wr, wc, g, q, pf);
        const unsigned onmask = tile_mask(i0, j0, wr, wc, n, diag);
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
        if (C.col0)
            apply_j_col0(strip, v.jn + (size_t)C.b * 16, diag, wr, wc, g, q, tid, pf,
                         [] { asm volatile("bar.sync 1, 256;" ::: "memory"); });  // the 8 consumer warps only
        store_tile(P, ld, n, i0, j0, wr, wc, g, q, diag, onmask, pf, acc);
    }
}

void launch_downdate(ekfslam_ctx* c, int slot) {
    DevView& v = c->v;
    static int mode = -1, sms = 0;
    if (mode < 0) {
        const char* e = getenv("EKFSLAM_DOWNDATE");
        mode = (e && !strcmp(e, "tile")) ? 0 : 1;  // default: warp-specialised persistent kernel
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    }
    const int nt = (v.nmax + TM - 1) / TM;
    const int T = nt * (nt + 1) / 2;
    KScope ks(c, slot);
    if (mode == 1) {
        const long long total = (long long)T * v.B;
        const long long ctas = total < (long long)sms * 2 ? total : (long long)sms * 2;
        const int M = (int)((total + ctas - 1) / ctas);
        const size_t ws_sm = sizeof(double) * (2 * WS_STAGES * TK * TPAD + TM * 9) + sizeof(unsigned long long) * 2 * WS_STAGES +
                             sizeof(int2) * M + sizeof(unsigned) * T;
        if (ws_sm <= 110 * 1024) {  // two CTAs per SM
            static size_t ws_cfg = 0;
            if (ws_sm > ws_cfg) {
                cudaFuncSetAttribute(k_downdate_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ws_sm);
                ws_cfg = ws_sm;
            }
            k_downdate_ws<<<(unsigned)ctas, WS_THREADS, ws_sm, c->stream>>>(v, T, total, M);
            return;
        }
    }
    const size_t sm = sizeof(double) * (2 * NSTAGE * TK * TPAD + TM * 9);
    static bool cfg = false;
    if (!cfg) {
        cudaFuncSetAttribute(k_downdate_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        cfg = true;
    }
    dim3 gd(T, v.B);
    k_downdate_tile<<<gd, 256, sm, c->stream>>>(v);
}
