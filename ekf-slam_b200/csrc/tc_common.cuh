// Shared building blocks of the tensor-core kernels (k_w, k_downdate*): DMMA wrapper, cp.async ring
// helpers, mbarrier / bulk-copy wrappers, lower-triangle tile decoding.
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------
// fp64 tensor-core building blocks.  mma.sync m8n8k4 f64 (SASS DMMA.8x8x4) fragment layout, lane
// l, g = l>>2, q = l&3:  A[g][q],  B[q][g],  C[g][2q], C[g][2q+1].
// Both GEMMs below use 64x64 block tiles, 8 warps in a 2x4 grid, 32x16 per warp (4x2 DMMA tiles),
// K panels of 16 staged through a 3-deep cp.async (LDGSTS) ring.
// ---------------------------------------------------------------------------------------
#define TM 64
#define TK 16
#define TPAD 68    // row stride (doubles) of a K-major panel [TK][64]: stride % 16 == 4 -> conflict-free frags
#define APAD 20    // row stride of a row-major A panel [64][TK]: same property
#define NSTAGE 3

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gmem_src, int src_bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }


struct DTile {
    int b, i0, j0, k, n, nk;
    bool diag, col0;
};

// tile m of this CTA (global tile t = blockIdx.x + m*G): metadata comes from shared memory
// (meta[m] = {k, n} of the tile's filter, lut[e] = ti<<16|tj), never from a dependent global load
__device__ __forceinline__ DTile decode_tile(const int2* meta, const unsigned* lut, int m, int M, long long t, int T,
                                             int tile = TM) {
    DTile d;
    d.nk = 0; d.b = 0; d.i0 = 0; d.j0 = 0; d.k = 0; d.n = 0; d.diag = false; d.col0 = false;
    if (m >= M) return d;
    d.b = (int)(t / T);
    const unsigned e = lut[(int)(t - (long long)d.b * T)];
    const int ti = (int)(e >> 16), tj = (int)(e & 0xffffu);
    d.i0 = ti * tile; d.j0 = tj * tile;
    d.diag = (ti == tj); d.col0 = (tj == 0);
    const int2 kn = meta[m];
    d.k = kn.x; d.n = kn.y;
    d.nk = (d.i0 < d.n) ? (d.k + TK - 1) / TK : 0;
    return d;
}


__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(double* smem_dst, const double* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

