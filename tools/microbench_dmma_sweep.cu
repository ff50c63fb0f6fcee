// DMMA.8x8x4 issue-rate sweep: throughput vs resident warps per SM and independent accumulator
// chains per warp, with register-resident operands and with shared-memory-fed operands (the access
// pattern of k_downdate).  Answers: how many warps in the K loop does the fp64 tensor pipe need?
#include <cstdio>
#include <cuda_runtime.h>

template <int NCH, bool LDSFEED>
__global__ void k(double* out, int iters) {
    __shared__ double sm[16 * 68 * 2];
    for (int i = threadIdx.x; i < 16 * 68 * 2; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double c[NCH][2];
#pragma unroll
    for (int i = 0; i < NCH; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    double a = 1.0 + lane * 1e-9, b = 1.0 - lane * 1e-9;
    for (int it = 0; it < iters; ++it) {
        if (LDSFEED) {
            // 4 k4 steps per "stage": per k4 load NCH/2 A frags + 2 B frags like the real kernel (4x2 tiles)
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
                double af[4], bf[2];
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) af[mt] = sm[(k4 * 4 + q) * 68 + mt * 8 + g];
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) bf[nt] = sm[16 * 68 + (k4 * 4 + q) * 68 + nt * 8 + g];
#pragma unroll
                for (int i = 0; i < NCH; ++i)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(af[i % 4]), "d"(bf[(i / 4) % 2]));
            }
        } else {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
#pragma unroll
                for (int i = 0; i < NCH; ++i)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                 : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NCH, bool LDSFEED>
static double run(double* out, int sms, int warps_per_sm, int iters) {
    const int threads = 32 * (warps_per_sm < 8 ? warps_per_sm : 8);
    const int bps = warps_per_sm <= 8 ? 1 : warps_per_sm / 8;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a);
        k<NCH, LDSFEED><<<sms * bps, threads>>>(out, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (r >= 1 && ms < best) best = ms;
    }
    const double dmma = (double)sms * warps_per_sm * iters * 4.0 * NCH;
    return dmma * 512.0 / (best * 1e-3) / 1e12;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
    const int iters = 2000;
    printf("{\"gpu\": \"%s\", \"unit\": \"TFLOP/s fp64 DMMA\", \"rows\": [\n", prop.name);
    const int ws[] = {4, 8, 16, 24, 32, 64};
    for (int wi = 0; wi < 6; ++wi) {
        const int w = ws[wi];
        printf("  {\"warps_per_sm\": %d, \"reg_1chain\": %.2f, \"reg_2chain\": %.2f, \"reg_4chain\": %.2f, \"reg_8chain\": %.2f, "
               "\"lds_8chain\": %.2f}%s\n", w, run<1, false>(out, sms, w, iters), run<2, false>(out, sms, w, iters),
               run<4, false>(out, sms, w, iters), run<8, false>(out, sms, w, iters), run<8, true>(out, sms, w, iters),
               wi < 5 ? "," : "");
    }
    printf("]}\n");
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}
