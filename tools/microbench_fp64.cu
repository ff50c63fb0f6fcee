// Measures the fp64 ceilings of the GPU it runs on (MEASURED_PEAKS.json has no fp64 figure):
//   * DFMA  — dependent-chain-free vector fp64 FMA throughput
//   * DMMA  — mma.sync.aligned.m8n8k4.row.col.f64 throughput (SASS DMMA.8x8x4)
//   * copy  — double2 streaming copy (read + write bytes), as a cross-check of hbm_gbs
// Prints one JSON object.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_fp64 microbench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters) {
    double a[16];
    const double x = 1.0000001, y = 1e-9 * threadIdx.x;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = i + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dmma(double* out, int iters) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_copy(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = in[i];
}

static float best_ms(void (*launch)(void*), void* arg, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < reps + 2; ++r) {
        cudaEventRecord(a);
        launch(arg);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (r >= 2 && ms < best) best = ms;
    }
    return best;
}

struct Args { double* out; int iters; int blocks; int threads; const double2* in; double2* o2; size_t n; };
static void l_dfma(void* p) { Args* a = (Args*)p; k_dfma<<<a->blocks, a->threads>>>(a->out, a->iters); }
static void l_dmma(void* p) { Args* a = (Args*)p; k_dmma<<<a->blocks, a->threads>>>(a->out, a->iters); }
static void l_copy(void* p) { Args* a = (Args*)p; k_copy<<<a->blocks, a->threads>>>(a->in, a->o2, a->n); }

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    Args a;
    a.threads = 256; a.blocks = sms * 8; a.iters = 4096;
    cudaMalloc(&a.out, sizeof(double) * a.blocks * a.threads);
    const float ms_f = best_ms(l_dfma, &a, 8);
    const double dfma_tf = 2.0 * 16.0 * a.iters * (double)a.blocks * a.threads / (ms_f * 1e-3) / 1e12;
    const float ms_m = best_ms(l_dmma, &a, 8);
    const double dmma_tf = 2.0 * 8 * 8 * 4 * 8.0 * a.iters * (double)a.blocks * (a.threads / 32) / (ms_m * 1e-3) / 1e12;
    const size_t n = (size_t)1 << 28;  // 4 GiB in, 4 GiB out
    double2 *in, *o2;
    cudaMalloc(&in, n * sizeof(double2)); cudaMalloc(&o2, n * sizeof(double2));
    cudaMemset(in, 0, n * sizeof(double2));
    a.in = in; a.o2 = o2; a.n = n; a.blocks = sms * 16; a.threads = 512;
    const float ms_c = best_ms(l_copy, &a, 6);
    const double copy_gbs = 2.0 * n * sizeof(double2) / (ms_c * 1e-3) / 1e9;
    // sustained: the same DMMA / DFMA kernels back to back for ~2 s each (power / clock steady state)
    a.threads = 256; a.blocks = sms * 8; a.iters = 4096;
    double sus[2];
    for (int which = 0; which < 2; ++which) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int reps = which == 0 ? (int)(2000.0f / ms_m) : (int)(2000.0f / ms_f);
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) { if (which == 0) l_dmma(&a); else l_dfma(&a); }
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double per = ms / reps;
        sus[which] = which == 0 ? 2.0 * 8 * 8 * 4 * 8.0 * a.iters * (double)a.blocks * (a.threads / 32) / (per * 1e-3) / 1e12
                                : 2.0 * 16.0 * a.iters * (double)a.blocks * a.threads / (per * 1e-3) / 1e12;
    }
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_khz_max\": %d, \"dfma_tflops\": %.3f, \"dmma_m8n8k4_tflops\": %.3f, "
           "\"copy_gbs\": %.1f, \"dfma_ms\": %.4f, \"dmma_ms\": %.4f, \"copy_ms\": %.4f, "
           "\"dmma_sustained_2s_tflops\": %.3f, \"dfma_sustained_2s_tflops\": %.3f}\n",
           prop.name, sms, clk, dfma_tf, dmma_tf, copy_gbs, ms_f, ms_m, ms_c, sus[0], sus[1]);
    cudaError_t e = cudaDeviceSynchronize();
    return e == cudaSuccess ? 0 : 1;
}
