// Microbenchmark: issue rate of scalar FADD/FMUL/FFMA against the packed FADD2/FMUL2/FFMA2 forms on sm_100a.
// Each thread runs `iters` iterations over 8 independent chains; reports warp-instructions per clock per SM
// and "fp32 lane-ops per clock per SM". Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o packed packed_f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b)
{
    float x[CHAINS * 2];
#pragma unroll
    for (int c = 0; c < CHAINS * 2; ++c) x[c] = a + (float)threadIdx.x * 1e-3f + c;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (MODE == 0) {  // scalar add, 2 per chain pair
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[2 * c]) : "f"(b));
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[2 * c + 1]) : "f"(b));
            } else if (MODE == 1) {  // packed add
                unsigned long long v, w;
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(v) : "f"(x[2 * c]), "f"(x[2 * c + 1]));
                asm volatile("mov.b64 %0, {%1,%1};" : "=l"(w) : "f"(b));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(w));
                asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(x[2 * c]), "=f"(x[2 * c + 1]) : "l"(v));
            } else if (MODE == 2) {  // scalar mul
                asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[2 * c]) : "f"(b));
                asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[2 * c + 1]) : "f"(b));
            } else if (MODE == 3) {  // packed mul
                unsigned long long v, w;
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(v) : "f"(x[2 * c]), "f"(x[2 * c + 1]));
                asm volatile("mov.b64 %0, {%1,%1};" : "=l"(w) : "f"(b));
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(w));
                asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(x[2 * c]), "=f"(x[2 * c + 1]) : "l"(v));
            } else if (MODE == 4) {  // scalar fma
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[2 * c]) : "f"(b), "f"(a));
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[2 * c + 1]) : "f"(b), "f"(a));
            } else {  // packed fma
                unsigned long long v, w, u;
                asm volatile("mov.b64 %0, {%1,%2};" : "=l"(v) : "f"(x[2 * c]), "f"(x[2 * c + 1]));
                asm volatile("mov.b64 %0, {%1,%1};" : "=l"(w) : "f"(b));
                asm volatile("mov.b64 %0, {%1,%1};" : "=l"(u) : "f"(a));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(w), "l"(u));
                asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(x[2 * c]), "=f"(x[2 * c + 1]) : "l"(v));
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS * 2; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char *name, int sms, int khz, float *d)
{
    const int iters = 4096, blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(d, 64, 1.0f, 1.0000001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, iters, 1.0f, 1.0000001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double lane_ops = (double)blocks * threads * iters * CHAINS * 2;  // fp32 results produced
    const double clocks = ms * 1e-3 * khz * 1e3;
    printf("{\"op\": \"%s\", \"ms\": %.4f, \"fp32_results_per_clk_per_sm\": %.1f}\n", name, ms, lane_ops / clocks / sms);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *d;
    cudaMalloc(&d, sizeof(float) * p.multiProcessorCount * 8 * 256);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, p.multiProcessorCount, khz);
    run<0>("FADD", p.multiProcessorCount, khz, d);
    run<1>("FADD2", p.multiProcessorCount, khz, d);
    run<2>("FMUL", p.multiProcessorCount, khz, d);
    run<3>("FMUL2", p.multiProcessorCount, khz, d);
    run<4>("FFMA", p.multiProcessorCount, khz, d);
    run<5>("FFMA2", p.multiProcessorCount, khz, d);
    return 0;
}
