// Throughput of packed FP32 (FFMA2/FADD2) vs scalar FFMA/FADD on sm_100a.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scratch/ubench_f32x2 scratch/ubench_f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up(u64 v, float& a, float& b){ asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b){ u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    u64 p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
    const u64 ss = pk(s, s * 0.5f);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {           // 16 scalar FFMA
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], s, 0.25f);
        } else if (MODE == 1) {    // 8 FFMA2 (same flops)
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], ss, ss);
        } else if (MODE == 2) {    // 16 scalar FADD
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = a[i] + s;
        } else if (MODE == 3) {    // 8 FADD2
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = add2(p[i], ss);
        } else if (MODE == 4) {    // 8 FADD2 with swizzle+neg operands
#pragma unroll
            for (int i = 0; i < 8; ++i) { float x, y; up(p[(i + 1) & 7], x, y); p[i] = add2(p[i], pk(y, -x)); }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) { float x, y; up(p[i], x, y); acc += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, int instr_per_iter) {
    int dev_sms; cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, dev_sms * 8 * 256 * sizeof(float));
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps = 1; warps <= 8; warps *= 2) {      // CTAs per SM (8 warps each)
        k<MODE><<<dev_sms * warps, 256>>>(out, 100, 1.0001f);
        cudaEventRecord(e0);
        k<MODE><<<dev_sms * warps, 256>>>(out, iters, 1.0001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double winstr = (double)warps * 8 * iters * instr_per_iter;     // warp-instr per SM
        printf("%-28s ctas/sm %d  %.3f ms  %.2f warp-instr/ns/SM  (%.2f per clk at %d MHz nominal)\n", name, warps, ms,
               winstr / (ms * 1e6), winstr / (ms * 1e6) / (clk * 1e-6), clk / 1000);
    }
    cudaFree(out);
}
int main() {
    run<0>("FFMA x16", 16);
    run<1>("FFMA2 x8", 8);
    run<2>("FADD x16", 16);
    run<3>("FADD2 x8", 8);
    run<4>("FADD2 x8 swz/neg", 8);
    return 0;
}
