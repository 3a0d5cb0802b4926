// Micro-benchmark (developer tool, not product code): what does a strided 16/32/64-byte
// access cost on B200 through the LSU versus through TMA?  It decides how the
// inverse-column planes travel from k_conv_cols to k_fit_rows.
//
// Matrix: ROWS x PITCH bytes (4096 rows, pitch 2056*16 B like gbuf).  A CTA owns a group of
// adjacent 16-byte columns and touches every row of them.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o ubench_scatter ubench_scatter.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int ROWS = 4096;
constexpr int COLS16 = 2048;            // 16-byte columns used
constexpr int PITCH16 = 2056;           // pitch in 16-byte units
constexpr int NMAT = 8;                 // matrices (templates) -> 8 * 135 MB, beyond L2

// ---- LSU stores: W = adjacent 16-byte columns written per row by one warp instruction ----
// CTA = 256 threads; CTA handles W columns; thread layout: lane%W -> column, (warp*32+lane)/W -> row slot
template <int W>
__global__ void k_store(float4* base, int ncolgroups) {
    const int cg = blockIdx.x % ncolgroups, mat = blockIdx.x / ncolgroups;
    float4* m = base + (size_t)mat * ROWS * PITCH16 + (size_t)cg * W;
    const int c = threadIdx.x % W, r0 = threadIdx.x / W;
    constexpr int RPI = 256 / W;        // rows per CTA-wide instruction
    const float4 v = make_float4(threadIdx.x, blockIdx.x, 1.f, 2.f);
#pragma unroll 4
    for (int r = r0; r < ROWS; r += RPI) m[(size_t)r * PITCH16 + c] = v;
}
template <int W>
__global__ void k_load(const float4* base, int ncolgroups, float* sink) {
    const int cg = blockIdx.x % ncolgroups, mat = blockIdx.x / ncolgroups;
    const float4* m = base + (size_t)mat * ROWS * PITCH16 + (size_t)cg * W;
    const int c = threadIdx.x % W, r0 = threadIdx.x / W;
    constexpr int RPI = 256 / W;
    float acc = 0.f;
#pragma unroll 8
    for (int r = r0; r < ROWS; r += RPI) {
        float4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(m + (size_t)r * PITCH16 + c));
        acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 1.2345e-30f) sink[0] = acc;
}
// 32 bytes per lane (256-bit access): one full sector per lane, 2 adjacent columns
__global__ void k_store256(float4* base, int ncolgroups) {
    const int cg = blockIdx.x % ncolgroups, mat = blockIdx.x / ncolgroups;
    float4* m = base + (size_t)mat * ROWS * PITCH16 + (size_t)cg * 2;
    const float a = threadIdx.x, b = blockIdx.x;
#pragma unroll 4
    for (int r = threadIdx.x; r < ROWS; r += 256) {
        asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(m + (size_t)r * PITCH16), "f"(a), "f"(b), "f"(a), "f"(b), "f"(a), "f"(b), "f"(a), "f"(b) : "memory");
    }
}
__global__ void k_load256(const float4* base, int ncolgroups, float* sink) {
    const int cg = blockIdx.x % ncolgroups, mat = blockIdx.x / ncolgroups;
    const float4* m = base + (size_t)mat * ROWS * PITCH16 + (size_t)cg * 2;
    float acc = 0.f;
#pragma unroll 8
    for (int r = threadIdx.x; r < ROWS; r += 256) {
        float v0, v1, v2, v3, v4, v5, v6, v7;
        asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v0), "=f"(v1), "=f"(v2), "=f"(v3), "=f"(v4), "=f"(v5), "=f"(v6), "=f"(v7) : "l"(m + (size_t)r * PITCH16));
        acc += v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
    }
    if (acc == 1.2345e-30f) sink[0] = acc;
}

// ---- TMA: tile [ROWS_PER_BOX rows][W*16 bytes] between shared memory and the matrix ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int W, int RB>
__global__ void k_tma_store(const __grid_constant__ CUtensorMap tm, int ncolgroups) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int cg = blockIdx.x % ncolgroups, mat = blockIdx.x / ncolgroups;
    // fill the staging tile once (contents irrelevant)
    float4* s = (float4*)smem;
    for (int i = threadIdx.x; i < RB * W; i += blockDim.x) s[i] = make_float4(i, 1.f, 2.f, 3.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int r = 0; r < ROWS; r += RB) {
            const int c0 = cg * W * 4;                // inner coordinate in floats
            const int c1 = mat * ROWS + r;            // row
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                         ::"l"(&tm), "r"(c0), "r"(c1), "r"(smem_u32(smem)) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <int W, int RB>
__global__ void k_tma_load(const __grid_constant__ CUtensorMap tm, int ncolgroups, float* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar[2];
    const int cg = blockIdx.x % ncolgroups, mat = blockIdx.x / ncolgroups;
    constexpr uint32_t BYTES = RB * W * 16;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float acc = 0.f;
    if (threadIdx.x == 0) {
        // two-deep pipeline: issue box i+1 while waiting for box i
        const int nbox = ROWS / RB;
        auto issue = [&](int i) {
            const int b = i & 1;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[b])), "r"(BYTES) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(smem + (size_t)b * BYTES)), "l"(&tm), "r"(cg * W * 4), "r"(mat * ROWS + i * RB), "r"(smem_u32(&bar[b])) : "memory");
        };
        issue(0);
        for (int i = 0; i < nbox; ++i) {
            if (i + 1 < nbox) issue(i + 1);
            const int b = i & 1;
            const uint32_t parity = (i >> 1) & 1;
            uint32_t done = 0;
            while (!done) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(smem_u32(&bar[b])), "r"(parity) : "memory");
            }
            acc += ((float*)(smem + (size_t)b * BYTES))[0];
        }
    }
    if (acc == 1.2345e-30f) sink[0] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    return (EncodeFn)fn;
}

static CUtensorMap make_map(void* base, int W, int RB) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)PITCH16 * 4, (cuuint64_t)ROWS * NMAT};     // floats, rows
    cuuint64_t strides[1] = {(cuuint64_t)PITCH16 * 16};                            // bytes between rows
    cuuint32_t box[2] = {(cuuint32_t)W * 4, (cuuint32_t)RB};
    cuuint32_t es[2] = {1, 1};
    CUresult r = get_encode()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d (W=%d RB=%d)\n", (int)r, W, RB); exit(1); }
    return tm;
}

template <typename F>
static void timeit(const char* name, double bytes, F&& launch) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int it = 0; it < 3; ++it) {
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    printf("%-28s %8.3f ms  %8.1f GB/s\n", name, best, bytes / best / 1e6);
}

int main() {
    float4* base; float* sink;
    const size_t bytes_total = (size_t)NMAT * ROWS * PITCH16 * 16;
    CK(cudaMalloc(&base, bytes_total));
    CK(cudaMemset(base, 0, bytes_total));
    CK(cudaMalloc(&sink, 16));
    const double moved = (double)NMAT * ROWS * COLS16 * 16;
    printf("matrix %d x %d x 16 B, %d matrices, %.2f GB moved per test\n", ROWS, COLS16, NMAT, moved / 1e9);

#define LSU(W) \
    timeit("LSU store W=" #W " (x16B/row)", moved, [&] { k_store<W><<<NMAT * (COLS16 / W), 256>>>(base, COLS16 / W); }); \
    timeit("LSU load  W=" #W " (x16B/row)", moved, [&] { k_load<W><<<NMAT * (COLS16 / W), 256>>>(base, COLS16 / W, sink); });
    LSU(1) LSU(2) LSU(4) LSU(8) LSU(32)
    timeit("LSU store 256-bit/lane", moved, [&] { k_store256<<<NMAT * (COLS16 / 2), 256>>>(base, COLS16 / 2); });
    timeit("LSU load  256-bit/lane", moved, [&] { k_load256<<<NMAT * (COLS16 / 2), 256>>>(base, COLS16 / 2, sink); });

#define TMA(W, RB) { \
    CUtensorMap tm = make_map(base, W, RB); \
    CK(cudaFuncSetAttribute(k_tma_store<W, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * RB * W * 16)); \
    CK(cudaFuncSetAttribute(k_tma_load<W, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * RB * W * 16)); \
    timeit("TMA store W=" #W " RB=" #RB, moved, [&] { k_tma_store<W, RB><<<NMAT * (COLS16 / W), 128, 2 * RB * W * 16>>>(tm, COLS16 / W); }); \
    timeit("TMA load  W=" #W " RB=" #RB, moved, [&] { k_tma_load<W, RB><<<NMAT * (COLS16 / W), 128, 2 * RB * W * 16>>>(tm, COLS16 / W, sink); }); }
    TMA(1, 256) TMA(2, 256) TMA(4, 256) TMA(8, 128) TMA(8, 256)
    printf("done\n");
    return 0;
}
