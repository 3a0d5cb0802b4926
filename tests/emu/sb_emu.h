// CPU emulator for the CUDA kernels of scarplet_b200 — TEST INFRASTRUCTURE ONLY.
//
// Compiles the unmodified kernel source (scarplet_b200/csrc/*.cuh, *.cu) with
// g++ -DSB_EMU so the index logic of every kernel (FFT exchanges, Hermitian
// unpacking, tile/halo bookkeeping, masks, argmax) can be exercised on a host
// without a GPU.  Each CUDA thread of a block is a ucontext fiber; a block
// barrier is a yield to the block scheduler; blocks are spread over OS threads.
// Nothing here is product code: scarplet_b200/_lib.py never opens the emulator
// library, and the parity claims are made by the `-m gpu` tests on a B200.
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline float4 make_float4(float x, float y, float z, float w) {
    float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r;
}
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
struct alignas(16) double4 { double x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
static inline double4 make_double4(double x, double y, double z, double w) {
    double4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r;
}

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace sbemu {

struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    int tid = 0;
    bool done = false;
    int wait_id = -1;      // barrier the fiber is parked at (-1: runnable)
    int wait_count = 0;    // threads that barrier expects (0: every live thread of the block)
};

struct Block {
    int bx = 0, by = 0, bz = 0;
    dim3 grid;
    void* smem = nullptr;
    ucontext_t sched;
    Fiber* cur = nullptr;
    const std::function<void()>* fn = nullptr;
    int arrived[17] = {0};   // bar.arrive counts per named barrier (threads that signalled and went on)
};

inline Block*& current() {
    static thread_local Block* b = nullptr;
    return b;
}

inline void fiber_entry() {
    Block* b = current();
    (*b->fn)();
    b->cur->done = true;
    swapcontext(&b->cur->ctx, &b->sched);
}

constexpr size_t kStack = 256 * 1024;

inline void run_blocks(dim3 grid, int nthreads, size_t smem_bytes,
                       const std::function<void()>& fn, std::atomic<long>& next) {
    std::vector<Fiber> fibers(nthreads);
    for (auto& f : fibers) f.stack = (char*)std::malloc(kStack);
    void* smem = nullptr;
    if (posix_memalign(&smem, 128, smem_bytes + 128) != 0) std::abort();
    Block blk;
    blk.grid = grid;
    blk.smem = smem;
    blk.fn = &fn;
    current() = &blk;
    const long total = (long)grid.x * grid.y * grid.z;
    for (;;) {
        long b = next.fetch_add(1);
        if (b >= total) break;
        blk.bx = (int)(b % grid.x);
        blk.by = (int)((b / grid.x) % grid.y);
        blk.bz = (int)(b / ((long)grid.x * grid.y));
        std::memset(smem, 0xCD, smem_bytes);  // poison: shared memory is uninitialised
        std::memset(blk.arrived, 0, sizeof(blk.arrived));
        for (int t = 0; t < nthreads; ++t) {
            Fiber& f = fibers[t];
            f.tid = t;
            f.done = false;
            f.wait_id = -1;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = kStack;
            f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        }
        // run every runnable fiber to its next barrier (or to the end), then open the barriers
        // whose expected number of threads has arrived
        for (;;) {
            int alive = 0, ran = 0;
            for (int t = 0; t < nthreads; ++t) {
                Fiber& f = fibers[t];
                if (f.done) continue;
                ++alive;
                if (f.wait_id >= 0) continue;
                blk.cur = &f;
                swapcontext(&blk.sched, &f.ctx);
                ++ran;
            }
            if (alive == 0) break;
            int live = 0, opened = 0;
            int waiting[17] = {0};
            for (int t = 0; t < nthreads; ++t) {
                Fiber& f = fibers[t];
                if (f.done) continue;
                ++live;
                if (f.wait_id >= 0) ++waiting[f.wait_id];
            }
            int used[17] = {0};
            for (int t = 0; t < nthreads; ++t) {
                Fiber& f = fibers[t];
                if (f.done || f.wait_id < 0) continue;
                const int expect = f.wait_count > 0 ? f.wait_count : live;
                if (waiting[f.wait_id] + blk.arrived[f.wait_id] >= expect) {
                    used[f.wait_id] = expect - waiting[f.wait_id];      // arrivals this phase consumed
                    f.wait_id = -1 - 100 - f.wait_id; ++opened;   // mark, open below
                }
            }
            for (int i = 0; i < 17; ++i) blk.arrived[i] -= used[i] > 0 ? used[i] : 0;
            for (int t = 0; t < nthreads; ++t)
                if (!fibers[t].done && fibers[t].wait_id < -1) fibers[t].wait_id = -1;
            if (ran == 0 && opened == 0) std::abort();     // barrier deadlock in the kernel under test
        }
    }
    current() = nullptr;
    std::free(smem);
    for (auto& f : fibers) std::free(f.stack);
}

inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& fn) {
    const long total = (long)grid.x * grid.y * grid.z;
    if (total <= 0) return;
    int nthreads = (int)(block.x * block.y * block.z);
    int workers = (int)std::thread::hardware_concurrency();
    if (const char* e = std::getenv("SB_EMU_WORKERS")) workers = std::atoi(e);
    workers = (int)std::max<long>(1, std::min<long>(workers, total));
    std::atomic<long> next(0);
    if (workers == 1) {
        run_blocks(grid, nthreads, smem_bytes, fn, next);
        return;
    }
    std::vector<std::thread> pool;
    for (int w = 0; w < workers; ++w)
        pool.emplace_back([&] { run_blocks(grid, nthreads, smem_bytes, fn, next); });
    for (auto& t : pool) t.join();
}

}  // namespace sbemu

#define SB_GLOBAL static void
#define SB_DEVICE inline
#define SB_HOSTDEV inline
#define SB_CONSTEXPR constexpr
#define SB_LAUNCH_BOUNDS(t, b)
#define SB_RESTRICT __restrict__

SB_DEVICE int sb_tid() { return sbemu::current()->cur->tid; }
SB_DEVICE int sb_bx() { return sbemu::current()->bx; }
SB_DEVICE int sb_by() { return sbemu::current()->by; }
SB_DEVICE int sb_bz() { return sbemu::current()->bz; }
SB_DEVICE int sb_nbx() { return (int)sbemu::current()->grid.x; }
SB_DEVICE int sb_nby() { return (int)sbemu::current()->grid.y; }
// named barrier: `count` threads of the block meet at barrier `id` (1..15); id 0 with count 0
// is the block-wide barrier over the threads that are still running
SB_DEVICE void sb_bar(int id, int count) {
    sbemu::Block* b = sbemu::current();
    b->cur->wait_id = id;
    b->cur->wait_count = count;
    swapcontext(&b->cur->ctx, &b->sched);
}
// bar.arrive: signal barrier `id` and carry on
SB_DEVICE void sb_bar_arrive(int id, int) { sbemu::current()->arrived[id]++; }
SB_DEVICE void sb_sync() { sb_bar(0, 0); }
SB_DEVICE void* sb_shared() { return sbemu::current()->smem; }
template <typename T> SB_DEVICE T sb_ldg(const T* p) { return *p; }
SB_DEVICE void sb_prefetch_l2(const void*) {}
SB_DEVICE float2 sb_ld_stream(const float2* p) { return *p; }
SB_DEVICE float4 sb_ld_stream(const float4* p) { return *p; }
SB_DEVICE float4 sb_ld_shared_soon(const float4* p) { return *p; }
SB_DEVICE void sb_st_stream(float4* p, float4 v) { *p = v; }
SB_DEVICE void sb_st_pair(float4* p, float4 a, float4 b) { p[0] = a; p[1] = b; }
SB_DEVICE void sb_st_pair(double4* p, double4 a, double4 b) { p[0] = a; p[1] = b; }
SB_DEVICE void sb_ld_sector(const float4* p, float4& a, float4& b) { a = p[0]; b = p[1]; }
SB_DEVICE float sb_fdiv_fast(float a, float b) { return a / b; }

// built with -ffp-contract=off, so these stay separate IEEE operations
SB_DEVICE double sb_mul(double a, double b) { return a * b; }
SB_DEVICE double sb_add(double a, double b) { return a + b; }
SB_DEVICE double sb_sub(double a, double b) { return a - b; }
SB_DEVICE double sb_div(double a, double b) { return a / b; }

typedef void* sb_stream_t;

#define SB_LAUNCH(kern, grid, block, smem, stream, ...) \
    sbemu::launch((grid), (block), (smem), [=]() { kern(__VA_ARGS__); })

inline int sb_rt_malloc(void** p, size_t n) {
    if (posix_memalign(p, 256, n ? n : 256) != 0) return 2;
    std::memset(*p, 0xFF, n);   // poison (NaN): device memory is uninitialised
    return 0;
}
inline int sb_rt_free(void* p) { std::free(p); return 0; }
inline int sb_rt_h2d(void* d, const void* h, size_t n, sb_stream_t) { std::memcpy(d, h, n); return 0; }
inline int sb_rt_d2h(void* h, const void* d, size_t n, sb_stream_t) { std::memcpy(h, d, n); return 0; }
inline int sb_rt_d2d(void* d, const void* s, size_t n, sb_stream_t) { std::memcpy(d, s, n); return 0; }
inline int sb_rt_memset(void* d, int v, size_t n, sb_stream_t) { std::memset(d, v, n); return 0; }
inline int sb_rt_sync(sb_stream_t) { return 0; }
inline int sb_rt_mem_info(size_t* free_b, size_t* total_b) { *free_b = *total_b = (size_t)2 << 30; return 0; }
inline int sb_rt_last_error() { return 0; }
typedef int sb_event_t;
inline int sb_rt_event_create(sb_event_t* e) { *e = 0; return 0; }
inline int sb_rt_event_destroy(sb_event_t) { return 0; }
inline int sb_rt_event_record(sb_event_t, sb_stream_t) { return 0; }
inline float sb_rt_event_ms(sb_event_t, sb_event_t) { return 0.f; }
inline const char* sb_rt_error_string(int) { return "emulator"; }
