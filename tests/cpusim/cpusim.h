// tests/cpusim/cpusim.h -- TEST INFRASTRUCTURE ONLY.
//
// A small SIMT emulator: lets g++ compile the product's CUDA sources
// (nuts333_b200/csrc/*.cu, with -DNUTSB_CPUSIM) so that `pytest -m "not gpu"`
// can exercise the *kernel logic* on a box without a GPU, under ASan/TSan if
// wanted.  It is NOT a CPU fallback: the product library (libnutsb200.so) is
// built by nvcc only, has no CPU path, and nothing in the nuts333_b200 package
// can load the library built from this header (tests/cpusim/build_sim.py puts
// it under tests/cpusim/_build/).
//
// Model: blocks run one after another; the threads of a block are OS threads
// (so `__shared__` maps to `static`), __syncthreads() is a drop-out barrier,
// warp collectives rendezvous the lanes named in the mask.  1-D grids/blocks,
// blockDim a multiple of 32.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)
#define __restrict__
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static

struct dim3 { unsigned x = 1, y = 1, z = 1; dim3() {} dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

namespace cpusim {

struct Idx { unsigned x = 0, y = 0, z = 0; };

struct DynBarrier {
    std::mutex m; std::condition_variable cv; int expected = 0, arrived = 0; uint64_t gen = 0;
    void reset(int n) { expected = n; arrived = 0; }
    void arrive_wait() {
        std::unique_lock<std::mutex> l(m);
        uint64_t g = gen;
        if (++arrived >= expected) { arrived = 0; ++gen; cv.notify_all(); }
        else if (!cv.wait_for(l, std::chrono::seconds(60), [&] { return gen != g; })) {
            fprintf(stderr, "cpusim: __syncthreads() deadlock\n"); abort();
        }
    }
    void drop() {
        std::unique_lock<std::mutex> l(m);
        --expected;
        if (expected > 0 && arrived >= expected) { arrived = 0; ++gen; cv.notify_all(); }
    }
};

struct Slot { uint64_t v; int aux; };

struct WarpState {
    std::mutex m; std::condition_variable cv;
    unsigned cur_mask = 0, arrived = 0; uint64_t gen = 0;
    Slot in[32]; uint64_t out[32];
};

struct BlockState {
    int nthreads = 0;
    DynBarrier bar;         // __syncthreads
    DynBarrier endbar;      // all threads, between blocks
    std::vector<WarpState> warps;
};

extern thread_local Idx tl_threadIdx, tl_blockIdx;
extern thread_local BlockState *tl_block;
extern Idx g_blockDim, g_gridDim;
extern unsigned char *g_dyn_smem;      // blocks run one at a time: one buffer serves them all
void set_dyn_smem(size_t bytes);

template <class Combine>
static inline uint64_t collective(unsigned mask, uint64_t v, int aux, Combine combine)
{
    int lane = (int)(tl_threadIdx.x & 31);
    WarpState &w = tl_block->warps[tl_threadIdx.x >> 5];
    std::unique_lock<std::mutex> l(w.m);
    if (!(mask >> lane & 1)) { fprintf(stderr, "cpusim: lane %d not in its own mask %08x\n", lane, mask); abort(); }
    while (w.arrived != 0 && w.cur_mask != mask)
        if (w.cv.wait_for(l, std::chrono::seconds(60)) == std::cv_status::timeout) { fprintf(stderr, "cpusim: warp collective mask conflict\n"); abort(); }
    w.cur_mask = mask; w.in[lane].v = v; w.in[lane].aux = aux; w.arrived |= 1u << lane;
    if (w.arrived == mask) {
        for (int k = 0; k < 32; ++k) if (mask >> k & 1) w.out[k] = combine(w.in, mask, k);
        w.arrived = 0; ++w.gen; w.cv.notify_all();
    } else {
        uint64_t g = w.gen;
        if (!w.cv.wait_for(l, std::chrono::seconds(60), [&] { return w.gen != g; })) {
            fprintf(stderr, "cpusim: warp collective deadlock (mask %08x arrived %08x)\n", mask, w.arrived); abort();
        }
    }
    return w.out[lane];
}

template <class F>
static inline void launch(dim3 grid, dim3 block, F body)
{
    if (block.x % 32 != 0 || block.y != 1 || block.z != 1 || grid.y != 1 || grid.z != 1) {
        fprintf(stderr, "cpusim: only 1-D launches with blockDim %% 32 == 0\n"); abort();
    }
    if (grid.x == 0) return;
    BlockState st; st.nthreads = (int)block.x;
    st.warps = std::vector<WarpState>(block.x / 32);
    st.bar.reset((int)block.x); st.endbar.reset((int)block.x);
    g_blockDim.x = block.x; g_gridDim.x = grid.x;
    std::vector<std::thread> th;
    th.reserve(block.x);
    for (unsigned t = 0; t < block.x; ++t)
        th.emplace_back([&, t] {
            tl_block = &st;
            for (unsigned b = 0; b < grid.x; ++b) {
                tl_threadIdx.x = t; tl_blockIdx.x = b;
                body();
                st.bar.drop();              // a finished thread no longer takes part in __syncthreads
                st.endbar.arrive_wait();    // block ends when every thread has left
                if (t == 0) st.bar.reset((int)block.x);
                st.endbar.arrive_wait();
            }
        });
    for (auto &x : th) x.join();
}

} // namespace cpusim

#define threadIdx (cpusim::tl_threadIdx)
#define blockIdx  (cpusim::tl_blockIdx)
#define blockDim  (cpusim::g_blockDim)
#define gridDim   (cpusim::g_gridDim)

static inline void __syncthreads() { cpusim::tl_block->bar.arrive_wait(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu)
{ cpusim::collective(mask, 0, 0, [](const cpusim::Slot *, unsigned, int) { return (uint64_t)0; }); }
static inline unsigned __ballot_sync(unsigned mask, int pred)
{
    return (unsigned)cpusim::collective(mask, pred ? 1 : 0, 0, [](const cpusim::Slot *in, unsigned m, int) {
        uint64_t r = 0; for (int k = 0; k < 32; ++k) if ((m >> k & 1) && in[k].v) r |= 1ull << k; return r; });
}
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v)
{
    return (unsigned)cpusim::collective(mask, v, 0, [](const cpusim::Slot *in, unsigned m, int) {
        uint64_t r = 0; for (int k = 0; k < 32; ++k) if (m >> k & 1) r |= in[k].v; return r; });
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
static inline unsigned __match_any_sync(unsigned mask, unsigned long long v)
{
    return (unsigned)cpusim::collective(mask, v, 0, [](const cpusim::Slot *in, unsigned m, int me) {
        uint64_t r = 0; for (int k = 0; k < 32; ++k) if ((m >> k & 1) && in[k].v == in[me].v) r |= 1ull << k; return r; });
}
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
    static_assert(sizeof(T) <= 8, "shfl");
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
    int lane = (int)(threadIdx.x & 31);
    int s = (lane & ~(width - 1)) | (src & (width - 1));
    uint64_t r = cpusim::collective(mask, raw, s, [](const cpusim::Slot *in, unsigned m, int me) {
        int s2 = in[me].aux; return (m >> s2 & 1) ? in[s2].v : in[me].v; });
    T o; memcpy(&o, &r, sizeof(T)); return o;
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32)
{
    int lane = (int)(threadIdx.x & 31);
    int s = lane - (int)d;
    if (s < (lane & ~(width - 1))) s = lane;
    return __shfl_sync(mask, v, s, 32);
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32)
{
    int lane = (int)(threadIdx.x & 31);
    int s = lane + (int)d;
    if (s > (lane | (width - 1))) s = lane;
    return __shfl_sync(mask, v, s, 32);
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32)
{
    (void)width;
    return __shfl_sync(mask, v, (int)((threadIdx.x & 31) ^ (unsigned)x), 32);
}

static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s)
{ return (unsigned)((((uint64_t)hi << 32) | lo) >> (s & 31)); }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s)
{
    uint64_t t = ((uint64_t)b << 32) | a; unsigned r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned)((t >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned __vcmpeq4(unsigned a, unsigned b)
{
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) if (((a >> (8 * i)) & 0xff) == ((b >> (8 * i)) & 0xff)) r |= 0xffu << (8 * i);
    return r;
}

static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline int atomicMax(int *p, int v)
{ int o = __atomic_load_n(p, __ATOMIC_RELAXED); while (o < v && !__atomic_compare_exchange_n(p, &o, v, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
static inline unsigned atomicMax(unsigned *p, unsigned v)
{ unsigned o = __atomic_load_n(p, __ATOMIC_RELAXED); while (o < v && !__atomic_compare_exchange_n(p, &o, v, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// ---- the slice of the CUDA runtime the library's host code uses ------------
typedef int cudaError_t;
typedef void *cudaStream_t;
struct cpusim_event { std::chrono::steady_clock::time_point t; };
typedef cpusim_event *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaEventDefault = 0 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
static inline const char *cudaGetErrorString(cudaError_t e) { return e ? "cpusim error" : "no error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = 4; return cudaSuccess; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int *least, int *greatest) { *least = 0; *greatest = -5; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new cpusim_event; return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new cpusim_event; return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b)
{ *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }

#define NUTSB_LAUNCH(grid, block, stream, kern, ...) \
    cpusim::launch(dim3(grid), dim3(block), [&] { kern(__VA_ARGS__); })
#define NUTSB_LAUNCH_SMEM(grid, block, smem, stream, kern, ...) \
    do { cpusim::set_dyn_smem(smem); cpusim::launch(dim3(grid), dim3(block), [&] { kern(__VA_ARGS__); }); } while (0)
#define NUTSB_DYN_SMEM(name) unsigned char *name = cpusim::g_dyn_smem
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

#ifdef CPUSIM_IMPLEMENTATION
namespace cpusim {
thread_local Idx tl_threadIdx, tl_blockIdx;
thread_local BlockState *tl_block = nullptr;
Idx g_blockDim, g_gridDim;
unsigned char *g_dyn_smem = nullptr;
void set_dyn_smem(size_t bytes)
{
    static size_t cap = 0;
    if (bytes > cap) { free(g_dyn_smem); g_dyn_smem = (unsigned char *)aligned_alloc(64, (bytes + 63) & ~(size_t)63); cap = bytes; }
}
}
#endif
