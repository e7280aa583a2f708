// nutsb_common.cuh -- shared definitions for the sm_100a kernels of the NUTS
// message path.  Compiled by nvcc for the product; compiled by g++ against
// tests/cpusim/cpusim.h (-DNUTSB_CPUSIM) for the CPU kernel-logic tests only.
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef NUTSB_CPUSIM
#include "cpusim.h"
#else
#include <cuda_runtime.h>
#define NUTSB_LAUNCH(grid, block, stream, kern, ...) \
    kern<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define NUTSB_LAUNCH_SMEM(grid, block, smem, stream, kern, ...) \
    kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
// dynamic shared memory of the running block, 16-byte aligned
#define NUTSB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

#include "../../include/nutsb200.h"

typedef uint64_t           u64;
typedef int64_t            i64;
typedef uint32_t           u32;
typedef int32_t            i32;
typedef uint16_t           u16;
typedef uint8_t            u8;

#define NUTSB_FULL 0xffffffffu

// bring the 128-byte lines that hold [p, p+n) towards the SM ahead of their use (n <= 2048 + 3)
#ifdef NUTSB_CPUSIM
static inline void nutsb_prefetch(const void *, u32) {}
#else
__device__ __forceinline__ void nutsb_prefetch(const void *p, u32 n)
{
    const char *a = (const char *)((size_t)p & ~(size_t)127), *e = (const char *)p + n;
    for (; a < e; a += 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(a));
}
#endif

__device__ __forceinline__ void nutsb_add64(u64 *p, u64 v) { atomicAdd((unsigned long long *)p, (unsigned long long)v); }

// recipients that do not simply take every op of their room: they go through nutsb_class_delivers
#define NUTSB_UF_FILTERED (NUTSB_UF_LOGIN | NUTSB_UF_IGNALL | NUTSB_UF_IGNSHOUT | NUTSB_UF_CLONE | NUTSB_UF_REMOTE)

// Status bits raised by kernels (ctx->d_status), decoded on the host.
#define NUTSB_ST_TEXT_TOO_LONG 0x01u
#define NUTSB_ST_BAD_KIND      0x02u
#define NUTSB_ST_BAD_INDEX     0x04u
#define NUTSB_ST_HAS_LEVEL     0x08u   // not an error: batch contains write_level ops
#define NUTSB_ST_RENDER_MISMATCH 0x10u // internal consistency check failed
#define NUTSB_ST_BAD_OFFSETS   0x20u

// ---- the colour table, nuts333.h:237-255 ---------------------------------------
// 26x26 table indexed by the two command letters: value = index+1, 0 = not a
// command.  Built on the host (nutsb_lib.cu) and staged into shared memory by
// every kernel that parses '~XX'.
#define NUTSB_CODETAB_BYTES 676

__device__ __forceinline__ int nutsb_code(const u8 *tab, u8 a, u8 b)
{
    u32 ua = (u32)a - 'A', ub = (u32)b - 'A';
    if (ua < 26u && ub < 26u) return (int)tab[ua * 26u + ub] - 1;
    return -1;
}
// Length of colcode[k]: ESC [ d m (k<5) or ESC [ 3|4 d m.
__device__ __forceinline__ u32 nutsb_code_len(int k) { return k < 5 ? 4u : 5u; }

// Recipient filter for the ops that reach a whole class of users (room and
// level ops); the room match itself is established by bucketing.
//   ROOM : nuts333.c:1410-1415      LEVEL: nuts333.c:1379-1383
__device__ __forceinline__ bool nutsb_class_delivers(u32 cflags, u32 clevel, u32 kind, u32 oflags, i32 target)
{
    if (cflags & (NUTSB_UF_LOGIN | NUTSB_UF_REMOTE)) return false;   // remote: framed for its netlink instead (nutsb_q_*)
    if (kind == NUTSB_OP_ROOM) {
        if (cflags & NUTSB_UF_CLONE) return false;              // c:1416: relayed to the owner instead (nutsb_q_*)
        if ((cflags & NUTSB_UF_IGNALL) && !(oflags & NUTSB_OF_FORCE_LISTEN)) return false;
        if ((cflags & NUTSB_UF_IGNSHOUT) && (oflags & NUTSB_OF_SHOUT)) return false;
        return true;
    }
    if (cflags & NUTSB_UF_CLONE) return false;
    return (oflags & NUTSB_OF_ABOVE) ? ((i32)clevel >= target) : ((i32)clevel <= target);
}

// ---- device-wide exclusive scan (three kernels, u64) ----------------------------
#define NUTSB_SCAN_THREADS 256
#define NUTSB_SCAN_ITEMS   16
#define NUTSB_SCAN_TILE    (NUTSB_SCAN_THREADS * NUTSB_SCAN_ITEMS)

// Block-wide exclusive scan of one value per thread (256 threads).  Returns the
// exclusive prefix; *total gets the block sum.  Contains __syncthreads().
__device__ __forceinline__ u64 nutsb_block_excl_scan(u64 v, u64 *total)
{
    __shared__ u64 s_warp[NUTSB_SCAN_THREADS / 32];
    __shared__ u64 s_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 inc = v;
    for (int d = 1; d < 32; d <<= 1) {
        u64 t = __shfl_up_sync(NUTSB_FULL, inc, d);
        if (lane >= d) inc += t;
    }
    __syncthreads();                       // protects s_warp/s_total against the previous call
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        u64 w = lane < NUTSB_SCAN_THREADS / 32 ? s_warp[lane] : 0;
        u64 winc = w;
        for (int d = 1; d < NUTSB_SCAN_THREADS / 32; d <<= 1) {
            u64 t = __shfl_up_sync(NUTSB_FULL, winc, d);
            if (lane >= d) winc += t;
        }
        if (lane < NUTSB_SCAN_THREADS / 32) s_warp[lane] = winc - w;
        if (lane == NUTSB_SCAN_THREADS / 32 - 1) s_total = winc;
    }
    __syncthreads();
    *total = s_total;
    return s_warp[warp] + inc - v;
}

template <class In>
__global__ void __launch_bounds__(NUTSB_SCAN_THREADS)
k_scan_reduce(In in, i64 n_host, const u32 *n_dev, u64 *block_sums)
{
    const i64 n = n_dev ? (i64)*n_dev : n_host;
    const i64 base = (i64)blockIdx.x * NUTSB_SCAN_TILE + (i64)threadIdx.x * NUTSB_SCAN_ITEMS;
    u64 s = 0;
    for (int k = 0; k < NUTSB_SCAN_ITEMS; ++k) { i64 i = base + k; if (i < n) s += in(i); }
    u64 total;
    (void)nutsb_block_excl_scan(s, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// One block: exclusive scan of the block sums in place; total -> sums[nb].
__global__ void __launch_bounds__(NUTSB_SCAN_THREADS)
k_scan_sums(u64 *sums, i64 nb)
{
    u64 carry = 0;
    for (i64 base = 0; base < nb; base += NUTSB_SCAN_THREADS) {
        i64 i = base + threadIdx.x;
        u64 v = i < nb ? sums[i] : 0, total;
        u64 ex = nutsb_block_excl_scan(v, &total);
        if (i < nb) sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) sums[nb] = carry;
}

// fused != 0: block_sums holds the raw per-block sums (k_scan_sums was skipped) and every
// block adds up the sums of the blocks before it by itself (cheap while there are few blocks).
template <class In, class Out>
__global__ void __launch_bounds__(NUTSB_SCAN_THREADS)
k_scan_apply(In in, Out out, i64 n_host, const u32 *n_dev, const u64 *block_sums, int fused)
{
    const i64 n = n_dev ? (i64)*n_dev : n_host;
    const i64 base = (i64)blockIdx.x * NUTSB_SCAN_TILE + (i64)threadIdx.x * NUTSB_SCAN_ITEMS;
    u64 v[NUTSB_SCAN_ITEMS], s = 0;
    for (int k = 0; k < NUTSB_SCAN_ITEMS; ++k) { i64 i = base + k; v[k] = i < n ? in(i) : 0; s += v[k]; }
    u64 before;
    if (fused) {
        u64 part = 0;
        for (u32 j = threadIdx.x; j < blockIdx.x; j += NUTSB_SCAN_THREADS) part += block_sums[j];
        (void)nutsb_block_excl_scan(part, &before);
    } else before = block_sums[blockIdx.x];
    u64 total;
    u64 ex = nutsb_block_excl_scan(s, &total) + before;
    for (int k = 0; k < NUTSB_SCAN_ITEMS; ++k) { i64 i = base + k; if (i <= n) out(i, ex); ex += v[k]; }   // out(n) = total
}

// ---- the same scan in ONE kernel: decoupled look-back ------------------------------------------
// Every block takes the next tile (a ticket, so that tiles are started in order), scans it, publishes
// its aggregate, adds up the aggregates / inclusive prefixes of the tiles before it as they appear and
// writes the tile's exclusive prefixes: the input is read once and there is one launch instead of two
// or three.  The per-tile state lives in a context buffer that is never cleared: a state word carries
// the epoch of the scan it belongs to (`epoch`, bumped by the host per scan), and the block that
// finishes last resets the ticket.  state[t] = { flag (epoch << 2 | 1 aggregate, | 2 inclusive), aggregate, inclusive }.
// (the aggregate and the inclusive prefix have a word each: a reader that has seen "aggregate" must still find the
// aggregate there when the writer has moved on to "inclusive")
struct ScanState { unsigned long long flag, aggregate, inclusive, pad_; };

// ITEMS elements per thread: 16 for a long input, 1 for a short one (<= 16k elements: more blocks, each thread waits
// for one input value instead of sixteen -- the per-user stream length is a chain of five dependent loads).
template <class In, class Out, int ITEMS>
__global__ void __launch_bounds__(NUTSB_SCAN_THREADS)
k_scan1(In in, Out out, i64 n_host, const u32 *n_dev, ScanState *state, u32 *ticket, u32 epoch, u32 nb)
{
    __shared__ u32 s_tile;
    __shared__ u64 s_before;
    const i64 n = n_dev ? (i64)*n_dev : n_host;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const u32 tile = s_tile;
    // the tile goes through shared memory both ways: element tid + 256 k is read / written by thread tid (coalesced),
    // elements 16 tid .. 16 tid + 15 are scanned by thread tid (one pad word per 16 keeps both patterns off each
    // other's banks).  With each thread reading and writing its own 16 consecutive elements straight from / to global
    // memory every load touched 32 sectors for 4 bytes each, and the 3M-element scan took 40 us.
    __shared__ u64 s_val[ITEMS > 1 ? NUTSB_SCAN_THREADS * (ITEMS + 1) : 1];
    const i64 tbase = (i64)tile * (NUTSB_SCAN_THREADS * ITEMS);
    u64 v[ITEMS], s = 0;
    if (ITEMS > 1) {
        for (int k = 0; k < ITEMS; ++k) {                                  // (all the loads first, then the stores)
            const i64 i = tbase + (i64)k * NUTSB_SCAN_THREADS + threadIdx.x;
            v[k] = i < n ? in(i) : 0;
        }
        for (int k = 0; k < ITEMS; ++k) {
            const u32 j = (u32)k * NUTSB_SCAN_THREADS + threadIdx.x;
            s_val[j + j / ITEMS] = v[k];
        }
        __syncthreads();
        for (int k = 0; k < ITEMS; ++k) { v[k] = s_val[threadIdx.x * (ITEMS + 1) + k]; s += v[k]; }
    } else {
        const i64 i = tbase + threadIdx.x;
        v[0] = s = i < n ? in(i) : 0;
    }
    u64 total;
    u64 ex = nutsb_block_excl_scan(s, &total);
    {
        // look-back by the whole block, 256 tiles at a time: thread i looks at tile (hi - 1 - i); the window's sum is
        // everything up to and including the nearest tile that already knows its inclusive prefix.  (One warp and 32
        // tiles a step was the first form; by itself the wider window measured no faster -- the blocks were waiting
        // for their own uncoalesced loads and stores, see above -- it is kept because it bounds the number of trips:
        // 3 instead of 23 for the 733 tiles of a 3M-element scan.)
        __shared__ u32 s_incl[NUTSB_SCAN_THREADS / 32];
        __shared__ u64 s_part[NUTSB_SCAN_THREADS / 32];
        const unsigned long long fa = ((unsigned long long)epoch << 2) | 1ull, fi = ((unsigned long long)epoch << 2) | 2ull;
        volatile ScanState *st = state;
        const int lane = (int)(threadIdx.x & 31), warp = (int)(threadIdx.x >> 5);
        u64 before = 0;                                                    // (thread 0's copy is the one that counts)
        if (tile > 0) {
            if (threadIdx.x == 0) { st[tile].aggregate = total; __threadfence(); st[tile].flag = fa; }   // my aggregate, for the tiles after me
            for (u32 hi = tile; hi > 0; ) {
                const bool have = threadIdx.x < hi;
                const u32 t = have ? hi - 1 - threadIdx.x : 0;
                unsigned long long f = fi;
                if (have) do { f = st[t].flag; } while (f != fa && f != fi);              // (tiles before me were started before me)
                __threadfence();
                const u32 incl = __ballot_sync(NUTSB_FULL, have && f == fi);
                if (lane == 0) s_incl[warp] = incl;
                __syncthreads();
                int stop = NUTSB_SCAN_THREADS;                                            // nearest inclusive tile in the window
                for (int w = NUTSB_SCAN_THREADS / 32 - 1; w >= 0; --w) if (s_incl[w]) stop = 32 * w + __ffs((int)s_incl[w]) - 1;
                u64 pv = (have && (int)threadIdx.x <= stop) ? (f == fi ? st[t].inclusive : st[t].aggregate) : 0;
                for (int d = 16; d; d >>= 1) pv += __shfl_xor_sync(NUTSB_FULL, pv, d);
                if (lane == 0) s_part[warp] = pv;
                __syncthreads();
                if (threadIdx.x == 0) for (int w = 0; w < NUTSB_SCAN_THREADS / 32; ++w) before += s_part[w];
                if (stop < NUTSB_SCAN_THREADS) break;                                     // (the same in every thread)
                hi = hi > NUTSB_SCAN_THREADS ? hi - NUTSB_SCAN_THREADS : 0;
            }
        }
        if (threadIdx.x == 0) { st[tile].inclusive = before + total; __threadfence(); st[tile].flag = fi; s_before = before; }
    }
    __syncthreads();
    ex += s_before;
    if (ITEMS > 1) {
        for (int k = 0; k < ITEMS; ++k) { s_val[threadIdx.x * (ITEMS + 1) + k] = ex; ex += v[k]; }
        __syncthreads();
        for (int k = 0; k < ITEMS; ++k) {
            const u32 j = (u32)k * NUTSB_SCAN_THREADS + threadIdx.x;
            const i64 i = tbase + j;
            if (i <= n) out(i, s_val[j + j / ITEMS]);                                     // out(n) = total
        }
    } else {
        const i64 i = tbase + threadIdx.x;
        if (i <= n) out(i, ex);
    }
    // the last block to get here puts the ticket back for the next scan
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); if (atomicAdd(ticket + 1, 1u) == nb - 1) { ticket[0] = 0; ticket[1] = 0; } }
}
