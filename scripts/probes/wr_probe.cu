// Write-pattern probe (development aid, not part of the product): how fast can B200 absorb
// many concurrent sequential write streams, as a function of how far apart they are?
//   every warp writes `piece`-byte pieces (16-byte stores, 2 in flight per lane);
//   mode 0: all warps sweep memory together (piece p of warp w at (p*nwarps + w)*piece)
//   mode 1: warp w owns a private contiguous region of `len` bytes (nwarps regions in flight)
//   mode 2: like 1, but the regions in flight at any time are confined to a window of `win` warps'
//           worth of regions: warp w writes region (w % win) + win * (w / win) ... i.e. blocks
//           are launched in waves; here simply grid = win warps per launch, several launches.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o wr_probe wr_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;

__global__ void k_sweep(uint4 *out, u64 total_vec, u64 nwarps)
{
    const u64 w = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; const int lane = threadIdx.x & 31;
    const uint4 v = make_uint4(lane, 1, 2, 3);
    for (u64 p = w; p * 64 + 64 <= total_vec; p += nwarps) { out[p * 64 + lane] = v; out[p * 64 + 32 + lane] = v; }
}
// region r of `len_vec` 16-byte vectors; warp w handles regions w, w+nwarps, ...
__global__ void k_regions(uint4 *out, u64 n_regions, u64 len_vec, u64 nwarps, u64 misalign_vec)
{
    const u64 w = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; const int lane = threadIdx.x & 31;
    const uint4 v = make_uint4(lane, 1, 2, 3);
    for (u64 r = w; r < n_regions; r += nwarps) {
        uint4 *d = out + r * len_vec + misalign_vec;
        u64 i = lane;
        for (; i + 32 < len_vec - misalign_vec; i += 64) { d[i] = v; d[i + 32] = v; }
    }
}
// "ours": item = (room, tile); block's 8 warps write run (tile) of every user of the room: user stream = tiles*run bytes
__global__ void k_items(uint4 *out, u64 tiles, u64 users, u64 run_vec)
{
    const u64 room = blockIdx.x / tiles, t = blockIdx.x % tiles; const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint4 v = make_uint4(lane, 1, 2, 3);
    for (u64 u = warp; u < users; u += blockDim.x / 32) {
        uint4 *d = out + ((room * users + u) * tiles + t) * run_vec;
        u64 i = lane;
        for (; i + 32 < run_vec; i += 64) { d[i] = v; d[i + 32] = v; }
        if (i < run_vec) d[i] = v;
    }
}
// tile-major order within a group of G rooms: consecutive blocks = same tile index, different rooms/users closer in time
__global__ void k_items_u(uint4 *out, u64 tiles, u64 users, u64 run_vec, u64 uchunk)
{
    const u64 chunks = users / uchunk;
    const u64 room = blockIdx.x / (tiles * chunks), rem = blockIdx.x % (tiles * chunks), c = rem / tiles, t = rem % tiles;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint4 v = make_uint4(lane, 1, 2, 3);
    for (u64 u = c * uchunk + warp; u < (c + 1) * uchunk; u += blockDim.x / 32) {
        uint4 *d = out + ((room * users + u) * tiles + t) * run_vec;
        u64 i = lane;
        for (; i + 32 < run_vec; i += 64) { d[i] = v; d[i + 32] = v; }
        if (i < run_vec) d[i] = v;
    }
}

// contiguous runs of run_vec 16-byte vectors (not a multiple of a line): ALIGN = 1 stores start at the run start;
// ALIGN = 2 / 8: a first partial store brings the body to a 32-byte / 128-byte boundary
template <int ALIGN>
__global__ void k_items_al(uint4 *out, u64 tiles, u64 users, u64 run_vec)
{
    const u64 room = blockIdx.x / tiles, t = blockIdx.x % tiles; const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint4 v = make_uint4(lane, 1, 2, 3);
    for (u64 u = warp; u < users; u += blockDim.x / 32) {
        const u64 base = ((room * users + u) * tiles + t) * run_vec;
        uint4 *d = out + base;
        u64 n = run_vec;
        const u64 pre = (ALIGN - (base % ALIGN)) % ALIGN;
        if (pre) { if ((u64)lane < pre) d[lane] = v; d += pre; n -= pre; }
        u64 i = lane;
        for (; i + 32 < n; i += 64) { d[i] = v; d[i + 32] = v; }
        if (i < n) d[i] = v;
    }
}

template <class F> static float timeit(F f)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f(); cudaDeviceSynchronize();
    float best = 1e9f;
    for (int i = 0; i < 5; ++i) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}
int main()
{
    const u64 total = 5ull << 30; uint4 *out; cudaMalloc(&out, total + (1 << 20));
    const u64 total_vec = total / 16;
    const int blocks = 148 * 5, threads = 256; const u64 nwarps = (u64)blocks * threads / 32;
    float ms = timeit([&] { k_sweep<<<blocks, threads>>>(out, total_vec, nwarps); });
    printf("sweep (all warps together, 1 KB pieces): %.1f GB/s\n", total / ms / 1e6);
    for (u64 len : {1024ull, 4096ull, 16384ull, 65536ull, 262144ull, 1048576ull}) {
        const u64 len_vec = len / 16, n_regions = total_vec / len_vec;
        ms = timeit([&] { k_regions<<<blocks, threads>>>(out, n_regions, len_vec, nwarps, 0); });
        printf("regions of %7llu B, consecutive regions to consecutive warps (%llu warps): %.1f GB/s\n", len, nwarps, total / ms / 1e6);
    }
    for (u64 len : {4096ull, 65536ull}) {
        const u64 len_vec = len / 16, n_regions = total_vec / len_vec;
        ms = timeit([&] { k_regions<<<blocks, threads>>>(out, n_regions, len_vec, nwarps, 3); });
        printf("regions of %7llu B, start misaligned by 48 B: %.1f GB/s\n", len, total / ms / 1e6);
    }
    {   // ours: 100 rooms x 100 users x 59 tiles x 9.3 KB runs
        const u64 tiles = 59, users = 100, rooms = 96, run_vec = 9344 / 16;
        const double bytes = (double)rooms * users * tiles * run_vec * 16;
        ms = timeit([&] { k_items<<<rooms * tiles, threads>>>(out, tiles, users, run_vec); });
        printf("items (room,tile) x 100 users, 9.3 KB runs, room-major: %.1f GB/s\n", bytes / ms / 1e6);
        for (u64 uc : {50ull, 25ull, 10ull}) {
            ms = timeit([&] { k_items_u<<<rooms * tiles * (users / uc), threads>>>(out, tiles, users, run_vec, uc); });
            printf("items (room, chunk of %llu users, tile), 9.3 KB runs: %.1f GB/s\n", uc, bytes / ms / 1e6);
        }
        for (u64 rv : {585ull, 586ull, 588ull, 181ull}) {
            const double b2 = (double)rooms * users * tiles * rv * 16;
            ms = timeit([&] { k_items_al<1><<<rooms * tiles, threads>>>(out, tiles, users, rv); });
            printf("runs of %llu B back to back, stores from the run start: %.1f GB/s\n", rv * 16, b2 / ms / 1e6);
            ms = timeit([&] { k_items_al<2><<<rooms * tiles, threads>>>(out, tiles, users, rv); });
            printf("runs of %llu B back to back, body aligned to 32 B: %.1f GB/s\n", rv * 16, b2 / ms / 1e6);
            ms = timeit([&] { k_items_al<8><<<rooms * tiles, threads>>>(out, tiles, users, rv); });
            printf("runs of %llu B back to back, body aligned to 128 B: %.1f GB/s\n", rv * 16, b2 / ms / 1e6);
        }
    }
    return 0;
}
