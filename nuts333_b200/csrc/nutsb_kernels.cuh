// nutsb_kernels.cuh -- hand-written sm_100a kernels of the write path:
// measure -> bucket by room -> prefix sums -> events -> stream offsets ->
// render + fan-out.  See DESIGN.md for the data layout and the roofline of each.
//
// Reference semantics: nuts333.c:1291-1366 (write_user), :1372-1385
// (write_level), :1401-1429 (write_room_except).
#pragma once
#include "nutsb_common.cuh"

// Device views -------------------------------------------------------------------

struct OpsView {                 // the batch, device pointers (nutsb_ops)
    i64 n;
    const u8  *text;
    const u64 *toff;
    const u8  *kind;
    const i32 *target;
    const i32 *except_user;
    const u8  *flags;
    const i32 *gate;             // may be null
    const u8  *verdict;          // may be null
};

struct PopView {                 // the population, built by nutsb_set_users
    i32 n_users, n_rooms, n_rooms_tot;   // n_rooms_tot = n_rooms + 1 (the last is "no room")
    const i32 *user_room;        // [U]  room' in [0, n_rooms_tot)
    const i32 *user_cls;         // [U]  global class id
    const i32 *user_slot;        // [U]  position in (room, class, index) order
    const i32 *slot_user;        // [U]
    const u8  *slot_cf;          // [U]  recipient flags by slot (NUTSB_UF_*)
    const u8  *slot_lv;          // [U]  level by slot
    const i32 *room_slot_off;    // [Rt+1]
    const i32 *room_cls_off;     // [Rt+1]
    const u8  *cls_flags;        // [K]
    const u8  *cls_level;        // [K]
    const u8  *codetab;          // [676]
};

#ifndef NUTSB_TILE_OPS
#define NUTSB_TILE_OPS   128     // room-list ops per fan-out tile (power of two; k_fanout runs 2 threads per op)
#endif
#define NUTSB_UCHUNK     (NUTSB_TILE_OPS < 128 ? NUTSB_TILE_OPS : 128)   // recipients per fan-out work item
#define NUTSB_MAX2(a, b) ((a) > (b) ? (a) : (b))
#define NUTSB_TEXT_CAP   NUTSB_MAX2(2048 + 64, 88 * NUTSB_TILE_OPS)   // staged source bytes per (sub)tile   (>= 2000+6)
#define NUTSB_ON_CAP     NUTSB_MAX2(12288, 96 * NUTSB_TILE_OPS)       // rendered bytes per (sub)tile, colour on  (>= 6*2000+4)
#define NUTSB_OFF_CAP    NUTSB_MAX2(4096, 80 * NUTSB_TILE_OPS)        // rendered bytes per (sub)tile, colour off (>= 2*2000)
#define NUTSB_EV_CAP     256     // events of a tile's recipients prefetched into shared memory
#define NUTSB_RUN_CAP    512     // planned copy runs per (sub)tile and recipient chunk

// ---- A. measure ------------------------------------------------------------------
// Rendered length of every op for both colour settings, liveness (gate), validation
// and the number of room lists the op enters.  A warp owns 32 consecutive ops = one
// contiguous byte range of the packed text: it is staged into shared memory with
// coalesced 16-byte loads, then scanned WORD-parallel (lane l takes words l, l+32,
// ...: balanced whatever the string lengths): exact SWAR masks find the '\n' and
// '~' bytes, their positions go to a per-warp list, and the list is then resolved
// 32 entries at a time (owner string by binary search over the 33 offsets, '/~'
// escape and command lookup in the shared-memory table) into per-string counters.
#define NUTSB_MEASURE_THREADS 256
#define NUTSB_MEASURE_WARP_BYTES 4096
#define NUTSB_MEASURE_LIST 256

// 0x80 in every byte of y that is zero, exact (no borrow between bytes)
__device__ __forceinline__ u32 nutsb_zero_bytes(u32 y)
{
    return ~(((y & 0x7f7f7f7fu) + 0x7f7f7f7fu) | y | 0x7f7f7f7fu);
}

__global__ void __launch_bounds__(NUTSB_MEASURE_THREADS)
k_measure(OpsView ops, PopView pop, u32 *len_on, u32 *len_off, u32 *nrep, u32 *status)
{
    __shared__ __align__(16) u8 s_stage[NUTSB_MEASURE_THREADS / 32][NUTSB_MEASURE_WARP_BYTES + 32];
    __shared__ u32 s_list[NUTSB_MEASURE_THREADS / 32][NUTSB_MEASURE_LIST];
    __shared__ u32 s_p0[NUTSB_MEASURE_THREADS / 32][33];
    __shared__ u32 s_ca[NUTSB_MEASURE_THREADS / 32][32], s_cb[NUTSB_MEASURE_THREADS / 32][32];
    __shared__ u32 s_nlist[NUTSB_MEASURE_THREADS / 32];
    __shared__ u8 s_tab[NUTSB_CODETAB_BYTES];
    for (int i = threadIdx.x; i < NUTSB_CODETAB_BYTES; i += blockDim.x) s_tab[i] = pop.codetab[i];
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const i64 wbase = (i64)blockIdx.x * NUTSB_MEASURE_THREADS + warp * 32;
    if (wbase >= ops.n) return;                    // whole warp leaves together
    const i64 wend = (wbase + 32 < ops.n) ? wbase + 32 : ops.n;
    const u32 nops = (u32)(wend - wbase);
    const u64 b0 = ops.toff[wbase], b1 = ops.toff[wend];
    // 16-byte aligned window around the warp's bytes (any cudaMalloc'd buffer is
    // readable up to the next 16-byte boundary)
    const u8 *pa = (const u8 *)((size_t)(ops.text + b0) & ~(size_t)15);
    const u64 span = (u64)((ops.text + b1) - pa);
    u8 *stage = s_stage[warp];

    // -- per-op facts (lane = op)
    const i64 i = wbase + lane;
    const bool have = i < ops.n;
    u64 o0 = b1, o1 = b1;
    if (have) { o0 = ops.toff[i]; o1 = ops.toff[i + 1]; }
    const bool bad = have && o1 < o0;
    const bool toolong = have && !bad && o1 - o0 > NUTSB_MAX_TEXT;
    bool live = have && !bad && !toolong;
    if (live && ops.gate && ops.gate[i] >= 0) {    // a gated-off op is neither measured nor bucketed
        const bool v = ops.verdict[ops.gate[i]] != 0;
        live = ((ops.flags[i] & NUTSB_OF_GATE_IF_SET) != 0) == v;
    }
    const u32 any_bad = __ballot_sync(NUTSB_FULL, bad);
    const u32 livemask = __ballot_sync(NUTSB_FULL, live);
    const bool staged = !any_bad && b1 >= b0 && span <= NUTSB_MEASURE_WARP_BYTES;
    const u32 n = (have && !bad) ? (u32)(o1 - o0) : 0;
    u32 nl = 0, drops = 0, m4 = 0, m5 = 0;

    if (staged && livemask) {
        const u32 nvec = (u32)((span + 15) >> 4);
        for (u32 v = lane; v < nvec; v += 32) *(uint4 *)(stage + 16 * v) = __ldg((const uint4 *)pa + v);
        s_p0[warp][lane] = (u32)((ops.text + o0) - pa);
        if (lane == 0) { s_p0[warp][32] = (u32)((ops.text + b1) - pa); s_nlist[warp] = 0; }
        s_ca[warp][lane] = 0; s_cb[warp][lane] = 0;
        __syncwarp();
        // -- word-parallel scan: record the positions of '\n' and '~'
        const u32 r0 = (u32)((ops.text + b0) - pa), r1 = (u32)span;
        for (u32 w = (r0 >> 2) + lane; w < (r1 + 3) >> 2; w += 32) {
            const u32 x = *(const u32 *)(stage + 4 * w);
            u32 vm = 0x80808080u;                     // bytes of the word inside the warp's range
            if (4 * w < r0) vm &= 0xffffffffu << (8 * (r0 - 4 * w));
            if (4 * w + 4 > r1) vm &= 0xffffffffu >> (8 * (4 * w + 4 - r1));
            const u32 mn = nutsb_zero_bytes(x ^ 0x0a0a0a0au) & vm, mt = nutsb_zero_bytes(x ^ 0x7e7e7e7eu) & vm;
            u32 m = mn | mt;
            while (m) {
                const u32 bit = (u32)__ffs((int)m) - 1;
                m &= m - 1;
                const u32 slot = atomicAdd(&s_nlist[warp], 1u);
                if (slot < NUTSB_MEASURE_LIST) s_list[warp][slot] = (4 * w + (bit >> 3)) | ((mn >> bit & 1u) << 31);
            }
        }
        __syncwarp();
        const u32 cnt = s_nlist[warp];
        if (cnt <= NUTSB_MEASURE_LIST) {
            // -- resolve the list, 32 entries at a time
            for (u32 e = lane; e < cnt; e += 32) {
                const u32 ent = s_list[warp][e];
                const u32 j = ent & 0x7fffffffu;
                u32 lo = 0, hi = nops;                 // owner: last q with p0[q] <= j
                while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (s_p0[warp][mid] <= j) lo = mid; else hi = mid; }
                const u32 q = lo;
                if (!(livemask >> q & 1u)) continue;
                const u32 q0 = s_p0[warp][q], q1 = s_p0[warp][q + 1];
                if (ent >> 31) atomicAdd(&s_ca[warp][q], 1u);                               // '\n'
                else if (j > q0 && stage[j - 1] == '/') atomicAdd(&s_ca[warp][q], 0x10000u); // "/~": slash dropped
                else if (j + 2 < q1) {
                    const int k = nutsb_code(s_tab, stage[j + 1], stage[j + 2]);
                    if (k >= 0) atomicAdd(&s_cb[warp][q], k < 5 ? 1u : 0x10000u);
                }
            }
            __syncwarp();
            const u32 ca = s_ca[warp][lane], cb = s_cb[warp][lane];
            nl = ca & 0xffffu; drops = ca >> 16; m4 = cb & 0xffffu; m5 = cb >> 16;
        }
    }
    if (live && !(staged && s_nlist[warp] <= NUTSB_MEASURE_LIST)) {
        // strings too long for the staging window, or too many special bytes: byte loop
        const u8 *s = ops.text + o0;
        nl = drops = m4 = m5 = 0;
        for (u32 j = 0; j < n; ++j) {
            const u8 c = s[j];
            if (c == '\n') ++nl;
            else if (c == '~') {
                if (j > 0 && s[j - 1] == '/') ++drops;
                else if (j + 2 < n) {
                    int k = nutsb_code(s_tab, s[j + 1], s[j + 2]);
                    if (k >= 0) { if (k < 5) ++m4; else ++m5; }
                }
            }
        }
    }
    if (!have) return;
    if (bad) { atomicOr(status, NUTSB_ST_BAD_OFFSETS); len_on[i] = len_off[i] = nrep[i] = 0; return; }
    if (toolong) { atomicOr(status, NUTSB_ST_TEXT_TOO_LONG); len_on[i] = len_off[i] = nrep[i] = 0; return; }
    u32 st = 0;
    const u32 loff = n - drops - 3 * (m4 + m5) + nl;
    const u32 of = ops.flags[i];
    len_off[i] = loff;
    len_on[i]  = (of & NUTSB_OF_PLAIN) ? loff                                   // colour setting ignored
               : loff + 4 * nl + 4 * m4 + 5 * m5 + ((of & NUTSB_OF_PAGER) ? 0u : 4u);   // pager lines carry no final reset

    // fan-in: how many room lists this op enters
    const u32 kind = ops.kind[i];
    const i32 tgt = ops.target[i], exc = ops.except_user[i];
    u32 rep = 0;
    if (kind == NUTSB_OP_USER) {
        if (tgt >= pop.n_users) st |= NUTSB_ST_BAD_INDEX; else if (tgt >= 0) rep = 1;
    } else if (kind == NUTSB_OP_ROOM) {
        if (tgt >= pop.n_rooms || tgt < -1) st |= NUTSB_ST_BAD_INDEX;
        else rep = tgt >= 0 ? 1u : (u32)pop.n_rooms;
    } else if (kind == NUTSB_OP_LEVEL) {
        rep = (u32)pop.n_rooms_tot; st |= NUTSB_ST_HAS_LEVEL;
    } else if (kind != NUTSB_OP_NONE) st |= NUTSB_ST_BAD_KIND;
    if (kind != NUTSB_OP_NONE && (exc >= pop.n_users || exc < -1)) st |= NUTSB_ST_BAD_INDEX;
    if (st & ~NUTSB_ST_HAS_LEVEL) rep = 0;
    nrep[i] = live ? rep : 0;
    if (st) atomicOr(status, st);
}

// ---- B. expand ops into (room, op) entries ------------------------------------------
__global__ void __launch_bounds__(256)
k_expand(OpsView ops, PopView pop, const u32 *nrep, const u64 *eoff, u32 *e_room, u32 *e_op)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ops.n) return;
    const u32 rep = nrep[i];
    if (!rep) return;
    const u64 base = eoff[i];
    const u32 kind = ops.kind[i];
    const i32 tgt = ops.target[i];
    if (rep == 1 && kind != NUTSB_OP_LEVEL) {
        u32 r = kind == NUTSB_OP_USER ? (u32)pop.user_room[tgt] : (tgt >= 0 ? (u32)tgt : 0u);
        e_room[base] = r; e_op[base] = (u32)i;
    } else {
        for (u32 j = 0; j < rep; ++j) { e_room[base + j] = j; e_op[base + j] = (u32)i; }
    }
}

// ---- stable LSD radix sort of (key,val) pairs, 11-bit digits ------------------------
#define NUTSB_RS_BITS    11
#define NUTSB_RS_DIGITS  (1 << NUTSB_RS_BITS)
#define NUTSB_RS_THREADS 256
#define NUTSB_RS_ROUNDS  16
#define NUTSB_RS_CHUNK   (NUTSB_RS_THREADS * NUTSB_RS_ROUNDS)

// hist[digit * nblocks + block]
__global__ void __launch_bounds__(NUTSB_RS_THREADS)
k_rs_hist(const u32 *keys, i64 n_host, const u32 *n_dev, int shift, u32 bits, u32 *hist, u32 nblocks)
{
    const i64 n = n_dev ? (i64)*n_dev : n_host;
    const u32 digits = 1u << bits;
    __shared__ u32 s_h[NUTSB_RS_DIGITS];
    for (u32 d = threadIdx.x; d < digits; d += blockDim.x) s_h[d] = 0;
    __syncthreads();
    const i64 base = (i64)blockIdx.x * NUTSB_RS_CHUNK;
    for (int r = 0; r < NUTSB_RS_ROUNDS; ++r) {
        const i64 i = base + r * NUTSB_RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&s_h[(keys[i] >> shift) & (digits - 1)], 1u);
    }
    __syncthreads();
    for (u32 d = threadIdx.x; d < digits; d += blockDim.x)
        hist[(size_t)d * nblocks + blockIdx.x] = s_h[d];
}

// offs = exclusive scan of hist (same layout).  vals_in == nullptr means iota.
__global__ void __launch_bounds__(NUTSB_RS_THREADS)
k_rs_scatter(const u32 *keys_in, const u32 *vals_in, i64 n_host, const u32 *n_dev, int shift, u32 bits,
             const u64 *offs, u32 nblocks, u32 *keys_out, u32 *vals_out)
{
    const i64 n = n_dev ? (i64)*n_dev : n_host;
    const u32 digits = 1u << bits;
    __shared__ u32 s_run[NUTSB_RS_DIGITS];
    for (u32 d = threadIdx.x; d < digits; d += blockDim.x)
        s_run[d] = (u32)offs[(size_t)d * nblocks + blockIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const i64 base = (i64)blockIdx.x * NUTSB_RS_CHUNK;
    for (int r = 0; r < NUTSB_RS_ROUNDS; ++r) {
        const i64 i = base + r * NUTSB_RS_THREADS + threadIdx.x;
        const bool valid = i < n;
        const u32 key = valid ? keys_in[i] : 0;
        const u32 d = valid ? ((key >> shift) & (digits - 1)) : 0xffffffffu;
        const u32 grp = __match_any_sync(NUTSB_FULL, d);
        const int leader = __ffs((int)grp) - 1;
        const u32 rank = (u32)__popc(grp & ((1u << lane) - 1));
        u32 gbase = 0;
        // warps take turns so that equal digits keep their input order
        for (int w = 0; w < NUTSB_RS_THREADS / 32; ++w) {
            if (warp == w && valid && lane == leader) { gbase = s_run[d]; s_run[d] = gbase + (u32)__popc(grp); }
            __syncthreads();
        }
        gbase = __shfl_sync(NUTSB_FULL, gbase, leader);
        if (valid) {
            keys_out[gbase + rank] = key;
            vals_out[gbase + rank] = vals_in ? vals_in[i] : (u32)i;
        }
    }
}

// seg_off[k] = first index i with keys[i] >= k, for k in [0, nkeys]; keys sorted.
__global__ void __launch_bounds__(256)
k_seg_bounds(const u32 *keys, i64 n_host, const u32 *n_dev, u32 nkeys, u32 *seg_off)
{
    const i64 n = n_dev ? (i64)*n_dev : n_host;
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const i64 lo = (i == 0) ? 0 : (i64)keys[i - 1] + 1;
    const i64 hi = (i == n) ? (i64)nkeys : (i64)keys[i];
    for (i64 k = lo; k <= hi && k <= (i64)nkeys; ++k) seg_off[k] = (u32)i;
}

// ---- C. per-entry classification -----------------------------------------------------
// e_info packs, per room-list entry: bit0 = enters the room slab (room/level op),
// bits 1-2 = event kind (0 none, 1 direct write_user op, 2 excluded recipient).
struct EntryArrays {
    const u32 *e_room, *e_op;    // sorted by room, op order inside
    u8  *e_info;
    i32 *e_delta;                // event: signed byte delta on the user's stream
    u32 *e_slot;                 // event: user slot
};

__global__ void __launch_bounds__(256)
k_entry_info(OpsView ops, PopView pop, EntryArrays ea, i64 n_ent, const u32 *len_on, const u32 *len_off)
{
    const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_ent) return;
    const u32 op = ea.e_op[e], room = ea.e_room[e];
    const u32 kind = ops.kind[op];
    u8 info = 0; i32 delta = 0; u32 slot = 0;
    if (kind == NUTSB_OP_USER) {
        const i32 u = ops.target[op];
        const u32 cf = pop.cls_flags[pop.user_cls[u]];
        info = 1u << 1;
        delta = (i32)((cf & NUTSB_UF_COLOUR) ? len_on[op] : len_off[op]);
        slot = (u32)pop.user_slot[u];
    } else {
        info = 1;
        const i32 x = ops.except_user[op];
        if (x >= 0 && (u32)pop.user_room[x] == room) {
            const i32 k = pop.user_cls[x];
            const u32 cf = pop.cls_flags[k];
            if (nutsb_class_delivers(cf, pop.cls_level[k], kind, ops.flags[op], ops.target[op])) {
                info |= 2u << 1;
                delta = -(i32)((cf & NUTSB_UF_COLOUR) ? len_on[op] : len_off[op]);
                slot = (u32)pop.user_slot[x];
            }
        }
    }
    ea.e_info[e] = info; ea.e_delta[e] = delta; ea.e_slot[e] = slot;
}

// After the packed scan over entries (lo32 = slab ops before, hi32 = events before):
// scatter the slab list and the event list.
struct EntryScatter {
    const u32 *e_room, *e_op; const u8 *e_info; const i32 *e_delta; const u32 *e_slot;
    const u32 *room_ent_off;     // [Rt+1] entry offset of each room
    u64 *e_scan;                 // [n_ent+1] packed exclusive scan (written by the scan's Out)
    u32 *bl_op, *bl_room;        // slab list
    u32 *ev_slot, *ev_ukey, *ev_op; i32 *ev_delta;
};

__global__ void __launch_bounds__(256)
k_entry_scatter(EntryScatter s, i64 n_ent)
{
    const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_ent) return;
    const u8 info = s.e_info[e];
    const u64 sc = s.e_scan[e];
    const u32 rankB = (u32)sc, evi = (u32)(sc >> 32);
    const u32 room = s.e_room[e];
    if (info & 1) { s.bl_op[rankB] = s.e_op[e]; s.bl_room[rankB] = room; }
    const u32 ek = info >> 1;
    if (ek) {
        const u32 roomB0 = (u32)s.e_scan[s.room_ent_off[room]];     // slab rank of the room's first entry
        s.ev_slot[evi]  = s.e_slot[e];
        s.ev_ukey[evi]  = 2u * (rankB - roomB0) + (ek == 2 ? 1u : 0u);
        s.ev_delta[evi] = s.e_delta[e];
        s.ev_op[evi]    = s.e_op[e];
    }
}

// room_b_off[r] = slab rank of room r's first entry, r in [0, Rt]
__global__ void __launch_bounds__(256)
k_room_b_off(const u32 *room_ent_off, const u64 *e_scan, u32 n_rooms_tot, u32 *room_b_off)
{
    const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rooms_tot) return;
    room_b_off[r] = (u32)e_scan[room_ent_off[r]];
}

// gather the sorted event arrays through the sort permutation
__global__ void __launch_bounds__(256)
k_ev_gather(const u32 *perm, const u32 *n_dev, const u32 *ev_ukey, const i32 *ev_delta, const u32 *ev_op,
            u32 *sv_ukey, i32 *sv_delta, u32 *sv_op)
{
    const i64 n_ev = (i64)*n_dev;
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ev) return;
    const u32 p = perm[i];
    sv_ukey[i] = ev_ukey[p]; sv_delta[i] = ev_delta[p]; sv_op[i] = ev_op[p];
}

// ---- class prefix access ---------------------------------------------------------------
// Bytes of the room slab a class receives before slab rank g.  In alias mode
// (every class takes every room op: no login/ignall/ignshout users, no level
// ops) this is the slab prefix itself; otherwise one scan per class column.
struct ClassPrefix {
    const u64 *vp_on, *vp_off;   // [nB+1] exclusive prefix of rendered lengths over the slab list
    const u64 *cp;               // [J][nB+1] or null (alias mode)
    u64 stride;                  // nB+1
    const i32 *room_cls_off;
    const u8  *cls_flags;
    __device__ __forceinline__ u64 at(i32 k, u32 room, u32 g) const
    {
        if (cp) return cp[(u64)(k - room_cls_off[room]) * stride + g];
        return (cls_flags[k] & NUTSB_UF_COLOUR) ? vp_on[g] : vp_off[g];
    }
};

// ---- E. per-user stream length --------------------------------------------------------
struct UserLenIn {
    ClassPrefix cpx; PopView pop;
    const u32 *room_b_off, *ev_off;     // ev_off: [U+1] by slot
    const u64 *sv_pre;                  // [n_ev+1] exclusive scan of sorted deltas (two's complement)
    __device__ u64 operator()(i64 u) const
    {
        const u32 room = (u32)pop.user_room[u];
        const i32 k = pop.user_cls[u];
        const u32 s = (u32)pop.user_slot[u];
        const u64 cls = cpx.at(k, room, room_b_off[room + 1]) - cpx.at(k, room, room_b_off[room]);
        return cls + (sv_pre[ev_off[s + 1]] - sv_pre[ev_off[s]]);
    }
};

__global__ void __launch_bounds__(128)
k_user_len(UserLenIn in, i64 n, u64 *len)
{
    const i64 u = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n) len[u] = in(u);
}

// ---- F. stream position of every (tile, recipient) cell ---------------------------------
struct Geometry {
    const u32 *room_b_off;       // [Rt+1]
    const u32 *room_tile_off;    // [Rt+1] tiles before room r
    const u64 *room_cell_off;    // [Rt+1] cells before room r; room r has (tiles_r+1)*users_r cells
    const u32 *room_item_off;    // [Rt+1] fan-out work items before room r
};

// One thread per cell: where in the recipient's stream the tile's first op
// starts (class prefix + the recipient's own exclusions / direct ops before the
// tile), and which of the recipient's events is the first one inside the tile.
__global__ void __launch_bounds__(256)
k_fill_pos(PopView pop, Geometry geo, ClassPrefix cpx, const u64 *stream_off,
           const u32 *ev_off, const u32 *sv_ukey, const u64 *sv_pre, u64 n_cells, u64 *cell_pos, u32 *cell_evi)
{
    const u64 cell = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= n_cells) return;
    u32 lo = 0, hi = (u32)pop.n_rooms_tot;               // last room with room_cell_off[r] <= cell
    while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (geo.room_cell_off[mid] <= cell) lo = mid; else hi = mid; }
    const u32 room = lo;
    const u32 users_r = (u32)(pop.room_slot_off[room + 1] - pop.room_slot_off[room]);
    const u64 local = cell - geo.room_cell_off[room];
    const u32 t = (u32)(local / users_r), ls = (u32)(local % users_r);
    const u32 s = (u32)pop.room_slot_off[room] + ls;
    const u32 b0 = geo.room_b_off[room], nb = geo.room_b_off[room + 1] - b0;
    const u32 tiles = (nb + NUTSB_TILE_OPS - 1) / NUTSB_TILE_OPS;
    const u32 e0 = ev_off[s], e1 = ev_off[s + 1];
    if (t >= tiles) { cell_pos[cell] = 0; cell_evi[cell] = e1; return; }      // sentinel row
    const u32 a0 = t * NUTSB_TILE_OPS, thr = 2 * a0 + 1;
    u32 l = e0, h = e1;                                   // first event with ukey >= thr
    while (l < h) { const u32 mid = (l + h) >> 1; if (sv_ukey[mid] < thr) l = mid + 1; else h = mid; }
    const i32 u = pop.slot_user[s];
    const i32 k = pop.user_cls[u];
    cell_pos[cell] = stream_off[u] + (cpx.at(k, room, b0 + a0) - cpx.at(k, room, b0)) + (sv_pre[l] - sv_pre[e0]);
    cell_evi[cell] = l;
}

// ---- write_user's byte machine, one thread per string ------------------------------------
// Restates nuts333.c:1315-1365.  The string sits in shared memory; words that
// hold none of '~' '/' '\n' are copied four bytes at a time, the rest goes
// through the byte-wise machine.  Emits the colour-on rendering, the colour-off
// rendering, or both in one pass (the parse is shared).
__device__ __forceinline__ u32 nutsb_special_mask(u32 w)
{
    // 0x80 in every byte of w that is '~', '/' or '\n' -- exact per byte (the cheaper
    // (y-0x01..)&~y form can flag the byte above a match, e.g. the '.' of "/.")
    const u32 M = 0x7f7f7f7fu;
    const u32 y0 = w ^ 0x7e7e7e7eu, y1 = w ^ 0x2f2f2f2fu, y2 = w ^ 0x0a0a0a0au;
    const u32 t0 = ((y0 & M) + M) | y0, t1 = ((y1 & M) + M) | y1, t2 = ((y2 & M) + M) | y2;
    return ~((t0 & t1 & t2) | M);
}

// Bytes of colcode[k] (nuts333.h:237-246), little-endian in one register pair.
__device__ __forceinline__ u64 nutsb_code_pack(int k)
{
    if (k < 5) return 0x1bull | ((u64)'[' << 8) | ((u64)('0' + ((0x75410u >> (4 * k)) & 0xf)) << 16) | ((u64)'m' << 24);
    return 0x1bull | ((u64)'[' << 8) | ((u64)(k < 13 ? '3' : '4') << 16) | ((u64)('0' + ((k - 5) & 7)) << 24) | ((u64)'m' << 32);
}
#define NUTSB_RESET_PACK 0x6d305b1bull     /* ESC [ 0 m */

// One step consumes the rest of the current aligned word up to and including its
// first special byte ('~', '/', '\n'): the plain bytes before it pass through
// (nuts333.c:1355), the special byte goes through the machine (c:1316-1354).
// What a step emits is "up to 4 plain bytes, then up to 6 bytes" chosen by the
// recipient's colour setting, stored with predicated byte stores -- one code
// path, so the strings of a warp stay converged; only a '~' that is not escaped
// takes a branch (the table lookup).  The string must sit in a window that
// starts at a 4-byte boundary (nutsb_lane_stage / nutsb_warp_stage); bytes of
// the window outside the string are never interpreted.  Returns the rendered length.
__device__ __forceinline__ u32 nutsb_render1(const u8 *s, u32 n, bool colour, u32 oflags, u8 *out, const u8 *tab)
{
    if (oflags & NUTSB_OF_PLAIN) colour = false;                 // more(NULL,...): c:2259
    u32 i = 0, o = 0;
    while (i < n) {
        const u8 *p = s + i;
        const u32 al = (u32)(size_t)p & 3u;
        const u32 w = *(const u32 *)(p - al);
        const u32 rem = n - i;
        const u32 top = al + rem < 4 ? al + rem : 4;             // bytes [al, top) of w belong to the string
        const u32 range = (0xffffffffu << (8 * al)) & (0xffffffffu >> (8 * (4 - top)));
        const u32 sm = nutsb_special_mask(w) & range;
        const u32 q = sm ? ((u32)(__ffs((int)sm) - 1) >> 3) : top;   // first special byte, or end of the word
        const u32 np = q - al;                                   // plain bytes passed through
        const u32 pv = w >> (8 * al);
        u32 adv = np, len = 0;
        u64 val = 0;
        if (sm) {
            const u32 c = (w >> (8 * q)) & 0xffu;
            const u32 j = i + np;                                // index of the special byte
            adv = np + 1; len = 1; val = c;
            if (c == '\n') {                                                   /* c:1316-1326 */
                len = colour ? 6u : 2u;
                val = colour ? (NUTSB_RESET_PACK | ((u64)'\n' << 32) | ((u64)'\r' << 40)) : (u64)((u32)'\n' | ((u32)'\r' << 8));
            } else if (c == '/') {                                             /* c:1330 */
                if (j + 1 < n && s[j + 1] == '~') len = 0;
            } else if (c == '~') {                                             /* c:1331-1354 */
                if (!(j > 0 && s[j - 1] == '/') && j + 2 < n) {
                    const int k = nutsb_code(tab, s[j + 1], s[j + 2]);
                    if (k >= 0) { adv = np + 3; len = colour ? nutsb_code_len(k) : 0u; val = nutsb_code_pack(k); }
                }
            }
        }
        u8 *d = out + o;
        if (np > 0) d[0] = (u8)pv;
        if (np > 1) d[1] = (u8)(pv >> 8);
        if (np > 2) d[2] = (u8)(pv >> 16);
        if (np > 3) d[3] = (u8)(pv >> 24);
        d += np;
        if (len > 0) d[0] = (u8)val;
        if (len > 1) d[1] = (u8)(val >> 8);
        if (len > 2) d[2] = (u8)(val >> 16);
        if (len > 3) d[3] = (u8)(val >> 24);
        if (len > 4) d[4] = (u8)(val >> 32);
        if (len > 5) d[5] = (u8)(val >> 40);
        o += np + len; i += adv;
    }
    if (colour && !(oflags & NUTSB_OF_PAGER)) {                  /* c:1365; the pager has no such reset */
        out[o] = 0x1b; out[o + 1] = '['; out[o + 2] = '0'; out[o + 3] = 'm'; o += 4;
    }
    return o;
}

// One lane stages its own string (32-bit loads; lanes of a warp walk 32 different
// strings, L1 absorbs the overlap).
__device__ __forceinline__ void nutsb_lane_stage(u8 *dst, const u8 *src, u32 n)
{
    const u32 a = (u32)((size_t)src & 3);
    const u32 *g = (const u32 *)(src - a);
    const u32 nw = (a + n + 3) >> 2;
    for (u32 w = 0; w < nw; ++w) ((u32 *)dst)[w] = __ldg(g + w);
}

// One lane copies its own rendered string from shared memory to an arbitrary
// byte address: head bytes to a 4-byte boundary, realigned 32-bit stores, tail.
__device__ __forceinline__ void nutsb_lane_copy(u8 *dst, const u8 *src, u32 n)
{
    u32 head = (u32)((4 - ((size_t)dst & 3)) & 3);
    if (head > n) head = n;
    for (u32 q = 0; q < head; ++q) dst[q] = src[q];
    dst += head; src += head; n -= head;
    const u32 nw = n >> 2;
    const u32 sm = (u32)((size_t)src & 3);
    const u32 *sa = (const u32 *)(src - sm);
    const u32 bsh = 8 * sm;
    u32 prev = nw ? sa[0] : 0;
    for (u32 j = 0; j < nw; ++j) {
        const u32 nxt = sa[j + 1];
        ((u32 *)dst)[j] = __funnelshift_r(prev, nxt, bsh);
        prev = nxt;
    }
    for (u32 q = 4 * nw; q < n; ++q) dst[q] = src[q];
}

// Staging convention: a string is copied into shared memory with 32-bit loads as the window
// of aligned words that holds it, so the window starts 4-byte aligned and the string begins
// at window + ((size_t)src & 3).  Whole words are read: up to 3 bytes either side of the
// string (inside the packed text allocation).  Size of that window:
__device__ __forceinline__ u32 nutsb_stage_bytes(const u8 *src, u32 n) { return (((u32)((size_t)src & 3)) + n + 3) & ~3u; }

// ---- warp copy: shared -> global at arbitrary byte alignment --------------------------
// The destination's first partial 16-byte chunk and its last are stored byte by
// byte (lanes 0-15 the head, lanes 16-31 the tail, one pass); the body is
// 16-byte stores, fully coalesced, two in flight per lane.  The source is
// realigned in registers from two 16-byte shared loads; the word part of the
// shift is a template parameter so no selects are executed.  src buffers carry
// >= 48 bytes of readable padding.
#ifndef NUTSB_COPY_UNROLL4
#define NUTSB_COPY_UNROLL4 0
#endif
template <int WSH>
__device__ __forceinline__ uint4 nutsb_realign(const uint4 &a, const uint4 &b, u32 bsh)
{
    const u32 W[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
    uint4 o;
    o.x = __funnelshift_r(W[WSH], W[WSH + 1], bsh);     o.y = __funnelshift_r(W[WSH + 1], W[WSH + 2], bsh);
    o.z = __funnelshift_r(W[WSH + 2], W[WSH + 3], bsh); o.w = __funnelshift_r(W[WSH + 3], W[WSH + 4], bsh);
    return o;
}
template <int WSH>
__device__ __forceinline__ void nutsb_copy_body(u8 *dst, const uint4 *sa, u32 nvec, u32 bsh, int lane)
{
    u32 v = (u32)lane;
#if NUTSB_COPY_UNROLL4
    for (; v + 96 < nvec; v += 128) {                   // four coalesced 512-byte stores in flight per lane
        const uint4 a0 = sa[v], b0 = sa[v + 1], a1 = sa[v + 32], b1 = sa[v + 33];
        const uint4 a2 = sa[v + 64], b2 = sa[v + 65], a3 = sa[v + 96], b3 = sa[v + 97];
        *(uint4 *)(dst + 16 * (size_t)v) = nutsb_realign<WSH>(a0, b0, bsh);
        *(uint4 *)(dst + 16 * (size_t)(v + 32)) = nutsb_realign<WSH>(a1, b1, bsh);
        *(uint4 *)(dst + 16 * (size_t)(v + 64)) = nutsb_realign<WSH>(a2, b2, bsh);
        *(uint4 *)(dst + 16 * (size_t)(v + 96)) = nutsb_realign<WSH>(a3, b3, bsh);
    }
#endif
    for (; v + 32 < nvec; v += 64) {
        const uint4 a0 = sa[v], b0 = sa[v + 1], a1 = sa[v + 32], b1 = sa[v + 33];
        *(uint4 *)(dst + 16 * (size_t)v) = nutsb_realign<WSH>(a0, b0, bsh);
        *(uint4 *)(dst + 16 * (size_t)(v + 32)) = nutsb_realign<WSH>(a1, b1, bsh);
    }
    if (v < nvec) *(uint4 *)(dst + 16 * (size_t)v) = nutsb_realign<WSH>(sa[v], sa[v + 1], bsh);
}

__device__ __forceinline__ void nutsb_warp_copy(u8 *dst, const u8 *src, u32 n, int lane)
{
    if (n == 0) return;
    u32 head = (u32)((16 - ((size_t)dst & 15)) & 15);
    if (head > n) head = n;
    const u32 nvec = (n - head) >> 4, tail = (n - head) & 15, toff = head + 16 * nvec;
    {   // head (lanes 0-15) and tail (lanes 16-31) bytes in one pass
        const u32 l2 = (u32)lane & 15;
        const bool is_tail = lane >= 16;
        const u32 idx = is_tail ? toff + l2 : l2;
        if (l2 < (is_tail ? tail : head)) dst[idx] = src[idx];
    }
    if (nvec == 0) return;
    dst += head; src += head;
    const u32 sm = (u32)((size_t)src & 15);
    const uint4 *sa = (const uint4 *)(src - sm);
    const u32 bsh = (sm & 3) * 8;
    switch (sm >> 2) {
    case 0:  nutsb_copy_body<0>(dst, sa, nvec, bsh, lane); break;
    case 1:  nutsb_copy_body<1>(dst, sa, nvec, bsh, lane); break;
    case 2:  nutsb_copy_body<2>(dst, sa, nvec, bsh, lane); break;
    default: nutsb_copy_body<3>(dst, sa, nvec, bsh, lane); break;
    }
}

// ---- H. render + fan-out ----------------------------------------------------------------
// One work item = (room, tile of <=64 slab ops, chunk of <=128 recipients).
// The block stages the tile's source strings in shared memory, renders each once
// (one thread per op, both colour settings in one pass: the byte machine is
// sequential, shared -> shared), prefetches the chunk's per-recipient cells, then
// every warp takes recipients in turn and copies that recipient's view of the
// tile -- normally ONE contiguous run of the rendered slab, cut only where the
// recipient is the excluded speaker or has a direct write_user op in between --
// to its place in the user's stream.
struct FanoutArgs {
    OpsView ops; PopView pop; Geometry geo; ClassPrefix cpx;
    const u32 *bl_op;
    const u32 *len_on, *len_off;
    const u64 *cell_pos; const u32 *cell_evi;
    const u32 *sv_ukey; const i32 *sv_delta;
    u8 *out;
    u64 *n_deliveries;           // [0] deliveries, [1] (k_direct), [2] source bytes staged (once per tile)
    u32 *status;
    u32 has_level;
};

#define NUTSB_FAN_THREADS (2 * NUTSB_TILE_OPS)
// -DNUTSB_FAN_PROFILE=1: thread 0 of every block adds the cycles it spent in each phase to
// counters[3..7] (setup, stage, render+plan, copy, barrier at the end of copy) -- a development aid
#ifndef NUTSB_FAN_PROFILE
#define NUTSB_FAN_PROFILE 0
#endif
#if NUTSB_FAN_PROFILE && !defined(NUTSB_CPUSIM)
#define NUTSB_PHASE(slot) do { if (tid == 0) { const long long t_ = clock64(); nutsb_add64(A.n_deliveries + (slot), (u64)(t_ - t_prev)); t_prev = t_; } } while (0)
#else
#define NUTSB_PHASE(slot) do { } while (0)
#endif
#ifndef NUTSB_FAN_TMA
#define NUTSB_FAN_TMA 0          // 1: copy runs with TMA bulk stores from 16 pre-shifted slab pieces (measured slower: DESIGN.md 4.1)
#endif
#ifndef NUTSB_FAN_PIECE
#define NUTSB_FAN_PIECE 2048     // slab bytes per TMA staging round (multiple of 512)
#endif
#ifndef NUTSB_FAN_DB
#define NUTSB_FAN_DB 1           // two staging areas: the next piece is built while the TMA reads this one
#endif
#define NUTSB_FAN_STAGE (NUTSB_FAN_TMA ? (NUTSB_FAN_DB ? 2 : 1) * 16 * NUTSB_FAN_PIECE : 0)
#define NUTSB_FAN_SMEM (NUTSB_TEXT_CAP + 32 + NUTSB_ON_CAP + 64 + NUTSB_OFF_CAP + 64 + NUTSB_FAN_STAGE)
static_assert(NUTSB_FAN_THREADS == 2 * NUTSB_TILE_OPS && (NUTSB_TILE_OPS & (NUTSB_TILE_OPS - 1)) == 0 &&
              NUTSB_UCHUNK <= NUTSB_FAN_THREADS / 2, "k_fanout thread mapping");

__global__ void __launch_bounds__(NUTSB_FAN_THREADS)
k_fanout(FanoutArgs A)
{
    NUTSB_DYN_SMEM(s_dyn);                       // NUTSB_FAN_SMEM bytes: staged source + the two rendered slabs
    u8 *const s_text = s_dyn;
    u8 *const s_on = s_dyn + NUTSB_TEXT_CAP + 32;
    u8 *const s_off = s_on + NUTSB_ON_CAP + 64;
    __shared__ u8  s_tab[NUTSB_CODETAB_BYTES];
    __shared__ u64 s_src[NUTSB_TILE_OPS];        // byte offset of each string in the packed text
    __shared__ u32 s_tlen[NUTSB_TILE_OPS];
    __shared__ u32 s_toff[NUTSB_TILE_OPS + 1];   // prefix of staged (word-aligned) sizes
    __shared__ u32 s_oon[NUTSB_TILE_OPS + 1];
    __shared__ u32 s_ooff[NUTSB_TILE_OPS + 1];
    __shared__ u8  s_kind[NUTSB_TILE_OPS];
    __shared__ u8  s_flags[NUTSB_TILE_OPS];
    __shared__ i32 s_target[NUTSB_TILE_OPS];
    __shared__ u64 s_upos[NUTSB_UCHUNK];         // per recipient of the chunk
    __shared__ u32 s_uev0[NUTSB_UCHUNK], s_uev1[NUTSB_UCHUNK], s_uevb[NUTSB_UCHUNK];
    __shared__ u8  s_ucf[NUTSB_UCHUNK], s_ulv[NUTSB_UCHUNK];
    __shared__ u32 s_evk[NUTSB_EV_CAP];          // the chunk's events inside this tile
    __shared__ i32 s_evd[NUTSB_EV_CAP];
    __shared__ u32 s_evn;
    __shared__ uint4 s_run[NUTSB_RUN_CAP];       // planned copy runs of the current (sub)tile:
                                                 // x,y = destination byte offset, z = slab offset | colour<<31, w = length
    __shared__ u8  s_ulegacy[NUTSB_UCHUNK];
    __shared__ u32 s_nruns;
    __shared__ u32 s_sub_b;
    __shared__ u32 s_room;
    __shared__ u32 s_deliv;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#if NUTSB_FAN_PROFILE && !defined(NUTSB_CPUSIM)
    long long t_prev = clock64();
#endif

    // -- decode the work item
    if (tid == 0) {
        const u32 item = blockIdx.x;
        u32 lo = 0, hi = (u32)A.pop.n_rooms_tot;        // last r with room_item_off[r] <= item
        while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (A.geo.room_item_off[mid] <= item) lo = mid; else hi = mid; }
        s_room = lo; s_deliv = 0; s_evn = 0;
    }
    for (int i = tid; i < NUTSB_CODETAB_BYTES; i += NUTSB_FAN_THREADS) s_tab[i] = A.pop.codetab[i];
    __syncthreads();
    const u32 room = s_room;
    const u32 slot0 = (u32)A.pop.room_slot_off[room];
    const u32 users_r = (u32)A.pop.room_slot_off[room + 1] - slot0;
    const u32 chunks = (users_r + NUTSB_UCHUNK - 1) / NUTSB_UCHUNK;
    const u32 local = blockIdx.x - A.geo.room_item_off[room];
    const u32 t = local / chunks, chunk = local % chunks;
    const u32 rb0 = A.geo.room_b_off[room];
    const u32 g0 = rb0 + t * NUTSB_TILE_OPS;
    const u32 gend = A.geo.room_b_off[room + 1];
    const u32 nb = (gend - g0 < NUTSB_TILE_OPS) ? gend - g0 : NUTSB_TILE_OPS;
    const u32 a0 = t * NUTSB_TILE_OPS;                  // room-local slab rank of the tile's first op
    const u32 ls_begin = chunk * NUTSB_UCHUNK;
    const u32 ls_end = (ls_begin + NUTSB_UCHUNK < users_r) ? ls_begin + NUTSB_UCHUNK : users_r;
    const u64 cell_row = A.geo.room_cell_off[room] + (u64)t * users_r;

    // -- per-op metadata (threads 0..128) and per-recipient cells + events (threads 128..255)
    if ((u32)tid < nb) {
        const u32 op = A.bl_op[g0 + tid];
        const u64 t0 = A.ops.toff[op];
        s_src[tid] = t0; s_tlen[tid] = (u32)(A.ops.toff[op + 1] - t0);
        s_kind[tid] = A.ops.kind[op]; s_flags[tid] = A.ops.flags[op]; s_target[tid] = A.ops.target[op];
    }
    if ((u32)tid <= nb) {
        s_oon[tid]  = (u32)(A.cpx.vp_on[g0 + tid]  - A.cpx.vp_on[g0]);
        s_ooff[tid] = (u32)(A.cpx.vp_off[g0 + tid] - A.cpx.vp_off[g0]);
    }
    if (tid >= NUTSB_FAN_THREADS - NUTSB_UCHUNK) {
        const u32 q = (u32)tid - (NUTSB_FAN_THREADS - NUTSB_UCHUNK);
        const u32 ls = ls_begin + q;
        if (ls < ls_end) {
            const u32 e0 = A.cell_evi[cell_row + ls], e1 = A.cell_evi[cell_row + users_r + ls];
            s_upos[q] = A.cell_pos[cell_row + ls];
            s_uev0[q] = e0; s_uev1[q] = e1;
            s_ucf[q] = A.pop.slot_cf[slot0 + ls];
            s_ulv[q] = A.pop.slot_lv[slot0 + ls];
            u32 base = 0xffffffffu;
            if (e1 > e0) {
                const u32 cnt = e1 - e0;
                const u32 got = atomicAdd(&s_evn, cnt);
                if (got + cnt <= NUTSB_EV_CAP) {
                    base = got;
                    for (u32 j = 0; j < cnt; ++j) { s_evk[base + j] = A.sv_ukey[e0 + j]; s_evd[base + j] = A.sv_delta[e0 + j]; }
                }
            }
            s_uevb[q] = base;
        }
    }
    __syncthreads();
    if (warp == 0) {                                     // prefix of the staged sizes
        u32 carry = 0;
        for (u32 base = 0; base < NUTSB_TILE_OPS; base += 32) {
            const u32 i = base + lane;
            const u32 v = i < nb ? nutsb_stage_bytes(A.ops.text + s_src[i], s_tlen[i]) : 0;
            u32 inc = v;
            for (int d = 1; d < 32; d <<= 1) { const u32 x = __shfl_up_sync(NUTSB_FULL, inc, d); if (lane >= d) inc += x; }
            if (i < nb) s_toff[i] = carry + inc - v;
            carry += __shfl_sync(NUTSB_FULL, inc, 31);
        }
        if (lane == 0) s_toff[nb] = carry;
    }
    __syncthreads();

    NUTSB_PHASE(3);
    u32 my_deliv = 0;
    u32 a = 0;
    while (a < nb) {
        // -- largest sub-tile [a,b) whose source and both renderings fit in shared memory
        if (tid == 0) {
            u32 b = nb;
            if (s_toff[nb] - s_toff[a] > NUTSB_TEXT_CAP || s_oon[nb] - s_oon[a] > NUTSB_ON_CAP ||
                s_ooff[nb] - s_ooff[a] > NUTSB_OFF_CAP) {
                b = a + 1;
                while (b < nb && s_toff[b + 1] - s_toff[a] <= NUTSB_TEXT_CAP &&
                       s_oon[b + 1] - s_oon[a] <= NUTSB_ON_CAP && s_ooff[b + 1] - s_ooff[a] <= NUTSB_OFF_CAP) ++b;
            }
            s_sub_b = b; s_nruns = 0;
        }
        __syncthreads();
        const u32 b = s_sub_b;

        // -- stage: two threads per op (tid and tid+128) copy alternate 32-bit words
        {
            const u32 i = a + ((u32)tid & (NUTSB_TILE_OPS - 1));
            if (i < b) {
                const u8 *src = A.ops.text + s_src[i];
                const u32 al = (u32)((size_t)src & 3);
                const u32 *g = (const u32 *)(src - al);
                const u32 nw = (al + s_tlen[i] + 3) >> 2;
                u32 *win = (u32 *)(s_text + (s_toff[i] - s_toff[a]));
                for (u32 w = (u32)tid / NUTSB_TILE_OPS; w < nw; w += 2) win[w] = __ldg(g + w);
            }
        }
        __syncthreads();
        NUTSB_PHASE(4);
        // -- render: threads 0..127 run the byte machine for colour-on recipients, threads
        //    128..255 for colour-off ones (the setting is uniform per warp)
        {
            const u32 i = a + ((u32)tid & (NUTSB_TILE_OPS - 1));
            const bool on = tid < NUTSB_TILE_OPS;
            if (i < b) {
                const u8 *src = A.ops.text + s_src[i];
                const u8 *str = s_text + (s_toff[i] - s_toff[a]) + ((u32)(size_t)src & 3u);
                const u32 *offc = on ? s_oon : s_ooff;
                const u32 got = nutsb_render1(str, s_tlen[i], on, s_flags[i], (on ? s_on : s_off) + (offc[i] - offc[a]), s_tab);
                if (got != offc[i + 1] - offc[i]) atomicOr(A.status, NUTSB_ST_RENDER_MISMATCH);
            }
        }

        // -- plan: one thread per recipient walks the recipient's events inside the tile and
        //    queues its copy runs (destination, slab offset, length).  Needs only the
        //    offsets, not the rendered bytes, so it runs before the barrier.
        if ((u32)tid < ls_end - ls_begin) {
            const u32 q = (u32)tid;
            const u32 cf = s_ucf[q];
            const bool full = !A.has_level && !(cf & (NUTSB_UF_LOGIN | NUTSB_UF_IGNALL | NUTSB_UF_IGNSHOUT));
            bool legacy = !full;
            if (full) {
                const bool colour = (cf & NUTSB_UF_COLOUR) != 0;
                const u32 *offc = colour ? s_oon : s_ooff;
                u64 p = s_upos[q];
                const u32 e0 = s_uev0[q], e1 = s_uev1[q], evb = s_uevb[q];
                u32 e = e0, cur = 0, deliv = 0;
                for (;;) {
                    u32 j = nb, ek = 0; i32 dlt = 0;
                    if (e < e1) {
                        const u32 uk = evb != 0xffffffffu ? s_evk[evb + (e - e0)] : A.sv_ukey[e];
                        const u32 jj = (uk >> 1) - a0;
                        if (jj < nb || (jj == nb && !(uk & 1))) {
                            j = jj; ek = (uk & 1) ? 2 : 1;
                            dlt = evb != 0xffffffffu ? s_evd[evb + (e - e0)] : A.sv_delta[e];
                        }
                    }
                    const u32 xs = cur > a ? cur : a, ye = j < b ? j : b;
                    if (xs < ye && offc[ye] != offc[xs]) {
                        const u32 r = atomicAdd(&s_nruns, 1u);
                        if (r < NUTSB_RUN_CAP) {
                            const u64 d = p + (offc[xs] - offc[cur]);
                            s_run[r] = make_uint4((u32)d, (u32)(d >> 32), (offc[xs] - offc[a]) | (colour ? 0x80000000u : 0u),
                                                  offc[ye] - offc[xs]);
                            deliv += ye - xs;
                        } else legacy = true;              // queue full: this recipient goes the slow way
                    } else if (xs < ye) deliv += ye - xs;      // zero-length renderings still count as deliveries
                    p += offc[j] - offc[cur];
                    if (!ek) break;
                    if (ek == 2) cur = j + 1; else { p += (u64)(i64)dlt; cur = j; }
                    ++e;
                }
                if (!legacy && deliv) atomicAdd(&s_deliv, deliv);
            }
            s_ulegacy[q] = legacy ? 1 : 0;
        }
        NUTSB_PHASE(5);                 // thread 0's own render + plan
        __syncthreads();
        NUTSB_PHASE(6);                 // waiting for the slowest renderer

#if NUTSB_FAN_TMA
        // -- copy with the TMA: a run's destination is byte-aligned, a bulk copy wants 16-byte
        //    aligned source, destination and size.  So: (1) every run's head and tail (< 16
        //    bytes each) are stored byte-wise by one thread per run; (2) the slab is taken a
        //    piece at a time; each piece is written 16 times into a staging area, copy r
        //    shifted left by r bytes, so whatever (source - destination) mod 16 a run has,
        //    there is a copy in which its body is 16-byte aligned; (3) one thread per run
        //    issues the body as one bulk store per piece; the SM moves no body bytes itself.
        {
            const u32 nruns = s_nruns < NUTSB_RUN_CAP ? s_nruns : NUTSB_RUN_CAP;
            u8 *const s_stg0 = s_off + NUTSB_OFF_CAP + 64;
            u32 round = 0;
            for (u32 r = (u32)tid; r < nruns; r += NUTSB_FAN_THREADS) {
                const uint4 run = s_run[r];
                const u32 so = run.z, len = run.w;
                const u8 *src = ((so >> 31) ? s_on : s_off) + (so & 0x7fffffffu);
                u8 *dst = A.out + (((u64)run.y << 32) | run.x);
                u32 head = (u32)((16 - ((size_t)dst & 15)) & 15);
                if (head > len) head = len;
                for (u32 q = 0; q < head; ++q) dst[q] = src[q];
                for (u32 q = head + ((len - head) & ~15u); q < len; ++q) dst[q] = src[q];
            }
            bool issued = false;
            for (int colour = 1; colour >= 0; --colour) {
                const u8 *slab = colour ? s_on : s_off;
                const u32 L = colour ? s_oon[b] - s_oon[a] : s_ooff[b] - s_ooff[a];
                for (u32 P0 = 0; P0 < L; P0 += NUTSB_FAN_PIECE, ++round) {
                    u8 *const s_stg = s_stg0 + (NUTSB_FAN_DB ? (round & 1) * 16 * NUTSB_FAN_PIECE : 0);
                    // (2) the 16 shifted copies of this piece: copy r holds slab[P0+r .. P0+r+PIECE)
                    for (u32 idx = (u32)tid; idx < NUTSB_FAN_PIECE; idx += NUTSB_FAN_THREADS) {
                        const u32 r = idx / (NUTSB_FAN_PIECE / 16), v = idx % (NUTSB_FAN_PIECE / 16);   // r is uniform per warp
                        if (P0 + r + 16 * v >= L) continue;
                        const uint4 *sa = (const uint4 *)(slab + P0) + v;
                        const uint4 x = sa[0], y = sa[1];
                        const u32 bsh = (r & 3) * 8;
                        uint4 o;
                        switch (r >> 2) {
                        case 0:  o = nutsb_realign<0>(x, y, bsh); break;
                        case 1:  o = nutsb_realign<1>(x, y, bsh); break;
                        case 2:  o = nutsb_realign<2>(x, y, bsh); break;
                        default: o = nutsb_realign<3>(x, y, bsh); break;
                        }
                        ((uint4 *)(s_stg + r * NUTSB_FAN_PIECE))[v] = o;
                    }
                    nutsb_fence_async_smem();
                    __syncthreads();
                    // (3) bodies: one bulk store per (run, piece)
                    for (u32 q = (u32)tid; q < nruns; q += NUTSB_FAN_THREADS) {
                        const uint4 run = s_run[q];
                        const u32 so = run.z;
                        if ((int)(so >> 31) != colour) continue;
                        const u32 S = so & 0x7fffffffu, len = run.w;
                        const u64 D = ((u64)run.y << 32) | run.x;
                        u32 head = (u32)((16 - (((size_t)A.out + D) & 15)) & 15);
                        if (head > len) head = len;
                        const u32 nvec = (len - head) >> 4;
                        if (!nvec) continue;
                        const u32 Sp = S + head, r = Sp & 15;             // slab bases are 16-byte aligned
                        const u32 lo = P0 + r, hi = lo + NUTSB_FAN_PIECE;
                        const u32 x0 = Sp > lo ? Sp : lo, x1 = Sp + 16 * nvec < hi ? Sp + 16 * nvec : hi;
                        if (x1 > x0) {
                            nutsb_bulk_s2g(A.out + D + head + (x0 - Sp), s_stg + r * NUTSB_FAN_PIECE + (x0 - lo), x1 - x0);
                            issued = true;
                        }
                    }
                    nutsb_bulk_commit();
                    if (NUTSB_FAN_DB) nutsb_bulk_wait_read1();   // the other staging area is free again
                    else nutsb_bulk_wait_read();                  // the staging area is rewritten by the next round
                    __syncthreads();
                }
            }
            nutsb_bulk_wait_read();                  // slabs and staging are reused by the next sub-tile
            if (issued) nutsb_bulk_wait_all();
            __syncthreads();
        }
#else
        // -- copy: a run is one contiguous piece of a recipient's stream; warps take runs in turn
        {
            const u32 nruns = s_nruns < NUTSB_RUN_CAP ? s_nruns : NUTSB_RUN_CAP;
            for (u32 r = (u32)warp; r < nruns; r += NUTSB_FAN_THREADS / 32) {
                const uint4 run = s_run[r];
                // the colour-off slab follows the colour-on one in the dynamic window
                const u32 so = (run.z & 0x7fffffffu) + ((run.z >> 31) ? 0u : (u32)(NUTSB_ON_CAP + 64));
                nutsb_warp_copy(A.out + (((u64)run.y << 32) | run.x), s_on + so, run.w, lane);
            }
        }
#endif
        // -- recipients that are not plain listeners (login / ignall / ignshout, level ops in the
        //    batch) or did not fit the queue: one warp per recipient, op by op where needed
        for (u32 ls = ls_begin + warp; ls < ls_end; ls += NUTSB_FAN_THREADS / 32) {
            const u32 q = ls - ls_begin;
            if (!s_ulegacy[q]) continue;
            const u32 cf = s_ucf[q], clv = s_ulv[q];
            const bool colour = (cf & NUTSB_UF_COLOUR) != 0;
            const bool full = !A.has_level && !(cf & (NUTSB_UF_LOGIN | NUTSB_UF_IGNALL | NUTSB_UF_IGNSHOUT));
            const u32 *offc = colour ? s_oon : s_ooff;
            const u8 *slab = colour ? s_on : s_off;
            u64 p = s_upos[q];
            const u32 e0 = s_uev0[q], e1 = s_uev1[q], evb = s_uevb[q];
            u32 e = e0;
            u32 cur = 0;
            for (;;) {
                u32 j = nb; u32 ek = 0; i32 dlt = 0;
                if (e < e1) {
                    const u32 uk = evb != 0xffffffffu ? s_evk[evb + (e - e0)] : A.sv_ukey[e];
                    const u32 jj = (uk >> 1) - a0;
                    if (jj < nb || (jj == nb && !(uk & 1))) {
                        j = jj; ek = (uk & 1) ? 2 : 1;
                        dlt = evb != 0xffffffffu ? s_evd[evb + (e - e0)] : A.sv_delta[e];
                    }
                }
                if (full) {
                    const u32 xs = cur > a ? cur : a, ye = j < b ? j : b;
                    if (xs < ye) {
                        nutsb_warp_copy(A.out + p + (offc[xs] - offc[cur]), slab + (offc[xs] - offc[a]),
                                        offc[ye] - offc[xs], lane);
                        my_deliv += ye - xs;
                    }
                    p += offc[j] - offc[cur];
                } else {
                    for (u32 i = cur; i < j; ++i) {
                        if (!nutsb_class_delivers(cf, clv, s_kind[i], s_flags[i], s_target[i])) continue;
                        const u32 len = offc[i + 1] - offc[i];
                        if (i >= a && i < b) { nutsb_warp_copy(A.out + p, slab + (offc[i] - offc[a]), len, lane); ++my_deliv; }
                        p += len;
                    }
                }
                if (!ek) break;
                if (ek == 2) cur = j + 1;                 // excluded from op j: nothing emitted for it
                else { p += (u64)(i64)dlt; cur = j; }      // a direct op's bytes go here (k_direct writes them)
                ++e;
            }
        }
        NUTSB_PHASE(7);                 // warp 0's share of the copy
        __syncthreads();
        NUTSB_PHASE(8);                 // waiting for the slowest copier
        a = b;
    }
    if (lane == 0 && my_deliv) atomicAdd(&s_deliv, my_deliv);
    __syncthreads();
    if (tid == 0) {
        if (s_deliv) nutsb_add64(A.n_deliveries, (u64)s_deliv);
        if (chunk == 0) {                                  // each slab op's source is read once per tile
            u32 tb = 0; for (u32 i = 0; i < nb; ++i) tb += s_tlen[i];
            nutsb_add64(A.n_deliveries + 2, (u64)tb);
        }
    }
}

// ---- I. direct ops (write_user) -----------------------------------------------------------
// One thread per event (events are sorted by recipient).  The direct ops among
// a block's 256 events are staged, rendered in the recipient's colour setting and
// copied into the recipient's stream at the offset the event prefix gives -- all
// three by the op's own thread, in sub-batches sized to shared memory.
struct DirectArgs {
    OpsView ops; PopView pop; ClassPrefix cpx;
    const u32 *room_b_off, *ev_off;
    const u32 *sv_ukey, *sv_op; const u64 *sv_pre;
    const u64 *stream_off;
    const u32 *ev_slot_sorted;      // user slot of each sorted event
    const i32 *sv_delta;
    u8 *out; i64 n_ev;
    u64 *n_deliveries;              // [0] deliveries, [1] bytes written by k_direct, [2] (k_fanout)
    u32 *status;
};

#define NUTSB_DIRECT_THREADS 256
#define NUTSB_DIR_WTEXT 2560        // per-warp staging bytes  (>= 2000 + 12)
#define NUTSB_DIR_WOUT  3072        // per-warp rendered bytes

// Warps are independent (no block barrier in the loop): a warp takes 32 consecutive
// events, the direct ops among them are staged, rendered and copied out by their own
// lanes through the warp's private shared-memory windows, in sub-batches that fit the
// windows.  A rendering too large for the window (a 2000-byte string of newlines is
// 12 KB) is written to the stream directly by its lane.
__global__ void __launch_bounds__(NUTSB_DIRECT_THREADS)
k_direct(DirectArgs A)
{
    __shared__ __align__(16) u8 s_text[NUTSB_DIRECT_THREADS / 32][NUTSB_DIR_WTEXT + 32];
    __shared__ __align__(16) u8 s_out[NUTSB_DIRECT_THREADS / 32][NUTSB_DIR_WOUT + 64];
    __shared__ u8  s_tab[NUTSB_CODETAB_BYTES];
    __shared__ u32 s_cnt;
    __shared__ unsigned long long s_bytes;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < NUTSB_CODETAB_BYTES; i += NUTSB_DIRECT_THREADS) s_tab[i] = A.pop.codetab[i];
    if (tid == 0) { s_cnt = 0; s_bytes = 0; }
    __syncthreads();

    const i64 e = (i64)blockIdx.x * NUTSB_DIRECT_THREADS + tid;
    bool isw = false, colour = false;
    u64 p = 0; const u8 *src = A.ops.text; u32 n = 0, tsz = 0, osz = 0, oflags = 0;
    if (e < A.n_ev) {
        const u32 uk = A.sv_ukey[e];
        if (!(uk & 1)) {                                   // an odd key is an exclusion
            const u32 s = A.ev_slot_sorted[e];
            const i32 u = A.pop.slot_user[s];
            const u32 room = (u32)A.pop.user_room[u];
            const i32 k = A.pop.user_cls[u];
            const u32 b0 = A.room_b_off[room];
            const u32 op = A.sv_op[e];
            const u64 t0 = A.ops.toff[op];
            p = A.stream_off[u] + (A.cpx.at(k, room, b0 + (uk >> 1)) - A.cpx.at(k, room, b0))
              + (A.sv_pre[e] - A.sv_pre[A.ev_off[s]]);
            src = A.ops.text + t0;
            n = (u32)(A.ops.toff[op + 1] - t0);
            colour = (A.pop.slot_cf[s] & NUTSB_UF_COLOUR) != 0;
            oflags = A.ops.flags[op];
            isw = true;
            tsz = nutsb_stage_bytes(src, n);
            osz = (u32)A.sv_delta[e];
        }
    }
    // inclusive prefixes over the warp's lanes
    u32 it = tsz, io = osz;
    for (int d = 1; d < 32; d <<= 1) {
        const u32 xt = __shfl_up_sync(NUTSB_FULL, it, d), xo = __shfl_up_sync(NUTSB_FULL, io, d);
        if (lane >= d) { it += xt; io += xo; }
    }
    u8 *const wtext = s_text[warp];
    u8 *const wout = s_out[warp];
    u32 a = 0;
    while (a < 32) {
        const u32 bt = __shfl_sync(NUTSB_FULL, it - tsz, (int)a), bo = __shfl_sync(NUTSB_FULL, io - osz, (int)a);   // exclusive at lane a
        const bool fits = (u32)lane >= a && it - bt <= NUTSB_DIR_WTEXT && io - bo <= NUTSB_DIR_WOUT;
        const u32 nofit = __ballot_sync(NUTSB_FULL, (u32)lane >= a && !fits);
        u32 b = nofit ? (u32)__ffs((int)nofit) - 1 : 32u;          // [a,b) fits the windows
        if (b == a) {
            // lane a's rendering alone exceeds the window: stage it, render straight into the stream
            if ((u32)lane == a && isw) {
                nutsb_lane_stage(wtext, src, n);
                if (nutsb_render1(wtext + ((u32)(size_t)src & 3u), n, colour, oflags, A.out + p, s_tab) != osz)
                    atomicOr(A.status, NUTSB_ST_RENDER_MISMATCH);
            }
            b = a + 1;
        } else if (isw && (u32)lane >= a && (u32)lane < b) {
            u8 *win = wtext + (it - tsz - bt);
            u8 *dst = wout + (io - osz - bo);
            nutsb_lane_stage(win, src, n);
            if (nutsb_render1(win + ((u32)(size_t)src & 3u), n, colour, oflags, dst, s_tab) != osz)
                atomicOr(A.status, NUTSB_ST_RENDER_MISMATCH);
            nutsb_lane_copy(A.out + p, dst, osz);
        }
        __syncwarp();                                       // the windows are reused by the next sub-batch
        a = b;
    }
    const u32 wcnt = (u32)__popc(__ballot_sync(NUTSB_FULL, isw));
    u32 wbytes = isw ? osz : 0;
    for (int d = 16; d; d >>= 1) wbytes += __shfl_xor_sync(NUTSB_FULL, wbytes, d);
    if (lane == 0 && wcnt) { atomicAdd(&s_cnt, wcnt); atomicAdd(&s_bytes, (unsigned long long)wbytes); }
    __syncthreads();
    if (tid == 0 && s_cnt) { nutsb_add64(A.n_deliveries, (u64)s_cnt); nutsb_add64(A.n_deliveries + 1, (u64)s_bytes); }
}

// ---- stream digests ------------------------------------------------------------------------
// h = fold(h * P + byte), h0 = FNV offset basis.  A block per user: each thread
// folds a contiguous slice as an affine map (mult, add), the block composes them
// in order.
#define NUTSB_DIGEST_P  0x100000001b3ull
#define NUTSB_DIGEST_H0 0xcbf29ce484222325ull

__global__ void __launch_bounds__(256)
k_digest(const u8 *bytes, const u64 *off, i32 n_users, u64 *digest)
{
    __shared__ u64 s_m[256], s_a[256];
    for (i32 u = blockIdx.x; u < n_users; u += gridDim.x) {
        const u64 b = off[u], n = off[u + 1] - b;
        const u64 per = (n + 255) / 256;
        u64 lo = (u64)threadIdx.x * per, hi = lo + per;
        if (lo > n) lo = n;
        if (hi > n) hi = n;
        u64 m = 1, a = 0;                               // x -> x*m + a
        for (u64 i = lo; i < hi; ++i) { m *= NUTSB_DIGEST_P; a = a * NUTSB_DIGEST_P + bytes[b + i]; }
        s_m[threadIdx.x] = m; s_a[threadIdx.x] = a;
        __syncthreads();
        if (threadIdx.x == 0) {
            u64 h = NUTSB_DIGEST_H0;
            for (int q = 0; q < 256; ++q) h = h * s_m[q] + s_a[q];
            digest[u] = h;
        }
        __syncthreads();
    }
}

// ---- small bookkeeping kernels -----------------------------------------------------------
// counts[0] = slab ops, counts[1] = events (from the packed entry scan's total)
__global__ void k_counts(const u64 *e_scan, i64 n_ent, u32 *counts)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) { const u64 t = e_scan[n_ent]; counts[0] = (u32)t; counts[1] = (u32)(t >> 32); }
}

struct Sizes {                   // read back by the host before the fan-out is launched
    u64 total_bytes;
    u64 cells;
    u32 n_slab, n_events, items, tiles;
};

// Per room: tiles, (tile, recipient) cells and fan-out work items; exclusive
// prefixes over rooms.  One block.
__global__ void __launch_bounds__(NUTSB_SCAN_THREADS)
k_geometry(PopView pop, const u32 *room_b_off, const u64 *stream_off, const u32 *counts,
           u32 *room_tile_off, u64 *room_cell_off, u32 *room_item_off, Sizes *sz)
{
    u64 c_tiles = 0, c_cells = 0, c_items = 0;
    const i32 rt = pop.n_rooms_tot;
    for (i32 base = 0; base < rt; base += NUTSB_SCAN_THREADS) {
        const i32 r = base + (i32)threadIdx.x;
        u64 tiles = 0, cells = 0, items = 0;
        if (r < rt) {
            const u64 users = (u64)(pop.room_slot_off[r + 1] - pop.room_slot_off[r]);
            const u64 nb = room_b_off[r + 1] - room_b_off[r];
            tiles = (nb + NUTSB_TILE_OPS - 1) / NUTSB_TILE_OPS;
            cells = (tiles + 1) * users;
            items = tiles * ((users + NUTSB_UCHUNK - 1) / NUTSB_UCHUNK);
        }
        u64 t1, t2, t3;
        const u64 e1 = nutsb_block_excl_scan(tiles, &t1);
        const u64 e2 = nutsb_block_excl_scan(cells, &t2);
        const u64 e3 = nutsb_block_excl_scan(items, &t3);
        if (r < rt) {
            room_tile_off[r] = (u32)(c_tiles + e1);
            room_cell_off[r] = c_cells + e2;
            room_item_off[r] = (u32)(c_items + e3);
        }
        c_tiles += t1; c_cells += t2; c_items += t3;
    }
    if (threadIdx.x == 0) {
        room_tile_off[rt] = (u32)c_tiles; room_cell_off[rt] = c_cells; room_item_off[rt] = (u32)c_items;
        sz->total_bytes = stream_off[pop.n_users];
        sz->cells = c_cells; sz->n_slab = counts[0]; sz->n_events = counts[1];
        sz->items = (u32)c_items; sz->tiles = (u32)c_tiles;
    }
}
