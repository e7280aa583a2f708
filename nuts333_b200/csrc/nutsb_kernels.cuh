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
#define NUTSB_TILE_OPS   128     // slab ops per fan-out tile
#endif
#define NUTSB_UCHUNK     128     // recipients per fan-out work item
#ifndef NUTSB_FAN_ON_CAP
#define NUTSB_FAN_ON_CAP  (96 * NUTSB_TILE_OPS)   // shared-memory window for a tile's colour-on rendering
#endif
#ifndef NUTSB_FAN_OFF_CAP
#define NUTSB_FAN_OFF_CAP (80 * NUTSB_TILE_OPS)   // ... and its colour-off rendering (larger tiles are copied slab -> stream directly)
#endif
#define NUTSB_FAN_RUN_CAP 256    // runs staged per round

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
    u32 *bl_meta;                // kind | flags << 8 | clamped target << 16 of each slab op (k_plan's filter walk)
    OpsView ops;
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
    if (info & 1) {
        const u32 op = s.e_op[e];
        i32 tg = s.ops.target[op]; tg = tg < -32768 ? -32768 : (tg > 32767 ? 32767 : tg);   // only compared with a level (u8)
        s.bl_op[rankB] = op; s.bl_room[rankB] = room;
        s.bl_meta[rankB] = (u32)s.ops.kind[op] | ((u32)s.ops.flags[op] << 8) | ((u32)(tg & 0xffff) << 16);
    }
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

// ---- F. copy plan ---------------------------------------------------------------------------
// A room's slab ops are cut into tiles of NUTSB_TILE_OPS; a cell is one (tile,
// recipient) pair.  Every recipient's view of a tile is a short list of RUNS:
// contiguous pieces of the rendered slab (k_render), cut only where the
// recipient is the excluded speaker (c:1415), where a direct write_user op's
// bytes interleave, or -- for recipients behind a filter (login / ignall /
// ignshout, write_level) -- where an op is not delivered to the recipient's
// class.  The plan is made in two passes over the cells (count, exclusive scan,
// fill) and leaves a flat run list in (room, tile, recipient) order plus one
// descriptor per fan-out work item, so that the kernel that moves the bytes
// (k_fanout) has nothing to decide.
struct Geometry {
    const u32 *room_b_off;       // [Rt+1] slab rank of room r's first op
    const u32 *room_tile_off;    // [Rt+1] tiles before room r
    const u64 *room_cell_off;    // [Rt+1] cells before room r; room r has tiles_r * users_r cells
    const u32 *room_item_off;    // [Rt+1] fan-out work items before room r
};

struct ItemDesc {                // one fan-out work item = (room, tile, chunk of <= NUTSB_UCHUNK recipients)
    u64 on_src, off_src;         // the tile's two renderings: byte offsets into the slab buffer
    u32 on_len, off_len;
    u32 run_begin, run_cnt;      // the item's runs in the run list
};

// A run, 16 bytes: x,y = destination byte offset in the stream buffer, z = source
// byte offset in the slab buffer (low 32 bits), w = length (24 bits) | source bits 32..39 << 24.
__device__ __forceinline__ uint4 nutsb_run_pack(u64 dst, u64 src, u32 len)
{
    return make_uint4((u32)dst, (u32)(dst >> 32), (u32)src, (len & 0xffffffu) | ((u32)(src >> 32) << 24));
}
#define NUTSB_RUN_MAX_LEN 0xffffffu

struct PlanArgs {
    PopView pop; Geometry geo; ClassPrefix cpx;
    const u64 *stream_off;
    const u32 *ev_off, *sv_ukey; const i32 *sv_delta; const u64 *sv_pre;
    const u32 *bl_meta;          // per slab op: kind | flags << 8 | clamped target << 16
    u64 n_cells, off_base;       // off_base: where the colour-off renderings start in the slab buffer
    u32 has_level;
    u32 *cell_nruns;             // count pass: out
    const u64 *run_off;          // fill pass: exclusive scan of cell_nruns, [n_cells+1]
    uint4 *runs; ItemDesc *items;
    u64 *counters;               // [0] deliveries
    u32 *status;
};

template <bool FILL>
__global__ void __launch_bounds__(256)
k_plan(PlanArgs A)
{
    __shared__ u32 s_deliv;
    if (FILL) { if (threadIdx.x == 0) s_deliv = 0; __syncthreads(); }
    const u64 cell = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u32 deliv = 0;
    if (cell < A.n_cells) {
        u32 lo = 0, hi = (u32)A.pop.n_rooms_tot;            // last room with room_cell_off[r] <= cell
        while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (A.geo.room_cell_off[mid] <= cell) lo = mid; else hi = mid; }
        const u32 room = lo;
        const u32 slot0 = (u32)A.pop.room_slot_off[room];
        const u32 users_r = (u32)A.pop.room_slot_off[room + 1] - slot0;
        const u64 local = cell - A.geo.room_cell_off[room];
        const u32 t = (u32)(local / users_r), ls = (u32)(local % users_r);
        const u32 s = slot0 + ls;
        const u32 b0 = A.geo.room_b_off[room], nb_room = A.geo.room_b_off[room + 1] - b0;
        const u32 a0 = t * NUTSB_TILE_OPS;                   // room-local slab rank of the tile's first op
        const u32 nb = nb_room - a0 < NUTSB_TILE_OPS ? nb_room - a0 : NUTSB_TILE_OPS;
        const u32 g0 = b0 + a0;
        const u32 e0 = A.ev_off[s], e1 = A.ev_off[s + 1];
        // the recipient's events inside the tile: keys in [2*a0+1, 2*(a0+nb)]
        u32 l = e0, h = e1;
        { const u32 thr = 2 * a0 + 1; while (l < h) { const u32 mid = (l + h) >> 1; if (A.sv_ukey[mid] < thr) l = mid + 1; else h = mid; } }
        u32 l_end = l; h = e1;
        { const u32 thr = 2 * (a0 + nb) + 1; while (l_end < h) { const u32 mid = (l_end + h) >> 1; if (A.sv_ukey[mid] < thr) l_end = mid + 1; else h = mid; } }
        const i32 u = A.pop.slot_user[s];
        const i32 k = A.pop.user_cls[u];
        const u32 cf = A.pop.slot_cf[s], clv = A.pop.slot_lv[s];
        const bool colour = (cf & NUTSB_UF_COLOUR) != 0;
        const bool full = !A.has_level && !(cf & (NUTSB_UF_LOGIN | NUTSB_UF_IGNALL | NUTSB_UF_IGNSHOUT));
        const u64 *vp = (colour ? A.cpx.vp_on : A.cpx.vp_off) + g0;      // tile-local prefix of rendered lengths
        const u64 sb = colour ? 0 : A.off_base;
        // stream position of the tile's first op: class prefix + the recipient's own events before the tile
        u64 p = 0;
        if (FILL) p = A.stream_off[u] + (A.cpx.at(k, room, g0) - A.cpx.at(k, room, b0)) + (A.sv_pre[l] - A.sv_pre[e0]);
        u64 r_out = FILL ? A.run_off[cell] : 0;
        u32 nruns = 0;
        u32 cur = 0;
        for (u32 e = l; ; ++e) {
            u32 j = nb; bool excl = false; i32 dlt = 0;
            if (e < l_end) { const u32 uk = A.sv_ukey[e]; j = (uk >> 1) - a0; excl = (uk & 1) != 0; if (FILL && !excl) dlt = A.sv_delta[e]; }
            if (full) {
                if (cur < j) {
                    const u64 v0 = vp[cur], v1 = vp[j];
                    deliv += j - cur;                                  // zero-length renderings are deliveries too
                    if (v1 > v0) {
                        if (FILL) { A.runs[r_out + nruns] = nutsb_run_pack(p, sb + v0, (u32)(v1 - v0)); p += v1 - v0; }
                        ++nruns;
                    }
                }
            } else {
                // behind a filter: op by op, maximal stretches of delivered ops make a run
                u64 run_dst = 0, run_src = 0; u32 run_len = 0;
                for (u32 i = cur; i < j; ++i) {
                    const u32 m = A.bl_meta[g0 + i];
                    const bool del = nutsb_class_delivers(cf, clv, m & 0xffu, (m >> 8) & 0xffu, (i32)(int16_t)(m >> 16));
                    if (del) {
                        const u64 v0 = vp[i]; const u32 len = (u32)(vp[i + 1] - v0);
                        ++deliv;
                        if (!run_len) { run_dst = p; run_src = sb + v0; }
                        run_len += len; p += len;
                    }
                    if ((!del || i + 1 == j) && run_len) {
                        if (FILL) A.runs[r_out + nruns] = nutsb_run_pack(run_dst, run_src, run_len);
                        ++nruns; run_len = 0;
                    }
                }
            }
            if (e >= l_end) break;
            if (excl) cur = j + 1;                            // excluded from op j: nothing emitted for it
            else { p += (u64)(i64)dlt; cur = j; }             // a direct op's bytes go here (k_direct writes them)
        }
        if (!FILL) A.cell_nruns[cell] = nruns;
        else if (ls % NUTSB_UCHUNK == 0) {
            const u32 chunks = (users_r + NUTSB_UCHUNK - 1) / NUTSB_UCHUNK;
            const u32 ls_end = ls + NUTSB_UCHUNK < users_r ? ls + NUTSB_UCHUNK : users_r;
            const u64 r_end = A.run_off[cell + (ls_end - ls)];
            ItemDesc d;
            d.on_src = A.cpx.vp_on[g0]; d.on_len = (u32)(A.cpx.vp_on[g0 + nb] - d.on_src);
            const u64 o0 = A.cpx.vp_off[g0];
            d.off_src = A.off_base + o0; d.off_len = (u32)(A.cpx.vp_off[g0 + nb] - o0);
            d.run_begin = (u32)r_out; d.run_cnt = (u32)(r_end - r_out);
            A.items[A.geo.room_item_off[room] + t * chunks + ls / NUTSB_UCHUNK] = d;
        }
    }
    if (FILL) {
        for (int d = 16; d; d >>= 1) deliv += __shfl_xor_sync(NUTSB_FULL, deliv, d);
        if ((threadIdx.x & 31) == 0 && deliv) atomicAdd(&s_deliv, deliv);
        __syncthreads();
        if (threadIdx.x == 0 && s_deliv) nutsb_add64(A.counters, (u64)s_deliv);
    }
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
template <int WSH>
__device__ __forceinline__ uint4 nutsb_realign(const uint4 &a, const uint4 &b, u32 bsh)
{
    const u32 W[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
    uint4 o;
    o.x = __funnelshift_r(W[WSH], W[WSH + 1], bsh);     o.y = __funnelshift_r(W[WSH + 1], W[WSH + 2], bsh);
    o.z = __funnelshift_r(W[WSH + 2], W[WSH + 3], bsh); o.w = __funnelshift_r(W[WSH + 3], W[WSH + 4], bsh);
    return o;
}
template <int WSH, int WIDTH>
__device__ __forceinline__ void nutsb_copy_body(u8 *dst, const uint4 *sa, u32 nvec, u32 bsh, int idx)
{
    u32 v = (u32)idx;
    for (; v + WIDTH < nvec; v += 2 * WIDTH) {          // two coalesced stores in flight per lane
        const uint4 a0 = sa[v], b0 = sa[v + 1], a1 = sa[v + WIDTH], b1 = sa[v + WIDTH + 1];
        *(uint4 *)(dst + 16 * (size_t)v) = nutsb_realign<WSH>(a0, b0, bsh);
        *(uint4 *)(dst + 16 * (size_t)(v + WIDTH)) = nutsb_realign<WSH>(a1, b1, bsh);
    }
    if (v < nvec) *(uint4 *)(dst + 16 * (size_t)v) = nutsb_realign<WSH>(sa[v], sa[v + 1], bsh);
}

// WIDTH threads (a warp, or several warps of one block) copy one run; idx = the thread's
// index in the group.
template <int WIDTH>
__device__ __forceinline__ void nutsb_group_copy(u8 *dst, const u8 *src, u32 n, int idx)
{
    if (n == 0) return;
    u32 head = (u32)((16 - ((size_t)dst & 15)) & 15);
    if (head > n) head = n;
    const u32 nvec = (n - head) >> 4, tail = (n - head) & 15, toff = head + 16 * nvec;
    if (idx < 32) {   // head (threads 0-15) and tail (threads 16-31) bytes in one pass
        const u32 l2 = (u32)idx & 15;
        const bool is_tail = idx >= 16;
        const u32 at = is_tail ? toff + l2 : l2;
        if (l2 < (is_tail ? tail : head)) dst[at] = src[at];
    }
    if (nvec == 0) return;
    dst += head; src += head;
    const u32 sm = (u32)((size_t)src & 15);
    const uint4 *sa = (const uint4 *)(src - sm);
    const u32 bsh = (sm & 3) * 8;
    switch (sm >> 2) {
    case 0:  nutsb_copy_body<0, WIDTH>(dst, sa, nvec, bsh, idx); break;
    case 1:  nutsb_copy_body<1, WIDTH>(dst, sa, nvec, bsh, idx); break;
    case 2:  nutsb_copy_body<2, WIDTH>(dst, sa, nvec, bsh, idx); break;
    default: nutsb_copy_body<3, WIDTH>(dst, sa, nvec, bsh, idx); break;
    }
}
__device__ __forceinline__ void nutsb_warp_copy(u8 *dst, const u8 *src, u32 n, int lane) { nutsb_group_copy<32>(dst, src, n, lane); }

// ---- G. render the slab ---------------------------------------------------------------------
// Every slab op is rendered ONCE per colour setting (the reference renders it once
// per recipient, c:1427 -> c:1315) into the slab buffer: the colour-on rendering of
// slab rank g at slab + vp_on[g], the colour-off one at slab + off_base + vp_off[g].
//
// write_user's byte machine (c:1315-1365) is position-local: what
// byte i emits depends on bytes i-3..i+2 only.  So a warp takes NUTSB_REN_OPS
// consecutive slab ops as ONE flat array of aligned 32-bit words and renders 32
// words per round, lane = word, whatever the string lengths: every lane classifies
// its four bytes, a warp scan of the emitted lengths places them, and the bytes go
// into the warp's private shared-memory windows, flushed to the slab with 16-byte
// stores whenever a window could overflow in the next round.  No text staging: a
// lane reads its word and the two neighbouring words straight from the packed text
// (consecutive lanes read consecutive words of the same string).
#define NUTSB_REN_THREADS 256
#define NUTSB_REN_OPS     16                       // slab ops per warp (<= 32)
#define NUTSB_REN_ON_WIN  3072                     // per-warp window, colour on  (a round emits <= 32*28 bytes)
#define NUTSB_REN_OFF_WIN 1536                     // per-warp window, colour off (a round emits <= 32*8 bytes)
#define NUTSB_REN_ON_ROUND  (32 * 28)
#define NUTSB_REN_OFF_ROUND (32 * 8)

struct RenderArgs {
    OpsView ops; const u8 *codetab;
    const u32 *bl_op; const u64 *vp_on, *vp_off;
    const u32 *n_slab;
    u8 *slab; u64 off_base;
    u64 *counters;               // [2] source bytes read
    u32 *status;
};

// 0xff in byte k of the result iff lo <= wb + k < hi (window byte coordinates)
__device__ __forceinline__ u32 nutsb_keep_mask(i32 wb, i32 lo, i32 hi)
{
    i32 dl = lo - wb; dl = dl < 0 ? 0 : dl;
    i32 dh = wb + 4 - hi; dh = dh < 0 ? 0 : dh;
    const u32 a = dl >= 4 ? 0u : 0xffffffffu << (8 * dl);
    const u32 b = dh >= 4 ? 0u : 0xffffffffu >> (8 * dh);
    return a & b;
}

// Context of one lane: pv | x | nx = window words wi-1, wi, wi+1 with every byte outside
// the string zeroed.  NUTSB_CTX(j), j in [-3, 5], is the byte j positions from x's byte 0.
#define NUTSB_CTX(j) ((j) < 0 ? (pv >> (8 * (((j) + 4) & 3))) & 0xffu : (j) < 4 ? (x >> (8 * ((j) & 3))) & 0xffu : (nx >> (8 * (((j) - 4) & 3))) & 0xffu)

// What position k of the lane's word emits (SURVEY.md A.1's table).  on_len/off_len in
// bytes; von = the colour-on bytes, little-endian; the colour-off bytes are "\n\r"
// (off_len 2) or the byte itself (off_len 1).
#define NUTSB_CLASSIFY(k, valid, colour, on_len, off_len, von)                                               \
    do {                                                                                                     \
        const u32 c_ = NUTSB_CTX(k);                                                                         \
        on_len = 0; off_len = 0; von = c_;                                                                   \
        if (valid) {                                                                                         \
            on_len = 1; off_len = 1;                                                                         \
            if (c_ == '\n') {                                                   /* c:1316-1326 */            \
                off_len = 2;                                                                                 \
                if (colour) { on_len = 6; von = NUTSB_RESET_PACK | ((u64)'\n' << 32) | ((u64)'\r' << 40); }   \
                else { on_len = 2; von = (u64)((u32)'\n' | ((u32)'\r' << 8)); }                               \
            } else if (c_ == '/') {                                             /* c:1330 */                 \
                if (NUTSB_CTX((k) + 1) == '~') { on_len = 0; off_len = 0; }                                  \
            } else if (c_ == '~') {                                             /* c:1331-1354 */            \
                if (NUTSB_CTX((k) - 1) != '/') {                                                             \
                    const int kk_ = nutsb_code(tab, (u8)NUTSB_CTX((k) + 1), (u8)NUTSB_CTX((k) + 2));         \
                    if (kk_ >= 0) { off_len = 0; on_len = colour ? nutsb_code_len(kk_) : 0u; von = nutsb_code_pack(kk_); } \
                }                                                                                            \
            } else if (c_ - 'A' < 26u) {                                        /* a command's two letters */ \
                const bool m1_ = NUTSB_CTX((k) - 1) == '~' && NUTSB_CTX((k) - 2) != '/' &&                   \
                                 nutsb_code(tab, (u8)c_, (u8)NUTSB_CTX((k) + 1)) >= 0;                       \
                const bool m2_ = NUTSB_CTX((k) - 2) == '~' && NUTSB_CTX((k) - 3) != '/' &&                   \
                                 nutsb_code(tab, (u8)NUTSB_CTX((k) - 1), (u8)c_) >= 0;                       \
                if (m1_ || m2_) { on_len = 0; off_len = 0; }                                                 \
            }                                                                                                \
        }                                                                                                    \
    } while (0)

#define NUTSB_EMIT_ON(d, len, v)                                                                             \
    do {                                                                                                     \
        if (len > 0) (d)[0] = (u8)(v);                                                                       \
        if (len > 1) (d)[1] = (u8)((v) >> 8);                                                                \
        if (len > 2) (d)[2] = (u8)((v) >> 16);                                                               \
        if (len > 3) (d)[3] = (u8)((v) >> 24);                                                               \
        if (len > 4) (d)[4] = (u8)((v) >> 32);                                                               \
        if (len > 5) (d)[5] = (u8)((v) >> 40);                                                               \
        (d) += len;                                                                                          \
    } while (0)
#define NUTSB_EMIT_OFF(d, len, c)                                                                            \
    do {                                                                                                     \
        if (len > 0) (d)[0] = len == 2 ? (u8)'\n' : (u8)(c);                                                 \
        if (len > 1) (d)[1] = (u8)'\r';                                                                      \
        (d) += len;                                                                                          \
    } while (0)

__global__ void __launch_bounds__(NUTSB_REN_THREADS)
k_render(RenderArgs A)
{
    __shared__ __align__(16) u8 s_on[NUTSB_REN_THREADS / 32][NUTSB_REN_ON_WIN + 64];
    __shared__ __align__(16) u8 s_off[NUTSB_REN_THREADS / 32][NUTSB_REN_OFF_WIN + 64];
    __shared__ u8 s_tab[NUTSB_CODETAB_BYTES];
    for (int i = threadIdx.x; i < NUTSB_CODETAB_BYTES; i += NUTSB_REN_THREADS) s_tab[i] = A.codetab[i];
    __syncthreads();
    const u8 *const tab = s_tab;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 n_slab = *A.n_slab;
    const u32 gbase = (blockIdx.x * (NUTSB_REN_THREADS / 32) + (u32)warp) * NUTSB_REN_OPS;
    if (gbase >= n_slab) return;                               // whole warp leaves together
    const u32 cnt = n_slab - gbase < NUTSB_REN_OPS ? n_slab - gbase : NUTSB_REN_OPS;

    // -- lane q < cnt holds op q: window = the aligned words that hold the string
    u32 al = 0, n = 0, fl = 0, nw = 0; u64 gw = 0;
    if ((u32)lane < cnt) {
        const u32 op = A.bl_op[gbase + lane];
        const u64 t0 = A.ops.toff[op];
        n = (u32)(A.ops.toff[op + 1] - t0);
        const u8 *src = A.ops.text + t0;
        al = (u32)((size_t)src & 3);
        gw = (u64)(size_t)(src - al);
        nw = (al + n + 3) >> 2; if (!nw) nw = 1;                // an empty string still ends in a reset (c:1365)
        fl = A.ops.flags[op];
    }
    u32 inc = nw;
    for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(NUTSB_FULL, inc, d); if (lane >= d) inc += t; }
    const u32 P = inc - nw;                                    // first flat word of op q
    const u32 W = __shfl_sync(NUTSB_FULL, inc, 31);
    const u64 on0 = A.vp_on[gbase], off0 = A.vp_off[gbase];
    u8 *g_on = A.slab + on0, *g_off = A.slab + A.off_base + off0;
    u8 *const w_on = s_on[warp], *const w_off = s_off[warp];
    u32 fill_on = 0, fill_off = 0, qcount = 0;

    for (u32 F = 0; F < W; F += 32) {
        // -- which op does flat word F + lane belong to
        const bool starts = (u32)lane < cnt && P >= F && P < F + 32;
        const u32 heads = __reduce_or_sync(NUTSB_FULL, starts ? 1u << (P - F) : 0u);
        u32 q = qcount + (u32)__popc(heads & (0xffffffffu >> (31 - lane))) - 1;
        qcount += (u32)__popc(heads);
        if (q >= cnt) q = cnt - 1;
        const u32 Pq = __shfl_sync(NUTSB_FULL, P, (int)q), alq = __shfl_sync(NUTSB_FULL, al, (int)q);
        const u32 nq = __shfl_sync(NUTSB_FULL, n, (int)q), flq = __shfl_sync(NUTSB_FULL, fl, (int)q);
        const u32 nwq = __shfl_sync(NUTSB_FULL, nw, (int)q);
        const u32 *gwq = (const u32 *)(size_t)__shfl_sync(NUTSB_FULL, gw, (int)q);
        const bool act = F + lane < W;
        const u32 wi = F + lane - Pq;
        u32 x = 0, pv = 0, nx = 0;
        const i32 lo = (i32)alq, hi = (i32)(alq + nq);
        if (act) {
            x = __ldg(gwq + wi) & nutsb_keep_mask((i32)(4 * wi), lo, hi);
            if (wi > 0) pv = __ldg(gwq + wi - 1) & nutsb_keep_mask((i32)(4 * wi) - 4, lo, hi);
            if (wi + 1 < nwq) nx = __ldg(gwq + wi + 1) & nutsb_keep_mask((i32)(4 * wi) + 4, lo, hi);
        }
        const bool colour = !(flq & NUTSB_OF_PLAIN);                        // more(NULL,...): c:2259
        const bool tail = act && wi == nwq - 1 && colour && !(flq & NUTSB_OF_PAGER);   // c:1365; the pager has no such reset
        const u32 vm = act ? nutsb_keep_mask((i32)(4 * wi), lo, hi) : 0u;
        u32 on0_, on1_, on2_, on3_, of0_, of1_, of2_, of3_; u64 v0_, v1_, v2_, v3_;
        NUTSB_CLASSIFY(0, (vm & 0xffu) != 0, colour, on0_, of0_, v0_);
        NUTSB_CLASSIFY(1, (vm & 0xff00u) != 0, colour, on1_, of1_, v1_);
        NUTSB_CLASSIFY(2, (vm & 0xff0000u) != 0, colour, on2_, of2_, v2_);
        NUTSB_CLASSIFY(3, (vm & 0xff000000u) != 0, colour, on3_, of3_, v3_);
        const u32 t_on = on0_ + on1_ + on2_ + on3_ + (tail ? 4u : 0u), t_off = of0_ + of1_ + of2_ + of3_;
        u32 sc = t_on | (t_off << 16);
        for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(NUTSB_FULL, sc, d); if (lane >= d) sc += t; }
        const u32 tot = __shfl_sync(NUTSB_FULL, sc, 31);
        {
            u8 *d = w_on + fill_on + (sc & 0xffffu) - t_on;
            NUTSB_EMIT_ON(d, on0_, v0_); NUTSB_EMIT_ON(d, on1_, v1_); NUTSB_EMIT_ON(d, on2_, v2_); NUTSB_EMIT_ON(d, on3_, v3_);
            if (tail) { d[0] = 0x1b; d[1] = '['; d[2] = '0'; d[3] = 'm'; }
            u8 *e = w_off + fill_off + (sc >> 16) - t_off;
            NUTSB_EMIT_OFF(e, of0_, v0_); NUTSB_EMIT_OFF(e, of1_, v1_); NUTSB_EMIT_OFF(e, of2_, v2_); NUTSB_EMIT_OFF(e, of3_, v3_);
        }
        fill_on += tot & 0xffffu; fill_off += tot >> 16;
        const bool last = F + 32 >= W;
        if (last || fill_on + NUTSB_REN_ON_ROUND > NUTSB_REN_ON_WIN) {
            __syncwarp();
            nutsb_warp_copy(g_on, w_on, fill_on, lane);
            g_on += fill_on; fill_on = 0;
            __syncwarp();
        }
        if (last || fill_off + NUTSB_REN_OFF_ROUND > NUTSB_REN_OFF_WIN) {
            __syncwarp();
            nutsb_warp_copy(g_off, w_off, fill_off, lane);
            g_off += fill_off; fill_off = 0;
            __syncwarp();
        }
    }
    // -- consistency with k_measure's lengths; bytes read
    if (lane == 0) {
        if ((u64)(g_on - A.slab) != A.vp_on[gbase + cnt] || (u64)(g_off - A.slab) != A.off_base + A.vp_off[gbase + cnt])
            atomicOr(A.status, NUTSB_ST_RENDER_MISMATCH);
    }
    u32 nsum = n;
    for (int d = 16; d; d >>= 1) nsum += __shfl_xor_sync(NUTSB_FULL, nsum, d);
    if (lane == 0) nutsb_add64(A.counters + 2, (u64)nsum);
}

// ---- H. fan-out -------------------------------------------------------------------------------
// The kernel that moves the bytes, and nothing else: one block per work item
// (room, tile, chunk of recipients).  It loads the tile's two renderings from the
// slab buffer into shared memory (contiguous, 16-byte loads), loads the item's runs,
// and its warps copy run after run to the recipients' streams with nutsb_warp_copy
// (coalesced 16-byte stores at whatever byte alignment the stream position has).
// A tile whose renderings exceed the shared-memory windows (strings of many hundred
// bytes) is copied slab -> stream directly, same routine, global source.
#ifndef NUTSB_FAN_THREADS
#define NUTSB_FAN_THREADS 256
#endif
#ifndef NUTSB_FAN_MINBLOCKS
#define NUTSB_FAN_MINBLOCKS 5
#endif
#ifndef NUTSB_FAN_GROUP
#define NUTSB_FAN_GROUP 1                          // warps that copy one run together
#endif
#define NUTSB_FAN_SMEM (NUTSB_FAN_ON_CAP + 80 + NUTSB_FAN_OFF_CAP + 80 + 16 * NUTSB_FAN_RUN_CAP)

struct FanoutArgs {
    const ItemDesc *items; const uint4 *runs;
    const u8 *slab; u64 off_base;
    u8 *out;
};

__global__ void __launch_bounds__(NUTSB_FAN_THREADS, NUTSB_FAN_MINBLOCKS)
k_fanout(FanoutArgs A)
{
    NUTSB_DYN_SMEM(s_dyn);
    u8 *const s_on = s_dyn;
    u8 *const s_off = s_dyn + NUTSB_FAN_ON_CAP + 80;
    uint4 *const s_run = (uint4 *)(s_dyn + NUTSB_FAN_ON_CAP + 80 + NUTSB_FAN_OFF_CAP + 80);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const ItemDesc d = A.items[blockIdx.x];
    const bool staged = d.on_len <= NUTSB_FAN_ON_CAP && d.off_len <= NUTSB_FAN_OFF_CAP;
    const u32 a_on = (u32)(d.on_src & 15), a_off = (u32)(d.off_src & 15);
    if (staged) {
        const uint4 *g = (const uint4 *)(A.slab + (d.on_src - a_on));
        for (u32 v = (u32)tid, nv = (a_on + d.on_len + 15) >> 4; v < nv; v += NUTSB_FAN_THREADS) ((uint4 *)s_on)[v] = __ldg(g + v);
        g = (const uint4 *)(A.slab + (d.off_src - a_off));
        for (u32 v = (u32)tid, nv = (a_off + d.off_len + 15) >> 4; v < nv; v += NUTSB_FAN_THREADS) ((uint4 *)s_off)[v] = __ldg(g + v);
    }
    for (u32 r0 = 0; r0 < d.run_cnt; r0 += NUTSB_FAN_RUN_CAP) {
        const u32 nr = d.run_cnt - r0 < NUTSB_FAN_RUN_CAP ? d.run_cnt - r0 : NUTSB_FAN_RUN_CAP;
        if (r0) __syncthreads();                               // the previous batch has been consumed
        for (u32 r = (u32)tid; r < nr; r += NUTSB_FAN_THREADS) s_run[r] = __ldg(A.runs + d.run_begin + r0 + r);
        __syncthreads();
        for (u32 r = (u32)warp / NUTSB_FAN_GROUP; r < nr; r += NUTSB_FAN_THREADS / 32 / NUTSB_FAN_GROUP) {
            const uint4 run = s_run[r];
            const u64 so = (u64)run.z | ((u64)(run.w >> 24) << 32);
            const u32 len = run.w & 0xffffffu;
            u8 *dst = A.out + (((u64)run.y << 32) | run.x);
            const u8 *src;
            if (!staged) src = A.slab + so;
            else if (so >= A.off_base) src = s_off + a_off + (u32)(so - d.off_src);
            else src = s_on + a_on + (u32)(so - d.on_src);
            nutsb_group_copy<32 * NUTSB_FAN_GROUP>(dst, src, len, tid % (32 * NUTSB_FAN_GROUP));
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
    u64 slab_on, slab_off;       // bytes of the slab's two renderings
    u32 n_slab, n_events, items, tiles;
};

// Per room: tiles, (tile, recipient) cells and fan-out work items; exclusive
// prefixes over rooms.  One block.
__global__ void __launch_bounds__(NUTSB_SCAN_THREADS)
k_geometry(PopView pop, const u32 *room_b_off, const u64 *stream_off, const u32 *counts, const u64 *vp_on, const u64 *vp_off,
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
            cells = tiles * users;
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
        sz->slab_on = vp_on[counts[0]]; sz->slab_off = vp_off[counts[0]];
    }
}
