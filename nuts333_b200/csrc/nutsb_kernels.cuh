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
    u32 has_clones;              // some user is flagged NUTSB_UF_CLONE
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
#define NUTSB_SLAB_TOT_WAYS 32
#define NUTSB_MEASURE_WARP_BYTES 4096
#define NUTSB_MEASURE_LIST 256

// 0x80 in every byte of y that is zero, exact (no borrow between bytes)
__device__ __forceinline__ u32 nutsb_zero_bytes(u32 y)
{
    return ~(((y & 0x7f7f7f7fu) + 0x7f7f7f7fu) | y | 0x7f7f7f7fu);
}

__global__ void __launch_bounds__(NUTSB_MEASURE_THREADS)
k_measure(OpsView ops, PopView pop, u32 *len_on, u32 *len_off, u32 *nrep, u32 *status, u64 *slab_tot)
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
    // persistent blocks: a warp takes 32 ops after 32 ops (the command table is loaded once per block, and a warp that
    // waits for its loads leaves the issue slots to the others instead of to a block still starting up)
    u64 s_on = 0, s_off = 0;
    for (i64 wbase = ((i64)blockIdx.x * (NUTSB_MEASURE_THREADS / 32) + warp) * 32; wbase < ops.n; wbase += (i64)gridDim.x * NUTSB_MEASURE_THREADS) {
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
    u32 st = 0, lon = 0, loff = 0, rep = 0;
    u32 kind = NUTSB_OP_NONE;
    if (have && bad) { atomicOr(status, NUTSB_ST_BAD_OFFSETS); len_on[i] = len_off[i] = nrep[i] = 0; }
    else if (have && toolong) { atomicOr(status, NUTSB_ST_TEXT_TOO_LONG); len_on[i] = len_off[i] = nrep[i] = 0; }
    else if (have) {
        loff = n - drops - 3 * (m4 + m5) + nl;
        const u32 of = ops.flags[i];
        if (of & NUTSB_OF_RAW) loff = n;                                       // no byte machine: c:1303
        lon = (of & (NUTSB_OF_PLAIN | NUTSB_OF_RAW)) ? loff                                    // colour setting ignored
            : loff + 4 * nl + 4 * m4 + 5 * m5 + ((of & NUTSB_OF_PAGER) ? 0u : 4u);   // pager lines carry no final reset
        len_off[i] = loff; len_on[i] = lon;

        // fan-in: how many room lists this op enters
        kind = ops.kind[i];
        const i32 tgt = ops.target[i], exc = ops.except_user[i];
        if (kind == NUTSB_OP_USER) {
            if (tgt >= pop.n_users) st |= NUTSB_ST_BAD_INDEX;
            else if (tgt >= 0) rep = (pop.has_clones && (pop.cls_flags[pop.user_cls[tgt]] & (NUTSB_UF_CLONE | NUTSB_UF_REMOTE))) ? 0u : 1u;   // no socket of its own
        } else if (kind == NUTSB_OP_ROOM) {
            if (tgt >= pop.n_rooms || tgt < -1) st |= NUTSB_ST_BAD_INDEX;
            else rep = tgt >= 0 ? 1u : (u32)pop.n_rooms;
        } else if (kind == NUTSB_OP_LEVEL) {
            rep = (u32)pop.n_rooms_tot; st |= NUTSB_ST_HAS_LEVEL;
        } else if (kind != NUTSB_OP_NONE) st |= NUTSB_ST_BAD_KIND;
        if (kind != NUTSB_OP_NONE && (exc >= pop.n_users || exc < -1)) st |= NUTSB_ST_BAD_INDEX;
        if (st & ~NUTSB_ST_HAS_LEVEL) rep = 0;
        if (!live) rep = 0;
        nrep[i] = rep;
        if (st) atomicOr(status, st);
    }
    // bytes of the slab's two renderings: every (room, op) entry of a room / level op is rendered once per
    // setting.  Known here already, so the slab buffer can be sized at the first read-back.
    const bool slab = rep && kind != NUTSB_OP_USER;
    if (slab) { s_on += (u64)lon * rep; s_off += (u64)loff * rep; }
    __syncwarp();                                              // (the staging window and the counters are reused)
    }
    for (int d = 16; d; d >>= 1) { s_on += __shfl_xor_sync(NUTSB_FULL, s_on, d); s_off += __shfl_xor_sync(NUTSB_FULL, s_off, d); }
    if (lane == 0 && (s_on | s_off)) {                         // NUTSB_SLAB_TOT_WAYS pairs: same-address atomics serialise
        u64 *t = slab_tot + 2 * (blockIdx.x % NUTSB_SLAB_TOT_WAYS);
        nutsb_add64(t, s_on); nutsb_add64(t + 1, s_off);
    }
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

// ---- stable LSD radix sort of (key,val) pairs, digits of up to 8 bits ------------------------
#define NUTSB_RS_BITS    8                         // digit width: 8 warps x 256 counters of shared memory in the scatter
#define NUTSB_RS_DIGITS  (1 << NUTSB_RS_BITS)
#define NUTSB_RS_THREADS 256
#define NUTSB_RS_ROUNDS  8
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
// Stable: a warp owns a contiguous piece of the block's chunk (NUTSB_RS_ROUNDS x 32 keys) and counts digits into
// its OWN row of shared-memory counters (match_any ranking inside a round), so the rounds need no barrier between
// the warps; one pass over the rows then turns the counts into each warp's first output slot per digit.
__global__ void __launch_bounds__(NUTSB_RS_THREADS)
k_rs_scatter(const u32 *keys_in, const u32 *vals_in, i64 n_host, const u32 *n_dev, int shift, u32 bits,
             const u64 *offs, u32 nblocks, u32 *keys_out, u32 *vals_out)
{
    const i64 n = n_dev ? (i64)*n_dev : n_host;
    const u32 digits = 1u << bits;
    __shared__ u32 s_cnt[NUTSB_RS_THREADS / 32][NUTSB_RS_DIGITS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (u32 i = threadIdx.x; i < (NUTSB_RS_THREADS / 32) * NUTSB_RS_DIGITS; i += blockDim.x) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const i64 base = (i64)blockIdx.x * NUTSB_RS_CHUNK + (i64)warp * (NUTSB_RS_ROUNDS * 32);
    u32 rk[NUTSB_RS_ROUNDS], ky[NUTSB_RS_ROUNDS];          // rank of the key among the warp's keys of the same digit; the keys
#pragma unroll
    for (int r = 0; r < NUTSB_RS_ROUNDS; ++r) {            // every load on its way before the first rank is taken (the kernel is latency-bound)
        const i64 i = base + r * 32 + lane;
        ky[r] = i < n ? keys_in[i] : 0u;
    }
#pragma unroll
    for (int r = 0; r < NUTSB_RS_ROUNDS; ++r) {
        const i64 i = base + r * 32 + lane;
        const bool valid = i < n;
        const u32 d = valid ? ((ky[r] >> shift) & (digits - 1)) : 0xffffffffu;
        const u32 grp = __match_any_sync(NUTSB_FULL, d);
        const int leader = __ffs((int)grp) - 1;
        u32 first = 0;
        if (valid && lane == leader) { first = s_cnt[warp][d]; s_cnt[warp][d] = first + (u32)__popc(grp); }
        first = __shfl_sync(NUTSB_FULL, first, leader);
        rk[r] = first + (u32)__popc(grp & ((1u << lane) - 1));
        __syncwarp();
    }
    __syncthreads();
    for (u32 d = threadIdx.x; d < digits; d += blockDim.x) {   // counts -> first slots: the block's base, then warp after warp
        u32 run = (u32)offs[(size_t)d * nblocks + blockIdx.x];
        for (int w = 0; w < NUTSB_RS_THREADS / 32; ++w) { const u32 t = s_cnt[w][d]; s_cnt[w][d] = run; run += t; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < NUTSB_RS_ROUNDS; ++r) {
        const i64 i = base + r * 32 + lane;
        if (i < n) {
            const u32 key = ky[r];
            const u32 dst = s_cnt[warp][(key >> shift) & (digits - 1)] + rk[r];
            keys_out[dst] = key;
            vals_out[dst] = vals_in ? vals_in[i] : (u32)i;
        }
    }
}

// seg_off[k] = first index i with keys[i] >= k, for k in [0, nkeys]; keys sorted.  Thread i in [0, n] writes the
// entries between its left neighbour's key and its own (folded into the kernels that walk the sorted keys anyway).
__device__ __forceinline__ void nutsb_seg_bounds(const u32 *keys, i64 n, i64 i, u32 nkeys, u32 *seg_off)
{
    if (i > n) return;
    const i64 lo = (i == 0) ? 0 : (i64)keys[i - 1] + 1;
    const i64 hi = (i == n) ? (i64)nkeys : (i64)keys[i];
    for (i64 k = lo; k <= hi && k <= (i64)nkeys; ++k) seg_off[k] = (u32)i;
}

// ---- C. per-entry classification -----------------------------------------------------
// e_info packs, per room-list entry: bit0 = enters the room slab (room/level op),
// bits 1-2 = event kind (0 none, 1 direct write_user op, 2 excluded recipient, 3 both in one:
// a write_user to u directly followed, in the same room list, by a room op that excludes u --
// say(), shout(), tell() ... all write "You ..." to the speaker and the line to everybody else).
// Event keys: 4 * (slab rank the event sits before) + 0 direct | 1 direct-and-skip | 2 skip.
#define NUTSB_EV_DIRECT  0u
#define NUTSB_EV_REPLACE 1u
#define NUTSB_EV_SKIP    2u
struct EntryArrays {
    const u32 *e_room, *e_op;    // sorted by room, op order inside
    u8  *e_info;
    i32 *e_delta;                // event: signed byte delta on the user's stream
    u32 *e_slot;                 // event: user slot
};

__global__ void __launch_bounds__(256)
k_entry_info(OpsView ops, PopView pop, EntryArrays ea, i64 n_ent, const u32 *len_on, const u32 *len_off, u32 *room_ent_off)
{
    const i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    nutsb_seg_bounds(ea.e_room, n_ent, e, (u32)pop.n_rooms_tot, room_ent_off);     // room_ent_off[r] = first entry of room r
    if (e >= n_ent) return;
    const u32 op = ea.e_op[e], room = ea.e_room[e];
    const u32 kind = ops.kind[op];
    u8 info = 0; i32 delta = 0; u32 slot = 0;
    if (kind == NUTSB_OP_USER) {
        const i32 u = ops.target[op];
        const i32 k = pop.user_cls[u];
        const u32 cf = pop.cls_flags[k];
        info = 1u << 1;
        delta = (i32)((cf & NUTSB_UF_COLOUR) ? len_on[op] : len_off[op]);
        slot = (u32)pop.user_slot[u];
        if (e + 1 < n_ent && ea.e_room[e + 1] == room) {     // followed by a room op that excludes u: one event
            const u32 op2 = ea.e_op[e + 1];
            const u32 kind2 = ops.kind[op2];
            if (kind2 != NUTSB_OP_USER && ops.except_user[op2] == u &&
                nutsb_class_delivers(cf, pop.cls_level[k], kind2, ops.flags[op2], ops.target[op2])) {
                info = 3u << 1;
                delta -= (i32)((cf & NUTSB_UF_COLOUR) ? len_on[op2] : len_off[op2]);
            }
        }
    } else {
        info = 1;
        const i32 x = ops.except_user[op];
        if (x >= 0 && (u32)pop.user_room[x] == room) {
            const i32 k = pop.user_cls[x];
            const u32 cf = pop.cls_flags[k];
            if (nutsb_class_delivers(cf, pop.cls_level[k], kind, ops.flags[op], ops.target[op])) {
                bool merged = false;                           // the write_user just before carries the exclusion
                if (e > 0 && ea.e_room[e - 1] == room) {
                    const u32 op0 = ea.e_op[e - 1];
                    merged = ops.kind[op0] == NUTSB_OP_USER && ops.target[op0] == x;
                }
                if (!merged) {
                    info |= 2u << 1;
                    delta = -(i32)((cf & NUTSB_UF_COLOUR) ? len_on[op] : len_off[op]);
                    slot = (u32)pop.user_slot[x];
                }
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
        s.ev_ukey[evi]  = 4u * (rankB - roomB0) + (ek == 1 ? NUTSB_EV_DIRECT : ek == 3 ? NUTSB_EV_REPLACE : NUTSB_EV_SKIP);
        s.ev_delta[evi] = s.e_delta[e];
        s.ev_op[evi]    = s.e_op[e];
    }
}

// room_b_off[r] = slab rank of room r's first entry, r in [0, Rt]; counts[0] = slab ops, counts[1] = events (the
// packed entry scan's total)
__global__ void __launch_bounds__(256)
k_room_b_off(const u32 *room_ent_off, const u64 *e_scan, i64 n_ent, u32 n_rooms_tot, u32 *room_b_off, u32 *counts)
{
    const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) { const u64 t = e_scan[n_ent]; counts[0] = (u32)t; counts[1] = (u32)(t >> 32); }
    if (r > n_rooms_tot) return;
    room_b_off[r] = (u32)e_scan[room_ent_off[r]];
}

// gather the sorted event arrays through the sort permutation
__global__ void __launch_bounds__(256)
k_ev_gather(const u32 *perm, const u32 *n_dev, const u32 *ev_ukey, const i32 *ev_delta, const u32 *ev_op,
            u32 *sv_ukey, i32 *sv_delta, u32 *sv_op, const u32 *sv_slot, u32 n_users, u32 *ev_off)
{
    const i64 n_ev = (i64)*n_dev;
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    nutsb_seg_bounds(sv_slot, n_ev, i, n_users, ev_off);                           // ev_off[s] = first event of slot s
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


// ---- F. copy plan ---------------------------------------------------------------------------
// A room's slab ops are cut into tiles of NUTSB_TILE_OPS; a cell is one (tile,
// recipient) pair.  Every recipient's view of a tile is a short list of RUNS:
// contiguous pieces of the rendered slab (k_render), cut only where the
// recipient is the excluded speaker (c:1415), where a direct write_user op's
// bytes interleave, or -- for recipients behind a filter (login / ignall /
// ignshout, write_level) -- where an op is not delivered to the recipient's
// class.  k_plan leaves a flat run list (each work item's runs contiguous, in
// recipient order) plus one descriptor per fan-out work item, so that the kernel
// that moves the bytes (k_fanout) has nothing to decide.
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
#ifdef NUTSB_EXPERIMENT_ALIGN     // timing experiment only (wrong bytes): every run starts and ends on a sector boundary
    dst &= ~(u64)(NUTSB_EXPERIMENT_ALIGN - 1); len = (len + NUTSB_EXPERIMENT_ALIGN - 1) & ~(u32)(NUTSB_EXPERIMENT_ALIGN - 1);
#endif
    return make_uint4((u32)dst, (u32)(dst >> 32), (u32)src, (len & 0xffffffu) | ((u32)(src >> 32) << 24));
}
#define NUTSB_RUN_MAX_LEN 0xffffffu

// What k_plan and k_direct need to know about a recipient, gathered once per batch (one 48-byte record
// instead of a chain of dependent loads through five arrays in every cell and every event).
struct SlotInfo {
    u64 so_u;                    // the recipient's stream starts here
    u64 end;                     // ... and ends here
    u64 base;                    // stream position of slab rank g after e events = base + cpx.at(k, room, g) + sv_pre[e]
    u32 b0, nb_room;             // the room's first slab rank, its slab ops
    u32 e0, e1;                  // the recipient's events
    u32 room; i32 k;             // room, class
    u32 cf_lv;                   // recipient flags | level << 8
    u32 pad_;
};

struct SlotInfoArgs {
    PopView pop; ClassPrefix cpx;
    const u32 *room_b_off, *ev_off; const u64 *sv_pre, *stream_off;
    SlotInfo *out;
};

__global__ void __launch_bounds__(128)
k_slot_info(SlotInfoArgs A)
{
    const u32 s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= (u32)A.pop.n_users) return;
    const i32 u = A.pop.slot_user[s];
    SlotInfo si;
    si.room = (u32)A.pop.user_room[u]; si.k = A.pop.user_cls[u];
    si.b0 = A.room_b_off[si.room]; si.nb_room = A.room_b_off[si.room + 1] - si.b0;
    si.e0 = A.ev_off[s]; si.e1 = A.ev_off[s + 1];
    si.so_u = A.stream_off[u]; si.end = A.stream_off[u + 1];
    si.base = si.so_u - A.cpx.at(si.k, si.room, si.b0) - A.sv_pre[si.e0];
    si.cf_lv = (u32)A.pop.slot_cf[s] | ((u32)A.pop.slot_lv[s] << 8);
    si.pad_ = 0;
    A.out[s] = si;
}

struct PlanArgs {
    PopView pop; Geometry geo; ClassPrefix cpx;
    const u64 *stream_off;
    const u32 *ev_off, *sv_ukey; const i32 *sv_delta; const u64 *sv_pre;
    const u32 *bl_meta;          // per slab op: kind | flags << 8 | clamped target << 16
    u64 n_cells, off_base;       // off_base: where the colour-off renderings start in the slab buffer
    u32 has_level;
    const SlotInfo *slots;
    u32 *run_cursor;             // runs reserved so far
    uint4 *runs; ItemDesc *items;
    u64 *counters;               // [0] deliveries
    u32 *status;
};

// The walk of one cell (tile t of the room, recipient ls of the room): counts its runs (FILL = false) or
// writes them from A.runs[r_out] on (FILL = true).
// `first`: index of the recipient's first event inside the tile -- found by binary search in the count pass
// (FILL = false) and handed back to the fill pass.
template <bool FILL>
__device__ __forceinline__ u32 nutsb_plan_cell(const PlanArgs &A, u32 room, u32 t, u32 ls, u64 r_out, u32 &deliv, u32 &first)
{
    const SlotInfo si = A.slots[(u32)A.pop.room_slot_off[room] + ls];
    const u32 b0 = si.b0, nb_room = si.nb_room;
    const u32 a0 = t * NUTSB_TILE_OPS;                   // room-local slab rank of the tile's first op
    const u32 nb = nb_room - a0 < NUTSB_TILE_OPS ? nb_room - a0 : NUTSB_TILE_OPS;
    const u32 g0 = b0 + a0;
    const u32 e0 = si.e0, e1 = si.e1;
    // the recipient's events inside the tile: keys in [4*a0+1, 4*(a0+nb)] (a direct op before the tile's first
    // slab op belongs to the tile before); the walk below stops at the first key beyond
    u32 l = first;
    if (!FILL) {
        u32 h = e1; l = e0;
        const u32 thr = 4 * a0 + 1;
        while (l < h) { const u32 mid = (l + h) >> 1; if (A.sv_ukey[mid] < thr) l = mid + 1; else h = mid; }
        first = l;
    }
    const u32 thr_end = 4 * (a0 + nb) + 1;
    const i32 k = si.k;
    const u32 cf = si.cf_lv & 0xffu, clv = si.cf_lv >> 8;
    const bool colour = (cf & NUTSB_UF_COLOUR) != 0;
    const bool full = !A.has_level && !(cf & NUTSB_UF_FILTERED);
    const u64 *vp = (colour ? A.cpx.vp_on : A.cpx.vp_off) + g0;      // tile-local prefix of rendered lengths
    const u64 sb = colour ? 0 : A.off_base;
    // stream position of the tile's first op: class prefix + the recipient's own events before the tile
    const u64 so_u = si.so_u;
    u64 p = si.base + A.cpx.at(k, room, g0) + A.sv_pre[l];
    // Plain listeners: a run covers only whole 32-byte sectors of the stream.  The pieces of a sector
    // that holds a discontinuity (an event of this recipient) are written together by k_direct's seam
    // pass: two partial writes of one sector far apart in time cost a read-modify-write in DRAM.
    // a_start = where the contiguous stretch of slab bytes this run belongs to begins in the stream.
    u64 a_start = so_u; bool first_seg = true;
    if (full && l > e0) {                                 // ... after the last event before the tile
        const u32 uk = A.sv_ukey[l - 1];
        a_start = si.base + A.cpx.at(k, room, b0 + (uk >> 2) + ((uk & 3u) != NUTSB_EV_DIRECT)) + A.sv_pre[l];
        first_seg = false;
    }
    const bool last_tile = a0 + nb == nb_room;
    u32 nruns = 0;
    u32 cur = 0;
    for (u32 e = l; ; ++e) {
        u32 j = nb, ek = NUTSB_EV_SKIP; i32 dlt = 0;
        bool in_tile = false;                                 // event e lies in this tile
        if (e < e1) {
            const u32 uk = A.sv_ukey[e];
            if (uk < thr_end) { in_tile = true; j = (uk >> 2) - a0; ek = uk & 3u; if (ek != NUTSB_EV_SKIP) dlt = A.sv_delta[e]; }
        }
        if (full) {
            if (cur < j) {
                const u64 v0 = vp[cur], v1 = vp[j];
                deliv += j - cur;                                  // zero-length renderings are deliveries too
                if (v1 > v0) {
                    const u64 x0 = p, x1 = p + (v1 - v0), f0 = x0 & ~(u64)31;
                    const u64 xs = a_start <= f0 ? f0 : (first_seg ? x0 : (x0 + 31) & ~(u64)31);
                    const u64 xe = (!in_tile && last_tile) ? x1 : x1 & ~(u64)31;        // the stream's end is kept exact
                    if (xs < xe) {
                        if (FILL) A.runs[r_out + nruns] = nutsb_run_pack(xs, sb + v0 + xs - x0, (u32)(xe - xs));
                        ++nruns;
                    }
                    p = x1;
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
        if (!in_tile) break;
        // a direct op's bytes go here (k_direct writes them); excluded from op j: nothing emitted for it
        if (ek == NUTSB_EV_DIRECT) { p += (u64)(i64)dlt; cur = j; }
        else {
            if (ek == NUTSB_EV_REPLACE) p += (u64)(i64)dlt + (A.cpx.at(k, room, g0 + j + 1) - A.cpx.at(k, room, g0 + j));
            cur = j + 1;
        }
        a_start = p; first_seg = false;
    }
    return nruns;
}

// One block per fan-out work item (room, tile, chunk of recipients), one thread per recipient: count the
// cell's runs, reserve the item's piece of the run list (one atomic per item: the order of the items in the
// list does not matter, each item's runs are contiguous), walk again and write them, write the descriptor.
// FILL = false only adds up the number of runs (sizes the list when recipients sit behind filters).
template <bool FILL>
__global__ void __launch_bounds__(NUTSB_UCHUNK)
k_plan(PlanArgs A)
{
    __shared__ u32 s_room, s_base, s_deliv;
    __shared__ u32 s_w[NUTSB_UCHUNK / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        const u32 item = blockIdx.x;
        u32 lo = 0, hi = (u32)A.pop.n_rooms_tot;            // last room with room_item_off[r] <= item
        while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (A.geo.room_item_off[mid] <= item) lo = mid; else hi = mid; }
        s_room = lo; s_deliv = 0;
    }
    __syncthreads();
    const u32 room = s_room;
    const u32 users_r = (u32)(A.pop.room_slot_off[room + 1] - A.pop.room_slot_off[room]);
    const u32 chunks = (users_r + NUTSB_UCHUNK - 1) / NUTSB_UCHUNK;
    const u32 local = blockIdx.x - A.geo.room_item_off[room];
    const u32 t = local / chunks, ls = (local % chunks) * NUTSB_UCHUNK + (u32)tid;
    const bool valid = ls < users_r;
    u32 deliv = 0;
    u32 first = 0;
    const u32 n = valid ? nutsb_plan_cell<false>(A, room, t, ls, 0, deliv, first) : 0u;
    u32 inc = n;
    for (int d = 1; d < 32; d <<= 1) { const u32 x = __shfl_up_sync(NUTSB_FULL, inc, d); if (lane >= d) inc += x; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    u32 before = 0, total = 0;
    for (int w = 0; w < NUTSB_UCHUNK / 32; ++w) { if (w < warp) before += s_w[w]; total += s_w[w]; }
    if (tid == 0) s_base = total ? atomicAdd(A.run_cursor, total) : 0u;
    if (!FILL) return;
    __syncthreads();
    const u32 base = s_base;
    deliv = 0;
    if (valid && n) (void)nutsb_plan_cell<true>(A, room, t, ls, (u64)base + before + inc - n, deliv, first);
    else if (valid) (void)nutsb_plan_cell<false>(A, room, t, ls, 0, deliv, first);      // deliveries of zero-length renderings
    if (tid == 0) {
        const u32 b0 = A.geo.room_b_off[room], nb_room = A.geo.room_b_off[room + 1] - b0;
        const u32 a0 = t * NUTSB_TILE_OPS;
        const u32 nb = nb_room - a0 < NUTSB_TILE_OPS ? nb_room - a0 : NUTSB_TILE_OPS;
        const u32 g0 = b0 + a0;
        ItemDesc d;                                           // a run may reach up to 31 bytes back into the previous tile
        const u64 n0 = A.cpx.vp_on[g0], o0 = A.cpx.vp_off[g0];
        const u32 xn = n0 < 32 ? (u32)n0 : 32u, xo = o0 < 32 ? (u32)o0 : 32u;
        d.on_src = n0 - xn; d.on_len = (u32)(A.cpx.vp_on[g0 + nb] - n0) + xn;
        d.off_src = A.off_base + o0 - xo; d.off_len = (u32)(A.cpx.vp_off[g0 + nb] - o0) + xo;
        d.run_begin = base; d.run_cnt = total;
        A.items[blockIdx.x] = d;
    }
    for (int d = 16; d; d >>= 1) deliv += __shfl_xor_sync(NUTSB_FULL, deliv, d);
    if (lane == 0 && deliv) atomicAdd(&s_deliv, deliv);
    __syncthreads();
    if (tid == 0 && s_deliv) nutsb_add64(A.counters, (u64)s_deliv);
}

// Bytes of colcode[k] (nuts333.h:237-246), little-endian in one register pair.
__device__ __forceinline__ u64 nutsb_code_pack(int k)
{
    if (k < 5) return 0x1bull | ((u64)'[' << 8) | ((u64)('0' + ((0x75410u >> (4 * k)) & 0xf)) << 16) | ((u64)'m' << 24);
    return 0x1bull | ((u64)'[' << 8) | ((u64)(k < 13 ? '3' : '4') << 16) | ((u64)('0' + ((k - 5) & 7)) << 24) | ((u64)'m' << 32);
}

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

// ---- G. render ------------------------------------------------------------------------------
// write_user's byte machine (c:1315-1365) is position-local: what
// byte i emits depends on bytes i-3..i+2 only.  So a warp takes up to 32 strings as
// ONE flat array of aligned 32-bit words and renders 32 words per round, lane =
// word, whatever the string lengths: every lane finds the '~' '/' '\n' bytes of its
// word with SWAR masks, resolves the (rare) '~XX' commands against the table, a warp
// scan of the emitted lengths places the bytes, and they go into the warp's private
// shared-memory windows, flushed through a sink whenever a window could overflow in
// the next round.  No text staging: a lane reads its word straight from the packed
// text (consecutive lanes read consecutive words of the same string) and gets the
// neighbouring words from the neighbouring lanes.
//
// BOTH = true : both renderings (colour on -> w_on, colour off -> w_off)   [k_render]
// BOTH = false: one rendering, in the colour setting of bit NUTSB_FL_COLOUR of fl  [k_direct]
#define NUTSB_REN_ON_WIN    3072                   // per-warp window, colour on  (a round emits <= 32*28 bytes)
#define NUTSB_REN_OFF_WIN   1536                   // per-warp window, colour off (a round emits <= 32*8 bytes)
#define NUTSB_REN_ON_ROUND  (32 * 28)
#define NUTSB_REN_OFF_ROUND (32 * 8)
#define NUTSB_FL_COLOUR     0x100u                 // internal: the rendering wanted is the colour-on one

// 0xff in byte k of the result iff lo <= wb + k < hi (window byte coordinates)
__device__ __forceinline__ u32 nutsb_keep_mask(i32 wb, i32 lo, i32 hi)
{
    i32 dl = lo - wb; dl = dl < 0 ? 0 : dl;
    i32 dh = wb + 4 - hi; dh = dh < 0 ? 0 : dh;
    const u32 a = dl >= 4 ? 0u : 0xffffffffu << (8 * dl);
    const u32 b = dh >= 4 ? 0u : 0xffffffffu >> (8 * dh);
    return a & b;
}

// Lane q < cnt describes string q: window = the aligned words that hold it (gw = address of
// the first, meta = NUTSB_FLAT_META(al, n, fl): al = offset of the string in it, n = length,
// fl = op flags | NUTSB_FL_COLOUR; nw = words, at least 1).  sink.on(window, bytes) /
// sink.off(window, bytes) take the rendered bytes in order; every string's rendering is
// contiguous in that byte sequence.
#define NUTSB_FLAT_META(al, n, fl) ((u32)(al) | ((u32)(n) << 2) | ((u32)(fl) << 14))     /* n <= NUTSB_MAX_TEXT < 4096 */

template <bool BOTH, class Sink>
__device__ __forceinline__ void nutsb_flat_render(u32 cnt, u32 meta, u64 gw, u32 nw,
                                                  u8 *w_on, u8 *w_off, const u8 *tab, int lane, Sink &sink)
{
    u32 inc = nw;
    for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(NUTSB_FULL, inc, d); if (lane >= d) inc += t; }
    const u32 P = inc - nw;                                    // first flat word of string q
    const u32 W = __shfl_sync(NUTSB_FULL, inc, 31);
    u32 fill_on = 0, fill_off = 0, qcount = 0;
    u32 carry_x = 0, carry_m = 0;                              // lane 31's word and command mask of the previous round

    for (u32 F = 0; F < W; F += 32) {
        // -- which string does flat word F + lane belong to
        const bool starts = (u32)lane < cnt && P >= F && P < F + 32;
        const u32 heads = __reduce_or_sync(NUTSB_FULL, starts ? 1u << (P - F) : 0u);
        u32 q = qcount + (u32)__popc(heads & (0xffffffffu >> (31 - lane))) - 1;
        qcount += (u32)__popc(heads);
        if (q >= cnt) q = cnt - 1;
        const u32 Pq = __shfl_sync(NUTSB_FULL, P, (int)q), mq = __shfl_sync(NUTSB_FULL, meta, (int)q);
        const u32 *gwq = (const u32 *)(size_t)__shfl_sync(NUTSB_FULL, gw, (int)q);
        const u32 alq = mq & 3u, nq = (mq >> 2) & 0xfffu, flq = mq >> 14;
        const u32 endq = alq + nq;                             // window byte after the string's last
        u32 nwq = (endq + 3) >> 2; if (!nwq) nwq = 1;
        const bool act = F + lane < W;
        const u32 wi = F + lane - Pq;
        // -- the lane's word and its neighbours, bytes outside the string zeroed (a string holds no NUL)
        const i32 left = (i32)endq - (i32)(4 * wi);            // bytes of this word before the string's end
        u32 km = left < 4 ? (left > 0 ? (1u << (8 * left)) - 1u : 0u) : 0xffffffffu;
        if (wi == 0) km &= 0xffffffffu << (8 * alq);
        if (!act) km = 0;
        const u32 x = act ? __ldg(gwq + wi) & km : 0u;
        u32 pv = __shfl_up_sync(NUTSB_FULL, x, 1), nx = __shfl_down_sync(NUTSB_FULL, x, 1);
        if (lane == 0) pv = carry_x;
        if (lane == 31) nx = (act && wi + 1 < nwq) ? __ldg(gwq + wi + 1) & (left < 8 ? (1u << (8 * (left - 4))) - 1u : 0xffffffffu) : 0u;
        if (wi == 0) pv = 0;
        if (wi + 1 >= nwq) nx = 0;
        const bool colour = (flq & NUTSB_FL_COLOUR) && !(flq & (NUTSB_OF_PLAIN | NUTSB_OF_RAW));   // more(NULL,...): c:2259
        const bool tail = act && wi == nwq - 1 && colour && !(flq & NUTSB_OF_PAGER);   // c:1365; the pager has no such reset

        // -- special bytes of the word: 0x80 per byte
        const u32 cook = (flq & NUTSB_OF_RAW) ? 0u : 0xffffffffu;               // raw strings have no special bytes
        const u32 tl = nutsb_zero_bytes(x ^ 0x7e7e7e7eu) & cook, sl = nutsb_zero_bytes(x ^ 0x2f2f2f2fu) & cook,
                  nl = nutsb_zero_bytes(x ^ 0x0a0a0a0au) & cook;
        const u32 t_next = (tl >> 8) | ((nx & 0xffu) == '~' ? 0x80000000u : 0u);      // byte k+1 is '~'
        const u32 s_prev = (sl << 8) | ((pv >> 24) == '/' ? 0x80u : 0u);             // byte k-1 is '/'
        const u32 slash_drop = sl & t_next;                                          // c:1330
        u32 ta = tl & ~s_prev;                                                       // '~' not after '/': c:1331-1333
        u32 mm = 0, m5 = 0, codes = 0;                    // matched commands: mask, mask of the 5-byte codes, index per byte
        if (ta) {
            const u64 xn = (u64)x | ((u64)nx << 32);
            do {
                const u32 bit = (u32)__ffs((int)ta) - 1;                            // 7, 15, 23 or 31
                ta &= ta - 1;
                const int kk = nutsb_code(tab, (u8)(xn >> (bit + 1)), (u8)(xn >> (bit + 9)));   // c:1337-1352
                if (kk >= 0) { mm |= 1u << bit; codes |= (u32)kk << (bit - 7); if (kk >= 5) m5 |= 1u << bit; }
            } while (ta);
        }
        u32 mp = __shfl_up_sync(NUTSB_FULL, mm, 1);
        if (lane == 0) mp = carry_m;
        if (wi == 0) mp = 0;
        const u32 consumed = (mm << 8) | (mm << 16) | (mp >> 24) | ((mp >> 16) & 0x8080u);   // a command's two letters
        const u32 plain = (km & 0x80808080u) & ~(slash_drop | consumed | mm | nl);
        // -- bytes emitted per position, one count per byte of the register (SURVEY.md A.1's table)
        const u32 l_off = (plain >> 7) + 2 * (nl >> 7);
        const u32 l_on = colour ? (plain >> 7) + 6 * (nl >> 7) + 4 * (mm >> 7) + (m5 >> 7) : l_off;
        const u32 t_off = (l_off * 0x01010101u) >> 24;
        const u32 t_on = ((l_on * 0x01010101u) >> 24) + (tail ? 4u : 0u);
        u32 sc = BOTH ? (t_on | (t_off << 16)) : t_on;
        for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(NUTSB_FULL, sc, d); if (lane >= d) sc += t; }
        const u32 tot = __shfl_sync(NUTSB_FULL, sc, 31);
        carry_x = __shfl_sync(NUTSB_FULL, x, 31); carry_m = __shfl_sync(NUTSB_FULL, mm, 31);

        // -- emit: the plain bytes at their offsets, then the few '\n' and commands
        {
            u8 *const d = w_on + fill_on + (BOTH ? (sc & 0xffffu) : sc) - t_on;
            u8 *const e = w_off + fill_off + (sc >> 16) - t_off;
            const u32 p_on = l_on * 0x01010100u, p_off = l_off * 0x01010100u;     // exclusive prefix per position
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (plain & (0x80u << (8 * k))) {
                    const u8 c = (u8)(x >> (8 * k));
                    d[(p_on >> (8 * k)) & 0xffu] = c;
                    if (BOTH) e[(p_off >> (8 * k)) & 0xffu] = c;
                }
            }
            u32 sp = colour ? (nl | mm) : nl;
            while (sp) {
                const u32 bit = (u32)__ffs((int)sp) - 1, k8 = bit - 7;
                sp &= sp - 1;
                const bool is_nl = (nl >> bit) & 1u;
                u8 *dd = d + ((p_on >> k8) & 0xffu);
                if (is_nl) {                                                     /* c:1316-1326 */
                    if (colour) { dd[0] = 0x1b; dd[1] = '['; dd[2] = '0'; dd[3] = 'm'; dd += 4; }
                    dd[0] = '\n'; dd[1] = '\r';
                    if (BOTH) { u8 *ee = e + ((p_off >> k8) & 0xffu); ee[0] = '\n'; ee[1] = '\r'; }
                } else {                                                         /* colcode[kk], h:237-246 */
                    const int kk = (int)((codes >> k8) & 0x1fu);
                    const u64 v = nutsb_code_pack(kk);
                    dd[0] = 0x1b; dd[1] = '['; dd[2] = (u8)(v >> 16); dd[3] = (u8)(v >> 24);
                    if (kk >= 5) dd[4] = 'm';
                }
            }
            if (tail) { u8 *dd = d + t_on - 4; dd[0] = 0x1b; dd[1] = '['; dd[2] = '0'; dd[3] = 'm'; }
        }
        fill_on += BOTH ? (tot & 0xffffu) : tot; if (BOTH) fill_off += tot >> 16;
        const bool last = F + 32 >= W;
        if (last || fill_on + NUTSB_REN_ON_ROUND > NUTSB_REN_ON_WIN) {
            __syncwarp();
            sink.on(w_on, fill_on, lane);
            fill_on = 0;
            __syncwarp();
        }
        if (BOTH && (last || fill_off + NUTSB_REN_OFF_ROUND > NUTSB_REN_OFF_WIN)) {
            __syncwarp();
            sink.off(w_off, fill_off, lane);
            fill_off = 0;
            __syncwarp();
        }
    }
}

// Every slab op is rendered ONCE per colour setting (the reference renders it once
// per recipient, c:1427 -> c:1315) into the slab buffer: the colour-on rendering of
// slab rank g at slab + vp_on[g], the colour-off one at slab + off_base + vp_off[g].
// A warp takes NUTSB_REN_OPS consecutive slab ops: its output is one contiguous
// piece of each rendering, written with 16-byte stores.
#define NUTSB_REN_THREADS 256
#define NUTSB_REN_OPS     32                       // slab ops per warp (<= 32)

struct RenderArgs {
    OpsView ops; const u8 *codetab;
    const u32 *bl_op; const u64 *vp_on, *vp_off;
    const u32 *n_slab;
    u8 *slab; u64 off_base;
    u64 *counters;               // [2] source bytes read
    u32 *status;
};

struct SlabSink {
    u8 *g_on, *g_off;
    __device__ __forceinline__ void on(const u8 *w, u32 nbytes, int lane) { nutsb_warp_copy(g_on, w, nbytes, lane); g_on += nbytes; }
    __device__ __forceinline__ void off(const u8 *w, u32 nbytes, int lane) { nutsb_warp_copy(g_off, w, nbytes, lane); g_off += nbytes; }
};

__global__ void __launch_bounds__(NUTSB_REN_THREADS)
k_render(RenderArgs A)
{
    __shared__ __align__(16) u8 s_on[NUTSB_REN_THREADS / 32][NUTSB_REN_ON_WIN + 64];
    __shared__ __align__(16) u8 s_off[NUTSB_REN_THREADS / 32][NUTSB_REN_OFF_WIN + 64];
    __shared__ u8 s_tab[NUTSB_CODETAB_BYTES];
    for (int i = threadIdx.x; i < NUTSB_CODETAB_BYTES; i += NUTSB_REN_THREADS) s_tab[i] = A.codetab[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 n_slab = *A.n_slab;
    // a warp takes units of NUTSB_REN_OPS slab ops, grid-strided (the grid may be smaller than the work:
    // beside the planning kernels k_render runs with a few blocks per SM)
    const u32 wstride = gridDim.x * (NUTSB_REN_THREADS / 32) * NUTSB_REN_OPS;
    u64 nsum = 0;
    for (u32 gbase = (blockIdx.x * (NUTSB_REN_THREADS / 32) + (u32)warp) * NUTSB_REN_OPS; gbase < n_slab; gbase += wstride) {
        const u32 cnt = n_slab - gbase < NUTSB_REN_OPS ? n_slab - gbase : NUTSB_REN_OPS;
        u32 meta = 0, n = 0, nw = 0; u64 gw = 0;
        if ((u32)lane < cnt) {
            const u32 op = A.bl_op[gbase + lane];
            const u64 t0 = A.ops.toff[op];
            n = (u32)(A.ops.toff[op + 1] - t0);
            const u8 *src = A.ops.text + t0;
            const u32 al = (u32)((size_t)src & 3);
            gw = (u64)(size_t)(src - al);
            nw = (al + n + 3) >> 2; if (!nw) nw = 1;            // an empty string still ends in a reset (c:1365)
            nutsb_prefetch(src, n);                             // all 32 strings on their way before the first round
            meta = NUTSB_FLAT_META(al, n, A.ops.flags[op] | NUTSB_FL_COLOUR);
        }
        SlabSink sink{ A.slab + A.vp_on[gbase], A.slab + A.off_base + A.vp_off[gbase] };
        nutsb_flat_render<true>(cnt, meta, gw, nw, s_on[warp], s_off[warp], s_tab, lane, sink);
        // consistency with k_measure's lengths
        if (lane == 0 && ((u64)(sink.g_on - A.slab) != A.vp_on[gbase + cnt] || (u64)(sink.g_off - A.slab) != A.off_base + A.vp_off[gbase + cnt]))
            atomicOr(A.status, NUTSB_ST_RENDER_MISMATCH);
        nsum += n;
    }
    for (int d = 16; d; d >>= 1) nsum += __shfl_xor_sync(NUTSB_FULL, nsum, d);
    if (lane == 0 && nsum) nutsb_add64(A.counters + 2, nsum);
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

// the work of one block: work item `item`, s_dyn = NUTSB_FAN_SMEM bytes of shared memory
__device__ __forceinline__ void nutsb_fanout_block(const FanoutArgs &A, u32 item, u8 *s_dyn)
{
    u8 *const s_on = s_dyn;
    u8 *const s_off = s_dyn + NUTSB_FAN_ON_CAP + 80;
    uint4 *const s_run = (uint4 *)(s_dyn + NUTSB_FAN_ON_CAP + 80 + NUTSB_FAN_OFF_CAP + 80);
    const int tid = threadIdx.x, warp = tid >> 5;
    const ItemDesc d = A.items[item];
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

__global__ void __launch_bounds__(NUTSB_FAN_THREADS, NUTSB_FAN_MINBLOCKS)
k_fanout(FanoutArgs A)
{
    NUTSB_DYN_SMEM(s_dyn);
    nutsb_fanout_block(A, blockIdx.x, s_dyn);
}

// ---- I. direct ops (write_user) -----------------------------------------------------------
// Events are sorted by recipient.  A warp takes 32 consecutive events; the direct ops among
// them (even keys) are rendered with the flat renderer in their recipient's colour setting
// and each rendering is copied into the recipient's stream at the offset the event prefix
// gives.
struct DirectArgs {
    OpsView ops; PopView pop; ClassPrefix cpx;
    const u32 *room_b_off, *ev_off;
    const u32 *sv_ukey, *sv_op; const u64 *sv_pre;
    const u64 *stream_off;
    const u32 *ev_slot_sorted;      // user slot of each sorted event
    const i32 *sv_delta;
    u8 *out; i64 n_ev;
    u64 *n_deliveries;              // [0] deliveries, [1] bytes written by k_direct, [2] (k_render)
    u32 *status;
    const u8 *slab; u64 off_base;   // the rendered slab (k_render): source of the seam bytes
    u32 has_level;
    const SlotInfo *slots;
    const u64 *dpre;                // gather-list mode (k_direct_compact): event e's rendering goes to out + dpre[e]
};

#define NUTSB_DIRECT_THREADS 256

struct DirectSink {                 // string q's rendering is bytes [O, O + osz) of the warp's output; it goes to out + p
    u8 *out; u64 p; u32 O, osz;     // lane q holds string q
    u32 cnt, done, qf;
    __device__ __forceinline__ void on(const u8 *w, u32 nbytes, int lane)
    {
        const u32 end = done + nbytes;
        while (qf < cnt) {
            const u32 Oq = __shfl_sync(NUTSB_FULL, O, (int)qf), zq = __shfl_sync(NUTSB_FULL, osz, (int)qf);
            const u64 pq = __shfl_sync(NUTSB_FULL, p, (int)qf);
            if (Oq >= end && zq) break;
            const u32 x0 = Oq > done ? Oq : done, x1 = Oq + zq < end ? Oq + zq : end;
            if (x1 > x0) {
                u8 *dst = out + pq + (x0 - Oq);
                const u8 *src = w + (x0 - done);
                const u32 len = x1 - x0;
                if (len < 192) { for (u32 i = (u32)lane; i < len; i += 32) dst[i] = src[i]; }
                else nutsb_warp_copy(dst, src, len, lane);
            }
            if (Oq + zq <= end) ++qf; else break;
        }
        done = end;
    }
    __device__ __forceinline__ void off(const u8 *, u32, int) {}
};

// shared memory of one k_direct block, carved from the dynamic window: per-warp render windows, per-warp
// compaction tables, the command table
#define NUTSB_DIR_WARPS (NUTSB_DIRECT_THREADS / 32)
#define NUTSB_DIR_SMEM  (NUTSB_DIR_WARPS * (NUTSB_REN_ON_WIN + 64) + NUTSB_DIR_WARPS * 32 * (8 + 8 + 8) + ((NUTSB_CODETAB_BYTES + 15) & ~15))

// the work of block `blk` of `nblk` (grid-strided over the events), smem = NUTSB_DIR_SMEM bytes.
// COMPACT (gather-list mode, nutsb_write_batch_iov): the renderings are packed one after the other in event
// order (out + dpre[e]) instead of going to their places in the streams, and there are no seams to write.
template <bool COMPACT>
__device__ __forceinline__ void nutsb_direct_block(const DirectArgs &A, u32 blk, u32 nblk, u8 *smem)
{
    typedef u8 OnWin[NUTSB_REN_ON_WIN + 64];
    OnWin *const s_on = (OnWin *)smem;
    u64 (*const s_gw)[32] = (u64 (*)[32])(smem + NUTSB_DIR_WARPS * (NUTSB_REN_ON_WIN + 64));
    u64 (*const s_p)[32] = s_gw + NUTSB_DIR_WARPS;
    u32 (*const s_par)[32][2] = (u32 (*)[32][2])(s_p + NUTSB_DIR_WARPS);
    u8 *const s_tab = (u8 *)(s_par + NUTSB_DIR_WARPS);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < NUTSB_CODETAB_BYTES; i += NUTSB_DIRECT_THREADS) s_tab[i] = A.pop.codetab[i];
    __syncthreads();

    u64 c_cnt = 0, c_bytes = 0; u32 seam_bytes = 0;
    for (i64 ebase = ((i64)blk * (NUTSB_DIRECT_THREADS / 32) + warp) * 32; ebase < A.n_ev;
         ebase += (i64)nblk * NUTSB_DIRECT_THREADS) {
        const i64 e = ebase + lane;
        bool isw = false, seam = false, first_ev = false, last_ev = false;
        u64 p = 0, q = 0, gw = 0; u32 n = 0, al = 0, osz = 0, fl = 0;
        u64 a_start = 0, b_end = 0; const u8 *sa = nullptr, *sb = nullptr;
        if (e < A.n_ev) {
            const u32 uk = A.sv_ukey[e];
            const SlotInfo si = A.slots[A.ev_slot_sorted[e]];
            const u32 room = si.room; const i32 k = si.k;
            const u32 b0 = si.b0, cf = si.cf_lv & 0xffu;
            const u32 e0 = si.e0, e1 = si.e1;
            const u64 so_u = si.so_u, base = si.base;
            const u32 j = uk >> 2, ek = uk & 3u, skip = ek != NUTSB_EV_DIRECT;
            p = base + A.cpx.at(k, room, b0 + j) + A.sv_pre[e];
            q = base + A.cpx.at(k, room, b0 + j + skip) + A.sv_pre[e + 1];     // where the stream goes on after the event
            if (ek != NUTSB_EV_SKIP) {                         // a direct op: its rendering fills [p, q)
                const u32 op = A.sv_op[e];
                const u64 t0 = A.ops.toff[op];
                const u8 *src = A.ops.text + t0;
                n = (u32)(A.ops.toff[op + 1] - t0);
                al = (u32)((size_t)src & 3);
                gw = (u64)(size_t)(src - al);
                nutsb_prefetch(src, n);                         // on its way while the seams are done
                fl = A.ops.flags[op] | ((cf & NUTSB_UF_COLOUR) ? NUTSB_FL_COLOUR : 0u);
                osz = (u32)(q - p);
                isw = true;
            }
            // -- seam: the event is a discontinuity in a plain listener's stream.  k_fanout's runs cover whole
            //    32-byte sectors only; the slab bytes that share a sector with the discontinuity -- the end of the
            //    stretch before it and the start of the stretch after it -- are written here, together.
            seam = !COMPACT && !A.has_level && !(cf & NUTSB_UF_FILTERED);
            if (COMPACT && isw) p = A.dpre[e];                  // (q - p was taken above)
            if (seam) {
                const bool colour = (cf & NUTSB_UF_COLOUR) != 0;
                const u64 *vp = colour ? A.cpx.vp_on : A.cpx.vp_off;
                const u8 *R = A.slab + (colour ? 0 : A.off_base);
                first_ev = (u32)e == e0; last_ev = (u32)e + 1 == e1;
                a_start = so_u; b_end = si.end;
                if (!first_ev && lane == 0) {                  // the other lanes take it from their neighbour
                    const u32 ukp = A.sv_ukey[e - 1];
                    a_start = base + A.cpx.at(k, room, b0 + (ukp >> 2) + ((ukp & 3u) != NUTSB_EV_DIRECT)) + A.sv_pre[e];
                }
                if (!last_ev && (lane == 31 || e + 1 >= A.n_ev))
                    b_end = base + A.cpx.at(k, room, b0 + (A.sv_ukey[e + 1] >> 2)) + A.sv_pre[e + 1];
                sa = R + vp[b0 + j];                           // the stretch before ends where slab op j begins
                sb = R + vp[b0 + j + skip];                    // after an exclusion the stream goes on with op j+1
            }
        }
        if (!COMPACT) {
            // the neighbouring events of the same recipient sit in the neighbouring lanes
            const u64 q_prev = __shfl_up_sync(NUTSB_FULL, q, 1), p_next = __shfl_down_sync(NUTSB_FULL, p, 1);
            if (seam && !first_ev && lane > 0) a_start = q_prev;
            if (seam && !last_ev && lane < 31 && e + 1 < A.n_ev) b_end = p_next;
            u64 t0 = p & ~(u64)31; if (t0 < a_start) t0 = a_start;
            u64 h1 = (q + 31) & ~(u64)31; if (h1 > b_end) h1 = b_end;
            const u32 tlen = seam ? (u32)(p - t0) : 0u, hlen = seam && h1 > q ? (u32)(h1 - q) : 0u;
            seam_bytes += tlen + hlen;
            // Every piece (< 32 bytes) is copied by the whole warp at once: one store request per piece.  All the
            // loads first (into the render window, free until the renderer starts), then all the stores: a load
            // and its store in one loop iteration would make every iteration wait a full trip to L2 / HBM.
            u8 *const stage = s_on[warp];                      // event i: tail piece at [64 i, +32), head piece at [64 i + 32, +32)
            const u32 lens = tlen | (hlen << 8);
            if (__any_sync(NUTSB_FULL, lens != 0)) {
#pragma unroll 8
                for (int i = 0; i < 32; ++i) {
                    const u32 l = __shfl_sync(NUTSB_FULL, lens, i);
                    const u8 *pa = (const u8 *)(size_t)__shfl_sync(NUTSB_FULL, (u64)(size_t)sa, i) - (l & 0xffu);
                    const u8 *pb = (const u8 *)(size_t)__shfl_sync(NUTSB_FULL, (u64)(size_t)sb, i);
                    u8 va = 0, vb = 0;
                    if ((u32)lane < (l & 0xffu)) va = pa[lane];
                    if ((u32)lane < (l >> 8)) vb = pb[lane];
                    stage[64 * i + lane] = va; stage[64 * i + 32 + lane] = vb;
                }
                __syncwarp();
#pragma unroll 4
                for (int i = 0; i < 32; ++i) {
                    const u32 l = __shfl_sync(NUTSB_FULL, lens, i);
                    const u64 da = __shfl_sync(NUTSB_FULL, t0, i), db = __shfl_sync(NUTSB_FULL, q, i);
                    if ((u32)lane < (l & 0xffu)) A.out[da + lane] = stage[64 * i + lane];
                    if ((u32)lane < (l >> 8)) A.out[db + lane] = stage[64 * i + 32 + lane];
                }
                __syncwarp();
            }
        }
        // -- compact the warp's direct ops to lanes 0..cnt-1
        const u32 wmask = __ballot_sync(NUTSB_FULL, isw);
        const u32 cnt = (u32)__popc(wmask);
        if (!cnt) continue;                                    // whole warp
        if (isw) {
            const u32 r = (u32)__popc(wmask & ((1u << lane) - 1));
            s_gw[warp][r] = gw; s_p[warp][r] = p;
            s_par[warp][r][0] = NUTSB_FLAT_META(al, n, fl); s_par[warp][r][1] = osz;
        }
        __syncwarp();
        u32 nw = 0, meta = 0;
        osz = 0; gw = 0; p = 0;
        if ((u32)lane < cnt) {
            gw = s_gw[warp][lane]; p = s_p[warp][lane];
            meta = s_par[warp][lane][0]; osz = s_par[warp][lane][1];
            nw = ((meta & 3u) + ((meta >> 2) & 0xfffu) + 3) >> 2; if (!nw) nw = 1;
        }
        __syncwarp();
        u32 inc = osz;
        for (int d = 1; d < 32; d <<= 1) { const u32 t = __shfl_up_sync(NUTSB_FULL, inc, d); if (lane >= d) inc += t; }
        const u32 total = __shfl_sync(NUTSB_FULL, inc, 31);
        DirectSink sink{ A.out, p, inc - osz, osz, cnt, 0u, 0u };
        nutsb_flat_render<false>(cnt, meta, gw, nw, s_on[warp], (u8 *)nullptr, s_tab, lane, sink);
        if (lane == 0 && sink.done != total) atomicOr(A.status, NUTSB_ST_RENDER_MISMATCH);
        c_cnt += cnt; c_bytes += total;
    }
    for (int d = 16; d; d >>= 1) seam_bytes += __shfl_xor_sync(NUTSB_FULL, seam_bytes, d);
    if (lane == 0 && (c_cnt | seam_bytes)) { nutsb_add64(A.n_deliveries, c_cnt); nutsb_add64(A.n_deliveries + 1, c_bytes + seam_bytes); }
}

#ifndef NUTSB_DIR_MINBLOCKS
#define NUTSB_DIR_MINBLOCKS 5       // <= 51 registers: 0.43 ms against 0.47 at 4 blocks per SM (the kernel is latency-bound)
#endif
__global__ void __launch_bounds__(NUTSB_DIRECT_THREADS, NUTSB_DIR_MINBLOCKS)
k_direct(DirectArgs A)
{
    NUTSB_DYN_SMEM(s_dyn);
    nutsb_direct_block<false>(A, blockIdx.x, gridDim.x, s_dyn);
}

__global__ void __launch_bounds__(NUTSB_DIRECT_THREADS, NUTSB_DIR_MINBLOCKS)
k_direct_compact(DirectArgs A)
{
    NUTSB_DYN_SMEM(s_dyn);
    nutsb_direct_block<true>(A, blockIdx.x, gridDim.x, s_dyn);
}

// ---- H+I in one launch ---------------------------------------------------------------------
// k_fanout keeps the store queues full and issues little (about a fifth of the issue slots); k_direct is
// bound by instruction issue and writes little.  Launched one after the other each has the machine to
// itself; launched as two kernels on two streams the first one's blocks fill every SM and the second
// waits.  So: ONE grid whose blocks are dealt between the two kinds by block index (a direct block every
// `stride` blocks until they are used up), which puts both kinds on every SM for as long as both last.
#define NUTSB_FD_SMEM (NUTSB_FAN_SMEM > NUTSB_DIR_SMEM ? NUTSB_FAN_SMEM : NUTSB_DIR_SMEM)
static_assert(NUTSB_FAN_THREADS == NUTSB_DIRECT_THREADS, "k_fanout_direct runs both bodies with one block size");
#ifndef NUTSB_FD_MINBLOCKS
#define NUTSB_FD_MINBLOCKS 5
#endif
__global__ void __launch_bounds__(NUTSB_FAN_THREADS, NUTSB_FD_MINBLOCKS)
k_fanout_direct(FanoutArgs F, DirectArgs D, u32 n_dir, u32 stride)
{
    NUTSB_DYN_SMEM(s_dyn);
    const u32 b = blockIdx.x;
    if (b % stride == 0 && b / stride < n_dir) { nutsb_direct_block<false>(D, b / stride, n_dir, s_dyn); return; }
    const u32 before = (b + stride - 1) / stride;          // direct blocks with a smaller index
    nutsb_fanout_block(F, b - (before < n_dir ? before : n_dir), s_dyn);
}

// the counters and the status at the end of a write batch, zero-copy as k_readback1
__global__ void k_readback_end(const u64 *counters, const u32 *status, u64 *h_counters, u32 *h_status)
{
    if (threadIdx.x < 3) h_counters[threadIdx.x] = counters[threadIdx.x];
    if (threadIdx.x == 0) *h_status = *status;
}

// ---- stream digests ------------------------------------------------------------------------
// h = fold(h * P + byte), h0 = FNV offset basis.  A block per user: each thread
// folds a contiguous slice as an affine map (mult, add), the block composes them
// in order.
#define NUTSB_DIGEST_P  0x100000001b3ull
#define NUTSB_DIGEST_H0 0xcbf29ce484222325ull

// init != null: the fold goes on from init[u] (the digest of what user u received in earlier batches)
__global__ void __launch_bounds__(256)
k_digest(const u8 *bytes, const u64 *off, i32 n_users, const u64 *init, u64 *digest)
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
            u64 h = init ? init[u] : NUTSB_DIGEST_H0;
            for (int q = 0; q < 256; ++q) h = h * s_m[q] + s_a[q];
            digest[u] = h;
        }
        __syncthreads();
    }
}

// ---- small bookkeeping kernels -----------------------------------------------------------
// read-back #1, written straight into pinned host memory (zero-copy): the number of (room, op) entries, the
// validation status and the slab totals -- one launch instead of three small copies in the copy queue
__global__ void __launch_bounds__(64)
k_readback1(const u64 *eoff_end, const u32 *status, const u64 *slab_tot, u64 *h_entries, u32 *h_status, u64 *h_slab_tot)
{
    const u32 t = threadIdx.x;
    if (t == 0) { *h_entries = *eoff_end; *h_status = *status; }
    if (t < 2 * NUTSB_SLAB_TOT_WAYS) h_slab_tot[t] = slab_tot[t];
}

struct Sizes {                   // read back by the host before the fan-out is launched
    u64 total_bytes;
    u64 cells;
    u64 slab_on, slab_off;       // bytes of the slab's two renderings
    u32 n_slab, n_events, items, tiles;
    u64 direct_bytes;            // gather-list mode: bytes of the direct ops' renderings
};

// Per room: tiles, (tile, recipient) cells and fan-out work items; exclusive
// prefixes over rooms.  One block.
__global__ void __launch_bounds__(NUTSB_SCAN_THREADS)
k_geometry(PopView pop, const u32 *room_b_off, const u64 *stream_off, const u32 *counts, const u64 *vp_on, const u64 *vp_off,
           u32 *room_tile_off, u64 *room_cell_off, u32 *room_item_off, Sizes *sz, const u64 *dpre)
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
        sz->direct_bytes = dpre ? dpre[counts[1]] : 0;
    }
}

// ---- gather lists (nutsb_write_batch_iov) ---------------------------------------------------
// A plain listener's stream is its room's slab in its colour setting, cut at its own events.  Instead of
// copying those bytes once per recipient (k_fanout), hand the host the slab itself (two renderings), the
// direct ops' renderings packed in event order (k_direct_compact) and, per recipient, the list of pieces:
// for every event the slab stretch before it and the direct op's rendering, then the stretch after the
// last event -- 2 * events + 1 entries, zero-length ones included so that every index is known up front
// (slot s, event e: entries 2e + s and 2e + 1 + s; the tail: 2 * e1 + s).  Entries are {address, length}
// in the HOST's copy of the pool: struct iovec as writev(2) takes it.
struct InDirectLen {                // bytes of sorted event e's direct rendering (0 for a pure exclusion)
    const u32 *sv_ukey, *sv_slot; const i32 *sv_delta; const SlotInfo *slots; const u64 *vp_on, *vp_off;
    __device__ u64 operator()(i64 e) const
    {
        const u32 uk = sv_ukey[e], ek = uk & 3u;
        if (ek == NUTSB_EV_SKIP) return 0;
        i64 d = sv_delta[e];                                // direct length minus the excluded op's
        if (ek == NUTSB_EV_REPLACE) {
            const SlotInfo *si = slots + sv_slot[e];
            const u64 *vp = (si->cf_lv & NUTSB_UF_COLOUR) ? vp_on : vp_off;
            const u32 g = si->b0 + (uk >> 2);
            d += (i64)(vp[g + 1] - vp[g]);
        }
        return (u64)d;
    }
};

struct __align__(16) IovEnt { u64 base, len; };    // struct iovec (LP64)
__device__ __forceinline__ IovEnt nutsb_iov_ent(u64 base, u64 len) { IovEnt x; x.base = base; x.len = len; return x; }

struct IovArgs {
    PopView pop; const SlotInfo *slots;
    const u64 *vp_on, *vp_off;
    const u32 *sv_ukey, *sv_slot; const u64 *dpre;
    u32 n_ev;
    u64 host_pool, host_dir;        // host addresses the slab's two renderings and the direct renderings are copied to
    u64 off_base;                   // where the colour-off renderings start in the slab
    IovEnt *iov; u64 *first; u32 *count;
    u64 *deliveries;
};

__global__ void __launch_bounds__(256)
k_iov(IovArgs A)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    i64 deliv = 0;
    if (t < A.n_ev) {
        const u32 e = t, s = A.sv_slot[e];
        const SlotInfo si = A.slots[s];
        const bool colour = (si.cf_lv & NUTSB_UF_COLOUR) != 0;
        const u64 *vp = (colour ? A.vp_on : A.vp_off) + si.b0;
        const u64 R = A.host_pool + (colour ? 0 : A.off_base);
        const u32 uk = A.sv_ukey[e], j = uk >> 2, skip = (uk & 3u) != NUTSB_EV_DIRECT;
        u32 jp = 0;                                         // slab rank the stretch before the event starts at
        if (e != si.e0) { const u32 ukp = A.sv_ukey[e - 1]; jp = (ukp >> 2) + ((ukp & 3u) != NUTSB_EV_DIRECT); }
        const u64 v0 = vp[jp], v1 = vp[j];
        const u64 d0 = A.dpre[e], d1 = A.dpre[e + 1];
        const u64 i0 = 2ull * e + s;
        A.iov[i0] = nutsb_iov_ent(R + v0, v1 - v0);
        A.iov[i0 + 1] = nutsb_iov_ent(A.host_dir + d0, d1 - d0);
        if (e + 1 == si.e1) { const u64 w0 = vp[j + skip]; A.iov[i0 + 2] = nutsb_iov_ent(R + w0, vp[si.nb_room] - w0); }
        deliv -= (i64)skip;
    }
    if (t < (u32)A.pop.n_users) {
        const u32 s = t;
        const SlotInfo si = A.slots[s];
        const i32 u = A.pop.slot_user[s];
        A.first[u] = 2ull * si.e0 + s; A.count[u] = 2u * (si.e1 - si.e0) + 1u;
        if (si.e0 == si.e1) {
            const bool colour = (si.cf_lv & NUTSB_UF_COLOUR) != 0;
            const u64 *vp = (colour ? A.vp_on : A.vp_off) + si.b0;
            A.iov[2ull * si.e0 + s] = nutsb_iov_ent(A.host_pool + (colour ? 0 : A.off_base) + vp[0], vp[si.nb_room] - vp[0]);
        }
        deliv += (i64)si.nb_room;                           // every slab op of the room but the ones it is excluded from
    }
    for (int d = 16; d; d >>= 1) deliv += __shfl_xor_sync(NUTSB_FULL, deliv, d);
    if ((threadIdx.x & 31) == 0 && deliv) nutsb_add64(A.deliveries, (u64)deliv);
}

// ---- parity digests ----------------------------------------------------------------
// d(m,u) = FNV-1a over the bytes of one delivery; per user D_u = the left fold over his deliveries in call order,
// per op D_m = the same fold over its recipients in user-list order:  D <- (D * K) ^ d ^ len,  D0 = 0; a delivery
// of no bytes is not folded (the reference makes no write(2) for it).  Computed from what the last write batch
// left in HBM: a room / level op has two renderings in the slab (so two d's, whoever receives it), a write_user's
// rendering sits in its recipient's stream.  Checking tools, not part of the hot path.
#define NUTSB_DG_K 0x9E3779B97F4A7C15ull

__device__ __forceinline__ u64 nutsb_fnv1a(const u8 *p, u64 n)
{
    u64 h = 0xcbf29ce484222325ull;
    for (u64 i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}

struct DgArgs {
    OpsView ops; PopView pop; ClassPrefix cpx;
    const u32 *bl_op, *bl_meta, *ev_slot_sorted, *sv_ukey, *sv_op; const u64 *sv_pre;
    const SlotInfo *slots; const u8 *slab, *out; u64 off_base;
    const u32 *counts;               // [0] slab ops, [1] events
    u32 has_level;
    const u32 *nrep; const i32 *room_users, *room_users_off;
    u64 *dg_on, *dg_off;             // per slab entry
    u64 *ev_d; u32 *ev_len;          // per event (direct kinds)
    u32 *op_g, *op_ev;               // per op: one of its slab entries / its event (~0: none)
    u64 *per_user, *per_op;
};

__global__ void __launch_bounds__(256) k_dg_slab(DgArgs A)
{
    const u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= A.counts[0]) return;
    const u64 a0 = A.cpx.vp_on[g], a1 = A.cpx.vp_on[g + 1], b0 = A.cpx.vp_off[g], b1 = A.cpx.vp_off[g + 1];
    A.dg_on[g] = nutsb_fnv1a(A.slab + a0, a1 - a0);
    A.dg_off[g] = nutsb_fnv1a(A.slab + A.off_base + b0, b1 - b0);
    A.op_g[A.bl_op[g]] = g;                          // (an all-room op has one entry per room, all with the same bytes)
}

__global__ void __launch_bounds__(256) k_dg_events(DgArgs A)
{
    const u32 e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= A.counts[1]) return;
    const u32 uk = A.sv_ukey[e], j = uk >> 2, ek = uk & 3u;
    if (ek == NUTSB_EV_SKIP) return;
    const SlotInfo si = A.slots[A.ev_slot_sorted[e]];
    const u32 skip = ek != NUTSB_EV_DIRECT;
    const u64 p = si.base + A.cpx.at(si.k, si.room, si.b0 + j) + A.sv_pre[e];
    const u64 q = si.base + A.cpx.at(si.k, si.room, si.b0 + j + skip) + A.sv_pre[e + 1];
    A.ev_d[e] = nutsb_fnv1a(A.out + p, q - p);
    A.ev_len[e] = (u32)(q - p);
    A.op_ev[A.sv_op[e]] = e;
}

__global__ void __launch_bounds__(128) k_dg_user(DgArgs A)
{
    const u32 s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= (u32)A.pop.n_users) return;
    const SlotInfo si = A.slots[s];
    const u32 cf = si.cf_lv & 0xffu, lv = si.cf_lv >> 8;
    const bool colour = (cf & NUTSB_UF_COLOUR) != 0;
    const bool full = !A.has_level && !(cf & NUTSB_UF_FILTERED);
    const u64 *vp = colour ? A.cpx.vp_on : A.cpx.vp_off;
    const u64 *dg = colour ? A.dg_on : A.dg_off;
    u32 e = si.e0;
    u64 D = 0;
    for (u32 a = 0; a <= si.nb_room; ++a) {
        bool skip = false;
        while (e < si.e1 && (A.sv_ukey[e] >> 2) == a) {              // the recipient's own events at this slab rank, in call order
            const u32 ek = A.sv_ukey[e] & 3u;
            if (ek != NUTSB_EV_SKIP) { const u64 len = A.ev_len[e]; if (len) D = (D * NUTSB_DG_K) ^ A.ev_d[e] ^ len; }
            if (ek != NUTSB_EV_DIRECT) skip = true;                   // ... and he is the excluded user of slab op a
            ++e;
        }
        if (a == si.nb_room || skip) continue;
        const u32 g = si.b0 + a;
        const u32 m = A.bl_meta[g];
        if (full || nutsb_class_delivers(cf, lv, m & 0xffu, (m >> 8) & 0xffu, (i32)(int16_t)(m >> 16))) {
            const u64 len = vp[g + 1] - vp[g];
            if (len) D = (D * NUTSB_DG_K) ^ dg[g] ^ len;
        }
    }
    A.per_user[A.pop.slot_user[s]] = D;
}

__global__ void __launch_bounds__(128) k_dg_op(DgArgs A)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.ops.n) return;
    u64 D = 0;
    if (A.nrep[i]) {
        const u32 kind = A.ops.kind[i];
        if (kind == NUTSB_OP_USER) {
            const u32 e = A.op_ev[i];
            if (e != 0xffffffffu && A.ev_len[e]) D = A.ev_d[e] ^ (u64)A.ev_len[e];
        } else if (A.op_g[i] != 0xffffffffu) {
            const u32 g = A.op_g[i];
            const u64 l_on = A.cpx.vp_on[g + 1] - A.cpx.vp_on[g], l_off = A.cpx.vp_off[g + 1] - A.cpx.vp_off[g];
            const u64 d_on = A.dg_on[g], d_off = A.dg_off[g];
            const i32 tgt = A.ops.target[i], exc = A.ops.except_user[i];
            const u32 of = A.ops.flags[i];
            const bool one_room = kind == NUTSB_OP_ROOM && tgt >= 0;
            const i32 n0 = one_room ? A.room_users_off[tgt] : 0, n1 = one_room ? A.room_users_off[tgt + 1] : A.pop.n_users;
            for (i32 x = n0; x < n1; ++x) {                           // recipients in user-list order (c:1409)
                const i32 u = one_room ? A.room_users[x] : x;
                if (u == exc) continue;
                if (kind == NUTSB_OP_ROOM && A.pop.user_room[u] >= A.pop.n_rooms) continue;      // u->room == NULL: c:1410
                const i32 s = A.pop.user_slot[u];
                const u32 cf = A.pop.slot_cf[s];
                if (!nutsb_class_delivers(cf, A.pop.slot_lv[s], kind, of, tgt)) continue;
                const u64 len = (cf & NUTSB_UF_COLOUR) ? l_on : l_off;
                if (len) D = (D * NUTSB_DG_K) ^ ((cf & NUTSB_UF_COLOUR) ? d_on : d_off) ^ len;
            }
        }
    }
    A.per_op[i] = D;
}
