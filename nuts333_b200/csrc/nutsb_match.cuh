// nutsb_match.cuh -- contains_swearing / site_banned / user_banned kernels.
//
//   contains_swearing (nuts333.c:2540-2559): "some list word is a substring of
//   the lower-cased string"  -> Aho-Corasick automaton over the word list, byte
//   classes fold A-Z onto a-z (tolower in the C locale, c:2657).
//   site_banned (c:330-345): "some tested token of datafiles/siteban is a
//   substring of the site" -> the same automaton, no case folding.
//   user_banned (c:349-364): "some tested token equals the name" -> hash set.
//
// One thread per string, persistent blocks.  The transition table lives in shared
// memory when it fits (the 64-word swear list: ~22 KB), else it is read through
// L1/L2 (the 10k-entry site list: ~20 MB, L2-resident).
#pragma once
#include "nutsb_common.cuh"

struct AcView {
    const u32 *trans;        // [nstates][ncls]; bit31 = target state is a match state
    const u8  *clsmap;       // [256] byte -> class (0 = byte absent from every pattern)
    u32 ncls, nstates;
    u32 root_match;          // an empty pattern: every string matches (strstr(s,"") != NULL)
    // the one-byte compact form k_ac_pair<false> walks (null when the automaton does not fit: then k_ac_match runs)
    const u16 *tr16;         // [nstates][ncls]: BYTE offset of the target state's row; match states have the highest rows
    const u8  *cls2;         // [256] 2 * class
    u32 thresh;              // a row offset >= thresh is a match state's
    u32 maxpat;              // longest pattern, bytes
    // the two-bytes-a-step form k_ac_pair walks (null when it does not fit 16 KB: short lists only)
    const u16 *t2;           // [nstates][t2_rs / 2]: entry (a, b) = row offset of delta(delta(s, a), b) | bit14: delta(s, a) is a match state | bit15: the target is
    const u16 *clsA;         // [256] 2 * ncls * class  (the pair's first byte)
    const u8  *clsB;         // [256] 2 * class         (its second)
    u32 t2_rs;               // bytes per row: a power of two >= 2 * ncls * ncls
};

#define NUTSB_AC_THREADS     256
#define NUTSB_AC_SMEM_ENTRIES 12288      // u16 entries (24 KB)

// One automaton step per byte: class lookup, then transition.
template <bool SMEM_DFA>
__device__ __forceinline__ bool nutsb_ac_step(u32 &st, u32 c, const u16 *s_tr, const AcView &ac)
{
    if (SMEM_DFA) {
        const u32 t = s_tr[st * ac.ncls + c];
        st = t & 0x7fffu;
        return (t & 0x8000u) != 0;
    } else {
        const u32 t = __ldg(ac.trans + (size_t)st * ac.ncls + c);
        st = t & 0x7fffffffu;
        return (t >> 31) != 0;
    }
}

// Persistent blocks: the table is loaded into shared memory once per block, then
// the block strides over the strings, one thread per string.  Bytes are fetched
// four at a time with aligned 32-bit loads, the next word requested before the
// current one is walked (the four class lookups of a word are independent; only
// the four transitions form a dependent chain).
template <bool SMEM_DFA>
__global__ void __launch_bounds__(NUTSB_AC_THREADS)
k_ac_match(const u8 *text, const u64 *off, i64 n, AcView ac, u8 *verdict)
{
    __shared__ u8 s_cls[256];
    __shared__ u16 s_tr[SMEM_DFA ? NUTSB_AC_SMEM_ENTRIES : 1];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_cls[i] = ac.clsmap[i];
    if (SMEM_DFA) {
        const u32 cnt = ac.nstates * ac.ncls;
        for (u32 i = threadIdx.x; i < cnt; i += blockDim.x) {
            const u32 t = ac.trans[i];
            s_tr[i] = (u16)((t & 0x7fffu) | ((t >> 31) << 15));
        }
    }
    __syncthreads();
    for (i64 i = (i64)blockIdx.x * NUTSB_AC_THREADS + threadIdx.x; i < n; i += (i64)gridDim.x * NUTSB_AC_THREADS) {
        const u64 o0 = off[i], o1 = off[i + 1];
        bool hit = ac.root_match != 0;
        if (!hit && o1 > o0) {
            const u8 *s = text + o0;
            const u32 len = (u32)(o1 - o0);
            u32 st = 0, j = 0;
            while (j < len && (((size_t)(s + j)) & 3) != 0 && !hit) {          // up to the first aligned word
                hit = nutsb_ac_step<SMEM_DFA>(st, s_cls[__ldg(s + j)], s_tr, ac); ++j;
            }
            const u32 nwords = hit ? 0 : (len - j) >> 2;
            const u32 *wp = (const u32 *)(s + j);
            u32 w = nwords ? __ldg(wp) : 0;
            u32 k = 0;
            for (; k < nwords && !hit; ++k) {
                const u32 wn = (k + 1 < nwords) ? __ldg(wp + k + 1) : 0;
                const u32 c0 = s_cls[w & 0xff], c1 = s_cls[(w >> 8) & 0xff], c2 = s_cls[(w >> 16) & 0xff], c3 = s_cls[w >> 24];
                hit = nutsb_ac_step<SMEM_DFA>(st, c0, s_tr, ac) || nutsb_ac_step<SMEM_DFA>(st, c1, s_tr, ac) ||
                      nutsb_ac_step<SMEM_DFA>(st, c2, s_tr, ac) || nutsb_ac_step<SMEM_DFA>(st, c3, s_tr, ac);
                w = wn;
            }
            j += 4 * nwords;
            while (j < len && !hit) { hit = nutsb_ac_step<SMEM_DFA>(st, s_cls[__ldg(s + j)], s_tr, ac); ++j; }
        }
        verdict[i] = hit ? 1 : 0;
    }
}

// ---- the warp-cooperative form --------------------------------------------------------------
// One thread per string leaves half the lanes idle (the longest string of a warp sets the pace) and reads
// the text with per-thread 4-byte loads.  Here a warp takes 32 consecutive strings -- one contiguous piece
// of the packed text -- stages it into shared memory with coalesced 16-byte loads and splits it into 32
// equal pieces BY BYTES, whatever the string lengths.  A lane starts its walk (maxpat - 1) bytes before its
// piece, in the root state: any match that ends inside the piece lies wholly in what the lane has walked,
// so every match is found by the lane whose piece it ends in (and a match found in the overlap is a true
// one as well).  The automaton is stored for a short step: u16 entries that hold the BYTE offset of the
// next state's row (no multiply), the class table pre-doubled, match states numbered last so that "hit" is
// one comparison of a running maximum per round.  A hit only sets the bit of its text position (one
// shared-memory atomic); whose string it is gets settled after the walk by the string's own lane.
#define NUTSB_ACW_THREADS 256
#define NUTSB_ACW_WIN     2560                     // text bytes staged per warp (32 strings of 80 bytes)
#define NUTSB_ACW_HITW    ((NUTSB_ACW_WIN + 48) / 32 + 1)                  // words of the per-warp hit bitmap

__device__ __forceinline__ u32 nutsb_acw_walk(const u8 *p, const u8 *e, const u8 *s_trb, const u8 *s_cls2, u32 &st)
{
    u32 acc = 0;
    for (; p < e; ++p) {
        st = *(const u16 *)(s_trb + st + s_cls2[*p]);
        acc = acc > st ? acc : st;
    }
    return acc;
}

// ---- no reset at string starts; two bytes a step ----------------------------------------------
// The walk's time goes into the chain of dependent table lookups (one per byte); a mask that reset the state at
// every string start (round 2's first form of this kernel) cost a second staged array and a third of the
// instructions on top.  Both are cut down:
//   * for a short word list the table is squared -- entry (s, a, b) is the state after the two bytes a b, with two
//     flag bits for "the state in between is a match state" and "the target is": half the dependent lookups, 16 KB
//     at most (PAIR; the 64-word list of config 3 does not fit and walks the one-byte table);
//   * the state is NOT reset at string starts.  A match found that way is a true occurrence in the packed text; it
//     belongs to string i iff it also STARTS at or after the string's first byte, and a pattern is at most maxpat
//     bytes, so a hit at least maxpat - 1 bytes into the string is good as it stands.  Only a hit inside a string's
//     first maxpat - 1 bytes can reach back into the previous string: the string's own lane settles that after the
//     walk by walking those few bytes again from the root (rare: the packed text has to break inside a word).
#define NUTSB_ACP_SMEM(t2_bytes) ((t2_bytes) + 512 + 256 + (NUTSB_ACW_THREADS / 32) * ((NUTSB_ACW_WIN + 48) + 34 * 4 + 4 * NUTSB_ACW_HITW))

__device__ __forceinline__ u32 nutsb_bits_any(const u32 *hits, u32 a0, u32 a1)              // any bit in [a0, a1) ?
{
    u32 any = 0;
    for (u32 wd = a0 >> 5; a1 > a0 && wd <= (a1 - 1) >> 5; ++wd) {
        u32 bits = hits[wd];
        if (wd == a0 >> 5) bits &= 0xffffffffu << (a0 & 31u);
        if (wd == (a1 - 1) >> 5) bits &= 0xffffffffu >> (31u - ((a1 - 1) & 31u));
        any |= bits;
    }
    return any;
}

// PAIR: the squared table; !PAIR: the one-byte table (tr16) walked the same way (no string-start mask, half the
// shared memory), for lists whose squared table does not fit.
template <bool PAIR>
__global__ void __launch_bounds__(NUTSB_ACW_THREADS)
k_ac_pair(const u8 *text, const u64 *off, i64 n, AcView ac, u32 t2_bytes, u8 *verdict)
{
    NUTSB_DYN_SMEM(smem);
    u8 *const s_t2 = smem;                                               // (!PAIR: the one-byte table, t2_bytes its size)
    u16 *const s_clsA = (u16 *)(smem + t2_bytes);
    u8 *const s_clsB = smem + t2_bytes + 512;
    u8 *const s_cls2 = smem + t2_bytes;                                  // (!PAIR)
    u8 *const s_stage_all = s_clsB + 256;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u8 *const stage = s_stage_all + warp * (NUTSB_ACW_WIN + 48);
    u32 *const s_p0 = (u32 *)(s_stage_all + (NUTSB_ACW_THREADS / 32) * (NUTSB_ACW_WIN + 48)) + warp * (34 + NUTSB_ACW_HITW);
    u32 *const hits = s_p0 + 34;
    if (PAIR) {
        const u32 cnt = ac.nstates * (ac.t2_rs >> 1);
        for (u32 i = threadIdx.x; i < cnt; i += blockDim.x) ((u16 *)s_t2)[i] = ac.t2[i];
        for (u32 i = threadIdx.x; i < 256; i += blockDim.x) { s_clsA[i] = ac.clsA[i]; s_clsB[i] = ac.clsB[i]; }
    } else {
        const u32 cnt = ac.nstates * ac.ncls;
        for (u32 i = threadIdx.x; i < cnt; i += blockDim.x) ((u16 *)s_t2)[i] = ac.tr16[i];
        for (u32 i = threadIdx.x; i < 256; i += blockDim.x) s_cls2[i] = ac.cls2[i];
    }
    __syncthreads();
    const u8 *const g_trb = (const u8 *)ac.tr16;                          // the one-byte table, for the rare second looks
    const u32 ov = ac.maxpat ? ac.maxpat - 1 : 0;
    const i64 ntask = (n + 31) >> 5;
    for (i64 task = (i64)blockIdx.x * (NUTSB_ACW_THREADS / 32) + warp; task < ntask; task += (i64)gridDim.x * (NUTSB_ACW_THREADS / 32)) {
        const i64 wbase = task << 5;
        const u32 nops = (u32)(n - wbase < 32 ? n - wbase : 32);
        u64 o0 = 0, o1 = 0;
        if ((u32)lane < nops) { o0 = off[wbase + lane]; o1 = off[wbase + lane + 1]; }
        const bool bad = __any_sync(NUTSB_FULL, o1 < o0);
        u32 mask = 0;
        if (ac.root_match) mask = 0xffffffffu;                           // strstr(s, "") != NULL
        else if (bad) {                                                  // offsets out of order: each string by itself
            u32 st = 0;
            if ((u32)lane < nops && o1 > o0 && nutsb_acw_walk(text + o0, text + o1, g_trb, ac.cls2, st) >= ac.thresh) mask = 1u << lane;
        } else {
            for (u32 qa = 0; qa < nops; ) {
                const u64 b0 = __shfl_sync(NUTSB_FULL, o0, (int)qa);
                const u8 *pa = (const u8 *)((size_t)(text + b0) & ~(size_t)15);
                const bool fit = (u32)lane >= qa && (u32)lane < nops && (u64)((text + o1) - pa) <= NUTSB_ACW_WIN;
                const u32 fm = __ballot_sync(NUTSB_FULL, fit) >> qa;
                const u32 cnt = fm == 0xffffffffu ? 32u : (u32)__ffs((int)~fm) - 1;
                if (cnt == 0) {                                          // one string longer than the window: its lane walks it
                    u32 st = 0;
                    if ((u32)lane == qa && nutsb_acw_walk(text + o0, text + o1, g_trb, ac.cls2, st) >= ac.thresh) mask |= 1u << lane;
                    ++qa;
                    continue;
                }
                const u32 qb = qa + cnt;
                const u64 b1 = __shfl_sync(NUTSB_FULL, o1, (int)qb - 1);
                const u32 span = (u32)((text + b1) - pa);
                const u32 nvec = (span + 15) >> 4;
                for (u32 v = lane; v < nvec; v += 32) *(uint4 *)(stage + 16 * v) = __ldg((const uint4 *)pa + v);
                for (u32 v = lane; v < (NUTSB_ACW_WIN + 48) / 32; v += 32) hits[v] = 0;
                const u32 r0 = (u32)((text + b0) - pa), r1 = span;
                if ((u32)lane >= qa && (u32)lane < qb) s_p0[lane - qa] = (u32)((text + o0) - pa);
                if (lane == 0) s_p0[cnt] = r1;
                __syncwarp();
                const u32 total = r1 - r0;
                const u32 chunk = 4u * ((((total + 31) >> 5) + 3) >> 2 | 1u);            // an odd number of words (banks)
                const u32 c0 = r0 + (u32)lane * chunk;
                u32 c1 = c0 + chunk; if (c1 > r1) c1 = r1;
                if (c0 < r1) {
                    // a word a round: its two column offsets are ready before the round starts (they do not depend on the
                    // state), the round itself is two dependent lookups.  Bytes walked outside [r0, r1) -- word alignment --
                    // only set hit bits nobody looks at, or lengthen the overlap.
                    const u32 js = (c0 >= r0 + ov ? c0 - ov : r0) & ~3u, je = (c1 + 3) & ~3u;
                    u32 st = 0, k01, k23;
                    if (!PAIR) {
                        u32 k0, k1, k2, k3;
                        { const u32 w = *(const u32 *)(stage + js);
                          k0 = s_cls2[w & 0xffu]; k1 = s_cls2[(w >> 8) & 0xffu]; k2 = s_cls2[(w >> 16) & 0xffu]; k3 = s_cls2[w >> 24]; }
                        const u32 T = ac.thresh;
                        for (u32 j = js; j < je; j += 4) {
                            const u32 jn = j + 4 < je ? j + 4 : j;
                            const u32 wn = *(const u32 *)(stage + jn);
                            const u32 n0 = s_cls2[wn & 0xffu], n1 = s_cls2[(wn >> 8) & 0xffu], n2 = s_cls2[(wn >> 16) & 0xffu], n3 = s_cls2[wn >> 24];
                            const u32 s1 = *(const u16 *)(s_t2 + st + k0);
                            const u32 s2 = *(const u16 *)(s_t2 + s1 + k1);
                            const u32 s3 = *(const u16 *)(s_t2 + s2 + k2);
                            st = *(const u16 *)(s_t2 + s3 + k3);
                            k0 = n0; k1 = n1; k2 = n2; k3 = n3;
                            u32 acc = s1 > s2 ? s1 : s2; acc = acc > s3 ? acc : s3; acc = acc > st ? acc : st;
                            if (acc >= T)
                                atomicOr(&hits[j >> 5], ((s1 >= T ? 1u : 0u) | (s2 >= T ? 2u : 0u) | (s3 >= T ? 4u : 0u) | (st >= T ? 8u : 0u)) << (j & 31u));
                        }
                    } else {
                    { const u32 w = *(const u32 *)(stage + js);
                      k01 = s_clsA[w & 0xffu] + s_clsB[(w >> 8) & 0xffu]; k23 = s_clsA[(w >> 16) & 0xffu] + s_clsB[w >> 24]; }
                    for (u32 j = js; j < je; j += 4) {
                        const u32 jn = j + 4 < je ? j + 4 : j;
                        const u32 wn = *(const u32 *)(stage + jn);
                        const u32 n01 = s_clsA[wn & 0xffu] + s_clsB[(wn >> 8) & 0xffu], n23 = s_clsA[(wn >> 16) & 0xffu] + s_clsB[wn >> 24];
                        const u32 s1 = *(const u16 *)(s_t2 + ((st & 0x3fffu) | k01));
                        st = *(const u16 *)(s_t2 + ((s1 & 0x3fffu) | k23));
                        k01 = n01; k23 = n23;
                        if ((s1 | st) >= 0x4000u)                            // bits 14/15 of the two entries = hits at bytes j .. j+3
                            atomicOr(&hits[j >> 5], ((s1 >> 14) | ((st >> 14) << 2)) << (j & 31u));
                    }
                    }
                }
                __syncwarp();
                if ((u32)lane < cnt) {                                    // lane i: string qa + i
                    const u32 a0 = s_p0[lane], a1 = s_p0[lane + 1];
                    const u32 am = a0 + ov < a1 ? a0 + ov : a1;           // hits from here on lie wholly inside the string
                    bool hit = nutsb_bits_any(hits, am, a1) != 0;
                    if (!hit && nutsb_bits_any(hits, a0, am)) {           // a hit in the first maxpat - 1 bytes: from inside, or from the string before?
                        u32 st = 0;
                        hit = nutsb_acw_walk(stage + a0, stage + am, g_trb, ac.cls2, st) >= ac.thresh;
                    }
                    if (hit) mask |= 1u << (qa + (u32)lane);
                }
                __syncwarp();
                qa = qb;
            }
        }
        mask = __reduce_or_sync(NUTSB_FULL, mask);
        if ((u32)lane < nops) verdict[wbase + lane] = (u8)((mask >> lane) & 1u);
    }
}

// Exact-match set: open addressing, 32-bit FNV-1a, linear probing.
struct SetView {
    const u32 *slot_off;     // [nslots] offset into pool, 0xffffffff = empty
    const u32 *slot_len;     // [nslots]
    const u8  *pool;
    u32 mask;                // nslots - 1 (power of two), 0 when the set is empty
    u32 count;
};

__device__ __forceinline__ u32 nutsb_fnv32(const u8 *p, u32 n)
{
    u32 h = 0x811c9dc5u;
    for (u32 i = 0; i < n; ++i) { h ^= p[i]; h *= 0x01000193u; }
    return h;
}

__global__ void __launch_bounds__(256)
k_set_match(const u8 *text, const u64 *off, i64 n, SetView set, u8 *verdict)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 hit = 0;
    if (set.count) {
        const u8 *s = text + off[i];
        const u32 len = (u32)(off[i + 1] - off[i]);
        u32 h = nutsb_fnv32(s, len) & set.mask;
        for (;;) {
            const u32 so = set.slot_off[h];
            if (so == 0xffffffffu) break;
            if (set.slot_len[h] == len) {
                u32 j = 0;
                while (j < len && set.pool[so + j] == s[j]) ++j;
                if (j == len) { hit = 1; break; }
            }
            h = (h + 1) & set.mask;
        }
    }
    verdict[i] = (u8)hit;
}

// ---- colour_com_count / colour_com_strip (nuts333.c:2563-2610) -----------------------------
// Small text utilities of the reference: `.who` pads its columns by the number of colour
// commands in a line (c:4846), peers older than 3.2 get lines with the commands removed
// (c:1301).  One thread per string; the strings are short.
//   count: after a hit the scan pointer moves on ONE byte and the loop over the remaining
//          table entries goes on from there (the `continue` belongs to the for, c:2575), so
//          "~FBK" counts 2.  strip: "~XX" removed for known XX, everything else kept; no '/~'
//          escape, no CR.
__global__ void __launch_bounds__(256)
k_colour_com_count(const u8 *text, const u64 *off, i64 n, const u8 *codetab, i32 *count, u32 *strip_len)
{
    __shared__ u8 s_tab[NUTSB_CODETAB_BYTES];
    for (int i = threadIdx.x; i < NUTSB_CODETAB_BYTES; i += blockDim.x) s_tab[i] = codetab[i];
    __syncthreads();
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u8 *s = text + off[i];
    const u32 len = (u32)(off[i + 1] - off[i]);
    i32 cnt = 0; u32 kept = 0;
    // strip length: the plain left-to-right rule
    for (u32 p = 0; p < len;) {
        if (s[p] == '~' && p + 2 < len && nutsb_code(s_tab, s[p + 1], s[p + 2]) >= 0) p += 3; else { ++kept; ++p; }
    }
    // count: the table is walked in its own order k = 0..20 with a moving pointer
    for (u32 p = 0; p < len;) {
        if (s[p] != '~') { ++p; continue; }
        ++p;
        int k = 0;
        while (k < 21 && p + 1 < len) {
            const int hit = nutsb_code(s_tab, s[p], s[p + 1]);     // index of the pair in the table, or -1
            if (hit >= k) { ++cnt; ++p; k = hit + 1; } else break;   // only entries not yet passed can still match
        }
    }
    if (count) count[i] = cnt;
    if (strip_len) strip_len[i] = kept;
}

__global__ void __launch_bounds__(256)
k_colour_com_strip(const u8 *text, const u64 *off, i64 n, const u8 *codetab, const u64 *out_off, u8 *out)
{
    __shared__ u8 s_tab[NUTSB_CODETAB_BYTES];
    for (int i = threadIdx.x; i < NUTSB_CODETAB_BYTES; i += blockDim.x) s_tab[i] = codetab[i];
    __syncthreads();
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u8 *s = text + off[i];
    const u32 len = (u32)(off[i + 1] - off[i]);
    u8 *d = out + out_off[i];
    for (u32 p = 0; p < len;) {
        if (s[p] == '~' && p + 2 < len && nutsb_code(s_tab, s[p + 1], s[p + 2]) >= 0) p += 3; else *d++ = s[p++];
    }
}
