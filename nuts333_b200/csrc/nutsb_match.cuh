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
    // the compact form k_ac_warp walks (null when the automaton does not fit: then k_ac_match runs)
    const u16 *tr16;         // [nstates][ncls]: BYTE offset of the target state's row; match states have the highest rows
    const u8  *cls2;         // [256] 2 * class
    u32 thresh;              // a row offset >= thresh is a match state's
    u32 maxpat;              // longest pattern, bytes
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
// one as well).  The state falls back to the root at every string start; a hit is credited to the string
// the walk is in.  The automaton is stored for a five-instruction step: u16 entries that hold the BYTE
// offset of the next state's row (no multiply), the class table pre-doubled, match states numbered last so
// that "hit" is a running maximum compared once per string piece.
#define NUTSB_ACW_THREADS 256
#define NUTSB_ACW_WIN     2560                     // text bytes staged per warp (32 strings of 80 bytes), and as many mask bytes
// dynamic shared memory: the table (tr_bytes, a multiple of 16), the class table, the warps' windows and string starts
#define NUTSB_ACW_HITW    ((NUTSB_ACW_WIN + 48) / 32 + 1)                  // words of the per-warp hit bitmap
#define NUTSB_ACW_SMEM(tr_bytes) ((tr_bytes) + 256 + (NUTSB_ACW_THREADS / 32) * (2 * (NUTSB_ACW_WIN + 48) + 34 * 4 + 4 * NUTSB_ACW_HITW))

__device__ __forceinline__ u32 nutsb_acw_walk(const u8 *p, const u8 *e, const u8 *s_trb, const u8 *s_cls2, u32 &st)
{
    u32 acc = 0;
    for (; p < e; ++p) {
        st = *(const u16 *)(s_trb + st + s_cls2[*p]);
        acc = acc > st ? acc : st;
    }
    return acc;
}

__global__ void __launch_bounds__(NUTSB_ACW_THREADS)
k_ac_warp(const u8 *text, const u64 *off, i64 n, AcView ac, u32 tr_bytes, u8 *verdict)
{
    NUTSB_DYN_SMEM(smem);
    u8 *const s_trb = smem;                                              // the transition table, bytes
    u8 *const s_cls2 = smem + tr_bytes;
    u8 *const s_stage_all = s_cls2 + 256;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u8 *const stage = s_stage_all + warp * 2 * (NUTSB_ACW_WIN + 48);
    u8 *const pm = stage + NUTSB_ACW_WIN + 48;
    u32 *const s_p0 = (u32 *)(s_stage_all + (NUTSB_ACW_THREADS / 32) * 2 * (NUTSB_ACW_WIN + 48)) + warp * (34 + NUTSB_ACW_HITW);
    u32 *const hits = s_p0 + 34;
    {
        const u32 cnt = ac.nstates * ac.ncls;
        for (u32 i = threadIdx.x; i < cnt; i += blockDim.x) ((u16 *)s_trb)[i] = ac.tr16[i];
        for (u32 i = threadIdx.x; i < 256; i += blockDim.x) s_cls2[i] = ac.cls2[i];
    }
    __syncthreads();
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
            if ((u32)lane < nops && o1 > o0 && nutsb_acw_walk(text + o0, text + o1, s_trb, s_cls2, st) >= ac.thresh) mask = 1u << lane;
        } else {
            // the warp's strings in one window, or -- when they add up to more than the window -- in several
            for (u32 qa = 0; qa < nops; ) {
                const u64 b0 = __shfl_sync(NUTSB_FULL, o0, (int)qa);
                const u8 *pa = (const u8 *)((size_t)(text + b0) & ~(size_t)15);
                const bool fit = (u32)lane >= qa && (u32)lane < nops && (u64)((text + o1) - pa) <= NUTSB_ACW_WIN;
                const u32 fm = __ballot_sync(NUTSB_FULL, fit) >> qa;
                const u32 cnt = fm == 0xffffffffu ? 32u : (u32)__ffs((int)~fm) - 1;   // strings qa .. qa+cnt-1 fit together (fit is monotone)
                if (cnt == 0) {                                          // one string longer than the window: its lane walks it
                    u32 st = 0;
                    if ((u32)lane == qa && nutsb_acw_walk(text + o0, text + o1, s_trb, s_cls2, st) >= ac.thresh) mask |= 1u << lane;
                    ++qa;
                    continue;
                }
                const u32 qb = qa + cnt;
                const u64 b1 = __shfl_sync(NUTSB_FULL, o1, (int)qb - 1);
                const u32 span = (u32)((text + b1) - pa);
                const u32 nvec = (span + 15) >> 4;
                for (u32 v = lane; v < nvec; v += 32) {
                    *(uint4 *)(stage + 16 * v) = __ldg((const uint4 *)pa + v);
                    *(uint4 *)(pm + 16 * v) = make_uint4(~0u, ~0u, ~0u, ~0u);
                }
                for (u32 v = lane; v < (NUTSB_ACW_WIN + 48) / 32; v += 32) hits[v] = 0;
                const u32 r0 = (u32)((text + b0) - pa), r1 = span;
                // s_p0[i] = start of string qa + i in the window, i = 0 .. cnt (the end)
                if ((u32)lane >= qa && (u32)lane < qb) s_p0[lane - qa] = (u32)((text + o0) - pa);
                if (lane == 0) s_p0[cnt] = r1;
                __syncwarp();
                // string starts as DATA: pm[j] = 0 where a string begins (the state falls back to the root there), 0xff
                // elsewhere -- a lane meets a string start once in ~64 bytes, but some lane of the warp meets one every
                // other step, so a branch for it would run all the time
                if ((u32)lane < cnt) pm[s_p0[lane]] = 0;
                __syncwarp();
                // a lane's piece is an ODD number of words: the lanes walk in step, byte i of every piece at once, and
                // with an odd word stride those 32 bytes sit in 32 different banks (64-byte pieces would share two)
                const u32 total = r1 - r0;
                const u32 chunk = 4u * ((((total + 31) >> 5) + 3) >> 2 | 1u);
                const u32 c0 = r0 + (u32)lane * chunk;
                u32 c1 = c0 + chunk; if (c1 > r1) c1 = r1;
                if (c0 < r1) {
                    // four bytes per round: the text word and the mask word with one load each, the four class lookups
                    // together, then the four dependent transitions; hits are looked for once per round (a running
                    // maximum) and, when there is one, the round is walked again byte by byte to credit the right string.
                    // The walk starts on a word boundary at or before its first byte and ends on one at or after its last:
                    // the extra bytes only lengthen the overlap (hits outside [r0, r1) are dropped).
                    const u32 js = (c0 >= r0 + ov ? c0 - ov : r0) & ~3u, je = (c1 + 3) & ~3u;
                    u32 st = 0;
                    // (the next round's words and classes are fetched before this round's transitions are walked: they do
                    // not depend on the state, and the kernel is bound by the latency of the four dependent lookups)
                    u32 m = *(const u32 *)(pm + js);
                    u32 k0, k1, k2, k3;
                    { const u32 w = *(const u32 *)(stage + js);
                      k0 = s_cls2[w & 0xffu]; k1 = s_cls2[(w >> 8) & 0xffu]; k2 = s_cls2[(w >> 16) & 0xffu]; k3 = s_cls2[w >> 24]; }
                    for (u32 j = js; j < je; j += 4) {                    // the same trip count in every lane (+-1)
                        const u32 jn = j + 4 < je ? j + 4 : j;
                        const u32 wn = *(const u32 *)(stage + jn), mn = *(const u32 *)(pm + jn);
                        const u32 n0 = s_cls2[wn & 0xffu], n1 = s_cls2[(wn >> 8) & 0xffu], n2 = s_cls2[(wn >> 16) & 0xffu], n3 = s_cls2[wn >> 24];
                        const u32 s1 = *(const u16 *)(s_trb + (st & __byte_perm(m, 0, 0x4400)) + k0);
                        const u32 s2 = *(const u16 *)(s_trb + (s1 & __byte_perm(m, 0, 0x4511)) + k1);
                        const u32 s3 = *(const u16 *)(s_trb + (s2 & __byte_perm(m, 0, 0x4622)) + k2);
                        st = *(const u16 *)(s_trb + (s3 & __byte_perm(m, 0, 0x4733)) + k3);
                        m = mn; k0 = n0; k1 = n1; k2 = n2; k3 = n3;
                        u32 acc = s1 > s2 ? s1 : s2; acc = acc > s3 ? acc : s3; acc = acc > st ? acc : st;
                        if (acc >= ac.thresh) {
                            // a hit: only its POSITION is noted here (one bit per text byte, one shared-memory atomic per
                            // round) -- whose string it is gets settled after the walk, by the string's own lane
                            const u32 T = ac.thresh;
                            const u32 hb = (s1 >= T ? 1u : 0u) | (s2 >= T ? 2u : 0u) | (s3 >= T ? 4u : 0u) | (st >= T ? 8u : 0u);
                            atomicOr(&hits[j >> 5], hb << (j & 31u));     // (j is a multiple of 4: the four bits stay in one word)
                        }
                    }
                }
                __syncwarp();
                if ((u32)lane < cnt) {                                    // lane i: any hit inside string qa + i ?
                    const u32 a0 = s_p0[lane], a1 = s_p0[lane + 1];
                    u32 any = 0;
                    for (u32 wd = a0 >> 5; a1 > a0 && wd <= (a1 - 1) >> 5; ++wd) {
                        u32 bits = hits[wd];
                        if (wd == a0 >> 5) bits &= 0xffffffffu << (a0 & 31u);
                        if (wd == (a1 - 1) >> 5) bits &= 0xffffffffu >> (31u - ((a1 - 1) & 31u));
                        any |= bits;
                    }
                    if (any) mask |= 1u << (qa + (u32)lane);
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
