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
