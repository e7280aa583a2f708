/* nutsb_gen.c -- synthetic input generator, host C.
 *
 * Shared by the oracle legs and the GPU driver: every batch is generated once on
 * the host and handed identically to both.  PRNG = splitmix64; entity i of kind
 * k under seed S has its own stream seeded S ^ (k<<56) ^ i*0x9E3779B97F4A7C15,
 * so generation is order-independent and shardable by room/rank.
 *
 * Built by gcc into nuts333_b200/_lib/libnutsgen.so.  Pure data generation: no
 * rendering, matching or fan-out logic lives here.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdio.h>

#define API __attribute__((visibility("default")))

typedef struct { uint64_t s; } rng_t;
static uint64_t rng_next(rng_t *r)
{
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static rng_t rng_for(uint64_t seed, uint64_t kind, uint64_t i)
{
    rng_t r; r.s = seed ^ (kind << 56) ^ (i * 0x9E3779B97F4A7C15ull);
    rng_next(&r);
    return r;
}
static uint32_t rng_below(rng_t *r, uint32_t n) { return (uint32_t)((rng_next(r) >> 11) % (n ? n : 1)); }
static int rng_chance(rng_t *r, uint32_t num, uint32_t den) { return rng_below(r, den) < num; }

enum { K_USER = 1, K_BODY = 2, K_MSG = 3, K_SWEAR = 4, K_SITE = 5, K_BANSITE = 6, K_BANUSER = 7, K_QNAME = 8 };

static const char *g_codes[21] = { "RS","OL","UL","LI","RV","FK","FR","FG","FY","FB","FM","FT","FW",
                                   "BK","BR","BG","BY","BB","BM","BT","BW" };

/* name = 'U' + 4 lower-case base-26 digits of u (letters only, first upper) */
API size_t nutsgen_name(int32_t u, char *out)
{
    uint32_t v = (uint32_t)u;
    out[0] = 'U';
    for (int d = 4; d >= 1; --d) { out[d] = (char)('a' + v % 26); v /= 26; }
    out[5] = 0;
    return 5;
}

/* users: room = u / users_per_room (users_per_room <= 0: one room); colour bit
 * random; level 1 (USER); stress: 2 % ignall, 1 % login, 1 % ignshout. */
API void nutsgen_users(uint64_t seed, int32_t u0, int32_t n, int32_t users_per_room, int stress,
                       int32_t *room, uint8_t *flags, uint8_t *level)
{
    for (int32_t k = 0; k < n; ++k) {
        int32_t u = u0 + k;
        rng_t r = rng_for(seed, K_USER, (uint64_t)u);
        uint8_t f = (uint8_t)(rng_next(&r) & 1);
        if (stress) {
            uint32_t x = rng_below(&r, 100);
            if (x < 2) f |= 4; else if (x < 3) f |= 2; else if (x < 4) f |= 8;
        }
        room[k] = users_per_room > 0 ? k / users_per_room : 0;
        flags[k] = f;
        level[k] = stress ? (uint8_t)rng_below(&r, 5) : 1;
    }
}

/* swear list: fuck shit cunt + (n-3) distinct random lower-case words of 4-8 letters.
 * out is n rows of 16 bytes (NUL-terminated). */
API void nutsgen_swear_words(uint64_t seed, int n, char *out)
{
    static const char *stock[3] = { "fuck", "shit", "cunt" };
    for (int i = 0; i < n; ++i) {
        char *w = out + 16 * i;
        if (i < 3) { strcpy(w, stock[i]); continue; }
        for (uint64_t attempt = 0;; ++attempt) {
            rng_t r = rng_for(seed, K_SWEAR, (uint64_t)i + (attempt << 32));
            int len = 4 + (int)rng_below(&r, 5);
            for (int j = 0; j < len; ++j) w[j] = (char)('a' + rng_below(&r, 26));
            w[len] = 0;
            int dup = 0;
            for (int q = 0; q < i && !dup; ++q) dup = strcmp(out + 16 * q, w) == 0;
            if (!dup) break;
        }
    }
}

/* One chat body: 3-16 words of 1-8 lower-case letters (5 % fully upper-case),
 * single spaces; before a word with p=1/8 a token: 60 % one of ~OL ~FR ~RS,
 * 30 % uniform over the 21 codes, 4 % escaped /~XX, 3 % bare ~, 3 % ~ + two
 * random capitals; ending ?/!/none with p 1/4,1/4,1/2.  Swear injection (when
 * n_swear>0): p=1/32 a word is replaced by a list word in random case, p=1/64 a
 * near-miss is inserted (list word split by ~RS, or with one letter changed).
 * Length <= 200 bytes.  Returns the length. */
static size_t gen_body(uint64_t seed, uint64_t m, const char *swear, int n_swear, uint8_t *out)
{
    rng_t r = rng_for(seed, K_BODY, m);
    size_t o = 0;
    int words = 3 + (int)rng_below(&r, 14);
    for (int w = 0; w < words && o < 170; ++w) {
        if (w) out[o++] = ' ';
        if (rng_chance(&r, 1, 8)) {
            uint32_t x = rng_below(&r, 100);
            if (x < 60) { static const char *c3[3] = { "OL", "FR", "RS" }; const char *c = c3[rng_below(&r, 3)]; out[o++] = '~'; out[o++] = c[0]; out[o++] = c[1]; }
            else if (x < 90) { const char *c = g_codes[rng_below(&r, 21)]; out[o++] = '~'; out[o++] = c[0]; out[o++] = c[1]; }
            else if (x < 94) { const char *c = g_codes[rng_below(&r, 21)]; out[o++] = '/'; out[o++] = '~'; out[o++] = c[0]; out[o++] = c[1]; }
            else if (x < 97) { out[o++] = '~'; }
            else { out[o++] = '~'; out[o++] = (uint8_t)('A' + rng_below(&r, 26)); out[o++] = (uint8_t)('A' + rng_below(&r, 26)); }
        }
        int inj = n_swear > 0 ? (int)rng_below(&r, 64) : 99;
        if (inj < 2) {                                   /* p = 1/32: a list word, random case */
            const char *sw = swear + 16 * rng_below(&r, (uint32_t)n_swear);
            for (size_t j = 0; sw[j]; ++j) out[o++] = (uint8_t)((rng_next(&r) & 1) ? sw[j] - 32 : sw[j]);
        } else if (inj == 2) {                           /* p = 1/64: a near miss */
            const char *sw = swear + 16 * rng_below(&r, (uint32_t)n_swear);
            size_t L = strlen(sw);
            if (rng_next(&r) & 1) {
                size_t cut = 1 + rng_below(&r, (uint32_t)(L - 1));
                for (size_t j = 0; j < L; ++j) { if (j == cut) { out[o++] = '~'; out[o++] = 'R'; out[o++] = 'S'; } out[o++] = (uint8_t)sw[j]; }
            } else {
                size_t pos = rng_below(&r, (uint32_t)L);
                for (size_t j = 0; j < L; ++j) out[o++] = (uint8_t)(j == pos ? (sw[j] == 'z' ? 'y' : sw[j] + 1) : sw[j]);
            }
        } else {
            int len = 1 + (int)rng_below(&r, 8);
            int upper = rng_chance(&r, 1, 20);
            for (int j = 0; j < len; ++j) out[o++] = (uint8_t)((upper ? 'A' : 'a') + rng_below(&r, 26));
        }
    }
    uint32_t e = rng_below(&r, 4);
    if (e == 0) out[o++] = '?'; else if (e == 1) out[o++] = '!';
    return o;
}

/* Bodies m0..m0+n-1 packed into bytes (cap >= 208*n), off[n+1] (off[0] = base). */
API int64_t nutsgen_bodies(uint64_t seed, int64_t m0, int64_t n, const char *swear, int n_swear,
                           uint8_t *bytes, uint64_t *off)
{
    uint64_t o = 0;
    off[0] = 0;
    for (int64_t k = 0; k < n; ++k) { o += gen_body(seed, (uint64_t)(m0 + k), swear, n_swear, bytes + o); off[k + 1] = o; }
    return (int64_t)o;
}

/* say() batch (nuts333.c:4062-4100).  Message m: room uniform, speaker uniform in
 * the room (users are laid out room-major, users_per_room each; <= 0: one room
 * of n_users).  mode 0: one op per message, write_room_except(room, "<Name>
 * <type>s: <body>\n", speaker).  mode 1 (ban_swearing): three ops per message,
 * gated on the body's swear verdict m:  write_user(speaker, noswearing) if
 * dirty; write_user(speaker, "You <type>: <body>\n") and the room line if clean.
 * Arrays must hold n*(mode?3:1) ops; text cap >= n * (mode ? 520 : 240).
 * Also returns each message's speaker and room. */
API int64_t nutsgen_say_ops(uint64_t seed, int64_t m0, int64_t n, int32_t n_users, int32_t users_per_room, int mode,
                            const uint8_t *bodies, const uint64_t *body_off,
                            uint8_t *text, uint64_t *text_off, uint8_t *kind, int32_t *target,
                            int32_t *except_user, uint8_t *flags, int32_t *gate,
                            int32_t *speaker, int32_t *msg_room)
{
    static const char noswearing[] = "Swearing is not allowed here.\n";      /* nuts333.h:151 */
    const int32_t upr = users_per_room > 0 ? users_per_room : n_users;
    const int32_t rooms = (n_users + upr - 1) / upr;
    uint64_t o = 0; int64_t q = 0;
    text_off[0] = 0;
    for (int64_t k = 0; k < n; ++k) {
        rng_t r = rng_for(seed, K_MSG, (uint64_t)(m0 + k));
        int32_t room = (int32_t)rng_below(&r, (uint32_t)rooms);
        int32_t first = room * upr, cnt = (first + upr <= n_users) ? upr : n_users - first;
        int32_t spk = first + (int32_t)rng_below(&r, (uint32_t)cnt);
        const uint8_t *b = bodies + body_off[k]; size_t bl = (size_t)(body_off[k + 1] - body_off[k]);
        const char *type = "say";                                          /* c:4080-4084 */
        if (bl && b[bl - 1] == '?') type = "ask"; else if (bl && b[bl - 1] == '!') type = "exclaim";
        char name[8]; nutsgen_name(spk, name);
        if (speaker) speaker[k] = spk;
        if (msg_room) msg_room[k] = room;
        if (mode) {
            memcpy(text + o, noswearing, sizeof noswearing - 1); o += sizeof noswearing - 1;
            kind[q] = 0; target[q] = spk; except_user[q] = -1; flags[q] = 8; gate[q] = (int32_t)k; text_off[++q] = o;
            o += (uint64_t)sprintf((char *)text + o, "You %s: ", type);
            memcpy(text + o, b, bl); o += bl; text[o++] = '\n';
            kind[q] = 0; target[q] = spk; except_user[q] = -1; flags[q] = 0; gate[q] = (int32_t)k; text_off[++q] = o;
        }
        o += (uint64_t)sprintf((char *)text + o, "%s %ss: ", name, type);
        memcpy(text + o, b, bl); o += bl; text[o++] = '\n';
        kind[q] = 1; target[q] = room; except_user[q] = spk; flags[q] = 0; gate[q] = mode ? (int32_t)k : -1; text_off[++q] = o;
    }
    return q;
}

/* Site strings: "h<a>.pool<b>.isp<c>.<tld>" (80 %) or a dotted quad (20 %). */
static size_t gen_site(uint64_t seed, uint64_t kind, uint64_t i, char *out)
{
    static const char *tld[4] = { "com", "org", "net", "edu" };
    rng_t r = rng_for(seed, kind, i);
    if (rng_chance(&r, 1, 5))
        return (size_t)sprintf(out, "%u.%u.%u.%u", 1 + rng_below(&r, 222), rng_below(&r, 256), rng_below(&r, 256), 1 + rng_below(&r, 254));
    return (size_t)sprintf(out, "h%u.pool%u.isp%u.%s", rng_below(&r, 1000), rng_below(&r, 100), rng_below(&r, 500), tld[rng_below(&r, 4)]);
}

API int64_t nutsgen_sites(uint64_t seed, int64_t i0, int64_t n, uint8_t *bytes, uint64_t *off)
{
    uint64_t o = 0; off[0] = 0;
    for (int64_t k = 0; k < n; ++k) { o += gen_site(seed, K_SITE, (uint64_t)(i0 + k), (char *)bytes + o); off[k + 1] = o; }
    return (int64_t)o;
}

/* Names for user_banned queries: the user's own name. */
API int64_t nutsgen_names(int64_t u0, int64_t n, uint8_t *bytes, uint64_t *off)
{
    uint64_t o = 0; off[0] = 0;
    for (int64_t k = 0; k < n; ++k) { o += nutsgen_name((int32_t)(u0 + k), (char *)bytes + o); off[k + 1] = o; }
    return (int64_t)o;
}

/* Ban files, one token per line.  which 0 = siteban: 50 % full host names drawn
 * from the query generator's space (so ~some hit), 30 % domain suffixes
 * ".isp<c>.<tld>", 20 % numeric prefixes "a.b.".  which 1 = userban: names, 5 %
 * drawn from live users [0,n_users), the rest from the same alphabet beyond it.
 * trailing_newline 0 leaves the last token unterminated (never tested, c:339). */
API size_t nutsgen_ban_file(uint64_t seed, int which, int32_t n_entries, int32_t n_users, int64_t n_query_sites,
                            int trailing_newline, uint8_t *out)
{
    static const char *tld[4] = { "com", "org", "net", "edu" };
    size_t o = 0;
    for (int32_t i = 0; i < n_entries; ++i) {
        rng_t r = rng_for(seed, which ? K_BANUSER : K_BANSITE, (uint64_t)i);
        if (!which) {
            uint32_t x = rng_below(&r, 100);
            if (x < 50) {
                /* 1 in 10 is an actual query site, the rest are plausible but absent */
                uint64_t src = rng_chance(&r, 1, 10) ? rng_below(&r, (uint32_t)(n_query_sites ? n_query_sites : 1)) : (1ull << 40) + i;
                o += gen_site(seed, K_SITE, src, (char *)out + o);
            } else if (x < 80) o += (size_t)sprintf((char *)out + o, ".isp%u.%s", 500 + rng_below(&r, 40) - (rng_chance(&r, 1, 8) ? 500 : 0), tld[rng_below(&r, 4)]);
            else o += (size_t)sprintf((char *)out + o, "%u.%u.", 1 + rng_below(&r, 254), rng_below(&r, 256));
        } else {
            int32_t u = rng_chance(&r, 1, 20) ? (int32_t)rng_below(&r, (uint32_t)(n_users ? n_users : 1)) : n_users + (int32_t)rng_below(&r, 300000);
            o += nutsgen_name(u, (char *)out + o);
        }
        if (i + 1 < n_entries || trailing_newline) out[o++] = '\n';
    }
    return o;
}
