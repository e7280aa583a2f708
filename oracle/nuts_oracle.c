/* oracle/nuts_oracle.c -- TEST INFRASTRUCTURE ONLY (see nuts_oracle.h).
 *
 * Plain-C restatement of the NUTS 3.3.3 per-message byte transform and
 * fan-out.  It is the checker the CUDA path is compared with, and the "port"
 * CPU baseline of bench.py; the product never links or loads it.
 *
 * Pinned by tests/test_oracle_golden.py against tests/golden/ (minted from the
 * unmodified reference) and by tests/test_oracle_vs_ref.py against
 * oracle/_ref/libnutsref.so when that was built.
 *
 * c: = /root/reference/nuts333.c, h: = /root/reference/nuts333.h
 */
#include "nuts_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* h:237-255 -- the 21 two-letter commands, in the reference's order.  The
 * ANSI strings are generated instead of listed: entries 0-4 are ESC[<d>m with
 * d = 0,1,4,5,7; entries 5-12 are ESC[3<d>m; entries 13-20 are ESC[4<d>m. */
static const char orc_codes[21][3] = {
    "RS","OL","UL","LI","RV",
    "FK","FR","FG","FY","FB","FM","FT","FW",
    "BK","BR","BG","BY","BB","BM","BT","BW"
};

static size_t orc_ansi(int k, uint8_t *out)
{
    static const char attr[5] = { '0','1','4','5','7' };
    size_t o = 0;
    out[o++] = 0x1b; out[o++] = '[';
    if (k < 5) out[o++] = (uint8_t)attr[k];
    else { out[o++] = (uint8_t)(k < 13 ? '3' : '4'); out[o++] = (uint8_t)('0' + (k - 5) % 8); }
    out[o++] = 'm';
    return o;
}

/* The strncmp(str,colcom[i],2) scan of c:1341-1352: both letters must lie
 * inside the string (strncmp stops at the terminating NUL). */
static int orc_code_at(const uint8_t *s, size_t n, size_t j)
{
    if (j + 1 >= n) return -1;
    for (int k = 0; k < 21; ++k)
        if (s[j] == (uint8_t)orc_codes[k][0] && s[j + 1] == (uint8_t)orc_codes[k][1]) return k;
    return -1;
}

/* c:1291-1366, USER_TYPE recipient.  The reference's 1000-byte staging buffer
 * and its flush points (c:1317,1338,1359) only split the stream across write()
 * calls, so they do not appear here. */
size_t orc_render(const uint8_t *s, size_t n, int colour, uint8_t *out)
{
    return orc_render_ex(s, n, colour, 0, out);
}

/* The same machine as run by write_user (c:1315-1365) and, with ORC_OF_PAGER, by the
 * pager on one fgets() chunk of a file (c:2254-2300: identical loop, no reset after
 * the string); ORC_OF_PLAIN = more(NULL,...): colour is `user!=NULL && user->colour`. */
size_t orc_render_ex(const uint8_t *s, size_t n, int colour, unsigned oflags, uint8_t *out)
{
    size_t i = 0, o = 0;
    if (oflags & ORC_OF_RAW) { memcpy(out, s, n); return n; }      /* write_sock, c:1281: no byte machine */
    if (oflags & ORC_OF_PLAIN) colour = 0;
    while (i < n) {
        uint8_t c = s[i];
        if (c == '\n') {                               /* c:1316-1326 */
            if (colour) o += orc_ansi(0, out + o);
            out[o++] = '\n'; out[o++] = '\r';
            ++i;
        } else if (c == '/' && i + 1 < n && s[i + 1] == '~') {
            ++i;                                       /* c:1330: slash dropped */
        } else if (c == '~' && i > 0 && s[i - 1] == '/') {
            out[o++] = '~'; ++i;                       /* c:1331-1333 */
        } else if (c == '~') {                         /* c:1337-1354 */
            int k = orc_code_at(s, n, i + 1);
            if (k >= 0) { if (colour) o += orc_ansi(k, out + o); i += 3; }
            else { out[o++] = '~'; ++i; }
        } else {
            out[o++] = c; ++i;                         /* c:1355 */
        }
    }
    if (colour && !(oflags & ORC_OF_PAGER)) o += orc_ansi(0, out + o);   /* c:1365 */
    return o;
}

static uint8_t orc_lower(uint8_t b) { return (b >= 'A' && b <= 'Z') ? (uint8_t)(b + 32) : b; }

/* memmem on a lower-cased haystack; needle of length 0 matches (strstr). */
static int orc_has(const uint8_t *h, size_t n, const uint8_t *w, size_t m)
{
    if (m == 0) return 1;
    if (m > n) return 0;
    for (size_t i = 0; i + m <= n; ++i) {
        size_t j = 0;
        while (j < m && orc_lower(h[i + j]) == w[j]) ++j;
        if (j == m) return 1;
    }
    return 0;
}

/* c:2540-2559: lower-case copy (C locale: only A-Z move, c:2657), then strstr
 * per list word until the entry that starts with '*'. */
int orc_contains_swearing(const uint8_t *s, size_t n, const char *const *words)
{
    if (!words) return 0;
    for (size_t w = 0; words[w] && words[w][0] != '*'; ++w)
        if (orc_has(s, n, (const uint8_t *)words[w], strlen(words[w]))) return 1;
    return 0;
}

/* c:2563-2583.  After a hit the code pointer moves on ONE byte and the scan
 * over the remaining table entries carries on from there (the inner
 * 'continue' belongs to the for), hence "~FBK" counts 2. */
int orc_colour_com_count(const uint8_t *s, size_t n)
{
    size_t p = 0; int cnt = 0;
    while (p < n) {
        if (s[p] != '~') { ++p; continue; }
        ++p;
        for (int k = 0; k < 21; ++k)
            if (p + 1 < n && s[p] == (uint8_t)orc_codes[k][0] && s[p + 1] == (uint8_t)orc_codes[k][1]) {
                ++cnt; ++p;
            }
    }
    return cnt;
}

/* c:2588-2610: drops ~XX for known XX, everything else (incl. '/') verbatim. */
size_t orc_colour_com_strip(const uint8_t *s, size_t n, uint8_t *out)
{
    size_t p = 0, o = 0;
    while (p < n) {
        if (s[p] == '~' && orc_code_at(s, n, p + 1) >= 0) p += 3;
        else out[o++] = s[p++];
    }
    return o;
}

static int orc_isspace(uint8_t b)
{
    return b == ' ' || (b >= '\t' && b <= '\r');       /* C-locale isspace */
}

/* c:338-342 / c:357-361: fscanf("%s") then while(!feof): a token is tested only
 * if the scan that produced it stopped on a whitespace byte, i.e. the token
 * does not run into end-of-file. */
size_t orc_ban_tokens(const uint8_t *f, size_t n, uint32_t *tok_off, uint32_t *tok_len, size_t cap)
{
    size_t p = 0, cnt = 0;
    for (;;) {
        while (p < n && orc_isspace(f[p])) ++p;
        if (p >= n) break;                              /* EOF while skipping: feof set */
        size_t b = p;
        while (p < n && !orc_isspace(f[p])) ++p;
        if (p >= n) break;                              /* token hit EOF: never tested */
        if (cnt < cap) { tok_off[cnt] = (uint32_t)b; tok_len[cnt] = (uint32_t)(p - b); }
        ++cnt;
    }
    return cnt;
}

/* A token is handed to strstr/strcmp as a C string: it ends at an embedded NUL. */
static size_t orc_cstr_len(const uint8_t *p, size_t n)
{
    const uint8_t *z = memchr(p, 0, n);
    return z ? (size_t)(z - p) : n;
}

static int orc_ban_scan(const uint8_t *f, size_t n, const uint8_t *q, size_t qn, int substring)
{
    size_t p = 0;
    for (;;) {
        while (p < n && orc_isspace(f[p])) ++p;
        if (p >= n) return 0;
        size_t b = p;
        while (p < n && !orc_isspace(f[p])) ++p;
        if (p >= n) return 0;
        size_t m = orc_cstr_len(f + b, p - b);
        if (substring) {                                /* c:340 strstr(site,line) */
            if (m == 0) return 1;
            for (size_t i = 0; i + m <= qn; ++i)
                if (memcmp(q + i, f + b, m) == 0) return 1;
        } else if (m == qn && memcmp(q, f + b, m) == 0) return 1;   /* c:359 */
    }
}

int orc_site_banned(const uint8_t *f, size_t n, int present, const uint8_t *site, size_t sn)
{
    if (!present) return 0;                             /* c:337 */
    return orc_ban_scan(f, n, site, orc_cstr_len(site, sn), 1);
}

int orc_user_banned(const uint8_t *f, size_t n, int present, const uint8_t *name, size_t nn)
{
    if (!present) return 0;                             /* c:356 */
    return orc_ban_scan(f, n, name, orc_cstr_len(name, nn), 0);
}

/* ---- recipient filter ---------------------------------------------------- */

int orc_delivers(uint8_t kind, int32_t target, int32_t except_user, uint8_t of,
                 int32_t u, int32_t u_room, uint8_t uf, uint8_t u_level)
{
    switch (kind) {
    case ORC_OP_USER:                                   /* c:1298: NULL user = no-op */
        return target >= 0 && u == target;
    case ORC_OP_ROOM:                                   /* c:1410-1415 */
        if (uf & ORC_UF_LOGIN) return 0;
        if (u_room < 0) return 0;
        if (target >= 0 && u_room != target) return 0;
        if ((uf & ORC_UF_IGNALL) && !(of & ORC_OF_FORCE_LISTEN)) return 0;
        if ((uf & ORC_UF_IGNSHOUT) && (of & ORC_OF_SHOUT)) return 0;
        return u != except_user;
    case ORC_OP_LEVEL:                                  /* c:1379-1383 */
        if (u == except_user || (uf & ORC_UF_LOGIN) || (uf & ORC_UF_CLONE)) return 0;
        return (of & ORC_OF_ABOVE) ? (int32_t)u_level >= target : (int32_t)u_level <= target;
    }
    return 0;
}

/* Clones, c:1416-1426: a CLONE_TYPE user never receives anything itself; what write_room_except would
 * have sent it is relayed to its owner as "~FT[ <room name> ]:~RS <str>" -- only for calls that name
 * the clone's own room (rm==NULL reaches the owner anyway), not when clone_hear is NOTHING or the
 * owner ignores everything, and with clone_hear SWEARS only when the string swears.  Rooms are named
 * "room<index>" as in the harness.  State set by orc_set_clones (NULL clears it). */
static const int32_t *orc_remote_link = NULL; static const uint8_t *orc_remote_old = NULL;
static const uint8_t *orc_remote_names = NULL; static const uint64_t *orc_remote_name_off = NULL;
void orc_set_remotes(const int32_t *link, const uint8_t *old, const uint8_t *names, const uint64_t *name_off)
{
    orc_remote_link = link; orc_remote_old = old; orc_remote_names = names; orc_remote_name_off = name_off;
}
/* c:1299-1306: "MSG <name>\n<str>[\n]EMSG\n", str stripped of colour commands for a peer < 3.2 */
static size_t orc_remote_frame(int32_t u, const uint8_t *s, size_t n, uint8_t *out)
{
    size_t o = 0;
    const size_t nl = (size_t)(orc_remote_name_off[u + 1] - orc_remote_name_off[u]);
    memcpy(out, "MSG ", 4); o = 4;
    memcpy(out + o, orc_remote_names + orc_remote_name_off[u], nl); o += nl;
    out[o++] = '\n';
    size_t bl;
    if (orc_remote_old[u]) bl = orc_colour_com_strip(s, n, out + o);
    else { memcpy(out + o, s, n); bl = n; }
    o += bl;
    if (!bl || out[o - 1] != '\n') out[o++] = '\n';
    memcpy(out + o, "EMSG\n", 5); o += 5;
    return o;
}

static const int32_t *orc_clone_owner = NULL; static const uint8_t *orc_clone_hear = NULL;
static const char *const *orc_clone_words = NULL;
void orc_set_clones(const int32_t *owner, const uint8_t *hear, const char *const *words)
{
    orc_clone_owner = owner; orc_clone_hear = hear; orc_clone_words = words;
}

static int orc_live(int64_t i, const uint8_t *of, const int32_t *gate, const uint8_t *verdict)
{
    if (!gate || gate[i] < 0) return 1;
    int v = verdict[gate[i]] != 0;
    return (of[i] & ORC_OF_GATE_IF_SET) ? v : !v;
}

uint64_t orc_fnv1a(const uint8_t *p, size_t n)
{
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}

/* ---- batch drivers ------------------------------------------------------- */

void orc_streams_free(orc_streams *s)
{
    if (!s) return;
    free(s->off); free(s->bytes); free(s->n_deliveries);
    s->off = NULL; s->bytes = NULL; s->n_deliveries = NULL;
}

int orc_write_batch(int64_t n_ops, const uint8_t *text, const uint64_t *toff,
                    const uint8_t *kind, const int32_t *target,
                    const int32_t *except_user, const uint8_t *of,
                    const int32_t *gate, const uint8_t *verdict,
                    int32_t n_users, const int32_t *room, const uint8_t *uf,
                    const uint8_t *ul,
                    const int32_t *only, int32_t n_only, orc_streams *out)
{
    uint8_t *want = NULL;
    if (only) {
        want = calloc((size_t)n_users + 1, 1);
        if (!want) return -1;
        for (int32_t k = 0; k < n_only; ++k)
            if (only[k] >= 0 && only[k] < n_users) want[only[k]] = 1;
    }
    size_t maxn = 0;
    for (int64_t i = 0; i < n_ops; ++i) {
        size_t n = (size_t)(toff[i + 1] - toff[i]);
        if (n > maxn) maxn = n;
    }
    uint8_t *ron = malloc(6 * maxn + 8), *roff = malloc(2 * maxn + 8);
    char *relay = malloc(maxn + 64); uint8_t *rrel = malloc(6 * (maxn + 64) + 8);
    out->n_users = n_users;
    out->off = calloc((size_t)n_users + 1, sizeof(uint64_t));
    out->n_deliveries = calloc((size_t)n_users + 1, sizeof(uint64_t));
    uint64_t *cur = calloc((size_t)n_users + 1, sizeof(uint64_t));
    out->bytes = NULL;
    if (!ron || !roff || !out->off || !out->n_deliveries || !cur) {
        free(ron); free(roff); free(cur); free(want); orc_streams_free(out); return -1;
    }
    /* two passes in the reference's order: ops outer, user list inner (c:1409) */
    for (int pass = 0; pass < 2; ++pass) {
        for (int64_t i = 0; i < n_ops; ++i) {
            if (!orc_live(i, of, gate, verdict)) continue;
            const uint8_t *s = text + toff[i];
            size_t n = (size_t)(toff[i + 1] - toff[i]);
            size_t lon = 0, loff = 0; int have_on = 0, have_off = 0;
            int32_t u0 = 0, u1 = n_users;
            if (kind[i] == ORC_OP_USER) {
                if (target[i] < 0 || target[i] >= n_users) continue;
                u0 = target[i]; u1 = u0 + 1;
            }
            for (int32_t u = u0; u < u1; ++u) {
                if (want && !want[u]) continue;
                if (!orc_delivers(kind[i], target[i], except_user[i], of[i], u, room[u], uf[u], ul[u]))
                    continue;
                if (kind[i] == ORC_OP_ROOM && (uf[u] & ORC_UF_CLONE)) {                 /* c:1416-1426 */
                    if (!orc_clone_owner || !relay || !rrel) continue;
                    const int32_t o = orc_clone_owner[u];
                    if (o < 0 || o >= n_users) continue;
                    if (orc_clone_hear[u] == 0 || (uf[o] & ORC_UF_IGNALL)) continue;      /* c:1417 */
                    if (target[i] != room[u]) continue;                                   /* c:1420: rm!=u->room */
                    if (orc_clone_hear[u] == 1 && !orc_contains_swearing(s, n, orc_clone_words)) continue;
                    if (want && !want[o]) continue;
                    const int pl = sprintf(relay, "~FT[ room%d ]:~RS ", (int)room[u]);    /* c:1424 */
                    memcpy(relay + pl, s, n);
                    size_t rl; int32_t dst = o;
                    if (uf[o] & ORC_UF_REMOTE) {                                          /* write_user(owner): c:1299 */
                        if (!orc_remote_link || orc_remote_link[o] < 0) continue;
                        dst = orc_remote_link[o];
                        if (want && !want[dst]) continue;
                        rl = orc_remote_frame(o, (const uint8_t *)relay, (size_t)pl + n, rrel);
                    } else rl = orc_render_ex((const uint8_t *)relay, (size_t)pl + n, (uf[o] & ORC_UF_COLOUR) != 0, 0, rrel);
                    if (pass == 0) { out->off[dst + 1] += rl; out->n_deliveries[dst] += 1; }
                    else { memcpy(out->bytes + cur[dst], rrel, rl); cur[dst] += rl; }
                    continue;
                }
                if (kind[i] == ORC_OP_USER && (uf[u] & ORC_UF_CLONE)) continue;       /* a clone has no socket of its own */
                if (uf[u] & ORC_UF_REMOTE) {                                            /* c:1299-1307 */
                    if (!orc_remote_link || !rrel) continue;
                    const int32_t l = orc_remote_link[u];
                    if (l < 0 || l >= n_users || (want && !want[l])) continue;
                    const size_t rl = orc_remote_frame(u, s, n, rrel);
                    if (pass == 0) { out->off[l + 1] += rl; out->n_deliveries[l] += 1; }
                    else { memcpy(out->bytes + cur[l], rrel, rl); cur[l] += rl; }
                    continue;
                }
                int c = (uf[u] & ORC_UF_COLOUR) != 0;
                if (c && !have_on)  { lon  = orc_render_ex(s, n, 1, of[i], ron);  have_on = 1; }
                if (!c && !have_off){ loff = orc_render_ex(s, n, 0, of[i], roff); have_off = 1; }
                size_t len = c ? lon : loff;
                if (pass == 0) { out->off[u + 1] += len; out->n_deliveries[u] += 1; }
                else { memcpy(out->bytes + cur[u], c ? ron : roff, len); cur[u] += len; }
            }
        }
        if (pass == 0) {
            for (int32_t u = 0; u < n_users; ++u) out->off[u + 1] += out->off[u];
            for (int32_t u = 0; u < n_users; ++u) cur[u] = out->off[u];
            out->bytes = malloc(out->off[n_users] + 1);
            if (!out->bytes) { free(ron); free(roff); free(cur); free(want); orc_streams_free(out); return -1; }
        }
    }
    free(ron); free(roff); free(cur); free(want); free(relay); free(rrel);
    return 0;
}

/* Parity digests: d = FNV-1a over one delivery's rendered bytes; per user, the left fold over his
 * deliveries in call order, per op the same fold over its recipients in user-list order (the order of the loop at
 * nuts333.c:1409):  D <- (D * 0x9E3779B97F4A7C15) ^ d ^ len,  D0 = 0.  A delivery of no bytes (an empty string for a
 * colour-off recipient) makes no write(2) at all in the reference and is not folded.  Populations without clones /
 * remote users (their relays are deliveries of other strings to other sockets).  Either array may be NULL. */
#define ORC_DG_K 0x9E3779B97F4A7C15ull
int orc_delivery_digests(int64_t n_ops, const uint8_t *text, const uint64_t *toff,
                         const uint8_t *kind, const int32_t *target,
                         const int32_t *except_user, const uint8_t *of,
                         const int32_t *gate, const uint8_t *verdict,
                         int32_t n_users, const int32_t *room, const uint8_t *uf, const uint8_t *ul,
                         uint64_t *per_user, uint64_t *per_op)
{
    size_t maxn = 0;
    for (int64_t i = 0; i < n_ops; ++i) { size_t n = (size_t)(toff[i + 1] - toff[i]); if (n > maxn) maxn = n; }
    for (int32_t u = 0; u < n_users; ++u) if (uf[u] & (ORC_UF_CLONE | ORC_UF_REMOTE)) return -2;
    uint8_t *buf = malloc(6 * maxn + 8);
    if (!buf) return -1;
    if (per_user) memset(per_user, 0, sizeof(uint64_t) * (size_t)n_users);
    for (int64_t i = 0; i < n_ops; ++i) {
        uint64_t D = 0;
        if (per_op) per_op[i] = 0;
        if (kind[i] > ORC_OP_LEVEL || !orc_live(i, of, gate, verdict)) continue;
        const uint8_t *s = text + toff[i];
        const size_t n = (size_t)(toff[i + 1] - toff[i]);
        uint64_t dv[2] = { 0, 0 }; size_t lv[2] = { 0, 0 }; int have[2] = { 0, 0 };
        int32_t u0 = 0, u1 = n_users;
        if (kind[i] == ORC_OP_USER) {
            if (target[i] < 0 || target[i] >= n_users) continue;
            u0 = target[i]; u1 = u0 + 1;
        }
        for (int32_t u = u0; u < u1; ++u) {
            if (!orc_delivers(kind[i], target[i], except_user[i], of[i], u, room[u], uf[u], ul[u])) continue;
            const int c = (uf[u] & ORC_UF_COLOUR) != 0;
            if (!have[c]) { lv[c] = orc_render_ex(s, n, c, of[i], buf); dv[c] = orc_fnv1a(buf, lv[c]); have[c] = 1; }
            if (!lv[c]) continue;
            D = (D * ORC_DG_K) ^ dv[c] ^ (uint64_t)lv[c];
            if (per_user) per_user[u] = (per_user[u] * ORC_DG_K) ^ dv[c] ^ (uint64_t)lv[c];
        }
        if (per_op) per_op[i] = D;
    }
    free(buf);
    return 0;
}

int64_t orc_write_batch_count(int64_t n_ops, const uint8_t *text, const uint64_t *toff,
                    const uint8_t *kind, const int32_t *target,
                    const int32_t *except_user, const uint8_t *of,
                    const int32_t *gate, const uint8_t *verdict,
                    int32_t n_users, const int32_t *room, const uint8_t *uf,
                    const uint8_t *ul, uint64_t *out_bytes)
{
    /* Faithful cost model: the reference renders the string again for EVERY
     * recipient (write_user is called per user, c:1427), so does this loop. */
    size_t maxn = 0;
    for (int64_t i = 0; i < n_ops; ++i) {
        size_t n = (size_t)(toff[i + 1] - toff[i]);
        if (n > maxn) maxn = n;
    }
    uint8_t *buf = malloc(6 * maxn + 8);
    if (!buf) return -1;
    int64_t deliveries = 0; uint64_t bytes = 0; volatile uint8_t sink = 0;
    for (int64_t i = 0; i < n_ops; ++i) {
        if (!orc_live(i, of, gate, verdict)) continue;
        const uint8_t *s = text + toff[i];
        size_t n = (size_t)(toff[i + 1] - toff[i]);
        int32_t u0 = 0, u1 = n_users;
        if (kind[i] == ORC_OP_USER) {
            if (target[i] < 0 || target[i] >= n_users) continue;
            u0 = target[i]; u1 = u0 + 1;
        }
        for (int32_t u = u0; u < u1; ++u) {
            if (!orc_delivers(kind[i], target[i], except_user[i], of[i], u, room[u], uf[u], ul[u]))
                continue;
            size_t len = orc_render_ex(s, n, (uf[u] & ORC_UF_COLOUR) != 0, of[i], buf);
            sink ^= buf[len ? len - 1 : 0];
            bytes += len; ++deliveries;
        }
    }
    free(buf);
    if (out_bytes) *out_bytes = bytes;
    return deliveries;
}

void orc_contains_swearing_batch(int64_t n, const uint8_t *text, const uint64_t *off,
                                 const char *const *words, uint8_t *verdict)
{
    for (int64_t i = 0; i < n; ++i)
        verdict[i] = (uint8_t)orc_contains_swearing(text + off[i], (size_t)(off[i + 1] - off[i]), words);
}

void orc_site_banned_batch(const uint8_t *f, size_t fn, int present, int64_t n,
                           const uint8_t *text, const uint64_t *off, uint8_t *verdict)
{
    for (int64_t i = 0; i < n; ++i)
        verdict[i] = (uint8_t)orc_site_banned(f, fn, present, text + off[i], (size_t)(off[i + 1] - off[i]));
}

void orc_user_banned_batch(const uint8_t *f, size_t fn, int present, int64_t n,
                           const uint8_t *text, const uint64_t *off, uint8_t *verdict)
{
    for (int64_t i = 0; i < n; ++i)
        verdict[i] = (uint8_t)orc_user_banned(f, fn, present, text + off[i], (size_t)(off[i + 1] - off[i]));
}

/* more(user, sock, filename), c:2205-2322, local recipient.  file/n = the file's bytes
 * (present == 0: fopen failed).  Appends what the reference writes to the socket to out
 * (capacity: 6*n + 256 is always enough), updates *filepos like user->filepos and
 * returns more()'s return value.  user_null = the login-stage call more(NULL,sock,file). */
int orc_more(const uint8_t *f, size_t n, int present, int user_null, int colour,
             int64_t *filepos, uint8_t *out, size_t *out_len)
{
    static const char prompt[] = "           ~BB*** Press <return> to continue, 'e'<return> to exit ***";
    size_t o = 0, pos = 0;
    *out_len = 0;
    if (!present) { if (!user_null) *filepos = 0; return 0; }            /* c:2214-2217 */
    if (!user_null && *filepos > 0) pos = (size_t)*filepos < n ? (size_t)*filepos : n;   /* c:2219 */
    uint8_t text[2000]; size_t tl = 0;         /* text[ARR_SIZE*2], fgets(text,1999,fp) */
    int eof = 0, lines = 0; int64_t num_chars = 0;
#define ORC_FGETS() do { size_t got = 0; uint8_t tmp[2000]; \
        while (got < 1998) { if (pos >= n) { eof = 1; break; } tmp[got] = f[pos++]; if (tmp[got++] == '\n') break; } \
        if (got) { memcpy(text, tmp, got); tl = got; } } while (0)
    ORC_FGETS();
    while (!eof && (lines < 23 || user_null)) {                          /* c:2237 */
        size_t len = 0; while (len < tl && text[len]) ++len;             /* while(*str), strlen(text) */
        o += orc_render_ex(text, len, colour, ORC_OF_PAGER | (user_null ? ORC_OF_PLAIN : 0), out + o);
        num_chars += (int64_t)len;
        lines += (int)(len / 80) + (len < 80);                           /* c:2303 */
        ORC_FGETS();
    }
#undef ORC_FGETS
    *out_len = o;
    if (user_null) return 2;                                             /* c:2309 */
    if (eof) { *filepos = 0; return 2; }                                 /* c:2310-2312 */
    *filepos += num_chars;                                               /* c:2316 */
    o += orc_render((const uint8_t *)prompt, sizeof prompt - 1, colour, out + o);   /* c:2321 write_user */
    *out_len = o;
    return 1;
}

/* ---- the callers: say / shout / emote / semote / echo / bcast ---------------------------
 * Restated call by call from nuts333.c; each appends the write calls the reference makes
 * to an op list (orc_speech_emit).  user_name = user->name, vis/muzzled = the user's
 * fields, room = the user's room index or -1, ban_swearing = the global. */
#include <stdio.h>
typedef struct {
    uint8_t *text; size_t text_len, text_cap;
    uint64_t *off; uint8_t *kind; int32_t *target; int32_t *except_user; uint8_t *flags;
    int64_t n, cap;
} orc_oplist;

static void orc_emit(orc_oplist *L, uint8_t kind, int32_t target, int32_t exc, uint8_t flags, const char *str, size_t n)
{
    if (L->n >= L->cap || L->text_len + n > L->text_cap) { L->n = L->cap + 1; return; }   /* overflow marker */
    memcpy(L->text + L->text_len, str, n); L->text_len += n;
    L->kind[L->n] = kind; L->target[L->n] = target; L->except_user[L->n] = exc; L->flags[L->n] = flags;
    L->off[++L->n] = L->text_len;
}

static const char orc_noswearing[] = "Swearing is not allowed here.\n";      /* h:151 */
static const char orc_invisname[] = "A presence";                            /* h:150 */

/* verbs: 0 say c:4062, 1 shout c:4105, 2 emote c:4188, 3 semote c:4213, 4 echo c:4289, 5 bcast c:4772 */
/* ---- ban-list maintenance ------------------------------------------------------------------
 * Both commands read the list with the same fscanf("%s") / while(!feof) loop as site_banned
 * (c:6232-6240, c:6359-6367): a last token that runs into EOF is never looked at -- so it can be
 * banned twice, and unban silently DROPS it from the rewritten file.  ban appends "token\n" with
 * fopen("a") (c:6250): after a file without a final newline the new token is glued to the last one.
 * unban rewrites every other tested token as "token\n"; an emptied list is unlinked (c:6376).
 * user tokens: first byte upper-cased (c:6269, c:6402).  Comparison is strcmp, both lists. */
int orc_ban_edit(const uint8_t *file, size_t n, int present, int is_user, int add, const char *token,
                 uint8_t *out, size_t *out_n, int *out_present)
{
    char tok[128]; size_t tl = strlen(token);
    if (tl > 80) tl = 80;
    memcpy(tok, token, tl); tok[tl] = 0;
    if (is_user && tok[0] >= 'a' && tok[0] <= 'z') tok[0] = (char)(tok[0] - 32);
    size_t p = 0, o = 0; int found = 0, cnt = 0;
    *out_present = present; *out_n = present ? n : 0;
    if (present) memcpy(out, file, n);
    if (!present && !add) return 1;                              /* c:6349: cannot open -> "not currently banned" */
    if (present) {
        for (;;) {
            while (p < n && (file[p] == ' ' || (file[p] >= 9 && file[p] <= 13))) ++p;
            if (p >= n) break;
            const size_t b = p;
            while (p < n && !(file[p] == ' ' || (file[p] >= 9 && file[p] <= 13))) ++p;
            if (p >= n) break;                                   /* ran into EOF: feof() is set, the loop ends */
            size_t m = 0; while (m < p - b && file[b + m]) ++m;  /* a C string: cut at NUL */
            const int same = m == tl && memcmp(file + b, tok, m) == 0;
            if (add) { if (same) return 1; }                     /* "already banned" */
            else if (same) found = 1;
            else { memcpy(out + o, file + b, m); o += m; out[o++] = '\n'; ++cnt; }
        }
    }
    if (add) {
        const size_t base = present ? n : 0;
        memcpy(out + base, tok, tl); out[base + tl] = '\n';
        *out_n = base + tl + 1; *out_present = 1;
        return 0;
    }
    if (!found) { memcpy(out, file, n); *out_n = n; return 1; }   /* tempfile unlinked, list untouched */
    *out_n = o; *out_present = cnt ? 1 : 0;
    if (!cnt) *out_n = 0;
    return 0;
}

/* Review buffers, c:2062-2071 (record) and c:5192-5222 (review): 15 lines of 200 bytes per room, a ring.
 * strncpy pads with NULs, then byte 200 is set to '\n' and byte 201 to NUL: a line of 200 bytes or more is
 * kept as its first 200 bytes plus a newline, a shorter one as it is (the '\n' sits behind the terminator). */
#define ORC_REVIEW_LINES 15
#define ORC_REVIEW_LEN   200
typedef struct { char buf[ORC_REVIEW_LINES][ORC_REVIEW_LEN + 2]; int line; } orc_revbuf;
static orc_revbuf *orc_rev = NULL; static int32_t orc_rev_rooms = 0;

static void orc_record(int32_t room, const char *str)
{
    if (!orc_rev || room < 0 || room >= orc_rev_rooms) return;
    orc_revbuf *rb = &orc_rev[room];
    strncpy(rb->buf[rb->line], str, ORC_REVIEW_LEN);
    rb->buf[rb->line][ORC_REVIEW_LEN] = '\n';
    rb->buf[rb->line][ORC_REVIEW_LEN + 1] = '\0';
    rb->line = (rb->line + 1) % ORC_REVIEW_LINES;
}

static void orc_emit(orc_oplist *L, uint8_t kind, int32_t target, int32_t exc, uint8_t flags, const char *str, size_t n);
/* review() with no argument: the user's own room (c:5200) */
static void orc_review(orc_oplist *L, int32_t user, int32_t room)
{
    char text[512]; int cnt = 0;
    if (!orc_rev || room < 0 || room >= orc_rev_rooms) return;
    const orc_revbuf *rb = &orc_rev[room];
    for (int i = 0; i < ORC_REVIEW_LINES; ++i) {
        const int line = (rb->line + i) % ORC_REVIEW_LINES;
        if (rb->buf[line][0]) {
            if (++cnt == 1) {
                const int len = snprintf(text, sizeof text, "\n~BB~FG*** Review buffer for the room%d ***\n\n", room);
                orc_emit(L, 0, user, -1, 0, text, (size_t)len);
            }
            orc_emit(L, 0, user, -1, 0, rb->buf[line], strlen(rb->buf[line]));
        }
    }
    if (!cnt) { const char *m = "Review buffer is empty.\n"; orc_emit(L, 0, user, -1, 0, m, strlen(m)); }
    else { const char *m = "\n~BB~FG*** End ***\n\n"; orc_emit(L, 0, user, -1, 0, m, strlen(m)); }
}

/* tell c:4128 / pemote c:4234 / wizshout c:6527 / revtell c:7699: targets of tell and pemote as the caller
 * resolved them (get_user), the target's name, the per-user revtell buffers (5 lines, record_tell c:2074). */
#define ORC_REVTELL_LINES 5
typedef struct { char buf[ORC_REVTELL_LINES][ORC_REVIEW_LEN + 2]; int line; } orc_tellbuf;
static orc_tellbuf *orc_tell = NULL; static int32_t orc_tell_users = 0;
static const int32_t *orc_speech_target = NULL;          /* per input line */
static const uint8_t *orc_names = NULL; static const uint64_t *orc_name_off = NULL;
void orc_set_speech_targets(const int32_t *target) { orc_speech_target = target; }

static void orc_record_tell(int32_t u, const char *str)
{
    if (!orc_tell || u < 0 || u >= orc_tell_users) return;
    orc_tellbuf *rb = &orc_tell[u];
    strncpy(rb->buf[rb->line], str, ORC_REVIEW_LEN);
    rb->buf[rb->line][ORC_REVIEW_LEN] = '\n';
    rb->buf[rb->line][ORC_REVIEW_LEN + 1] = '\0';
    rb->line = (rb->line + 1) % ORC_REVTELL_LINES;
}

static void orc_speech_private(orc_oplist *L, int verb, int32_t user, const char *uname, int vis, int muzzled,
                               int ban_swearing, const char *const *words, const uint8_t *in, size_t n, int64_t line)
{
    char text[4400], tname[64]; int len;
    const char *name = vis ? uname : orc_invisname;
    const int32_t t = orc_speech_target ? orc_speech_target[line] : -1;
    tname[0] = 0;
    if (t >= 0 && orc_names) {
        size_t nl = (size_t)(orc_name_off[t + 1] - orc_name_off[t]); if (nl > 63) nl = 63;
        memcpy(tname, orc_names + orc_name_off[t], nl); tname[nl] = 0;
    }
    switch (verb) {
    case 7: {                                                   /* tell, c:4128 */
        if (muzzled) { const char *m = "You are muzzled, you cannot tell anyone anything.\n"; orc_emit(L, 0, user, -1, 0, m, strlen(m)); return; }
        const char *type = (n && in[n - 1] == '?') ? "ask" : "tell";                  /* c:4175 */
        len = snprintf(text, sizeof text, "~OLYou %s %s:~RS %.*s\n", type, tname, (int)n, (const char *)in);
        orc_emit(L, 0, user, -1, 0, text, (size_t)len);
        len = snprintf(text, sizeof text, "~OL%s %ss you:~RS %.*s\n", name, type, (int)n, (const char *)in);
        orc_emit(L, 0, t, -1, 0, text, (size_t)len);
        orc_record_tell(t, text);
        return; }
    case 8:                                                     /* pemote, c:4234 */
        if (muzzled) { const char *m = "You are muzzled, you cannot emote.\n"; orc_emit(L, 0, user, -1, 0, m, strlen(m)); return; }
        len = snprintf(text, sizeof text, "~OL(To %s)~RS %s %.*s\n", tname, name, (int)n, (const char *)in);
        orc_emit(L, 0, user, -1, 0, text, (size_t)len);
        len = snprintf(text, sizeof text, "~OL>>~RS %s %.*s\n", name, (int)n, (const char *)in);
        orc_emit(L, 0, t, -1, 0, text, (size_t)len);
        orc_record_tell(t, text);
        return;
    case 9:                                                     /* wizshout without a level word, c:6527 */
        if (muzzled) { const char *m = "You are muzzled, you cannot wizshout.\n"; orc_emit(L, 0, user, -1, 0, m, strlen(m)); return; }
        if (ban_swearing && orc_contains_swearing(in, n, words)) { orc_emit(L, 0, user, -1, 0, orc_noswearing, sizeof orc_noswearing - 1); return; }
        len = snprintf(text, sizeof text, "~OLYou wizshout:~RS %.*s\n", (int)n, (const char *)in);
        orc_emit(L, 0, user, -1, 0, text, (size_t)len);
        len = snprintf(text, sizeof text, "~OL%s wizshouts:~RS %.*s\n", uname, (int)n, (const char *)in);
        orc_emit(L, 2, 2 /* WIZ */, user, ORC_OF_ABOVE, text, (size_t)len);          /* c:6564 write_level(WIZ,1,text,user) */
        return;
    case 10: {                                                  /* revtell, c:7699 */
        int cnt = 0;
        if (!orc_tell || user < 0 || user >= orc_tell_users) return;
        const orc_tellbuf *rb = &orc_tell[user];
        for (int i = 0; i < ORC_REVTELL_LINES; ++i) {
            const int ln = (rb->line + i) % ORC_REVTELL_LINES;
            if (rb->buf[ln][0]) {
                if (++cnt == 1) { const char *m = "\n~BB~FG*** Your revtell buffer ***\n\n"; orc_emit(L, 0, user, -1, 0, m, strlen(m)); }
                orc_emit(L, 0, user, -1, 0, rb->buf[ln], strlen(rb->buf[ln]));
            }
        }
        if (!cnt) { const char *m = "Revtell buffer is empty.\n"; orc_emit(L, 0, user, -1, 0, m, strlen(m)); }
        else { const char *m = "\n~BB~FG*** End ***\n\n"; orc_emit(L, 0, user, -1, 0, m, strlen(m)); }
        return; }
    }
}

static void orc_speech_one(orc_oplist *L, int verb, int32_t user, const char *uname, int vis, int muzzled, int32_t room,
                           int ban_swearing, const char *const *words, const uint8_t *in, size_t n)
{
    char text[4200]; int len;
    const char *name = vis ? uname : orc_invisname;
    const int sw = ban_swearing && orc_contains_swearing(in, n, words);
    switch (verb) {
    case 0: {
        const char *type = "say";
        if (muzzled) { const char *m = "You are muzzled, you cannot speak.\n"; orc_emit(L, 0, user, -1, 0, m, strlen(m)); return; }
        if (room < 0) return;                                   /* c:4071: relayed over the netlink */
        if (n && in[n - 1] == '?') type = "ask"; else if (n && in[n - 1] == '!') type = "exclaim";   /* c:4080 */
        if (sw) { orc_emit(L, 0, user, -1, 0, orc_noswearing, sizeof orc_noswearing - 1); return; }   /* c:4091 */
        len = snprintf(text, sizeof text, "You %s: %.*s\n", type, (int)n, (const char *)in);
        orc_emit(L, 0, user, -1, 0, text, (size_t)len);                              /* c:4095 */
        len = snprintf(text, sizeof text, "%s %ss: %.*s\n", name, type, (int)n, (const char *)in);
        orc_emit(L, 1, room, user, 0, text, (size_t)len);                            /* c:4098 */
        orc_record(room, text);                                                      /* c:4099 */
        return; }
    case 1:
        if (muzzled) { const char *m = "You are muzzled, you cannot shout.\n"; orc_emit(L, 0, user, -1, 0, m, strlen(m)); return; }
        if (sw) { orc_emit(L, 0, user, -1, 0, orc_noswearing, sizeof orc_noswearing - 1); return; }
        len = snprintf(text, sizeof text, "~OLYou shout:~RS %.*s\n", (int)n, (const char *)in);
        orc_emit(L, 0, user, -1, 0, text, (size_t)len);
        len = snprintf(text, sizeof text, "~OL%s shouts:~RS %.*s\n", name, (int)n, (const char *)in);
        orc_emit(L, 1, -1, user, ORC_OF_SHOUT, text, (size_t)len);                   /* c:4125, com_num==SHOUT */
        return;
    case 2:
        if (muzzled) { const char *m = "You are muzzled, you cannot emote.\n"; orc_emit(L, 0, user, -1, 0, m, strlen(m)); return; }
        if (room < 0) return;                                   /* not reached for users in no room (record(NULL)) */
        if (sw) { orc_emit(L, 0, user, -1, 0, orc_noswearing, sizeof orc_noswearing - 1); return; }
        if (n && in[0] == ';') len = snprintf(text, sizeof text, "%s%.*s\n", name, (int)(n - 1), (const char *)in + 1);
        else len = snprintf(text, sizeof text, "%s %.*s\n", name, (int)n, (const char *)in);
        orc_emit(L, 1, room, -1, 0, text, (size_t)len);                              /* c:4208 write_room(user->room) */
        orc_record(room, text);                                                      /* c:4209 */
        return;
    case 3:
        if (muzzled) { const char *m = "You are muzzled, you cannot emote.\n"; orc_emit(L, 0, user, -1, 0, m, strlen(m)); return; }
        if (n && in[0] == '#') len = snprintf(text, sizeof text, "~OL!!~RS %s%.*s\n", name, (int)(n - 1), (const char *)in + 1);
        else len = snprintf(text, sizeof text, "~OL!!~RS %s %.*s\n", name, (int)n, (const char *)in);
        orc_emit(L, 1, -1, -1, ORC_OF_SHOUT, text, (size_t)len);                     /* c:4231, com_num==SEMOTE */
        return;
    case 4:
        if (muzzled) { const char *m = "You are muzzled, you cannot echo.\n"; orc_emit(L, 0, user, -1, 0, m, strlen(m)); return; }
        if (room < 0) return;                                   /* as emote */
        len = snprintf(text, sizeof text, "(%s) ", uname);
        orc_emit(L, 2, 2 /* WIZ */, -1, ORC_OF_ABOVE, text, (size_t)len);            /* c:4301 write_level(WIZ,1,text,NULL) */
        len = snprintf(text, sizeof text, "- %.*s\n", (int)n, (const char *)in);
        orc_emit(L, 1, room, -1, 0, text, (size_t)len);                              /* c:4303 */
        orc_record(room, text);                                                      /* c:4304 */
        return;
    case 6:                                                     /* review, c:5192: the speaker looks at his own room */
        orc_review(L, user, room);
        return;
    case 5:
        if (muzzled) { const char *m = "You are muzzled, you cannot broadcast anything.\n"; orc_emit(L, 0, user, -1, 0, m, strlen(m)); return; }
        if (vis) len = snprintf(text, sizeof text, "\07\n~BR*** Broadcast message from %s ***\n%.*s\n\n", uname, (int)n, (const char *)in);
        else len = snprintf(text, sizeof text, "\07\n~BR*** Broadcast message ***\n%.*s\n\n", (int)n, (const char *)in);
        orc_emit(L, 1, -1, -1, ORC_OF_FORCE_LISTEN, text, (size_t)len);              /* c:4783-4787 */
        return;
    }
}

/* n input lines -> the ops the reference's callers make, in order.  names = packed user
 * names with name_off[n_users+1]; sflags bit0 invisible, bit1 muzzled; room[u] or -1.
 * Arrays sized for cap ops / text_cap bytes.  Returns the number of ops, or -1 on overflow. */
int64_t orc_speech_ops(int64_t n, const uint8_t *verb, const int32_t *speaker, const uint8_t *bodies, const uint64_t *body_off,
                       const uint8_t *names, const uint64_t *name_off, const uint8_t *sflags, const int32_t *room,
                       int ban_swearing, const char *const *words,
                       uint8_t *text, size_t text_cap, uint64_t *off, uint8_t *kind, int32_t *target, int32_t *except_user,
                       uint8_t *flags, int64_t cap)
{
    orc_oplist L = { text, 0, text_cap, off, kind, target, except_user, flags, 0, cap };
    off[0] = 0;
    /* review buffers start empty (c:2799) and live for the n lines of this call */
    int32_t max_room = -1, n_users = 0;
    for (int64_t m = 0; m < n; ++m) if (speaker[m] >= n_users) n_users = speaker[m] + 1;
    for (int32_t u = 0; u < n_users; ++u) if (room[u] > max_room) max_room = room[u];
    orc_rev_rooms = max_room + 1;
    orc_rev = orc_rev_rooms ? (orc_revbuf *)calloc((size_t)orc_rev_rooms, sizeof(orc_revbuf)) : NULL;
    if (orc_speech_target) for (int64_t m = 0; m < n; ++m) if (orc_speech_target[m] >= n_users) n_users = orc_speech_target[m] + 1;
    orc_tell_users = n_users;
    orc_tell = n_users ? (orc_tellbuf *)calloc((size_t)n_users, sizeof(orc_tellbuf)) : NULL;
    orc_names = names; orc_name_off = name_off;
    for (int64_t m = 0; m < n; ++m) {
        const int32_t u = speaker[m];
        char uname[64]; size_t nl = (size_t)(name_off[u + 1] - name_off[u]);
        if (nl > 63) nl = 63;
        memcpy(uname, names + name_off[u], nl); uname[nl] = 0;
        if (verb[m] >= 7) orc_speech_private(&L, verb[m], u, uname, !(sflags[u] & 1), (sflags[u] & 2) != 0, ban_swearing, words,
                                             bodies + body_off[m], (size_t)(body_off[m + 1] - body_off[m]), m);
        else
        orc_speech_one(&L, verb[m], u, uname, !(sflags[u] & 1), (sflags[u] & 2) != 0, room[u], ban_swearing, words,
                       bodies + body_off[m], (size_t)(body_off[m + 1] - body_off[m]));
        if (L.n > L.cap) { free(orc_rev); orc_rev = NULL; free(orc_tell); orc_tell = NULL; return -1; }
    }
    free(orc_rev); orc_rev = NULL; free(orc_tell); orc_tell = NULL;
    return L.n;
}
