/* oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Drives the UNMODIFIED reference (nuts333.c, pulled in by #include from where
 * it lies under /root/reference -- never copied into this repo) so that its own
 * write_user / write_room_except / write_level / contains_swearing /
 * site_banned / user_banned can be called in-process and their socket output
 * captured.  Built by oracle/Makefile into oracle/_ref/libnutsref.so (git-ignored,
 * travels to the GPU box as a binary).
 *
 * The three tricks:
 *   - <unistd.h> first, then `#define write nutsref_write` so every write(2)
 *     in the reference lands in the capture sink below;
 *   - `#define main nutsref_main` so the talker's main() is just a function;
 *   - users/rooms are made with the reference's own create_user()/create_room()
 *     with user->socket = the user's index.
 * The one permitted edit -- the swear_words[] initializer (h:271-277 invites
 * it) -- is done by the Makefile on a scratch copy of nuts333.h so that the
 * table has NUTSREF_MAX_SWEAR slots that ref_set_swear_words() can fill.
 */
#include <unistd.h>
#include <stdint.h>
#include <stddef.h>
#include <sys/stat.h>
#include <sys/types.h>

static ssize_t nutsref_write(int fd, const void *buf, size_t n);

#define write nutsref_write
#define main  nutsref_main
#include "nuts333.c"
#undef main
#undef write

/* ---- capture sink -------------------------------------------------------- */

typedef struct { uint8_t *p; size_t n, cap; uint64_t calls; } sink_t;
static sink_t  *g_sink = NULL;
static int      g_nsink = 0;
static int      g_sink_mode = 0;      /* 0 keep bytes, 1 count only, 2 real write to /dev/null */
static int      g_devnull = -1;
static uint64_t g_write_calls = 0, g_write_bytes = 0;

/* parity digests taken from the reference's own write(2) calls: one delivery = the 1-3 calls
 * write_user makes for one recipient (buffer flushes c:1318,1339,1360,1363 and the closing reset c:1365), so the
 * running FNV-1a of a socket is carried across calls and closed when the harness moves on to the next op */
static int       g_dg_on = 0, g_dg_ntouched = 0;
static uint64_t *g_dg_fnv = NULL, *g_dg_len = NULL; static int *g_dg_order = NULL; static uint8_t *g_dg_touched = NULL;

static ssize_t nutsref_write(int fd, const void *buf, size_t n)
{
    ++g_write_calls; g_write_bytes += n;
    if (g_dg_on && fd >= 0 && fd < g_nsink && n) {
        if (!g_dg_touched[fd]) { g_dg_touched[fd] = 1; g_dg_order[g_dg_ntouched++] = fd; g_dg_fnv[fd] = 0xcbf29ce484222325ull; g_dg_len[fd] = 0; }
        uint64_t h = g_dg_fnv[fd];
        for (size_t k = 0; k < n; ++k) { h ^= ((const uint8_t *)buf)[k]; h *= 0x100000001b3ull; }
        g_dg_fnv[fd] = h; g_dg_len[fd] += n;
    }
    if (g_sink_mode == 2) return write(g_devnull, buf, n);   /* the real write(2) */
    if (g_sink_mode == 1 || fd < 0 || fd >= g_nsink) return (ssize_t)n;
    sink_t *s = &g_sink[fd];
    if (s->n + n > s->cap) {
        size_t nc = s->cap ? s->cap * 2 : 256;
        while (nc < s->n + n) nc *= 2;
        uint8_t *np = realloc(s->p, nc);
        if (!np) return -1;
        s->p = np; s->cap = nc;
    }
    memcpy(s->p + s->n, buf, n);
    s->n += n; s->calls++;
    return (ssize_t)n;
}

/* ---- population ---------------------------------------------------------- */

static NL_OBJECT g_links[4096];           /* netlink objects of the link pseudo-users (ref_set_remote) */
static UR_OBJECT *g_users = NULL;
static RM_OBJECT *g_rooms = NULL;
static int g_nusers = 0, g_nrooms = 0;

void ref_reset(void)
{
    for (int i = 0; i < g_nusers; ++i) free(g_users[i]);
    for (int i = 0; i < g_nrooms; ++i) free(g_rooms[i]);
    for (int i = 0; i < g_nsink; ++i) free(g_sink[i].p);
    free(g_users); free(g_rooms); free(g_sink);
    g_users = NULL; g_rooms = NULL; g_sink = NULL;
    g_nusers = g_nrooms = g_nsink = 0;
    g_write_calls = g_write_bytes = 0;
    memset(g_links, 0, sizeof g_links);
    init_globals();                    /* c:1032 */
    system_logging = 0;                /* keep write_syslog() away from the CWD */
    force_listen = 0;
    com_num = -1;
}

int ref_sizeof_user(void) { return (int)sizeof(struct user_struct); }

/* sink mode: 0 capture bytes, 1 count only, 2 write(2) to /dev/null */
void ref_set_sink_mode(int mode)
{
    g_sink_mode = mode;
    if (mode == 2 && g_devnull < 0) g_devnull = open("/dev/null", O_WRONLY);
}

int ref_add_rooms(int n)
{
    g_rooms = realloc(g_rooms, sizeof(RM_OBJECT) * (size_t)(g_nrooms + n));
    for (int i = 0; i < n; ++i) {
        RM_OBJECT r = create_room();   /* c:2776 */
        sprintf(r->name, "room%d", g_nrooms);
        g_rooms[g_nrooms++] = r;
    }
    return g_nrooms;
}

/* flags as ORC_UF_*: 1 colour, 2 login, 4 ignall, 8 ignshout, 16 clone. */
int ref_add_users(int n, const int32_t *room, const uint8_t *flags, const uint8_t *level)
{
    g_users = realloc(g_users, sizeof(UR_OBJECT) * (size_t)(g_nusers + n));
    g_sink  = realloc(g_sink, sizeof(sink_t) * (size_t)(g_nusers + n));
    for (int i = 0; i < n; ++i) {
        UR_OBJECT u = create_user();   /* c:2673: appended to user_first..user_last */
        if (!u) return -1;
        u->socket   = g_nusers;
        u->room     = (room[i] >= 0 && room[i] < g_nrooms) ? g_rooms[room[i]] : NULL;
        u->colour   = (flags[i] & 1) ? 1 : 0;
        u->login    = (flags[i] & 2) ? 3 : 0;
        u->ignall   = (flags[i] & 4) ? 1 : 0;
        u->ignshout = (flags[i] & 8) ? 1 : 0;
        u->level    = level[i];
        sprintf(u->name, "U%d", g_nusers);
        memset(&g_sink[g_nusers], 0, sizeof(sink_t));
        g_users[g_nusers++] = u;
    }
    g_nsink = g_nusers;
    return g_nusers;
}

/* user u becomes a clone of `owner` (c:7022-7036: type, owner, clone_hear; a clone has no socket) */
void ref_set_clone(int u, int owner, int hear)
{
    if (u < 0 || u >= g_nusers || owner < 0 || owner >= g_nusers) return;
    g_users[u]->type = CLONE_TYPE; g_users[u]->owner = g_users[owner]; g_users[u]->clone_hear = hear;
}

/* user u becomes a REMOTE_TYPE user on a netlink whose socket is the sink of pseudo-user `link`
 * (one netlink object per link user, created on first use); old != 0: a peer of version 3.1 */
void ref_set_remote(int u, int link, int old)
{
    if (u < 0 || u >= g_nusers || link < 0 || link >= g_nusers || link >= 4096) return;
    if (!g_links[link]) {
        g_links[link] = create_netlink();
        g_links[link]->socket = link;
    }
    g_links[link]->ver_major = 3; g_links[link]->ver_minor = old ? 1 : 3;
    g_users[u]->type = REMOTE_TYPE; g_users[u]->netlink = g_links[link];
}

/* ---- the reference's own entry points, one call each --------------------- */

static void ref_ambient(uint8_t oflags)
{
    force_listen = (oflags & 1) ? 1 : 0;
    com_num = (oflags & 2) ? SHOUT : SAY;
}

void ref_write_user(int u, const char *str)
{
    write_user((u >= 0 && u < g_nusers) ? g_users[u] : NULL, (char *)str);
}

void ref_write_room_except(int rm, const char *str, int except_user, uint8_t oflags)
{
    ref_ambient(oflags);
    write_room_except((rm >= 0 && rm < g_nrooms) ? g_rooms[rm] : NULL, (char *)str,
                      (except_user >= 0 && except_user < g_nusers) ? g_users[except_user] : NULL);
}

void ref_write_level(int level, int above, const char *str, int except_user)
{
    write_level(level, above, (char *)str,
                (except_user >= 0 && except_user < g_nusers) ? g_users[except_user] : NULL);
}

int ref_contains_swearing(const char *str) { return contains_swearing((char *)str); }
int ref_colour_com_count(const char *str)  { return colour_com_count((char *)str); }
size_t ref_colour_com_strip(const char *str, char *out)
{
    char *r = colour_com_strip((char *)str);
    size_t n = strlen(r);
    memcpy(out, r, n);
    return n;
}

#ifdef NUTSREF_MAX_SWEAR
/* words: NULL-terminated array WITHOUT the "*" sentinel; it is appended here. */
int ref_set_swear_words(const char *const *words)
{
    static char *own[NUTSREF_MAX_SWEAR];
    int n = 0;
    for (int i = 0; i < NUTSREF_MAX_SWEAR; ++i) { free(own[i]); own[i] = NULL; }
    while (words && words[n]) {
        if (n >= NUTSREF_MAX_SWEAR - 1) return -1;
        own[n] = strdup(words[n]); swear_words[n] = own[n]; ++n;
    }
    own[n] = strdup("*"); swear_words[n] = own[n];
    return n;
}
#endif

/* Ban files: the reference opens datafiles/siteban|userban relative to the CWD
 * (c:336,355).  dir must exist; pass data==NULL to remove the file. */
int ref_set_ban_file(const char *dir, int which, const void *data, size_t n)
{
    char path[512];
    if (chdir(dir) != 0) return -1;
    mkdir(DATAFILES, 0777);
    snprintf(path, sizeof path, "%s/%s", DATAFILES, which ? USERBAN : SITEBAN);
    if (!data) { unlink(path); return 0; }
    FILE *fp = fopen(path, "wb");
    if (!fp) return -1;
    if (n && fwrite(data, 1, n, fp) != n) { fclose(fp); return -1; }
    fclose(fp);
    return 0;
}
/* The reference's own ban / unban commands (c:6216-6429) run by user `by` in the CWD (a scratch dir
 * with datafiles/): which 0 site, 1 user; add 1 ban, 0 unban.  An offline user is banned only if his
 * userfile says a lower level (c:6287-6297): the harness puts a level-0 userfile in place first. */
int ref_ban_command(const char *dir, int by, int which, int add, const char *token)
{
    char path[600];
    if (by < 0 || by >= g_nusers || chdir(dir) != 0) return -1;
    mkdir(DATAFILES, 0777); mkdir(USERFILES, 0777);
    strncpy(word[2], token, WORD_LEN); word[2][WORD_LEN] = 0;
    word_count = 3;
    if (which && add) {
        char nm[WORD_LEN + 2]; strcpy(nm, word[2]); nm[0] = (char)toupper((unsigned char)nm[0]);
        snprintf(path, sizeof path, "%s/%s.D", USERFILES, nm);
        FILE *fp = fopen(path, "w");
        if (fp) { fputs("x\n0 0 0 0 0\n", fp); fclose(fp); }
    }
    if (!which) { if (add) ban_site(g_users[by]); else unban_site(g_users[by]); }
    else        { if (add) ban_user(g_users[by]); else unban_user(g_users[by]); }
    return 0;
}
/* the ban file as it now stands on disk: returns its length, -1 if it does not exist */
long ref_get_ban_file(const char *dir, int which, void *out, size_t cap)
{
    char path[600];
    snprintf(path, sizeof path, "%s/%s/%s", dir, DATAFILES, which ? USERBAN : SITEBAN);
    FILE *fp = fopen(path, "rb");
    if (!fp) return -1;
    const size_t n = fread(out, 1, cap, fp);
    fclose(fp);
    return (long)n;
}
int ref_site_banned(const char *site) { return site_banned((char *)site); }
int ref_user_banned(const char *name) { return user_banned((char *)name); }

/* The reference's own say()/shout()/emote()/semote()/echo()/bcast() (c:4062-4305, c:4772), one
 * input line each, with the ambient state the main loop / exec_com would have set. */
void ref_set_user_speech(int u, const char *name, int vis, int muzzled)
{
    if (u < 0 || u >= g_nusers) return;
    strncpy(g_users[u]->name, name, USER_NAME_LEN); g_users[u]->name[USER_NAME_LEN] = 0;
    g_users[u]->vis = vis; g_users[u]->muzzled = muzzled;
}
void ref_set_ban_swearing(int on) { ban_swearing = on; }
void ref_speech(int verb, int u, const char *inpstr)
{
    static char line[ARR_SIZE * 2];
    if (u < 0 || u >= g_nusers) return;
    strncpy(line, inpstr, sizeof line - 1); line[sizeof line - 1] = 0;
    force_listen = 0;                      /* cleared per input line, c:154 */
    word_count = 2;                        /* the callers' "say what?" checks are the caller's business */
    switch (verb) {
    case 0: com_num = SAY;    say(g_users[u], line); break;
    case 1: com_num = SHOUT;  shout(g_users[u], line); break;
    case 2: com_num = EMOTE;  emote(g_users[u], line); break;
    case 3: com_num = SEMOTE; semote(g_users[u], line); break;
    case 4: com_num = ECHO;   echo(g_users[u], line); break;
    case 5: com_num = BCAST;  bcast(g_users[u], line); break;
    case 6: com_num = REVIEW; word_count = 1; review(g_users[u]); break;   /* c:5192, no argument: the user's own room */
    case 9: com_num = WIZSHOUT; { char *sp; strncpy(word[1], line, WORD_LEN); word[1][WORD_LEN] = 0;
                                  if ((sp = strchr(word[1], ' '))) *sp = 0;     /* word[1] = first word (exec_com's split) */
                                  if (sp && *remove_first(line)) word_count = 3; }   /* ".wizshout <level> <message>" */
            wizshout(g_users[u], line); break;
    case 10: com_num = REVTELL; revtell(g_users[u]); break;
    }
    force_listen = 0;
}

/* tell() c:4128 / pemote() c:4234 to user t: the input line is "<name> <message>", word[1] the name */
void ref_speech_to(int verb, int u, int t, const char *msg)
{
    static char line[ARR_SIZE * 2];
    if (u < 0 || u >= g_nusers || t < 0 || t >= g_nusers) return;
    snprintf(line, sizeof line, "%s %s", g_users[t]->name, msg);
    strcpy(word[1], g_users[t]->name);
    word_count = 3; force_listen = 0;
    if (verb == 7) { com_num = TELL; tell(g_users[u], line); }
    else { com_num = PEMOTE; pemote(g_users[u], line); }
}

/* more(), c:2205: the reference's own pager on a file on disk, socket = the user's index */
int ref_more(int u, int null_user, const char *filename)
{
    if (u < 0 || u >= g_nusers) return -1;
    return more(null_user ? NULL : g_users[u], g_users[u]->socket, (char *)filename);
}
long ref_get_filepos(int u) { return (u >= 0 && u < g_nusers) ? (long)g_users[u]->filepos : -1; }

/* ---- stream access ------------------------------------------------------- */

size_t ref_stream_len(int u) { return (u >= 0 && u < g_nsink) ? g_sink[u].n : 0; }
const uint8_t *ref_stream_ptr(int u) { return (u >= 0 && u < g_nsink) ? g_sink[u].p : NULL; }
uint64_t ref_stream_calls(int u) { return (u >= 0 && u < g_nsink) ? g_sink[u].calls : 0; }
void ref_stream_clear(int u) { if (u >= 0 && u < g_nsink) { g_sink[u].n = 0; g_sink[u].calls = 0; } }
uint64_t ref_total_write_calls(void) { return g_write_calls; }
uint64_t ref_total_write_bytes(void) { return g_write_bytes; }

/* ---- batch drivers (same op encoding as oracle/nuts_oracle.h) ------------ */

/* Runs the ops through the reference's functions.  Strings are copied into a
 * NUL-terminated scratch (the batch CSR carries no terminators).  Returns the
 * number of reference calls made. */
int64_t ref_write_batch(int64_t n_ops, const uint8_t *text, const uint64_t *toff,
                        const uint8_t *kind, const int32_t *target,
                        const int32_t *except_user, const uint8_t *oflags,
                        const int32_t *gate, const uint8_t *verdict)
{
    size_t maxn = 0;
    for (int64_t i = 0; i < n_ops; ++i)
        if (toff[i + 1] - toff[i] > maxn) maxn = (size_t)(toff[i + 1] - toff[i]);
    char *str = malloc(maxn + 1);
    if (!str) return -1;
    int64_t calls = 0;
    for (int64_t i = 0; i < n_ops; ++i) {
        if (gate && gate[i] >= 0) {
            int v = verdict[gate[i]] != 0;
            if (((oflags[i] & 8) != 0) != v) continue;
        }
        size_t n = (size_t)(toff[i + 1] - toff[i]);
        memcpy(str, text + toff[i], n); str[n] = 0;
        switch (kind[i]) {
        case 0: ref_write_user(target[i], str); break;
        case 1: ref_write_room_except(target[i], str, except_user[i], oflags[i]); break;
        case 2: ref_write_level(target[i], (oflags[i] & 4) != 0, str, except_user[i]); break;
        }
        ++calls;
    }
    free(str);
    return calls;
}

/* ref_write_batch with the parity digests of SURVEY.md 8(d) folded from what the reference writes: per user
 * (sockets = user indices) in call order, per op over its recipients in the order the reference reached them */
int64_t ref_write_batch_digests(int64_t n_ops, const uint8_t *text, const uint64_t *toff,
                                const uint8_t *kind, const int32_t *target,
                                const int32_t *except_user, const uint8_t *oflags,
                                const int32_t *gate, const uint8_t *verdict, uint64_t *per_user, uint64_t *per_op)
{
    const uint64_t K = 0x9E3779B97F4A7C15ull;
    g_dg_fnv = calloc((size_t)g_nsink + 1, 8); g_dg_len = calloc((size_t)g_nsink + 1, 8);
    g_dg_order = calloc((size_t)g_nsink + 1, sizeof(int)); g_dg_touched = calloc((size_t)g_nsink + 1, 1);
    if (!g_dg_fnv || !g_dg_len || !g_dg_order || !g_dg_touched) return -1;
    for (int u = 0; u < g_nusers; ++u) per_user[u] = 0;
    const int mode = g_sink_mode;
    g_sink_mode = 1; g_dg_on = 1;
    int64_t calls = 0;
    for (int64_t i = 0; i < n_ops; ++i) {
        g_dg_ntouched = 0;
        calls += ref_write_batch(1, text, toff + i, kind + i, target + i, except_user + i, oflags + i, gate ? gate + i : NULL, verdict);
        uint64_t D = 0;
        for (int k = 0; k < g_dg_ntouched; ++k) {
            const int fd = g_dg_order[k];
            const uint64_t d = g_dg_fnv[fd], len = g_dg_len[fd];
            D = (D * K) ^ d ^ len;
            if (fd < g_nusers) per_user[fd] = (per_user[fd] * K) ^ d ^ len;
            g_dg_touched[fd] = 0;
        }
        per_op[i] = D;
    }
    g_dg_on = 0; g_sink_mode = mode;
    free(g_dg_fnv); free(g_dg_len); free(g_dg_order); free(g_dg_touched);
    g_dg_fnv = g_dg_len = NULL; g_dg_order = NULL; g_dg_touched = NULL;
    return calls;
}

void ref_contains_swearing_batch(int64_t n, const uint8_t *text, const uint64_t *off, uint8_t *verdict)
{
    char str[ARR_SIZE * 2 + 1];
    for (int64_t i = 0; i < n; ++i) {
        size_t m = (size_t)(off[i + 1] - off[i]);
        if (m > ARR_SIZE * 2) m = ARR_SIZE * 2;
        memcpy(str, text + off[i], m); str[m] = 0;
        verdict[i] = (uint8_t)contains_swearing(str);
    }
}

void ref_ban_batch(int which, int64_t n, const uint8_t *text, const uint64_t *off, uint8_t *verdict)
{
    char str[256];
    for (int64_t i = 0; i < n; ++i) {
        size_t m = (size_t)(off[i + 1] - off[i]);
        if (m > 255) m = 255;
        memcpy(str, text + off[i], m); str[m] = 0;
        verdict[i] = (uint8_t)(which ? user_banned(str) : site_banned(str));
    }
}
