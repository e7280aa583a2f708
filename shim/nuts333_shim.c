/* shim/nuts333_shim.c -- the drop-in: nuts333.c's write layer on libnutsb200.so.
 *
 * The reference reaches its write path by direct C calls inside one translation unit; the
 * boundary is the eight functions it declares in nuts333.h:307, 313, 324-325, 478-483.  This
 * file holds replacement BODIES for them with the reference's exact names and signatures:
 *
 *     void write_sock(int sock, char *str);                              nuts333.c:1281
 *     void write_user(UR_OBJECT user, char *str);                        nuts333.c:1291
 *     void write_level(int level, int above, char *str, UR_OBJECT user); nuts333.c:1372
 *     void write_room(RM_OBJECT rm, char *str);                          nuts333.c:1390
 *     void write_room_except(RM_OBJECT rm, char *str, UR_OBJECT user);   nuts333.c:1401
 *     int  contains_swearing(char *str);                                 nuts333.c:2540
 *     int  site_banned(char *site);                                      nuts333.c:330
 *     int  user_banned(char *name);                                      nuts333.c:349
 *
 * It is C, written against the reference's own types and globals (UR_OBJECT, RM_OBJECT,
 * NL_OBJECT, user_first, room_first, nl_first, force_listen, com_num, swear_words ...), so it
 * is compiled IN the translation unit of nuts333.c: a maintainer deletes the eight original
 * bodies and adds `#include "nuts333_shim.c"` at the end of nuts333.c (nuts333.h defines its
 * globals and has no include guard, so a second translation unit cannot include it).  The
 * test harness (tests/dropin/) does the same to the UNMODIFIED file: it compiles nuts333.c as
 * it is, weakens the eight symbols in the object (objcopy) and links these bodies over them.
 *
 * What the talker gains: every write_user / write_room / write_level call of one main-loop
 * iteration (or of many) is queued and rendered in ONE device batch; the host then makes one
 * write(2) (or writev(2)) per socket.  What it must do in exchange -- three lines:
 *
 *     main():                 nb_init(0);                       after init_globals()
 *     main loop, do_events(): nb_flush_to_sockets();            once per iteration (c:235, c:7721)
 *     compile with            -Dwrite=nb_write_through -Dclose=nb_close_through
 *
 * The last line keeps the byte order on every socket: code that still write()s by itself
 * (more() c:2205, the telnet option strings) first flushes what is queued, and a socket is
 * flushed before it is closed (disconnect_user c:1771).
 *
 * Users and rooms cross the C-ABI as their index in the reference's lists.  The shim keeps a
 * snapshot of the fields the path reads (nuts333.h:67-85: room, type, login, ignall, ignshout,
 * colour, level, socket, owner, clone_hear, netlink) and checks it against the lists at every
 * call -- the reference evaluates its filters at call time, so the moment anything differs the
 * queue is flushed under the old population and the new one goes to the device.  The check
 * walks the user list, which is what the reference's own write_room_except does per call
 * (c:1409); a talker that wants it cheaper sets nb_dirty where it changes those fields.
 *
 * Errors never reach the caller (the reference ignores write(2)'s result, c:1285): a failed
 * call is counted in nb_errors and the last message kept in nb_last_error.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "nutsb200.h"

#ifndef NB_NAME                      /* the harness links the bodies under other names (see above) */
#define NB_NAME(f) f
#endif
#ifndef NB_SOCK_WRITE                /* the write(2) that reaches the socket */
#ifdef write
#undef write
#endif
#include <unistd.h>
#define NB_SOCK_WRITE(fd, buf, n) write((fd), (buf), (n))
#endif
#ifndef NB_SOCK_CLOSE
#ifdef close
#undef close
#endif
#define NB_SOCK_CLOSE(fd) close(fd)
#endif

static nutsb_ctx *nb;
static int  nb_use_iov = 0;          /* nb_flush_to_sockets: gather lists (nutsb_flush_iov) instead of streams */
static int  nb_in_flush = 0;
int         nb_errors = 0;
char        nb_last_error[256];
/* counters a test (or a curious maintainer) can read */
uint64_t    nb_stat_flushes = 0, nb_stat_ops = 0, nb_stat_syncs = 0, nb_stat_sock_writes = 0;

static void nb_fail(const char *what, int rc)
{
    ++nb_errors;
    snprintf(nb_last_error, sizeof nb_last_error, "%s: %s (%s)", what, nutsb_strerror(rc), nb ? nutsb_last_error(nb) : "");
}
#define NB_TRY(call) do { int rc_ = (call); if (rc_ < 0) nb_fail(#call, rc_); } while (0)

/* ---- the population snapshot ------------------------------------------------------------ */
struct nb_urec {
    UR_OBJECT u; RM_OBJECT room; UR_OBJECT owner; NL_OBJECT nl;
    int type, level, socket, clone_hear, flags;
    char name[USER_NAME_LEN + 1];
};
struct nb_lrec { NL_OBJECT nl; int socket, old_peer; };
static struct nb_urec *nb_u;  static int nb_nu, nb_cap_u;
static RM_OBJECT      *nb_r;  static int nb_nr, nb_cap_r;
static char          (*nb_rname)[ROOM_NAME_LEN + 1];
static struct nb_lrec *nb_l;  static int nb_nl, nb_cap_l;     /* netlinks: pseudo-users nb_nu .. nb_nu+nb_nl-1 */
static int nb_valid = 0;
int nb_dirty = 0;                    /* a talker may set this instead of relying on the per-call check */

/* pointer -> index, open addressing */
static struct { const void *k; int v; } *nb_h; static unsigned nb_hmask;
static unsigned nb_hash(const void *p) { return (unsigned)(((uintptr_t)p >> 4) * 0x9E3779B1u); }
static int nb_user_index(UR_OBJECT u)
{
    unsigned i;
    if (!u || !nb_h) return -1;
    for (i = nb_hash(u) & nb_hmask; nb_h[i].k; i = (i + 1) & nb_hmask) if (nb_h[i].k == u) return nb_h[i].v;
    return -1;
}
static int nb_room_index(RM_OBJECT rm)
{
    int i;
    for (i = 0; i < nb_nr; ++i) if (nb_r[i] == rm) return i;
    return -1;
}
static int nb_link_index(NL_OBJECT nl)
{
    int i;
    for (i = 0; i < nb_nl; ++i) if (nb_l[i].nl == nl) return i;
    return -1;
}

static int nb_user_flags(UR_OBJECT u)
{
    return (u->colour ? NUTSB_UF_COLOUR : 0) | (u->login ? NUTSB_UF_LOGIN : 0)
         | (u->ignall ? NUTSB_UF_IGNALL : 0) | (u->ignshout ? NUTSB_UF_IGNSHOUT : 0)
         | (u->type == CLONE_TYPE ? NUTSB_UF_CLONE : 0) | (u->type == REMOTE_TYPE && u->netlink ? NUTSB_UF_REMOTE : 0);
}
static int nb_old_peer(NL_OBJECT nl) { return nl->ver_major <= 3 && nl->ver_minor < 2; }      /* c:1300 */

/* Is the snapshot still what the lists say?  (the fields of nuts333.h:67-85 the path reads) */
static int nb_population_current(void)
{
    UR_OBJECT u; RM_OBJECT rm; NL_OBJECT nl; int i;
    if (!nb_valid || nb_dirty) return 0;
    for (u = user_first, i = 0; u != NULL; u = u->next, ++i) {
        const struct nb_urec *r;
        if (i >= nb_nu) return 0;
        r = &nb_u[i];
        if (r->u != u || r->room != u->room || r->type != u->type || r->level != u->level || r->socket != u->socket
            || r->flags != nb_user_flags(u) || r->owner != u->owner || r->clone_hear != u->clone_hear || r->nl != u->netlink
            || strcmp(r->name, u->name)) return 0;
    }
    if (i != nb_nu) return 0;
    for (rm = room_first, i = 0; rm != NULL; rm = rm->next, ++i)
        if (i >= nb_nr || nb_r[i] != rm || strcmp(nb_rname[i], rm->name)) return 0;
    if (i != nb_nr) return 0;
    for (nl = nl_first, i = 0; nl != NULL; nl = nl->next, ++i)
        if (i >= nb_nl || nb_l[i].nl != nl || nb_l[i].socket != nl->socket || nb_l[i].old_peer != nb_old_peer(nl)) return 0;
    return i == nb_nl;
}

void nb_flush_to_sockets(void);

/* The lists -> SoA arrays -> nutsb_set_users & co.  Called with an empty queue. */
static void nb_sync_population(void)
{
    UR_OBJECT u; RM_OBJECT rm; NL_OBJECT nl;
    int i, n = 0, nr = 0, nl_n = 0, tot;
    int32_t *room, *prev, *owner, *link; uint8_t *flags, *level, *hear, *oldp, *sfl, *names; uint64_t *noff;
    unsigned hs;
    for (u = user_first; u != NULL; u = u->next) ++n;
    for (rm = room_first; rm != NULL; rm = rm->next) ++nr;
    for (nl = nl_first; nl != NULL; nl = nl->next) ++nl_n;
    tot = n + nl_n;
    room = malloc(sizeof *room * (size_t)(tot + 1)); prev = malloc(sizeof *prev * (size_t)(tot + 1));
    owner = malloc(sizeof *owner * (size_t)(tot + 1)); link = malloc(sizeof *link * (size_t)(tot + 1));
    flags = malloc((size_t)tot + 1); level = malloc((size_t)tot + 1); hear = malloc((size_t)tot + 1);
    oldp = malloc((size_t)tot + 1); sfl = calloc((size_t)tot + 1, 1);
    names = malloc((size_t)(tot + 1) * (USER_NAME_LEN + 1)); noff = malloc(sizeof *noff * (size_t)(tot + 2));
    if (!room || !prev || !owner || !link || !flags || !level || !hear || !oldp || !sfl || !names || !noff) { nb_fail("nb_sync_population", NUTSB_E_NOMEM); goto out; }

    /* where everybody was before (review / revtell buffers follow their owner: nutsb_set_users_remap) */
    for (u = user_first, i = 0; u != NULL; u = u->next, ++i) prev[i] = nb_user_index(u);
    for (nl = nl_first; nl != NULL; nl = nl->next, ++i) { const int k = nb_link_index(nl); prev[i] = k < 0 ? -1 : nb_nu + k; }

    /* rooms */
    if (nr > nb_cap_r) { nb_cap_r = nr + 16; nb_r = realloc(nb_r, sizeof *nb_r * (size_t)nb_cap_r); nb_rname = realloc(nb_rname, sizeof *nb_rname * (size_t)nb_cap_r); }
    for (rm = room_first, i = 0; rm != NULL; rm = rm->next, ++i) { nb_r[i] = rm; strcpy(nb_rname[i], rm->name); }
    nb_nr = nr;
    /* netlinks */
    if (nl_n > nb_cap_l) { nb_cap_l = nl_n + 8; nb_l = realloc(nb_l, sizeof *nb_l * (size_t)nb_cap_l); }
    for (nl = nl_first, i = 0; nl != NULL; nl = nl->next, ++i) { nb_l[i].nl = nl; nb_l[i].socket = nl->socket; nb_l[i].old_peer = nb_old_peer(nl); }
    nb_nl = nl_n;
    /* users */
    if (n > nb_cap_u) { nb_cap_u = n + n / 2 + 64; nb_u = realloc(nb_u, sizeof *nb_u * (size_t)nb_cap_u); }
    for (hs = 16; hs < 2u * (unsigned)n + 2u; hs <<= 1) ;
    free(nb_h); nb_h = calloc(hs, sizeof *nb_h); nb_hmask = hs - 1;
    for (u = user_first, i = 0; u != NULL; u = u->next, ++i) {
        struct nb_urec *r = &nb_u[i]; unsigned h;
        r->u = u; r->room = u->room; r->owner = u->owner; r->nl = u->netlink; r->type = u->type; r->level = u->level;
        r->socket = u->socket; r->clone_hear = u->clone_hear; r->flags = nb_user_flags(u);
        strncpy(r->name, u->name, USER_NAME_LEN); r->name[USER_NAME_LEN] = 0;
        for (h = nb_hash(u) & nb_hmask; nb_h[h].k; h = (h + 1) & nb_hmask) ;
        nb_h[h].k = u; nb_h[h].v = i;
    }
    nb_nu = n;
    noff[0] = 0;
    for (i = 0; i < n; ++i) {
        const struct nb_urec *r = &nb_u[i];
        const size_t ln = strlen(r->name);
        room[i] = r->room ? nb_room_index(r->room) : -1;
        flags[i] = (uint8_t)r->flags; level[i] = (uint8_t)r->level;
        owner[i] = r->type == CLONE_TYPE ? nb_user_index(r->owner) : -1;
        hear[i] = (uint8_t)(r->type == CLONE_TYPE ? r->clone_hear : 0);
        link[i] = (r->flags & NUTSB_UF_REMOTE) ? n + nb_link_index(r->nl) : -1;
        oldp[i] = (uint8_t)((r->flags & NUTSB_UF_REMOTE) ? nb_old_peer(r->nl) : 0);
        memcpy(names + noff[i], r->name, ln); noff[i + 1] = noff[i] + ln;
    }
    /* one pseudo-user per netlink: in no room, at a login stage (nothing reaches it by itself); its stream
       is what goes to the link's socket -- the MSG/EMSG frames (c:1299-1307) and write_sock's own lines */
    for (i = n; i < tot; ++i) {
        room[i] = -1; flags[i] = NUTSB_UF_LOGIN; level[i] = 0; owner[i] = -1; hear[i] = 0; link[i] = -1; oldp[i] = 0;
        noff[i + 1] = noff[i];
    }
    NB_TRY(nutsb_set_users_remap(nb, tot, nr, room, flags, level, prev));
    NB_TRY(nutsb_set_user_names(nb, tot, names, noff, sfl));
    {
        uint8_t *rn = malloc((size_t)(nr + 1) * (ROOM_NAME_LEN + 1)); uint64_t *ro = malloc(sizeof *ro * (size_t)(nr + 2));
        if (rn && ro) {
            ro[0] = 0;
            for (i = 0; i < nr; ++i) { const size_t ln = strlen(nb_rname[i]); memcpy(rn + ro[i], nb_rname[i], ln); ro[i + 1] = ro[i] + ln; }
            NB_TRY(nutsb_set_room_names(nb, nr, rn, ro));
        }
        free(rn); free(ro);
    }
    NB_TRY(nutsb_set_clones(nb, tot, owner, hear));
    NB_TRY(nutsb_set_remotes(nb, tot, link, oldp));
    nb_valid = 1; nb_dirty = 0; ++nb_stat_syncs;
out:
    free(room); free(prev); free(owner); free(link); free(flags); free(level); free(hear); free(oldp); free(sfl); free(names); free(noff);
}

/* every entry point: the device knows the population the reference would filter against right now */
static int nb_ready(void)
{
    if (!nb) return 0;
    if (!nb_population_current()) {
        nb_flush_to_sockets();                 /* what is queued was said under the old population */
        nb_sync_population();
    }
    return nb_valid;
}

/* ---- lifecycle ------------------------------------------------------------------------------ */
static char *nb_banfile[2]; static long nb_banlen[2] = { -2, -2 };     /* the ban files the device holds; -1: no file, -2: none yet */
void nb_reload_swear_words(void) { if (nb) NB_TRY(nutsb_set_swear_words(nb, (const char *const *)swear_words)); }   /* h:275-277 */

int nb_init(int device)
{
    const int rc = nutsb_create(&nb, device);
    if (rc < 0) { nb = NULL; nb_fail("nutsb_create", rc); return rc; }
    nb_reload_swear_words();
    nb_valid = 0; nb_banlen[0] = nb_banlen[1] = -2;
    return 0;
}
void nb_set_iov(int on) { nb_use_iov = on != 0; }
void nb_shutdown(void)
{
    if (!nb) return;
    nb_flush_to_sockets();
    nutsb_destroy(nb); nb = NULL; nb_valid = 0;
    free(nb_u); nb_u = NULL; nb_nu = nb_cap_u = 0; free(nb_r); nb_r = NULL; free(nb_rname); nb_rname = NULL; nb_nr = nb_cap_r = 0;
    free(nb_l); nb_l = NULL; nb_nl = nb_cap_l = 0; free(nb_h); nb_h = NULL;
}

/* ---- the sink: what the reference's write(2) calls were (c:1318, 1339, 1360, 1363, 1365) ------ */
void nb_flush_to_sockets(void)
{
    int i;
    if (!nb || nb_in_flush || nutsb_q_pending(nb) == 0) return;
    nb_in_flush = 1;
    ++nb_stat_flushes;
    if (nb_use_iov) {
        nutsb_iov_streams s;
        const int rc = nutsb_flush_iov(nb, &s);
        if (rc < 0) nb_fail("nutsb_flush_iov", rc);
        else for (i = 0; i < nb_nu + nb_nl; ++i) {
            const int sock = i < nb_nu ? nb_u[i].socket : nb_l[i - nb_nu].socket;
            const nutsb_iovec *v = s.iov + s.first[i]; uint32_t k;
            if (s.off[i + 1] == s.off[i]) continue;
            /* (a host with real sockets hands v[0 .. count) to writev(2) in chunks of IOV_MAX) */
            for (k = 0; k < s.count[i]; ++k) if (v[k].len) { NB_SOCK_WRITE(sock, v[k].base, v[k].len); ++nb_stat_sock_writes; }
        }
    } else {
        nutsb_streams s;
        const int rc = nutsb_flush(nb, &s);
        if (rc < 0) nb_fail("nutsb_flush", rc);
        else for (i = 0; i < nb_nu + nb_nl; ++i) {
            const int sock = i < nb_nu ? nb_u[i].socket : nb_l[i - nb_nu].socket;
            if (s.off[i + 1] > s.off[i]) { NB_SOCK_WRITE(sock, s.bytes + s.off[i], (size_t)(s.off[i + 1] - s.off[i])); ++nb_stat_sock_writes; }
        }
    }
    nb_in_flush = 0;
}

/* for code that still writes by itself (compile nuts333.c with -Dwrite=nb_write_through -Dclose=nb_close_through) */
long nb_write_through(int fd, const void *buf, size_t n) { nb_flush_to_sockets(); return (long)NB_SOCK_WRITE(fd, buf, n); }
int  nb_close_through(int fd) { nb_flush_to_sockets(); return NB_SOCK_CLOSE(fd); }

/* ---- the eight bodies ----------------------------------------------------------------------- */

/*** Write a NULL terminated string to a socket ***/
void NB_NAME(write_sock)(int sock, char *str)                          /* c:1281 */
{
    int i;
    if (nb_ready()) {
        for (i = 0; i < nb_nu; ++i)
            if (nb_u[i].socket == sock && nb_u[i].type == USER_TYPE) { ++nb_stat_ops; NB_TRY(nutsb_q_write_sock(nb, i, str)); return; }
        for (i = 0; i < nb_nl; ++i)
            if (nb_l[i].socket == sock) { ++nb_stat_ops; NB_TRY(nutsb_q_write_sock(nb, nb_nu + i, str)); return; }
    }
    nb_flush_to_sockets();                                             /* a socket that is nobody's yet (c:280) */
    NB_SOCK_WRITE(sock, str, strlen(str));
}

/*** Send message to user ***/
void NB_NAME(write_user)(UR_OBJECT user, char *str)                    /* c:1291 */
{
    if (user == NULL || !nb_ready()) return;                           /* c:1298 */
    ++nb_stat_ops;
    NB_TRY(nutsb_q_write_user(nb, nb_user_index(user), str));
}

/*** Write to users of level 'level' and above or below depending on above variable ***/
void NB_NAME(write_level)(int level, int above, char *str, UR_OBJECT user)     /* c:1372 */
{
    if (!nb_ready()) return;
    ++nb_stat_ops;
    NB_TRY(nutsb_q_write_level(nb, level, above, str, nb_user_index(user)));
}

/*** Write to everyone in room rm except for "user"; rm==NULL: every room ***/
void NB_NAME(write_room_except)(RM_OBJECT rm, char *str, UR_OBJECT user)       /* c:1401 */
{
    int r = -1;
    if (!nb_ready()) return;
    if (rm != NULL && (r = nb_room_index(rm)) < 0) return;             /* a room that is in no list has nobody in it */
    ++nb_stat_ops;
    /* the two globals the reference reads INSIDE the call (c:1413, c:1414) */
    NB_TRY(nutsb_q_write_room_except(nb, r, str, nb_user_index(user), force_listen, com_num == SHOUT || com_num == SEMOTE));
}

/*** Subsid function to below but this one is used the most ***/
void NB_NAME(write_room)(RM_OBJECT rm, char *str)                      /* c:1390 */
{
    NB_NAME(write_room_except)(rm, str, NULL);
}

/*** See if string contains any swearing ***/
int NB_NAME(contains_swearing)(char *str)                              /* c:2540 */
{
    int v;
    if (!nb) return 0;
    v = nutsb_contains_swearing(nb, str);
    if (v < 0) { nb_fail("nutsb_contains_swearing", v); return 0; }    /* c:2546-2549: a failure reads as clean */
    return v;
}

/* datafiles/siteban | userban as they stand on disk (the reference re-reads them per query, c:336, c:355) */
static void nb_refresh_ban_files(void)
{
    static const char *fname[2] = { SITEBAN, USERBAN };
    char *cur[2] = { NULL, NULL }; long len[2]; int w, changed = 0;
    for (w = 0; w < 2; ++w) {
        char path[256]; FILE *fp;
        snprintf(path, sizeof path, "%s/%s", DATAFILES, fname[w]);
        len[w] = -1;
        if ((fp = fopen(path, "rb")) != NULL) {
            long cap = 4096, n = 0; size_t got;
            cur[w] = malloc((size_t)cap);
            while (cur[w] && (got = fread(cur[w] + n, 1, (size_t)(cap - n), fp)) > 0) {
                n += (long)got;
                if (n == cap) { cap *= 2; cur[w] = realloc(cur[w], (size_t)cap); }
            }
            fclose(fp);
            len[w] = cur[w] ? n : -1;
        }
        if (len[w] != nb_banlen[w] || (len[w] > 0 && memcmp(cur[w], nb_banfile[w], (size_t)len[w]))) changed = 1;
    }
    if (changed) {
        static const char none = 0;
        for (w = 0; w < 2; ++w) { free(nb_banfile[w]); nb_banfile[w] = cur[w]; nb_banlen[w] = len[w]; }
        NB_TRY(nutsb_set_ban_files(nb, len[0] < 0 ? NULL : (len[0] ? nb_banfile[0] : &none), len[0] < 0 ? 0 : (size_t)len[0],
                                       len[1] < 0 ? NULL : (len[1] ? nb_banfile[1] : &none), len[1] < 0 ? 0 : (size_t)len[1]));
    } else { free(cur[0]); free(cur[1]); }
}

/*** See if users site is banned ***/
int NB_NAME(site_banned)(char *site)                                   /* c:330 */
{
    int v;
    if (!nb) return 0;
    nb_refresh_ban_files();
    v = nutsb_site_banned(nb, site);
    if (v < 0) { nb_fail("nutsb_site_banned", v); return 0; }          /* c:337: no list, no ban */
    return v;
}

/*** See if user is banned ***/
int NB_NAME(user_banned)(char *name)                                   /* c:349 */
{
    int v;
    if (!nb) return 0;
    nb_refresh_ban_files();
    v = nutsb_user_banned(nb, name);
    if (v < 0) { nb_fail("nutsb_user_banned", v); return 0; }
    return v;
}
