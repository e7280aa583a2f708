/* tests/dropin/dropin_harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * The drop-in test: the UNMODIFIED reference (nuts333.c, #included from where it lies under
 * /root/reference, never copied) is compiled twice from this one file --
 *
 *   oracle/_ref/libdropin_ref.so   as it is: its own write_user / write_room_except / ...
 *   oracle/_ref/libdropin_shim.so  with -DDROPIN_SHIM: the eight bodies of shim/nuts333_shim.c are
 *                                  compiled in (under the names nutsb_shim_*), the object's eight
 *                                  original symbols are weakened with objcopy and tests/dropin/dropin_alias.c
 *                                  -- eight strong one-line trampolines -- is linked over them, so that
 *                                  every one of the reference's ~600 call sites lands in the shim
 *                                  (tests/dropin/Makefile).  libnutsb200.so is resolved at load time.
 *
 * Both are driven with the same scripted session through the reference's OWN main-loop body and
 * exec_com() (say, .shout, .tell, .emote, .semote, .echo, .bcast, .wizshout, .pemote, .review, .revtell, .go,
 * .colour, .ignall, .look, .who, .help -> more(), .clone, .ban ...) and the bytes every socket received
 * are compared (tests/test_dropin.py).
 *
 * Hooks, so that the talker can run without sockets or a clock: write(2) -> per-fd capture (in the
 * shim build everything queued is flushed first: shim/nuts333_shim.c's -Dwrite=nb_write_through),
 * close(2) -> no-op (+ flush), time(2) -> a settable fake clock.
 */
#include <unistd.h>
#include <stdint.h>
#include <stddef.h>
#include <time.h>
#include <sys/stat.h>
#include <sys/types.h>

static ssize_t dropin_write(int fd, const void *buf, size_t n);
static int     dropin_close(int fd);
static time_t  dropin_time(time_t *t);

#define write    dropin_write
#define close    dropin_close
#define time(x)  dropin_time(x)
#define main     nutsref_main
#include "nuts333.c"
#undef main
#undef time
#undef close
#undef write

/* ---- capture sink -------------------------------------------------------------------------- */
#define FD_BASE 1000
#define FD_MAX  4096
typedef struct { uint8_t *p; size_t n, cap; uint64_t calls; } sink_t;
static sink_t   g_sink[FD_MAX];
static time_t   g_now = 850000000;
static uint64_t g_write_calls = 0;

static ssize_t dropin_capture(int fd, const void *buf, size_t n)
{
    ++g_write_calls;
    if (fd < FD_BASE || fd >= FD_BASE + FD_MAX) return (ssize_t)n;
    sink_t *s = &g_sink[fd - FD_BASE];
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
static time_t dropin_time(time_t *t) { if (t) *t = g_now; return g_now; }

#ifdef DROPIN_SHIM
#define NB_NAME(f) nutsb_shim_##f
#define NB_SOCK_WRITE(fd, buf, n) dropin_capture((fd), (buf), (n))
#define NB_SOCK_CLOSE(fd) 0
#include "../../shim/nuts333_shim.c"
/* what -Dwrite=nb_write_through -Dclose=nb_close_through do in a real build */
static ssize_t dropin_write(int fd, const void *buf, size_t n) { return (ssize_t)nb_write_through(fd, buf, n); }
static int dropin_close(int fd) { return nb_close_through(fd); }
#else
static ssize_t dropin_write(int fd, const void *buf, size_t n) { return dropin_capture(fd, buf, n); }
static int dropin_close(int fd) { (void)fd; return 0; }
#endif

/* ---- session set-up -------------------------------------------------------------------------- */
#define MAX_H 4096
static UR_OBJECT g_users[MAX_H]; static int g_nusers = 0;
static int g_user_fd[MAX_H];       /* the socket a handle was given: never reused, so it tells a user from a later one at the same address */
static RM_OBJECT g_rooms[256];   static int g_nrooms = 0;
static NL_OBJECT g_links[64];    static int g_nlinks = 0;
static int g_next_fd = FD_BASE;

int dropin_is_shim(void)
{
#ifdef DROPIN_SHIM
    return 1;
#else
    return 0;
#endif
}

/* scratch directory with datafiles/ helpfiles/ userfiles/ in it: the talker's CWD */
int dropin_reset(const char *dir, int device, int use_iov)
{
    int i;
    if (chdir(dir) != 0) return -1;
#ifdef DROPIN_SHIM
    nb_shutdown();
#endif
    while (user_first) destruct_user(user_first);
    for (i = 0; i < g_nrooms; ++i) free(g_rooms[i]);
    for (i = 0; i < g_nlinks; ++i) free(g_links[i]);
    room_first = room_last = NULL; nl_first = nl_last = NULL;
    for (i = 0; i < FD_MAX; ++i) { free(g_sink[i].p); memset(&g_sink[i], 0, sizeof g_sink[i]); }
    g_nusers = g_nrooms = g_nlinks = 0; g_next_fd = FD_BASE; g_now = 850000000; g_write_calls = 0;
    init_globals();                    /* c:1032 */
    system_logging = 0;                /* keep write_syslog() away from the CWD */
    set_date_time();
    force_listen = 0; com_num = -1; no_prompt = 0; destructed = 0;
#ifdef DROPIN_SHIM
    if (nb_init(device) != 0) return -2;
    nb_set_iov(use_iov);
    nb_errors = 0;
#else
    (void)device; (void)use_iov;
#endif
    return 0;
}

int dropin_errors(void)
{
#ifdef DROPIN_SHIM
    return nb_errors;
#else
    return 0;
#endif
}
const char *dropin_last_error(void)
{
#ifdef DROPIN_SHIM
    return nb_last_error;
#else
    return "";
#endif
}
/* [0] flushes [1] ops queued [2] population uploads [3] socket writes made by the shim */
void dropin_stats(uint64_t *out)
{
#ifdef DROPIN_SHIM
    out[0] = nb_stat_flushes; out[1] = nb_stat_ops; out[2] = nb_stat_syncs; out[3] = nb_stat_sock_writes;
#else
    out[0] = out[1] = out[2] = 0; out[3] = g_write_calls;
#endif
}

void dropin_flush(void)
{
#ifdef DROPIN_SHIM
    nb_flush_to_sockets();
#endif
}

int dropin_add_room(const char *name, const char *label, const char *desc, int access)
{
    RM_OBJECT r = create_room();       /* c:2776 */
    if (!r || g_nrooms >= 256) return -1;
    strncpy(r->name, name, ROOM_NAME_LEN); r->name[ROOM_NAME_LEN] = 0;
    strncpy(r->label, label, ROOM_LABEL_LEN); r->label[ROOM_LABEL_LEN] = 0;
    strncpy(r->desc, desc, ROOM_DESC_LEN); r->desc[ROOM_DESC_LEN] = 0;
    r->access = access;
    g_rooms[g_nrooms] = r;
    return g_nrooms++;
}
/* a <-> b */
void dropin_link_rooms(int a, int b)
{
    int i;
    if (a < 0 || b < 0 || a >= g_nrooms || b >= g_nrooms) return;
    for (i = 0; i < MAX_LINKS; ++i) if (!g_rooms[a]->link[i]) { g_rooms[a]->link[i] = g_rooms[b]; break; }
    for (i = 0; i < MAX_LINKS; ++i) if (!g_rooms[b]->link[i]) { g_rooms[b]->link[i] = g_rooms[a]; break; }
}

/* a netlink object up and verified, as connect_to_site / accept_server_connection leave it */
int dropin_add_netlink(const char *service, int room, int ver_minor)
{
    NL_OBJECT nl = create_netlink();
    if (!nl || g_nlinks >= 64) return -1;
    strncpy(nl->service, service, SERV_NAME_LEN);
    nl->socket = g_next_fd++; nl->type = OUTGOING; nl->stage = UP; nl->connected = 1; nl->allow = ALL;
    nl->ver_major = 3; nl->ver_minor = ver_minor; nl->ver_patch = 0;
    nl->connect_room = room >= 0 ? g_rooms[room] : NULL;
    nl->last_recvd = g_now;
    if (room >= 0) g_rooms[room]->netlink = nl;
    g_links[g_nlinks] = nl;
    return g_nlinks++;
}

/* a logged-in user as connect_user (c:1677) leaves one; returns the handle (the socket is FD_BASE + handle
 * order of creation).  login != 0: still at a login stage. */
int dropin_add_user(const char *name, int room, int level, int colour, int login, int prompt_on, int command_mode)
{
    UR_OBJECT u = create_user();       /* c:2673 */
    if (!u || g_nusers >= MAX_H) return -1;
    strncpy(u->name, name, USER_NAME_LEN); u->name[USER_NAME_LEN] = 0;
    strcpy(u->desc, "is a test user");
    strcpy(u->in_phrase, "enters"); strcpy(u->out_phrase, "goes");
    strcpy(u->site, "test.site"); strcpy(u->last_site, "test.site");
    u->socket = g_next_fd++;
    u->room = room >= 0 ? g_rooms[room] : NULL;
    u->level = level; u->colour = colour; u->login = login; u->prompt = prompt_on; u->command_mode = command_mode;
    u->last_login = g_now - 3600; u->last_input = g_now;
    if (!login) num_of_users++; else num_of_logins++;
    g_users[g_nusers] = u; g_user_fd[g_nusers] = u->socket;
    return g_nusers++;
}

/* a user who came over a netlink (nl_transfer, c:3077-3165): REMOTE_TYPE, no socket of its own */
int dropin_add_remote_user(const char *name, int room, int level, int link)
{
    const int h = dropin_add_user(name, room, level, 1, 0, 0, 0);
    if (h < 0 || link < 0 || link >= g_nlinks) return -1;
    --g_next_fd;
    g_users[h]->type = REMOTE_TYPE; g_users[h]->socket = -1; g_users[h]->netlink = g_links[link]; g_user_fd[h] = -1;
    return h;
}

static int dropin_alive(int h)
{
    UR_OBJECT u;
    if (h < 0 || h >= g_nusers || !g_users[h]) return 0;
    for (u = user_first; u; u = u->next) if (u == g_users[h] && u->socket == g_user_fd[h] && u->type != CLONE_TYPE) return 1;
    g_users[h] = NULL;
    return 0;
}
int dropin_user_alive(int h) { return dropin_alive(h); }
int dropin_user_fd(int h) { return dropin_alive(h) ? g_users[h]->socket : -1; }
int dropin_link_fd(int l) { return (l >= 0 && l < g_nlinks) ? g_links[l]->socket : -1; }
int dropin_user_room(int h)
{
    int i;
    if (!dropin_alive(h) || !g_users[h]->room) return -1;
    for (i = 0; i < g_nrooms; ++i) if (g_rooms[i] == g_users[h]->room) return i;
    return -1;
}
/* direct edits of the fields the path reads, for what no command sets (0 colour 1 ignall 2 ignshout 3 vis 4 muzzled 5 level 6 login) */
void dropin_set_field(int h, int field, int value)
{
    if (!dropin_alive(h)) return;
    switch (field) {
    case 0: g_users[h]->colour = value; break;   case 1: g_users[h]->ignall = value; break;
    case 2: g_users[h]->ignshout = value; break; case 3: g_users[h]->vis = value; break;
    case 4: g_users[h]->muzzled = value; break;  case 5: g_users[h]->level = value; break;
    case 6: g_users[h]->login = value; break;
    }
}
void dropin_set_globals(int ban_swearing_, int now_delta) { ban_swearing = ban_swearing_; g_now += now_delta; set_date_time(); }

#ifdef NUTSREF_MAX_SWEAR
int dropin_set_swear_words(const char *const *words)
{
    static char *own[NUTSREF_MAX_SWEAR];
    int n = 0, i;
    for (i = 0; i < NUTSREF_MAX_SWEAR; ++i) { free(own[i]); own[i] = NULL; }
    while (words && words[n]) {
        if (n >= NUTSREF_MAX_SWEAR - 1) return -1;
        own[n] = strdup(words[n]); swear_words[n] = own[n]; ++n;
    }
    own[n] = strdup("*"); swear_words[n] = own[n];
#ifdef DROPIN_SHIM
    nb_reload_swear_words();
#endif
    return n;
}
#endif

/* ---- one input line, as the main loop handles it from GOT_LINE on (c:150-234) ---------------------- */
int dropin_input(int h, const char *line)
{
    static char inpstr[ARR_SIZE];
    UR_OBJECT user;
    if (!dropin_alive(h)) return -1;
    user = g_users[h];
    if (user->type != USER_TYPE) return -1;                        /* c:129 */
    strncpy(inpstr, line, ARR_SIZE - 1); inpstr[ARR_SIZE - 1] = 0;
    terminate(inpstr);                                             /* c:149 */
    no_prompt = 0; com_num = -1; force_listen = 0; destructed = 0; /* c:151-154 */
    user->buff[0] = '\0'; user->buffpos = 0; user->last_input = time(0);
    if (user->login) { login(user, inpstr); return 0; }            /* c:158-160 */
    if (!user->misc_op) {                                          /* c:164-174 */
        if (!strcmp(inpstr, ".") && user->inpstr_old[0]) {
            strcpy(inpstr, user->inpstr_old);
            sprintf(text, "%s\n", inpstr);
            write_user(user, text);
        } else if (inpstr[0]) strncpy(user->inpstr_old, inpstr, REVIEW_LEN);
    }
    clear_words();                                                 /* c:177-178 */
    word_count = wordfind(inpstr);
    if (user->afk) {                                               /* c:179-203, without the session lock */
        write_user(user, "You are no longer AFK.\n");
        user->afk_mesg[0] = '\0';
        if (user->vis) { sprintf(text, "%s comes back from being AFK.\n", user->name); write_room_except(user->room, text, user); }
        user->afk = 0;
    }
    if (!word_count) {                                             /* c:204-212 */
        if (misc_ops(user, inpstr)) return 0;
        if (user->command_mode) prompt(user);
        return 0;
    }
    if (misc_ops(user, inpstr)) return 0;                          /* c:213 */
    com_num = -1;
    if (user->command_mode || strchr(".;!<>-#", inpstr[0])) exec_com(user, inpstr);   /* c:215-217 */
    else say(user, inpstr);
    /* c:218-233.  disconnect_user() ends with destructed=0 (c:1809), so after .quit the reference's main loop calls
     * prompt() on the freed user object and writes to the closed socket: unobservable there, undefined here --
     * the harness asks the user list instead */
    if (!destructed && dropin_alive(h)) {
        if (user->room != NULL) prompt(user);
        else switch ((int)com_num) {
            case -1: case HOME: case QUIT: case MODE: case PROMPT: case SUICIDE: case REBOOT: case SHUTDOWN: prompt(user); break;
            default: break;
        }
    }
    return 0;
}

/* the admission calls: accept_connection's site check (c:278-285) ... */
int dropin_site_banned(const char *site) { return site_banned((char *)site); }
int dropin_user_banned(const char *name) { return user_banned((char *)name); }
int dropin_contains_swearing(const char *s) { return contains_swearing((char *)s); }
/* ... and login()'s name stage (c:1462-1520) on a fresh connection: returns the fd the answer went to */
int dropin_login_attempt(const char *name)
{
    static char line[ARR_SIZE];
    UR_OBJECT u = create_user(), v;
    int fd;
    if (!u) return -1;
    fd = u->socket = g_next_fd++;
    u->login = 3; num_of_logins++;
    strcpy(u->site, "new.site");
    strncpy(line, name, ARR_SIZE - 1); line[ARR_SIZE - 1] = 0;
    no_prompt = 0; com_num = -1; force_listen = 0; destructed = 0;
    login(u, line);
    for (v = user_first; v; v = v->next) if (v == u) break;
    if (v) { dropin_close(u->socket); destruct_user(u); num_of_logins--; }      /* not refused: hang up */
    return fd;
}
/* the heartbeat's entries into the write layer (check_idle_and_timeout c:7770, reboot / shutdown countdown c:7741) */
void dropin_events(int now_delta) { g_now += now_delta; set_date_time(); check_reboot_shutdown(); check_idle_and_timeout(); }

/* ---- stream access ------------------------------------------------------------------------- */
size_t dropin_stream_len(int fd) { return (fd >= FD_BASE && fd < FD_BASE + FD_MAX) ? g_sink[fd - FD_BASE].n : 0; }
const uint8_t *dropin_stream_ptr(int fd) { return (fd >= FD_BASE && fd < FD_BASE + FD_MAX) ? g_sink[fd - FD_BASE].p : NULL; }
uint64_t dropin_stream_calls(int fd) { return (fd >= FD_BASE && fd < FD_BASE + FD_MAX) ? g_sink[fd - FD_BASE].calls : 0; }
int dropin_fd_base(void) { return FD_BASE; }
int dropin_next_fd(void) { return g_next_fd; }
