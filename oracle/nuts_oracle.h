/* oracle/nuts_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the NUTS 3.3.3 write path (the "port" oracle) and the
 * batch conventions shared with the reference harness (oracle/ref_harness.c,
 * which drives the UNMODIFIED nuts333.c).  Nothing under oracle/ is part of the
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load these libraries, and only as the checker or
 * as the reported CPU baseline.
 *
 * Parity pin: every function here is checked against tests/golden/ (vectors
 * minted from the unmodified reference, see tests/golden/make_golden.py) and,
 * when oracle/_ref/libnutsref.so is present, differentially against it.
 *
 * Citations: c: = /root/reference/nuts333.c, h: = /root/reference/nuts333.h
 */
#ifndef NUTS_ORACLE_H
#define NUTS_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Recipient flag bits -- identical to include/nutsb200.h (NUTSB_UF_*). */
#define ORC_UF_COLOUR   0x01u  /* user->colour != 0          h:80     */
#define ORC_UF_LOGIN    0x02u  /* user->login  != 0          c:1410   */
#define ORC_UF_IGNALL   0x04u  /* user->ignall               c:1413   */
#define ORC_UF_IGNSHOUT 0x08u  /* user->ignshout             c:1414   */
#define ORC_UF_CLONE    0x10u  /* type==CLONE_TYPE  (write_level skips; relay is "next") */
#define ORC_UF_REMOTE   0x20u  /* type==REMOTE_TYPE (netlink framing is "next")          */

/* Op kinds -- one op is one call of the reference's write surface. */
#define ORC_OP_USER  0   /* write_user(target,str)                    c:1291 */
#define ORC_OP_ROOM  1   /* write_room_except(target,str,except)      c:1401 */
#define ORC_OP_LEVEL 2   /* write_level(target,above,str,except)      c:1372 */

/* Op flag bits -- the ambient globals the reference reads inside the call. */
#define ORC_OF_FORCE_LISTEN 0x01u /* force_listen            h:293, c:1413 */
#define ORC_OF_SHOUT        0x02u /* com_num in {SHOUT,SEMOTE}       c:1414 */
#define ORC_OF_ABOVE        0x04u /* write_level 'above' argument    c:1381 */
#define ORC_OF_GATE_IF_SET  0x08u /* op is live iff gate verdict==1 (else ==0) */
#define ORC_OF_PAGER        0x10u /* pager line: no reset after the string   c:2254-2300 */
#define ORC_OF_RAW          0x40u  /* bytes as they are: what goes to a netlink socket (c:1303) */
#define ORC_OF_PLAIN        0x20u /* colour taken as off: more(NULL,...)      c:2259      */

/* ---- single-string primitives ------------------------------------------ */

/* c:1291-1366 for a USER_TYPE recipient: returns bytes written to out.
 * out must hold 6*n+4 bytes. */
size_t orc_render(const uint8_t *s, size_t n, int colour, uint8_t *out);
size_t orc_render_ex(const uint8_t *s, size_t n, int colour, unsigned oflags, uint8_t *out);

/* c:2540-2559 + c:2654-2658.  words = NULL-or-'*'-terminated list, exactly as
 * swear_words[] (h:275-277). */
int orc_contains_swearing(const uint8_t *s, size_t n, const char *const *words);

/* c:2563-2583 (incl. the double-count quirk) and c:2588-2610. */
int    orc_colour_com_count(const uint8_t *s, size_t n);
size_t orc_colour_com_strip(const uint8_t *s, size_t n, uint8_t *out);

/* Ban lists: the fscanf("%s")/feof loop of c:330-364 applied to raw file
 * bytes.  Returns the number of TESTED tokens; tok_off/tok_len index file[].
 * Pass cap=0 to count only. */
size_t orc_ban_tokens(const uint8_t *file, size_t n, uint32_t *tok_off,
                      uint32_t *tok_len, size_t cap);
int orc_site_banned(const uint8_t *file, size_t n, int file_present,
                    const uint8_t *site, size_t site_len);
int orc_user_banned(const uint8_t *file, size_t n, int file_present,
                    const uint8_t *name, size_t name_len);

/* ---- batch drivers ------------------------------------------------------ */

/* Recipient filter of one op for one user (c:1410-1415, c:1379-1383, c:1298).
 * room[u] < 0 means user->room==NULL. */
int orc_delivers(uint8_t kind, int32_t target, int32_t except_user, uint8_t oflags,
                 int32_t u, int32_t u_room, uint8_t u_flags, uint8_t u_level);

/* Remote users (REMOTE_TYPE, c:1299-1307) inside orc_write_batch: link[u] = the pseudo-user whose stream
 * stands for the netlink socket of remote user u (-1 for everybody else), old[u] != 0 for a peer older than
 * 3.2 (colour commands stripped), names of the users (packed) for the "MSG <name>" line.  NULLs: none. */
void orc_set_remotes(const int32_t *link, const uint8_t *old, const uint8_t *names, const uint64_t *name_off);

/* Clone relay (c:1416-1426) inside orc_write_batch: owner[u] (-1 unless u is a clone), hear[u] =
 * clone_hear (0 nothing, 1 swears, 2 all), the swear list for hear == 1.  NULL, NULL, NULL: no clones. */
void orc_set_clones(const int32_t *owner, const uint8_t *hear, const char *const *words);

typedef struct {
    int64_t   n_users;
    uint64_t *off;     /* n_users+1 stream offsets            */
    uint8_t  *bytes;   /* concatenated per-user socket bytes  */
    uint64_t *n_deliveries; /* per user                       */
} orc_streams;

/* Parity digests of SURVEY.md 8(d): per user and per op (see nuts_oracle.c).  -2: clones / remote users present. */
int orc_delivery_digests(int64_t n_ops, const uint8_t *text, const uint64_t *text_off,
                         const uint8_t *kind, const int32_t *target, const int32_t *except_user, const uint8_t *oflags,
                         const int32_t *gate, const uint8_t *verdict,
                         int32_t n_users, const int32_t *room, const uint8_t *uflags, const uint8_t *ulevel,
                         uint64_t *per_user, uint64_t *per_op);

/* Runs n_ops calls in order against the population and returns, per user, the
 * exact byte stream the reference would have written to that user's socket.
 * gate[i] >= 0 makes op i conditional on verdict[gate[i]] (say(), c:4091).
 * only_users (optional, may be NULL): restrict materialisation to these users
 * (sampled-user parity at full size); streams for others are empty. */
int orc_write_batch(int64_t n_ops, const uint8_t *text, const uint64_t *text_off,
                    const uint8_t *kind, const int32_t *target,
                    const int32_t *except_user, const uint8_t *oflags,
                    const int32_t *gate, const uint8_t *verdict,
                    int32_t n_users, const int32_t *room, const uint8_t *uflags,
                    const uint8_t *ulevel,
                    const int32_t *only_users, int32_t n_only,
                    orc_streams *out);
void orc_streams_free(orc_streams *s);

/* Timing leg for bench.py: same loop, bytes rendered into a reused scratch
 * buffer (no materialised streams).  Returns deliveries made; *out_bytes gets
 * the rendered byte count. */
int64_t orc_write_batch_count(int64_t n_ops, const uint8_t *text, const uint64_t *text_off,
                    const uint8_t *kind, const int32_t *target,
                    const int32_t *except_user, const uint8_t *oflags,
                    const int32_t *gate, const uint8_t *verdict,
                    int32_t n_users, const int32_t *room, const uint8_t *uflags,
                    const uint8_t *ulevel, uint64_t *out_bytes);

void orc_contains_swearing_batch(int64_t n, const uint8_t *text, const uint64_t *off,
                                 const char *const *words, uint8_t *verdict);
void orc_site_banned_batch(const uint8_t *file, size_t fn, int present, int64_t n,
                           const uint8_t *text, const uint64_t *off, uint8_t *verdict);
void orc_user_banned_batch(const uint8_t *file, size_t fn, int present, int64_t n,
                           const uint8_t *text, const uint64_t *off, uint8_t *verdict);

/* Ban-list maintenance on file bytes (see nuts_oracle.c): ban_site c:6216 / ban_user c:6262 (add = 1),
 * unban_site c:6341 / unban_user c:6385 (add = 0).  out receives the file as it stands afterwards
 * (cap >= n + strlen(token) + 2), *out_present whether it exists.  Returns 0 done, 1 nothing to do
 * ("already banned" / "not currently banned"). */
int orc_ban_edit(const uint8_t *file, size_t n, int present, int is_user, int add, const char *token,
                 uint8_t *out, size_t *out_n, int *out_present);

/* The callers say/shout/emote/semote/echo/bcast restated (see nuts_oracle.c): input lines -> ops. */
/* verbs: 0 say 1 shout 2 emote 3 semote 4 echo 5 bcast 6 review (own room) 7 tell 8 pemote 9 wizshout
 * 10 revtell; for 7 and 8 the target user of input line m is target[m] (orc_set_speech_targets). */
void orc_set_speech_targets(const int32_t *target);
int64_t orc_speech_ops(int64_t n, const uint8_t *verb, const int32_t *speaker, const uint8_t *bodies, const uint64_t *body_off,
                       const uint8_t *names, const uint64_t *name_off, const uint8_t *sflags, const int32_t *room,
                       int ban_swearing, const char *const *words,
                       uint8_t *text, size_t text_cap, uint64_t *off, uint8_t *kind, int32_t *target, int32_t *except_user,
                       uint8_t *flags, int64_t cap);

/* The pager, c:2205-2322 (see nuts_oracle.c). */
int orc_more(const uint8_t *file, size_t n, int present, int user_null, int colour,
             int64_t *filepos, uint8_t *out, size_t *out_len);

/* 64-bit FNV-1a, the digest of SURVEY.md section 8(d). */
uint64_t orc_fnv1a(const uint8_t *p, size_t n);

#ifdef __cplusplus
}
#endif
#endif
