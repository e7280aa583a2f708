"""The drop-in, end to end: the UNMODIFIED nuts333.c driven through its own main-loop body and exec_com()
(say, .shout, .tell, .emote, .semote, .pemote, .echo, .bcast, .wizshout, .review, .revtell, .go, .colour, .ignall,
.look, .who, .help -> more(), .clone / .csay, .ban / .unban, .quit ...), once with its own write layer and once with
the eight bodies of shim/nuts333_shim.c linked over it (tests/dropin/): every socket must receive the same bytes.

The shim build resolves nutsb_* at load time: the product library on the GPU (`-m gpu`), the emulator build here."""
import ctypes as C
import os
import random
import shutil
import subprocess
import tempfile
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
REFDIR = ROOT / "oracle" / "_ref"
WORDS64 = None


def _build():
    if Path("/root/reference/nuts333.c").exists():
        subprocess.run(["make", "-C", str(ROOT / "tests" / "dropin")], check=True, stdout=subprocess.DEVNULL)
    if not (REFDIR / "libdropin_ref.so").exists() or not (REFDIR / "libdropin_shim.so").exists():
        pytest.skip("oracle/_ref/libdropin_*.so not built (reference sources absent)")


def _bind(lib):
    lib.dropin_reset.argtypes = [C.c_char_p, C.c_int, C.c_int]
    lib.dropin_add_room.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    lib.dropin_add_netlink.argtypes = [C.c_char_p, C.c_int, C.c_int]
    lib.dropin_add_user.argtypes = [C.c_char_p] + [C.c_int] * 6
    lib.dropin_add_remote_user.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
    lib.dropin_input.argtypes = [C.c_int, C.c_char_p]
    lib.dropin_site_banned.argtypes = [C.c_char_p]
    lib.dropin_user_banned.argtypes = [C.c_char_p]
    lib.dropin_contains_swearing.argtypes = [C.c_char_p]
    lib.dropin_login_attempt.argtypes = [C.c_char_p]
    lib.dropin_set_swear_words.argtypes = [C.POINTER(C.c_char_p)]
    lib.dropin_stream_len.restype = C.c_size_t
    lib.dropin_stream_ptr.restype = C.POINTER(C.c_uint8)
    lib.dropin_stream_calls.restype = C.c_uint64
    lib.dropin_last_error.restype = C.c_char_p
    lib.dropin_stats.argtypes = [C.POINTER(C.c_uint64)]
    return lib


def _scratch(tmp):
    d = Path(tmp)
    for sub in ("datafiles", "helpfiles", "userfiles", "mailspool"):
        (d / sub).mkdir(parents=True, exist_ok=True)
    # helpfiles with colour commands; "colour" is longer than a page (23 lines): more() pages it
    (d / "helpfiles" / "colour").write_bytes(b"".join(b"~FR line %02d ~OLbold~RS /~FG escaped ~BBblue~RS ~XX ~\n" % i for i in range(40)))
    (d / "helpfiles" / "say").write_bytes(b"~FTUsage:~RS say <text>\n\nSays something.\n")
    (d / "datafiles" / "siteban").write_bytes(b"evil.com\n.badnet.org\n10.1.\nlast.noeol")
    (d / "datafiles" / "userban").write_bytes(b"Troll\nSpammer\nNoeol")
    return d


def make_script(seed, n_users, n_rooms, n_lines):
    """A deterministic session: set-up records and input lines (bytes), independent of which talker runs it."""
    rng = random.Random(seed)
    name = lambda i: "U" + "".join(chr(97 + (i // 26 ** k) % 26) for k in (2, 1, 0))
    users = []
    for i in range(n_users):
        users.append(dict(name=name(i), room=rng.randrange(n_rooms), level=rng.choice([1, 1, 1, 2, 3, 4]), colour=rng.randrange(2),
                          login=0, prompt=rng.random() < 0.3, cmode=rng.random() < 0.1,
                          ignall=rng.random() < 0.05, ignshout=rng.random() < 0.08, vis=rng.random() > 0.07, muzzled=rng.random() < 0.04))
    users[0].update(level=4, cmode=False, muzzled=False)              # a GOD who may do everything
    users[1].update(level=4, cmode=False, muzzled=False, colour=1)
    vocab = ["hello", "there", "~FRred", "~OLbold~RS", "what", "shit", "FUCK", "a/~b", "ok", "~", "~FX", "x" * 30, "/~FG", "~BK~FWinv",
             "scunthorpe", "sh~RSit", "really", "the", "quick", "brown", "fox", "~LIblink", "~UL_~RS", "tail~"]
    body = lambda: " ".join(rng.choice(vocab) for _ in range(rng.randint(1, 9))) + rng.choice(["", "", "?", "!"])
    lines = []
    for _ in range(n_lines):
        u = rng.randrange(n_users)
        r = rng.random()
        other = name(rng.randrange(n_users))
        room = "room%d" % rng.randrange(n_rooms)
        if r < 0.38: ln = body()
        elif r < 0.46: ln = ".shout " + body()
        elif r < 0.54: ln = ".tell %s %s" % (other, body())
        elif r < 0.60: ln = rng.choice([".emote ", ";"]) + body()
        elif r < 0.64: ln = rng.choice([".semote ", "#"]) + body()
        elif r < 0.68: ln = ".pemote %s %s" % (other, body())
        elif r < 0.71: ln = rng.choice([".echo ", "-"]) + body()
        elif r < 0.73: ln = ".bcast " + body()
        elif r < 0.76: ln = ".wizshout " + rng.choice(["", "", "GOD ", "ARCH ", "WIZ "]) + body()
        elif r < 0.79: ln = rng.choice([".review", ".revtell", ".review " + room])
        elif r < 0.83: ln = ".go " + room
        elif r < 0.86: ln = rng.choice([".colour", ".ignall", ".ignshout", ".igntell", ".vis", ".invis", ".mode", ".prompt"])
        # (.people is left out: its sockstr[3] (c:4799, sprintf "%2d" c:4835) overflows for the harness's socket numbers)
        elif r < 0.89: ln = rng.choice([".look", ".who", ".who", ".help", ".help colour", ".help say", ".version", ".ranks", ".exits", "."])
        elif r < 0.91: ln = rng.choice([".clone " + room, ".myclones", ".csay %s %s" % (room, body()), ".chear %s %s" % (room, rng.choice(["all", "swears", "nothing"])),
                                        ".destroy " + room, ".cemote %s %s" % (room, body())])
        elif r < 0.93: ln = rng.choice([".muzzle " + other, ".unmuzzle " + other, ".wake " + other, ".topic " + body(), ".desc " + body()[:30], ".cbuff", ".afk", ".afk gone ~FRfishing"])
        elif r < 0.95: ln = rng.choice([".ban site s%d.example.org" % rng.randrange(5), ".unban site s%d.example.org" % rng.randrange(5), ".ban user " + other,
                                        ".unban user " + other, ".listbans sites", ".listbans users", ".swban"])
        elif r < 0.96: ln = rng.choice(["", "e", ".write " + body(), ".read", ".wipe all"])
        elif r < 0.965: ln = ".quit"
        else: ln = body()
        lines.append((u, ln.encode()[:900]))
    return dict(users=users, n_rooms=n_rooms, lines=lines, seed=seed)


def run_session(lib, scratch, sc, device, use_iov, swear_words, flush_rng=None, checks=True):
    """Runs the script on one talker; returns {fd: bytes}, the verdict list and the stats."""
    assert lib.dropin_reset(str(scratch).encode(), device, 1 if use_iov else 0) == 0, lib.dropin_last_error()
    arr = (C.c_char_p * (len(swear_words) + 1))(*[w.encode() for w in swear_words], None)
    lib.dropin_set_swear_words(arr)
    lib.dropin_set_globals(1, 0)                                            # ban_swearing on (datafiles/config:13)
    n_rooms = sc["n_rooms"]
    for r in range(n_rooms):
        lib.dropin_add_room(b"room%d" % r, b"r%d" % r, b"A ~FGgreen~RS room, number %d.\n~OLTwo lines of it.\n" % r, 0)
    for r in range(n_rooms):
        lib.dropin_link_rooms(r, (r + 1) % n_rooms)
        if r + 2 < n_rooms: lib.dropin_link_rooms(r, r + 2)
    l_new = lib.dropin_add_netlink(b"newtalker", 0, 3)
    l_old = lib.dropin_add_netlink(b"oldtalker", 1 % n_rooms, 1)            # older than 3.2: colour commands stripped (c:1300)
    handles = []
    for i, u in enumerate(sc["users"]):
        h = lib.dropin_add_user(u["name"].encode(), u["room"], u["level"], u["colour"], u["login"], int(u["prompt"]), int(u["cmode"]))
        for f, key in ((1, "ignall"), (2, "ignshout"), (3, "vis"), (4, "muzzled")):
            lib.dropin_set_field(h, f, int(u[key]))
        handles.append(h)
        if i == 3: lib.dropin_add_remote_user(b"Remy", 0, 2, l_new)
        if i == 5: lib.dropin_add_remote_user(b"Oldie", 1 % n_rooms, 1, l_old)
        if i == 7: lib.dropin_add_user(b"Loggingin", -1, 0, 0, 3, 0, 0)     # sits at a login stage: receives nothing
    verdicts = []
    rng = random.Random(sc["seed"] * 7 + 1)
    frng = flush_rng or random.Random(99)
    next_flush = frng.randint(1, 40)
    for k, (u, ln) in enumerate(sc["lines"]):
        lib.dropin_input(handles[u], ln)
        if not lib.dropin_user_alive(handles[u]) or rng.random() < 0.004:   # somebody left or joins: a newcomer takes the handle
            nu = sc["users"][u]
            handles[u] = lib.dropin_add_user(("N%s%d" % (nu["name"][1:], k)).encode()[:12].rstrip(b"0123456789") + bytes([97 + k % 26]),
                                              rng.randrange(n_rooms), nu["level"], rng.randrange(2), 0, 0, 0)
        if checks and k % 97 == 0:                                          # admission: c:279, c:1496 and the raw verdicts
            for s in (b"host.evil.com", b"x.badnet.org", b"badnet.org", b"110.1.2.3", b"last.noeol", b"good.org", b"s3.example.org"):
                verdicts.append(lib.dropin_site_banned(s))
            for nme in (b"Troll", b"troll", b"Noeol", b"Spammer", sc["users"][k % len(sc["users"])]["name"].encode()):
                verdicts.append(lib.dropin_user_banned(nme))
            verdicts.append(lib.dropin_contains_swearing(ln or b"x"))
            lib.dropin_login_attempt(rng.choice([b"troll", b"Spammer", b"newbie", b"ab", sc["users"][2]["name"].encode()]))
        if k % 211 == 0:
            lib.dropin_events(rng.randint(1, 30))
        next_flush -= 1
        if next_flush <= 0:
            lib.dropin_flush()                                              # "once per main-loop iteration" -- here every 1..40 lines
            next_flush = frng.randint(1, 40)
    lib.dropin_flush()
    assert lib.dropin_errors() == 0, lib.dropin_last_error()
    out = {}
    for fd in range(lib.dropin_fd_base(), lib.dropin_next_fd()):
        n = lib.dropin_stream_len(fd)
        out[fd] = bytes(np.ctypeslib.as_array(lib.dropin_stream_ptr(fd), shape=(n,))) if n else b""
    st = (C.c_uint64 * 4)()
    lib.dropin_stats(st)
    return out, verdicts, list(st)


def _compare(sc, nutsb_lib_path, device, use_iov, swear_words):
    _build()
    C.CDLL(str(nutsb_lib_path), mode=C.RTLD_GLOBAL)                      # nutsb_* for the shim build, bound at load time
    ref = _bind(C.CDLL(str(REFDIR / "libdropin_ref.so")))
    shim = _bind(C.CDLL(str(REFDIR / "libdropin_shim.so")))
    assert not ref.dropin_is_shim() and shim.dropin_is_shim()
    cwd = os.getcwd()
    ta, tb = tempfile.mkdtemp(prefix="dropin_ref_"), tempfile.mkdtemp(prefix="dropin_shim_")
    try:
        a, va, sa = run_session(ref, _scratch(ta), sc, device, use_iov, swear_words)
        b, vb, sb = run_session(shim, _scratch(tb), sc, device, use_iov, swear_words)
    finally:
        os.chdir(cwd)
        shutil.rmtree(ta, ignore_errors=True); shutil.rmtree(tb, ignore_errors=True)
    assert va == vb, "admission / swear verdicts differ"
    assert a.keys() == b.keys()
    for fd in sorted(a):
        if a[fd] != b[fd]:
            i = next((j for j in range(min(len(a[fd]), len(b[fd]))) if a[fd][j] != b[fd][j]), min(len(a[fd]), len(b[fd])))
            raise AssertionError(f"socket {fd}: {len(a[fd])} vs {len(b[fd])} bytes, first difference at {i}: "
                                 f"{a[fd][max(0, i - 60):i + 60]!r} vs {b[fd][max(0, i - 60):i + 60]!r}")
    total = sum(len(x) for x in a.values())
    assert total > 0 and sum(1 for x in a.values() if x) >= len(sc["users"]) // 2
    return total, sa, sb


def _words64():
    from nuts333_b200 import synth
    return [w for w in synth.swear_words(64) if w != "*"]


def test_dropin_on_emulator(sim_lib):
    from cpusim.build_sim import build_sim
    sc = make_script(5, 12, 3, 170)
    total, sa, sb = _compare(sc, build_sim(), 0, False, ["fuck", "shit", "cunt"])
    assert sb[0] > 0 and sb[1] > 170 and sb[3] < sa[3]                    # batched: fewer socket writes than the reference made


def test_dropin_gather_lists_on_emulator(sim_lib):
    from cpusim.build_sim import build_sim
    sc = make_script(6, 10, 2, 90)
    _compare(sc, build_sim(), 0, True, ["fuck", "shit", "cunt"])


@pytest.mark.gpu
@pytest.mark.parametrize("use_iov", [False, True])
def test_dropin_on_gpu(use_iov):
    """12,000 input lines, 130 users (colour on and off, ignall / ignshout, invisible, muzzled, command mode), 6 rooms,
    clones made by .clone, two remote users on netlinks (one older than 3.2), a 64-word swear list, logins and
    logouts in between: byte-identical on every socket."""
    from nuts333_b200 import api, build
    build.build()
    sc = make_script(11 + int(use_iov), 130, 6, 12000)
    total, sa, sb = _compare(sc, api.library_path(), 0, use_iov, _words64())
    assert total > 10_000_000 and sb[1] > 30000
