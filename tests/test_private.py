"""tell c:4128, pemote c:4234, wizshout c:6527, revtell c:7699 (the rest of the callers).

  * the oracle restatement against the reference's OWN functions driven in-process (the target found by the
    reference's get_user from the name in the input line),
  * the queue tier (Talker.tell / pemote / wizshout / revtell) against the oracle: emulator here, GPU with `-m gpu`.
"""
import random

import numpy as np
import pytest

import oracle_lib as O
from nuts333_b200 import api

STOCK = ["fuck", "shit", "cunt", "*"]
SAY, TELL, PEMOTE, WIZSHOUT, REVTELL = 0, 7, 8, 9, 10


def make_script(seed, U, N):
    rng = random.Random(seed)
    room = np.array([rng.randint(0, 1) for _ in range(U)], np.int32)
    flags = np.array([rng.choice([0, 1]) for _ in range(U)], np.uint8)
    level = np.array([rng.randint(0, 4) for _ in range(U)], np.uint8)
    names = [("U%c%c" % (chr(97 + u // 26), chr(97 + u % 26)) + "".join(rng.choice("xyz") for _ in range(rng.randint(0, 6)))).encode() for u in range(U)]
    sflags = np.array([rng.choice([0, 0, 0, 1, 2]) for _ in range(U)], np.uint8)
    words = ["hello", "there", "~FRred", "~OLbold~RS", "what", "shit", "ok", "x" * 70, "~", "/~FG", "y" * 95]
    verbs, speakers, targets, bodies = [], [], [], []
    for _ in range(N):
        v = rng.choice([SAY, TELL, TELL, TELL, PEMOTE, PEMOTE, WIZSHOUT, REVTELL])
        s = rng.randint(0, U - 1)
        t = rng.choice([x for x in range(U) if x != s])
        b = " ".join(rng.choice(words) for _ in range(rng.randint(1, 7)))
        if rng.random() < 0.3: b += rng.choice("?!")
        verbs.append(v); speakers.append(s); targets.append(t if v in (TELL, PEMOTE) else -1)
        bodies.append(b"" if v == REVTELL else b.encode())
    bt, bo = O.pack(bodies)
    return dict(users=dict(room=room, flags=flags, level=level), n_rooms=2, names=names, sflags=sflags, verb=np.array(verbs, np.uint8),
                speaker=np.array(speakers, np.int32), target=np.array(targets, np.int32), bodies=bodies, bt=bt, bo=bo)


def port_streams(port, c, ban):
    nt, no = O.pack(c["names"])
    ops = port.speech_ops(c["verb"], c["speaker"], c["bt"], c["bo"], nt, no, c["sflags"], c["users"]["room"], ban, STOCK, target=c["target"])
    return port.write_batch(ops, c["users"])


def test_private_oracle_vs_reference(port, ref):
    for seed, ban in ((41, True), (42, False)):
        c = make_script(seed, 14, 300)
        off, data, nd = port_streams(port, c, ban)
        ref.reset(c["n_rooms"], c["users"])
        ref.set_swear_words(STOCK[:-1])
        ref.lib.ref_set_ban_swearing(int(ban))
        for u, nm in enumerate(c["names"]):
            ref.lib.ref_set_user_speech(u, nm, int(not (c["sflags"][u] & 1)), int((c["sflags"][u] & 2) != 0))
        for v, s, t, b in zip(c["verb"], c["speaker"], c["target"], c["bodies"]):
            if int(v) in (TELL, PEMOTE):
                ref.lib.ref_speech_to(int(v), int(s), int(t), b)
            else:
                ref.lib.ref_speech(int(v), int(s), b)
        for u in range(14):
            assert data[int(off[u]):int(off[u + 1])].tobytes() == ref.stream(u), (seed, u)
    ref.lib.ref_set_ban_swearing(0)


def _check_queue_tier(ctx, port, seed, U, N):
    for ban in (True, False):
        c = make_script(seed, U, N)
        off, data, nd = port_streams(port, c, ban)
        ctx.set_swear_words(STOCK)
        # a fresh set of users (create_user() starts with empty revtell buffers, c:2747)
        ctx.set_users(c["users"]["room"], c["users"]["flags"], c["users"]["level"], c["n_rooms"], prev_index=np.full(U, -1, np.int32))
        ctx.set_user_names(c["names"], c["sflags"])
        ctx.set_ban_swearing(ban)
        t = api.Talker(ctx)
        for v, s, tg, b in zip(c["verb"], c["speaker"], c["target"], c["bodies"]):
            v, s, tg = int(v), int(s), int(tg)
            if v == SAY: t.say(s, b)
            elif v == TELL: t.tell(s, tg, b)
            elif v == PEMOTE: t.pemote(s, tg, b)
            elif v == WIZSHOUT: t.wizshout(s, b)
            else: t.revtell(s)
        st = t.flush()
        assert (st.off == off).all() and (st.data == data).all()
    ctx.set_ban_swearing(False)


def test_wizshout_to_a_level(sim_lib, port):
    """the form with a level word: "~OLYou wizshout to level ARCH:~RS ..." and write_level(ARCH, 1, ..., user)"""
    ctx = api.Context(0, sim_lib)
    users = dict(room=np.zeros(4, np.int32), flags=np.array([1, 0, 1, 0], np.uint8), level=np.array([4, 3, 2, 1], np.uint8))
    ctx.set_users(users["room"], users["flags"], users["level"], 1)
    ctx.set_user_names([b"God", b"Arch", b"Wiz", b"User"], np.zeros(4, np.uint8))
    t = api.Talker(ctx)
    t.wizshout(0, b"ARCH  meeting ~FRnow", lev=3, level_name=b"ARCH")     # the whole line: the level word is taken off inside
    st = t.flush()
    assert st.user(0) == port.render(b"~OLYou wizshout to level ARCH:~RS meeting ~FRnow\n", 1)
    assert st.user(1) == port.render(b"~OLGod wizshouts to level ARCH:~RS meeting ~FRnow\n", 0)
    assert st.user(2) == b"" and st.user(3) == b""
    ctx.close()


def test_private_on_emulator(sim_lib, port):
    ctx = api.Context(0, sim_lib)
    _check_queue_tier(ctx, port, 43, 10, 120)
    ctx.close()


@pytest.mark.gpu
def test_private_on_gpu(gpu_ctx, port):
    _check_queue_tier(gpu_ctx, port, 44, 40, 700)


def test_wizshout_swear_check_sees_the_level_word(sim_lib, ref):
    """c:6541 asks contains_swearing of the WHOLE line, before remove_first (c:6553) takes the level word off: a list
    word that matches the level word, or straddles it and the message, refuses the line."""
    words = ["arch", "h mee", "*"]
    ctx = api.Context(0, sim_lib)
    ctx.set_swear_words(words)
    users = dict(room=np.zeros(3, np.int32), flags=np.array([1, 0, 0], np.uint8), level=np.array([4, 3, 2], np.uint8))
    names = [b"God", b"Arch", b"Wiz"]
    ctx.set_users(users["room"], users["flags"], users["level"], 1)
    ctx.set_user_names(names, np.zeros(3, np.uint8))
    ctx.set_ban_swearing(True)
    lines = [(b"GOD hello all", 4, b"GOD"), (b"ARCH hello", 3, b"ARCH"), (b"WIZ meeting", 2, b"WIZ"), (b"GOD h meet", 4, b"GOD"),
             (b"plain words", -1, None), (b"search party", -1, None)]
    ref.reset(1, users)
    ref.set_swear_words(words[:-1])
    ref.lib.ref_set_ban_swearing(1)
    for u, nm in enumerate(names):
        ref.lib.ref_set_user_speech(u, nm, 1, 0)
    t = api.Talker(ctx)
    for line, lev, lname in lines:
        t.wizshout(0, line, lev=lev, level_name=lname)
        ref.lib.ref_speech(WIZSHOUT, 0, line)
    st = t.flush()
    for u in range(3):
        assert st.user(u) == ref.stream(u), u
    assert b"Swearing is not allowed" in st.user(0) and b"wizshout to level GOD" in st.user(0)
    ref.lib.ref_set_ban_swearing(0)
    ctx.close()
