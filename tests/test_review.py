"""Review buffers: record() c:2062 as called by say / emote / echo, review() c:5192.

  * the oracle restatement against the reference's OWN say()/emote()/echo()/review() driven in-process,
  * the queue tier (Talker.say/... + Talker.review) against the oracle: on the SIMT emulator here, on the
    GPU with `-m gpu` (the buffers are host state; the swear verdicts they wait for and the replay through
    write_user are the device's).
"""
import random

import numpy as np
import pytest

import oracle_lib as O
from nuts333_b200 import api

STOCK = ["fuck", "shit", "cunt", "*"]
REVIEW = 6


def make_script(seed, U, NR, N):
    rng = random.Random(seed)
    room = np.array([rng.randint(0, NR - 1) for _ in range(U)], np.int32)
    flags = np.array([rng.choice([0, 1]) for _ in range(U)], np.uint8)
    level = np.array([rng.randint(0, 4) for _ in range(U)], np.uint8)
    names = [("U" + "".join(rng.choice("abcdefghij") for _ in range(rng.randint(2, 10)))).encode() for _ in range(U)]
    sflags = np.array([rng.choice([0, 0, 0, 1, 2]) for _ in range(U)], np.uint8)
    words = ["hello", "there", "~FRred", "~OLbold~RS", "what", "shit", "ok", "x" * 60, "~", "/~FG", "y" * 97]
    verbs, speakers, bodies = [], [], []
    for _ in range(N):
        r = rng.random()
        if r < 0.12:
            verbs.append(REVIEW); speakers.append(rng.randint(0, U - 1)); bodies.append(b"")
            continue
        b = " ".join(rng.choice(words) for _ in range(rng.randint(1, 7)))
        if rng.random() < 0.1: b = ";" + b
        if rng.random() < 0.3: b += rng.choice("?!")
        verbs.append(rng.choice([0, 0, 0, 1, 2, 2, 3, 4, 5])); speakers.append(rng.randint(0, U - 1)); bodies.append(b.encode())
    bt, bo = O.pack(bodies)
    return dict(users=dict(room=room, flags=flags, level=level), n_rooms=NR, names=names, sflags=sflags,
                verb=np.array(verbs, np.uint8), speaker=np.array(speakers, np.int32), bodies=bodies, bt=bt, bo=bo)


def port_streams(port, c, ban):
    nt, no = O.pack(c["names"])
    ops = port.speech_ops(c["verb"], c["speaker"], c["bt"], c["bo"], nt, no, c["sflags"], c["users"]["room"], ban, STOCK)
    return port.write_batch(ops, c["users"])


def test_review_oracle_vs_reference(port, ref):
    for seed, ban in ((11, True), (12, False)):
        c = make_script(seed, 12, 2, 260)
        off, data, nd = port_streams(port, c, ban)
        ref.reset(c["n_rooms"], c["users"])
        ref.set_swear_words(STOCK[:-1])
        ref.lib.ref_set_ban_swearing(int(ban))
        for u, nm in enumerate(c["names"]):
            ref.lib.ref_set_user_speech(u, nm, int(not (c["sflags"][u] & 1)), int((c["sflags"][u] & 2) != 0))
        for v, s, b in zip(c["verb"], c["speaker"], c["bodies"]):
            ref.lib.ref_speech(int(v), int(s), b)
        for u in range(12):
            assert data[int(off[u]):int(off[u + 1])].tobytes() == ref.stream(u), (seed, u)
    ref.lib.ref_set_ban_swearing(0)


def test_long_lines_are_cut_at_200_bytes(port, ref):
    users = dict(room=np.zeros(2, np.int32), flags=np.array([1, 0], np.uint8), level=np.ones(2, np.uint8))
    bodies = [b"a" * 250, b"b" * 190, b"c" * 191, b"~FR" * 70, b""]
    verb = np.array([0, 0, 0, 2, REVIEW], np.uint8); spk = np.array([0, 1, 0, 1, 1], np.int32)
    bt, bo = O.pack(bodies)
    c = dict(users=users, n_rooms=1, names=[b"Ua", b"Ub"], sflags=np.zeros(2, np.uint8), verb=verb, speaker=spk, bt=bt, bo=bo)
    off, data, _ = port_streams(port, c, False)
    ref.reset(1, users)
    for u, nm in enumerate(c["names"]):
        ref.lib.ref_set_user_speech(u, nm, 1, 0)
    for v, s, b in zip(verb, spk, bodies):
        ref.lib.ref_speech(int(v), int(s), b)
    for u in range(2):
        assert data[int(off[u]):int(off[u + 1])].tobytes() == ref.stream(u)


def _check_queue_tier(ctx, port, seed, U, NR, N):
    for ban in (True, False):
        c = make_script(seed, U, NR, N)
        off, data, nd = port_streams(port, c, ban)
        ctx.set_swear_words(STOCK)
        ctx.set_users(c["users"]["room"], c["users"]["flags"], c["users"]["level"], c["n_rooms"])
        ctx.set_user_names(c["names"], c["sflags"])
        ctx.set_ban_swearing(ban)
        t = api.Talker(ctx)
        for r in range(c["n_rooms"]):          # the rooms' buffers outlive nutsb_set_users (only create_room / clear_revbuff empty them)
            t.clear_revbuff(r)
        fn = [t.say, t.shout, t.emote, t.semote, t.echo, t.bcast]
        for v, s, b in zip(c["verb"], c["speaker"], c["bodies"]):
            if int(v) == REVIEW:
                t.review(int(s), int(c["users"]["room"][s]))
            else:
                fn[int(v)](int(s), b)
        st = t.flush()
        assert (st.off == off).all() and (st.data == data).all()
    ctx.set_ban_swearing(False)


def test_review_on_emulator(sim_lib, port):
    ctx = api.Context(0, sim_lib)
    _check_queue_tier(ctx, port, 13, 10, 2, 120)
    ctx.close()


@pytest.mark.gpu
def test_review_on_gpu(gpu_ctx, port):
    _check_queue_tier(gpu_ctx, port, 14, 40, 3, 600)


def test_population_updates_keep_the_buffers_and_wait_for_the_flush(sim_lib, port):
    """nutsb_set_users between two flushes only (queued ops name users by index: NUTSB_E_STATE otherwise), and it
    leaves the review / revtell buffers alone: a logout shifts the list, prev_index carries the revtell buffers over."""
    ctx = api.Context(0, sim_lib)
    room, flags, level = np.zeros(3, np.int32), np.array([0, 1, 0], np.uint8), np.ones(3, np.uint8)
    ctx.set_users(room, flags, level, 1)
    ctx.set_user_names([b"Ann", b"Bob", b"Cy"], np.zeros(3, np.uint8))
    t = api.Talker(ctx)
    t.say(0, b"first line")
    t.tell(0, 2, b"psst")
    with pytest.raises(api.NutsbError) as e:
        ctx.set_users(room[:2], flags[:2], level[:2], 1)
    assert e.value.code == api.E_STATE and t.pending() > 0
    st = t.flush()
    assert st.user(2) == port.render(b"Ann says: first line\n", 0) + port.render(b"~OLAnn tells you:~RS psst\n", 0)
    # Bob logs out: Cy moves from index 2 to 1 and keeps his revtell buffer; the room keeps its review buffer
    ctx.set_users(room[:2], flags[[0, 2]], level[:2], 1, prev_index=np.array([0, 2], np.int32))
    ctx.set_user_names([b"Ann", b"Cy"], np.zeros(2, np.uint8))
    t.revtell(1)
    t.review(0, 0)
    st = t.flush()
    assert b"Ann tells you:" in st.user(1) and b"Review buffer is empty" not in st.user(0) and b"Ann says: first line" in st.user(0)
    # a new user at index 1 instead: empty revtell buffer
    ctx.set_users(room[:2], flags[:2], level[:2], 1, prev_index=np.array([0, -1], np.int32))
    ctx.set_user_names([b"Ann", b"Dee"], np.zeros(2, np.uint8))
    t.revtell(1)
    assert t.flush().user(1) == port.render(b"Revtell buffer is empty.\n", 1)
    ctx.close()
