"""The callers' composition row: say / shout / emote / semote / echo /
bcast as input lines.

  * the oracle restatement of the six callers against the reference's OWN functions driven
    in-process (say(), shout(), ... from nuts333.c),
  * the device composer (nutsb_speech_batch) and the queue tier (nutsb_q_speech through
    Talker.say/...) against the oracle: on the SIMT emulator here, on the GPU with `-m gpu`.
"""
import random

import numpy as np
import pytest

import oracle_lib as O
from nuts333_b200 import api

STOCK = ["fuck", "shit", "cunt", "*"]


def make_case(seed, U, NR, N):
    rng = random.Random(seed)
    room = np.array([rng.randint(-1, NR - 1) if rng.random() < 0.1 else rng.randint(0, NR - 1) for _ in range(U)], np.int32)
    flags = np.array([rng.choice([0, 1, 1, 0, 4, 8, 5, 9, 2]) for _ in range(U)], np.uint8)
    level = np.array([rng.randint(0, 4) for _ in range(U)], np.uint8)
    names = [("U" + "".join(rng.choice("abcdefghij") for _ in range(rng.randint(2, 10)))).encode() for _ in range(U)]
    sflags = np.array([rng.choice([0, 0, 0, 1, 2, 3]) for _ in range(U)], np.uint8)
    words = ["hello", "there", "~FRred", "~OLbold~RS", "what", "shit", "FUCK", "sh~RSit", "a/~b", "ok", "x" * 30, "~", "/~FG"]
    verbs, speakers, bodies = [], [], []
    for _ in range(N):
        b = " ".join(rng.choice(words) for _ in range(rng.randint(1, 8)))
        r = rng.random()
        if r < 0.1: b = ";" + b
        elif r < 0.2: b = "#" + b
        r = rng.random()
        if r < 0.25: b += "?"
        elif r < 0.5: b += "!"
        verbs.append(rng.randint(0, 5)); speakers.append(rng.randint(0, U - 1)); bodies.append(b.encode())
    bt, bo = O.pack(bodies)
    return dict(users=dict(room=room, flags=flags, level=level), n_rooms=NR, names=names, sflags=sflags,
                verb=np.array(verbs, np.uint8), speaker=np.array(speakers, np.int32), bodies=bodies, bt=bt, bo=bo)


def port_streams(port, c, ban):
    nt, no = O.pack(c["names"])
    ops = port.speech_ops(c["verb"], c["speaker"], c["bt"], c["bo"], nt, no, c["sflags"], c["users"]["room"], ban, STOCK)
    return port.write_batch(ops, c["users"])


def test_callers_oracle_vs_reference(port, ref):
    for seed, ban in ((1, True), (2, False), (3, True)):
        c = make_case(seed, 40, 3, 300)
        off, data, nd = port_streams(port, c, ban)
        ref.reset(c["n_rooms"], c["users"])
        ref.set_swear_words(STOCK[:-1])
        ref.lib.ref_set_ban_swearing(int(ban))
        for u, nm in enumerate(c["names"]):
            ref.lib.ref_set_user_speech(u, nm, int(not (c["sflags"][u] & 1)), int((c["sflags"][u] & 2) != 0))
        for v, s, b in zip(c["verb"], c["speaker"], c["bodies"]):
            # say/emote/echo by a user in no room never reach these functions in the talker
            # (they would touch user->netlink / user->room): nothing is written either way
            if c["users"]["room"][s] < 0 and int(v) in (0, 2, 4) and not (c["sflags"][s] & 2):
                continue
            ref.lib.ref_speech(int(v), int(s), b)
        for u in range(40):
            assert data[int(off[u]):int(off[u + 1])].tobytes() == ref.stream(u), (seed, u)
    ref.lib.ref_set_ban_swearing(0)


def _check_composer(ctx, port, seed, U, NR, N):
    for ban in (True, False):
        c = make_case(seed, U, NR, N)
        off, data, nd = port_streams(port, c, ban)
        ctx.set_swear_words(STOCK)
        ctx.set_users(c["users"]["room"], c["users"]["flags"], c["users"]["level"], c["n_rooms"])
        ctx.set_user_names(c["names"], c["sflags"])
        ctx.set_ban_swearing(ban)
        # batch tier: composed on the device
        st = ctx.speech_batch(c["verb"], c["speaker"], c["bt"], c["bo"])
        assert (st.off == off).all() and (st.data == data).all() and st.n_deliveries == int(nd.sum())
        # ... and with gather lists as the result (nutsb_speech_batch_iov)
        if ban:
            iv = ctx.speech_batch_iov(c["verb"], c["speaker"], c["bt"], c["bo"])
            assert (iv.off == off).all() and iv.n_deliveries == int(nd.sum())
            for u in range(U):
                assert iv.user(u) == data[int(off[u]):int(off[u + 1])].tobytes(), (seed, u)
        # queue tier: the reference's own names, one call per line
        t = api.Talker(ctx)
        fn = [t.say, t.shout, t.emote, t.semote, t.echo, t.bcast]
        for v, s, b in zip(c["verb"], c["speaker"], c["bodies"]):
            fn[int(v)](int(s), b)
        st2 = t.flush()
        assert (st2.off == off).all() and (st2.data == data).all()
    ctx.set_ban_swearing(False)


def _check_bad_body_offsets(ctx):
    """offsets that are not monotone are refused (caught on the device), for both result forms"""
    ctx.set_users(np.zeros(2, np.int32), np.zeros(2, np.uint8), np.ones(2, np.uint8), 1)
    ctx.set_user_names([b"Ua", b"Ub"], np.zeros(2, np.uint8))
    bt = np.frombuffer(b"hello there, world", np.uint8).copy()
    verb, spk = np.zeros(3, np.uint8), np.zeros(3, np.int32)
    for bo in ([0, 12, 5, 18], [0, 40, 12, 18], [4, 2, 12, 18]):
        for call in (ctx.speech_batch, ctx.speech_batch_iov):
            with pytest.raises(api.NutsbError) as e:
                call(verb, spk, bt, np.array(bo, np.uint64))
            assert e.value.code == api.E_INVAL
    assert ctx.speech_batch(verb, spk, bt, np.array([0, 5, 12, 18], np.uint64)).total_bytes > 0


def test_composer_bad_offsets_on_emulator(sim_lib):
    ctx = api.Context(0, sim_lib)
    _check_bad_body_offsets(ctx)
    ctx.close()


@pytest.mark.gpu
def test_composer_bad_offsets_on_gpu(gpu_ctx):
    _check_bad_body_offsets(gpu_ctx)


def test_composer_on_emulator(sim_lib, port):
    ctx = api.Context(0, sim_lib)
    _check_composer(ctx, port, 4, 30, 3, 120)
    ctx.close()


@pytest.mark.gpu
def test_composer_on_gpu(gpu_ctx, port):
    _check_composer(gpu_ctx, port, 5, 60, 4, 800)
    _check_composer(gpu_ctx, port, 6, 500, 3, 4000)


@pytest.mark.gpu
def test_composer_errors(gpu_ctx):
    gpu_ctx.set_users(np.zeros(2, np.int32), np.zeros(2, np.uint8), np.ones(2, np.uint8), 1)
    gpu_ctx.set_user_names([b"Ua", b"Ub"], np.zeros(2, np.uint8))
    bt, bo = O.pack([b"hi"])
    with pytest.raises(api.NutsbError) as e:
        gpu_ctx.speech_batch(np.array([0], np.uint8), np.array([7], np.int32), bt, bo)
    assert e.value.code == api.E_RANGE
    with pytest.raises(api.NutsbError) as e:
        gpu_ctx.speech_batch(np.array([9], np.uint8), np.array([0], np.int32), bt, bo)
    assert e.value.code == api.E_INVAL
    gpu_ctx.set_users(np.zeros(3, np.int32), np.zeros(3, np.uint8), np.ones(3, np.uint8), 1)
    with pytest.raises(api.NutsbError) as e:                  # names no longer match the population
        gpu_ctx.speech_batch(np.array([0], np.uint8), np.array([0], np.int32), bt, bo)
    assert e.value.code == api.E_STATE
