"""Clone relay (nuts333.c:1416-1426): a CLONE_TYPE user receives nothing itself; what
write_room_except would have sent it goes to its owner as "~FT[ <room> ]:~RS <str>".

  * the oracle restatement against the reference's OWN write_room_except / write_user / write_level with
    real clone users (type, owner, clone_hear set as create_clone does),
  * the queue tier (Talker.write_*) with nutsb_set_clones against the oracle: emulator here, GPU with `-m gpu`.
"""
import ctypes as C
import random

import numpy as np
import pytest

import oracle_lib as O
from nuts333_b200 import api

STOCK = ["fuck", "shit", "cunt", "*"]


def make_case(seed, U, NR, N):
    rng = random.Random(seed)
    room = np.array([rng.randint(0, NR - 1) for _ in range(U)], np.int32)
    flags = np.array([rng.choice([0, 1, 1, 0, 4, 8, 1, 0, 2]) for _ in range(U)], np.uint8)
    level = np.array([rng.randint(0, 4) for _ in range(U)], np.uint8)
    owner = np.full(U, -1, np.int32); hear = np.zeros(U, np.uint8)
    real = [u for u in range(U) if rng.random() < 0.6] or [0]
    for u in range(U):
        if u not in real:
            flags[u] |= 0x10
            owner[u] = rng.choice(real); hear[u] = rng.choice([0, 1, 2, 2])
    words = ["hello", "~FRred", "~OLbold~RS", "what", "shit", "FUCK", "a/~b", "ok", "line\n", "~", "x" * 40]
    texts, kind, target, exc, fl = [], [], [], [], []
    for _ in range(N):
        texts.append((" ".join(rng.choice(words) for _ in range(rng.randint(1, 6))) + "\n").encode())
        k = rng.choice([0, 1, 1, 1, 1, 2])
        kind.append(k)
        if k == 0:
            target.append(rng.choice(real)); exc.append(-1); fl.append(0)
        elif k == 1:
            target.append(rng.choice([-1] + list(range(NR)) * 3)); exc.append(rng.randint(-1, U - 1)); fl.append(rng.choice([0, 0, 1, 2]))
        else:
            target.append(rng.randint(0, 4)); exc.append(rng.randint(-1, U - 1)); fl.append(rng.choice([0, 4]))
    text, off = O.pack(texts)
    ops = dict(text=text, off=off, kind=np.array(kind, np.uint8), target=np.array(target, np.int32),
               except_user=np.array(exc, np.int32), flags=np.array(fl, np.uint8))
    return dict(users=dict(room=room, flags=flags, level=level), owner=owner, hear=hear, n_rooms=NR, ops=ops, texts=texts)


def port_streams(port, c):
    w = port._words(STOCK)
    port.lib.orc_set_clones(c["owner"].ctypes.data_as(C.c_void_p), c["hear"].ctypes.data_as(C.c_void_p), w)
    try:
        return port.write_batch(c["ops"], c["users"])
    finally:
        port.lib.orc_set_clones(None, None, None)


def test_clone_relay_oracle_vs_reference(port, ref):
    for seed in (31, 32, 33):
        c = make_case(seed, 24, 3, 220)
        off, data, nd = port_streams(port, c)
        ref.reset(c["n_rooms"], c["users"])
        ref.set_swear_words(STOCK[:-1])
        for u in range(24):
            if c["owner"][u] >= 0:
                ref.lib.ref_set_clone(u, int(c["owner"][u]), int(c["hear"][u]))
        o = c["ops"]
        ref.lib.ref_write_batch(len(o["kind"]), O._ptr(o["text"], O.u8p), O._ptr(o["off"], O.u64p), O._ptr(o["kind"], O.u8p),
                                O._ptr(o["target"], O.i32p), O._ptr(o["except_user"], O.i32p), O._ptr(o["flags"], O.u8p), None, None)
        for u in range(24):
            exp = b"" if c["owner"][u] >= 0 else ref.stream(u)     # what the reference "writes" to a clone's socket (-1) is lost
            assert data[int(off[u]):int(off[u + 1])].tobytes() == exp, (seed, u)


def _check_queue_tier(ctx, port, seed, U, NR, N):
    c = make_case(seed, U, NR, N)
    off, data, nd = port_streams(port, c)
    ctx.set_swear_words(STOCK)
    ctx.set_users(c["users"]["room"], c["users"]["flags"], c["users"]["level"], c["n_rooms"])
    ctx.set_clones(c["owner"], c["hear"])
    t = api.Talker(ctx)
    o = c["ops"]
    for i, s in enumerate(c["texts"]):
        k, tg, ex, f = int(o["kind"][i]), int(o["target"][i]), int(o["except_user"][i]), int(o["flags"][i])
        if k == 0:
            t.write_user(tg, s)
        elif k == 1:
            t.force_listen, t.com_num = f & 1, (api.SHOUT if f & 2 else -1)
            t.write_room_except(None if tg < 0 else tg, s, None if ex < 0 else ex)
        else:
            t.write_level(tg, bool(f & 4), s, None if ex < 0 else ex)
    st = t.flush()
    assert (st.off == off).all() and (st.data == data).all()
    # the host-buffer batch tier makes the same relays (it expands every op as the queue tier does) ...
    st = ctx.write_batch(o)
    assert (st.off == off).all() and (st.data == data).all()
    iv = ctx.write_batch_iov(o)
    assert (iv.off == off).all() and all(iv.user(u) == st.user(u) for u in range(U))
    # ... flagged clones without their owner table are refused, not silently skipped
    ctx.set_users(c["users"]["room"], c["users"]["flags"], c["users"]["level"], c["n_rooms"])
    with pytest.raises(api.NutsbError) as e:
        ctx.write_batch(o)
    assert e.value.code == api.E_STATE
    with pytest.raises(api.NutsbError) as e:
        t.write_room_except(0, b"x\n", None)
    assert e.value.code == api.E_STATE and t.pending() == 0
    ctx.set_clones(c["owner"], c["hear"])


def test_clone_relay_on_emulator(sim_lib, port):
    ctx = api.Context(0, sim_lib)
    _check_queue_tier(ctx, port, 34, 16, 3, 120)
    # a call whose relay would not fit text[] queues nothing at all (not the op, not the relays made before)
    ctx.set_users(np.array([0, 0, 0], np.int32), np.array([0x10, 0, 0x10], np.uint8), np.ones(3, np.uint8), 1)
    ctx.set_clones(np.array([1, -1, 1], np.int32), np.array([2, 0, 2], np.uint8))
    t = api.Talker(ctx)
    t.write_room_except(0, b"ok\n", None)
    n0 = t.pending()
    with pytest.raises(api.NutsbError) as e:
        t.write_room_except(0, b"y" * 1995, None)
    assert e.value.code == api.E_RANGE and t.pending() == n0
    assert t.flush().user(1).count(b"ok") == 3
    with pytest.raises(api.NutsbError):
        ctx.set_clones(np.full(16, -1, np.int32), np.zeros(16, np.uint8))      # disagrees with the CLONE flags
    ctx.close()


@pytest.mark.gpu
def test_clone_relay_on_gpu(gpu_ctx, port):
    _check_queue_tier(gpu_ctx, port, 35, 60, 4, 900)
    _check_queue_tier(gpu_ctx, port, 36, 300, 2, 2500)
