"""Netlink framing (nuts333.c:1299-1307): what write_user hands to write_sock for a
REMOTE_TYPE user -- "MSG <name>\\n<str>[\\n]EMSG\\n", colour commands stripped for a peer older than 3.2,
unrendered -- lands in the stream of the pseudo-user that stands for the netlink socket, in call order and,
within one call, in user-list order.  Mixed with clones (a clone's owner may itself be remote).

  * the oracle restatement against the reference's OWN write_user / write_room_except / write_level with
    real REMOTE_TYPE users on real netlink objects (write_sock's write(2) captured per socket),
  * the queue tier with nutsb_set_remotes (+ nutsb_set_clones) against the oracle: emulator here, GPU with `-m gpu`.
"""
import ctypes as C
import random

import numpy as np
import pytest

import oracle_lib as O
from nuts333_b200 import api

STOCK = ["fuck", "shit", "cunt", "*"]


def make_case(seed, U, NR, N, n_links=2):
    rng = random.Random(seed)
    room = np.array([rng.randint(0, NR - 1) for _ in range(U)], np.int32)
    flags = np.array([rng.choice([0, 1, 1, 0, 4, 8, 1, 0]) for _ in range(U)], np.uint8)
    level = np.array([rng.randint(0, 4) for _ in range(U)], np.uint8)
    names = [("U%c%c" % (chr(97 + u // 26), chr(97 + u % 26)) + "".join(rng.choice("xyz") for _ in range(rng.randint(0, 5)))).encode() for u in range(U)]
    link = np.full(U, -1, np.int32); old = np.zeros(U, np.uint8)
    owner = np.full(U, -1, np.int32); hear = np.zeros(U, np.uint8)
    links = list(range(n_links))                      # the first users stand for the netlink sockets
    for l in links:
        room[l] = -1; flags[l] = api.UF_LOGIN         # in no room, and nothing reaches them by itself
    rest = list(range(n_links, U))
    remote = [u for u in rest if rng.random() < 0.3]
    real = [u for u in rest if u not in remote]
    clones = [u for u in real if rng.random() < 0.2]
    real = [u for u in real if u not in clones] or [rest[0]]
    link_old = [l % 2 == 0 for l in links]            # the peer's version belongs to the netlink, not to the user
    for u in remote:
        flags[u] = (int(flags[u]) & 0xfe) | api.UF_REMOTE; link[u] = rng.choice(links); old[u] = link_old[link[u]]
    for u in clones:
        flags[u] |= api.UF_CLONE; owner[u] = rng.choice(real + remote); hear[u] = rng.choice([1, 2, 2])
    words = ["hello", "~FRred", "~OLbold~RS", "what", "shit", "a/~b", "ok", "~", "x" * 30, "~FX"]
    texts, kind, target, exc, fl = [], [], [], [], []
    for _ in range(N):
        t = " ".join(rng.choice(words) for _ in range(rng.randint(1, 6)))
        texts.append((t + ("\n" if rng.random() < 0.7 else "")).encode())
        k = rng.choice([0, 0, 1, 1, 1, 2])
        kind.append(k)
        if k == 0:
            target.append(rng.choice(real + remote)); exc.append(-1); fl.append(0)
        elif k == 1:
            target.append(rng.choice([-1] + list(range(NR)) * 3)); exc.append(rng.randint(-1, U - 1)); fl.append(rng.choice([0, 0, 1, 2]))
        else:
            target.append(rng.randint(0, 4)); exc.append(rng.randint(-1, U - 1)); fl.append(rng.choice([0, 4]))
    text, off = O.pack(texts)
    ops = dict(text=text, off=off, kind=np.array(kind, np.uint8), target=np.array(target, np.int32),
               except_user=np.array(exc, np.int32), flags=np.array(fl, np.uint8))
    return dict(users=dict(room=room, flags=flags, level=level), names=names, link=link, old=old, owner=owner, hear=hear,
                n_rooms=NR, ops=ops, texts=texts, U=U)


def port_streams(port, c):
    nt, no = O.pack(c["names"])
    w = port._words(STOCK)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    port.lib.orc_set_clones(vp(c["owner"]), vp(c["hear"]), w)
    port.lib.orc_set_remotes(vp(c["link"]), vp(c["old"]), vp(nt), vp(no))
    try:
        return port.write_batch(c["ops"], c["users"])
    finally:
        port.lib.orc_set_clones(None, None, None)
        port.lib.orc_set_remotes(None, None, None, None)


def test_remote_oracle_vs_reference(port, ref):
    for seed in (51, 52, 53):
        c = make_case(seed, 26, 3, 240)
        off, data, nd = port_streams(port, c)
        ref.reset(c["n_rooms"], c["users"])
        ref.set_swear_words(STOCK[:-1])
        for u in range(c["U"]):
            ref.lib.ref_set_user_speech(u, c["names"][u], 1, 0)
            if c["owner"][u] >= 0:
                ref.lib.ref_set_clone(u, int(c["owner"][u]), int(c["hear"][u]))
            if c["link"][u] >= 0:
                ref.lib.ref_set_remote(u, int(c["link"][u]), int(c["old"][u]))
        o = c["ops"]
        ref.lib.ref_write_batch(len(o["kind"]), O._ptr(o["text"], O.u8p), O._ptr(o["off"], O.u64p), O._ptr(o["kind"], O.u8p),
                                O._ptr(o["target"], O.i32p), O._ptr(o["except_user"], O.i32p), O._ptr(o["flags"], O.u8p), None, None)
        for u in range(c["U"]):
            if c["owner"][u] >= 0 or c["link"][u] >= 0:
                continue                                   # no socket of their own
            assert data[int(off[u]):int(off[u + 1])].tobytes() == ref.stream(u), (seed, u)


def _check_queue_tier(ctx, port, seed, U, NR, N):
    c = make_case(seed, U, NR, N)
    off, data, nd = port_streams(port, c)
    ctx.set_swear_words(STOCK)
    ctx.set_users(c["users"]["room"], c["users"]["flags"], c["users"]["level"], c["n_rooms"])
    ctx.set_user_names(c["names"], np.zeros(U, np.uint8))
    ctx.set_clones(c["owner"], c["hear"])
    ctx.set_remotes(c["link"], c["old"])
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
    # the host-buffer batch tier frames and relays like the queue tier; the device tier refuses the population
    st = ctx.write_batch(o)
    assert (st.off == off).all() and (st.data == data).all()
    iv = ctx.write_batch_iov(o)
    assert (iv.off == off).all() and all(iv.user(u) == st.user(u) for u in range(U))
    with pytest.raises(api.NutsbError) as e:
        ctx.write_batch_dev(0, 0, 0, 0, 0, 0, 0)
    assert e.value.code == api.E_UNSUPPORTED


def test_remote_on_emulator(sim_lib, port):
    ctx = api.Context(0, sim_lib)
    _check_queue_tier(ctx, port, 54, 18, 3, 140)
    ctx.close()


@pytest.mark.gpu
def test_remote_on_gpu(gpu_ctx, port):
    _check_queue_tier(gpu_ctx, port, 55, 60, 4, 900)
    _check_queue_tier(gpu_ctx, port, 56, 200, 2, 2000)
