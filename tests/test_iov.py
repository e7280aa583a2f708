"""Gather lists (nutsb_write_batch_iov): user u's pieces, concatenated, are byte for byte the stream the
oracle gives for u (= what write_user / write_room_except, nuts333.c:1291-1429, would have written to
u's socket), and the same as nutsb_write_batch returns.  The bodies run on the SIMT emulator (CPU tier)
and on the device (-m gpu)."""
import hashlib
import random

import numpy as np
import pytest

import oracle_lib as O
from golden_util import golden
from nuts333_b200 import api, synth

ALPHA = [bytes([c]) for c in b"abcdefghijklmnopqrstuvwxyzRSOLFBKGTWYMUIV ~~~~//\n\n"]


def _random_batch(rng, U, NR, N, maxlen, simple):
    room = np.array([rng.randint(0 if rng.random() < 0.5 else -1, NR - 1) for _ in range(U)], np.int32)
    flags = np.array([rng.choice([0, 1] if simple else [0, 1, 1, 0, 2, 4, 8, 5, 9]) for _ in range(U)], np.uint8)
    level = np.array([rng.randint(0, 4) for _ in range(U)], np.uint8)
    texts, kind, target, exc, fl = [], [], [], [], []
    for i in range(N):
        k = rng.choice([0, 0, 1, 1, 1] if simple else [0, 1, 1, 1, 1, 2])
        n = rng.randint(0, maxlen) if rng.random() < 0.9 else 0
        texts.append(b"".join(rng.choice(ALPHA + [b"~FR", b"~RS", b"~OL", b"word ", b"/~", b"\xfe"]) for _ in range(n))[:2000])
        kind.append(k)
        f = rng.choice([0, 0, 0, api.OF_PAGER, api.OF_PLAIN])
        if k == 0:
            target.append(rng.randint(-1, U - 1)); exc.append(-1); fl.append(f)
        elif k == 1:
            target.append(rng.randint(-1, NR - 1)); exc.append(rng.randint(-1, U - 1)); fl.append(f | rng.choice([0, 0, 1, 2]))
        else:
            target.append(rng.randint(0, 4)); exc.append(rng.randint(-1, U - 1)); fl.append(f | rng.choice([0, 4]))
    text, off = O.pack(texts)
    ops = dict(text=text, off=off, kind=np.array(kind, np.uint8), target=np.array(target, np.int32),
               except_user=np.array(exc, np.int32), flags=np.array(fl, np.uint8))
    return ops, dict(room=room, flags=flags, level=level)


def _check(ctx, port, ops, users, n_rooms, verdict=None, expect_compact=None, again=True):
    ctx.set_users(users["room"], users["flags"], users["level"], n_rooms)
    full = dict(ops, verdict=verdict) if verdict is not None else ops
    iv = ctx.write_batch_iov(full)
    o, d, nd = port.write_batch(ops, users, verdict=verdict)
    U = len(users["room"])
    assert iv.n_users == U and (iv.off == o).all() and iv.n_deliveries == int(nd.sum())
    for u in range(U):
        p = iv.pieces(u)
        assert int(p[:, 1].sum()) == int(o[u + 1] - o[u]), u
        for a, n in p:                                  # every piece lies inside one of the pools
            assert n == 0 or any(lo <= int(a) and int(a) + int(n) <= lo + sz for lo, sz in iv.pools), u
        assert iv.user(u) == d[int(o[u]):int(o[u + 1])].tobytes(), u
    if expect_compact is False:                         # recipients behind filters: the streams, one piece per user
        assert iv.n_iov == U and (iv.count == 1).all()
    if expect_compact and U and int(o[-1]):
        assert (iv.count % 2 == 1).all() and iv.n_iov == int(iv.count.sum())
    # ... and the streams call still gives the same bytes afterwards (shared scratch is not left dirty)
    if again:
        st = ctx.write_batch(full)
        assert (st.off == o).all() and (st.data == d).all() and st.n_deliveries == int(nd.sum())
    return iv


def _body_say_pipeline(ctx, port, n_users, per_room, n_msgs):
    words = synth.swear_words(64)
    ctx.set_swear_words(words)
    us, n_rooms = synth.users(n_users, per_room)
    bt, bo = synth.bodies(n_msgs, words)
    v = ctx.contains_swearing_batch(bt, bo)
    ops, spk, rm = synth.say_ops(n_msgs, n_users, per_room, bt, bo, gated=True)
    iv = _check(ctx, port, ops, us, n_rooms, verdict=v, expect_compact=True)
    # plain listeners: the pool is far smaller than the streams it describes
    assert iv.pool_bytes < iv.total_bytes / 4
    ctx.set_swear_words(["fuck", "shit", "cunt", "*"])


def _body_random(ctx, port, seeds, sizes, again=True):
    for seed in seeds:
        rng = random.Random(seed)
        U, NR, N, maxlen = rng.choice(sizes)
        simple = seed % 3 != 0
        ops, users = _random_batch(rng, U, NR, N, maxlen, simple)
        plain = not (users["flags"] & 0x3e).any() and not (ops["kind"] == 2).any()      # every recipient a plain listener
        assert plain or not simple
        _check(ctx, port, ops, users, NR, expect_compact=plain, again=again)


def _body_long_strings(ctx, port, n_texts):
    """strings of up to 2000 bytes dense in newlines / commands: renderings of up to 12 KB cross the
    renderer's shared-memory windows in k_direct_compact and k_render; plain listeners, so the lists are compact"""
    rng = np.random.default_rng(11)
    alphabet = np.frombuffer(b"~/\n" + b"FRSOLKBGTWYMUIV" + b"xy z", np.uint8)
    texts = []
    for i in range(n_texts):
        n = int(rng.integers(0, 2001)) if i % 3 else int(rng.integers(0, 40))
        w = np.ones(len(alphabet)); w[:3] = (8, 2, 6) if i % 2 else (1, 1, 1)
        texts.append(rng.choice(alphabet, size=n, p=w / w.sum()).tobytes())
    texts[1] = b"\n" * 2000
    texts[2] = b"~FR" * 666
    text, off = O.pack(texts)
    n = len(texts)
    kind = rng.integers(0, 2, n).astype(np.uint8)
    ops = dict(text=text, off=off, kind=kind,
               target=np.where(kind == 0, rng.integers(0, 7, n), rng.integers(-1, 2, n)).astype(np.int32),
               except_user=rng.integers(-1, 7, n).astype(np.int32), flags=np.zeros(n, np.uint8))
    users = dict(room=np.array([0, 0, 0, 1, 1, 0, -1], np.int32), flags=np.array([1, 0, 1, 0, 1, 1, 0], np.uint8),
                 level=np.ones(7, np.uint8))
    _check(ctx, port, ops, users, 2, expect_compact=True, again=False)


def _body_edges(ctx, port):
    users = dict(room=np.array([0, 0, -1, 1], np.int32), flags=np.array([1, 0, 1, 0], np.uint8), level=np.ones(4, np.uint8))
    e = dict(text=np.zeros(0, np.uint8), off=np.zeros(1, np.uint64), kind=np.zeros(0, np.uint8), target=np.zeros(0, np.int32),
             except_user=np.zeros(0, np.int32), flags=np.zeros(0, np.uint8))
    iv = _check(ctx, port, e, users, 2)                          # empty batch
    assert iv.total_bytes == 0 and all(iv.user(u) == b"" for u in range(4))
    # direct ops only (a user in no room among them), then room ops only, then two direct ops before one room op
    for texts, kind, target, exc in (
            ([b"~FRhello\n", b"", b"x"], [0, 0, 0], [2, 0, 1], [-1, -1, -1]),
            ([b"~OLroom line\n", b"all\n"], [1, 1], [0, -1], [1, 3]),
            ([b"a\n", b"b\n", b"~FGc\n", b"d"], [0, 0, 1, 0], [0, 0, 0, 0], [-1, -1, 0, -1])):
        text, off = O.pack(texts)
        ops = dict(text=text, off=off, kind=np.array(kind, np.uint8), target=np.array(target, np.int32),
                   except_user=np.array(exc, np.int32), flags=np.zeros(len(kind), np.uint8))
        _check(ctx, port, ops, users, 2, expect_compact=True)
    # the streams were never built in HBM: nothing to digest (the streams call builds them again)
    ctx.set_users(users["room"], users["flags"], users["level"], 2)
    ctx.write_batch_iov(ops)
    with pytest.raises(api.NutsbError) as e:
        ctx.stream_digests()
    assert e.value.code == api.E_STATE
    ctx.write_batch(ops)
    assert len(ctx.stream_digests()) == 4
    # no users at all
    ctx.set_users(np.zeros(0, np.int32), np.zeros(0, np.uint8), np.zeros(0, np.uint8), 2)
    iv = ctx.write_batch_iov(dict(ops, kind=np.ones(4, np.uint8), target=np.array([0, 1, -1, 0], np.int32),
                                  except_user=np.full(4, -1, np.int32)))
    assert iv.n_users == 0 and iv.total_bytes == 0 and iv.n_iov == 0


def _body_c1_through_the_queue_tier(ctx, n):
    """BASELINE config 1 through the reference's call surface (write_room per line), flushed into gather
    lists; the SHA-256 of the gathered stream is the one minted from the unmodified reference."""
    for colour in (0, 1):
        ctx.set_users(np.zeros(1, np.int32), np.array([colour], np.uint8), np.array([api.GOD], np.uint8), 1)
        t = api.Talker(ctx)
        for i in range(n):
            t.write_room(0, b"Fred says: ~OLline %04d ~FRred~RS done\n" % i)
        iv = t.flush_iov()
        assert t.pending() == 0 and iv.n_deliveries == n
        if n == 1000:
            g = golden()["c1"][str(colour)]
            assert iv.total_bytes == g["n"] and hashlib.sha256(iv.user(0)).hexdigest() == g["sha256"]
        else:
            assert iv.total_bytes == n * (52 if colour else 31)
        # speech through the queue tier: say() = write_user + write_room_except (c:4094-4098)
    ctx.set_users(np.zeros(3, np.int32), np.array([1, 0, 1], np.uint8), np.full(3, api.GOD, np.uint8), 1)
    ctx.set_user_names([b"Fred", b"Wilma", b"Barney"], np.zeros(3, np.uint8))
    t = api.Talker(ctx)
    for i in range(5):
        t.say(i % 3, b"hello ~FRthere~RS %d" % i)
    a = t.flush_iov()
    got = [a.user(u) for u in range(3)]                  # the pieces point into the context's pool: gather before the next batch
    for i in range(5):
        t.say(i % 3, b"hello ~FRthere~RS %d" % i)
    b = t.flush()
    assert (a.off == b.off).all() and a.n_deliveries == b.n_deliveries
    for u in range(3):
        assert got[u] == b.user(u) and len(got[u]) > 0


# ---- CPU tier: the product's sources on the SIMT emulator ------------------------------------
def test_iov_say_pipeline_sim(sim_lib, port):
    ctx = api.Context(0, sim_lib)
    _body_say_pipeline(ctx, port, 60, 20, 150)
    ctx.close()


def test_iov_random_and_edges_sim(sim_lib, port):
    ctx = api.Context(0, sim_lib)
    _body_edges(ctx, port)
    _body_random(ctx, port, range(3), [(5, 1, 40, 30), (40, 3, 100, 40), (33, 2, 120, 12)], again=False)
    ctx.close()


def test_iov_long_strings_sim(sim_lib, port):
    ctx = api.Context(0, sim_lib)
    _body_long_strings(ctx, port, 24)
    ctx.close()


def test_iov_c1_queue_tier_sim(sim_lib):
    ctx = api.Context(0, sim_lib)
    _body_c1_through_the_queue_tier(ctx, 40)
    ctx.close()


# ---- device tier --------------------------------------------------------------------------------
@pytest.mark.gpu
def test_iov_say_pipeline_gpu(gpu_ctx, port):
    _body_say_pipeline(gpu_ctx, port, 2000, 100, 20000)


@pytest.mark.gpu
def test_iov_random_and_edges_gpu(gpu_ctx, port):
    _body_edges(gpu_ctx, port)
    _body_random(gpu_ctx, port, range(200, 230),
                 [(1, 1, 5, 20), (40, 3, 400, 40), (300, 2, 900, 30), (1000, 1, 300, 60), (64, 70, 5000, 25),
                  (7, 1, 2000, 1990), (129, 5, 257, 400)])


@pytest.mark.gpu
def test_iov_long_strings_gpu(gpu_ctx, port):
    _body_long_strings(gpu_ctx, port, 300)


@pytest.mark.gpu
def test_iov_c1_queue_tier_gpu(gpu_ctx):
    _body_c1_through_the_queue_tier(gpu_ctx, 1000)
