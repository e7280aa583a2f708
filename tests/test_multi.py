"""nutsb_multi: one population + one batch over several shards == the same batch on one context == the
oracle.  The batches hold what the sharding has to get right: shout / bcast (write_room_except(NULL, ...), replicated
to every shard), write_level (replicated), recipients behind filters (login / ignall / ignshout), users in no room,
excluded users who live on another shard, gated ops.  Shards on one device exercise the same host logic as shards on
several (the emulator here; with `-m gpu` two shards on cuda:0, and cuda:0 + cuda:1 when the box has two)."""
import random

import numpy as np
import pytest

import oracle_lib as O
from nuts333_b200 import api


def make_case(seed, U, NR, N):
    rng = random.Random(seed)
    room = np.array([rng.choice([-1] + list(range(NR)) * 6) for _ in range(U)], np.int32)
    flags = np.array([rng.choice([0, 1, 1, 0, 4, 8, 1, 0, 2, 5, 9]) for _ in range(U)], np.uint8)
    level = np.array([rng.randint(0, 4) for _ in range(U)], np.uint8)
    words = ["hello", "~FRred", "~OLbold~RS", "what", "a/~b", "ok", "line\n", "~", "x" * 40, "~FX", "/~FG"]
    texts, kind, target, exc, fl, gate = [], [], [], [], [], []
    for i in range(N):
        texts.append((" ".join(rng.choice(words) for _ in range(rng.randint(1, 7))) + "\n").encode())
        k = rng.choice([0, 0, 1, 1, 1, 1, 2])
        kind.append(k)
        if k == 0:
            target.append(rng.randint(-1, U - 1)); exc.append(-1); fl.append(0)
        elif k == 1:
            target.append(rng.choice([-1, -1] + list(range(NR)) * 2)); exc.append(rng.randint(-1, U - 1)); fl.append(rng.choice([0, 0, 1, 2, 3]))
        else:
            target.append(rng.randint(0, 4)); exc.append(rng.randint(-1, U - 1)); fl.append(rng.choice([0, 4]))
        gate.append(rng.choice([-1, -1, i % 50]))
        if gate[-1] >= 0 and rng.random() < 0.5: fl[-1] |= 8
    text, off = O.pack(texts)
    ops = dict(text=text, off=off, kind=np.array(kind, np.uint8), target=np.array(target, np.int32),
               except_user=np.array(exc, np.int32), flags=np.array(fl, np.uint8), gate=np.array(gate, np.int32),
               verdict=np.array([rng.randint(0, 1) for _ in range(50)], np.uint8))
    return dict(users=dict(room=room, flags=flags, level=level), n_rooms=NR, ops=ops)


def _check(lib, port, devices, seed, U, NR, N, one_ctx=None):
    c = make_case(seed, U, NR, N)
    o, us = c["ops"], c["users"]
    off, data, nd = port.write_batch(o, us, verdict=o["verdict"])
    want = [data[int(off[u]):int(off[u + 1])].tobytes() for u in range(U)]
    m = api.MultiContext(devices, lib)
    m.set_users(us["room"], us["flags"], us["level"], NR)
    rs, ush, ul = m.plan()
    assert set(ush.tolist()) <= set(range(len(devices))) and (len(devices) == 1 or len(set(rs.tolist())) > 1)
    assert all(ush[u] == rs[us["room"][u]] for u in range(U) if us["room"][u] >= 0)      # a user lives with his room
    got, total, deliv = m.write_batch(o)
    for u in range(U):
        assert got[u] == want[u], (devices, u)
    assert total == int(off[-1]) and deliv == int(nd.sum())
    # the same kept in HBM: per-user digests in global order == the digests of the one-context run
    _, total2, _ = m.write_batch(o, keep=True)
    dg = m.stream_digests()
    assert total2 == total
    if one_ctx is not None:
        one_ctx.set_users(us["room"], us["flags"], us["level"], NR)
        st = one_ctx.write_batch(o)
        assert all(st.user(u) == want[u] for u in range(U))
        assert (one_ctx.stream_digests() == dg).all()
        # a job in two chunks: the digests fold on
        half = N // 2
        sl = lambda a, b: dict(text=o["text"], off=o["off"][a:b + 1], kind=o["kind"][a:b], target=o["target"][a:b], except_user=o["except_user"][a:b],
                               flags=o["flags"][a:b], gate=o["gate"][a:b], verdict=o["verdict"])
        m.write_batch(sl(0, half), keep=True); d1 = m.stream_digests()
        m.write_batch(sl(half, N), keep=True); d2 = m.stream_digests(d1)
        assert (d2 == dg).all()
    # verdict batches split over the shards
    strs = [("site%d.evil.com" % i if i % 3 == 0 else "host%d.good.org" % i).encode() for i in range(41)]
    bt, bo = O.pack(strs)
    m.set_ban_files(b"evil.com\n", b"Troll\n")
    assert (m.site_banned_batch(bt, bo) == np.array([i % 3 == 0 for i in range(41)], np.uint8)).all()
    m.close()


def test_multi_on_emulator(sim_lib, port):
    one = api.Context(0, sim_lib)
    _check(sim_lib, port, [0, 0, 0], 61, 30, 5, 140, one_ctx=one)
    _check(sim_lib, port, [0], 62, 12, 2, 60)
    one.close()


def test_multi_rank_view_on_emulator(sim_lib, port):
    """one process per GPU: every rank plans and routes alike, runs its own shard, and the ranks' users partition the population"""
    c = make_case(63, 24, 4, 90)
    o, us = c["ops"], c["users"]
    off, data, _ = port.write_batch(o, us, verdict=o["verdict"])
    seen = np.zeros(24, np.int32)
    for rank in range(2):
        m = api.MultiContext(lib=sim_lib, rank=(2, rank, 0))
        m.set_users(us["room"], us["flags"], us["level"], 4)
        _, ush, _ = m.plan()
        got, _, _ = m.write_batch(o)
        for u in range(24):
            if ush[u] == rank:
                seen[u] += 1
                assert got[u] == data[int(off[u]):int(off[u + 1])].tobytes()
            else:
                assert got[u] == b""
        # the routed ops of the rank's shard, run on a plain context, give the same streams in local order
        r = m.route(o, rank)
        assert len(r["kind"]) <= len(o["kind"])
        m.close()
    assert (seen == 1).all()


def test_multi_refuses_clones(sim_lib):
    m = api.MultiContext([0, 0], sim_lib)
    with pytest.raises(api.NutsbError) as e:
        m.set_users(np.zeros(2, np.int32), np.array([0x10, 0], np.uint8), np.ones(2, np.uint8), 1)
    assert e.value.code == api.E_UNSUPPORTED
    m.close()


@pytest.mark.gpu
def test_multi_on_gpu(gpu_ctx, port):
    import ctypes
    lib = gpu_ctx.lib
    _check(lib, port, [0, 0], 64, 400, 7, 3000, one_ctx=gpu_ctx)
    _check(lib, port, [0, 0, 0, 0], 65, 1500, 40, 6000, one_ctx=gpu_ctx)
    import torch
    n = torch.cuda.device_count()
    if n >= 2:
        _check(lib, port, list(range(min(n, 8))), 66, 3000, 64, 20000, one_ctx=gpu_ctx)
