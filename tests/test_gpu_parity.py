"""Parity of the CUDA path, through the C-ABI (libnutsb200.so), against

  * tests/golden/golden.json   -- vectors minted from the unmodified reference,
  * oracle/liboracle.so        -- the C restatement, on the same seeded inputs,
  * size-independent properties at BASELINE.json's full sizes.

Bit-exact: every comparison is on bytes / 0-1 verdicts.  Nothing here reads
/root/reference.
"""
import hashlib
import random

import numpy as np
import pytest

import oracle_lib as O
from golden_util import golden, kat_render_batch, mixed_batch
from nuts333_b200 import api, synth

pytestmark = pytest.mark.gpu

ALPHA = [b"~", b"/", b"\n", b"F", b"R", b"S", b"O", b"L", b"K", b"B", b"G", b"T", b"W", b"Y", b"M", b"U", b"I", b"V",
         b"x", b"y", b" "]


def _check_streams(st, off, data):
    assert (st.off == off).all()
    assert st.data.shape == data.shape and (st.data == data).all()


def test_render_kat_vectors(gpu_ctx):
    ops, users, exp = kat_render_batch()
    gpu_ctx.set_users(users["room"], users["flags"], users["level"], 1)
    st = gpu_ctx.write_batch(ops)
    assert st.user(0) == exp[0]
    assert st.user(1) == exp[1]


def test_render_fuzz_vectors(gpu_ctx):
    g = golden()["render_fuzz"]
    strings = [bytes.fromhex(v["s"]) for v in g if b"\0" not in bytes.fromhex(v["s"])]
    # one user pair per string so every rendering is its own stream
    n = len(strings)
    text, off = O.pack([s for s in strings for _ in (0, 1)])
    ops = dict(text=text, off=off, kind=np.zeros(2 * n, np.uint8), target=np.arange(2 * n, dtype=np.int32),
               except_user=np.full(2 * n, -1, np.int32), flags=np.zeros(2 * n, np.uint8))
    gpu_ctx.set_users(np.zeros(2 * n, np.int32), np.tile(np.array([0, 1], np.uint8), n), np.ones(2 * n, np.uint8), 1)
    st = gpu_ctx.write_batch(ops)
    for i, v in enumerate(g):
        a, b = st.user(2 * i), st.user(2 * i + 1)
        assert (len(a), len(b)) == (v["n0"], v["n1"])
        assert hashlib.sha256(a).hexdigest() == v["c0"] and hashlib.sha256(b).hexdigest() == v["c1"]


def test_mixed_batch_reference_digests(gpu_ctx, port):
    ops, users, n_rooms, verdict, lens, shas = mixed_batch()
    gpu_ctx.set_users(users["room"], users["flags"], users["level"], n_rooms)
    st = gpu_ctx.write_batch(dict(ops, verdict=verdict))
    for u in range(len(lens)):
        s = st.user(u)
        assert len(s) == lens[u] and hashlib.sha256(s).hexdigest() == shas[u], u
    _, _, nd = port.write_batch(ops, users, verdict=verdict)
    assert st.n_deliveries == int(nd.sum())


def test_c1_config_through_the_reference_surface(gpu_ctx):
    """BASELINE config 1: 1,000 write_room('say' lines with ~OL/~FR/~RS), colour off and on,
    issued one call at a time through the write_room() mirror."""
    for colour in (0, 1):
        gpu_ctx.set_users(np.zeros(1, np.int32), np.array([colour], np.uint8), np.array([api.GOD], np.uint8), 1)
        t = api.Talker(gpu_ctx)
        for i in range(1000):
            t.write_room(0, b"Fred says: ~OLline %04d ~FRred~RS done\n" % i)
        assert t.pending() == 1000
        st = t.flush()
        g = golden()["c1"][str(colour)]
        assert st.total_bytes == g["n"] and hashlib.sha256(st.user(0)).hexdigest() == g["sha256"]
        assert st.n_deliveries == 1000 and t.pending() == 0


def test_empty_and_ragged(gpu_ctx, port):
    gpu_ctx.set_users(np.array([0, 0, -1], np.int32), np.array([1, 0, 1], np.uint8), np.ones(3, np.uint8), 1)
    users = dict(room=np.array([0, 0, -1], np.int32), flags=np.array([1, 0, 1], np.uint8), level=np.ones(3, np.uint8))
    # no ops at all
    e = dict(text=np.zeros(0, np.uint8), off=np.zeros(1, np.uint64), kind=np.zeros(0, np.uint8),
             target=np.zeros(0, np.int32), except_user=np.zeros(0, np.int32), flags=np.zeros(0, np.uint8))
    st = gpu_ctx.write_batch(e)
    assert st.total_bytes == 0 and st.n_users == 3
    # ops that reach nobody: write_user(NULL), a room op whose only listener is excluded
    texts = [b"x", b"", b"~RS", b""]
    text, off = O.pack(texts)
    ops = dict(text=text, off=off, kind=np.array([0, 1, 1, 0], np.uint8), target=np.array([-1, 0, 0, 2], np.int32),
               except_user=np.array([-1, 1, -1, -1], np.int32), flags=np.zeros(4, np.uint8))
    st = gpu_ctx.write_batch(ops)
    o, d, nd = port.write_batch(ops, users)
    _check_streams(st, o, d)
    assert st.user(0) == b"\x1b[0m" + b"\x1b[0m\x1b[0m" and st.user(1) == b"" and st.user(2) == b"\x1b[0m"
    # no users
    gpu_ctx.set_users(np.zeros(0, np.int32), np.zeros(0, np.uint8), np.zeros(0, np.uint8), 2)
    st = gpu_ctx.write_batch(dict(ops, kind=np.array([1, 1, 1, 1], np.uint8), target=np.array([0, 1, -1, 0], np.int32),
                                  except_user=np.full(4, -1, np.int32)))
    assert st.total_bytes == 0 and st.n_users == 0


def test_errors_leave_the_host_alive(gpu_ctx):
    gpu_ctx.set_users(np.zeros(2, np.int32), np.zeros(2, np.uint8), np.ones(2, np.uint8), 1)
    one = lambda **k: dict(dict(text=np.frombuffer(b"x" * 2001, np.uint8), off=np.array([0, 1], np.uint64),
                                kind=np.array([1], np.uint8), target=np.array([0], np.int32),
                                except_user=np.array([-1], np.int32), flags=np.zeros(1, np.uint8)), **k)
    with pytest.raises(api.NutsbError) as e:
        gpu_ctx.write_batch(one(off=np.array([0, 2001], np.uint64)))
    assert e.value.code == api.E_RANGE
    for bad in (one(target=np.array([5], np.int32)), one(except_user=np.array([2], np.int32)),
                one(kind=np.array([0], np.uint8), target=np.array([2], np.int32))):
        with pytest.raises(api.NutsbError) as e:
            gpu_ctx.write_batch(bad)
        assert e.value.code == api.E_RANGE
    with pytest.raises(api.NutsbError) as e:
        gpu_ctx.write_batch(one(kind=np.array([9], np.uint8)))
    assert e.value.code == api.E_INVAL
    # clones / remote users without their owner / link tables: refused (their relays and frames would be missing),
    # and the device tier refuses such a population altogether
    gpu_ctx.set_users(np.zeros(3, np.int32), np.array([api.UF_REMOTE, api.UF_CLONE, 0], np.uint8), np.ones(3, np.uint8), 1)
    with pytest.raises(api.NutsbError) as e:
        gpu_ctx.write_batch(one())
    assert e.value.code == api.E_STATE
    with pytest.raises(api.NutsbError) as e:
        gpu_ctx.write_batch_dev(0, 0, 0, 0, 0, 0, 0)
    assert e.value.code == api.E_UNSUPPORTED
    with pytest.raises(api.NutsbError) as e:
        gpu_ctx.set_clones(np.array([-1, -1, -1], np.int32), np.zeros(3, np.uint8))     # disagrees with the CLONE flag
    assert e.value.code == api.E_INVAL
    # and the context still works
    gpu_ctx.set_users(np.zeros(2, np.int32), np.zeros(2, np.uint8), np.ones(2, np.uint8), 1)
    assert gpu_ctx.write_batch(one()).total_bytes == 2


@pytest.mark.parametrize("seed,U,NR,N,maxlen", [(1, 40, 3, 400, 40), (2, 300, 2, 900, 30), (3, 1000, 1, 300, 60),
                                                (4, 64, 70, 5000, 25), (5, 7, 1, 3000, 1990)])
def test_random_batches_vs_oracle(gpu_ctx, port, seed, U, NR, N, maxlen):
    """every op kind, every recipient filter, excepts, all-room ops, users in no room,
    big rooms (several recipient chunks), many rooms, strings up to the 2000-byte limit"""
    rng = random.Random(seed)
    room = np.array([rng.randint(-1, NR - 1) for _ in range(U)], np.int32)
    flags = np.array([rng.choice([0, 1, 1, 0, 1, 0, 2, 4, 8, 5, 9, 12, 13]) for _ in range(U)], np.uint8)
    level = np.array([rng.randint(0, 4) for _ in range(U)], np.uint8)
    texts, kind, target, exc, fl = [], [], [], [], []
    for i in range(N):
        k = rng.choice([0, 1, 1, 1, 1, 2])
        n = rng.randint(0, maxlen) if rng.random() < 0.9 else rng.randint(0, 8)
        s = b"".join(rng.choice(ALPHA + [b"~FR", b"~RS", b"~OL", b"word ", b"/~", b"\xfe"]) for _ in range(n))[:2000]
        texts.append(s); kind.append(k)
        if k == 0:
            target.append(rng.randint(-1, U - 1)); exc.append(-1); fl.append(0)
        elif k == 1:
            target.append(rng.randint(-1, NR - 1)); exc.append(rng.randint(-1, U - 1)); fl.append(rng.choice([0, 0, 0, 1, 2, 3]))
        else:
            target.append(rng.randint(0, 4)); exc.append(rng.randint(-1, U - 1)); fl.append(rng.choice([0, 4]))
    text, off = O.pack(texts)
    ops = dict(text=text, off=off, kind=np.array(kind, np.uint8), target=np.array(target, np.int32),
               except_user=np.array(exc, np.int32), flags=np.array(fl, np.uint8))
    users = dict(room=room, flags=flags, level=level)
    gpu_ctx.set_users(room, flags, level, NR)
    st = gpu_ctx.write_batch(ops)
    o, d, nd = port.write_batch(ops, users)
    _check_streams(st, o, d)
    assert st.n_deliveries == int(nd.sum())
    # the device digest is the fold of the bytes the host received
    dg = gpu_ctx.stream_digests()
    for u in random.Random(0).sample(range(U), min(U, 5)):
        h = 0xcbf29ce484222325
        for x in st.user(u):
            h = (h * 0x100000001b3 + x) & (2 ** 64 - 1)
        assert int(dg[u]) == h


def test_fuzz_many_shapes_vs_oracle(gpu_ctx, port):
    """40 random (population, batch) shapes: rooms of 1..600 users, empty rooms, tiles that are
    exactly full / one over, strings that force sub-tiles, all flags, pager and plain ops."""
    for seed in range(100, 140):
        rng = random.Random(seed)
        U = rng.choice([1, 2, 31, 32, 33, 127, 128, 129, 257, 600])
        NR = rng.choice([1, 1, 2, 5, 40])
        N = rng.choice([1, 127, 128, 129, 256, 257, 700, 2000])
        simple = rng.random() < 0.5
        room = np.array([rng.randint(0 if simple else -1, NR - 1) for _ in range(U)], np.int32)
        flags = np.array([rng.choice([0, 1] if simple else [0, 1, 1, 0, 2, 4, 8, 5, 9]) for _ in range(U)], np.uint8)
        level = np.array([rng.randint(0, 4) for _ in range(U)], np.uint8)
        big = rng.random() < 0.3
        texts, kind, target, exc, fl = [], [], [], [], []
        for i in range(N):
            k = rng.choice([0, 1, 1, 1, 1] if simple else [0, 1, 1, 1, 1, 2])
            n = rng.randint(0, 400 if big else 40)
            texts.append(b"".join(rng.choice(ALPHA + [b"~FR", b"~RS", b"~OL", b"word ", b"/~", b"\n"]) for _ in range(n))[:2000])
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
        users = dict(room=room, flags=flags, level=level)
        gpu_ctx.set_users(room, flags, level, NR)
        st = gpu_ctx.write_batch(ops)
        o, d, nd = port.write_batch(ops, users)
        assert (st.off == o).all() and (st.data == d).all() and st.n_deliveries == int(nd.sum()), seed


def test_simple_population_fast_path_vs_oracle(gpu_ctx, port):
    """colour-only population (no login/ignall/ignshout, no level ops): the path the
    benchmark takes, at a size the oracle finishes in seconds (config 2, scaled)."""
    us, n_rooms = synth.users(1000, 0)
    bt, bo = synth.bodies(3000)
    ops, spk, rm = synth.say_ops(3000, 1000, 0, bt, bo, gated=False)
    gpu_ctx.set_users(us["room"], us["flags"], us["level"], n_rooms)
    st = gpu_ctx.write_batch(ops)
    o, d, nd = port.write_batch(ops, us)
    _check_streams(st, o, d)
    assert st.n_deliveries == 3000 * 999


def test_say_pipeline_scaled_config3(gpu_ctx, port):
    """config 3 scaled down: 64-word swear list, say() gated on the DEVICE's verdicts."""
    words = synth.swear_words(64)
    gpu_ctx.set_swear_words(words)
    us, n_rooms = synth.users(2000, 100)
    bt, bo = synth.bodies(20000, words)
    v = gpu_ctx.contains_swearing_batch(bt, bo)
    assert (v == port.contains_swearing_batch(bt, bo, words)).all()
    ops, spk, rm = synth.say_ops(20000, 2000, 100, bt, bo, gated=True)
    gpu_ctx.set_users(us["room"], us["flags"], us["level"], n_rooms)
    st = gpu_ctx.write_batch(dict(ops, verdict=v))
    o, d, nd = port.write_batch(ops, us, verdict=v)
    _check_streams(st, o, d)
    clean = int((v == 0).sum())
    assert st.n_deliveries == clean * 100 + (len(v) - clean)
    gpu_ctx.set_swear_words(["fuck", "shit", "cunt", "*"])


def test_swear_vectors_and_list_semantics(gpu_ctx, port):
    g = golden()
    t = api.Talker(gpu_ctx)
    gpu_ctx.set_swear_words(["fuck", "shit", "cunt", "*"])
    for v in g["swear_stock"]:
        assert t.contains_swearing(bytes.fromhex(v["s"])) == v["v"]
    g64 = g["swear64"]
    gpu_ctx.set_swear_words(g64["words"])
    bt, bo = synth.bodies(g64["n"], g64["words"], seed=g64["seed"])
    v = gpu_ctx.contains_swearing_batch(bt, bo)
    assert int(v.sum()) == g64["dirty"] and hashlib.sha256(v.tobytes()).hexdigest() == g64["verdict_sha256"]
    # list ends at '*'; an empty word matches everything; upper-case words never match
    for words, s, want in ((["*", "abc"], b"abc", 0), (["", "*"], b"abc", 1), (["ABC", "*"], b"ABC abc", 0),
                           (["a", "*"], b"", 0), (["", "*"], b"", 1)):
        gpu_ctx.set_swear_words(words)
        assert t.contains_swearing(s) == want == port.contains_swearing(s, words)
    # a list too big for shared memory takes the L2-resident table
    big = ["w%04dxyz" % i for i in range(3000)] + ["*"]
    gpu_ctx.set_swear_words(big)
    strings = [b"say W0007XYZ now", b"w3000xyz", b"xw2999xyzx", b"w12", b""]
    bt2, bo2 = O.pack(strings)
    assert (gpu_ctx.contains_swearing_batch(bt2, bo2) == port.contains_swearing_batch(bt2, bo2, big)).all()
    # strings longer than a staging window
    long_s = [b"a" * 1999 + b"~", b"b" * 1500 + b"SHIT" + b"c" * 400] * 40
    gpu_ctx.set_swear_words(["fuck", "shit", "cunt", "*"])
    bt3, bo3 = O.pack(long_s)
    assert (gpu_ctx.contains_swearing_batch(bt3, bo3) == port.contains_swearing_batch(bt3, bo3, ["shit", "*"])).all()


def test_ban_vectors(gpu_ctx, port):
    b = golden()["bans"]
    t = api.Talker(gpu_ctx)
    gpu_ctx.set_ban_files(bytes.fromhex(b["site_file"]), bytes.fromhex(b["user_file"]))
    for v in b["site"]:
        assert t.site_banned(bytes.fromhex(v["q"])) == v["v"]
    for v in b["user"]:
        assert t.user_banned(bytes.fromhex(v["q"])) == v["v"]
    gpu_ctx.set_ban_files(bytes.fromhex(b["site2_file"]), b"")
    for v in b["site2"]:
        assert t.site_banned(bytes.fromhex(v["q"])) == v["v"]
    assert t.user_banned(b"Troll") == 0
    gpu_ctx.set_ban_files(b"", None)
    for v in b["empty"] + b["missing"]:
        assert t.site_banned(bytes.fromhex(v["q"])) == v["v"] == 0
    for tn in (0, 1):
        g = b["generated"][str(tn)]
        gpu_ctx.set_ban_files(synth.ban_file(0, 300, 2000, 2000, bool(tn)), synth.ban_file(1, 300, 2000, 2000, bool(tn)))
        st, so = synth.sites(2000)
        nt, no = synth.names(2000)
        vs, vu = gpu_ctx.site_banned_batch(st, so), gpu_ctx.user_banned_batch(nt, no)
        assert (int(vs.sum()), hashlib.sha256(vs.tobytes()).hexdigest()) == (g["site_hits"], g["site_sha256"])
        assert (int(vu.sum()), hashlib.sha256(vu.tobytes()).hexdigest()) == (g["user_hits"], g["user_sha256"])
    with pytest.raises(api.NutsbError) as e:
        gpu_ctx.set_ban_files(b"x" * 82 + b"\n", None)
    assert e.value.code == api.E_RANGE


def test_config4_bans_full_size(gpu_ctx, port):
    """100k sites + 100k names vs 10k-entry lists, with and without the trailing newline.
    Full verdict vectors are checked against the oracle on a 3,000-query sample, the
    rest through properties: user verdicts equal set membership of the tested tokens;
    dropping the trailing newline can only clear verdicts (the last token goes untested)."""
    st, so = synth.sites(100000)
    nt, no = synth.names(100000)
    res = {}
    for tn in (True, False):
        sf, uf = synth.ban_file(0, 10000, 100000, 100000, tn), synth.ban_file(1, 10000, 100000, 100000, tn)
        gpu_ctx.set_ban_files(sf, uf)
        vs, vu = gpu_ctx.site_banned_batch(st, so), gpu_ctx.user_banned_batch(nt, no)
        idx = np.random.RandomState(4).choice(100000, 3000, replace=False)
        sub_s = [st[int(so[i]):int(so[i + 1])].tobytes() for i in idx]
        sub_n = [nt[int(no[i]):int(no[i + 1])].tobytes() for i in idx]
        a, ao = O.pack(sub_s)
        c, co = O.pack(sub_n)
        assert (vs[idx] == port.ban_batch(0, sf, a, ao)).all()
        assert (vu[idx] == port.ban_batch(1, uf, c, co)).all()
        tested = set(port.ban_tokens(uf))
        member = np.array([nt[int(no[i]):int(no[i + 1])].tobytes() in tested for i in range(100000)], np.uint8)
        assert (vu == member).all()
        assert 0.005 < vs.mean() < 0.5
        res[tn] = (vs, vu)
    assert ((res[False][0] <= res[True][0]).all() and (res[False][1] <= res[True][1]).all())


def _full_size_properties(gpu_ctx, port, ops, us, n_rooms, verdict, expect_deliveries):
    gpu_ctx.set_users(us["room"], us["flags"], us["level"], n_rooms)
    st = gpu_ctx.write_batch(dict(ops, verdict=verdict) if verdict is not None else ops)
    U = len(us["room"])
    assert st.n_deliveries == expect_deliveries
    # sampled users: whole streams, byte for byte, against the oracle
    sample = sorted(np.random.RandomState(3).choice(U, 24, replace=False).tolist())
    o, d, nd = port.write_batch(ops, us, verdict=verdict, only_users=sample)
    for u in sample:
        assert st.user(u) == d[int(o[u]):int(o[u + 1])].tobytes(), u
    # every stream ends with the reference's terminal reset / CR and has no NUL byte
    assert (st.data != 0).all()
    on = (us["flags"] & 1) != 0
    lens = np.diff(st.off.astype(np.int64))
    assert (lens > 0).all()
    last4 = np.stack([st.data[(st.off[1:] - k).astype(np.int64)] for k in (4, 3, 2, 1)], 1)
    assert (last4[on] == np.frombuffer(b"\x1b[0m", np.uint8)).all()
    assert (last4[~on][:, 2:] == np.frombuffer(b"\n\r", np.uint8)).all()
    # device digests == digests of what the host received (the D2H copy is the stream)
    dg = gpu_ctx.stream_digests()
    for u in sample[:6]:
        h = 0xcbf29ce484222325
        for x in st.user(u)[:200000]:
            h = (h * 0x100000001b3 + x) & (2 ** 64 - 1)
        if len(st.user(u)) <= 200000:
            assert int(dg[u]) == h
    # idempotence: the same batch again gives the same bytes
    st2 = gpu_ctx.write_batch(dict(ops, verdict=verdict) if verdict is not None else ops)
    assert (st2.off == st.off).all() and (gpu_ctx.stream_digests() == dg).all()
    # the same batch as gather lists (nutsb_write_batch_iov): every user's pieces add up to its stream's
    # length, the sampled users' gathered bytes are the oracle's, and the pool is a fraction of the streams
    iv = gpu_ctx.write_batch_iov(dict(ops, verdict=verdict) if verdict is not None else ops)
    assert (iv.off == st.off).all() and iv.n_deliveries == expect_deliveries
    piece_user = np.repeat(np.arange(U), iv.count.astype(np.int64))
    order = (np.repeat(iv.first.astype(np.int64), iv.count.astype(np.int64))
             + np.arange(len(piece_user)) - np.repeat(np.cumsum(iv.count.astype(np.int64)) - iv.count.astype(np.int64), iv.count.astype(np.int64)))
    per_user = np.bincount(piece_user, weights=iv.iov[order, 1].astype(np.float64), minlength=U).astype(np.int64)
    assert (per_user == lens).all()
    for u in sample:
        assert iv.user(u) == d[int(o[u]):int(o[u + 1])].tobytes(), u
    assert iv.pool_bytes * 8 < st.total_bytes
    return st


def test_config2_full_size(gpu_ctx, port):
    """100k messages x 1k users in one room, colour expand/strip only (~1e8 deliveries)."""
    us, n_rooms = synth.users(1000, 0)
    bt, bo = synth.bodies(100000)
    ops, spk, rm = synth.say_ops(100000, 1000, 0, bt, bo, gated=False)
    st = _full_size_properties(gpu_ctx, port, ops, us, n_rooms, None, 100000 * 999)
    # linearity: stream lengths are sums of the two renderings' lengths minus own lines
    lens_on = np.array([len(port.render(ops["text"][int(ops["off"][i]):int(ops["off"][i + 1])].tobytes(), 1)) for i in range(0, 100000, 997)])
    assert lens_on.min() > 0
    tot_in = int(ops["off"][-1])
    assert st.total_bytes > 900 * tot_in


def test_config3_full_size(gpu_ctx, port):
    """1M messages x 10k users (100 rooms of 100), 64-word list, say() under ban_swearing."""
    words = synth.swear_words(64)
    gpu_ctx.set_swear_words(words)
    us, n_rooms = synth.users(10000, 100)
    bt, bo = synth.bodies(1000000, words)
    v = gpu_ctx.contains_swearing_batch(bt, bo)
    assert (v == port.contains_swearing_batch(bt, bo, words)).all()
    ops, spk, rm = synth.say_ops(1000000, 10000, 100, bt, bo, gated=True)
    clean = int((v == 0).sum())
    _full_size_properties(gpu_ctx, port, ops, us, n_rooms, v, clean * 100 + (1000000 - clean))
    gpu_ctx.set_swear_words(["fuck", "shit", "cunt", "*"])


def test_serial_and_overlapped_schedules_agree(gpu_ctx, port):
    """every kernel on one stream (what bench.py times kernels in) against the default schedule
    (k_render on the side stream, fan-out and direct blocks dealt over one grid)"""
    words = synth.swear_words(64)
    us, n_rooms = synth.users(3000, 100, stress=True)
    bt, bo = synth.bodies(60000, words)
    gpu_ctx.set_swear_words(words)
    v = gpu_ctx.contains_swearing_batch(bt, bo)
    sops, _, _ = synth.say_ops(60000, 3000, 100, bt, bo, gated=True)
    gpu_ctx.set_users(us["room"], us["flags"], us["level"], n_rooms)
    res = []
    for overlap in (False, True, False):
        gpu_ctx.set_overlap(overlap)
        st = gpu_ctx.write_batch(dict(sops, verdict=v))
        res.append((st.off.copy(), hashlib.sha256(st.data.tobytes()).hexdigest(), st.n_deliveries))
    gpu_ctx.set_overlap(True)
    assert (res[0][0] == res[1][0]).all() and res[0][1] == res[1][1] == res[2][1] and res[0][2] == res[1][2]
    only = np.arange(0, 3000, 97, dtype=np.int32)
    off, data, nd = port.write_batch(sops, us, verdict=v, only_users=only)
    st = gpu_ctx.write_batch(dict(sops, verdict=v))
    for u in only:
        assert st.user(int(u)) == data[int(off[u]):int(off[u + 1])].tobytes()


def _config5_pipeline(ctx, port, n_users, n_msgs, per_room, n_ban, sample):
    """BASELINE config 5, one rank's share: stage A admission (site_banned or user_banned -> never a user of the
    talker), stage B swear verdicts, stage C say() composed lines rendered and fanned out in rooms of `per_room`."""
    words = synth.swear_words(64)
    ctx.set_swear_words(words)
    sf, uf = synth.ban_file(0, n_ban, n_users, n_users, True), synth.ban_file(1, n_ban, n_users, n_users, False)
    ctx.set_ban_files(sf, uf)
    st_, so_ = synth.sites(n_users)
    nt_, no_ = synth.names(n_users)
    vs, vu = ctx.site_banned_batch(st_, so_), ctx.user_banned_batch(nt_, no_)
    assert (vs == port.ban_batch(0, sf, st_, so_)).all() and (vu == port.ban_batch(1, uf, nt_, no_)).all()
    admitted = np.nonzero((vs | vu) == 0)[0]
    assert 0 < len(admitted) < n_users                       # some connections are refused (c:279, c:1496)
    U = len(admitted) - len(admitted) % per_room              # whole rooms of admitted users
    us, n_rooms = synth.users(U, per_room)
    bt, bo = synth.bodies(n_msgs, words)
    v = ctx.contains_swearing_batch(bt, bo)
    assert (v == port.contains_swearing_batch(bt, bo, words)).all()
    ops, spk, rm = synth.say_ops(n_msgs, U, per_room, bt, bo, gated=True)
    ctx.set_users(us["room"], us["flags"], us["level"], n_rooms)
    st = ctx.write_batch(dict(ops, verdict=v))
    clean = int((v == 0).sum())
    assert st.n_deliveries == clean * per_room + (n_msgs - clean)
    pick = sorted(np.random.RandomState(5).choice(U, sample, replace=False).tolist())
    o, d, nd = port.write_batch(ops, us, verdict=v, only_users=pick)
    for u in pick:
        assert st.user(u) == d[int(o[u]):int(o[u + 1])].tobytes(), u
    ctx.set_swear_words(["fuck", "shit", "cunt", "*"])
    return st


def test_config5_pipeline_scaled(gpu_ctx, port):
    _config5_pipeline(gpu_ctx, port, 6000, 60000, 100, 600, 40)


def test_config5_pipeline_one_rank_full_size(gpu_ctx, port):
    """10M messages x 100k users over 8 ranks = 1.25M messages x 12.5k connecting users per rank"""
    st = _config5_pipeline(gpu_ctx, port, 12500, 1250000, 100, 1250, 16)
    dg = gpu_ctx.stream_digests()
    assert len(set(dg.tolist())) == len(dg)                   # every user's stream is its own
