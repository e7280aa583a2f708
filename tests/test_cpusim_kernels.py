"""Kernel LOGIC on a box without a GPU: the product's .cu sources compiled against the
SIMT emulator (tests/cpusim) and compared with the oracle.  This library is test
infrastructure; the product never loads it.  Kept small: the emulator runs every CUDA
thread as an OS thread."""
import hashlib

import numpy as np

import oracle_lib as O
from golden_util import kat_render_batch, mixed_batch
from nuts333_b200 import api, synth


def _ctx(sim_lib):
    return api.Context(0, sim_lib)


def test_mixed_batch_matches_reference_digests(sim_lib, port):
    ops, users, n_rooms, verdict, lens, shas = mixed_batch()
    ctx = _ctx(sim_lib)
    ctx.set_users(users["room"], users["flags"], users["level"], n_rooms)
    st = ctx.write_batch(dict(ops, verdict=verdict))
    for u in range(len(lens)):
        s = st.user(u)
        assert len(s) == lens[u] and hashlib.sha256(s).hexdigest() == shas[u], u
    _, _, nd = port.write_batch(ops, users, verdict=verdict)
    assert st.n_deliveries == int(nd.sum())
    ctx.close()


def test_render_kat_and_say_pipeline(sim_lib, port):
    ctx = _ctx(sim_lib)
    ops, users, exp = kat_render_batch()
    ctx.set_users(users["room"], users["flags"], users["level"], 1)
    st = ctx.write_batch(ops)
    assert st.user(0) == exp[0] and st.user(1) == exp[1]
    # say() pipeline, gated on the device's own swear verdicts
    words = synth.swear_words(64)
    ctx.set_swear_words(words)
    us, n_rooms = synth.users(60, 20)
    bt, bo = synth.bodies(150, words)
    v = ctx.contains_swearing_batch(bt, bo)
    assert (v == port.contains_swearing_batch(bt, bo, words)).all()
    sops, spk, rm = synth.say_ops(150, 60, 20, bt, bo, gated=True)
    ctx.set_users(us["room"], us["flags"], us["level"], n_rooms)
    st = ctx.write_batch(dict(sops, verdict=v))
    off, data, nd = port.write_batch(sops, us, verdict=v)
    assert (st.off == off).all() and (st.data == data).all() and st.n_deliveries == int(nd.sum())
    ctx.close()


def test_ban_verdicts(sim_lib, port):
    ctx = _ctx(sim_lib)
    for tn in (False, True):
        sf, uf = synth.ban_file(0, 200, 1000, 1000, tn), synth.ban_file(1, 200, 1000, 1000, tn)
        ctx.set_ban_files(sf, uf)
        st, so = synth.sites(1000)
        nt, no = synth.names(1000)
        assert (ctx.site_banned_batch(st, so) == port.ban_batch(0, sf, st, so)).all()
        assert (ctx.user_banned_batch(nt, no) == port.ban_batch(1, uf, nt, no)).all()
    ctx.close()


def test_long_strings_cross_the_render_windows(sim_lib, port):
    """Strings of up to 2000 bytes dense in newlines / commands: renderings of up to 12 KB, so a warp's
    shared-memory windows are flushed in the middle of a string (k_render, k_direct) and a tile's
    renderings exceed the fan-out's windows (slab -> stream copy)."""
    rng = np.random.default_rng(7)
    alphabet = np.frombuffer(b"~/\n" + b"FRSOLKBGTWYMUIV" + b"xy z", np.uint8)
    texts = []
    for i in range(90):
        n = int(rng.integers(0, 2001)) if i % 3 else int(rng.integers(0, 40))
        w = np.ones(len(alphabet)); w[:3] = (8, 2, 6) if i % 2 else (1, 1, 1)
        texts.append(rng.choice(alphabet, size=n, p=w / w.sum()).tobytes())
    texts[5] = b"\n" * 2000
    texts[11] = b"~FR" * 666
    text, off = O.pack(texts)
    n_users, n = 7, len(texts)
    kind = rng.integers(0, 2, n).astype(np.uint8)
    target = np.where(kind == 0, rng.integers(0, n_users, n), rng.integers(-1, 2, n)).astype(np.int32)
    ops = dict(text=text, off=off, kind=kind, target=target,
               except_user=rng.integers(-1, n_users, n).astype(np.int32), flags=np.zeros(n, np.uint8))
    users = dict(room=np.array([0, 0, 0, 1, 1, 0, 1], np.int32), flags=np.array([1, 0, 1, 0, 1, 5, 0], np.uint8),
                 level=np.ones(n_users, np.uint8))
    ctx = _ctx(sim_lib)
    ctx.set_users(users["room"], users["flags"], users["level"], 2)
    st = ctx.write_batch(ops)
    eoff, data, nd = port.write_batch(ops, users)
    assert (st.off == eoff).all() and (st.data == data).all() and st.n_deliveries == int(nd.sum())
    ctx.close()


def test_many_tiles_per_room_seams(sim_lib, port):
    """say() traffic long enough for several fan-out tiles per room: runs are cut at tile boundaries and at
    every speaker's exclusion / "You say" line, and k_direct's seam pass writes the sectors around the cuts."""
    ctx = _ctx(sim_lib)
    words = synth.swear_words(8)
    for n_users, per_room, n_msgs, stress in ((24, 12, 700, False), (9, 3, 500, True)):
        us, n_rooms = synth.users(n_users, per_room, stress=stress)
        bt, bo = synth.bodies(n_msgs, words)
        v = port.contains_swearing_batch(bt, bo, words)
        sops, _, _ = synth.say_ops(n_msgs, n_users, per_room, bt, bo, gated=True)
        ctx.set_users(us["room"], us["flags"], us["level"], n_rooms)
        st = ctx.write_batch(dict(sops, verdict=v))
        off, data, nd = port.write_batch(sops, us, verdict=v)
        assert (st.off == off).all() and st.n_deliveries == int(nd.sum())
        bad = np.nonzero(st.data != data)[0]
        assert bad.size == 0, (n_users, bad[:8], int(np.searchsorted(off, bad[0], side="right")) - 1)
    ctx.close()


def test_serial_and_overlapped_schedules_agree(sim_lib, port):
    """nutsb_set_overlap(0): k_render, k_direct and k_fanout one after the other on one stream (the schedule
    bench.py times kernels in); default: k_render on the side stream, k_fanout_direct in one launch."""
    ctx = _ctx(sim_lib)
    words = synth.swear_words(8)
    us, n_rooms = synth.users(40, 10, stress=True)
    bt, bo = synth.bodies(400, words)
    v = port.contains_swearing_batch(bt, bo, words)
    sops, _, _ = synth.say_ops(400, 40, 10, bt, bo, gated=True)
    off, data, nd = port.write_batch(sops, us, verdict=v)
    ctx.set_users(us["room"], us["flags"], us["level"], n_rooms)
    for overlap in (False, True):
        ctx.set_overlap(overlap)
        st = ctx.write_batch(dict(sops, verdict=v))
        assert (st.off == off).all() and (st.data == data).all() and st.n_deliveries == int(nd.sum()), overlap
    ctx.close()


def test_config5_pipeline_on_emulator(sim_lib, port):
    """admission by the ban verdicts, swear verdicts, say() fan-out: BASELINE config 5 in miniature"""
    from test_gpu_parity import _config5_pipeline
    ctx = _ctx(sim_lib)
    _config5_pipeline(ctx, port, 90, 300, 10, 12, 20)
    ctx.close()


def test_flat_renderer_fuzz_vs_oracle(sim_lib, port):
    """The position-parallel renderer against the byte machine's restatement on strings dense in everything the
    machine looks at -- '~', '/', newlines, command letters, every other byte value 1..255 -- of lengths that put
    commands across word, lane and round boundaries, as write_user ops (k_direct) and room ops (k_render)."""
    rng = np.random.default_rng(2024)
    hot = np.frombuffer(b"~/\n" + b"FRSOLKBGTWYMUIV" + b"x ", np.uint8)
    texts = []
    for i in range(700):
        n = int(rng.integers(0, 40)) if i % 4 else int(rng.integers(100, 420))
        s = rng.choice(hot, size=n)
        noise = rng.random(n) < 0.12
        s = np.where(noise, rng.integers(1, 256, n).astype(np.uint8), s).astype(np.uint8)
        texts.append(s.tobytes())
    texts += [b"", b"~", b"~F", b"~FR", b"/~FR", b"//~FR", b"~/~", b"\n", b"~OL\n", b"x/~", b"~FBK", b"\xff~RS\x80", b"~FX~RS", b"tail~"]
    text, off = O.pack(texts)
    n = len(texts)
    kind = (np.arange(n) % 2).astype(np.uint8)                       # alternately write_user and write_room
    target = np.where(kind == 0, np.arange(n) % 4, 0).astype(np.int32)
    ops = dict(text=text, off=off, kind=kind, target=target, except_user=np.full(n, -1, np.int32),
               flags=np.where(np.arange(n) % 7 == 0, api.OF_PAGER, 0).astype(np.uint8))
    users = dict(room=np.zeros(4, np.int32), flags=np.array([1, 0, 1, 0], np.uint8), level=np.ones(4, np.uint8))
    ctx = _ctx(sim_lib)
    ctx.set_users(users["room"], users["flags"], users["level"], 1)
    st = ctx.write_batch(ops)
    eoff, data, nd = port.write_batch(ops, users)
    assert (st.off == eoff).all()
    bad = np.nonzero(st.data != data)[0]
    assert bad.size == 0, bad[:8]
    ctx.close()


def test_long_scan_tiles(sim_lib, port):
    """The per-op scans at the edge between the two forms of k_scan1: n + 1 = 16384 elements is the last input of the
    one-item-per-thread form (64 full tiles), n + 1 = 16385 the first of the 16-items form (tiles through shared
    memory, several tiles of look-back, a last tile that holds the total alone)."""
    U = 6
    users = dict(room=np.array([0, 0, 1, 1, 0, -1], np.int32), flags=np.array([1, 0, 1, 0, 5, 1], np.uint8), level=np.ones(U, np.uint8))
    ctx = _ctx(sim_lib)
    ctx.set_users(users["room"], users["flags"], users["level"], 2)
    for n in (16383, 16384):
        rng = np.random.default_rng(n)
        texts = [b"x" * int(k) for k in rng.integers(0, 3, n)]
        texts[7] = b"~FRred\n"; texts[n - 1] = b"last/~ one\n"
        text, off = O.pack(texts)
        kind = np.zeros(n, np.uint8); kind[::1000] = 1
        ops = dict(text=text, off=off, kind=kind, target=np.where(kind == 0, rng.integers(-1, U, n), rng.integers(-1, 2, n)).astype(np.int32),
                   except_user=np.full(n, -1, np.int32), flags=np.zeros(n, np.uint8))
        st = ctx.write_batch(ops)
        eoff, data, nd = port.write_batch(ops, users)
        assert (st.off == eoff).all() and (st.data == data).all() and st.n_deliveries == int(nd.sum()), n
    ctx.close()
