"""Differential fuzz: the oracle restatement against the unmodified reference driven
in-process (oracle/_ref).  Skipped where the reference binary was never built."""
import random

import numpy as np

import oracle_lib as O
from nuts333_b200 import synth

ALPHA = [b"~", b"/", b"\n", b"F", b"R", b"S", b"O", b"L", b"K", b"B", b"G", b"T", b"W", b"Y", b"M", b"U", b"I", b"V",
         b"x", b"y", b" "]


def test_render_fuzz(port, ref):
    rng = random.Random(2024)
    for i in range(3000):
        n = rng.randint(0, 39) if i % 50 else rng.randint(900, 1990)
        s = bytearray()
        for _ in range(n):
            s += bytes([rng.randint(1, 255)]) if rng.random() < 0.14 else rng.choice(ALPHA)
        s = bytes(s[:2000])
        for c in (0, 1):
            assert port.render(s, c) == ref.render(s, c), s


def test_write_batch_fuzz(port, ref):
    rng = random.Random(7)
    for case in range(6):
        U, NR, N = rng.randint(1, 80), rng.randint(1, 5), rng.randint(1, 300)
        room = np.array([rng.randint(-1, NR - 1) for _ in range(U)], np.int32)
        flags = np.array([rng.choice([0, 1, 1, 0, 2, 4, 8, 5, 9, 12, 13]) for _ in range(U)], np.uint8)
        level = np.array([rng.randint(0, 4) for _ in range(U)], np.uint8)
        texts, kind, target, exc, fl = [], [], [], [], []
        for i in range(N):
            k = rng.choice([0, 1, 1, 1, 2])
            texts.append(b"".join(rng.choice(ALPHA + [b"~FR", b"~RS", b"word "]) for _ in range(rng.randint(0, 40))))
            kind.append(k)
            if k == 0:
                target.append(rng.randint(-1, U - 1)); exc.append(-1); fl.append(0)
            elif k == 1:
                target.append(rng.randint(-1, NR - 1)); exc.append(rng.randint(-1, U - 1)); fl.append(rng.choice([0, 1, 2, 3]))
            else:
                target.append(rng.randint(0, 4)); exc.append(rng.randint(-1, U - 1)); fl.append(rng.choice([0, 4]))
        text, off = O.pack(texts)
        ops = dict(text=text, off=off, kind=np.array(kind, np.uint8), target=np.array(target, np.int32),
                   except_user=np.array(exc, np.int32), flags=np.array(fl, np.uint8))
        users = dict(room=room, flags=flags, level=level)
        o, data, nd = port.write_batch(ops, users)
        ref.write_batch(ops, NR, users)
        for u in range(U):
            assert data[int(o[u]):int(o[u + 1])].tobytes() == ref.stream(u), (case, u)


def test_swear_and_bans(port, ref):
    words = synth.swear_words(64)
    ref.set_swear_words(words[:-1])
    bt, bo = synth.bodies(5000, words, seed=11)
    assert (port.contains_swearing_batch(bt, bo, words) == ref.contains_swearing_batch(bt, bo)).all()
    ref.set_swear_words(["fuck", "shit", "cunt"])
    for tn in (False, True):
        sf, uf = synth.ban_file(0, 500, 3000, 3000, tn, seed=5), synth.ban_file(1, 500, 3000, 3000, tn, seed=5)
        ref.set_ban_file(0, sf); ref.set_ban_file(1, uf)
        st, so = synth.sites(3000, seed=5)
        nt, no = synth.names(3000)
        assert (port.ban_batch(0, sf, st, so) == ref.ban_batch(0, st, so)).all()
        assert (port.ban_batch(1, uf, nt, no) == ref.ban_batch(1, nt, no)).all()
    ref.set_ban_file(0, None); ref.set_ban_file(1, None)
