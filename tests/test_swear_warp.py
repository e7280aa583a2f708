"""contains_swearing's warp-cooperative matcher (k_ac_warp, nutsb_match.cuh) against the oracle restatement of
nuts333.c:2540-2559: a warp's 32 strings are one piece of text cut into 32 equal pieces by bytes, so matches that
straddle two lanes' pieces, strings that start or end inside a piece, empty strings, pieces larger than the staging
window and pattern lists of every shape all have to come out as strstr() would have it."""
import numpy as np
import pytest

import oracle_lib as O
from nuts333_b200 import api, synth


def _case(rng, n, maxlen, words, p_word=0.25, alphabet=b"abcdefgh ~RS"):
    ws = [w.encode() for w in words if w != "*"]
    out = []
    for _ in range(n):
        ln = int(rng.integers(0, maxlen + 1))
        s = bytearray(rng.choice(np.frombuffer(alphabet, np.uint8), size=ln).tobytes())
        if ws and ln and rng.random() < p_word:                      # plant a word (random case), or a near miss
            w = bytearray(ws[int(rng.integers(len(ws)))])
            if rng.random() < 0.3 and len(w) > 1: w[int(rng.integers(len(w)))] ^= 1
            if rng.random() < 0.5: w = bytearray(bytes(w).upper())
            at = int(rng.integers(0, ln))
            s[at:at + len(w)] = w
        out.append(bytes(s[:999]))
    return out


LISTS = [
    ["fuck", "shit", "cunt", "*"],
    ["a", "*"],
    ["abcabcabd", "bca", "cab", "dd", "*"],
    ["x" * 40 + "y", "hh", "*"],                                      # a pattern longer than a lane's piece of text
    ["*"],                                                            # empty list: nothing swears
    ["", "zzz", "*"],                                                 # an empty word: strstr(s, "") matches everything
    ["UPPER", "ab", "*"],                                             # upper-case list words never match (the haystack is lower-cased)
]


def _check(ctx, port, seed, sizes):
    rng = np.random.default_rng(seed)
    for words in LISTS + [synth.swear_words(64)]:
        ctx.set_swear_words(words)
        for n, maxlen in sizes:
            strs = _case(rng, n, maxlen, words)
            if n > 40:
                strs[3] = b""; strs[4] = b""; strs[35] = b""          # empty strings inside and at the edge of a warp's piece
            text, off = O.pack(strs)
            got = ctx.contains_swearing_batch(text, off)
            want = port.contains_swearing_batch(text, off, words)
            bad = np.nonzero(got != want)[0]
            assert bad.size == 0, (words[:3], n, maxlen, int(bad[0]), strs[int(bad[0])], int(got[bad[0]]), int(want[bad[0]]))
    ctx.set_swear_words(["fuck", "shit", "cunt", "*"])


def test_warp_matcher_on_emulator(sim_lib, port):
    ctx = api.Context(0, sim_lib)
    # 70 strings of <= 60 bytes: staged pieces; 40 of <= 400: pieces beyond the 3 KB window (one string per lane)
    _check(ctx, port, 3, [(70, 60), (33, 5), (40, 400)])
    ctx.close()


@pytest.mark.gpu
def test_warp_matcher_on_gpu(gpu_ctx, port):
    _check(gpu_ctx, port, 4, [(20000, 90), (5000, 12), (3000, 999), (31, 200), (1, 50)])
