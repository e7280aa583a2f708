"""contains_swearing's warp-cooperative matcher (k_ac_pair, nutsb_match.cuh) against the oracle restatement of
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


def _cut_case(rng, n, words):
    """One long text full of list words, cut into n strings at random places: words broken across two (or, with empty
    and one-byte strings, several) strings must not be found, words whole at a string's very start or end must be."""
    ws = [w.encode() for w in words if w not in ("*", "")] or [b"q"]
    parts, have = [], 0
    fill = [b"a", b"b ", b"xh", b" ", b"h", b"ab x", b"x"]
    pick = rng.integers(0, 1 << 30, 12 * n + 8)
    while have < 12 * n:
        r = int(pick[len(parts)])
        parts.append(ws[(r >> 8) % len(ws)] if (r & 255) < 154 else fill[(r >> 8) % len(fill)])
        have += len(parts[-1])
    text = b"".join(parts)
    cuts = np.sort(rng.integers(0, len(text) + 1, n - 1))
    cuts = [0] + cuts.tolist() + [len(text)]
    return [text[cuts[i]:cuts[i + 1]] for i in range(n)]


def _check(ctx, port, seed, sizes):
    rng = np.random.default_rng(seed)
    for words in LISTS + [synth.swear_words(64)]:
        ctx.set_swear_words(words)
        for n, maxlen in sizes:
            strs = _case(rng, n, maxlen, words) if maxlen >= 0 else _cut_case(rng, n, words)
            if n > 40 and maxlen >= 0:
                strs[3] = b""; strs[4] = b""; strs[35] = b""          # empty strings inside and at the edge of a warp's piece
            text, off = O.pack(strs)
            got = ctx.contains_swearing_batch(text, off)
            want = port.contains_swearing_batch(text, off, words)
            bad = np.nonzero(got != want)[0]
            assert bad.size == 0, (words[:3], n, maxlen, int(bad[0]), strs[int(bad[0])], int(got[bad[0]]), int(want[bad[0]]))
    ctx.set_swear_words(["fuck", "shit", "cunt", "*"])


def test_warp_matcher_on_emulator(sim_lib, port):
    # 70 strings of <= 60 bytes: staged pieces; 40 of <= 400: pieces beyond the 3 KB window (one string per lane);
    # (n, -1): one text cut into n strings at random places.  The short lists walk the squared table (two bytes a
    # step), the 64-word list the one-byte table.
    ctx = api.Context(0, sim_lib)
    _check(ctx, port, 3, [(70, 60), (33, 5), (40, 400), (90, -1), (300, -1)])
    ctx.close()


@pytest.mark.gpu
def test_warp_matcher_on_gpu(gpu_ctx, port):
    _check(gpu_ctx, port, 4, [(20000, 90), (5000, 12), (3000, 999), (31, 200), (1, 50), (20000, -1), (777, -1)])
