"""colour_com_count / colour_com_strip (nuts333.c:2563-2610): the count's
double-count quirk and the strip, against golden vectors from the reference and the oracle."""
import random

import numpy as np
import pytest

import oracle_lib as O
from golden_util import golden
from nuts333_b200 import api

CODES = [b"RS", b"OL", b"UL", b"LI", b"RV", b"FK", b"FR", b"FG", b"FY", b"FB", b"FM", b"FT", b"FW",
         b"BK", b"BR", b"BG", b"BY", b"BB", b"BM", b"BT", b"BW"]


def _strings(seed, n):
    rng = random.Random(seed)
    toks = [b"~", b"/", b"x", b" ", b"\n", b"~F", b"K", b"B", b"R", b"S", b"W"] + [b"~" + c for c in CODES] + CODES
    out = [bytes.fromhex(v["s"]) for v in golden()["colour_com"]]
    out += [b"".join(rng.choice(toks) for _ in range(rng.randint(0, 25))) for _ in range(n)]
    out += [b"~FBKRS", b"~BBKRVIOL", b"~FB", b"~F", b"~", b"", b"~FBW~FBW", b"x" * 1990 + b"~FBK~RS"]
    return out


def _check(ctx, port):
    strings = _strings(3, 600)
    text, off = O.pack(strings)
    cnt = ctx.colour_com_count_batch(text, off)
    d, o = ctx.colour_com_strip_batch(text, off)
    for i, s in enumerate(strings):
        assert int(cnt[i]) == port.colour_com_count(s), s
        assert d[int(o[i]):int(o[i + 1])].tobytes() == port.colour_com_strip(s), s
    for v in golden()["colour_com"]:                 # the reference's own answers
        i = strings.index(bytes.fromhex(v["s"]))
        assert int(cnt[i]) == v["count"] and d[int(o[i]):int(o[i + 1])].tobytes().hex() == v["strip"]
    e = np.zeros(1, np.uint64)
    assert len(ctx.colour_com_count_batch(np.zeros(0, np.uint8), e)) == 0


def test_oracle_vs_reference(port, ref):
    for s in _strings(4, 400):
        if b"\0" in s or len(s) > 900:        # colour_com_strip writes into a static text2[1000] (c:2592)
            continue
        assert port.colour_com_count(s) == ref.colour_com_count(s), s
        assert port.colour_com_strip(s) == ref.colour_com_strip(s), s


def test_on_emulator(sim_lib, port):
    ctx = api.Context(0, sim_lib)
    _check(ctx, port)
    ctx.close()


@pytest.mark.gpu
def test_on_gpu(gpu_ctx, port):
    _check(gpu_ctx, port)
