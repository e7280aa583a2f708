"""Ban-list maintenance: ban_site c:6216, ban_user c:6262, unban_site c:6341,
unban_user c:6385 as edits of the two token files.

  * the oracle restatement (bytes -> bytes) against the reference's OWN commands run in a scratch
    directory, the resulting files compared byte for byte -- including the feof() quirks: a last token
    without a newline can be banned twice, is glued to the next ban, and is dropped by an unban;
  * the library (nutsb_ban_edit on the context's lists, matchers rebuilt) against the oracle, and the
    verdicts after every edit against the oracle's matchers: emulator here, GPU with `-m gpu`.
"""
import random

import numpy as np
import pytest

import oracle_lib as O
from nuts333_b200 import api

SITES = [b"evil.com", b".badnet.org", b"10.1.", b"last.noeol", b"x.y", b"host.evil.com", b"EVIL.com"]
NAMES = [b"Troll", b"troll", b"Spammer", b"Noeol", b"Bob", b"al"]


def script(seed, n):
    rng = random.Random(seed)
    return [(rng.randint(0, 1), rng.random() < 0.55, 0) for _ in range(n)], rng


def start_files(rng):
    files = [None, b"", b"evil.com\n.badnet.org\n10.1.\nlast.noeol", b"evil.com\n\n  x.y \n", b"Troll\nSpammer\nNoeol", b"Bob\n"]
    return rng.choice(files[:4]), rng.choice([files[0], files[1], files[4], files[5]])


def test_ban_edit_oracle_vs_reference(port, ref):
    users = dict(room=np.zeros(1, np.int32), flags=np.zeros(1, np.uint8), level=np.array([4], np.uint8))
    for seed in range(6):
        steps, rng = script(seed, 40)
        cur = list(start_files(rng))
        ref.reset(1, users)
        ref.lib.ref_set_user_speech(0, b"Wizard", 1, 0)
        ref.set_ban_file(0, cur[0]); ref.set_ban_file(1, cur[1])
        for which, add, _ in steps:
            tok = rng.choice(NAMES if which else SITES)
            r, nf = port.ban_edit(cur[which], bool(which), add, tok)
            disk = ref.ban_command(0, which, add, tok)
            assert nf == disk, (seed, which, add, tok, cur[which], nf, disk)
            cur[which] = nf
    ref.set_ban_file(0, None); ref.set_ban_file(1, None)


def test_ban_edit_quirks(port):
    f = b"evil.com\nlast.noeol"
    assert port.ban_edit(f, False, True, b"last.noeol") == (0, b"evil.com\nlast.noeollast.noeol\n")   # never tested, glued
    assert port.ban_edit(f, False, True, b"evil.com") == (1, f)
    assert port.ban_edit(f, False, False, b"evil.com") == (0, None)        # the untested token is dropped, list emptied
    assert port.ban_edit(f, False, False, b"last.noeol") == (1, f)
    assert port.ban_edit(None, False, False, b"x") == (1, None)
    assert port.ban_edit(None, True, True, b"troll") == (0, b"Troll\n")
    assert port.ban_edit(b"Troll\n", True, True, b"troll") == (1, b"Troll\n")


def _check_library(ctx, port, seeds):
    st, so = O.pack([b"host.evil.com", b"a.badnet.org", b"10.1.2.3", b"good.org", b"last.noeol", b"x.y.z", b"EVIL.com.au"])
    nt, no = O.pack(NAMES + [b"Al"])
    for seed in seeds:
        steps, rng = script(seed, 30)
        cur = list(start_files(rng))
        ctx.set_ban_files(cur[0], cur[1])
        for which, add, _ in steps:
            tok = rng.choice(NAMES if which else SITES)
            r, nf = port.ban_edit(cur[which], bool(which), add, tok)
            assert ctx.ban_edit(which, add, tok) == r
            assert ctx.ban_file(which) == nf, (seed, which, add, tok, cur[which])
            cur[which] = nf
            assert (ctx.site_banned_batch(st, so) == port.ban_batch(0, cur[0], st, so)).all()
            assert (ctx.user_banned_batch(nt, no) == port.ban_batch(1, cur[1], nt, no)).all()
    with pytest.raises(api.NutsbError):
        ctx.ban_edit(0, True, b"two words")
    with pytest.raises(api.NutsbError):
        ctx.ban_edit(1, True, b"Averyveryverylongname")


def test_ban_edit_on_emulator(sim_lib, port):
    ctx = api.Context(0, sim_lib)
    _check_library(ctx, port, (20, 21))
    ctx.close()


@pytest.mark.gpu
def test_ban_edit_on_gpu(gpu_ctx, port):
    _check_library(gpu_ctx, port, (22, 23, 24))
