"""Pins the oracle restatement (oracle/nuts_oracle.c) to vectors minted from the
unmodified reference (tests/golden/golden.json), and -- where oracle/_ref was built --
the reference harness to the same vectors."""
import hashlib

import numpy as np
import pytest

import oracle_lib as O
from golden_util import golden, mixed_batch
from nuts333_b200 import synth


def _impls(port, want_ref=True):
    r = O.ref() if want_ref else None
    return [("port", port)] + ([("ref", r)] if r is not None else [])


def test_render_kat(port):
    for name, imp in _impls(port):
        for v in golden()["render"]:
            s = bytes.fromhex(v["s"])
            assert imp.render(s, 0).hex() == v["c0"], (name, s[:40])
            assert imp.render(s, 1).hex() == v["c1"], (name, s[:40])


def test_render_fuzz(port):
    for v in golden()["render_fuzz"]:
        s = bytes.fromhex(v["s"])
        a, b = port.render(s, 0), port.render(s, 1)
        assert (len(a), len(b)) == (v["n0"], v["n1"])
        assert hashlib.sha256(a).hexdigest() == v["c0"] and hashlib.sha256(b).hexdigest() == v["c1"]


def test_render_bounds(port):
    # <= 6n+4 with colour, <= 2n without
    for s in (b"\n" * 100, b"~FR" * 50, b"", b"/~" * 30):
        assert len(port.render(s, 1)) <= 6 * len(s) + 4 and len(port.render(s, 0)) <= 2 * len(s)


def test_swear(port):
    stock = ["fuck", "shit", "cunt", "*"]
    for v in golden()["swear_stock"]:
        assert port.contains_swearing(bytes.fromhex(v["s"]), stock) == v["v"]
    g = golden()["swear64"]
    assert synth.swear_words(64) == g["words"]
    bt, bo = synth.bodies(g["n"], g["words"], seed=g["seed"])
    v = port.contains_swearing_batch(bt, bo, g["words"])
    assert int(v.sum()) == g["dirty"] and hashlib.sha256(v.tobytes()).hexdigest() == g["verdict_sha256"]
    # list semantics: ends at '*', empty word matches everything, upper-case words never match
    assert port.contains_swearing(b"abc", ["*", "abc"]) == 0
    assert port.contains_swearing(b"abc", ["", "*"]) == 1
    assert port.contains_swearing(b"ABC abc", ["ABC", "*"]) == 0


def test_colour_com(port):
    for v in golden()["colour_com"]:
        s = bytes.fromhex(v["s"])
        assert port.colour_com_count(s) == v["count"] and port.colour_com_strip(s).hex() == v["strip"]


def test_bans(port):
    b = golden()["bans"]
    sf, uf = bytes.fromhex(b["site_file"]), bytes.fromhex(b["user_file"])
    for v in b["site"]:
        assert port.site_banned(sf, bytes.fromhex(v["q"])) == v["v"]
    for v in b["user"]:
        assert port.user_banned(uf, bytes.fromhex(v["q"])) == v["v"]
    s2 = bytes.fromhex(b["site2_file"])
    for v in b["site2"]:
        assert port.site_banned(s2, bytes.fromhex(v["q"])) == v["v"]
    for v in b["empty"]:
        assert port.site_banned(b"", bytes.fromhex(v["q"])) == v["v"]
    for v in b["missing"]:
        assert port.site_banned(None, bytes.fromhex(v["q"])) == v["v"]
    assert port.ban_tokens(sf) == [b"evil.com", b".badnet.org", b"10.1."]     # last token untested (feof)
    for tn in (0, 1):
        g = b["generated"][str(tn)]
        sfile, ufile = synth.ban_file(0, 300, 2000, 2000, bool(tn)), synth.ban_file(1, 300, 2000, 2000, bool(tn))
        st, so = synth.sites(2000)
        nt, no = synth.names(2000)
        vs, vu = port.ban_batch(0, sfile, st, so), port.ban_batch(1, ufile, nt, no)
        assert (int(vs.sum()), hashlib.sha256(vs.tobytes()).hexdigest()) == (g["site_hits"], g["site_sha256"])
        assert (int(vu.sum()), hashlib.sha256(vu.tobytes()).hexdigest()) == (g["user_hits"], g["user_sha256"])


def test_c1_config(port):
    """BASELINE config 1: 1,000 write_room lines to Fred, colour off / on."""
    texts = [b"Fred says: ~OLline %04d ~FRred~RS done\n" % i for i in range(1000)]
    text, off = O.pack(texts)
    n = len(texts)
    ops = dict(text=text, off=off, kind=np.full(n, O.OP_ROOM, np.uint8), target=np.zeros(n, np.int32),
               except_user=np.full(n, -1, np.int32), flags=np.zeros(n, np.uint8))
    for colour in (0, 1):
        users = dict(room=np.zeros(1, np.int32), flags=np.array([colour], np.uint8), level=np.array([4], np.uint8))
        so, data, nd = port.write_batch(ops, users)
        g = golden()["c1"][str(colour)]
        assert len(data) == g["n"] and hashlib.sha256(data.tobytes()).hexdigest() == g["sha256"] and nd[0] == 1000


def test_mixed_batch(port):
    ops, users, n_rooms, verdict, lens, shas = mixed_batch()
    off, data, nd = port.write_batch(ops, users, verdict=verdict)
    for u in range(len(lens)):
        s = data[int(off[u]):int(off[u + 1])].tobytes()
        assert len(s) == lens[u] and hashlib.sha256(s).hexdigest() == shas[u], u
    # only_users restriction gives the same streams for the chosen users
    off2, data2, _ = port.write_batch(ops, users, verdict=verdict, only_users=[3, 17])
    for u in (3, 17):
        assert data2[int(off2[u]):int(off2[u + 1])].tobytes() == data[int(off[u]):int(off[u + 1])].tobytes()


def test_reference_files(port):
    """Whole reference data files as one string (inputs are read from /root/reference
    when it exists; only digests are committed)."""
    from pathlib import Path
    base = Path("/root/reference")
    if not base.exists():
        pytest.skip("/root/reference not present")
    for f, g in golden()["files"].items():
        s = (base / f).read_bytes()
        assert hashlib.sha256(s).hexdigest() == g["in_sha256"]
        a, b = port.render(s, 0), port.render(s, 1)
        assert (len(a), hashlib.sha256(a).hexdigest()) == (g["n0"], g["c0"])
        assert (len(b), hashlib.sha256(b).hexdigest()) == (g["n1"], g["c1"])
