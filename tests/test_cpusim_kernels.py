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
