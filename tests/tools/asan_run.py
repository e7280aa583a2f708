"""Run by tests/test_asan.py in a python started under AddressSanitizer: the product's kernels compiled against the
SIMT emulator with -fsanitize=address, on batches that cross every window (tiles, render windows, seams)."""
import sys, ctypes
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
import numpy as np
from cpusim.build_sim import build_sim
from nuts333_b200 import api, synth
import oracle_lib as O
lib = api.bind(ctypes.CDLL(str(build_sim(sanitize="address"))))
port = O.port()
ctx = api.Context(0, lib)
words = synth.swear_words(8)
ctx.set_swear_words(words)
for n_users, per_room, n_msgs, stress in ((24, 12, 200, False), (9, 3, 120, True)):
    us, n_rooms = synth.users(n_users, per_room, stress=stress)
    bt, bo = synth.bodies(n_msgs, words)
    v = ctx.contains_swearing_batch(bt, bo)
    sops, _, _ = synth.say_ops(n_msgs, n_users, per_room, bt, bo, gated=True)
    ctx.set_users(us["room"], us["flags"], us["level"], n_rooms)
    for ov in ((True, False) if not stress else (True,)):
        ctx.set_overlap(ov)
        st = ctx.write_batch(dict(sops, verdict=v))
        off, data, nd = port.write_batch(sops, us, verdict=v)
        assert (st.off == off).all() and (st.data == data).all()
    # the gather-list kernels (k_direct_compact, k_iov) / the fallback behind filters
    iv = ctx.write_batch_iov(dict(sops, verdict=v))
    assert (iv.off == off).all() and all(iv.user(u) == data[int(off[u]):int(off[u + 1])].tobytes() for u in range(n_users))
# long strings
rng = np.random.default_rng(7)
alphabet = np.frombuffer(b"~/\n" + b"FRSOLKBGTWYMUIV" + b"xy z", np.uint8)
texts = [rng.choice(alphabet, size=int(rng.integers(0, 2001))).tobytes() for _ in range(60)]
text, toff = O.pack(texts)
n = len(texts)
kind = rng.integers(0, 2, n).astype(np.uint8)
ops = dict(text=text, off=toff, kind=kind, target=np.where(kind == 0, rng.integers(0, 7, n), rng.integers(-1, 2, n)).astype(np.int32),
           except_user=rng.integers(-1, 7, n).astype(np.int32), flags=np.zeros(n, np.uint8))
users = dict(room=np.array([0, 0, 0, 1, 1, 0, 1], np.int32), flags=np.array([1, 0, 1, 0, 1, 5, 0], np.uint8), level=np.ones(7, np.uint8))
ctx.set_users(users["room"], users["flags"], users["level"], 2)
st = ctx.write_batch(ops)
eoff, data, nd = port.write_batch(ops, users)
assert (st.off == eoff).all() and (st.data == data).all()
# round 2 kernels: the warp matcher across its staging windows (pieces beyond the window, empty strings, a one-string
# call), the digest kernels, the ban matchers, two room shards behind nutsb_multi
from test_swear_warp import _case as swear_case, LISTS
from test_multi import make_case
for wl in LISTS[:4]:
    ctx.set_swear_words(wl)
    for n, maxlen in ((70, 60), (40, 400), (1, 50), (33, 0), (3, 999)):
        strs = swear_case(rng, n, maxlen, wl)
        if n > 40: strs[3] = b""; strs[35] = b""
        t2, o2 = O.pack(strs)
        assert (ctx.contains_swearing_batch(t2, o2) == port.contains_swearing_batch(t2, o2, wl)).all()
c = make_case(91, 24, 3, 150)
o, us = c["ops"], c["users"]
pu, po = port.delivery_digests(o, us, verdict=o["verdict"])
ctx.set_users(us["room"], us["flags"], us["level"], c["n_rooms"])
ctx.write_batch(o)
du, do = ctx.delivery_digests(len(o["kind"]))
assert (du == pu).all() and (do == po).all()
off, data, nd = port.write_batch(o, us, verdict=o["verdict"])
def fnv(b):
    h = 0xcbf29ce484222325
    for x in b: h = ((h * 0x100000001b3) + x) & 0xFFFFFFFFFFFFFFFF
    return h
sd = np.array([fnv(data[int(off[u]):int(off[u + 1])].tobytes()) for u in range(24)], np.uint64)
assert (ctx.stream_digests() == sd).all()
sf, uf = synth.ban_file(0, 50, 300, 300, False), synth.ban_file(1, 50, 300, 300, True)
ctx.set_ban_files(sf, uf)
stx, sox = synth.sites(300)
ntx, nox = synth.names(300)
assert (ctx.site_banned_batch(stx, sox) == port.ban_batch(0, sf, stx, sox)).all()
assert (ctx.user_banned_batch(ntx, nox) == port.ban_batch(1, uf, ntx, nox)).all()
ctx.close()
m = api.MultiContext([0, 0], lib)
m.set_users(us["room"], us["flags"], us["level"], c["n_rooms"])
got, total, deliv = m.write_batch(o)
assert total == int(off[-1]) and all(got[u] == data[int(off[u]):int(off[u + 1])].tobytes() for u in range(24))
m.write_batch(o, keep=True)
assert (m.stream_digests() == sd).all()
m.close()
print("asan run ok")
