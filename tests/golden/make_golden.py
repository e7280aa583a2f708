"""Mints tests/golden/*.json from the UNMODIFIED reference (oracle/_ref/libnutsref.so,
built from /root/reference by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

The vectors pin the oracle restatement (tests/test_oracle_golden.py) and the CUDA
path (tests/test_gpu_parity.py) to the reference's own compiled code.
"""
import hashlib
import json
import random
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))
import oracle_lib as O  # noqa: E402

R = O.ref()
assert R is not None, "build oracle/_ref first (make -C oracle)"
hx = lambda b: bytes(b).hex()

KAT_STRINGS = [
    b"Fred says: ~OLhi ~FRred~RS there\n", b"/~FR literal", b"~", b"~~", b"~/~", b"//~FR", b"~F", b"~FX~RS",
    b"a\nb", b"~FBK", b"tail~", b"~OL\n", b"x/~", b"\xff~RS\x80", b"~rs lower", b"~RS", b"",
    b"\n", b"\n\n\n", b"/", b"//", b"/~", b"/~~RS", b"~/", b"/\n~RS", b"~R", b"~RS~", b"~R~RS", b"~~RS", b"~\nRS",
    b"~FK~FR~FG~FY~FB~FM~FT~FW~BK~BR~BG~BY~BB~BM~BT~BW~RS~OL~UL~LI~RV", b"\x07\x07beep\x07",
    b"~" * 40, b"/~" * 20, b"\n" * 33, b"/" * 31 + b"~RS", b"x" * 31 + b"~RS", b"x" * 30 + b"~RS", b"x" * 29 + b"~RS",
    b"x" * 62 + b"~O", b"x" * 63 + b"~OL", b"x" * 64 + b"~OL!",
    # bytes one bit away from the special ones, right after them (SWAR borrow traps)
    b"a/.b", b"~\x7fRS", b"x\n\x0by", b"/.~RS", b"//.", b"~\x7f", b"abc/.~FRdef\n\x0b~\x7f/./~", b"\x7f\x7e\x2e\x2f\x0b\x0a" * 9,
    b"http://a/./b/~FRc/~d", b"...///...~~~...\n\n\x0b",
]
# codes straddling the reference's 1000-byte buffer flush points and the string end
for pos in (990, 993, 994, 995, 996, 997, 998, 999, 1000, 1001, 1010):
    KAT_STRINGS.append(b"y" * pos + b"~FR" + b"z" * 7)
    KAT_STRINGS.append(b"y" * pos + b"\n~OL\n")
KAT_STRINGS.append(b"q" * 1997 + b"~RS")
KAT_STRINGS.append(b"q" * 1998 + b"~R")
KAT_STRINGS.append((b"~FR\n" * 500))

rng = random.Random(0x333)
ALPHA = [b"~", b"/", b"\n", b"F", b"R", b"S", b"O", b"L", b"K", b"B", b"G", b"T", b"W", b"Y", b"M", b"U", b"I", b"V",
         b"x", b"y", b" "]
fuzz = []
for i in range(400):
    n = rng.randint(0, 39) if i % 50 else rng.randint(900, 1990)
    s = bytearray()
    for _ in range(n):
        s += bytes([rng.randint(1, 255)]) if rng.random() < 0.14 else rng.choice(ALPHA)
    fuzz.append(bytes(s[:2000]))

render = [dict(s=hx(s), c0=hx(R.render(s, 0)), c1=hx(R.render(s, 1))) for s in KAT_STRINGS]
render_fuzz = [dict(s=hx(s), c0=hashlib.sha256(R.render(s, 0)).hexdigest(), c1=hashlib.sha256(R.render(s, 1)).hexdigest(),
                    n0=len(R.render(s, 0)), n1=len(R.render(s, 1))) for s in fuzz]

# contains_swearing, stock list and a 64-word list
R.set_swear_words(["fuck", "shit", "cunt"])
SW = [b"what the FUCK", b"sh~RSit", b"~OLShIt", b"Scunthorpe", b"clean", b"fu ck", b"", b"FUCK", b"f", b"shi", b"xshitx" * 3]
swear_stock = [dict(s=hx(s), v=R.contains_swearing(s)) for s in SW]
from nuts333_b200 import synth  # noqa: E402
w64 = synth.swear_words(64)
R.set_swear_words(w64[:-1])
bt, bo = synth.bodies(3000, w64)
v64 = R.contains_swearing_batch(bt, bo)
swear64 = dict(words=w64, n=3000, seed=synth.SEED, verdict_sha256=hashlib.sha256(v64.tobytes()).hexdigest(),
               dirty=int(v64.sum()))
R.set_swear_words(["fuck", "shit", "cunt"])

cc = [dict(s=hx(s), count=R.colour_com_count(s), strip=hx(R.colour_com_strip(s)))
      for s in [b"~FBK", b"~FB", b"~FR~RS", b"/~FRa~OLb~c~", b"~", b"~~RS", b"plain"]]

# ban files
site_file = b"evil.com\n.badnet.org\n10.1.\nlast.noeol"
user_file = b"Troll\nSpammer\nNoeol"
R.set_ban_file(0, site_file)
R.set_ban_file(1, user_file)
SQ = [b"host.evil.com", b"evil.com.au", b"x.badnet.org", b"badnet.org", b"10.1.2.3", b"110.1.2.3", b"last.noeol", b"good.org", b""]
UQ = [b"Troll", b"troll", b"Trol", b"Trolls", b"Spammer", b"Noeol", b""]
bans = dict(site_file=hx(site_file), user_file=hx(user_file),
            site=[dict(q=hx(q), v=R.site_banned(q)) for q in SQ], user=[dict(q=hx(q), v=R.user_banned(q)) for q in UQ])
# whitespace variety + trailing newline + empty + missing
site2 = b"  a.b \t c.d\r\n\n e.f\x0bg.h\x0c i.j \n"
R.set_ban_file(0, site2)
bans["site2_file"] = hx(site2)
bans["site2"] = [dict(q=hx(q), v=R.site_banned(q)) for q in [b"xa.bx", b"c.d", b"e.f", b"g.h", b"i.j", b"e.fg.h", b"zz"]]
R.set_ban_file(0, b"")
bans["empty"] = [dict(q=hx(q), v=R.site_banned(q)) for q in [b"a", b""]]
R.set_ban_file(0, None)
bans["missing"] = [dict(q=hx(q), v=R.site_banned(q)) for q in [b"a", b""]]
# generated 300-entry lists, both newline variants
gb = {}
for tn in (0, 1):
    sf = synth.ban_file(0, 300, 2000, 2000, bool(tn)); uf = synth.ban_file(1, 300, 2000, 2000, bool(tn))
    R.set_ban_file(0, sf); R.set_ban_file(1, uf)
    st, so = synth.sites(2000); nt, no = synth.names(2000)
    vs = R.ban_batch(0, st, so); vu = R.ban_batch(1, nt, no)
    gb[str(tn)] = dict(site_sha256=hashlib.sha256(vs.tobytes()).hexdigest(), site_hits=int(vs.sum()),
                       user_sha256=hashlib.sha256(vu.tobytes()).hexdigest(), user_hits=int(vu.sum()))
bans["generated"] = gb
R.set_ban_file(0, None); R.set_ban_file(1, None)

# config C1: 1,000 x write_room(drive, "Fred says: ~OLline %04d ~FRred~RS done\n"), colour off / on
c1 = {}
for colour in (0, 1):
    users = dict(room=np.zeros(1, np.int32), flags=np.array([colour], np.uint8), level=np.array([4], np.uint8))
    R.reset(1, users)
    for i in range(1000):
        R.lib.ref_write_room_except(0, b"Fred says: ~OLline %04d ~FRred~RS done\n" % i, -1, 0)
    s = R.stream(0)
    c1[str(colour)] = dict(n=len(s), sha256=hashlib.sha256(s).hexdigest())

# a mixed batch: every op kind, filters, excepts, all-room ops, no-room users, gates
rng = random.Random(99)
U, NR, N = 60, 4, 500
room = np.array([rng.randint(-1, NR - 1) for _ in range(U)], np.int32)
flags = np.array([rng.choice([0, 1, 1, 0, 2, 4, 8, 5, 9, 12, 13, 1, 0]) for _ in range(U)], np.uint8)
level = np.array([rng.randint(0, 4) for _ in range(U)], np.uint8)
texts, kind, target, exc, fl, gate = [], [], [], [], [], []
verdict = np.array([rng.randint(0, 1) for _ in range(64)], np.uint8)
for i in range(N):
    k = rng.choice([0, 1, 1, 1, 2])
    n = rng.randint(0, 60)
    texts.append(b"".join(rng.choice(ALPHA + [b"~FR", b"~RS", b"~OL", b"word ", b"/~"]) for _ in range(n))[:1500])
    kind.append(k)
    if k == 0:
        target.append(rng.randint(-1, U - 1)); exc.append(-1); f = 0
    elif k == 1:
        target.append(rng.randint(-1, NR - 1)); exc.append(rng.randint(-1, U - 1)); f = rng.choice([0, 0, 1, 2, 3])
    else:
        target.append(rng.randint(0, 4)); exc.append(rng.randint(-1, U - 1)); f = rng.choice([0, 4])
    g = rng.choice([-1, -1, -1, rng.randint(0, 63)])
    gate.append(g)
    fl.append(f | (rng.choice([0, 8]) if g >= 0 else 0))
text, off = O.pack(texts)
ops = dict(text=text, off=off, kind=np.array(kind, np.uint8), target=np.array(target, np.int32),
           except_user=np.array(exc, np.int32), flags=np.array(fl, np.uint8), gate=np.array(gate, np.int32))
users = dict(room=room, flags=flags, level=level)
R.write_batch(ops, NR, users, verdict=verdict)
streams = R.streams(U)
mixed = dict(n_rooms=NR, room=room.tolist(), flags=flags.tolist(), level=level.tolist(),
             texts=[hx(t) for t in texts], kind=kind, target=target, except_user=exc, op_flags=fl, gate=gate,
             verdict=verdict.tolist(), stream_len=[len(s) for s in streams],
             stream_sha256=[hashlib.sha256(s).hexdigest() for s in streams])

# reference data files rendered whole (inputs stay in /root/reference; only digests are kept)
files = {}
for f in ("helpfiles/colour", "helpfiles/mainhelp", "datafiles/drive.R", "datafiles/wizroom.R", "motd2"):
    p = Path("/root/reference") / f
    s = p.read_bytes()
    files[f] = dict(in_sha256=hashlib.sha256(s).hexdigest(), n0=len(R.render(s, 0)), n1=len(R.render(s, 1)),
                    c0=hashlib.sha256(R.render(s, 0)).hexdigest(), c1=hashlib.sha256(R.render(s, 1)).hexdigest())

out = dict(render=render, render_fuzz=render_fuzz, swear_stock=swear_stock, swear64=swear64, colour_com=cc,
           bans=bans, c1=c1, mixed=mixed, files=files,
           note="minted from the unmodified reference by tests/golden/make_golden.py")
(HERE / "golden.json").write_text(json.dumps(out, indent=0))
print("wrote", HERE / "golden.json", (HERE / "golden.json").stat().st_size, "bytes")
print("C1", c1)
print({k: (v["n0"], v["c0"][:8], v["n1"], v["c1"][:8]) for k, v in files.items()})
