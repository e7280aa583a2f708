"""Loads tests/golden/golden.json (vectors minted from the unmodified reference)."""
import json
from pathlib import Path

import numpy as np

import oracle_lib as O

_G = None


def golden():
    global _G
    if _G is None:
        _G = json.loads((Path(__file__).resolve().parent / "golden" / "golden.json").read_text())
    return _G


def mixed_batch():
    """-> ops dict, users dict, n_rooms, verdict, expected lens, expected sha256s"""
    m = golden()["mixed"]
    texts = [bytes.fromhex(t) for t in m["texts"]]
    text, off = O.pack(texts)
    ops = dict(text=text, off=off, kind=np.array(m["kind"], np.uint8), target=np.array(m["target"], np.int32),
               except_user=np.array(m["except_user"], np.int32), flags=np.array(m["op_flags"], np.uint8),
               gate=np.array(m["gate"], np.int32))
    users = dict(room=np.array(m["room"], np.int32), flags=np.array(m["flags"], np.uint8),
                 level=np.array(m["level"], np.uint8))
    return ops, users, m["n_rooms"], np.array(m["verdict"], np.uint8), m["stream_len"], m["stream_sha256"]


def kat_render_batch():
    """Every golden render string as a write_user op to a colour-off user (0) and a
    colour-on user (1).  -> ops, users, expected streams [c0 concat, c1 concat]"""
    g = golden()["render"]
    texts, kind, target = [], [], []
    exp = [b"", b""]
    for v in g:
        s = bytes.fromhex(v["s"])
        for u, key in ((0, "c0"), (1, "c1")):
            texts.append(s); kind.append(0); target.append(u)
            exp[u] += bytes.fromhex(v[key])
    text, off = O.pack(texts)
    n = len(texts)
    ops = dict(text=text, off=off, kind=np.array(kind, np.uint8), target=np.array(target, np.int32),
               except_user=np.full(n, -1, np.int32), flags=np.zeros(n, np.uint8))
    users = dict(room=np.zeros(2, np.int32), flags=np.array([0, 1], np.uint8), level=np.ones(2, np.uint8))
    return ops, users, exp
