"""The pager row: more() = the write_user byte machine applied to a
file one fgets() chunk at a time, no closing reset, 23 screen lines per page.

  * the oracle restatement (orc_more) against the reference's own more() in-process,
  * the C-ABI mirror (nutsb_q_more + NUTSB_OF_PAGER/PLAIN ops) against the oracle:
    on the SIMT emulator here, on the GPU with `-m gpu`.
"""
import os
import tempfile

import numpy as np
import pytest

import oracle_lib as O
from nuts333_b200 import api
from pager_util import CASES, make_file


def _pages_port(port, data, user_null, colour):
    """every page until more() reports the end -> list of (retval, bytes, filepos)"""
    out, pos = [], 0
    for _ in range(60):
        rv, b, pos = port.more(data, user_null, colour, pos)
        out.append((rv, b, pos))
        if rv != 1:
            break
    return out


def test_port_vs_reference_pager(port, ref):
    with tempfile.TemporaryDirectory() as d:
        for ci, case in enumerate(CASES):
            data = make_file(**case)
            path = os.path.join(d, f"f{ci}")
            with open(path, "wb") as fh:
                fh.write(data)
            for colour in (0, 1):
                for user_null in (False, True):
                    users = dict(room=np.zeros(1, np.int32), flags=np.array([colour], np.uint8), level=np.ones(1, np.uint8))
                    ref.reset(1, users)
                    for rv, b, pos in _pages_port(port, data, user_null, colour):
                        ref.lib.ref_stream_clear(0)
                        rrv, rpos = ref.more(0, user_null, path)
                        assert (rrv, ref.stream(0)) == (rv, b), (ci, colour, user_null)
                        if not user_null:
                            assert rpos == pos
        # missing file
        users = dict(room=np.zeros(1, np.int32), flags=np.array([1], np.uint8), level=np.ones(1, np.uint8))
        ref.reset(1, users)
        assert ref.more(0, False, os.path.join(d, "nope"))[0] == 0 == port.more(None, False, 1, 7)[0]


def _check_mirror(ctx, port):
    ctx.set_users(np.zeros(3, np.int32), np.array([0, 1, 1], np.uint8), np.ones(3, np.uint8), 1)
    with tempfile.TemporaryDirectory() as d:
        for ci, case in enumerate(CASES):
            data = make_file(**case)
            path = os.path.join(d, f"f{ci}")
            with open(path, "wb") as fh:
                fh.write(data)
            t = api.Talker(ctx)
            # users 0 (colour off) and 1 (colour on) page through the file; user 2 gets the login-stage call
            exp = {u: _pages_port(port, data, False, u) for u in (0, 1)}
            for page in range(max(len(exp[0]), len(exp[1]))):
                want = {}
                for u in (0, 1):
                    if page < len(exp[u]):
                        rv = t.more(u, u, path)
                        assert rv == exp[u][page][0] and t.filepos[u] == exp[u][page][2]
                        want[u] = exp[u][page][1]
                t.write_user(0, "after the page ~FRok\n")          # ordinary ops interleave with pager lines
                st = t.flush()
                assert st.user(0) == want.get(0, b"") + b"after the page ok\n\r"
                assert st.user(1) == want.get(1, b"")
            assert t.more(None, 2, path) == 2
            st = t.flush()
            assert st.user(2) == port.more(data, True, 1, 0)[1] and st.user(0) == b""
        t = api.Talker(ctx)
        t.filepos[1] = 99
        assert t.more(1, 1, os.path.join(d, "nope")) == 0 and t.filepos[1] == 0 and t.pending() == 0


def test_pager_mirror_on_emulator(sim_lib, port):
    ctx = api.Context(0, sim_lib)
    _check_mirror(ctx, port)
    ctx.close()


@pytest.mark.gpu
def test_pager_mirror_on_gpu(gpu_ctx, port):
    _check_mirror(gpu_ctx, port)


@pytest.mark.gpu
def test_pager_flags_in_batches(gpu_ctx, port):
    """NUTSB_OF_PAGER / NUTSB_OF_PLAIN on every op kind, mixed with ordinary ops."""
    import random
    rng = random.Random(5)
    U, NR, N = 50, 2, 600
    room = np.array([rng.randint(0, NR - 1) for _ in range(U)], np.int32)
    flags = np.array([rng.choice([0, 1]) for _ in range(U)], np.uint8)
    level = np.ones(U, np.uint8)
    toks = [b"~OL", b"~FR", b"~RS", b"/~", b"\n", b"abc ", b"~", b"/"]
    texts = [b"".join(rng.choice(toks) for _ in range(rng.randint(0, 20))) for _ in range(N)]
    kind = np.array([rng.choice([0, 1, 1]) for _ in range(N)], np.uint8)
    target = np.array([rng.randint(0, U - 1) if k == 0 else rng.randint(-1, NR - 1) for k in kind], np.int32)
    exc = np.array([-1 if k == 0 else rng.randint(-1, U - 1) for k in kind], np.int32)
    of = np.array([rng.choice([0, api.OF_PAGER, api.OF_PLAIN, api.OF_PAGER | api.OF_PLAIN]) for _ in range(N)], np.uint8)
    text, off = O.pack(texts)
    ops = dict(text=text, off=off, kind=kind, target=target, except_user=exc, flags=of)
    users = dict(room=room, flags=flags, level=level)
    gpu_ctx.set_users(room, flags, level, NR)
    st = gpu_ctx.write_batch(ops)
    o, d, nd = port.write_batch(ops, users)
    assert (st.off == o).all() and (st.data == d).all() and st.n_deliveries == int(nd.sum())
