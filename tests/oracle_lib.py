"""ctypes bindings for the CHECKERS under oracle/ (test infrastructure only).

`port()`  -> oracle/liboracle.so      (plain-C restatement, always built)
`ref()`   -> oracle/_ref/libnutsref.so (the unmodified reference driven in-process;
             present only where /root/reference was available at build time, or
             where the prebuilt binary travelled with the snapshot)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"

u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)
i32p = C.POINTER(C.c_int32)

OP_USER, OP_ROOM, OP_LEVEL = 0, 1, 2
OF_FORCE_LISTEN, OF_SHOUT, OF_ABOVE, OF_GATE_IF_SET = 1, 2, 4, 8
UF_COLOUR, UF_LOGIN, UF_IGNALL, UF_IGNSHOUT, UF_CLONE, UF_REMOTE = 1, 2, 4, 8, 16, 32


def build_oracle(force: bool = False) -> None:
    """make -C oracle (restatement always; _ref only when the reference is present)."""
    so = ORACLE_DIR / "liboracle.so"
    src_new = max((ORACLE_DIR / f).stat().st_mtime for f in ("nuts_oracle.c", "nuts_oracle.h"))
    if force or not so.exists() or so.stat().st_mtime < src_new:
        subprocess.run(["make", "-C", str(ORACLE_DIR), "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    refso = ORACLE_DIR / "_ref" / "libnutsref.so"
    if Path("/root/reference/nuts333.c").exists():
        h = ORACLE_DIR / "ref_harness.c"
        if force or not refso.exists() or refso.stat().st_mtime < h.stat().st_mtime:
            subprocess.run(["make", "-C", str(ORACLE_DIR), "ref"], check=True,
                           stdout=subprocess.DEVNULL)


def _ptr(a, typ):
    if a is None:
        return None
    return a.ctypes.data_as(typ)


def pack(strings):
    """list[bytes] -> (bytes u8 array, off u64 array) CSR."""
    off = np.zeros(len(strings) + 1, dtype=np.uint64)
    if strings:
        off[1:] = np.cumsum([len(s) for s in strings], dtype=np.uint64)
    buf = np.frombuffer(b"".join(strings), dtype=np.uint8).copy() if strings else np.zeros(0, np.uint8)
    if buf.size == 0:
        buf = np.zeros(1, np.uint8)[:0].copy()
    return buf, off


class _Streams(C.Structure):
    _fields_ = [("n_users", C.c_int64), ("off", u64p), ("bytes", u8p), ("n_deliveries", u64p)]


class Port:
    """oracle/liboracle.so"""

    def __init__(self):
        build_oracle()
        self.lib = L = C.CDLL(str(ORACLE_DIR / "liboracle.so"))
        L.orc_render.restype = C.c_size_t
        L.orc_render.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_char_p]
        L.orc_contains_swearing.restype = C.c_int
        L.orc_contains_swearing.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_char_p)]
        L.orc_colour_com_count.restype = C.c_int
        L.orc_colour_com_count.argtypes = [C.c_char_p, C.c_size_t]
        L.orc_colour_com_strip.restype = C.c_size_t
        L.orc_colour_com_strip.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        L.orc_site_banned.restype = C.c_int
        L.orc_site_banned.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_char_p, C.c_size_t]
        L.orc_user_banned.restype = C.c_int
        L.orc_user_banned.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_char_p, C.c_size_t]
        L.orc_ban_tokens.restype = C.c_size_t
        L.orc_ban_tokens.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_size_t]
        L.orc_write_batch.restype = C.c_int
        L.orc_write_batch.argtypes = [C.c_int64, u8p, u64p, u8p, i32p, i32p, u8p, i32p, u8p,
                                      C.c_int32, i32p, u8p, u8p, i32p, C.c_int32, C.POINTER(_Streams)]
        L.orc_write_batch_count.restype = C.c_int64
        L.orc_write_batch_count.argtypes = [C.c_int64, u8p, u64p, u8p, i32p, i32p, u8p, i32p, u8p,
                                            C.c_int32, i32p, u8p, u8p, u64p]
        L.orc_streams_free.argtypes = [C.POINTER(_Streams)]
        L.orc_delivery_digests.argtypes = [C.c_int64, u8p, u64p, u8p, i32p, i32p, u8p, i32p, u8p,
                                           C.c_int32, i32p, u8p, u8p, u64p, u64p]
        L.orc_contains_swearing_batch.argtypes = [C.c_int64, u8p, u64p, C.POINTER(C.c_char_p), u8p]
        L.orc_site_banned_batch.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int64, u8p, u64p, u8p]
        L.orc_user_banned_batch.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int64, u8p, u64p, u8p]
        L.orc_speech_ops.restype = C.c_int64
        L.orc_speech_ops.argtypes = [C.c_int64, u8p, i32p, u8p, u64p, u8p, u64p, u8p, i32p, C.c_int, C.POINTER(C.c_char_p),
                                     u8p, C.c_size_t, u64p, u8p, i32p, i32p, u8p, C.c_int64]
        L.orc_more.restype = C.c_int
        L.orc_more.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_char_p,
                               C.POINTER(C.c_size_t)]
        L.orc_ban_edit.restype = C.c_int
        L.orc_ban_edit.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p,
                                   C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
        L.orc_fnv1a.restype = C.c_uint64
        L.orc_fnv1a.argtypes = [C.c_char_p, C.c_size_t]

    @staticmethod
    def _words(words):
        arr = (C.c_char_p * (len(words) + 1))(*[w if isinstance(w, bytes) else w.encode() for w in words], None)
        return arr

    def render(self, s: bytes, colour: int) -> bytes:
        out = C.create_string_buffer(6 * len(s) + 8)
        n = self.lib.orc_render(s, len(s), colour, out)
        return out.raw[:n]

    def contains_swearing(self, s: bytes, words) -> int:
        """words: list WITH the '*' sentinel semantics applied by the callee (list ends at '*' or end)."""
        return self.lib.orc_contains_swearing(s, len(s), self._words(words))

    def colour_com_count(self, s: bytes) -> int:
        return self.lib.orc_colour_com_count(s, len(s))

    def colour_com_strip(self, s: bytes) -> bytes:
        out = C.create_string_buffer(len(s) + 8)
        n = self.lib.orc_colour_com_strip(s, len(s), out)
        return out.raw[:n]

    def site_banned(self, file, site: bytes) -> int:
        return self.lib.orc_site_banned(file or b"", len(file or b""), int(file is not None), site, len(site))

    def user_banned(self, file, name: bytes) -> int:
        return self.lib.orc_user_banned(file or b"", len(file or b""), int(file is not None), name, len(name))

    def ban_tokens(self, file: bytes):
        n = self.lib.orc_ban_tokens(file, len(file), None, None, 0)
        o = (C.c_uint32 * max(n, 1))()
        l = (C.c_uint32 * max(n, 1))()
        self.lib.orc_ban_tokens(file, len(file), o, l, n)
        return [file[o[i]:o[i] + l[i]] for i in range(n)]

    def write_batch(self, ops, users, verdict=None, only_users=None):
        """ops: dict of numpy arrays (text, off, kind, target, except_user, flags[, gate]);
        users: dict (room, flags, level).  Returns (off u64[U+1], bytes u8[], n_deliveries u64[U])."""
        st = _Streams()
        only = None if only_users is None else np.ascontiguousarray(only_users, dtype=np.int32)
        gate = ops.get("gate")
        rc = self.lib.orc_write_batch(
            len(ops["kind"]), _ptr(ops["text"], u8p), _ptr(ops["off"], u64p), _ptr(ops["kind"], u8p),
            _ptr(ops["target"], i32p), _ptr(ops["except_user"], i32p), _ptr(ops["flags"], u8p),
            _ptr(gate, i32p), _ptr(verdict, u8p),
            len(users["room"]), _ptr(users["room"], i32p), _ptr(users["flags"], u8p), _ptr(users["level"], u8p),
            _ptr(only, i32p), 0 if only is None else len(only), C.byref(st))
        if rc != 0:
            raise MemoryError("orc_write_batch")
        U = len(users["room"])
        off = np.ctypeslib.as_array(st.off, shape=(U + 1,)).copy()
        total = int(off[U])
        data = np.ctypeslib.as_array(st.bytes, shape=(max(total, 1),))[:total].copy()
        nd = np.ctypeslib.as_array(st.n_deliveries, shape=(U + 1,))[:U].copy()
        self.lib.orc_streams_free(C.byref(st))
        return off, data, nd

    def delivery_digests(self, ops, users, verdict=None):
        """SURVEY.md 8(d) parity digests -> (per_user u64[U], per_op u64[n_ops])"""
        U, n = len(users["room"]), len(ops["kind"])
        pu, po = np.zeros(max(U, 1), np.uint64), np.zeros(max(n, 1), np.uint64)
        rc = self.lib.orc_delivery_digests(n, _ptr(ops["text"], u8p), _ptr(ops["off"], u64p), _ptr(ops["kind"], u8p), _ptr(ops["target"], i32p),
                                           _ptr(ops["except_user"], i32p), _ptr(ops["flags"], u8p), _ptr(ops.get("gate"), i32p), _ptr(verdict, u8p),
                                           U, _ptr(users["room"], i32p), _ptr(users["flags"], u8p), _ptr(users["level"], u8p), _ptr(pu, u64p), _ptr(po, u64p))
        assert rc == 0, rc
        return pu[:U], po[:n]

    def write_batch_count(self, ops, users, verdict=None):
        nb = C.c_uint64(0)
        gate = ops.get("gate")
        d = self.lib.orc_write_batch_count(
            len(ops["kind"]), _ptr(ops["text"], u8p), _ptr(ops["off"], u64p), _ptr(ops["kind"], u8p),
            _ptr(ops["target"], i32p), _ptr(ops["except_user"], i32p), _ptr(ops["flags"], u8p),
            _ptr(gate, i32p), _ptr(verdict, u8p),
            len(users["room"]), _ptr(users["room"], i32p), _ptr(users["flags"], u8p), _ptr(users["level"], u8p),
            C.byref(nb))
        return int(d), int(nb.value)

    def contains_swearing_batch(self, text, off, words):
        n = len(off) - 1
        v = np.zeros(max(n, 1), np.uint8)
        self.lib.orc_contains_swearing_batch(n, _ptr(text, u8p), _ptr(off, u64p), self._words(words), _ptr(v, u8p))
        return v[:n]

    def ban_batch(self, which, file, text, off):
        n = len(off) - 1
        v = np.zeros(max(n, 1), np.uint8)
        fn = self.lib.orc_user_banned_batch if which else self.lib.orc_site_banned_batch
        fn(file or b"", len(file or b""), int(file is not None), n, _ptr(text, u8p), _ptr(off, u64p), _ptr(v, u8p))
        return v[:n]

    def ban_edit(self, file, is_user: bool, add: bool, token: bytes):
        """ban / unban restated on file bytes (None = no file) -> (result, file afterwards or None)"""
        data = file or b""
        out = C.create_string_buffer(len(data) + len(token) + 8)
        on, op = C.c_size_t(0), C.c_int(0)
        r = self.lib.orc_ban_edit(data, len(data), int(file is not None), int(is_user), int(add), token, out, C.byref(on), C.byref(op))
        return r, (out.raw[:on.value] if op.value else None)

    def speech_ops(self, verb, speaker, bodies, body_off, names, name_off, sflags, room, ban_swearing, words, target=None):
        """input lines -> ops dict (the callers restated); target[m] = the user tell / pemote line m is for"""
        n = len(verb)
        tgt = None if target is None else np.ascontiguousarray(target, np.int32)
        self.lib.orc_set_speech_targets(None if tgt is None else tgt.ctypes.data_as(C.c_void_p))
        n_rev = int(((np.asarray(verb) == 6) | (np.asarray(verb) == 10)).sum())   # a review replays up to 15 lines + header + footer
        cap = 3 * n + 17 * n_rev + 1
        tcap = int(body_off[-1]) * 2 + 256 * n + 17 * 256 * n_rev + 64
        text = np.zeros(tcap, np.uint8); off = np.zeros(cap + 1, np.uint64); kind = np.zeros(cap, np.uint8)
        target = np.zeros(cap, np.int32); exc = np.zeros(cap, np.int32); flags = np.zeros(cap, np.uint8)
        q = self.lib.orc_speech_ops(n, _ptr(np.ascontiguousarray(verb, np.uint8), u8p), _ptr(np.ascontiguousarray(speaker, np.int32), i32p),
                                    _ptr(bodies, u8p), _ptr(body_off, u64p), _ptr(names, u8p), _ptr(name_off, u64p),
                                    _ptr(np.ascontiguousarray(sflags, np.uint8), u8p), _ptr(np.ascontiguousarray(room, np.int32), i32p),
                                    int(ban_swearing), self._words(words), _ptr(text, u8p), tcap, _ptr(off, u64p), _ptr(kind, u8p),
                                    _ptr(target, i32p), _ptr(exc, i32p), _ptr(flags, u8p), cap)
        self.lib.orc_set_speech_targets(None)
        assert q >= 0
        return dict(text=text[:int(off[q])].copy(), off=off[:q + 1].copy(), kind=kind[:q].copy(), target=target[:q].copy(),
                    except_user=exc[:q].copy(), flags=flags[:q].copy())

    def more(self, data, user_null: bool, colour: int, filepos: int):
        """-> (retval, bytes written to the socket, new filepos)"""
        n = len(data or b"")
        out = C.create_string_buffer(6 * n + 512)
        pos, ol = C.c_int64(filepos), C.c_size_t(0)
        rv = self.lib.orc_more(data or b"", n, int(data is not None), int(user_null), colour, C.byref(pos), out, C.byref(ol))
        return rv, out.raw[:ol.value], pos.value

    def fnv1a(self, b: bytes) -> int:
        return self.lib.orc_fnv1a(b, len(b))


class Ref:
    """oracle/_ref/libnutsref.so -- the unmodified reference, in-process.
    One population at a time (the reference keeps its users in globals)."""

    def __init__(self):
        build_oracle()
        so = ORACLE_DIR / "_ref" / "libnutsref.so"
        if not so.exists():
            raise FileNotFoundError(str(so))
        self.lib = L = C.CDLL(str(so))
        L.ref_add_users.argtypes = [C.c_int, i32p, u8p, u8p]
        L.ref_write_user.argtypes = [C.c_int, C.c_char_p]
        L.ref_write_room_except.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_uint8]
        L.ref_write_level.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.ref_contains_swearing.argtypes = [C.c_char_p]
        L.ref_colour_com_count.argtypes = [C.c_char_p]
        L.ref_colour_com_strip.restype = C.c_size_t
        L.ref_colour_com_strip.argtypes = [C.c_char_p, C.c_char_p]
        L.ref_set_swear_words.argtypes = [C.POINTER(C.c_char_p)]
        L.ref_set_ban_file.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
        L.ref_site_banned.argtypes = [C.c_char_p]
        L.ref_user_banned.argtypes = [C.c_char_p]
        L.ref_stream_len.restype = C.c_size_t
        L.ref_stream_ptr.restype = C.POINTER(C.c_uint8)
        L.ref_stream_calls.restype = C.c_uint64
        L.ref_total_write_calls.restype = C.c_uint64
        L.ref_total_write_bytes.restype = C.c_uint64
        L.ref_set_user_speech.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_int]
        L.ref_speech.argtypes = [C.c_int, C.c_int, C.c_char_p]
        L.ref_speech_to.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p]
        L.ref_more.argtypes = [C.c_int, C.c_int, C.c_char_p]
        L.ref_get_filepos.restype = C.c_long
        L.ref_ban_command.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p]
        L.ref_get_ban_file.restype = C.c_long
        L.ref_get_ban_file.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
        L.ref_write_batch.restype = C.c_int64
        L.ref_write_batch.argtypes = [C.c_int64, u8p, u64p, u8p, i32p, i32p, u8p, i32p, u8p]
        L.ref_contains_swearing_batch.argtypes = [C.c_int64, u8p, u64p, u8p]
        L.ref_ban_batch.argtypes = [C.c_int, C.c_int64, u8p, u64p, u8p]
        self._tmp = tempfile.mkdtemp(prefix="nutsref_cwd_")
        self._cwd = os.getcwd()
        L.ref_reset()

    def reset(self, n_rooms=0, users=None, sink_mode=0):
        L = self.lib
        L.ref_reset()
        L.ref_set_sink_mode(sink_mode)
        if n_rooms:
            L.ref_add_rooms(n_rooms)
        if users is not None:
            L.ref_add_users(len(users["room"]), _ptr(users["room"], i32p), _ptr(users["flags"], u8p),
                            _ptr(users["level"], u8p))

    def render(self, s: bytes, colour: int) -> bytes:
        """write_user() to a single USER_TYPE recipient; s must not contain NUL."""
        assert b"\0" not in s
        self.reset(1, dict(room=np.zeros(1, np.int32), flags=np.array([1 if colour else 0], np.uint8),
                           level=np.ones(1, np.uint8)))
        self.lib.ref_write_user(0, s)
        return self.stream(0)

    def stream(self, u: int) -> bytes:
        n = self.lib.ref_stream_len(u)
        if n == 0:
            return b""
        return bytes(np.ctypeslib.as_array(self.lib.ref_stream_ptr(u), shape=(n,)))

    def set_swear_words(self, words):
        arr = (C.c_char_p * (len(words) + 1))(*[w if isinstance(w, bytes) else w.encode() for w in words], None)
        return self.lib.ref_set_swear_words(arr)

    def contains_swearing(self, s: bytes) -> int:
        assert b"\0" not in s
        return self.lib.ref_contains_swearing(s)

    def colour_com_count(self, s: bytes) -> int:
        return self.lib.ref_colour_com_count(s)

    def colour_com_strip(self, s: bytes) -> bytes:
        out = C.create_string_buffer(len(s) + 8)
        n = self.lib.ref_colour_com_strip(s, out)
        return out.raw[:n]

    def set_ban_file(self, which: int, data):
        cwd = os.getcwd()
        try:
            rc = self.lib.ref_set_ban_file(self._tmp.encode(), which, data, 0 if data is None else len(data))
        finally:
            os.chdir(cwd)
        assert rc == 0
        return rc

    def _in_tmp(self, fn, *a):
        # the reference opens datafiles/<list> relative to the CWD (c:336,355)
        cwd = os.getcwd()
        os.chdir(self._tmp)
        try:
            return fn(*a)
        finally:
            os.chdir(cwd)

    def ban_command(self, by: int, which: int, add: bool, token: bytes):
        """the reference's own ban_site / ban_user / unban_site / unban_user -> the list on disk afterwards"""
        cwd = os.getcwd()
        try:
            rc = self.lib.ref_ban_command(self._tmp.encode(), by, which, int(add), token)
        finally:
            os.chdir(cwd)
        assert rc == 0
        buf = C.create_string_buffer(1 << 20)
        n = self.lib.ref_get_ban_file(self._tmp.encode(), which, buf, len(buf))
        return None if n < 0 else buf.raw[:n]

    def site_banned(self, site: bytes) -> int:
        return self._in_tmp(self.lib.ref_site_banned, site)

    def user_banned(self, name: bytes) -> int:
        return self._in_tmp(self.lib.ref_user_banned, name)

    def more(self, u: int, null_user: bool, filename: str):
        """the reference's pager on a file on disk -> (retval, user->filepos)"""
        rv = self.lib.ref_more(u, int(null_user), filename.encode())
        return rv, int(self.lib.ref_get_filepos(u))

    def write_batch(self, ops, n_rooms, users, verdict=None, sink_mode=0):
        self.reset(n_rooms, users, sink_mode)
        gate = ops.get("gate")
        calls = self.lib.ref_write_batch(
            len(ops["kind"]), _ptr(ops["text"], u8p), _ptr(ops["off"], u64p), _ptr(ops["kind"], u8p),
            _ptr(ops["target"], i32p), _ptr(ops["except_user"], i32p), _ptr(ops["flags"], u8p),
            _ptr(gate, i32p), _ptr(verdict, u8p))
        return calls

    def streams(self, n_users):
        return [self.stream(u) for u in range(n_users)]

    def delivery_digests(self, ops, n_rooms, users, verdict=None):
        """the same digests folded from the reference's own write(2) calls"""
        self.reset(n_rooms, users, 1)
        U, n = len(users["room"]), len(ops["kind"])
        pu, po = np.zeros(max(U, 1), np.uint64), np.zeros(max(n, 1), np.uint64)
        self.lib.ref_write_batch_digests.restype = C.c_int64
        rc = self.lib.ref_write_batch_digests(n, _ptr(ops["text"], u8p), _ptr(ops["off"], u64p), _ptr(ops["kind"], u8p), _ptr(ops["target"], i32p),
                                              _ptr(ops["except_user"], i32p), _ptr(ops["flags"], u8p), _ptr(ops.get("gate"), i32p), _ptr(verdict, u8p),
                                              _ptr(pu, u64p), _ptr(po, u64p))
        assert rc >= 0
        return pu[:U], po[:n]

    def contains_swearing_batch(self, text, off):
        n = len(off) - 1
        v = np.zeros(max(n, 1), np.uint8)
        self.lib.ref_contains_swearing_batch(n, _ptr(text, u8p), _ptr(off, u64p), _ptr(v, u8p))
        return v[:n]

    def ban_batch(self, which, text, off):
        n = len(off) - 1
        v = np.zeros(max(n, 1), np.uint8)
        self._in_tmp(self.lib.ref_ban_batch, which, n, _ptr(text, u8p), _ptr(off, u64p), _ptr(v, u8p))
        return v[:n]


_PORT = None
_REF = None


def port() -> Port:
    global _PORT
    if _PORT is None:
        _PORT = Port()
    return _PORT


def ref():
    """Returns the Ref singleton or None when oracle/_ref was never built."""
    global _REF
    if _REF is None:
        try:
            _REF = Ref()
        except (FileNotFoundError, OSError):
            return None
    return _REF
