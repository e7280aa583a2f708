"""Debug aid for tests/test_dropin.py: runs both talkers in lockstep and reports the first input line after which
some socket differs.  usage: dropin_debug.py <seed> <users> <rooms> <lines> <iov 0|1> [sim]"""
import ctypes as C, os, sys, tempfile, random
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import test_dropin as T

seed, U, R, N, iov = (int(x) for x in sys.argv[1:6])
if len(sys.argv) > 6:
    from cpusim.build_sim import build_sim
    libp = build_sim()
else:
    from nuts333_b200 import api, build
    build.build(); libp = api.library_path()
C.CDLL(str(libp), mode=C.RTLD_GLOBAL)
libs = [T._bind(C.CDLL(str(T.REFDIR / "libdropin_ref.so"))), T._bind(C.CDLL(str(T.REFDIR / "libdropin_shim.so")))]
sc = T.make_script(seed, U, R, N)
words = T._words64() if len(sys.argv) <= 6 else ["fuck", "shit", "cunt"]

class Stepper:
    """run_session as a generator-like object: the same steps, one input line at a time"""
    def __init__(self, lib, d):
        self.lib, self.d = lib, d
    def streams(self):
        lib = self.lib
        out = {}
        for fd in range(lib.dropin_fd_base(), lib.dropin_next_fd()):
            n = lib.dropin_stream_len(fd)
            out[fd] = bytes(np.ctypeslib.as_array(lib.dropin_stream_ptr(fd), shape=(n,))) if n else b""
        return out

# monkeypatch: intercept dropin_input to stop after each line
import threading
dirs = [T._scratch(tempfile.mkdtemp()) for _ in libs]
# run both sessions fully but record, per line index, the stream lengths: then diff
def run(lib, d):
    lens = []
    orig_input = lib.dropin_input
    def hook(h, ln):
        r = orig_input(h, ln)
        lib.dropin_flush()
        lens.append((h, ln, {fd: lib.dropin_stream_len(fd) for fd in range(lib.dropin_fd_base(), lib.dropin_next_fd())}))
        return r
    lib.dropin_input = hook
    out = T.run_session(lib, d, sc, 0, bool(iov), words)
    return lens, out
la, oa = run(libs[0], dirs[0])
lb, ob = run(libs[1], dirs[1])
print("lines", len(la), len(lb), "fds", len(oa[0]), len(ob[0]))
for k, (a, b) in enumerate(zip(la, lb)):
    if a[2] != b[2] or a[0] != b[0]:
        print("first difference after line", k, a[0], a[1], "handle b", b[0])
        bad = [fd for fd in sorted(set(a[2]) | set(b[2])) if a[2].get(fd) != b[2].get(fd)]
        print("fds", bad[:10])
        for fd in bad[:3]:
            x, y = oa[0].get(fd, b""), ob[0].get(fd, b"")
            p0 = la[k - 1][2].get(fd, 0) if k else 0
            print(fd, "ref:", x[max(0, p0 - 80):p0 + 300]); print(fd, "shim:", y[max(0, p0 - 80):p0 + 300])
        for lib, o, nm in ((libs[0], oa, "ref"), (libs[1], ob, "shim")):
            pass
        sfd_a = [fd for fd in a[2] if a[2][fd] != (la[k - 1][2].get(fd, 0) if k else 0)]
        sfd_b = [fd for fd in b[2] if b[2][fd] != (lb[k - 1][2].get(fd, 0) if k else 0)]
        print("fds that grew on this line: ref", sfd_a[:8], "...", len(sfd_a), " shim", sfd_b[:8], "...", len(sfd_b))
        for fd in sorted(set(sfd_a) ^ set(sfd_b))[:4] + sorted(set(sfd_b))[:2]:
            pa = la[k - 1][2].get(fd, 0) if k else 0; pb = lb[k - 1][2].get(fd, 0) if k else 0
            print(fd, "ref+:", oa[0].get(fd, b"")[pa:a[2].get(fd, 0)][:300]); print(fd, "shim+:", ob[0].get(fd, b"")[pb:b[2].get(fd, 0)][:300])
        for j in range(max(0, k - 6), k + 1): print("  line", j, la[j][0], la[j][1])
        break
else:
    print("no difference; equal streams:", oa[0] == ob[0])
