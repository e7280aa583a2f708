"""The kernels under AddressSanitizer (no compute-sanitizer on this pool): the same .cu sources compiled for the
SIMT emulator with -fsanitize=address, every device buffer a host malloc of its exact requested size class, shared
memory a static / aligned_alloc array.  Catches out-of-bounds reads and writes of the staging windows, the slab
buffer, the run list and the streams on batches that cross tiles, render windows and seams."""
import os
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_emulated_kernels_under_asan():
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    libasan = subprocess.run([gcc, "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not libasan or not Path(libasan).exists():
        pytest.skip("no libasan")
    env = dict(os.environ, LD_PRELOAD=libasan, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1")
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "tools" / "asan_run.py")], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "asan run ok" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
