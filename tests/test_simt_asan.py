"""The emulated kernels under AddressSanitizer.  In the SIMT build shared memory and device scratch are heap blocks, so an
out-of-bounds shared-memory index that a GPU would silently absorb aborts here (this caught a 1 KiB overrun of the aliased
Huffman scratch when the hash table shrank to 2 KiB).  Complements compute-sanitizer on the real device."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_emulated_kernels_are_clean_under_asan():
    libasan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(libasan) or not os.path.exists(libasan):
        pytest.skip("libasan not available")
    subprocess.check_call(["make", "-s", "-C", ROOT, "tests/simt/libtsgpu_simt_asan.so"])
    env = dict(os.environ, LD_PRELOAD=libasan, ASAN_OPTIONS="detect_leaks=0")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "simt", "asan_sweep.py")], env=env,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "asan sweep ok" in out.stdout, (out.stdout + out.stderr)[-4000:]
