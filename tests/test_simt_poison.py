"""The emulated kernels with garbage instead of zeros in shared memory and in freshly allocated device memory
(TSGPU_SIMT_POISON=1): the plain emulator hands out zeroed memory, a GPU does not, so a kernel that forgets to initialise
its scratch passes one and fails the other.  Runs the emulator suites again in a child process with the switch on."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


import pytest


@pytest.mark.parametrize("order", ["", "reverse"])
def test_emulated_kernels_do_not_rely_on_zeroed_memory_or_lane_order(order):
    # order "reverse": lanes run from 31 down to 0 between synchronisation points — with the default ascending order lane 0
    # always runs first, which would hide a missing __syncwarp() after a "lane 0 writes, every lane reads" section.
    # (TSGPU_SIMT_ORDER=random:<seed> draws the next lane at random; slower, for manual runs.)
    env = dict(os.environ, TSGPU_SIMT_POISON="1", TSGPU_SIMT_ORDER=order)
    out = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "not gpu", "-p", "no:cacheprovider",
                          "tests/test_simt_zstd.py", "tests/test_simt_aesgcm.py", "tests/test_simt_fuzz.py", "tests/test_golden_fixtures.py"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, (out.stdout + out.stderr)[-4000:]
