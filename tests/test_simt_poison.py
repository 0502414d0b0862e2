"""The emulated kernels with garbage instead of zeros in shared memory and in freshly allocated device memory
(TSGPU_SIMT_POISON=1): the plain emulator hands out zeroed memory, a GPU does not, so a kernel that forgets to initialise
its scratch passes one and fails the other.  Runs the emulator suites again in a child process with the switch on."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_emulated_kernels_do_not_rely_on_zeroed_memory():
    env = dict(os.environ, TSGPU_SIMT_POISON="1")
    out = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "not gpu", "-p", "no:cacheprovider",
                          "tests/test_simt_zstd.py", "tests/test_simt_aesgcm.py", "tests/test_golden_fixtures.py"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, (out.stdout + out.stderr)[-4000:]
