"""The emulated kernels under AddressSanitizer and UndefinedBehaviorSanitizer.  In the SIMT build shared memory and device scratch are heap blocks, so an
out-of-bounds shared-memory index that a GPU would silently absorb aborts here (this caught a 1 KiB overrun of the aliased
Huffman scratch when the hash table shrank to 2 KiB).  Complements compute-sanitizer on the real device."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("san,lib", [("asan", "libtsgpu_simt_asan.so"), ("ubsan", "libtsgpu_simt_ubsan.so")])
def test_emulated_kernels_are_clean_under_sanitizers(san, lib):
    # ubsan: a shift by >= the width is masked by x86 and clamped by a GPU, a misaligned 128-bit access is tolerated by x86
    # and faults on a GPU — the emulator alone would not notice either
    runtime = subprocess.run(["gcc", "-print-file-name=lib%s.so" % san], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(runtime) or not os.path.exists(runtime):
        pytest.skip("lib%s not available" % san)
    subprocess.check_call(["make", "-s", "-C", ROOT, "tests/simt/" + lib])
    env = dict(os.environ, LD_PRELOAD=runtime, ASAN_OPTIONS="detect_leaks=0", UBSAN_OPTIONS="print_stacktrace=1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "simt", "asan_sweep.py"), lib], env=env,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "sanitizer sweep ok" in out.stdout, (out.stdout + out.stderr)[-4000:]


def test_mutated_frames_never_overrun_under_asan():
    # bit flips, stomps, truncations, appended bytes on frames of both compressors and of libzstd (levels 1 / 3 / 19): any answer
    # but an out-of-bounds access; bytes after the frame are refused like zstd-jni refuses them ("Src size is incorrect")
    runtime = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(runtime) or not os.path.exists(runtime):
        pytest.skip("libasan not available")
    subprocess.check_call(["make", "-s", "-C", ROOT, "tests/simt/libtsgpu_simt_asan.so"])
    env = dict(os.environ, LD_PRELOAD=runtime, ASAN_OPTIONS="detect_leaks=0")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "simt", "asan_corrupt.py"), "libtsgpu_simt_asan.so", "8", "11"],
                         env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "corrupt-frame sweep ok" in out.stdout, (out.stdout + out.stderr)[-4000:]
