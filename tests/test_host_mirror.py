"""Runs the C++ host-mirror test binary (tests/cpp/test_host_mirror.cpp): the reference's operator-surface unit
tests restated against tiered-storage-for-apache-kafka_b200/host/chunk_transform.hpp."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(target, *args):
    subprocess.check_call(["make", "-s", "-C", ROOT, target])
    out = subprocess.run([os.path.join(ROOT, target), *args], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().splitlines()[-1].startswith("OK:")


def test_host_mirror_builders_and_chain_under_emulation():
    # CPU box: index builders / finisher / fetch plan natively, the batched chain through the test-only SIMT build
    _run("tests/cpp/test_host_mirror_simt", "--device")


@pytest.mark.gpu
def test_host_mirror_chain_on_gpu():
    _run("tests/cpp/test_host_mirror_gpu", "--device")
