"""The compiled host mirror (include/tuun_b200.hpp) over the C ABI: the reference's known-answer
generator tests restated in C++ (tests/cpp/golden_test.cpp), host half here, device half on the GPU."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "cpp", "golden_test")


def _build():
    subprocess.run(["make", "-C", os.path.join(HERE, "cpp")], check=True, stdout=subprocess.PIPE,
                   stderr=subprocess.STDOUT)


def test_cpp_host_flatten_and_lowering():
    _build()
    r = subprocess.run([BIN, "--host-only"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "33 cases, 0 failures" in r.stdout


@pytest.mark.gpu
def test_cpp_golden_vectors_on_device():
    if not os.path.exists(BIN):
        _build()
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr[-4000:]
    assert "0 failures" in r.stdout
