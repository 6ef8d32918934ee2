"""The C++ host mirror (include/prestige.hpp) and the C ABI, exercised from a compiled C++ program."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_test")


def _build():
    src = os.path.join(ROOT, "tests", "cpp", "host_test.cpp")
    lib = os.path.join(ROOT, "prestige_b200")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(ROOT, "include", "prestige.hpp"))):
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), src, "-L" + lib, "-lprestige_b200",
                        "-Wl,-rpath," + lib, "-o", EXE], check=True, capture_output=True)


def test_cpp_mirror_and_abi_cpu():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "for i in 0..n {" in r.stdout and r.stdout.strip().endswith("OK")


@pytest.mark.gpu
def test_cpp_eq1_through_abi_gpu():
    _build()
    r = subprocess.run([EXE, "--gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bit-exact" in r.stdout
