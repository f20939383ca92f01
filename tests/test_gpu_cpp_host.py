"""Builds and runs the C++ test of the GPUCompute host module (tests/cpp/test_gpucompute.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_gpucompute.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_gpucompute.bin")


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = os.path.join(ROOT, "urho3d_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", SRC, "-o", EXE, "-L" + lib, "-lu3dgpu", "-L" + os.path.join(ROOT, "oracle"),
                           "-loracle", "-Wl,-rpath," + lib, "-Wl,-rpath," + os.path.join(ROOT, "oracle")])


def test_cpp_host_module_compiles_and_links():
    """CPU-only check: the header-only module compiles against include/u3d_gpu.h and links to libu3dgpu.so."""
    build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_host_module_parity():
    build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    print(out.stdout[-2000:], out.stderr[-2000:])
    assert out.returncode == 0 and out.stdout.strip().endswith("OK")
