"""The in-engine side of GPUCompute.h (its URHO3D_API branch: Camera*, Matrix3x4, BoundingBox, Frustum, CullMode overloads and helpers)
compiled against the reference's own headers and linked with the reference's own objects (oracle/_ref) and libu3dgpu.so:
tests/cpp/test_engine_types.cpp instantiates the same call-site code with the reference's OcclusionBuffer and with GpuOcclusionBuffer.
The program is built where /root/reference exists; the binary travels to the GPU box (tests/cpp/*.bin is git-ignored, not gpurun-ignored)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_engine_types.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_engine_types.bin")
REF = os.environ.get("REF", "/root/reference")
REFLIB = os.path.join(ROOT, "oracle", "_ref")


def build():
    """Same include roots and float mode as oracle/build_ref.sh; -std=c++11 because that is what the engine is compiled with."""
    if not os.path.isdir(os.path.join(REF, "Source", "Urho3D")):
        return os.path.exists(EXE)
    subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")])
    if os.path.exists(EXE) and os.path.getmtime(EXE) >= max(os.path.getmtime(p) for p in (
            SRC, os.path.join(ROOT, "urho3d_b200", "GPUCompute", "GPUCompute.h"),
            os.path.join(ROOT, "urho3d_b200", "GPUCompute", "GpuOcclusionBufferObject.h"),
            os.path.join(ROOT, "urho3d_b200", "GPUCompute", "GpuOctreeMirror.h"), os.path.join(ROOT, "include", "u3d_gpu.h"),
            os.path.join(REFLIB, "libu3dref.so"))):
        return True
    src, tp, stub = os.path.join(REF, "Source", "Urho3D"), os.path.join(REF, "Source", "ThirdParty"), os.path.join(ROOT, "oracle", "ref_stub")
    lib = os.path.join(ROOT, "urho3d_b200", "lib")
    subprocess.check_call(
        ["g++", "-std=c++11", "-O1", "-msse3", "-ffp-contract=off", "-w", "-DURHO3D_STATIC_DEFINE", "-DURHO3D_THREADING", "-DURHO3D_LOGGING",
         "-DURHO3D_OPENGL", "-DURHO3D_SSE", "-I" + stub, "-I" + os.path.join(stub, "Urho3D"), "-I" + os.path.join(REFLIB, "inc"), "-I" + src,
         "-I" + os.path.join(REF, "Source"), "-I" + tp, SRC, "-o", EXE, "-L" + REFLIB, "-lu3dref", "-L" + lib, "-lu3dgpu",
         "-Wl,-rpath," + REFLIB, "-Wl,-rpath," + lib, "-lpthread", "-ldl"])
    return True


def run(expect_device):
    env = dict(os.environ)
    if expect_device:
        env["U3D_EXPECT_DEVICE"] = "1"
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120, env=env)
    print(out.stdout[-2000:], out.stderr[-2000:])
    return out


def test_engine_overloads_compile_link_and_run_against_the_reference_headers():
    if not build():
        pytest.skip("the reference tree is not present and no prebuilt binary travelled")
    out = run(expect_device=False)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK")


@pytest.mark.gpu
def test_gpu_class_matches_the_reference_class_in_one_process():
    if not build():
        pytest.skip("no prebuilt binary travelled")
    out = run(expect_device=True)
    assert out.returncode == 0 and "in-process" in out.stdout and out.stdout.strip().endswith("OK")
