"""The C++ host side (include/minarrow_b200.hpp) above the C ABI: the reference's leaf-kernel tests transcribed into C++
(tests/cpp/test_reference_kats.cpp) with the reference's function names.  CPU: the header compiles, the binary links
against every entry point it uses and the library loads.  GPU: the transcribed tests run through the CUDA path."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_reference_kats.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_reference_kats")


def build_cpp():
    lib = os.path.join(ROOT, "minarrow_b200", "libminarrow_b200.so")
    deps = [SRC, os.path.join(ROOT, "include", "minarrow_b200.hpp"), os.path.join(ROOT, "include", "minarrow_b200.h"), lib]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), SRC,
                               "-L" + os.path.join(ROOT, "minarrow_b200"), "-lminarrow_b200",
                               "-Wl,-rpath," + os.path.join(ROOT, "minarrow_b200"), "-o", EXE])
    return EXE


def test_cpp_host_layer_compiles_links_and_loads():
    exe = build_cpp()
    r = subprocess.run([exe, "--link"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.startswith("abi 1"), r.stdout + r.stderr


@pytest.mark.gpu
def test_reference_tests_transcribed_to_cpp_pass_on_the_gpu(gpu_ctx):
    exe = build_cpp()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " 0 failed" in r.stdout and "kernel launches" in r.stdout, r.stdout[-500:]


def test_division_by_invariant_scalar_arithmetic():
    """minarrow_b200/csrc/divmagic.h (the multiply-high replacement for `column / scalar`) against the CPU's own divide:
    edge divisors x edge dividends, 20 M random pairs per width, the 16-bit domain exhaustively.  Host-only, ~3 s."""
    src = os.path.join(ROOT, "tests", "cpp", "test_divmagic.cpp")
    exe = os.path.join(ROOT, "tests", "cpp", "test_divmagic")
    deps = [src, os.path.join(ROOT, "minarrow_b200", "csrc", "divmagic.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", src, "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-2000:]
    assert r.stdout.count(" 0 failed") == 5, r.stdout
