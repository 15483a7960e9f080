"""The C++ host side (include/minarrow_b200.hpp) above the C ABI: the reference's leaf-kernel tests transcribed into C++
(tests/cpp/test_reference_kats.cpp) with the reference's function names.  CPU: the header compiles, the binary links
against every entry point it uses and the library loads.  GPU: the transcribed tests run through the CUDA path."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_reference_kats.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_reference_kats")
ROUTES_SRC = os.path.join(ROOT, "tests", "cpp", "test_container_routes.cpp")
ROUTES_EXE = os.path.join(ROOT, "tests", "cpp", "test_container_routes")
GROUP_SRC = os.path.join(ROOT, "tests", "cpp", "test_shard_group.cpp")
GROUP_EXE = os.path.join(ROOT, "tests", "cpp", "test_shard_group")


def _build(src, exe):
    lib = os.path.join(ROOT, "minarrow_b200", "libminarrow_b200.so")
    deps = [src, lib] + [os.path.join(ROOT, "include", h) for h in ("minarrow_b200.hpp", "minarrow_b200_containers.hpp", "minarrow_b200.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        # $ORIGIN-relative rpath: the binary finds the library wherever the repo is copied (the GPU box uses another path)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), src,
                               "-L" + os.path.join(ROOT, "minarrow_b200"), "-lminarrow_b200",
                               "-Wl,-rpath,$ORIGIN/../../minarrow_b200", "-o", exe])
    return exe


def build_cpp():
    _build(ROUTES_SRC, ROUTES_EXE)
    _build(GROUP_SRC, GROUP_EXE)
    return _build(SRC, EXE)


def test_cpp_host_layer_compiles_links_and_loads():
    exe = build_cpp()
    r = subprocess.run([exe, "--link"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.startswith("abi 1"), r.stdout + r.stderr


def test_cpp_container_routes_compile_link_and_load():
    build_cpp()
    r = subprocess.run([ROUTES_EXE, "--link"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.startswith("abi 1"), r.stdout + r.stderr


def test_cpp_shard_group_compiles_links_and_loads():
    build_cpp()
    r = subprocess.run([GROUP_EXE, "--link"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.startswith("abi 1"), r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("world", [0, 1, 2, 5])
def test_shard_group_single_process_multi_rank(gpu_ctx, world):
    """mnr_shard_* / mnr_group_* from one process (tests/cpp/test_shard_group.cpp): SuperTable stats through the batched
    fused exchange + shard-local element-wise fan-out.  world 0 = one rank per GPU of the box (3 virtual ranks on a
    1-GPU box); explicit worlds map rank r to device r % n_devices, so the mailbox protocol runs on any box."""
    build_cpp()
    r = subprocess.run([GROUP_EXE] + ([str(world)] if world else []), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " 0 failed" in r.stdout and "kernel launches" in r.stdout, r.stdout[-500:]


@pytest.mark.gpu
def test_reference_router_and_container_tests_transcribed_to_cpp_pass_on_the_gpu(gpu_ctx):
    """resolve_binary_arithmetic, route_super_array_broadcast, Table / SuperTable routes: the reference's own tests
    (broadcast/super_array.rs:480-763, table.rs:432-566, super_table.rs:684-887) against include/minarrow_b200_containers.hpp."""
    build_cpp()
    r = subprocess.run([ROUTES_EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " 0 failed" in r.stdout and "kernel launches" in r.stdout, r.stdout[-500:]


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


def test_float_remainder_fast_path_is_exact():
    """minarrow_b200/csrc/fastmod.h (one division + one fma instead of the libm loop for float `%`) against libm's
    fmod / fmodf bit for bit: specials x specials, every exponent distance, exact multiples and their neighbours,
    subnormals.  Host-only, ~4 s; the device build of the same header is checked in tests/test_gpu_narrow_division.py."""
    src = os.path.join(ROOT, "tests", "cpp", "test_fastmod.cpp")
    exe = os.path.join(ROOT, "tests", "cpp", "test_fastmod")
    deps = [src, os.path.join(ROOT, "minarrow_b200", "csrc", "fastmod.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-ffp-contract=off", src, "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    assert r.stdout.count(", 0 failed") == 2, r.stdout


def test_float64_power_fast_path_stays_inside_the_reference_tolerance():
    """minarrow_b200/csrc/fastpow.h (f64 Power = exp(b ln a) in ~45 FP64 operations instead of two libm calls) against
    glibc's exp(b * log(a)) — the oracle's arithmetic — over 1.4e7 random pairs up to |b ln a| = 699, plus the special
    values that must fall through to libm.  Host-only, ~4 s."""
    src = os.path.join(ROOT, "tests", "cpp", "test_fastpow.cpp")
    exe = os.path.join(ROOT, "tests", "cpp", "test_fastpow")
    deps = [src, os.path.join(ROOT, "minarrow_b200", "csrc", "fastpow.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-ffp-contract=off", src, "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.rstrip().endswith("0 failed"), r.stdout[-2000:]
