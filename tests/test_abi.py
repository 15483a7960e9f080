"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/*.h declares
(and nothing the binding does not know), the status codes follow KernelError, and — with no GPU — the product
path fails loudly instead of falling back to anything on the CPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "minarrow_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mnr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from minarrow_b200 import _lib
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 60
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_lib.SIGNATURES), (set(names) ^ set(_lib.SIGNATURES))
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("mnr_")}
    assert exported == set(names), exported ^ set(names)


def test_rust_ffi_is_generated_from_the_header():
    """rust/minarrow-b200/src/ffi.rs (the `extern "C"` block of the binding crate; unbuilt here, no cargo/rustc) is
    generated from the header: it is up to date, declares every entry point, and agrees with the ctypes table on arity."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_rust_ffi
    from minarrow_b200 import _lib
    text = gen_rust_ffi.generate()
    assert open(gen_rust_ffi.OUT).read() == text, "stale: run python tools/gen_rust_ffi.py"
    decls = dict(re.findall(r"pub fn (mnr_[a-z0-9_]+)\((.*?)\)(?: -> [^;]+)?;", text))
    assert set(decls) == set(header_functions())
    for name, (res, args) in _lib.SIGNATURES.items():
        n_rust = 0 if not decls[name].strip() else decls[name].count(":")
        assert n_rust == len(args), (name, decls[name], len(args))
        assert ("->" in re.search(rf"pub fn {name}\(.*?\)([^;]*);", text).group(1)) == (res is not None), name


def test_abi_is_plain_c():
    src = open(HEADER).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    assert "torch" not in code.lower() and "cudaStream_t" not in code and "#include <cuda" not in code
    assert 'extern "C"' in src
    # compiles as C (not C++) with nothing but the standard headers
    subprocess.check_call(["gcc", "-std=c99", "-fsyntax-only", "-Wall", "-Werror", "-x", "c", HEADER])


def test_status_codes_follow_kernel_error_order():
    # src/enums/error.rs:157-187 declaration order
    variants = ["TYPE_MISMATCH", "LENGTH_MISMATCH", "BROADCASTING", "OPERATOR_MISMATCH", "UNSUPPORTED_TYPE",
                "COLUMN_NOT_FOUND", "INVALID_ARGUMENTS", "PLAN", "OUT_OF_BOUNDS", "DIVIDE_BY_ZERO"]
    src = open(HEADER).read()
    for i, v in enumerate(variants, start=1):
        assert re.search(rf"MNR_ERR_{v}\s*=\s*-{i}\b", src), v
    ops = re.search(r"typedef enum \{\s*MNR_ADD = 0, MNR_SUB = 1, MNR_MUL = 2, MNR_DIV = 3, MNR_REM = 4, MNR_POW = 5, "
                    r"MNR_FLOORDIV = 6", src)
    assert ops, "mnr_op must follow ArithmeticOperator order (src/enums/operators.rs:19-48)"


def test_host_helpers_without_gpu():
    from minarrow_b200 import _lib
    lib = _lib.load()
    assert lib.mnr_abi_version() == 1
    assert C.sizeof(_lib.Agg) == 32
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    assert lib.mnr_device_count() == 0
    h = C.c_void_p()
    rc = lib.mnr_ctx_create(0, C.byref(h))
    assert rc == -101 and b"no CPU fallback" in lib.mnr_last_error()


def test_agg_combine_and_mean_on_host():
    """mnr_agg_combine / mnr_agg_mean are host arithmetic (the rank-order float add + wrapping int add)."""
    from minarrow_b200 import _lib
    lib = _lib.load()
    A = _lib.Agg
    parts = (A * 3)()
    for i, (s, mn, mx, c) in enumerate([(2 ** 63 - 1, -5, 7, 3), (1, -9, 2, 1), (10, 4, 4, 0)]):
        parts[i].sum.i64, parts[i].min.i64, parts[i].max.i64, parts[i].count = s, mn, mx, c
    out = A()
    assert lib.mnr_agg_combine(2, parts, 3, C.byref(out)) == 0     # MNR_I64
    assert (out.sum.i64, out.min.i64, out.max.i64, out.count) == (-2 ** 63 + 10, -9, 7, 4)
    assert lib.mnr_agg_mean(2, C.byref(out)) == float(-2 ** 63 + 10) / 4
    nan = float("nan")
    for i, (s, mn, mx, c) in enumerate([(0.1, nan, nan, 2), (0.2, -0.0, 3.0, 2), (0.3, 0.0, -1.0, 1)]):
        parts[i].sum.f64, parts[i].min.f64, parts[i].max.f64, parts[i].count = s, mn, mx, c
    assert lib.mnr_agg_combine(5, parts, 3, C.byref(out)) == 0     # MNR_F64
    assert out.sum.f64 == (0.1 + 0.2) + 0.3 and out.count == 5
    import math
    assert out.min.f64 == 0.0 and math.copysign(1, out.min.f64) == -1.0 and out.max.f64 == 3.0
    empty = A()
    assert math.isnan(lib.mnr_agg_mean(5, C.byref(empty)))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "minarrow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                code = "\n".join(l for l in txt.splitlines() if not l.strip().startswith(("//", "#", "*", "/*")))
                assert "oracle" not in code.lower(), f"{f} references the oracle: the product must not route through it"


def test_datetime_delegation_argument_errors_without_gpu():
    """apply_datetime_* (dispatch.rs:300-372) rejects bad arguments before any device is touched: the reference's
    `confirm_equal_len` (LengthMismatch), a window outside the array, a mask shorter than the window."""
    import numpy as np
    import pytest
    import minarrow_b200 as mnr
    from minarrow_b200.kernels import arithmetic as ar
    A = mnr.ArithmeticOperator
    a = mnr.DatetimeArray.from_slice(np.array([1000, 2000], np.int64))
    b = mnr.DatetimeArray.from_slice(np.array([10], np.int64))
    with pytest.raises(mnr.KernelError) as ei:
        ar.apply_datetime_i64((a, 0, 2), (b, 0, 1), A.Add)
    assert ei.value.kind == "LengthMismatch" and "apply_datetime: length mismatch" in str(ei.value)
    with pytest.raises(mnr.KernelError) as ei:
        ar.apply_datetime_i64((a, 1, 2), (a, 0, 2), A.Add)
    assert ei.value.kind == "OutOfBounds"
    short = mnr.DatetimeArray(a.data, mnr.Bitmask.from_bools([True]))
    with pytest.raises(mnr.KernelError) as ei:
        ar.apply_datetime_i64((short, 0, 2), (a, 0, 2), A.Add)
    assert ei.value.kind == "InvalidArguments"
    e = mnr.DatetimeArray.from_slice(np.zeros(0, np.int64))
    assert ar.apply_datetime_i64((e, 0, 0), (e, 0, 0), A.Add).is_empty()      # empty in, empty out: no launch


def test_scalar_arithmetic_and_the_scalar_scalar_arm():
    """`scalar_arithmetic` (routing/arithmetic.rs:34-209) is host arithmetic on two scalars; `broadcast_value` on two Scalars
    calls it with ArithmeticOperator::Add whatever the operator (broadcast/mod.rs:161-163) — reproduced as is."""
    import numpy as np
    import pytest
    import minarrow_b200 as mnr
    from minarrow_b200.containers import broadcast_value, scalar_arithmetic
    A = mnr.ArithmeticOperator
    assert scalar_arithmetic(np.int32(7), np.int32(-2), A.Divide) == -3 and scalar_arithmetic(np.int32(7), np.int32(-2), A.Divide).dtype == np.int32
    assert scalar_arithmetic(np.int64(2 ** 62), np.int64(2 ** 62), A.Add) == np.int64(-2 ** 63)           # wraps like the release build
    assert scalar_arithmetic(np.int64(3), np.float64(0.5), A.Multiply) == 1.5                             # Int + Float = Float
    assert scalar_arithmetic(np.float32(1), np.float32(3), A.Divide).dtype == np.float32
    assert scalar_arithmetic(np.uint32(5), np.uint32(7), A.Subtract) == np.uint32(2 ** 32 - 2)
    for bad in ((np.int32(1), np.int64(2), A.Add), (np.int64(1), np.int64(2), A.Power)):
        with pytest.raises(mnr.KernelError) as ei:
            scalar_arithmetic(*bad)
        assert ei.value.kind == "NotImplemented"
    assert broadcast_value(A.Multiply, np.int64(6), np.int64(7)) == 13
    assert broadcast_value(A.Subtract, 1.5, 2) == 3.5


def test_documents_cite_evidence_files_that_exist():
    """Every `profiles/...` path DESIGN.md / README.md / INTEGRATION.md / profiles/README.md cite is in the tree, and
    `profiles/traffic.json` (read by bench.py for `roofline.traffic`) names a source capture that exists."""
    import json
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    missing = []
    for doc in ("DESIGN.md", "README.md", "INTEGRATION.md", os.path.join("profiles", "README.md")):
        text = open(os.path.join(root, doc)).read()
        for m in re.finditer(r"profiles/[A-Za-z0-9_./-]+", text):
            p = m.group(0).rstrip(".,)/")
            if "..." in p or p.endswith("_"):
                continue
            if not os.path.exists(os.path.join(root, p)):
                missing.append((doc, p))
    assert not missing, missing
    t = json.load(open(os.path.join(root, "profiles", "traffic.json")))
    src = re.search(r"profiles/[A-Za-z0-9_./-]+", t["source"]).group(0)
    assert os.path.exists(os.path.join(root, src)) and t["reduce_stats_kernel_i64_masked"] > 8_000_000_000
