"""Runs the reference's known-answer tests (tests/golden/reference_kats.json) against a backend.

A backend is any object with the methods used below; tests/backends.py provides one over the CPU
oracle and one over the CUDA path (through the C ABI), so the same transcribed reference tests pin both.
"""
import json
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NP = {"i8": np.int8, "u8": np.uint8, "i16": np.int16, "u16": np.uint16, "i32": np.int32, "u32": np.uint32,
      "i64": np.int64, "u64": np.uint64, "f32": np.float32, "f64": np.float64}
OPS = {"add": 0, "subtract": 1, "multiply": 2, "divide": 3, "remainder": 4, "power": 5, "floordiv": 6}
LOPS = {"and": 0, "or": 1, "xor": 2}


def load_cases():
    with open(os.path.join(HERE, "golden", "reference_kats.json")) as f:
        return json.load(f)["cases"]


def case_id(c):
    return f'{c["kind"]}-{c.get("dtype", "")}-{c.get("op", "")}-{c["ref"].split("/")[-1]}'


def _bools(x):
    return None if x is None else np.asarray(x, dtype=bool)


def _fill(c):
    if "a_fill" in c:
        n, v = c["a_fill"]
        a = np.full(n, bool(v))
        for i in c.get("clear", []):
            a[i] = False
        return a
    return _bools(c["a"])


def run_case(c, be, allow_heavy=False):
    kind = c["kind"]
    if c.get("heavy") and not allow_heavy:
        return "skipped"
    if kind in ("apply_int", "apply_float"):
        dt = NP[c["dtype"]]
        if not be.supports_dtype(c["dtype"]):
            return "skipped"
        lhs, rhs = np.array(c["lhs"], dtype=dt), np.array(c["rhs"], dtype=dt)
        mask = _bools(c["mask"])
        if c.get("expect_error"):
            try:
                be.apply(lhs, rhs, OPS[c["op"]], mask)
            except Exception as e:  # noqa: BLE001
                assert getattr(e, "kind", None) == c["expect_error"], (c["ref"], repr(e))
                return "ok"
            raise AssertionError(f'{c["ref"]}: expected {c["expect_error"]}')
        data, valid = be.apply(lhs, rhs, OPS[c["op"]], mask)
        assert data.dtype == dt and data.shape == lhs.shape, c["ref"]
        if c.get("expect_special") == "all_inf":
            assert np.all(np.isinf(data)), c["ref"]
        elif c.get("expect_special") == "all_nan":
            assert np.all(np.isnan(data)), c["ref"]
        else:
            exp = np.array(c["expect_data"], dtype=dt)
            eps = c.get("eps", 0.0)
            if eps == 0.0 or kind == "apply_int":
                assert np.array_equal(data, exp), (c["ref"], data, exp)
            elif c.get("rel"):
                assert np.all(np.abs(data - exp) <= eps * np.maximum(1.0, np.abs(exp))), (c["ref"], data, exp)
            else:
                assert np.all(np.abs(data - exp) < eps), (c["ref"], data, exp)
        if c["expect_valid"] is None:
            assert valid is None, (c["ref"], "dense call must not produce a mask (dispatch.rs:99-104)")
        else:
            assert valid is not None and np.array_equal(valid, _bools(c["expect_valid"])), (c["ref"], valid)
        return "ok"
    if kind == "apply_fma":
        dt = NP[c["dtype"]]
        lhs, rhs, acc = (np.array(c[k], dtype=dt) for k in ("lhs", "rhs", "acc"))
        data, valid = be.apply_fma(lhs, rhs, acc, _bools(c["mask"]))
        assert np.array_equal(data, np.array(c["expect_data"], dtype=dt)), (c["ref"], data)
        if c["expect_valid"] is None:
            assert valid is None, c["ref"]
        else:
            assert np.array_equal(valid, _bools(c["expect_valid"])), c["ref"]
        return "ok"
    if kind == "merge_and":
        out = be.merge_and(_bools(c["a"]), _bools(c["b"]))
        assert np.array_equal(out, _bools(c["expect"])), c["ref"]
        return "ok"
    if kind == "bits_binop":
        out = be.bits_binop(LOPS[c["op"]], _bools(c["a"]), _bools(c["b"]))
        assert np.array_equal(out, _bools(c["expect"])), (c["ref"], out)
        return "ok"
    if kind in ("bits_not", "bits_invert"):
        out = be.bits_not(_bools(c["a"])) if kind == "bits_not" else be.bits_invert(_bools(c["a"]))
        assert np.array_equal(out, _bools(c["expect"])), c["ref"]
        return "ok"
    if kind in ("bits_in", "bits_not_in", "bits_eq", "bits_ne", "bits_union", "bits_intersect"):
        a, b = _bools(c["a"]), _bools(c["b"])
        n = c.get("len", len(a))
        out = getattr(be, kind)(a, b, n)
        assert np.array_equal(out, _bools(c["expect"])), (c["ref"], out)
        return "ok"
    if kind in ("bits_all_eq", "bits_all_ne"):
        assert getattr(be, kind)(_bools(c["a"]), _bools(c["b"])) == c["expect"], c["ref"]
        return "ok"
    if kind == "bits_popcount":
        assert be.bits_popcount(_bools(c["a"])) == c["expect"], c["ref"]
        return "ok"
    if kind in ("bits_all_true", "bits_all_false"):
        assert getattr(be, kind)(_fill(c)) == c["expect"], c["ref"]
        return "ok"
    if kind == "bits_count":
        a = _fill(c)
        ones = be.bits_popcount(a)
        assert ones == c["expect_ones"] and len(a) - ones == c["expect_zeros"], c["ref"]
        return "ok"
    if kind == "bytes_set_all":
        assert list(be.bytes_set_all(c["len"], c["value"])) == c["expect_bytes"], c["ref"]
        return "ok"
    if kind == "bytes_from_bools":
        by = be.bytes_from_bools(_bools(c["bools"]))
        words = np.frombuffer(bytes(by), dtype="<u8")
        assert [int(w) for w in words] == c["expect_words"], c["ref"]
        return "ok"
    if kind == "route":
        lhs, rhs = np.array(c["lhs"], dtype=NP[c["ldtype"]]), np.array(c["rhs"], dtype=NP[c["rdtype"]])
        data, valid = be.route(OPS[c["op"]], lhs, rhs)
        assert data.dtype == NP[c["expect_dtype"]], c["ref"]
        assert np.array_equal(data, np.array(c["expect_data"], dtype=data.dtype)), (c["ref"], data)
        assert valid is None, c["ref"]
        return "ok"
    if kind == "super_route":
        dt = NP[c["dtype"]]
        lc = [np.array(x, dtype=dt) for x in c["lhs_chunks"]]
        rc = [np.array(x, dtype=dt) for x in c["rhs_chunks"]]
        if c.get("expect_error"):
            try:
                be.super_route(OPS[c["op"]], lc, rc)
            except Exception as e:  # noqa: BLE001
                assert getattr(e, "kind", None) == c["expect_error"], (c["ref"], repr(e))
                return "ok"
            raise AssertionError(f'{c["ref"]}: expected {c["expect_error"]}')
        out = be.super_route(OPS[c["op"]], lc, rc)
        assert len(out) == len(c["expect_chunks"]), c["ref"]
        for got, exp in zip(out, c["expect_chunks"]):
            assert np.array_equal(got, np.array(exp, dtype=dt)), (c["ref"], got)
        return "ok"
    if kind == "sum_arange":
        n = c["n"]
        if c["dtype"] == "i64":
            assert be.sum_i64(np.arange(n, dtype=np.int64)) == c["expect"], c["ref"]
        else:
            got = be.sum_f64(np.arange(n, dtype=np.float64))
            assert math.isclose(got, c["expect"], rel_tol=1e-12), (c["ref"], got)
        return "ok"
    raise AssertionError(f"unknown case kind {kind}")
