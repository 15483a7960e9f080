#!/usr/bin/env python3
"""Writes tests/golden/reference_kats.json — the reference's own known-answer tests for the hot path.

The reference (pbower/minarrow v0.10.1) is Rust and cannot be executed in this image, so these vectors
are TRANSCRIBED from its in-tree `#[test]`s / examples; every case carries the reference `file:line` it
was read from.  Literals only: nothing here is computed by this repo's oracle or kernels, so the file
pins the oracle (tests/test_oracle_golden.py) and, through the C ABI, the CUDA path
(tests/test_gpu_golden.py).  Re-run after editing:  python tests/golden/make_golden.py
"""
import json
import os

T, F = True, False
INT_TYPES = ["i32", "u32", "i64", "u64", "i8", "u8", "i16", "u16"]
cases = []


def add(**kw):
    cases.append(kw)


# ---- src/kernels/arithmetic/mod.rs:117-230  int_kernel_suite! (instantiated :232-291) --------------
for ty in INT_TYPES:
    bits = int(ty[1:])
    src = "src/kernels/arithmetic/mod.rs"
    lhs, rhs = [1, 4, 9, 16], [1, 2, 3, 4]
    for op, exp, line in [("add", [2, 6, 12, 20], "124-129"), ("subtract", [0, 2, 6, 12], "131-132"),
                          ("multiply", [1, 8, 27, 64], "134-139"), ("divide", [1, 2, 3, 4], "141-142"),
                          ("remainder", [0, 0, 0, 0], "144-145")]:
        add(kind="apply_int", dtype=ty, op=op, lhs=lhs, rhs=rhs, mask=None, expect_data=exp,
            expect_valid=None, ref=f"{src}:{line}")
    # :147-159 Power == repeated wrapping_mul: 1^1, 4^2, 9^3, 16^4 (wraps for 8/16-bit types)
    powv = [(1 ** 1) % (1 << bits), (4 ** 2) % (1 << bits), (9 ** 3) % (1 << bits), (16 ** 4) % (1 << bits)]
    if ty[0] == "i":
        powv = [v - (1 << bits) if v >= (1 << (bits - 1)) else v for v in powv]
    add(kind="apply_int", dtype=ty, op="power", lhs=lhs, rhs=rhs, mask=None, expect_data=powv,
        expect_valid=None, ref=f"{src}:147-159")
    # :161-177 dense /0 and %0 must panic
    for op in ("divide", "remainder"):
        add(kind="apply_int", dtype=ty, op=op, lhs=lhs, rhs=[0, 0, 0, 0], mask=None,
            expect_error="DivideByZero", ref=f"{src}:161-177")
    # :182-201 masked Div/Rem
    add(kind="apply_int", dtype=ty, op="divide", lhs=[10, 20, 30, 40], rhs=[2, 0, 3, 5], mask=[T, F, T, F],
        expect_data=[5, 0, 10, 0], expect_valid=[T, F, T, F], ref=f"{src}:182-192")
    add(kind="apply_int", dtype=ty, op="remainder", lhs=[10, 20, 30, 40], rhs=[2, 0, 3, 5], mask=[T, F, T, F],
        expect_data=[0, 0, 0, 0], expect_valid=[T, F, T, F], ref=f"{src}:194-201")
    # :203-219 all-valid mask, zero divisors -> value 0 + null
    add(kind="apply_int", dtype=ty, op="divide", lhs=[100, 100, 100, 100], rhs=[1, 0, 2, 0], mask=[T, T, T, T],
        expect_data=[100, 0, 50, 0], expect_valid=[T, F, T, F], ref=f"{src}:203-219")
    # :222-228 empty
    add(kind="apply_int", dtype=ty, op="add", lhs=[], rhs=[], mask=None, expect_data=[], expect_valid=None,
        ref=f"{src}:222-228")

# ---- src/kernels/arithmetic/mod.rs:293-370  float_kernel_suite! -----------------------------------
for ty, eps in (("f32", 1e-6), ("f64", 1e-12)):
    src = "src/kernels/arithmetic/mod.rs"
    lhs, rhs = [1.0, 4.0, 9.0, 16.0], [0.5, 2.0, 3.0, 4.0]
    for op, exp, line in [("add", [1.5, 6.0, 12.0, 20.0], "303-304"), ("subtract", [0.5, 2.0, 6.0, 12.0], "306-307"),
                          ("multiply", [0.5, 8.0, 27.0, 64.0], "309-310"), ("divide", [2.0, 2.0, 3.0, 4.0], "312-313")]:
        add(kind="apply_float", dtype=ty, op=op, lhs=lhs, rhs=rhs, mask=None, expect_data=exp, eps=0.0,
            expect_valid=None, ref=f"{src}:{line}")
    add(kind="apply_float", dtype=ty, op="remainder", lhs=lhs, rhs=rhs, mask=None, expect_data=[0.0, 0.0, 0.0, 0.0],
        eps=eps, expect_valid=None, ref=f"{src}:315-326")
    # :328-340 Power vs exp(b*ln a) within eps: 1^0.5, 4^2, 9^3, 16^4
    add(kind="apply_float", dtype=ty, op="power", lhs=lhs, rhs=rhs, mask=None,
        expect_data=[1.0, 16.0, 729.0, 65536.0], eps=eps, rel=True, expect_valid=None, ref=f"{src}:328-340")
    add(kind="apply_float", dtype=ty, op="divide", lhs=lhs, rhs=[0.0] * 4, mask=None, expect_special="all_inf",
        expect_valid=None, ref=f"{src}:342-348")
    add(kind="apply_float", dtype=ty, op="remainder", lhs=lhs, rhs=[0.0] * 4, mask=None, expect_special="all_nan",
        expect_valid=None, ref=f"{src}:350-354")
    add(kind="apply_float", dtype=ty, op="multiply", lhs=lhs, rhs=rhs, mask=[T, F, T, F],
        expect_data=[0.5, 0.0, 27.0, 0.0], eps=0.0, expect_valid=[T, F, T, F], ref=f"{src}:356-360")
    add(kind="apply_float", dtype=ty, op="add", lhs=[], rhs=[], mask=None, expect_data=[], eps=0.0,
        expect_valid=None, ref=f"{src}:362-364")
    # ---- :372-399 FMA
    add(kind="apply_fma", dtype=ty, lhs=[1.0, 2.0, 3.0], rhs=[4.0, 5.0, 6.0], acc=[0.5, 0.5, 0.5], mask=None,
        expect_data=[4.5, 10.5, 18.5], expect_valid=None, ref=f"{src}:372-378,388-394")
    add(kind="apply_fma", dtype=ty, lhs=[1.0, 2.0, 3.0], rhs=[4.0, 5.0, 6.0], acc=[0.5, 0.5, 0.5], mask=[T, F, T],
        expect_data=[4.5, 0.0, 18.5], expect_valid=[T, F, T], ref=f"{src}:380-382,396-398")

# ---- :401-409 merge_bitmasks_to_new == AND
add(kind="merge_and", a=[T, F, T, T], b=[T, T, F, T], expect=[T, F, F, T],
    ref="src/kernels/arithmetic/mod.rs:401-409")
# ---- :507-537 SIMD int power short vs long
for n in (16, 128):
    add(kind="apply_int", dtype="u32", op="power", lhs=[2] * n, rhs=[10] * n, mask=None, expect_data=[1024] * n,
        expect_valid=None, ref="src/kernels/arithmetic/mod.rs:507-537")
# ---- datetime delegation (:430-494): integer kernels over i64 with AND-merged masks
add(kind="apply_int", dtype="i64", op="add", lhs=[1000, 2000, 3000], rhs=[10, 20, 30], mask=None,
    expect_data=[1010, 2020, 3030], expect_valid=None, ref="src/kernels/arithmetic/mod.rs:419-428")
for op, exp in [("add", [11, 22, 33, 44]), ("subtract", [9, 18, 27, 36]), ("multiply", [10, 40, 90, 160]),
                ("divide", [10, 10, 10, 10]), ("remainder", [0, 0, 0, 0]),
                ("power", [10, 400, 27000, 2560000])]:
    add(kind="apply_int", dtype="i64", op=op, lhs=[10, 20, 30, 40], rhs=[1, 2, 3, 4], mask=None, expect_data=exp,
        expect_valid=None, ref="src/kernels/arithmetic/mod.rs:430-458")
add(kind="apply_int", dtype="i64", op="add", lhs=[10, 20, 30, 40], rhs=[1, 2, 3, 4], mask=[T, F, T, T],
    expect_data=[11, 0, 33, 44], expect_valid=[T, F, T, T], ref="src/kernels/arithmetic/mod.rs:472-485")

# ---- src/kernels/bitmask/simd.rs:817-945  simd_bitmask_suite! ---------------------------------------
S = "src/kernels/bitmask/simd.rs"
a8 = [T, F, T, F, T, T, F, F]
b8 = [T, T, F, F, T, F, T, F]
add(kind="bits_binop", op="and", a=a8, b=b8, expect=[x & y for x, y in zip(a8, b8)], ref=f"{S}:817-824")
add(kind="bits_binop", op="or", a=a8, b=b8, expect=[x | y for x, y in zip(a8, b8)], ref=f"{S}:826-834")
add(kind="bits_binop", op="xor", a=a8, b=b8, expect=[x ^ y for x, y in zip(a8, b8)], ref=f"{S}:836-844")
add(kind="bits_not", a=[T, F, T, F], expect=[F, T, F, T], ref=f"{S}:846-853")
lhs4 = [T, F, T, F]
add(kind="bits_in", a=lhs4, b=[T] * 4, expect=lhs4, ref=f"{S}:856-863")
add(kind="bits_in", a=lhs4, b=[F] * 4, expect=[F, T, F, T], ref=f"{S}:864-869")
add(kind="bits_in", a=lhs4, b=[T, F, T, F], expect=[T] * 4, ref=f"{S}:870-875")
add(kind="bits_in", a=lhs4, b=[], len=0, expect=[], ref=f"{S}:876-879")
add(kind="bits_not_in", a=lhs4, b=lhs4, expect=[F] * 4, ref=f"{S}:882-891")
add(kind="bits_eq", a=[T, F, T, F], b=[T, F, F, T], expect=[T, T, F, F], ref=f"{S}:893-903")
add(kind="bits_ne", a=[T, F, T, F], b=[T, F, F, T], expect=[F, F, T, T], ref=f"{S}:893-903")
add(kind="bits_all_eq", a=a8, b=a8, expect=True, ref=f"{S}:905-913")
add(kind="bits_all_eq", a=a8, b=[F] + a8[1:], expect=False, ref=f"{S}:905-913")
add(kind="bits_all_ne", a=[T, F, T], b=[F, T, F], expect=True, ref=f"{S}:915-921")
add(kind="bits_all_ne", a=[T, F, T], b=[T, F, T], expect=False, ref=f"{S}:915-921")
add(kind="bits_popcount", a=[T, F, T, F, T, F, F, T], expect=4, ref=f"{S}:923-928")
for lanes in (8, 16, 32, 64):  # W8/W16/W32/W64 with AVX-512 lane table, build.rs:55-110
    n = 64 * lanes
    add(kind="bits_all_true", a_fill=[n, True], clear=[], expect=True, ref=f"{S}:930-937")
    add(kind="bits_all_true", a_fill=[n, True], clear=[3], expect=False, ref=f"{S}:930-937")
    add(kind="bits_all_false", a_fill=[n, True], clear=[], expect=False, ref=f"{S}:939-945")
    add(kind="bits_all_false", a_fill=[n, False], clear=[], expect=True, ref=f"{S}:939-945")

# ---- src/kernels/bitmask/std.rs:384-540 --------------------------------------------------------------
S = "src/kernels/bitmask/std.rs"
add(kind="bits_binop", op="and", a=[T, F, T, T, F, F, T, T], b=[F, F, T, F, T, F, T, F],
    expect=[F, F, T, F, F, F, T, F], ref=f"{S}:384-393")
add(kind="bits_binop", op="or", a=[T, F, T, T], b=[F, F, T, F], expect=[T, F, T, T], ref=f"{S}:395-404")
add(kind="bits_binop", op="xor", a=[T, F, T, F], b=[F, T, T, F], expect=[T, T, F, F], ref=f"{S}:406-415")
add(kind="bits_not", a=[T, F, T, F], expect=[F, T, F, T], ref=f"{S}:417-425")
add(kind="bits_in", a=[T, F, T], b=[T, F, T], expect=[T, T, T], ref=f"{S}:427-435")
add(kind="bits_in", a=[T, F, T], b=[T, T, T], expect=[T, F, T], ref=f"{S}:437-446")
add(kind="bits_in", a=[T, F, T], b=[F, F, F], expect=[F, T, F], ref=f"{S}:448-457")
add(kind="bits_not_in", a=[T, F], b=[T, F], expect=[F, F], ref=f"{S}:459-468")
add(kind="bits_eq", a=[T, F, T], b=[T, F, F], expect=[T, T, F], ref=f"{S}:470-479")
add(kind="bits_ne", a=[T, F, T], b=[T, T, F], expect=[F, T, T], ref=f"{S}:481-490")
add(kind="bits_all_eq", a=[T, F, T, F], b=[T, F, T, F], expect=True, ref=f"{S}:492-497")
add(kind="bits_all_eq", a=[T, F, T, F], b=[F, T, F, T], expect=False, ref=f"{S}:499-504")
add(kind="bits_all_ne", a=[T, F], b=[F, T], expect=True, ref=f"{S}:506-511")
add(kind="bits_all_ne", a=[T, F], b=[T, F], expect=False, ref=f"{S}:513-518")
add(kind="bits_popcount", a=[T, F, T, F, T, T], expect=4, ref=f"{S}:520-524")
add(kind="bits_all_true", a=[T, T, T, T], expect=True, ref=f"{S}:526-532")
add(kind="bits_all_true", a=[T, T, F, T], expect=False, ref=f"{S}:526-532")
add(kind="bits_all_false", a=[F, F, F, F], expect=True, ref=f"{S}:534-540")
add(kind="bits_all_false", a=[F, T, F, F], expect=False, ref=f"{S}:534-540")

# ---- src/kernels/bitmask/mod.rs:203-287 ----------------------------------------------------------------
add(kind="bytes_set_all", len=10, value=True, expect_bytes=[0xFF, 0x03], ref="src/kernels/bitmask/mod.rs:272-287")
add(kind="bytes_from_bools", bools=[T] + [F] * 62 + [T, T] + [F] * 62 + [T],
    expect_words=[(1 | (1 << 63)), (1 | (1 << 63))], ref="src/kernels/bitmask/mod.rs:214-228")

# ---- src/structs/bitmask.rs:943-1080 --------------------------------------------------------------------
S = "src/structs/bitmask.rs"
add(kind="bits_count", a_fill=[16, True], clear=[], expect_ones=16, expect_zeros=0, ref=f"{S}:943-946")
add(kind="bits_count", a_fill=[16, True], clear=[0], expect_ones=15, expect_zeros=1, ref=f"{S}:947-950")
ua = [F, T, F, T, F, F, F, F]
ub = [F, F, F, T, T, F, F, F]
add(kind="bits_union", a=ua, b=ub, expect=[x | y for x, y in zip(ua, ub)], ref=f"{S}:953-962")
add(kind="bits_intersect", a=ua, b=ub, expect=[x & y for x, y in zip(ua, ub)], ref=f"{S}:963-964")
add(kind="bits_invert", a=ua, expect=[not x for x in ua], ref=f"{S}:965-966")
add(kind="bits_union", a=[T, F, F, T], b=[F, T, F, T], expect=[T, T, F, T], ref=f"{S}:1071-1080")

# ---- routing / broadcast container-level vectors (all via resolve_binary_arithmetic) -------------------
R = "src/kernels/broadcast"
add(kind="route", op="add", ldtype="i32", lhs=[1, 2, 3], rdtype="i32", rhs=[5], expect_dtype="i32",
    expect_data=[6, 7, 8], ref=f"{R}/super_array.rs:479-491")
add(kind="route", op="add", ldtype="i32", lhs=[5], rdtype="i32", rhs=[1, 2, 3], expect_dtype="i32",
    expect_data=[6, 7, 8], ref=f"{R}/super_array.rs:493-506")
add(kind="route", op="add", ldtype="i32", lhs=[1, 2, 3], rdtype="i32", rhs=[4, 5, 6], expect_dtype="i32",
    expect_data=[5, 7, 9], ref=f"{R}/super_array.rs:508-521")
add(kind="route", op="divide", ldtype="i32", lhs=[50], rdtype="i32", rhs=[100, 200, 300], expect_dtype="i32",
    expect_data=[0, 0, 0], ref=f"{R}/scalar.rs:1332-1350")
add(kind="route", op="multiply", ldtype="i32", lhs=[10, 20, 30], rdtype="i32", rhs=[5], expect_dtype="i32",
    expect_data=[50, 100, 150], ref=f"{R}/scalar.rs:1352-1370")
add(kind="route", op="multiply", ldtype="i32", lhs=[2], rdtype="i32", rhs=[2, 4, 6], expect_dtype="i32",
    expect_data=[4, 8, 12], ref=f"{R}/scalar.rs:1284-1305")
for op, exp in [("add", [12, 24, 36]), ("subtract", [8, 16, 24]), ("multiply", [20, 80, 180]),
                ("divide", [5, 5, 5]), ("remainder", [0, 0, 0])]:
    add(kind="route", op=op, ldtype="i32", lhs=[10, 20, 30], rdtype="i32", rhs=[2, 4, 6], expect_dtype="i32",
        expect_data=exp, ref="examples/arithmetic.rs:19-58")
# SuperArray route, per chunk (super_array.rs:524-565, 596-640)
add(kind="super_route", op="add", dtype="i32", lhs_chunks=[[1, 2, 3], [4, 5, 6]],
    rhs_chunks=[[10, 10, 10], [20, 20, 20]], expect_chunks=[[11, 12, 13], [24, 25, 26]],
    ref=f"{R}/super_array.rs:524-565")
add(kind="super_route", op="multiply", dtype="i32", lhs_chunks=[[2, 3, 4], [5, 6, 7]],
    rhs_chunks=[[10, 10, 10], [2, 2, 2]], expect_chunks=[[20, 30, 40], [10, 12, 14]],
    ref=f"{R}/super_array.rs:596-640")
add(kind="super_route", op="divide", dtype="i32", lhs_chunks=[[100, 200, 300]], rhs_chunks=[[10, 20, 30]],
    expect_chunks=[[10, 10, 10]], ref=f"{R}/super_array.rs:642-672")
add(kind="super_route", op="add", dtype="i32", lhs_chunks=[[1, 2, 3]], rhs_chunks=[[10, 10]],
    expect_error="ShapeError", ref=f"{R}/super_array.rs:567-594")
# routing/binary_map.rs:72-152 (f64 add / mul with len-1 broadcast)
add(kind="route", op="add", ldtype="f64", lhs=[1.0, 2.0, 3.0], rdtype="f64", rhs=[10.0, 20.0, 30.0],
    expect_dtype="f64", expect_data=[11.0, 22.0, 33.0], ref="src/kernels/routing/binary_map.rs:84-97")
add(kind="route", op="multiply", ldtype="f64", lhs=[1.0, 2.0, 3.0], rdtype="f64", rhs=[10.0],
    expect_dtype="f64", expect_data=[10.0, 20.0, 30.0], ref="src/kernels/routing/binary_map.rs:99-112")

# ---- bench self-checks (benches/*.rs) ---------------------------------------------------------------------
add(kind="sum_arange", dtype="i64", n=1000, expect=499500, ref="benches/hotloop_benchmark_simd.rs:36,199-210")
add(kind="sum_arange", dtype="f64", n=1000, expect=499500.0, ref="benches/hotloop_benchmark_simd.rs:316-320")
add(kind="sum_arange", dtype="i64", n=1000000000, expect=499999999500000000, heavy=True,
    ref="benches/benchmark_parallel_simd.rs:39,103-112")

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")
with open(out, "w") as f:
    json.dump({"reference": "pbower/minarrow v0.10.1", "cases": cases}, f, indent=0)
print(f"wrote {len(cases)} cases -> {out}")
