"""Adapters giving the CPU oracle and the CUDA path one test-facing shape (see kat_runner.py)."""
import numpy as np

from oracle import oracle as orc


def _bits(b):
    return None if b is None else orc.Bits.from_bools(b)


class OracleBackend:
    name = "oracle"

    def supports_dtype(self, name):
        return True

    def apply(self, lhs, rhs, op, mask):
        data, m = orc.apply(lhs, rhs, op, _bits(mask))
        return data, (None if m is None else m.to_bools())

    def apply_fma(self, lhs, rhs, acc, mask):
        data, m = orc.apply_fma(lhs, rhs, acc, _bits(mask))
        return data, (None if m is None else m.to_bools())

    def merge_and(self, a, b):
        return orc.merge_bitmasks_to_new(_bits(a), _bits(b), len(a)).to_bools()

    def bits_binop(self, op, a, b):
        return orc.bitmask_binop((_bits(a), 0, len(a)), (_bits(b), 0, len(b)), op).to_bools()

    def bits_not(self, a):
        return orc.not_mask((_bits(a), 0, len(a))).to_bools()

    def bits_invert(self, a):
        return orc.invert(_bits(a)).to_bools()

    def bits_in(self, a, b, n):
        return orc.in_mask((_bits(a), 0, n), (_bits(b), 0, n)).to_bools()

    def bits_not_in(self, a, b, n):
        return orc.not_in_mask((_bits(a), 0, n), (_bits(b), 0, n)).to_bools()

    def bits_eq(self, a, b, n):
        return orc.eq_mask((_bits(a), 0, n), (_bits(b), 0, n)).to_bools()

    def bits_ne(self, a, b, n):
        return orc.ne_mask((_bits(a), 0, n), (_bits(b), 0, n)).to_bools()

    def bits_union(self, a, b, n):
        return orc.union(_bits(a), _bits(b)).to_bools()

    def bits_intersect(self, a, b, n):
        return orc.intersect(_bits(a), _bits(b)).to_bools()

    def bits_all_eq(self, a, b):
        return orc.all_eq((_bits(a), 0, len(a)), (_bits(b), 0, len(b)))

    def bits_all_ne(self, a, b):
        return orc.all_ne((_bits(a), 0, len(a)), (_bits(b), 0, len(b)))

    def bits_popcount(self, a):
        m = _bits(a)
        p = orc.popcount_mask((m, 0, len(a)))
        assert p == orc.count_ones(m)
        return p

    def bits_all_true(self, a):
        return orc.all_true_mask(_bits(a))

    def bits_all_false(self, a):
        return orc.all_false_mask(_bits(a))

    def bytes_set_all(self, n, value):
        return orc.new_set_all(n, value).bits

    def bytes_from_bools(self, b):
        return _bits(b).bits

    def route(self, op, lhs, rhs):
        data, m = orc.resolve_binary_arithmetic(op, lhs, rhs, None)
        return data, (None if m is None else m.to_bools())

    def super_route(self, op, lhs_chunks, rhs_chunks):
        """route_super_array_broadcast — src/kernels/broadcast/super_array.rs:180-249, no masks."""
        out = []
        for l, r in zip(lhs_chunks, rhs_chunks):
            if len(l) != len(r):
                raise orc.KernelError("ShapeError", "Super Array broadcasting error")
            out.append(orc.resolve_binary_arithmetic(op, l, r, None)[0])
        return out

    def sum_i64(self, d):
        a, b, c = orc.simd_sum_i64(d), orc.hotloop_sum_i64(d), orc.rayon_simd_sum_i64(d, threads=2)
        assert a == b == c
        return a

    def sum_f64(self, d):
        return orc.rayon_simd_sum_f64(d)
