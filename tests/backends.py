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


class GpuBackend:
    """The CUDA path through the C ABI, via the reference-named host mirror (minarrow_b200.kernels.*)."""
    name = "gpu"

    def __init__(self, ctx=None):
        import minarrow_b200 as mnr
        self.mnr = mnr
        self.ctx = ctx or mnr.default_context()
        self.ar = mnr.kernels.arithmetic
        self.bm = mnr.kernels.bitmask

    def _b(self, b):
        return None if b is None else self.mnr.Bitmask.from_bools(b)

    def supports_dtype(self, name):
        return True

    def apply(self, lhs, rhs, op, mask):
        out = self.ar.APPLY[lhs.dtype](lhs, rhs, op, self._b(mask), self.ctx)
        return out.data, (None if out.null_mask is None else out.null_mask.to_bools())

    def apply_fma(self, lhs, rhs, acc, mask):
        f = self.ar.apply_fma_f32 if lhs.dtype == np.float32 else self.ar.apply_fma_f64
        out = f(lhs, rhs, acc, self._b(mask), self.ctx)
        return out.data, (None if out.null_mask is None else out.null_mask.to_bools())

    def merge_and(self, a, b):
        return self.bm.merge_bitmasks_to_new(self._b(a), self._b(b), len(a), self.ctx).to_bools()

    def bits_binop(self, op, a, b):
        return self.bm.bitmask_binop((self._b(a), 0, len(a)), (self._b(b), 0, len(b)), op, self.ctx).to_bools()

    def bits_not(self, a):
        return self.bm.not_mask((self._b(a), 0, len(a)), self.ctx).to_bools()

    def bits_invert(self, a):
        return self.bm.invert(self._b(a), self.ctx).to_bools()

    def bits_in(self, a, b, n):
        return self.bm.in_mask((self._b(a), 0, n), (self._b(b), 0, n), self.ctx).to_bools()

    def bits_not_in(self, a, b, n):
        return self.bm.not_in_mask((self._b(a), 0, n), (self._b(b), 0, n), self.ctx).to_bools()

    def bits_eq(self, a, b, n):
        return self.bm.eq_mask((self._b(a), 0, n), (self._b(b), 0, n), self.ctx).to_bools()

    def bits_ne(self, a, b, n):
        return self.bm.ne_mask((self._b(a), 0, n), (self._b(b), 0, n), self.ctx).to_bools()

    def bits_union(self, a, b, n):
        return self.bm.union(self._b(a), self._b(b), self.ctx).to_bools()

    def bits_intersect(self, a, b, n):
        return self.bm.intersect(self._b(a), self._b(b), self.ctx).to_bools()

    def bits_all_eq(self, a, b):
        return self.bm.all_eq((self._b(a), 0, len(a)), (self._b(b), 0, len(b)), self.ctx)

    def bits_all_ne(self, a, b):
        return self.bm.all_ne((self._b(a), 0, len(a)), (self._b(b), 0, len(b)), self.ctx)

    def bits_popcount(self, a):
        m = self._b(a)
        p = self.bm.popcount_mask((m, 0, len(a)), self.ctx)
        assert p == self.bm.count_ones(m, self.ctx)
        return p

    def bits_all_true(self, a):
        return self.bm.all_true_mask(self._b(a), self.ctx)

    def bits_all_false(self, a):
        return self.bm.all_false_mask(self._b(a), self.ctx)

    def bytes_set_all(self, n, value):
        return self.mnr.DeviceBitmask.new_set_all(self.ctx, n, value).download().bits

    def bytes_from_bools(self, b):
        m = self._b(b)
        return self.mnr.DeviceBitmask.upload(self.ctx, m).download().bits

    def route(self, op, lhs, rhs):
        out = self.mnr.resolve_binary_arithmetic(op, lhs, rhs, None, self.ctx)
        return out.data, (None if out.null_mask is None else out.null_mask.to_bools())

    def super_route(self, op, lhs_chunks, rhs_chunks):
        SA, IA = self.mnr.SuperArray, self.mnr.IntegerArray
        out = self.mnr.route_super_array_broadcast(op, SA([IA(c) for c in lhs_chunks]), SA([IA(c) for c in rhs_chunks]),
                                                   None, self.ctx)
        return [c.data for c in out.chunks]

    def sum_i64(self, d):
        return self.mnr.kernels.reduce.sum(d, None, self.ctx)

    def sum_f64(self, d):
        return self.mnr.kernels.reduce.sum(d, None, self.ctx)
