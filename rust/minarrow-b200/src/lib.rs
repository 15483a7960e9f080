//! Safe layer over libminarrow_b200.so: a device-resident buffer type alongside `Vec64`, a device validity-bitmask
//! type, Arrow-layout-preserving upload/download, and the reference's leaf functions with identical signatures.
//!
//! NOT COMPILED IN THIS REPO (no Rust toolchain in the build image) — see INTEGRATION.md.  Every call maps 1:1 onto an
//! entry point of include/minarrow_b200.h whose behaviour is pinned by the GPU parity tests.
pub mod ffi;

use core::ffi::{c_int, c_void};
use core::marker::PhantomData;
use std::ffi::CStr;

use minarrow::enums::error::KernelError;
use minarrow::enums::operators::ArithmeticOperator;
use minarrow::{Bitmask, FloatArray, IntegerArray, Vec64};

/// Element types the device path carries (IntegerArray<T> / FloatArray<T>).
pub trait B200Dtype: Copy + 'static { const CODE: c_int; }
impl B200Dtype for i32 { const CODE: c_int = ffi::MNR_I32; }
impl B200Dtype for u32 { const CODE: c_int = ffi::MNR_U32; }
impl B200Dtype for i64 { const CODE: c_int = ffi::MNR_I64; }
impl B200Dtype for u64 { const CODE: c_int = ffi::MNR_U64; }
impl B200Dtype for f32 { const CODE: c_int = ffi::MNR_F32; }
impl B200Dtype for f64 { const CODE: c_int = ffi::MNR_F64; }

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::mnr_last_error()) }.to_string_lossy().into_owned()
}

/// Status codes -1..-10 are the KernelError variants in declaration order (src/enums/error.rs:157-187).
fn check(rc: c_int) -> Result<(), KernelError> {
    match rc {
        0 => Ok(()),
        -2 => Err(KernelError::LengthMismatch(last_error())),
        -5 => Err(KernelError::UnsupportedType(last_error())),
        -9 => Err(KernelError::OutOfBounds(last_error())),
        -10 => panic!("{}", last_error()),     // the dense integer kernels' divide-by-zero panic (std.rs:54-55)
        _ => Err(KernelError::InvalidArguments(last_error())),
    }
}

/// One device + stream.  One context per host thread at a time (header "Conventions").
pub struct Context { h: *mut ffi::mnr_ctx }
unsafe impl Send for Context {}
impl Context {
    pub fn new(device: i32) -> Result<Self, KernelError> {
        let mut h = core::ptr::null_mut();
        check(unsafe { ffi::mnr_ctx_create(device as c_int, &mut h) })?;
        Ok(Self { h })
    }
}
impl Drop for Context { fn drop(&mut self) { unsafe { ffi::mnr_ctx_destroy(self.h) } } }

/// Device-resident values buffer: the `Vec64<T>` / `Buffer<T>` analogue in HBM (src/structs/buffer.rs:126-139).
pub struct DeviceBuffer<'c, T: B200Dtype> { ctx: &'c Context, h: *mut ffi::mnr_buf, _t: PhantomData<T> }
/// Device-resident bit-packed mask: the `Bitmask` analogue (src/structs/bitmask.rs:66-71).
pub struct DeviceBitmask<'c> { ctx: &'c Context, h: *mut ffi::mnr_bits }

impl<'c, T: B200Dtype> DeviceBuffer<'c, T> {
    /// Arrow-layout-preserving upload: the bytes of `data` one for one.
    pub fn upload(ctx: &'c Context, data: &[T]) -> Result<Self, KernelError> {
        let mut h = core::ptr::null_mut();
        check(unsafe { ffi::mnr_buf_upload(ctx.h, T::CODE, data.as_ptr() as *const c_void, data.len(), &mut h) })?;
        Ok(Self { ctx, h, _t: PhantomData })
    }
    pub fn len(&self) -> usize { unsafe { ffi::mnr_buf_len(self.h) } }
    /// Download into a 64-byte aligned `Vec64<T>` (taken by `Buffer::from` without a copy, buffer.rs:168-219).
    pub fn download(&self) -> Result<Vec64<T>, KernelError> {
        let n = self.len();
        let mut out = Vec64::<T>::with_capacity(n);
        unsafe { out.set_len(n) };
        check(unsafe { ffi::mnr_buf_download(self.ctx.h, self.h, out.as_mut_ptr() as *mut c_void) })?;
        Ok(out)
    }
    /// Fused null-aware `self op rhs` with both validity masks merged inside the kernel (one HBM pass).
    pub fn binary(&self, op: ArithmeticOperator, rhs: &Self, lhs_mask: Option<&DeviceBitmask<'c>>,
                  rhs_mask: Option<&DeviceBitmask<'c>>, or_union: bool)
                  -> Result<(Self, Option<DeviceBitmask<'c>>), KernelError> {
        let (mut ob, mut om) = (core::ptr::null_mut(), core::ptr::null_mut());
        check(unsafe {
            ffi::mnr_ew_binary(self.ctx.h, op as c_int, self.h, rhs.h,
                               lhs_mask.map_or(core::ptr::null(), |m| m.h as *const _),
                               rhs_mask.map_or(core::ptr::null(), |m| m.h as *const _),
                               if or_union { ffi::MNR_MASK_OR } else { ffi::MNR_MASK_AND }, &mut ob, &mut om)
        })?;
        let mask = if om.is_null() { None } else { Some(DeviceBitmask { ctx: self.ctx, h: om }) };
        Ok((Self { ctx: self.ctx, h: ob, _t: PhantomData }, mask))
    }
    /// Null-aware {sum, min, max, count} in one pass.
    pub fn stats(&self, validity: Option<&DeviceBitmask<'c>>) -> Result<ffi::mnr_agg, KernelError> {
        let mut agg = core::mem::MaybeUninit::<ffi::mnr_agg>::uninit();
        check(unsafe { ffi::mnr_reduce_stats(self.ctx.h, self.h, validity.map_or(core::ptr::null(), |m| m.h as *const _), agg.as_mut_ptr()) })?;
        Ok(unsafe { agg.assume_init() })
    }
}
impl<'c, T: B200Dtype> Drop for DeviceBuffer<'c, T> { fn drop(&mut self) { unsafe { ffi::mnr_buf_free(self.h) } } }

impl<'c> DeviceBitmask<'c> {
    pub fn upload(ctx: &'c Context, mask: &Bitmask) -> Result<Self, KernelError> {
        let mut h = core::ptr::null_mut();
        check(unsafe { ffi::mnr_bits_upload(ctx.h, mask.bits.as_ptr(), mask.len(), &mut h) })?;
        Ok(Self { ctx, h })
    }
    pub fn len(&self) -> usize { unsafe { ffi::mnr_bits_len(self.h) } }
    pub fn count_ones(&self) -> Result<u64, KernelError> {
        let mut ones = 0u64;
        check(unsafe { ffi::mnr_bits_popcount(self.ctx.h, self.h, 0, self.len(), &mut ones) })?;
        Ok(ones)
    }
}
impl<'c> Drop for DeviceBitmask<'c> { fn drop(&mut self) { unsafe { ffi::mnr_bits_free(self.h) } } }

// Drop-ins for the reference's leaf functions (src/kernels/arithmetic/dispatch.rs:74-131,147-204,376-402): same names,
// same signatures (plus the context), same results — `Some(mask)` iff a mask was passed; masked zero divisors become
// nulls; dense integer zero divisors panic like the reference's kernels do.
macro_rules! impl_apply {
    ($name:ident, $ffi:ident, $t:ty, $arr:ident) => {
        pub fn $name(ctx: &Context, lhs: &[$t], rhs: &[$t], op: ArithmeticOperator, mask: Option<&Bitmask>)
                     -> Result<$arr<$t>, KernelError> {
            let len = lhs.len();
            let mut out = Vec64::<$t>::with_capacity(len);
            unsafe { out.set_len(len) };
            let mut out_mask = mask.map(|_| Bitmask::new_set_all(len, false));
            check(unsafe {
                ffi::$ffi(ctx.h, lhs.as_ptr(), len, rhs.as_ptr(), rhs.len(), op as c_int,
                          mask.map_or(core::ptr::null(), |m| m.bits.as_ptr()), out.as_mut_ptr(),
                          out_mask.as_mut().map_or(core::ptr::null_mut(), |m| m.bits.as_mut_ptr()))
            })?;
            Ok($arr { data: out.into(), null_mask: out_mask })
        }
    };
}
impl_apply!(apply_int_i32, mnr_apply_int_i32, i32, IntegerArray);
impl_apply!(apply_int_u32, mnr_apply_int_u32, u32, IntegerArray);
impl_apply!(apply_int_i64, mnr_apply_int_i64, i64, IntegerArray);
impl_apply!(apply_int_u64, mnr_apply_int_u64, u64, IntegerArray);
impl_apply!(apply_float_f32, mnr_apply_float_f32, f32, FloatArray);
impl_apply!(apply_float_f64, mnr_apply_float_f64, f64, FloatArray);

/// Drop-in for `apply_fma_f64` (dispatch.rs:221-290,411-418): `lhs.mul_add(rhs, acc)` with one rounding.
pub fn apply_fma_f64(ctx: &Context, lhs: &[f64], rhs: &[f64], acc: &[f64], mask: Option<&Bitmask>)
                     -> Result<FloatArray<f64>, KernelError> {
    let len = lhs.len();
    let mut out = Vec64::<f64>::with_capacity(len);
    unsafe { out.set_len(len) };
    let mut out_mask = mask.map(|_| Bitmask::new_set_all(len, false));
    check(unsafe {
        ffi::mnr_apply_fma_host(ctx.h, ffi::MNR_F64, lhs.as_ptr() as *const c_void, len, rhs.as_ptr() as *const c_void, rhs.len(),
                                acc.as_ptr() as *const c_void, acc.len(), mask.map_or(core::ptr::null(), |m| m.bits.as_ptr()),
                                out.as_mut_ptr() as *mut c_void, out_mask.as_mut().map_or(core::ptr::null_mut(), |m| m.bits.as_mut_ptr()))
    })?;
    Ok(FloatArray { data: out.into(), null_mask: out_mask })
}

impl<'c> DeviceBitmask<'c> {
    /// `and_masks` / `or_masks` / `xor_masks` over two device-resident masks (src/kernels/bitmask/dispatch.rs:96-131).
    pub fn binop(&self, op: minarrow::enums::operators::LogicalOperator, rhs: &Self) -> Result<Self, KernelError> {
        let mut h = core::ptr::null_mut();
        check(unsafe { ffi::mnr_bits_binop(self.ctx.h, op as c_int, self.h, 0, rhs.h, 0, self.len(), &mut h) })?;
        Ok(Self { ctx: self.ctx, h })
    }
    /// `not_mask` (dispatch.rs:133-145): trailing bits of the result stay zero.
    pub fn not(&self) -> Result<Self, KernelError> {
        let mut h = core::ptr::null_mut();
        check(unsafe { ffi::mnr_bits_not(self.ctx.h, self.h, 0, self.len(), &mut h) })?;
        Ok(Self { ctx: self.ctx, h })
    }
    pub fn download(&self) -> Result<Bitmask, KernelError> {
        let mut m = Bitmask::new_set_all(self.len(), false);
        check(unsafe { ffi::mnr_bits_download(self.ctx.h, self.h, m.bits.as_mut_ptr()) })?;
        Ok(m)
    }
}

impl<'c, T: B200Dtype> DeviceBuffer<'c, T> {
    /// `array op scalar` / `scalar op array` without materialising the length-1 operand
    /// (replaces broadcast_length_1_array + the binary kernel, src/kernels/routing/broadcast.rs:25-112).
    pub fn scalar(&self, op: ArithmeticOperator, scalar: T, scalar_is_lhs: bool, mask: Option<&DeviceBitmask<'c>>)
                  -> Result<(Self, Option<DeviceBitmask<'c>>), KernelError> {
        let (mut ob, mut om) = (core::ptr::null_mut(), core::ptr::null_mut());
        check(unsafe {
            ffi::mnr_ew_scalar(self.ctx.h, op as c_int, self.h, &scalar as *const T as *const c_void, scalar_is_lhs as c_int,
                               mask.map_or(core::ptr::null(), |m| m.h as *const _), &mut ob, &mut om)
        })?;
        let mask = if om.is_null() { None } else { Some(DeviceBitmask { ctx: self.ctx, h: om }) };
        Ok((Self { ctx: self.ctx, h: ob, _t: PhantomData }, mask))
    }
}

/// Null-aware {sum, min, max, count} of a SuperArray's chunks resident on one GPU: one batched launch per dtype class,
/// partials folded in chunk order (the device form of `par_chunks(1 << 20).map(sum).sum()`,
/// benches/benchmark_parallel_simd.rs:81-97).  Per-GPU results are combined by the caller's all-reduce.
pub fn chunk_stats<'c, T: B200Dtype>(ctx: &'c Context, chunks: &[(&DeviceBuffer<'c, T>, Option<&DeviceBitmask<'c>>)])
                                     -> Result<ffi::mnr_agg, KernelError> {
    let bufs: Vec<*const ffi::mnr_buf> = chunks.iter().map(|(b, _)| b.h as *const _).collect();
    let masks: Vec<*const ffi::mnr_bits> = chunks.iter().map(|(_, m)| m.map_or(core::ptr::null(), |m| m.h as *const _)).collect();
    let mut parts = vec![core::mem::MaybeUninit::<ffi::mnr_agg>::uninit(); chunks.len()];
    check(unsafe { ffi::mnr_reduce_stats_batch(ctx.h, chunks.len(), bufs.as_ptr(), masks.as_ptr(), 1, parts.as_mut_ptr() as *mut ffi::mnr_agg) })?;
    let mut out = core::mem::MaybeUninit::<ffi::mnr_agg>::uninit();
    check(unsafe { ffi::mnr_agg_combine(T::CODE, parts.as_ptr() as *const ffi::mnr_agg, chunks.len(), out.as_mut_ptr()) })?;
    Ok(unsafe { out.assume_init() })
}

// ---- SuperArray / SuperTable over all GPUs of the box, from one process (include/minarrow_b200.h "sharding") ----------------

/// All GPUs of the box: one context + NVLink mailbox per device (`mnr_group`).  Chunk `i` of `n` lives on rank
/// `floor(i * G / n)` (`mnr_shard_owner`) — contiguous blocks, so the SuperArray re-assembles with the same chunk
/// boundaries (src/structs/chunked/super_array.rs:96-103).
pub struct Group { h: *mut ffi::mnr_group, world: usize }
unsafe impl Send for Group {}

/// One chunk resident on its owning GPU (values + optional validity).  Freed with the group's per-rank context.
pub struct ShardedChunk { buf: *mut ffi::mnr_buf, validity: *mut ffi::mnr_bits }
impl Drop for ShardedChunk {
    fn drop(&mut self) { unsafe { ffi::mnr_buf_free(self.buf); ffi::mnr_bits_free(self.validity) } }
}

impl Group {
    /// `devices = None`: devices 0..world-1.
    pub fn new(world: usize, devices: Option<&[i32]>) -> Result<Self, KernelError> {
        let mut h = core::ptr::null_mut();
        let dv: Option<Vec<c_int>> = devices.map(|d| d.iter().map(|&x| x as c_int).collect());
        check(unsafe { ffi::mnr_group_create(world as c_int, dv.as_ref().map_or(core::ptr::null(), |v| v.as_ptr()), &mut h) })?;
        Ok(Self { h, world })
    }
    pub fn world(&self) -> usize { self.world }
    pub fn owner(&self, chunk: usize, n_chunks: usize) -> usize { unsafe { ffi::mnr_shard_owner(chunk, n_chunks, self.world as c_int) as usize } }

    /// `SuperArray -> shards`: chunk i is uploaded to its owner, all PCIe links copying at once.
    pub fn upload<T: B200Dtype>(&self, chunks: &[(&[T], Option<&Bitmask>)]) -> Result<Vec<ShardedChunk>, KernelError> {
        let n = chunks.len();
        let ptrs: Vec<*const c_void> = chunks.iter().map(|(d, _)| d.as_ptr() as *const c_void).collect();
        let lens: Vec<usize> = chunks.iter().map(|(d, _)| d.len()).collect();
        let masks: Vec<*const u8> = chunks.iter().map(|(_, m)| m.map_or(core::ptr::null(), |m| m.bits.as_ptr())).collect();
        let mut bufs = vec![core::ptr::null_mut(); n];
        let mut vals = vec![core::ptr::null_mut(); n];
        check(unsafe { ffi::mnr_group_upload(self.h, T::CODE, n, ptrs.as_ptr(), lens.as_ptr(), masks.as_ptr(), bufs.as_mut_ptr(), vals.as_mut_ptr()) })?;
        Ok(bufs.into_iter().zip(vals).map(|(buf, validity)| ShardedChunk { buf, validity }).collect())
    }

    /// `lhs[i] op rhs[i]` on the GPU that owns chunk pair i (route_super_array_broadcast, src/kernels/broadcast/super_array.rs:180-249:
    /// validity = union of the two chunks' masks, fused into the kernel); one batched launch per device, no communication.
    pub fn ew_binary(&self, op: ArithmeticOperator, lhs: &[ShardedChunk], rhs: &[ShardedChunk]) -> Result<Vec<ShardedChunk>, KernelError> {
        let n = lhs.len();
        let l: Vec<*const ffi::mnr_buf> = lhs.iter().map(|c| c.buf as *const _).collect();
        let r: Vec<*const ffi::mnr_buf> = rhs.iter().map(|c| c.buf as *const _).collect();
        let lm: Vec<*const ffi::mnr_bits> = lhs.iter().map(|c| c.validity as *const _).collect();
        let rm: Vec<*const ffi::mnr_bits> = rhs.iter().map(|c| c.validity as *const _).collect();
        let mut ob = vec![core::ptr::null_mut(); n];
        let mut om = vec![core::ptr::null_mut(); n];
        check(unsafe { ffi::mnr_group_ew_binary(self.h, op as c_int, n, l.as_ptr(), r.as_ptr(), lm.as_ptr(), rm.as_ptr(), ffi::MNR_MASK_OR,
                                                ob.as_mut_ptr(), om.as_mut_ptr()) })?;
        Ok(ob.into_iter().zip(om).map(|(buf, validity)| ShardedChunk { buf, validity }).collect())
    }

    /// Per-column {sum, min, max, count} of a sharded SuperTable (`columns[c]` = column c's chunks across the batches): batched
    /// kernels on every GPU, per-column fold + NVLink mailbox exchange in the kernel that finishes last, rank-order combine —
    /// one call, no NCCL, identical bits on every rank (benches/benchmark_parallel_simd.rs:81-97 over super_table.rs:38-73).
    pub fn column_stats(&self, columns: &[(c_int, &[ShardedChunk])]) -> Result<Vec<ffi::mnr_agg>, KernelError> {
        let mut bufs = Vec::new();
        let mut vals = Vec::new();
        let mut col = Vec::new();
        for (c, (_, chunks)) in columns.iter().enumerate() {
            for ch in chunks.iter() { bufs.push(ch.buf as *const ffi::mnr_buf); vals.push(ch.validity as *const ffi::mnr_bits); col.push(c as u32); }
        }
        let dts: Vec<c_int> = columns.iter().map(|(dt, _)| *dt).collect();
        let mut out = vec![core::mem::MaybeUninit::<ffi::mnr_agg>::uninit(); columns.len()];
        check(unsafe { ffi::mnr_group_reduce_stats(self.h, bufs.len(), bufs.as_ptr(), vals.as_ptr(), 1, columns.len(), col.as_ptr(), dts.as_ptr(),
                                                   out.as_mut_ptr() as *mut ffi::mnr_agg) })?;
        Ok(out.into_iter().map(|a| unsafe { a.assume_init() }).collect())
    }
}
impl Drop for Group { fn drop(&mut self) { unsafe { ffi::mnr_group_destroy(self.h) } } }
