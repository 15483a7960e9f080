//! `extern "C"` declarations of include/minarrow_b200.h (the subset the safe layer uses).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

#[repr(C)] pub struct mnr_ctx { _p: [u8; 0] }
#[repr(C)] pub struct mnr_buf { _p: [u8; 0] }
#[repr(C)] pub struct mnr_bits { _p: [u8; 0] }

#[repr(C)] #[derive(Clone, Copy)]
pub union mnr_scalar64 { pub i64_: i64, pub u64_: u64, pub f64_: f64 }

#[repr(C)] #[derive(Clone, Copy)]
pub struct mnr_agg { pub sum: mnr_scalar64, pub min: mnr_scalar64, pub max: mnr_scalar64, pub count: u64 }

/// mnr_dtype codes.
pub const MNR_I32: c_int = 0; pub const MNR_U32: c_int = 1; pub const MNR_I64: c_int = 2; pub const MNR_U64: c_int = 3;
pub const MNR_F32: c_int = 4; pub const MNR_F64: c_int = 5;
/// mnr_mask_mode.
pub const MNR_MASK_AND: c_int = 0; pub const MNR_MASK_OR: c_int = 1;

extern "C" {
    pub fn mnr_last_error() -> *const c_char;
    pub fn mnr_ctx_create(device: c_int, out: *mut *mut mnr_ctx) -> c_int;
    pub fn mnr_ctx_destroy(ctx: *mut mnr_ctx);
    pub fn mnr_ctx_synchronize(ctx: *mut mnr_ctx) -> c_int;

    pub fn mnr_buf_upload(ctx: *mut mnr_ctx, dtype: c_int, host: *const c_void, len: usize, out: *mut *mut mnr_buf) -> c_int;
    pub fn mnr_buf_download(ctx: *mut mnr_ctx, buf: *const mnr_buf, host: *mut c_void) -> c_int;
    pub fn mnr_buf_len(buf: *const mnr_buf) -> usize;
    pub fn mnr_buf_free(buf: *mut mnr_buf);
    pub fn mnr_bits_upload(ctx: *mut mnr_ctx, host: *const u8, len_bits: usize, out: *mut *mut mnr_bits) -> c_int;
    pub fn mnr_bits_download(ctx: *mut mnr_ctx, bits: *const mnr_bits, host: *mut u8) -> c_int;
    pub fn mnr_bits_len(bits: *const mnr_bits) -> usize;
    pub fn mnr_bits_free(bits: *mut mnr_bits);

    pub fn mnr_ew_binary(ctx: *mut mnr_ctx, op: c_int, lhs: *const mnr_buf, rhs: *const mnr_buf, lhs_mask: *const mnr_bits,
                         rhs_mask: *const mnr_bits, mode: c_int, out: *mut *mut mnr_buf, out_mask: *mut *mut mnr_bits) -> c_int;
    pub fn mnr_ew_scalar(ctx: *mut mnr_ctx, op: c_int, arr: *const mnr_buf, scalar: *const c_void, scalar_is_lhs: c_int,
                         mask: *const mnr_bits, out: *mut *mut mnr_buf, out_mask: *mut *mut mnr_bits) -> c_int;
    pub fn mnr_bits_binop(ctx: *mut mnr_ctx, op: c_int, lhs: *const mnr_bits, lhs_offset: usize, rhs: *const mnr_bits,
                          rhs_offset: usize, len: usize, out: *mut *mut mnr_bits) -> c_int;
    pub fn mnr_bits_not(ctx: *mut mnr_ctx, src: *const mnr_bits, offset: usize, len: usize, out: *mut *mut mnr_bits) -> c_int;
    pub fn mnr_bits_popcount(ctx: *mut mnr_ctx, mask: *const mnr_bits, offset: usize, len: usize, ones: *mut u64) -> c_int;
    pub fn mnr_bits_popcount_async(ctx: *mut mnr_ctx, mask: *const mnr_bits, offset: usize, len: usize, out_device: *mut c_void) -> c_int;
    pub fn mnr_reduce_stats(ctx: *mut mnr_ctx, buf: *const mnr_buf, validity: *const mnr_bits, out: *mut mnr_agg) -> c_int;
    pub fn mnr_reduce_stats_batch(ctx: *mut mnr_ctx, n: usize, bufs: *const *const mnr_buf, validities: *const *const mnr_bits,
                                  with_minmax: c_int, out: *mut mnr_agg) -> c_int;
    pub fn mnr_agg_mean(dtype: c_int, agg: *const mnr_agg) -> f64;

    // SuperArray chunks: consolidate / rechunk on the device, and shards fed straight from the reference's own Arrow C
    // stream export (`stream` = *mut minarrow::ffi::arrow_c_ffi::ArrowArrayStream, src/ffi/arrow_c_ffi.rs:153-168)
    pub fn mnr_concat(ctx: *mut mnr_ctx, n: usize, bufs: *const *const mnr_buf, validities: *const *const mnr_bits,
                      out: *mut *mut mnr_buf, out_validity: *mut *mut mnr_bits) -> c_int;
    pub fn mnr_bits_slice(ctx: *mut mnr_ctx, src: *const mnr_bits, offset: usize, len: usize, out: *mut *mut mnr_bits) -> c_int;
    pub fn mnr_arrow_stream_import(ctx: *mut mnr_ctx, stream: *mut c_void, chunk_lo: usize, chunk_hi: usize, capacity: usize,
                                   values: *mut *mut mnr_buf, validity: *mut *mut mnr_bits, n_imported: *mut usize,
                                   n_seen: *mut usize) -> c_int;

    // host-slice drop-ins: the reference leaf signatures (src/kernels/arithmetic/dispatch.rs:74-79,147-152)
    pub fn mnr_apply_int_i32(ctx: *mut mnr_ctx, lhs: *const i32, lhs_len: usize, rhs: *const i32, rhs_len: usize, op: c_int,
                             mask: *const u8, out: *mut i32, out_mask: *mut u8) -> c_int;
    pub fn mnr_apply_int_i64(ctx: *mut mnr_ctx, lhs: *const i64, lhs_len: usize, rhs: *const i64, rhs_len: usize, op: c_int,
                             mask: *const u8, out: *mut i64, out_mask: *mut u8) -> c_int;
    pub fn mnr_apply_float_f32(ctx: *mut mnr_ctx, lhs: *const f32, lhs_len: usize, rhs: *const f32, rhs_len: usize, op: c_int,
                               mask: *const u8, out: *mut f32, out_mask: *mut u8) -> c_int;
    pub fn mnr_apply_float_f64(ctx: *mut mnr_ctx, lhs: *const f64, lhs_len: usize, rhs: *const f64, rhs_len: usize, op: c_int,
                               mask: *const u8, out: *mut f64, out_mask: *mut u8) -> c_int;
    pub fn mnr_stats_host(ctx: *mut mnr_ctx, dtype: c_int, data: *const c_void, len: usize, validity: *const u8,
                          with_minmax: c_int, out: *mut mnr_agg) -> c_int;
}
