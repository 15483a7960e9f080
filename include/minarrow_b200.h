/*
 * minarrow_b200.h — C ABI of the B200-native drop-in for Minarrow's columnar compute hot path.
 *
 * Every entry point names the reference interface it replaces (paths relative to pbower/minarrow
 * v0.10.1).  The seam in the reference is the set of free functions in
 * src/kernels/arithmetic/dispatch.rs and src/kernels/bitmask/dispatch.rs, called from the router
 * src/kernels/routing/arithmetic.rs:278-339 and directly by users; INTEGRATION.md shows the Rust
 * `extern "C"` block + build.rs lines a maintainer adds to bind them.
 *
 * Conventions
 *  - Plain C: pointers and sizes only; no CUDA or torch types.  A CUDA stream crosses as `void*`.
 *  - Return 0 on success, a negative mnr_status otherwise; mnr_last_error() gives the message
 *    (thread-local, like ArrowArrayStream::get_last_error, src/ffi/arrow_c_ffi.rs:160-168).
 *    Codes -1..-10 are the KernelError variants in declaration order (src/enums/error.rs:157-187).
 *  - Layout is Arrow's, byte for byte: values = contiguous T[len]; validity /
 *    boolean data = ceil(len/8) bytes, bit i = byte i>>3 bit i&7 (LSB first), 1 = valid, bits >= len
 *    zero (Bitmask::mask_trailing_bits, src/structs/bitmask.rs:83-90).
 *  - Inputs are borrowed and never written; outputs are fresh and owned by the caller
 *    (dispatch.rs:88-97).  An output validity mask exists iff an input mask was passed
 *    (dispatch.rs:90-104).
 *  - A context owns one CUDA stream.  Device-resident calls are asynchronous on that stream unless
 *    stated; anything returning a host scalar synchronises.  One context per host thread at a time.
 *  - There is no CPU fallback: without a usable sm_100 device every call fails with MNR_ERR_NO_DEVICE.
 */
#ifndef MINARROW_B200_H
#define MINARROW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MNR_ABI_VERSION 1

typedef struct mnr_ctx mnr_ctx;   /* device + stream + scratch                                          */
typedef struct mnr_buf mnr_buf;   /* device-resident values buffer: the Vec64<T>/Buffer<T> analogue      */
typedef struct mnr_bits mnr_bits; /* device-resident bit-packed mask: the Bitmask analogue               */
typedef struct mnr_xchg mnr_xchg; /* cross-GPU mailbox set for the fused reduction + exchange kernel      */
typedef struct mnr_group mnr_group; /* the GPUs of one box driven from one process: one ctx + mailbox per device */

/* Element types of IntegerArray<T> / FloatArray<T> (src/structs/variants/{integer,float}.rs);
 * 8/16-bit integers are the reference's `extended_numeric_types` feature (dispatch.rs:380-387). */
typedef enum {
    MNR_I32 = 0, MNR_U32 = 1, MNR_I64 = 2, MNR_U64 = 3, MNR_F32 = 4, MNR_F64 = 5,
    MNR_I8 = 6, MNR_U8 = 7, MNR_I16 = 8, MNR_U16 = 9
} mnr_dtype;

/* ArithmeticOperator, declaration order (src/enums/operators.rs:19-48). */
typedef enum {
    MNR_ADD = 0, MNR_SUB = 1, MNR_MUL = 2, MNR_DIV = 3, MNR_REM = 4, MNR_POW = 5, MNR_FLOORDIV = 6
} mnr_op;

/* LogicalOperator (src/enums/operators.rs:88-104). */
typedef enum { MNR_AND = 0, MNR_OR = 1, MNR_XOR = 2 } mnr_logical_op;

/* How two input validity masks combine when both are given to a fused element-wise call.
 * AND = merge_bitmasks_to_new / Bitmask::intersect (src/kernels/bitmask/mod.rs:171-197; used by
 *       impl_apply_datetime!, dispatch.rs:321-322) — valid iff both valid.
 * OR  = Bitmask::union as used by route_super_array_broadcast
 *       (src/kernels/broadcast/super_array.rs:214-229; src/structs/bitmask.rs:661-669).
 * With one mask only, that mask is used as is in either mode (super_array.rs:218-220). */
typedef enum { MNR_MASK_AND = 0, MNR_MASK_OR = 1 } mnr_mask_mode;

typedef enum {
    MNR_OK = 0,
    MNR_ERR_TYPE_MISMATCH = -1,
    MNR_ERR_LENGTH_MISMATCH = -2,     /* confirm_equal_len, src/utils.rs:163-171                         */
    MNR_ERR_BROADCASTING = -3,
    MNR_ERR_OPERATOR_MISMATCH = -4,
    MNR_ERR_UNSUPPORTED_TYPE = -5,    /* routing/arithmetic.rs:403                                       */
    MNR_ERR_COLUMN_NOT_FOUND = -6,
    MNR_ERR_INVALID_ARGUMENTS = -7,
    MNR_ERR_PLAN = -8,
    MNR_ERR_OUT_OF_BOUNDS = -9,
    MNR_ERR_DIVIDE_BY_ZERO = -10,     /* the dense integer kernels' panic, std.rs:54-55,61-62,69-70      */
    MNR_ERR_CUDA = -100,
    MNR_ERR_NO_DEVICE = -101,
    MNR_ERR_OUT_OF_MEMORY = -102
} mnr_status;

/* 8-byte scalar whose active member follows the column dtype's accumulator type:
 * I8..I64 -> i64, U8..U64 -> u64, F32/F64 -> f64. */
typedef union { int64_t i64; uint64_t u64; double f64; } mnr_scalar64;

/* Null-aware aggregate of one column (32 bytes; also the per-GPU partial that is all-reduced). */
typedef struct {
    mnr_scalar64 sum;   /* wrapping two's complement for integers; f64 for floats                        */
    mnr_scalar64 min;   /* identity when nothing qualifies: INT_MAX-of-type / UINT_MAX-of-type / NaN     */
    mnr_scalar64 max;   /* identity: INT_MIN-of-type / 0 / NaN                                           */
    uint64_t count;     /* number of valid rows (popcount of validity, or len)                           */
} mnr_agg;

/* ---- library / context ----------------------------------------------------------------------------- */
int mnr_abi_version(void);
const char* mnr_last_error(void);
int mnr_device_count(void);
int mnr_ctx_create(int device, mnr_ctx** out);
/* Borrow the caller's stream (e.g. torch.cuda.current_stream().cuda_stream); `cuda_stream` = cudaStream_t. */
int mnr_ctx_create_on_stream(int device, void* cuda_stream, mnr_ctx** out);
void mnr_ctx_destroy(mnr_ctx* ctx);
int mnr_ctx_synchronize(mnr_ctx* ctx);
int mnr_ctx_device(const mnr_ctx* ctx);
void* mnr_ctx_stream(const mnr_ctx* ctx);
/* Number of kernels of this library launched through `ctx` so far. */
uint64_t mnr_ctx_launch_count(const mnr_ctx* ctx);
/* Per-context options; contexts are independent (a knob set on one never changes another's launches).
 *   "ew_grid_cap", "ew_max_tier", "ew_sdiv64_cfg", "ew_fdiv_cfg", "ew_heavy_cfg", "ew_cheap8_cfg": launch geometry of the element-wise
 *       kernels for tuning sweeps — results never depend on them;
 *   "host_chunk_rows": rows per staging chunk of the host-slice drop-ins (multiple of 1024; default 4 Mi).  A float sum
 *       through mnr_stats_host folds one partial per staging chunk in chunk order, so its bits depend on this value
 *       (integer results never do);
 *   "reduce_overlap" (0/1, default 0): consecutive mnr_reduce_stats_exchange calls on this context may overlap — the
 *       next reduction streams its column while the previous one finishes its cross-GPU exchange.  Opt-in because the
 *       caller vouches that the column being reduced is at rest (not written by work still in flight on the stream other
 *       than this library's own reductions).
 * Unknown keys return MNR_ERR_INVALID_ARGUMENTS. */
int mnr_ctx_set_option(mnr_ctx* ctx, const char* key, int64_t value);

/* ---- device-resident buffers: Vec64<T> / Buffer<T> (src/structs/buffer.rs:126-139) -------------------- */
int mnr_buf_alloc(mnr_ctx* ctx, mnr_dtype dtype, size_t len, mnr_buf** out);
/* Arrow-layout-preserving upload of `len` elements (Array::data_ptr_and_byte_len, src/enums/array.rs:2563). */
int mnr_buf_upload(mnr_ctx* ctx, mnr_dtype dtype, const void* host, size_t len, mnr_buf** out);
/* Same without the final synchronise: `host` (ideally pinned) must stay untouched until mnr_ctx_synchronize. */
int mnr_buf_upload_async(mnr_ctx* ctx, mnr_dtype dtype, const void* host, size_t len, mnr_buf** out);
/* Non-owning view of caller-owned device memory. */
int mnr_buf_wrap(mnr_ctx* ctx, mnr_dtype dtype, void* device_ptr, size_t len, mnr_buf** out);
/* ArrayV window `(array, offset, len)` (src/structs/views/array_view.rs:79-94): non-owning, parent must outlive it. */
int mnr_buf_slice(const mnr_buf* parent, size_t offset, size_t len, mnr_buf** out);
int mnr_buf_download(mnr_ctx* ctx, const mnr_buf* buf, void* host);   /* synchronises */
size_t mnr_buf_len(const mnr_buf* buf);
int mnr_buf_dtype(const mnr_buf* buf);
void* mnr_buf_device_ptr(const mnr_buf* buf);
void mnr_buf_free(mnr_buf* buf);

/* ---- device-resident bitmasks: Bitmask (src/structs/bitmask.rs:66-71) --------------------------------- */
int mnr_bits_alloc(mnr_ctx* ctx, size_t len_bits, mnr_bits** out);
/* Bitmask::new_set_all (bitmask.rs:94-105). */
int mnr_bits_new_set_all(mnr_ctx* ctx, size_t len_bits, int value, mnr_bits** out);
/* Upload ceil(len_bits/8) bytes (Array::null_mask_ptr_and_byte_len, src/enums/array.rs:2672); slack bits are cleared. */
int mnr_bits_upload(mnr_ctx* ctx, const uint8_t* host_bytes, size_t len_bits, mnr_bits** out);
int mnr_bits_upload_async(mnr_ctx* ctx, const uint8_t* host_bytes, size_t len_bits, mnr_bits** out);
int mnr_bits_wrap(mnr_ctx* ctx, void* device_ptr, size_t len_bits, mnr_bits** out);
int mnr_bits_download(mnr_ctx* ctx, const mnr_bits* bits, uint8_t* host_bytes);   /* synchronises */
size_t mnr_bits_len(const mnr_bits* bits);
void* mnr_bits_device_ptr(const mnr_bits* bits);
void mnr_bits_free(mnr_bits* bits);

/* ---- element-wise arithmetic, device-resident ----------------------------------------------------------
 * apply_int_{i32,u32,i64,u64,..} / apply_float_{f32,f64} (src/kernels/arithmetic/dispatch.rs:65-206,376-402)
 * with the caller-side mask merge fused in.  One HBM pass: out[i] = valid ? lhs[i] op rhs[i] : 0.
 *  - no mask: dense kernel, *out_mask = NULL; integer Div/Rem/FloorDiv with a zero divisor returns
 *    MNR_ERR_DIVIDE_BY_ZERO (the reference panics; output contents unspecified) — this one case synchronises.
 *  - mask(s): invalid row => value 0 + validity 0; integer Div/Rem/FloorDiv by zero => value 0 + validity 0
 *    (std.rs:96-136, simd.rs:268-328); float results keep the input validity (NaN/Inf stay valid).
 *  - integers wrap; MIN / -1 = MIN, MIN % -1 = 0 (assumption A.2); floats are IEEE with no FMA
 *    contraction; Power = exp(b * ln a) (tolerance parity only).
 * `mode` matters only when both masks are given.  lhs/rhs must share dtype and length. */
int mnr_ew_binary(mnr_ctx* ctx, mnr_op op, const mnr_buf* lhs, const mnr_buf* rhs, const mnr_bits* lhs_mask,
                  const mnr_bits* rhs_mask, mnr_mask_mode mode, mnr_buf** out, mnr_bits** out_mask);
/* Same, into caller-provided outputs (out_mask required iff a mask is given).  For every *_into entry point: an
 * output (values or mask) must not overlap any input of the same call — the kernels read inputs through the read-only
 * path — and an overlapping call is refused with MNR_ERR_INVALID_ARGUMENTS.  In-place `x = x op y` therefore takes a
 * fresh output (the allocating forms draw from a stream-ordered pool: no cudaMalloc per call). */
int mnr_ew_binary_into(mnr_ctx* ctx, mnr_op op, const mnr_buf* lhs, const mnr_buf* rhs, const mnr_bits* lhs_mask,
                       const mnr_bits* rhs_mask, mnr_mask_mode mode, mnr_buf* out, mnr_bits* out_mask);
/* Scalar broadcast without materialising the length-1 operand (replaces broadcast_length_1_array +
 * the binary kernel, src/kernels/routing/broadcast.rs:25-112).  `scalar` points at one host element of the
 * array's dtype; scalar_is_lhs selects `scalar op arr[i]` vs `arr[i] op scalar` (operand order is
 * significant: src/kernels/broadcast/scalar.rs:1332-1350). */
int mnr_ew_scalar(mnr_ctx* ctx, mnr_op op, const mnr_buf* arr, const void* scalar, int scalar_is_lhs,
                  const mnr_bits* mask, mnr_buf** out, mnr_bits** out_mask);
int mnr_ew_scalar_into(mnr_ctx* ctx, mnr_op op, const mnr_buf* arr, const void* scalar, int scalar_is_lhs,
                       const mnr_bits* mask, mnr_buf* out, mnr_bits* out_mask);
/* apply_fma_{f32,f64} (dispatch.rs:211-290,404-418): out = fma(a, b, c), single rounding. */
int mnr_ew_fma(mnr_ctx* ctx, const mnr_buf* a, const mnr_buf* b, const mnr_buf* c, const mnr_bits* mask,
               mnr_buf** out, mnr_bits** out_mask);
int mnr_ew_fma_into(mnr_ctx* ctx, const mnr_buf* a, const mnr_buf* b, const mnr_buf* c, const mnr_bits* mask,
                    mnr_buf* out, mnr_bits* out_mask);
/* Mixed-type promotion of the router (routing/arithmetic.rs:244-269,342-373): an I32 operand against an
 * F64 / F32 operand is cast on load (`as f64` / `as f32`) inside the same pass; output has the float dtype. */
int mnr_ew_binary_promote(mnr_ctx* ctx, mnr_op op, const mnr_buf* lhs, const mnr_buf* rhs, const mnr_bits* lhs_mask,
                          const mnr_bits* rhs_mask, mnr_mask_mode mode, mnr_buf** out, mnr_bits** out_mask);

/* Batched fan-out of the container routes: route_super_array_broadcast walks chunk pairs one leaf call at a time
 * (src/kernels/broadcast/super_array.rs:180-249, "TODO: Parallelise" :193; Table / SuperTable routes table.rs:31-62,
 * super_table.rs:38-73).  These take the whole chunk list: per-chunk length check as in super_array.rs:203-213, then
 * ONE launch per (dtype, alignment, masked) class.  Chunk i's result is bit-identical to mnr_ew_binary(lhs[i], rhs[i]).
 * Mask arrays may be NULL or hold NULL entries.  `scalars[i]` points at one host element of arrs[i]'s dtype. */
int mnr_ew_binary_batch(mnr_ctx* ctx, mnr_op op, size_t n, const mnr_buf* const* lhs, const mnr_buf* const* rhs,
                        const mnr_bits* const* lhs_mask, const mnr_bits* const* rhs_mask, mnr_mask_mode mode,
                        mnr_buf** out, mnr_bits** out_mask);
int mnr_ew_binary_batch_into(mnr_ctx* ctx, mnr_op op, size_t n, const mnr_buf* const* lhs, const mnr_buf* const* rhs,
                             const mnr_bits* const* lhs_mask, const mnr_bits* const* rhs_mask, mnr_mask_mode mode,
                             mnr_buf* const* out, mnr_bits* const* out_mask);
int mnr_ew_scalar_batch_into(mnr_ctx* ctx, mnr_op op, size_t n, const mnr_buf* const* arrs, const void* const* scalars,
                             int scalar_is_lhs, const mnr_bits* const* masks, mnr_buf* const* out,
                             mnr_bits* const* out_mask);

/* ---- bitmask kernels, device-resident (src/kernels/bitmask/dispatch.rs) ---------------------------------
 * Windows are BitmaskVT = (&Bitmask, offset, len).  Like the reference (bitmask_window_bytes,
 * src/kernels/bitmask/mod.rs:124-128) binop/not start at BYTE offset/8: sub-byte offsets are floored. */
/* and_masks / or_masks / xor_masks (dispatch.rs:96-131) -> bitmask_binop_simd (simd.rs:95-139). */
int mnr_bits_binop(mnr_ctx* ctx, mnr_logical_op op, const mnr_bits* lhs, size_t lhs_offset, const mnr_bits* rhs,
                   size_t rhs_offset, size_t len, mnr_bits** out);
int mnr_bits_binop_into(mnr_ctx* ctx, mnr_logical_op op, const mnr_bits* lhs, size_t lhs_offset,
                        const mnr_bits* rhs, size_t rhs_offset, size_t len, mnr_bits* out);
/* not_mask (dispatch.rs:135-144) -> bitmask_unop_simd (simd.rs:169-203); BooleanArray `!` (boolean.rs:853-866). */
int mnr_bits_not(mnr_ctx* ctx, const mnr_bits* src, size_t offset, size_t len, mnr_bits** out);
int mnr_bits_not_into(mnr_ctx* ctx, const mnr_bits* src, size_t offset, size_t len, mnr_bits* out);
/* popcount_mask (dispatch.rs:258-267 -> simd.rs:596-645): set bits of words [offset/64 ..) over `len` bits.
 * Bitmask::count_ones / null_count (bitmask.rs:393-417): offset 0, len = mask len; null_count = len - ones.
 * Blocks until the count has arrived in host memory (polled out of mapped pinned memory; work queued on the stream before
 * the call has completed by then, the stream itself may still be retiring the kernel). */
int mnr_bits_popcount(mnr_ctx* ctx, const mnr_bits* mask, size_t offset, size_t len, uint64_t* ones);
/* Asynchronous form for device-resident pipelines: the count is written to DEVICE memory `out_device` (one uint64,
 * 8-byte aligned) on the context stream; no synchronisation (null_count feeding a later kernel never visits the host). */
int mnr_bits_popcount_async(mnr_ctx* ctx, const mnr_bits* mask, size_t offset, size_t len, void* out_device);
/* all_true_mask / all_false_mask (dispatch.rs:273-295). Synchronise. */
int mnr_bits_all_true(mnr_ctx* ctx, const mnr_bits* mask, int* out);
int mnr_bits_all_false(mnr_ctx* ctx, const mnr_bits* mask, int* out);
/* Validity merge as a stand-alone mask (normally fused into mnr_ew_*): AND = merge_bitmasks_to_new
 * (bitmask/mod.rs:171-197), OR = Bitmask::union_opt (bitmask.rs:651-669).  Either side may be NULL
 * ("no nulls"); both NULL => *out = NULL. */
int mnr_bits_merge(mnr_ctx* ctx, const mnr_bits* lhs, const mnr_bits* rhs, size_t len, mnr_mask_mode mode,
                   mnr_bits** out);
/* eq_mask / ne_mask (dispatch.rs:178-200 -> simd.rs:402-472): offsets must be multiples of 64
 * (the reference panics otherwise => MNR_ERR_INVALID_ARGUMENTS). */
int mnr_bits_eq(mnr_ctx* ctx, const mnr_bits* a, size_t a_offset, const mnr_bits* b, size_t b_offset, size_t len,
                int negate, mnr_bits** out);
/* all_eq / all_ne (dispatch.rs:204-226 -> simd.rs:490-581); all_ne is !all_eq as in the reference. Synchronise. */
int mnr_bits_all_eq(mnr_ctx* ctx, const mnr_bits* a, size_t a_offset, const mnr_bits* b, size_t b_offset,
                    size_t len, int* out);
/* in_mask / not_in_mask (dispatch.rs:150-172 -> simd.rs:327-398). */
int mnr_bits_in(mnr_ctx* ctx, const mnr_bits* lhs, size_t lhs_offset, const mnr_bits* rhs, size_t rhs_offset,
                size_t len, int negate, mnr_bits** out);

/* Bitmask::slice_clone (src/structs/bitmask.rs:604-626): bits [offset, offset + len) as a fresh mask from bit 0. */
int mnr_bits_slice(mnr_ctx* ctx, const mnr_bits* src, size_t offset, size_t len, mnr_bits** out);
/* Device consolidate: SuperArray::consolidate / Array::concat = append_array over all chunks in row order
 * (src/traits/consolidate.rs:61-69, src/macros.rs:311-341; the building block of rechunk, super_array.rs:674-787).
 * Values are appended; validity is gathered at bit granularity; a chunk without a mask counts as all valid and
 * *out_validity is NULL iff no chunk has a mask.  All chunks must share a dtype (else MNR_ERR_TYPE_MISMATCH).
 * Re-splitting is zero-copy for values (mnr_buf_slice) and mnr_bits_slice for validity. */
int mnr_concat(mnr_ctx* ctx, size_t n, const mnr_buf* const* bufs, const mnr_bits* const* validities, mnr_buf** out,
               mnr_bits** out_validity);
/* simd_eq_mask_u8 / _u16 / _u32 / _u64 (src/kernels/bitmask/simd.rs:741-788): typed compare -> bitmask,
 * bit i = ((data[i] & *field_mask) == *target); `field_mask` / `target` point at one host element of data's dtype
 * (any integer dtype: the compare is on the raw lanes). */
int mnr_eq_mask(mnr_ctx* ctx, const mnr_buf* data, const void* field_mask, const void* target, mnr_bits** out);

/* ---- reductions, device-resident -------------------------------------------------------------------------
 * Sum follows the benches that define it in the reference (benches/benchmark_parallel_simd.rs:44-97,
 * benches/hotloop_benchmark_simd.rs:56-174): integer sums wrap and are bit-exact in any order; float sums
 * use the fixed order documented in DESIGN.md (<= 1e-12 relative to the reference order).  count / min /
 * max / mean and null-skipping are not in the reference tree (downstream `simd-kernels` crate); their
 * definition is DESIGN.md "A.6".  `validity` may be NULL (dense).  I32/U32 widen to 64-bit sums, F32 sums in f64. */
int mnr_reduce_stats(mnr_ctx* ctx, const mnr_buf* buf, const mnr_bits* validity, mnr_agg* out_host);  /* blocks until the aggregate
    has arrived in host memory (polled, sequence-framed stores into mapped pinned memory; no stream synchronisation) */
int mnr_reduce_sum(mnr_ctx* ctx, const mnr_buf* buf, const mnr_bits* validity, mnr_scalar64* out_sum,
                   uint64_t* out_count);                                                             /* syncs */
/* Asynchronous forms: the 32-byte mnr_agg is written to DEVICE memory `out_device` (16-byte aligned) on the
 * context stream — the per-GPU partial handed to the NCCL all-reduce.  with_minmax = 0 leaves min/max at
 * their identities and runs the cheaper sum+count kernel. */
int mnr_reduce_stats_async(mnr_ctx* ctx, const mnr_buf* buf, const mnr_bits* validity, int with_minmax,
                           void* out_device);
/* Batched fan-out: the SuperArray / Table / SuperTable routes call the leaf once per chunk and per column
 * (src/kernels/broadcast/super_array.rs:180-249 — "TODO: Parallelise" :193 —, table.rs:31-62, super_table.rs:38-73).
 * These take the whole list and issue ONE launch per (dtype, alignment, masked) class; aggregate i is bit-identical
 * to mnr_reduce_stats(bufs[i], validities[i]).  `validities` may be NULL (all dense) or hold NULL entries.
 * The async form writes n x 32 bytes to DEVICE memory `out_device` on the context stream. */
int mnr_reduce_stats_batch(mnr_ctx* ctx, size_t n, const mnr_buf* const* bufs, const mnr_bits* const* validities,
                           int with_minmax, mnr_agg* out_host);                                      /* syncs */
int mnr_reduce_stats_batch_async(mnr_ctx* ctx, size_t n, const mnr_buf* const* bufs, const mnr_bits* const* validities,
                                 int with_minmax, void* out_device);
/* ---- fused reduction + cross-GPU exchange (SuperArray shards over the GPUs of one box) -----------------------
 * Each rank (one per GPU) owns a mailbox in its HBM that every peer can store into.  One process per GPU: publish the
 * 64-byte CUDA IPC handle (any transport: the Python harness all-gathers them with torch.distributed) and call
 * mnr_xchg_connect.  One process, many GPUs: mnr_xchg_connect_local (peer access), or simply mnr_group_create below.
 * After that mnr_reduce_stats_exchange is ONE kernel per call: the shard's null-aware aggregate, P2P stores of the 32-byte
 * partial into every peer's mailbox through NVLink/NVSwitch, a flag wait, and the rank-order combine — every rank
 * ends with the same global aggregate (bit-identical, floats included) in `out_device`.  Collective: every rank of
 * the group must make the same sequence of exchange calls.  world <= 16.
 * Failure: if a peer's partial never arrives (~10 s) the kernel sets the exchange's error word and marks the result
 * unusable (count = UINT64_MAX).  The _sync forms return MNR_ERR_CUDA and clear the word; asynchronous callers poll
 * mnr_xchg_status.  An epoch is consumed only by a launch that actually happened. */
#define MNR_IPC_HANDLE_BYTES 64
#define MNR_XCHG_MAX_AGGS 64   /* aggregates (columns) one exchange can carry */
int mnr_xchg_create(mnr_ctx* ctx, int world, int rank, mnr_xchg** out);
int mnr_xchg_local_handle(mnr_xchg* x, uint8_t* handle64);
int mnr_xchg_connect(mnr_xchg* x, const uint8_t* handles /* world x MNR_IPC_HANDLE_BYTES, rank order */);
/* Same process: peers[r] = rank r's exchange (peers[own rank] ignored).  Enables peer access between the devices; two
 * ranks may share one device ("virtual ranks": the whole exchange path runs on a single-GPU box). */
int mnr_xchg_connect_local(mnr_xchg* x, mnr_xchg* const* peers);
/* *timed_out = 1 if an exchange on `x` gave up waiting since the last clear; synchronises the context stream. */
int mnr_xchg_status(mnr_xchg* x, int clear, int* timed_out);
void mnr_xchg_destroy(mnr_xchg* x);
int mnr_reduce_stats_exchange(mnr_ctx* ctx, mnr_xchg* x, const mnr_buf* buf, const mnr_bits* validity, int with_minmax,
                              void* out_device);
int mnr_reduce_stats_exchange_sync(mnr_ctx* ctx, mnr_xchg* x, const mnr_buf* buf, const mnr_bits* validity,
                                   int with_minmax, mnr_agg* out_host);
/* Sharded SuperArray / SuperTable reduction in one call per rank (benches/benchmark_parallel_simd.rs:81-97 par_chunks ->
 * chunk sums -> combine, over the chunk lists of broadcast/super_array.rs:180-249 and super_table.rs:38-73): the `n`
 * LOCAL chunks (chunk i belongs to column col_of_chunk[i] < n_cols, dtype col_dtypes[col]) are reduced by batched
 * launches; the block that writes the last chunk aggregate folds them per column in chunk order, exchanges the n_cols
 * per-column partials with every peer in ONE mailbox epoch and folds those in rank order.  out_device receives n_cols x
 * 32 bytes, identical on every rank.  A rank may hold no chunk at all (n = 0): it still takes part.  x = NULL: local
 * per-column fold only (single GPU).  n_cols <= MNR_XCHG_MAX_AGGS. */
int mnr_reduce_stats_batch_exchange(mnr_ctx* ctx, mnr_xchg* x, size_t n, const mnr_buf* const* bufs,
                                    const mnr_bits* const* validities, int with_minmax, size_t n_cols,
                                    const uint32_t* col_of_chunk, const mnr_dtype* col_dtypes, void* out_device);
int mnr_reduce_stats_batch_exchange_sync(mnr_ctx* ctx, mnr_xchg* x, size_t n, const mnr_buf* const* bufs,
                                         const mnr_bits* const* validities, int with_minmax, size_t n_cols,
                                         const uint32_t* col_of_chunk, const mnr_dtype* col_dtypes, mnr_agg* out_host);

/* ---- sharding: chunks -> GPUs --------------------------------------------------------------------------
 * Chunk i of n lives on rank floor(i * world / n): contiguous blocks, so global row order is preserved and results
 * re-assemble as a SuperArray with the same chunk boundaries.  Pure host arithmetic. */
int mnr_shard_owner(size_t chunk, size_t n_chunks, int world);   /* rank >= 0, or a negative mnr_status */
int mnr_shard_chunk_range(size_t n_chunks, int world, int rank, size_t* lo, size_t* hi);   /* chunks [lo, hi) */
/* One big Array cut into `world` contiguous windows on `align`-row boundaries (64 rows = one validity word). */
int mnr_shard_row_range(size_t n_rows, int world, int rank, size_t align, size_t* offset, size_t* len);

/* One process driving all GPUs of the box: `world` contexts (own streams) + mailboxes connected through peer access.
 * devices = NULL means devices 0..world-1; a device may be listed more than once (virtual ranks).  The mnr_group_*
 * calls route every chunk to the context that owns it (mnr_buf carries its context) and launch per device; the
 * per-rank handles are ordinary mnr_ctx / mnr_xchg for everything else in this header. */
int mnr_group_create(int world, const int* devices, mnr_group** out);
void mnr_group_destroy(mnr_group* g);
int mnr_group_world(const mnr_group* g);
mnr_ctx* mnr_group_ctx(mnr_group* g, int rank);
mnr_xchg* mnr_group_xchg(mnr_group* g, int rank);
int mnr_group_synchronize(mnr_group* g);
/* Upload the chunks of a host SuperArray: chunk i -> rank mnr_shard_owner(i, n_chunks, world), all links copying at the
 * same time (pinned host memory).  host_validity may be NULL or hold NULL entries; out_validity[i] is NULL for those. */
int mnr_group_upload(mnr_group* g, mnr_dtype dtype, size_t n_chunks, const void* const* host_chunks, const size_t* lens,
                     const uint8_t* const* host_validity, mnr_buf** out_bufs, mnr_bits** out_validity);
/* Shard-local element-wise fan-out (no communication): chunk pair i runs on the device that owns it; lhs[i] and rhs[i]
 * must live on the same rank.  One batched launch per device and (dtype, alignment, masked) class; fresh outputs on the
 * owning device.  Semantics per chunk = mnr_ew_binary / mnr_ew_scalar. */
int mnr_group_ew_binary(mnr_group* g, mnr_op op, size_t n, const mnr_buf* const* lhs, const mnr_buf* const* rhs,
                        const mnr_bits* const* lhs_mask, const mnr_bits* const* rhs_mask, mnr_mask_mode mode, mnr_buf** out,
                        mnr_bits** out_mask);
int mnr_group_ew_scalar(mnr_group* g, mnr_op op, size_t n, const mnr_buf* const* arrs, const void* const* scalars,
                        int scalar_is_lhs, const mnr_bits* const* masks, mnr_buf** out, mnr_bits** out_mask);
/* Per-column {sum, min, max, count} of a sharded SuperArray / SuperTable: mnr_reduce_stats_batch_exchange on every rank
 * (the fused NVLink exchange), then rank 0's n_cols aggregates are copied to out_host.  Synchronises the group. */
int mnr_group_reduce_stats(mnr_group* g, size_t n, const mnr_buf* const* bufs, const mnr_bits* const* validities,
                           int with_minmax, size_t n_cols, const uint32_t* col_of_chunk, const mnr_dtype* col_dtypes,
                           mnr_agg* out_host);
/* mean = (double)sum / (double)count on the host from a (combined) aggregate; NaN when count == 0. */
double mnr_agg_mean(mnr_dtype dtype, const mnr_agg* agg);
/* Combine per-chunk / per-GPU partials in index order (the documented rank-order float add). */
int mnr_agg_combine(mnr_dtype dtype, const mnr_agg* partials, size_t n, mnr_agg* out);

/* ---- host-slice drop-ins: exactly the reference leaf signatures ------------------------------------------
 * `fn apply_int_i64(lhs:&[i64], rhs:&[i64], op, mask:Option<&Bitmask>) -> Result<IntegerArray<i64>,KernelError>`
 * (dispatch.rs:74-79,147-152): host pointers in, host pointers out; upload, kernel and download are pipelined
 * in chunks on the context's copy streams.  `mask` is the single pre-merged validity (NULL = None); `out_mask`
 * (ceil(len/8) bytes) is written iff `mask` is given.  Fastest with pinned (page-locked) host memory. */
int mnr_apply_host(mnr_ctx* ctx, mnr_dtype dtype, mnr_op op, const void* lhs, size_t lhs_len, const void* rhs,
                   size_t rhs_len, const uint8_t* mask, void* out, uint8_t* out_mask);
int mnr_apply_int_i32(mnr_ctx*, const int32_t* lhs, size_t lhs_len, const int32_t* rhs, size_t rhs_len, mnr_op op,
                      const uint8_t* mask, int32_t* out, uint8_t* out_mask);
int mnr_apply_int_u32(mnr_ctx*, const uint32_t* lhs, size_t lhs_len, const uint32_t* rhs, size_t rhs_len, mnr_op op,
                      const uint8_t* mask, uint32_t* out, uint8_t* out_mask);
int mnr_apply_int_i64(mnr_ctx*, const int64_t* lhs, size_t lhs_len, const int64_t* rhs, size_t rhs_len, mnr_op op,
                      const uint8_t* mask, int64_t* out, uint8_t* out_mask);
int mnr_apply_int_u64(mnr_ctx*, const uint64_t* lhs, size_t lhs_len, const uint64_t* rhs, size_t rhs_len, mnr_op op,
                      const uint8_t* mask, uint64_t* out, uint8_t* out_mask);
int mnr_apply_float_f32(mnr_ctx*, const float* lhs, size_t lhs_len, const float* rhs, size_t rhs_len, mnr_op op,
                        const uint8_t* mask, float* out, uint8_t* out_mask);
int mnr_apply_float_f64(mnr_ctx*, const double* lhs, size_t lhs_len, const double* rhs, size_t rhs_len, mnr_op op,
                        const uint8_t* mask, double* out, uint8_t* out_mask);
/* apply_fma_f32 / apply_fma_f64 (dispatch.rs:221-226). */
int mnr_apply_fma_host(mnr_ctx* ctx, mnr_dtype dtype, const void* lhs, size_t lhs_len, const void* rhs,
                       size_t rhs_len, const void* acc, size_t acc_len, const uint8_t* mask, void* out,
                       uint8_t* out_mask);
/* Null-aware aggregate of a host column (the sum the benches time, plus count/min/max): chunked upload
 * overlapped with the reduction kernel; one 32-byte result comes back. */
int mnr_stats_host(mnr_ctx* ctx, mnr_dtype dtype, const void* data, size_t len, const uint8_t* validity,
                   int with_minmax, mnr_agg* out);
/* bitmask_binop over host bytes (BitmaskVT windows, byte-floored offsets as above). */
int mnr_bitmask_binop_host(mnr_ctx* ctx, mnr_logical_op op, const uint8_t* lhs, size_t lhs_offset,
                           const uint8_t* rhs, size_t rhs_offset, size_t len, uint8_t* out);
/* Page-lock / unlock a caller buffer so the drop-ins above copy at full PCIe speed. */
int mnr_host_register(void* ptr, size_t bytes);
int mnr_host_unregister(void* ptr);
/* Pinned host allocation (the download side of SharedBuffer::from_owner, src/structs/shared_buffer/mod.rs:187). */
int mnr_host_alloc(size_t bytes, void** out);
void mnr_host_free(void* ptr);

/* ---- Arrow C Data Interface <-> device buffers ----------------------------------------------------------------
 * The reference's own C ABI (src/ffi/arrow_c_ffi.rs:87-98,121-133; export_to_c :432-470 with buffers =
 * [validity | NULL, values], n_buffers = 2, :481-490; import_from_c :640).  Struct layouts are the Arrow spec's. */
#ifndef ARROW_C_DATA_INTERFACE
#define ARROW_C_DATA_INTERFACE
struct ArrowSchema {
    const char* format; const char* name; const char* metadata; int64_t flags; int64_t n_children;
    struct ArrowSchema** children; struct ArrowSchema* dictionary;
    void (*release)(struct ArrowSchema*); void* private_data;
};
struct ArrowArray {
    int64_t length; int64_t null_count; int64_t offset; int64_t n_buffers; int64_t n_children;
    const void** buffers; struct ArrowArray** children; struct ArrowArray* dictionary;
    void (*release)(struct ArrowArray*); void* private_data;
};
#endif
/* Upload a HOST ArrowArray of a Minarrow numeric type (formats c C s S i I l L f g) or boolean (b).  `offset` is
 * honoured: element offset for values, an arbitrary BIT offset for validity / boolean data (shifted on the device).
 * Numeric: *values is set, data_bits may be NULL.  Boolean: *data_bits is set, values may be NULL.  *validity is
 * NULL when the array has no validity buffer or null_count == 0.  The array is borrowed: release stays with the caller. */
int mnr_arrow_import(mnr_ctx* ctx, const struct ArrowArray* array, const struct ArrowSchema* schema, mnr_buf** values,
                     mnr_bits** data_bits, mnr_bits** validity);
/* Download into a fresh host ArrowArray (64-byte aligned buffers, offset 0, null_count filled in) that the consumer
 * releases through array->release, like export_to_c.  `out_schema` may be NULL. */
int mnr_arrow_export(mnr_ctx* ctx, const mnr_buf* values, const mnr_bits* validity, struct ArrowArray* out_array,
                     struct ArrowSchema* out_schema);
int mnr_arrow_export_bool(mnr_ctx* ctx, const mnr_bits* data_bits, const mnr_bits* validity,
                          struct ArrowArray* out_array, struct ArrowSchema* out_schema);

/* Arrow C Stream Interface (src/ffi/arrow_c_ffi.rs:153-168 ArrowArrayStream; the PyCapsule route a chunked producer uses,
 * :160-168) -> SuperArray chunks on this device.  The stream is drained to its end; chunk i is uploaded (exactly like
 * mnr_arrow_import, offsets honoured) iff chunk_lo <= i < chunk_hi — the caller passes the contiguous block of chunks
 * that `chunk i -> rank floor(i*G/n)` assigns to this rank (minarrow_b200/sharded.py), the other chunks are released
 * unread.  values[k] / validity[k] (k < *n_imported <= capacity) receive the uploaded chunks in stream order;
 * validity[k] is NULL for a chunk without nulls.  *n_seen = chunks in the stream.  Every consumed ArrowArray is
 * released here; the stream itself stays with the caller (release after the call).  Numeric formats only. */
#ifndef ARROW_C_STREAM_INTERFACE
#define ARROW_C_STREAM_INTERFACE
struct ArrowArrayStream {
    int (*get_schema)(struct ArrowArrayStream*, struct ArrowSchema* out);
    int (*get_next)(struct ArrowArrayStream*, struct ArrowArray* out);
    const char* (*get_last_error)(struct ArrowArrayStream*);
    void (*release)(struct ArrowArrayStream*);
    void* private_data;
};
#endif
int mnr_arrow_stream_import(mnr_ctx* ctx, struct ArrowArrayStream* stream, size_t chunk_lo, size_t chunk_hi, size_t capacity,
                            mnr_buf** values, mnr_bits** validity, size_t* n_imported, size_t* n_seen);

#ifdef __cplusplus
}
#endif
#endif /* MINARROW_B200_H */
