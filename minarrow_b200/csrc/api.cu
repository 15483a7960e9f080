// C ABI of libminarrow_b200.so (include/minarrow_b200.h): contexts, device-resident buffers / bitmasks,
// argument validation with the reference's error behaviour, and the host-slice drop-in pipelines.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <initializer_list>
#include <limits>
#include <string>
#include <algorithm>
#include <vector>

#include "internal.h"
#include "ew_kernels.cuh"
#include "reduce_kernels.cuh"

namespace mnr {
__global__ void clear_trailing_kernel(uint8_t* bits, uint64_t len) {
    if (len & 7) bits[(len - 1) >> 3] &= (uint8_t)((1u << (unsigned)(len & 7)) - 1u);
}
}  // namespace mnr

using namespace mnr;

// ---- errors ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
static int fail_cuda(cudaError_t e, const char* what) {
    const int code = (e == cudaErrorMemoryAllocation) ? MNR_ERR_OUT_OF_MEMORY
                     : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice)
                         ? MNR_ERR_NO_DEVICE
                         : MNR_ERR_CUDA;
    return fail(code, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}
#define CU(x)                                            \
    do {                                                 \
        cudaError_t e__ = (x);                           \
        if (e__ != cudaSuccess) return fail_cuda(e__, #x); \
    } while (0)
#define REQUIRE(cond, code, ...) \
    do {                         \
        if (!(cond)) return fail(code, __VA_ARGS__); \
    } while (0)

static bool valid_dtype(int d) { return d >= MNR_I32 && d <= MNR_U16; }
static bool is_float_dtype(int d) { return d == MNR_F32 || d == MNR_F64; }
static size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }
static size_t mask_bytes(size_t bits) { return (bits + 7) >> 3; }

namespace mnr {
int fail_public(int code, const char* msg) { return fail(code, "%s", msg); }   // error reporting for the other translation units
}  // namespace mnr

extern "C" {

int mnr_abi_version(void) { return MNR_ABI_VERSION; }
const char* mnr_last_error(void) { return g_err.c_str(); }

int mnr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ---- context -------------------------------------------------------------------------------------------------
static void ctx_release(mnr_ctx* c) {
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int i = 0; i < 3; ++i) {
        if (c->slot_stream[i]) { cudaStreamSynchronize(c->slot_stream[i]); cudaStreamDestroy(c->slot_stream[i]); }
        for (int j = 0; j < 6; ++j) if (c->stage[i][j]) cudaFree(c->stage[i][j]);
    }
    for (int i = 0; i < 4; ++i) { if (c->partials[i]) cudaFree(c->partials[i]); if (c->ticket[i]) cudaFree(c->ticket[i]); }
    if (c->chunk_aggs) cudaFree(c->chunk_aggs);
    if (c->ew_segs) cudaFree(c->ew_segs);
    if (c->batch_partials) cudaFree(c->batch_partials);
    if (c->batch_segs) cudaFree(c->batch_segs);
    if (c->batch_tickets) cudaFree(c->batch_tickets);
    if (c->fold_desc) cudaFree(c->fold_desc);
    if (c->fold_local) cudaFree(c->fold_local);
    if (c->fold_result) cudaFree(c->fold_result);
    if (c->d_agg) cudaFree(c->d_agg);
    if (c->d_count) cudaFree(c->d_count);
    if (c->h_scratch) cudaFreeHost(c->h_scratch);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    cudaGetLastError();
    delete c;
}

static int ctx_alloc(mnr_ctx* c, cudaStream_t stream, bool own) {
    if (own) CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    else c->stream = stream;
    for (int i = 0; i < 3; ++i) CU(cudaStreamCreateWithFlags(&c->slot_stream[i], cudaStreamNonBlocking));
    size_t pbytes = sizeof(AggRaw) * (size_t)reduce_max_grid();
    if (pbytes < sizeof(unsigned long long) * (size_t)popcount_max_grid()) pbytes = sizeof(unsigned long long) * (size_t)popcount_max_grid();
    for (int i = 0; i < 4; ++i) {
        CU(cudaMalloc(&c->partials[i], pbytes));
        CU(cudaMalloc(&c->ticket[i], 64));
        CU(cudaMemset(c->ticket[i], 0, 64));
    }
    CU(cudaMalloc(&c->d_agg, sizeof(AggRaw)));
    CU(cudaMalloc(&c->d_count, 64));
    CU(cudaMalloc(&c->fold_local, sizeof(AggRaw) * MNR_XCHG_MAX_AGGS));
    CU(cudaMalloc(&c->fold_result, sizeof(AggRaw) * MNR_XCHG_MAX_AGGS));
    CU(cudaHostAlloc(&c->h_scratch, 256, cudaHostAllocMapped | cudaHostAllocPortable));   // kernels store results here directly
    memset(c->h_scratch, 0, 256);
    cudaMemPool_t pool;
    CU(cudaDeviceGetDefaultMemPool(&pool, c->device));
    uint64_t thr = UINT64_MAX;   // keep freed blocks cached: fresh outputs per call without cudaMalloc cost
    CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    return MNR_OK;
}

static int ctx_init(int device, cudaStream_t stream, bool own, mnr_ctx** out) {
    REQUIRE(out, MNR_ERR_INVALID_ARGUMENTS, "mnr_ctx_create: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(MNR_ERR_NO_DEVICE, "no CUDA device (%s); minarrow_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    REQUIRE(device >= 0 && device < n, MNR_ERR_NO_DEVICE, "device %d out of range (have %d)", device, n);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    // The library carries sm_100a SASS only: arch-specific ("a") code runs on exactly that compute capability —
    // an sm_103 / sm_120 part would pass a ">= 10" test and fail at the first launch.
    REQUIRE(prop.major == 10 && prop.minor == 0, MNR_ERR_NO_DEVICE,
            "device %d is sm_%d%d; this library carries sm_100a code only (no fallback path)", device, prop.major,
            prop.minor);
    CU(cudaSetDevice(device));
    mnr_ctx* c = new mnr_ctx();
    c->device = device;
    c->own_stream = own;
    const int rc = ctx_alloc(c, stream, own);
    if (rc) { ctx_release(c); return rc; }   // nothing leaks when a later allocation fails
    *out = c;
    return MNR_OK;
}

int mnr_ctx_create(int device, mnr_ctx** out) { return ctx_init(device, nullptr, true, out); }
int mnr_ctx_create_on_stream(int device, void* cuda_stream, mnr_ctx** out) {
    return ctx_init(device, static_cast<cudaStream_t>(cuda_stream), false, out);
}

void mnr_ctx_destroy(mnr_ctx* c) {
    if (!c) return;
    ctx_release(c);
}

int mnr_ctx_synchronize(mnr_ctx* c) {
    REQUIRE(c, MNR_ERR_INVALID_ARGUMENTS, "ctx is NULL");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return MNR_OK;
}
int mnr_ctx_device(const mnr_ctx* c) { return c ? c->device : -1; }
void* mnr_ctx_stream(const mnr_ctx* c) { return c ? (void*)c->stream : nullptr; }
uint64_t mnr_ctx_launch_count(const mnr_ctx* c) { return c ? c->launches : 0; }

int mnr_ctx_set_option(mnr_ctx* c, const char* key, int64_t value) {
    REQUIRE(c && key, MNR_ERR_INVALID_ARGUMENTS, "ctx/key is NULL");
    if (!strcmp(key, "ew_grid_cap")) { c->knobs.grid_cap = (int)value; return MNR_OK; }
    if (!strcmp(key, "ew_max_tier")) { c->knobs.max_tier = (int)value; return MNR_OK; }
    if (!strcmp(key, "ew_sdiv64_cfg")) { c->knobs.sdiv64_cfg = (int)value; return MNR_OK; }
    if (!strcmp(key, "ew_fdiv_cfg")) { c->knobs.fdiv_cfg = (int)value; return MNR_OK; }
    if (!strcmp(key, "ew_heavy_cfg")) { c->knobs.heavy_cfg = (int)value; return MNR_OK; }
    if (!strcmp(key, "ew_cheap8_cfg")) { c->knobs.cheap8_cfg = (int)value; return MNR_OK; }
    if (!strcmp(key, "reduce_overlap")) { c->reduce_overlap = value != 0; return MNR_OK; }
    if (!strcmp(key, "host_chunk_rows")) {
        REQUIRE(value >= 1024 && value % 1024 == 0, MNR_ERR_INVALID_ARGUMENTS, "host_chunk_rows must be a multiple of 1024");
        c->host_chunk_rows = (size_t)value;
        return MNR_OK;
    }
    return fail(MNR_ERR_INVALID_ARGUMENTS, "unknown option '%s'", key);
}

// ---- buffers ---------------------------------------------------------------------------------------------------
static int dev_alloc(mnr_ctx* c, size_t bytes, void** p) {
    CU(cudaSetDevice(c->device));
    CU(cudaMallocAsync(p, pad256(bytes ? bytes : 1), c->stream));
    return MNR_OK;
}

int mnr_buf_alloc(mnr_ctx* c, mnr_dtype dtype, size_t len, mnr_buf** out) {
    REQUIRE(c && out, MNR_ERR_INVALID_ARGUMENTS, "ctx/out is NULL");
    REQUIRE(valid_dtype(dtype), MNR_ERR_UNSUPPORTED_TYPE, "unknown dtype %d", (int)dtype);
    void* p = nullptr;
    int rc = dev_alloc(c, len * dtype_size(dtype), &p);
    if (rc) return rc;
    *out = new mnr_buf{c, dtype, p, len, true};
    return MNR_OK;
}

int mnr_buf_upload_async(mnr_ctx* c, mnr_dtype dtype, const void* host, size_t len, mnr_buf** out) {
    REQUIRE(host || len == 0, MNR_ERR_INVALID_ARGUMENTS, "host pointer is NULL");
    int rc = mnr_buf_alloc(c, dtype, len, out);
    if (rc) return rc;
    if (len) {
        cudaError_t e = cudaMemcpyAsync((*out)->ptr, host, len * dtype_size(dtype), cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) { mnr_buf_free(*out); *out = nullptr; return fail_cuda(e, "cudaMemcpyAsync (upload)"); }
    }
    return MNR_OK;
}

int mnr_buf_upload(mnr_ctx* c, mnr_dtype dtype, const void* host, size_t len, mnr_buf** out) {
    int rc = mnr_buf_upload_async(c, dtype, host, len, out);
    if (rc) return rc;
    if (len) CU(cudaStreamSynchronize(c->stream));   // the caller may reuse `host` on return
    return MNR_OK;
}

int mnr_buf_wrap(mnr_ctx* c, mnr_dtype dtype, void* device_ptr, size_t len, mnr_buf** out) {
    REQUIRE(c && out, MNR_ERR_INVALID_ARGUMENTS, "ctx/out is NULL");
    REQUIRE(valid_dtype(dtype), MNR_ERR_UNSUPPORTED_TYPE, "unknown dtype %d", (int)dtype);
    REQUIRE(device_ptr || len == 0, MNR_ERR_INVALID_ARGUMENTS, "device pointer is NULL");
    REQUIRE((reinterpret_cast<uintptr_t>(device_ptr) % dtype_size(dtype)) == 0, MNR_ERR_INVALID_ARGUMENTS,
            "device pointer is not aligned to the element size");
    *out = new mnr_buf{c, dtype, device_ptr, len, false};
    return MNR_OK;
}

int mnr_buf_slice(const mnr_buf* parent, size_t offset, size_t len, mnr_buf** out) {
    REQUIRE(parent && out, MNR_ERR_INVALID_ARGUMENTS, "parent/out is NULL");
    REQUIRE(offset <= parent->len && len <= parent->len - offset, MNR_ERR_OUT_OF_BOUNDS,
            "slice [%zu, %zu) out of bounds for length %zu", offset, offset + len, parent->len);
    *out = new mnr_buf{parent->ctx, parent->dtype, static_cast<char*>(parent->ptr) + offset * dtype_size(parent->dtype),
                       len, false};
    return MNR_OK;
}

int mnr_buf_download(mnr_ctx* c, const mnr_buf* b, void* host) {
    REQUIRE(c && b && (host || b->len == 0), MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    CU(cudaSetDevice(c->device));
    if (b->len) CU(cudaMemcpyAsync(host, b->ptr, b->len * dtype_size(b->dtype), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return MNR_OK;
}
size_t mnr_buf_len(const mnr_buf* b) { return b ? b->len : 0; }
int mnr_buf_dtype(const mnr_buf* b) { return b ? (int)b->dtype : -1; }
void* mnr_buf_device_ptr(const mnr_buf* b) { return b ? b->ptr : nullptr; }
void mnr_buf_free(mnr_buf* b) {
    if (!b) return;
    if (b->owned && b->ptr) { cudaSetDevice(b->ctx->device); cudaFreeAsync(b->ptr, b->ctx->stream); }
    delete b;
}

int mnr_bits_alloc(mnr_ctx* c, size_t len_bits, mnr_bits** out) {
    REQUIRE(c && out, MNR_ERR_INVALID_ARGUMENTS, "ctx/out is NULL");
    void* p = nullptr;
    int rc = dev_alloc(c, mask_bytes(len_bits), &p);
    if (rc) return rc;
    *out = new mnr_bits{c, static_cast<uint8_t*>(p), len_bits, true};
    return MNR_OK;
}

int mnr_bits_new_set_all(mnr_ctx* c, size_t len_bits, int value, mnr_bits** out) {
    int rc = mnr_bits_alloc(c, len_bits, out);
    if (rc) return rc;
    if (len_bits) {
        CU(cudaMemsetAsync((*out)->ptr, value ? 0xFF : 0, mask_bytes(len_bits), c->stream));
        if (value && (len_bits & 7)) { clear_trailing_kernel<<<1, 1, 0, c->stream>>>((*out)->ptr, len_bits); c->launches++; }
        CU(cudaGetLastError());
    }
    return MNR_OK;
}

int mnr_bits_upload_async(mnr_ctx* c, const uint8_t* host_bytes, size_t len_bits, mnr_bits** out) {
    REQUIRE(host_bytes || len_bits == 0, MNR_ERR_INVALID_ARGUMENTS, "host pointer is NULL");
    int rc = mnr_bits_alloc(c, len_bits, out);
    if (rc) return rc;
    if (len_bits) {
        cudaError_t e = cudaMemcpyAsync((*out)->ptr, host_bytes, mask_bytes(len_bits), cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess && (len_bits & 7)) {
            clear_trailing_kernel<<<1, 1, 0, c->stream>>>((*out)->ptr, len_bits);
            c->launches++;
            e = cudaGetLastError();
        }
        if (e != cudaSuccess) { mnr_bits_free(*out); *out = nullptr; return fail_cuda(e, "mnr_bits_upload"); }
    }
    return MNR_OK;
}

int mnr_bits_upload(mnr_ctx* c, const uint8_t* host_bytes, size_t len_bits, mnr_bits** out) {
    int rc = mnr_bits_upload_async(c, host_bytes, len_bits, out);
    if (rc) return rc;
    if (len_bits) CU(cudaStreamSynchronize(c->stream));
    return MNR_OK;
}

int mnr_bits_wrap(mnr_ctx* c, void* device_ptr, size_t len_bits, mnr_bits** out) {
    REQUIRE(c && out, MNR_ERR_INVALID_ARGUMENTS, "ctx/out is NULL");
    REQUIRE(device_ptr || len_bits == 0, MNR_ERR_INVALID_ARGUMENTS, "device pointer is NULL");
    *out = new mnr_bits{c, static_cast<uint8_t*>(device_ptr), len_bits, false};
    return MNR_OK;
}

int mnr_bits_download(mnr_ctx* c, const mnr_bits* b, uint8_t* host_bytes) {
    REQUIRE(c && b && (host_bytes || b->len == 0), MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    CU(cudaSetDevice(c->device));
    if (b->len) CU(cudaMemcpyAsync(host_bytes, b->ptr, mask_bytes(b->len), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return MNR_OK;
}
size_t mnr_bits_len(const mnr_bits* b) { return b ? b->len : 0; }
void* mnr_bits_device_ptr(const mnr_bits* b) { return b ? b->ptr : nullptr; }
void mnr_bits_free(mnr_bits* b) {
    if (!b) return;
    if (b->owned && b->ptr) { cudaSetDevice(b->ctx->device); cudaFreeAsync(b->ptr, b->ctx->stream); }
    delete b;
}

// ---- element-wise ------------------------------------------------------------------------------------------------
static uint64_t scalar_to_bits(mnr_dtype dt, const void* scalar) {
    uint64_t b = 0;
    memcpy(&b, scalar, dtype_size(dt));
    return b;
}

// Integer Div / Rem / FloorDiv whose divisor is the broadcast scalar (rhs == NULL) and non-zero: hand the kernel the
// scalar's multiplicative inverse (divmagic.h) so no row executes a divide.  A zero scalar keeps the generic kernel,
// which nulls every row (masked) or raises the divide-by-zero flag (dense).
static void prepare_scalar_division(EwArgs& a) {
    a.sdiv = 0;
    a.nonzero_divisor = 0;
    if (is_float_dtype(a.dtype) || a.rhs != nullptr || a.lhs == nullptr) return;
    if (a.op != MNR_DIV && a.op != MNR_REM && a.op != MNR_FLOORDIV) return;
    const size_t sz = dtype_size(a.dtype);
    const bool is_signed = a.dtype == MNR_I8 || a.dtype == MNR_I16 || a.dtype == MNR_I32 || a.dtype == MNR_I64;
    uint64_t u = a.scalar_bits;
    if (sz < 8) u &= (1ull << (8 * sz)) - 1ull;
    if (u == 0) return;
    a.nonzero_divisor = 1;
    // 8/16-bit columns: the packed division path of the generic kernel takes the scalar as a broadcast word (its reciprocal
    // is loop-invariant) and beats the widened multiplicative inverse (r02e: i8 0.25 -> the two-column rate).
    if (sz <= 2) return;
    const int nbits = sz == 8 ? 64 : 32;   // 8/16-bit columns are widened to 32 bits in the kernel
    DivMagic k;
    if (is_signed) {
        const int64_t d = sz == 8 ? (int64_t)u : sz == 4 ? (int64_t)(int32_t)u : sz == 2 ? (int64_t)(int16_t)u : (int64_t)(int8_t)u;
        k = div_magic_signed(d, nbits);
    } else {
        k = div_magic_unsigned(u, nbits);
    }
    a.sdiv = 1;
    a.magic_m = k.m; a.magic_s1 = k.s1; a.magic_s2 = k.s2;
}

// Launch + (dense integer Div/Rem/FloorDiv only) the reference's divide-by-zero panic as an error code.
static int run_ew(mnr_ctx* c, EwArgs& a, cudaStream_t s, bool promote, mnr_dtype lt, mnr_dtype rt) {
    if (a.n == 0) return MNR_OK;
    if (!promote) prepare_scalar_division(a);
    const bool dense_int_div = !a.lmask && !a.rmask && !is_float_dtype(a.dtype) && !a.sdiv && !a.nonzero_divisor &&
                               (a.op == MNR_DIV || a.op == MNR_REM || a.op == MNR_FLOORDIV);
    a.div0_flag = c->ticket[0] + 8;
    if (dense_int_div) CU(cudaMemsetAsync(a.div0_flag, 0, 4, s));
    CU(promote ? launch_ew_promote(a, lt, rt, s) : launch_ew_binary(a, s));
    c->launches++;
    if (dense_int_div) {
        unsigned int* h = static_cast<unsigned int*>(c->h_scratch);
        CU(cudaMemcpyAsync(h, a.div0_flag, 4, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        if (*h) return fail(MNR_ERR_DIVIDE_BY_ZERO, "%s by zero in dense integer kernel (the reference panics here)",
                            a.op == MNR_DIV ? "Division" : a.op == MNR_REM ? "Remainder" : "Floor division");
    }
    return MNR_OK;
}

// The kernels read their inputs through the read-only (non-coherent) path and declare every pointer __restrict__: an
// output that overlaps an input is undefined behaviour, so the *_into entry points refuse it (in-place `x = x op y` needs
// a fresh output; the stream-ordered pool makes that allocation cheap).
static bool overlaps(const void* a, size_t abytes, const void* b, size_t bbytes) {
    if (!a || !b || !abytes || !bbytes) return false;
    const uintptr_t pa = reinterpret_cast<uintptr_t>(a), pb = reinterpret_cast<uintptr_t>(b);
    return pa < pb + bbytes && pb < pa + abytes;
}
static int check_no_alias(const mnr_buf* out, const mnr_bits* out_mask, std::initializer_list<const mnr_buf*> ins,
                          std::initializer_list<const mnr_bits*> in_masks) {
    const size_t ob = out ? out->len * dtype_size(out->dtype) : 0, omb = out_mask ? mask_bytes(out_mask->len) : 0;
    for (const mnr_buf* b : ins) {
        if (!b) continue;
        const size_t bb = b->len * dtype_size(b->dtype);
        REQUIRE(!(out && overlaps(out->ptr, ob, b->ptr, bb)) && !(out_mask && overlaps(out_mask->ptr, omb, b->ptr, bb)),
                MNR_ERR_INVALID_ARGUMENTS, "an output overlaps an input buffer (in-place operation is not supported: pass a fresh output)");
    }
    for (const mnr_bits* m : in_masks) {
        if (!m) continue;
        const size_t mb = mask_bytes(m->len);
        REQUIRE(!(out && overlaps(out->ptr, ob, m->ptr, mb)) && !(out_mask && overlaps(out_mask->ptr, omb, m->ptr, mb)),
                MNR_ERR_INVALID_ARGUMENTS, "an output overlaps an input mask (in-place operation is not supported: pass a fresh output)");
    }
    REQUIRE(!(out && out_mask && overlaps(out->ptr, ob, out_mask->ptr, omb)), MNR_ERR_INVALID_ARGUMENTS, "output values and output mask overlap");
    return MNR_OK;
}

static int check_masks(const mnr_bits* lm, const mnr_bits* rm, size_t n) {
    REQUIRE(!lm || lm->len >= n, MNR_ERR_INVALID_ARGUMENTS, "lhs mask has %zu bits, need %zu", lm->len, n);
    REQUIRE(!rm || rm->len >= n, MNR_ERR_INVALID_ARGUMENTS, "rhs mask has %zu bits, need %zu", rm->len, n);
    return MNR_OK;
}

int mnr_ew_binary_into(mnr_ctx* c, mnr_op op, const mnr_buf* lhs, const mnr_buf* rhs, const mnr_bits* lm,
                       const mnr_bits* rm, mnr_mask_mode mode, mnr_buf* out, mnr_bits* out_mask) {
    REQUIRE(c && lhs && rhs && out, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    REQUIRE(op >= MNR_ADD && op <= MNR_FLOORDIV, MNR_ERR_OPERATOR_MISMATCH, "unknown operator %d", (int)op);
    REQUIRE(lhs->len == rhs->len, MNR_ERR_LENGTH_MISMATCH, "apply numeric: length mismatch (lhs: %zu, rhs: %zu)",
            lhs->len, rhs->len);
    REQUIRE(lhs->dtype == rhs->dtype, MNR_ERR_UNSUPPORTED_TYPE,
            "Unsupported array type combination for arithmetic operations (dtypes %d, %d)", (int)lhs->dtype,
            (int)rhs->dtype);
    REQUIRE(out->dtype == lhs->dtype && out->len == lhs->len, MNR_ERR_INVALID_ARGUMENTS, "output buffer shape/dtype mismatch");
    const bool masked = lm || rm;
    REQUIRE(!masked || (out_mask && out_mask->len == lhs->len), MNR_ERR_INVALID_ARGUMENTS,
            "masked call needs an output mask of %zu bits", lhs->len);
    int rc = check_masks(lm, rm, lhs->len);
    if (rc) return rc;
    rc = check_no_alias(out, masked ? out_mask : nullptr, {lhs, rhs}, {lm, rm});
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    EwArgs a{};
    a.dtype = lhs->dtype; a.op = op; a.lhs = lhs->ptr; a.rhs = rhs->ptr;
    a.lmask = lm ? lm->ptr : nullptr; a.rmask = rm ? rm->ptr : nullptr; a.mask_or = mode == MNR_MASK_OR;
    a.out = out->ptr; a.out_mask = masked ? out_mask->ptr : nullptr; a.n = lhs->len; a.k = c->knobs;
    return run_ew(c, a, c->stream, false, lhs->dtype, rhs->dtype);
}

static int alloc_outputs(mnr_ctx* c, mnr_dtype dt, size_t n, bool masked, mnr_buf** out, mnr_bits** out_mask) {
    REQUIRE(out && out_mask, MNR_ERR_INVALID_ARGUMENTS, "out/out_mask is NULL");
    *out = nullptr; *out_mask = nullptr;
    int rc = mnr_buf_alloc(c, dt, n, out);
    if (rc) return rc;
    if (masked) {
        rc = mnr_bits_alloc(c, n, out_mask);
        if (rc) { mnr_buf_free(*out); *out = nullptr; return rc; }
    }
    return MNR_OK;
}
static int drop_outputs(int rc, mnr_buf** out, mnr_bits** out_mask) {
    if (rc) { mnr_buf_free(*out); mnr_bits_free(*out_mask); *out = nullptr; *out_mask = nullptr; }
    return rc;
}

int mnr_ew_binary(mnr_ctx* c, mnr_op op, const mnr_buf* lhs, const mnr_buf* rhs, const mnr_bits* lm, const mnr_bits* rm,
                  mnr_mask_mode mode, mnr_buf** out, mnr_bits** out_mask) {
    REQUIRE(c && lhs && rhs, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    REQUIRE(lhs->len == rhs->len, MNR_ERR_LENGTH_MISMATCH, "apply numeric: length mismatch (lhs: %zu, rhs: %zu)",
            lhs->len, rhs->len);
    int rc = alloc_outputs(c, lhs->dtype, lhs->len, lm || rm, out, out_mask);
    if (rc) return rc;
    return drop_outputs(mnr_ew_binary_into(c, op, lhs, rhs, lm, rm, mode, *out, *out_mask), out, out_mask);
}

int mnr_ew_scalar_into(mnr_ctx* c, mnr_op op, const mnr_buf* arr, const void* scalar, int scalar_is_lhs,
                       const mnr_bits* mask, mnr_buf* out, mnr_bits* out_mask) {
    REQUIRE(c && arr && scalar && out, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    REQUIRE(op >= MNR_ADD && op <= MNR_FLOORDIV, MNR_ERR_OPERATOR_MISMATCH, "unknown operator %d", (int)op);
    REQUIRE(out->dtype == arr->dtype && out->len == arr->len, MNR_ERR_INVALID_ARGUMENTS, "output buffer shape/dtype mismatch");
    REQUIRE(!mask || (out_mask && out_mask->len == arr->len), MNR_ERR_INVALID_ARGUMENTS,
            "masked call needs an output mask of %zu bits", arr->len);
    int rc = check_masks(mask, nullptr, arr->len);
    if (rc) return rc;
    rc = check_no_alias(out, mask ? out_mask : nullptr, {arr}, {mask});
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    EwArgs a{};
    a.dtype = arr->dtype; a.op = op;
    a.lhs = scalar_is_lhs ? nullptr : arr->ptr;
    a.rhs = scalar_is_lhs ? arr->ptr : nullptr;
    a.scalar_bits = scalar_to_bits(arr->dtype, scalar);
    a.lmask = mask ? mask->ptr : nullptr; a.rmask = nullptr; a.mask_or = 0;
    a.out = out->ptr; a.out_mask = mask ? out_mask->ptr : nullptr; a.n = arr->len; a.k = c->knobs;
    return run_ew(c, a, c->stream, false, arr->dtype, arr->dtype);
}

int mnr_ew_scalar(mnr_ctx* c, mnr_op op, const mnr_buf* arr, const void* scalar, int scalar_is_lhs, const mnr_bits* mask,
                  mnr_buf** out, mnr_bits** out_mask) {
    REQUIRE(c && arr, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    int rc = alloc_outputs(c, arr->dtype, arr->len, mask != nullptr, out, out_mask);
    if (rc) return rc;
    return drop_outputs(mnr_ew_scalar_into(c, op, arr, scalar, scalar_is_lhs, mask, *out, *out_mask), out, out_mask);
}

int mnr_ew_fma_into(mnr_ctx* c, const mnr_buf* a, const mnr_buf* b, const mnr_buf* acc, const mnr_bits* mask, mnr_buf* out,
                    mnr_bits* out_mask) {
    REQUIRE(c && a && b && acc && out, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    REQUIRE(a->len == b->len, MNR_ERR_LENGTH_MISMATCH, "apply numeric: length mismatch (lhs: %zu, rhs: %zu)", a->len, b->len);
    REQUIRE(a->len == acc->len, MNR_ERR_LENGTH_MISMATCH, "acc length mismatch (lhs: %zu, rhs: %zu)", a->len, acc->len);
    REQUIRE(is_float_dtype(a->dtype) && b->dtype == a->dtype && acc->dtype == a->dtype, MNR_ERR_UNSUPPORTED_TYPE,
            "fma needs three F32 or three F64 operands");
    REQUIRE(out->dtype == a->dtype && out->len == a->len, MNR_ERR_INVALID_ARGUMENTS, "output buffer shape/dtype mismatch");
    REQUIRE(!mask || (out_mask && out_mask->len == a->len), MNR_ERR_INVALID_ARGUMENTS,
            "masked call needs an output mask of %zu bits", a->len);
    int rc = check_masks(mask, nullptr, a->len);
    if (rc) return rc;
    rc = check_no_alias(out, mask ? out_mask : nullptr, {a, b, acc}, {mask});
    if (rc) return rc;
    if (a->len == 0) return MNR_OK;
    CU(cudaSetDevice(c->device));
    CU(launch_ew_fma(a->dtype, a->ptr, b->ptr, acc->ptr, mask ? mask->ptr : nullptr, out->ptr,
                     mask ? out_mask->ptr : nullptr, a->len, c->stream));
    c->launches++;
    return MNR_OK;
}

int mnr_ew_fma(mnr_ctx* c, const mnr_buf* a, const mnr_buf* b, const mnr_buf* acc, const mnr_bits* mask, mnr_buf** out,
               mnr_bits** out_mask) {
    REQUIRE(c && a && b && acc, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    REQUIRE(a->len == b->len, MNR_ERR_LENGTH_MISMATCH, "apply numeric: length mismatch (lhs: %zu, rhs: %zu)", a->len, b->len);
    REQUIRE(a->len == acc->len, MNR_ERR_LENGTH_MISMATCH, "acc length mismatch (lhs: %zu, rhs: %zu)", a->len, acc->len);
    int rc = alloc_outputs(c, a->dtype, a->len, mask != nullptr, out, out_mask);
    if (rc) return rc;
    return drop_outputs(mnr_ew_fma_into(c, a, b, acc, mask, *out, *out_mask), out, out_mask);
}

int mnr_ew_binary_promote(mnr_ctx* c, mnr_op op, const mnr_buf* lhs, const mnr_buf* rhs, const mnr_bits* lm,
                          const mnr_bits* rm, mnr_mask_mode mode, mnr_buf** out, mnr_bits** out_mask) {
    REQUIRE(c && lhs && rhs, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    REQUIRE(op >= MNR_ADD && op <= MNR_FLOORDIV, MNR_ERR_OPERATOR_MISMATCH, "unknown operator %d", (int)op);
    REQUIRE(lhs->len == rhs->len, MNR_ERR_LENGTH_MISMATCH, "arithmetic_dispatch => Length mismatch: LHS %zu RHS %zu",
            lhs->len, rhs->len);
    if (lhs->dtype == rhs->dtype) return mnr_ew_binary(c, op, lhs, rhs, lm, rm, mode, out, out_mask);
    mnr_dtype ot;
    const mnr_dtype l = lhs->dtype, r = rhs->dtype;
    if ((l == MNR_I32 && r == MNR_F64) || (l == MNR_F64 && r == MNR_I32)) ot = MNR_F64;
    else if ((l == MNR_I32 && r == MNR_F32) || (l == MNR_F32 && r == MNR_I32)) ot = MNR_F32;
    else return fail(MNR_ERR_UNSUPPORTED_TYPE, "Unsupported array type combination for arithmetic operations (dtypes %d, %d)",
                     (int)l, (int)r);
    int rc = check_masks(lm, rm, lhs->len);
    if (rc) return rc;
    rc = alloc_outputs(c, ot, lhs->len, lm || rm, out, out_mask);
    if (rc) return rc;
    EwArgs a{};
    a.dtype = ot; a.op = op; a.lhs = lhs->ptr; a.rhs = rhs->ptr;
    a.lmask = lm ? lm->ptr : nullptr; a.rmask = rm ? rm->ptr : nullptr; a.mask_or = mode == MNR_MASK_OR;
    a.out = (*out)->ptr; a.out_mask = (lm || rm) ? (*out_mask)->ptr : nullptr; a.n = lhs->len; a.k = c->knobs;
    return drop_outputs(run_ew(c, a, c->stream, true, l, r), out, out_mask);
}

// ---- batched element-wise fan-out -------------------------------------------------------------------------------------
static int ensure_ew_segs(mnr_ctx* c, size_t bytes) {
    if (c->ew_segs_bytes < bytes * 2) {
        if (c->ew_segs) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->ew_segs)); c->ew_segs = nullptr; }
        CU(cudaMalloc(&c->ew_segs, bytes * 2));
        c->ew_segs_bytes = bytes * 2;
    }
    return MNR_OK;
}

// items: fully validated EwArgs (one per chunk).  Aligned items of one (dtype, masked, tier) class share a launch;
// the rest go one by one.  Dense integer division keeps the reference's panic as ONE error for the whole batch.
static int run_ew_batch(mnr_ctx* c, std::vector<EwArgs>& items) {
    if (items.empty()) return MNR_OK;
    CU(cudaSetDevice(c->device));
    unsigned int* flag = c->ticket[0] + 8;
    bool any_dense_int_div = false;
    for (auto& a : items) {
        a.div0_flag = flag;
        prepare_scalar_division(a);
        if (!a.lmask && !a.rmask && !is_float_dtype(a.dtype) && !a.sdiv && !a.nonzero_divisor &&
            (a.op == MNR_DIV || a.op == MNR_REM || a.op == MNR_FLOORDIV))
            any_dense_int_div = true;
    }
    if (any_dense_int_div) CU(cudaMemsetAsync(flag, 0, 4, c->stream));
    struct Key { int dtype, tier, masked, sdiv; };
    std::vector<Key> keys;
    std::vector<std::vector<EwDev>> groups;
    std::vector<uint64_t> max_n;
    for (const auto& a : items) {
        if (a.n == 0) continue;
        const int tier = ew_batch_tier(a.dtype, a.op, a.sdiv != 0, a.lhs, a.rhs, a.out, c->knobs.max_tier);
        if (tier == 0) {
            CU(launch_ew_binary(a, c->stream));
            c->launches++;
            continue;
        }
        const Key k{(int)a.dtype, tier, (a.lmask || a.rmask) ? 1 : 0, a.sdiv};
        size_t g = 0;
        for (; g < keys.size(); ++g) if (keys[g].dtype == k.dtype && keys[g].tier == k.tier && keys[g].masked == k.masked && keys[g].sdiv == k.sdiv) break;
        if (g == keys.size()) { keys.push_back(k); groups.emplace_back(); max_n.push_back(0); }
        EwDev d;
        d.lhs = a.lhs; d.rhs = a.rhs; d.scalar_bits = a.scalar_bits; d.lmask = a.lmask; d.rmask = a.rmask; d.mask_or = a.mask_or;
        d.out = a.out; d.out_mask = a.out_mask; d.n = a.n; d.div0_flag = a.div0_flag; d.op = a.op;
        d.sdiv = a.sdiv; d.magic.m = a.magic_m; d.magic.s1 = a.magic_s1; d.magic.s2 = a.magic_s2;
        groups[g].push_back(d);
        max_n[g] = std::max<uint64_t>(max_n[g], a.n);
    }
    for (size_t g = 0; g < groups.size(); ++g) {
        for (size_t off = 0; off < groups[g].size(); off += 65535) {
            const size_t cnt = std::min<size_t>(65535, groups[g].size() - off);
            int rc = ensure_ew_segs(c, std::min<size_t>(65535, groups[g].size()) * sizeof(EwDev));
            if (rc) return rc;
            char* dst = static_cast<char*>(c->ew_segs) + (c->ew_flip ? c->ew_segs_bytes / 2 : 0);
            c->ew_flip ^= 1;
            CU(cudaMemcpyAsync(dst, groups[g].data() + off, cnt * sizeof(EwDev), cudaMemcpyHostToDevice, c->stream));
            CU(launch_ew_batch((mnr_dtype)keys[g].dtype, items[0].op, keys[g].tier, keys[g].masked != 0, keys[g].sdiv != 0,
                               reinterpret_cast<const EwDev*>(dst), (uint32_t)cnt, max_n[g], c->knobs.grid_cap, c->stream));
            c->launches++;
        }
    }
    if (any_dense_int_div) {
        unsigned int* h = static_cast<unsigned int*>(c->h_scratch);
        CU(cudaMemcpyAsync(h, flag, 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (*h) return fail(MNR_ERR_DIVIDE_BY_ZERO, "division by zero in dense integer kernel (the reference panics here)");
    }
    return MNR_OK;
}

int mnr_ew_binary_batch_into(mnr_ctx* c, mnr_op op, size_t n, const mnr_buf* const* lhs, const mnr_buf* const* rhs,
                             const mnr_bits* const* lmask, const mnr_bits* const* rmask, mnr_mask_mode mode,
                             mnr_buf* const* out, mnr_bits* const* out_mask) {
    REQUIRE(c && (n == 0 || (lhs && rhs && out)), MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    REQUIRE(op >= MNR_ADD && op <= MNR_FLOORDIV, MNR_ERR_OPERATOR_MISMATCH, "unknown operator %d", (int)op);
    std::vector<EwArgs> items(n);
    for (size_t i = 0; i < n; ++i) {
        const mnr_buf *l = lhs[i], *r = rhs[i];
        const mnr_bits* lm = lmask ? lmask[i] : nullptr;
        const mnr_bits* rm = rmask ? rmask[i] : nullptr;
        mnr_bits* om = out_mask ? out_mask[i] : nullptr;
        REQUIRE(l && r && out[i], MNR_ERR_INVALID_ARGUMENTS, "chunk %zu: NULL buffer", i);
        // per-chunk length check of the SuperArray route (broadcast/super_array.rs:203-213)
        REQUIRE(l->len == r->len, MNR_ERR_LENGTH_MISMATCH, "chunk %zu: length mismatch (lhs: %zu, rhs: %zu)", i, l->len, r->len);
        REQUIRE(l->dtype == r->dtype, MNR_ERR_UNSUPPORTED_TYPE,
                "chunk %zu: Unsupported array type combination for arithmetic operations (dtypes %d, %d)", i, (int)l->dtype, (int)r->dtype);
        REQUIRE(out[i]->dtype == l->dtype && out[i]->len == l->len, MNR_ERR_INVALID_ARGUMENTS, "chunk %zu: output shape/dtype mismatch", i);
        const bool masked = lm || rm;
        REQUIRE(!masked || (om && om->len == l->len), MNR_ERR_INVALID_ARGUMENTS, "chunk %zu: masked call needs an output mask of %zu bits", i, l->len);
        int rc = check_masks(lm, rm, l->len);
        if (rc) return rc;
        rc = check_no_alias(out[i], masked ? om : nullptr, {l, r}, {lm, rm});   // chunk i's outputs against chunk i's inputs
        if (rc) return rc;
        EwArgs& a = items[i];
        a = EwArgs{};
        a.dtype = l->dtype; a.op = op; a.lhs = l->ptr; a.rhs = r->ptr;
        a.lmask = lm ? lm->ptr : nullptr; a.rmask = rm ? rm->ptr : nullptr; a.mask_or = mode == MNR_MASK_OR;
        a.out = out[i]->ptr; a.out_mask = masked ? om->ptr : nullptr; a.n = l->len; a.k = c->knobs;
    }
    return run_ew_batch(c, items);
}

int mnr_ew_scalar_batch_into(mnr_ctx* c, mnr_op op, size_t n, const mnr_buf* const* arrs, const void* const* scalars,
                             int scalar_is_lhs, const mnr_bits* const* masks, mnr_buf* const* out, mnr_bits* const* out_mask) {
    REQUIRE(c && (n == 0 || (arrs && scalars && out)), MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    REQUIRE(op >= MNR_ADD && op <= MNR_FLOORDIV, MNR_ERR_OPERATOR_MISMATCH, "unknown operator %d", (int)op);
    std::vector<EwArgs> items(n);
    for (size_t i = 0; i < n; ++i) {
        const mnr_buf* arr = arrs[i];
        const mnr_bits* m = masks ? masks[i] : nullptr;
        mnr_bits* om = out_mask ? out_mask[i] : nullptr;
        REQUIRE(arr && scalars[i] && out[i], MNR_ERR_INVALID_ARGUMENTS, "chunk %zu: NULL argument", i);
        REQUIRE(out[i]->dtype == arr->dtype && out[i]->len == arr->len, MNR_ERR_INVALID_ARGUMENTS, "chunk %zu: output shape/dtype mismatch", i);
        REQUIRE(!m || (om && om->len == arr->len), MNR_ERR_INVALID_ARGUMENTS, "chunk %zu: masked call needs an output mask of %zu bits", i, arr->len);
        int rc = check_masks(m, nullptr, arr->len);
        if (rc) return rc;
        rc = check_no_alias(out[i], m ? om : nullptr, {arr}, {m});
        if (rc) return rc;
        EwArgs& a = items[i];
        a = EwArgs{};
        a.dtype = arr->dtype; a.op = op;
        a.lhs = scalar_is_lhs ? nullptr : arr->ptr;
        a.rhs = scalar_is_lhs ? arr->ptr : nullptr;
        a.scalar_bits = scalar_to_bits(arr->dtype, scalars[i]);
        a.lmask = m ? m->ptr : nullptr; a.out = out[i]->ptr; a.out_mask = m ? om->ptr : nullptr; a.n = arr->len; a.k = c->knobs;
    }
    return run_ew_batch(c, items);
}

// Fresh outputs for every chunk (the reference returns a new SuperArray), then the batched launch.
int mnr_ew_binary_batch(mnr_ctx* c, mnr_op op, size_t n, const mnr_buf* const* lhs, const mnr_buf* const* rhs,
                        const mnr_bits* const* lmask, const mnr_bits* const* rmask, mnr_mask_mode mode, mnr_buf** out,
                        mnr_bits** out_mask) {
    REQUIRE(c && (n == 0 || (lhs && rhs && out && out_mask)), MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    for (size_t i = 0; i < n; ++i) { out[i] = nullptr; out_mask[i] = nullptr; }
    int rc = MNR_OK;
    for (size_t i = 0; i < n && !rc; ++i) {
        if (!lhs[i] || !rhs[i]) { rc = fail(MNR_ERR_INVALID_ARGUMENTS, "chunk %zu: NULL buffer", i); break; }
        if (lhs[i]->len != rhs[i]->len) {
            rc = fail(MNR_ERR_LENGTH_MISMATCH, "chunk %zu: length mismatch (lhs: %zu, rhs: %zu)", i, lhs[i]->len, rhs[i]->len);
            break;
        }
        const bool masked = (lmask && lmask[i]) || (rmask && rmask[i]);
        rc = alloc_outputs(c, lhs[i]->dtype, lhs[i]->len, masked, &out[i], &out_mask[i]);
    }
    if (!rc) rc = mnr_ew_binary_batch_into(c, op, n, lhs, rhs, lmask, rmask, mode, out, out_mask);
    if (rc) for (size_t i = 0; i < n; ++i) { mnr_buf_free(out[i]); mnr_bits_free(out_mask[i]); out[i] = nullptr; out_mask[i] = nullptr; }
    return rc;
}

// ---- bitmask kernels ------------------------------------------------------------------------------------------------
static int check_window(const mnr_bits* m, uint64_t pos, size_t len, const char* what) {
    REQUIRE((pos >> 3) + mask_bytes(len) <= mask_bytes(m->len) || len == 0, MNR_ERR_OUT_OF_BOUNDS,
            "%s: window [%llu, +%zu) leaves the %zu-bit mask", what, (unsigned long long)pos, len, m->len);
    return MNR_OK;
}

int mnr_bits_binop_into(mnr_ctx* c, mnr_logical_op op, const mnr_bits* lhs, size_t lo, const mnr_bits* rhs, size_t ro,
                        size_t len, mnr_bits* out) {
    REQUIRE(c && lhs && rhs && out, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    REQUIRE(op >= MNR_AND && op <= MNR_XOR, MNR_ERR_OPERATOR_MISMATCH, "unknown logical operator %d", (int)op);
    REQUIRE(out->len == len, MNR_ERR_INVALID_ARGUMENTS, "output mask has %zu bits, need %zu", out->len, len);
    const uint64_t lp = (uint64_t)(lo / 8) * 8, rp = (uint64_t)(ro / 8) * 8;   // bitmask_window_bytes floors to bytes
    int rc = check_window(lhs, lp, len, "bitmask_binop lhs");
    if (rc) return rc;
    rc = check_window(rhs, rp, len, "bitmask_binop rhs");
    if (rc) return rc;
    rc = check_no_alias(nullptr, out, {}, {lhs, rhs});
    if (rc) return rc;
    if (len == 0) return MNR_OK;
    CU(cudaSetDevice(c->device));
    CU(launch_bits_op((int)op, lhs->ptr, lp, lhs->len, rhs->ptr, rp, rhs->len, len, out->ptr, c->stream));
    c->launches++;
    return MNR_OK;
}

int mnr_bits_binop(mnr_ctx* c, mnr_logical_op op, const mnr_bits* lhs, size_t lo, const mnr_bits* rhs, size_t ro,
                   size_t len, mnr_bits** out) {
    REQUIRE(c && out, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    *out = nullptr;
    int rc = mnr_bits_alloc(c, len, out);
    if (rc) return rc;
    rc = mnr_bits_binop_into(c, op, lhs, lo, rhs, ro, len, *out);
    if (rc) { mnr_bits_free(*out); *out = nullptr; }
    return rc;
}

int mnr_bits_not_into(mnr_ctx* c, const mnr_bits* src, size_t off, size_t len, mnr_bits* out) {
    REQUIRE(c && src && out, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    REQUIRE(out->len == len, MNR_ERR_INVALID_ARGUMENTS, "output mask has %zu bits, need %zu", out->len, len);
    const uint64_t sp = (uint64_t)(off / 8) * 8;
    int rc = check_window(src, sp, len, "bitmask_unop");
    if (rc) return rc;
    rc = check_no_alias(nullptr, out, {}, {src});
    if (rc) return rc;
    if (len == 0) return MNR_OK;
    CU(cudaSetDevice(c->device));
    CU(launch_bits_op(4, src->ptr, sp, src->len, nullptr, 0, 0, len, out->ptr, c->stream));
    c->launches++;
    return MNR_OK;
}

int mnr_bits_not(mnr_ctx* c, const mnr_bits* src, size_t off, size_t len, mnr_bits** out) {
    REQUIRE(c && out, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    *out = nullptr;
    int rc = mnr_bits_alloc(c, len, out);
    if (rc) return rc;
    rc = mnr_bits_not_into(c, src, off, len, *out);
    if (rc) { mnr_bits_free(*out); *out = nullptr; }
    return rc;
}

static int popcount_sync(mnr_ctx* c, const mnr_bits* a, uint64_t ap, const mnr_bits* b, uint64_t bp, size_t len,
                         uint64_t* ones) {
    CU(cudaSetDevice(c->device));
    if (len == 0) { *ones = 0; return MNR_OK; }
    // The kernel's last block stores the count straight into mapped pinned host memory with one 8-byte store (a single PCIe
    // write).  A count is at most `len`, so all-ones cannot be a result: the slot is armed with it and polled — the call
    // returns when the count lands instead of when the stream has drained (bounded; then cudaStreamSynchronize as before).
    volatile unsigned long long* h = reinterpret_cast<volatile unsigned long long*>(static_cast<char*>(c->h_scratch) + 64);
    *h = ~0ull;
    CU(launch_bits_popcount(a->ptr, ap, a->len, b ? b->ptr : nullptr, bp, b ? b->len : 0, len,
                            reinterpret_cast<unsigned long long*>(c->partials[3]), c->ticket[3] + 4, c->d_count,
                            const_cast<unsigned long long*>(h), c->stream));
    c->launches++;
    for (int spin = 0; spin < 100000; ++spin) {
        const unsigned long long r = *h;
        if (r != ~0ull) { *ones = r; return MNR_OK; }
    }
    CU(cudaStreamSynchronize(c->stream));
    REQUIRE(*h != ~0ull, MNR_ERR_CUDA, "popcount finished without a result");
    *ones = *h;
    return MNR_OK;
}

int mnr_bits_popcount(mnr_ctx* c, const mnr_bits* m, size_t off, size_t len, uint64_t* ones) {
    REQUIRE(c && m && ones, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    const uint64_t p = (uint64_t)(off / 64) * 64;   // popcount_mask_simd: word_start = offset / 64
    int rc = check_window(m, p, len, "popcount_mask");
    if (rc) return rc;
    return popcount_sync(c, m, p, nullptr, 0, len, ones);
}

int mnr_bits_popcount_async(mnr_ctx* c, const mnr_bits* m, size_t off, size_t len, void* out_device) {
    REQUIRE(c && m && out_device, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    REQUIRE((reinterpret_cast<uintptr_t>(out_device) & 7u) == 0, MNR_ERR_INVALID_ARGUMENTS, "out_device must be 8-byte aligned");
    const uint64_t p = (uint64_t)(off / 64) * 64;
    int rc = check_window(m, p, len, "popcount_mask");
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    if (len == 0) { CU(cudaMemsetAsync(out_device, 0, 8, c->stream)); return MNR_OK; }
    CU(launch_bits_popcount(m->ptr, p, m->len, nullptr, 0, 0, len, reinterpret_cast<unsigned long long*>(c->partials[3]),
                            c->ticket[3] + 4, static_cast<unsigned long long*>(out_device), nullptr, c->stream));
    c->launches++;
    return MNR_OK;
}

int mnr_bits_all_true(mnr_ctx* c, const mnr_bits* m, int* out) {
    REQUIRE(c && m && out, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    uint64_t ones = 0;
    int rc = popcount_sync(c, m, 0, nullptr, 0, m->len, &ones);
    if (rc) return rc;
    *out = ones == m->len;
    return MNR_OK;
}
int mnr_bits_all_false(mnr_ctx* c, const mnr_bits* m, int* out) {
    REQUIRE(c && m && out, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    uint64_t ones = 0;
    int rc = popcount_sync(c, m, 0, nullptr, 0, m->len, &ones);
    if (rc) return rc;
    *out = ones == 0;
    return MNR_OK;
}

int mnr_bits_merge(mnr_ctx* c, const mnr_bits* lhs, const mnr_bits* rhs, size_t len, mnr_mask_mode mode, mnr_bits** out) {
    REQUIRE(c && out, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    *out = nullptr;
    if (!lhs && !rhs) return MNR_OK;
    int rc = check_masks(lhs, rhs, len);
    if (rc) return rc;
    rc = mnr_bits_alloc(c, len, out);
    if (rc) return rc;
    if (len == 0) return MNR_OK;
    CU(cudaSetDevice(c->device));
    if (lhs && rhs)
        CU(launch_bits_op(mode == MNR_MASK_OR ? 1 : 0, lhs->ptr, 0, lhs->len, rhs->ptr, 0, rhs->len, len, (*out)->ptr, c->stream));
    else {
        const mnr_bits* m = lhs ? lhs : rhs;
        CU(launch_bits_op(5, m->ptr, 0, m->len, nullptr, 0, 0, len, (*out)->ptr, c->stream));
    }
    c->launches++;
    return MNR_OK;
}

int mnr_bits_eq(mnr_ctx* c, const mnr_bits* a, size_t ao, const mnr_bits* b, size_t bo, size_t len, int negate,
                mnr_bits** out) {
    REQUIRE(c && a && b && out, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    *out = nullptr;
    if (len == 0) return mnr_bits_alloc(c, 0, out);
    REQUIRE(ao % 64 == 0 && bo % 64 == 0, MNR_ERR_INVALID_ARGUMENTS,
            "eq_bits_mask: offsets must be 64-bit aligned (got a: %zu, b: %zu)", ao, bo);
    int rc = check_window(a, ao, len, "eq_mask a");
    if (rc) return rc;
    rc = check_window(b, bo, len, "eq_mask b");
    if (rc) return rc;
    rc = mnr_bits_alloc(c, len, out);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    CU(launch_bits_op(negate ? 2 : 3, a->ptr, ao, a->len, b->ptr, bo, b->len, len, (*out)->ptr, c->stream));
    c->launches++;
    return MNR_OK;
}

int mnr_bits_all_eq(mnr_ctx* c, const mnr_bits* a, size_t ao, const mnr_bits* b, size_t bo, size_t len, int* out) {
    REQUIRE(c && a && b && out, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    if (len == 0) { *out = 1; return MNR_OK; }
    REQUIRE(len < 64 || (ao % 64 == 0 && bo % 64 == 0), MNR_ERR_INVALID_ARGUMENTS,
            "all_eq_mask_simd: offsets must be 64-bit aligned (got a: %zu, b: %zu)", ao, bo);
    const uint64_t ap = (uint64_t)(ao / 64) * 64, bp = (uint64_t)(bo / 64) * 64;
    int rc = check_window(a, ap, len, "all_eq a");
    if (rc) return rc;
    rc = check_window(b, bp, len, "all_eq b");
    if (rc) return rc;
    uint64_t diff = 0;
    rc = popcount_sync(c, a, ap, b, bp, len, &diff);
    if (rc) return rc;
    *out = diff == 0;
    return MNR_OK;
}

int mnr_bits_in(mnr_ctx* c, const mnr_bits* lhs, size_t lo, const mnr_bits* rhs, size_t ro, size_t len, int negate,
                mnr_bits** out) {
    REQUIRE(c && lhs && rhs && out, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    *out = nullptr;
    if (len == 0) return mnr_bits_alloc(c, 0, out);
    const uint64_t rp = (uint64_t)(ro / 64) * 64;
    int rc = check_window(rhs, rp, len, "in_mask rhs");
    if (rc) return rc;
    uint64_t ones = 0;
    rc = popcount_sync(c, rhs, rp, nullptr, 0, len, &ones);
    if (rc) return rc;
    const bool has_true = ones != 0, has_false = ones != len;
    mnr_bits* r = nullptr;
    if (has_true && has_false) rc = mnr_bits_new_set_all(c, len, 1, &r);
    else if (has_true) {   // lhs.slice_clone(lhs_off, len): exact bit offset (bitmask.rs:604-626)
        // exact bit offset here (not floored): the window must end inside the mask bit-exactly
        rc = (lo <= lhs->len && len <= lhs->len - lo) ? MNR_OK
                 : fail(MNR_ERR_OUT_OF_BOUNDS, "in_mask lhs: window [%zu, +%zu) leaves the %zu-bit mask", lo, len, lhs->len);
        if (!rc) rc = mnr_bits_alloc(c, len, &r);
        if (!rc) {
            CU(launch_bits_op(5, lhs->ptr, lo, lhs->len, nullptr, 0, 0, len, r->ptr, c->stream));
            c->launches++;
        }
    } else rc = mnr_bits_not(c, lhs, lo, len, &r);
    if (rc) { mnr_bits_free(r); return rc; }
    if (negate) {   // not_in_mask = not_mask(in_mask)
        mnr_bits* nr = nullptr;
        rc = mnr_bits_not(c, r, 0, len, &nr);
        mnr_bits_free(r);
        if (rc) return rc;
        r = nr;
    }
    *out = r;
    return MNR_OK;
}

int mnr_bits_slice(mnr_ctx* c, const mnr_bits* src, size_t offset, size_t len, mnr_bits** out) {
    REQUIRE(c && src && out, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    *out = nullptr;
    REQUIRE(offset <= src->len && len <= src->len - offset, MNR_ERR_OUT_OF_BOUNDS, "slice [%zu, %zu) out of bounds for %zu bits",
            offset, offset + len, src->len);
    int rc = mnr_bits_alloc(c, len, out);
    if (rc || len == 0) return rc;
    CU(cudaSetDevice(c->device));
    CU(launch_bits_op(5, src->ptr, offset, src->len, nullptr, 0, 0, len, (*out)->ptr, c->stream));
    c->launches++;
    return MNR_OK;
}

int mnr_concat(mnr_ctx* c, size_t n, const mnr_buf* const* bufs, const mnr_bits* const* validities, mnr_buf** out,
               mnr_bits** out_validity) {
    REQUIRE(c && out && out_validity && (bufs || n == 0), MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    *out = nullptr; *out_validity = nullptr;
    REQUIRE(n >= 1, MNR_ERR_INVALID_ARGUMENTS, "concat needs at least one chunk (the dtype comes from the chunks)");
    std::vector<ConcatSeg> segs(n);
    uint64_t total = 0, max_rows = 0;
    bool any_mask = false;
    for (size_t i = 0; i < n; ++i) {
        REQUIRE(bufs[i], MNR_ERR_INVALID_ARGUMENTS, "chunk %zu is NULL", i);
        REQUIRE(bufs[i]->dtype == bufs[0]->dtype, MNR_ERR_TYPE_MISMATCH, "chunk %zu has dtype %d, chunk 0 has %d", i,
                (int)bufs[i]->dtype, (int)bufs[0]->dtype);
        const mnr_bits* v = validities ? validities[i] : nullptr;
        REQUIRE(!v || v->len >= bufs[i]->len, MNR_ERR_INVALID_ARGUMENTS, "chunk %zu: validity has %zu bits, need %zu", i, v->len, bufs[i]->len);
        segs[i] = ConcatSeg{bufs[i]->ptr, v ? v->ptr : nullptr, bufs[i]->len, total};
        total += bufs[i]->len;
        max_rows = std::max<uint64_t>(max_rows, bufs[i]->len);
        any_mask |= v != nullptr;
    }
    int rc = alloc_outputs(c, bufs[0]->dtype, total, any_mask, out, out_validity);
    if (rc) return rc;
    if (total == 0) return MNR_OK;
    CU(cudaSetDevice(c->device));
    rc = ensure_ew_segs(c, n * sizeof(ConcatSeg));
    if (rc) return drop_outputs(rc, out, out_validity);
    char* dst = static_cast<char*>(c->ew_segs) + (c->ew_flip ? c->ew_segs_bytes / 2 : 0);
    c->ew_flip ^= 1;
    CU(cudaMemcpyAsync(dst, segs.data(), n * sizeof(ConcatSeg), cudaMemcpyHostToDevice, c->stream));
    CU(launch_concat((int)dtype_size(bufs[0]->dtype), reinterpret_cast<const ConcatSeg*>(dst), (uint32_t)n, max_rows, total,
                     (*out)->ptr, any_mask ? (*out_validity)->ptr : nullptr, c->stream));
    c->launches += any_mask ? 2 : 1;
    return MNR_OK;
}

int mnr_eq_mask(mnr_ctx* c, const mnr_buf* data, const void* field_mask, const void* target, mnr_bits** out) {
    REQUIRE(c && data && field_mask && target && out, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    REQUIRE(!is_float_dtype(data->dtype), MNR_ERR_UNSUPPORTED_TYPE, "eq_mask works on integer lanes (u8/u16/u32/u64 and their signed twins)");
    *out = nullptr;
    int rc = mnr_bits_alloc(c, data->len, out);
    if (rc) return rc;
    if (data->len == 0) return MNR_OK;
    CU(cudaSetDevice(c->device));
    CU(launch_eq_mask((int)dtype_size(data->dtype), data->ptr, data->len, scalar_to_bits(data->dtype, field_mask),
                      scalar_to_bits(data->dtype, target), (*out)->ptr, c->stream));
    c->launches++;
    return MNR_OK;
}

// ---- reductions ------------------------------------------------------------------------------------------------------
static int check_reduce(const mnr_ctx* c, const mnr_buf* b, const mnr_bits* v) {
    REQUIRE(c && b, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    REQUIRE(!v || v->len >= b->len, MNR_ERR_INVALID_ARGUMENTS, "validity has %zu bits, need %zu", v->len, b->len);
    return MNR_OK;
}

int mnr_reduce_stats_async(mnr_ctx* c, const mnr_buf* b, const mnr_bits* v, int with_minmax, void* out_device) {
    int rc = check_reduce(c, b, v);
    if (rc) return rc;
    REQUIRE(out_device && (reinterpret_cast<uintptr_t>(out_device) & 15u) == 0, MNR_ERR_INVALID_ARGUMENTS,
            "out_device must be a 16-byte aligned device pointer");
    CU(cudaSetDevice(c->device));
    CU(launch_reduce_stats(b->dtype, b->ptr, v ? v->ptr : nullptr, b->len, with_minmax != 0, c->partials[3], c->ticket[3],
                           static_cast<AggRaw*>(out_device), nullptr, c->stream));
    c->launches++;
    return MNR_OK;
}

// Synchronous form: the kernel's finishing block also stores the aggregate into mapped pinned host memory, so the call
// is one launch + one stream synchronise (no D2H memcpy).
// Synchronous reduction: the kernel stores the aggregate into mapped pinned host memory, each 32-bit word framed with this
// call's sequence number (reduce_kernels.cuh, host_seq), and the host polls the slot instead of waiting for the stream to
// drain — on an 8 KB column (BASELINE configs[0]) the stream synchronisation was the larger half of the call.  The poll is
// bounded: after ~50 us without the result the call falls back to cudaStreamSynchronize (long reductions; a faulted kernel
// reports its error there).
static bool host_result_ready(const volatile uint64_t* slot, uint32_t seq, mnr_agg* out) {
    uint32_t w[8];
    for (int i = 0; i < 8; ++i) {
        const uint64_t x = slot[i];
        if ((uint32_t)(x >> 32) != seq) return false;
        w[i] = (uint32_t)x;
    }
    memcpy(out, w, sizeof(mnr_agg));
    return true;
}
static volatile uint64_t* host_result_slot(mnr_ctx* c) {
    return reinterpret_cast<volatile uint64_t*>(static_cast<char*>(c->h_scratch) + 192);
}
static uint32_t next_host_seq(mnr_ctx* c) {
    c->host_seq = (c->host_seq + 1) & 0xFFFFFFu;
    if (c->host_seq == 0) c->host_seq = 1;
    return c->host_seq;
}
static int await_host_result(mnr_ctx* c, cudaStream_t s, uint32_t seq, mnr_agg* out) {
    const volatile uint64_t* slot = host_result_slot(c);
    for (int spin = 0; spin < 20000; ++spin)
        if (host_result_ready(slot, seq, out)) return MNR_OK;
    CU(cudaStreamSynchronize(s));
    REQUIRE(host_result_ready(slot, seq, out), MNR_ERR_CUDA, "reduction finished without a result");
    return MNR_OK;
}
static int reduce_sync(mnr_ctx* c, const mnr_buf* b, const mnr_bits* v, bool minmax, mnr_agg* out) {
    static_assert(sizeof(mnr_agg) == 32, "eight 32-bit words");
    int rc = check_reduce(c, b, v);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    const uint32_t seq = next_host_seq(c);
    volatile uint64_t* slot = host_result_slot(c);
    CU(launch_reduce_stats(b->dtype, b->ptr, v ? v->ptr : nullptr, b->len, minmax, c->partials[3], c->ticket[3], c->d_agg,
                           reinterpret_cast<AggRaw*>(const_cast<uint64_t*>(slot)), c->stream, seq));
    c->launches++;
    return await_host_result(c, c->stream, seq, out);
}

int mnr_reduce_stats(mnr_ctx* c, const mnr_buf* b, const mnr_bits* v, mnr_agg* out_host) {
    REQUIRE(out_host, MNR_ERR_INVALID_ARGUMENTS, "out is NULL");
    return reduce_sync(c, b, v, true, out_host);
}

int mnr_reduce_sum(mnr_ctx* c, const mnr_buf* b, const mnr_bits* v, mnr_scalar64* out_sum, uint64_t* out_count) {
    REQUIRE(out_sum, MNR_ERR_INVALID_ARGUMENTS, "out_sum is NULL");
    mnr_agg a;
    int rc = reduce_sync(c, b, v, false, &a);
    if (rc) return rc;
    *out_sum = a.sum;
    if (out_count) *out_count = a.count;
    return MNR_OK;
}


// ---- batched reductions: one launch per (dtype, alignment tier, masked) class -----------------------------------------
static int ensure_batch_scratch(mnr_ctx* c, size_t nseg_total, size_t nseg_launch, size_t partial_slots) {
    const size_t need_p = partial_slots * sizeof(AggRaw), need_s = nseg_total * sizeof(ReduceSeg), need_t = nseg_launch * sizeof(unsigned int);
    if (c->batch_partials_bytes < need_p) {
        if (c->batch_partials) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->batch_partials)); c->batch_partials = nullptr; c->batch_partials_bytes = 0; }
        CU(cudaMalloc(&c->batch_partials, need_p));
        c->batch_partials_bytes = need_p;
    }
    if (c->batch_segs_bytes < need_s * 2) {
        if (c->batch_segs) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->batch_segs)); c->batch_segs = nullptr; c->batch_segs_bytes = 0; }
        CU(cudaMalloc(&c->batch_segs, need_s * 2));
        c->batch_segs_bytes = need_s * 2;
    }
    if (c->batch_tickets_bytes < need_t) {
        if (c->batch_tickets) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->batch_tickets)); c->batch_tickets = nullptr; c->batch_tickets_bytes = 0; }
        CU(cudaMalloc(&c->batch_tickets, need_t * 2));
        CU(cudaMemset(c->batch_tickets, 0, need_t * 2));   // tickets re-arm themselves after every launch
        c->batch_tickets_bytes = need_t * 2;
    }
    return MNR_OK;
}

// Launch plan of a batched reduction: the segments grouped by kernel instantiation (order inside a group is the caller's
// order), flattened in launch order with gridDim.y <= 65535 per launch.
struct ReduceLaunch { int dtype, tier, masked; size_t cnt; uint32_t max_blk; size_t seg0; size_t part0; };
struct ReducePlan {
    std::vector<ReduceLaunch> launches;
    std::vector<ReduceSeg> flat;
    size_t pblk_total = 0;   // partial slots over all launches of the call (every launch has its own region: they overlap)
};
static ReducePlan plan_reduce_batch(size_t n, const mnr_buf* const* bufs, const mnr_bits* const* validities, bool minmax) {
    struct Key { int dtype, tier, masked; };
    std::vector<Key> keys;
    std::vector<std::vector<ReduceSeg>> groups;
    for (size_t i = 0; i < n; ++i) {
        const mnr_buf* b = bufs[i];
        const mnr_bits* v = validities ? validities[i] : nullptr;
        const Key k{(int)b->dtype, reduce_tier(b->ptr, minmax), v ? 1 : 0};
        size_t g = 0;
        for (; g < keys.size(); ++g) if (keys[g].dtype == k.dtype && keys[g].tier == k.tier && keys[g].masked == k.masked) break;
        if (g == keys.size()) { keys.push_back(k); groups.emplace_back(); }
        ReduceSeg s;
        s.data = b->ptr; s.mask = v ? v->ptr : nullptr; s.n = b->len;
        s.nblk = reduce_nblk(b->dtype, b->len, k.tier, minmax); s.out_index = (uint32_t)i;
        groups[g].push_back(s);
    }
    ReducePlan p;
    for (size_t g = 0; g < groups.size(); ++g)
        for (size_t off = 0; off < groups[g].size(); off += 65535) {
            const size_t cnt = std::min<size_t>(65535, groups[g].size() - off);
            uint32_t max_blk = 1;
            for (size_t i = 0; i < cnt; ++i) max_blk = std::max(max_blk, groups[g][off + i].nblk);
            p.launches.push_back(ReduceLaunch{keys[g].dtype, keys[g].tier, keys[g].masked, cnt, max_blk, p.flat.size(), p.pblk_total});
            p.flat.insert(p.flat.end(), groups[g].begin() + off, groups[g].begin() + off + cnt);
            p.pblk_total += cnt * max_blk;
        }
    return p;
}

// Batched launch(es) over validated chunks.  `f` / `x`: optional second stage (per-column fold + cross-GPU exchange) run
// by the block that writes the last chunk aggregate of the call.  One descriptor upload for the whole call, so the
// launches sit back to back on the stream.  The descriptor area is double-buffered; a pageable-source cudaMemcpyAsync
// stages the bytes before it returns, so the host vector may die right after; stream order protects the device copy.
// The launches of one call carry the programmatic-launch attribute and are independent of each other (own partials region,
// own tickets — one per segment of the call, self re-arming), so launch i+1 fills the SMs while launch i drains; the
// descriptor copy in front keeps the call itself ordered after everything earlier on the stream.
static int reduce_batch_launch(mnr_ctx* c, const ReducePlan& p, bool minmax, AggRaw* outs, const FoldArgs& f, const XchgDev& x) {
    int rc = ensure_batch_scratch(c, p.flat.size(), p.flat.size(), p.pblk_total);
    if (rc) return rc;
    char* dst = static_cast<char*>(c->batch_segs) + (c->batch_flip ? c->batch_segs_bytes / 2 : 0);
    c->batch_flip ^= 1;
    CU(cudaMemcpyAsync(dst, p.flat.data(), p.flat.size() * sizeof(ReduceSeg), cudaMemcpyHostToDevice, c->stream));
    for (const ReduceLaunch& L : p.launches) {
        CU(launch_reduce_stats_batch((mnr_dtype)L.dtype, L.tier, L.masked != 0, minmax,
                                     reinterpret_cast<const ReduceSeg*>(dst) + L.seg0, (uint32_t)L.cnt, L.max_blk,
                                     static_cast<AggRaw*>(c->batch_partials) + L.part0, static_cast<unsigned int*>(c->batch_tickets) + L.seg0,
                                     outs, f, x, c->stream));
        c->launches++;
    }
    return MNR_OK;
}

int mnr_reduce_stats_batch_async(mnr_ctx* c, size_t n, const mnr_buf* const* bufs, const mnr_bits* const* validities,
                                 int with_minmax, void* out_device) {
    REQUIRE(c && (bufs || n == 0) && (out_device || n == 0), MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    REQUIRE((reinterpret_cast<uintptr_t>(out_device) & 15u) == 0, MNR_ERR_INVALID_ARGUMENTS, "out_device must be 16-byte aligned");
    if (n == 0) return MNR_OK;
    for (size_t i = 0; i < n; ++i) {
        int rc = check_reduce(c, bufs[i], validities ? validities[i] : nullptr);
        if (rc) return rc;
    }
    CU(cudaSetDevice(c->device));
    return reduce_batch_launch(c, plan_reduce_batch(n, bufs, validities, with_minmax != 0), with_minmax != 0,
                               static_cast<AggRaw*>(out_device), FoldArgs{}, XchgDev{});
}

int mnr_reduce_stats_batch(mnr_ctx* c, size_t n, const mnr_buf* const* bufs, const mnr_bits* const* validities,
                           int with_minmax, mnr_agg* out_host) {
    REQUIRE(c && (out_host || n == 0), MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    if (n == 0) return MNR_OK;
    CU(cudaSetDevice(c->device));
    if (c->chunk_aggs_cap < n) {
        if (c->chunk_aggs) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->chunk_aggs)); }
        c->chunk_aggs = nullptr; c->chunk_aggs_cap = 0;
        CU(cudaMalloc(&c->chunk_aggs, sizeof(AggRaw) * n));
        c->chunk_aggs_cap = n;
    }
    int rc = mnr_reduce_stats_batch_async(c, n, bufs, validities, with_minmax, c->chunk_aggs);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out_host, c->chunk_aggs, sizeof(mnr_agg) * n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return MNR_OK;
}

// ---- fused reduction + cross-GPU exchange ---------------------------------------------------------------------------------
// One mailbox per rank; peers map it through CUDA IPC (one process per GPU) or plain peer access (one process, many GPUs).
static constexpr unsigned kSlotBytes = (kXchgHeader + 32u * MNR_XCHG_MAX_AGGS + 63u) & ~63u;
static constexpr size_t kMailboxBytes = (size_t)2 * kMaxPeers * kSlotBytes;

int mnr_xchg_create(mnr_ctx* c, int world, int rank, mnr_xchg** out) {
    REQUIRE(c && out, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, MNR_ERR_INVALID_ARGUMENTS,
            "world %d / rank %d out of range (max %d peers)", world, rank, kMaxPeers);
    CU(cudaSetDevice(c->device));
    // Kernels of an exchange wait on their peers: everything they may need must be resident before the first one runs.
    CU(reduce_preload_all());
    mnr_xchg* x = new mnr_xchg();
    x->ctx = c; x->world = world; x->rank = rank;
    cudaError_t e = cudaMalloc(&x->mailbox, kMailboxBytes);
    if (e == cudaSuccess) e = cudaMemset(x->mailbox, 0, kMailboxBytes);
    if (e == cudaSuccess) e = cudaMalloc(&x->err, 128);          // [0] error word, [64] the `done` epoch
    if (e == cudaSuccess) e = cudaMemset(x->err, 0, 128);
    if (e == cudaSuccess) e = cudaMalloc(&x->partials, 2 * sizeof(AggRaw) * (size_t)reduce_max_grid());
    if (e == cudaSuccess) e = cudaMalloc(&x->ticket, 128);
    if (e == cudaSuccess) e = cudaMemset(x->ticket, 0, 128);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        if (x->mailbox) cudaFree(x->mailbox);
        if (x->err) cudaFree(x->err);
        if (x->partials) cudaFree(x->partials);
        if (x->ticket) cudaFree(x->ticket);
        delete x;
        return fail_cuda(e, "mnr_xchg_create");
    }
    x->done = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(x->err) + 64);
    x->peers[rank] = x->mailbox;
    x->connected = world == 1;
    *out = x;
    return MNR_OK;
}

int mnr_xchg_local_handle(mnr_xchg* x, uint8_t* handle64) {
    REQUIRE(x && handle64, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == MNR_IPC_HANDLE_BYTES, "IPC handle size");
    CU(cudaSetDevice(x->ctx->device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, x->mailbox));
    memcpy(handle64, &h, sizeof h);
    return MNR_OK;
}

int mnr_xchg_connect(mnr_xchg* x, const uint8_t* handles) {
    REQUIRE(x && handles, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    CU(cudaSetDevice(x->ctx->device));
    for (int r = 0; r < x->world; ++r) {
        if (r == x->rank || x->opened[r] || x->peers[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * MNR_IPC_HANDLE_BYTES, sizeof h);
        void* p = nullptr;
        CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        x->peers[r] = static_cast<char*>(p);
        x->opened[r] = true;
    }
    x->connected = true;
    return MNR_OK;
}

int mnr_xchg_connect_local(mnr_xchg* x, mnr_xchg* const* peers) {
    REQUIRE(x && peers, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    CU(cudaSetDevice(x->ctx->device));
    for (int r = 0; r < x->world; ++r) {
        if (r == x->rank) continue;
        REQUIRE(peers[r] && peers[r]->world == x->world && peers[r]->rank == r, MNR_ERR_INVALID_ARGUMENTS,
                "peers[%d] is not rank %d of a %d-rank exchange", r, r, x->world);
        const int pd = peers[r]->ctx->device;
        if (pd == x->ctx->device) x->shares_device = true;   // two "virtual ranks" on one GPU: nothing to enable
        else {
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, x->ctx->device, pd));
            REQUIRE(can, MNR_ERR_CUDA, "device %d cannot access device %d (no P2P path)", x->ctx->device, pd);
            cudaError_t e = cudaDeviceEnablePeerAccess(pd, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return fail_cuda(e, "cudaDeviceEnablePeerAccess");
        }
        x->peers[r] = peers[r]->mailbox;
    }
    x->connected = true;
    return MNR_OK;
}

int mnr_xchg_status(mnr_xchg* x, int clear, int* timed_out) {
    REQUIRE(x && timed_out, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    CU(cudaSetDevice(x->ctx->device));
    unsigned int h = 0;
    CU(cudaMemcpyAsync(&h, x->err, 4, cudaMemcpyDeviceToHost, x->ctx->stream));
    CU(cudaStreamSynchronize(x->ctx->stream));
    if (h && clear) { CU(cudaMemsetAsync(x->err, 0, 4, x->ctx->stream)); CU(cudaStreamSynchronize(x->ctx->stream)); }
    *timed_out = h != 0;
    return MNR_OK;
}

void mnr_xchg_destroy(mnr_xchg* x) {
    if (!x) return;
    cudaSetDevice(x->ctx->device);
    cudaStreamSynchronize(x->ctx->stream);
    for (int r = 0; r < x->world; ++r) if (x->opened[r]) cudaIpcCloseMemHandle(x->peers[r]);
    cudaFree(x->mailbox);
    cudaFree(x->err);
    cudaFree(x->partials);
    cudaFree(x->ticket);
    delete x;
}

static int check_xchg(const mnr_ctx* c, const mnr_xchg* x) {
    REQUIRE(x->ctx == c, MNR_ERR_INVALID_ARGUMENTS, "exchange handle belongs to another context");
    REQUIRE(x->connected, MNR_ERR_INVALID_ARGUMENTS, "mnr_xchg_connect has not been called");
    return MNR_OK;
}
static XchgDev xchg_dev(const mnr_xchg* x, unsigned long long epoch) {
    XchgDev d{};
    d.world = x->world; d.rank = x->rank; d.epoch = epoch; d.err = x->err; d.done = x->done; d.slot_bytes = kSlotBytes;
    for (int r = 0; r < x->world; ++r) d.mailbox[r] = x->peers[r];
    return d;
}

int mnr_reduce_stats_exchange(mnr_ctx* c, mnr_xchg* x, const mnr_buf* b, const mnr_bits* v, int with_minmax, void* out_device) {
    int rc = check_reduce(c, b, v);
    if (rc) return rc;
    REQUIRE(x, MNR_ERR_INVALID_ARGUMENTS, "exchange handle is NULL");
    rc = check_xchg(c, x);
    if (rc) return rc;
    REQUIRE(out_device && (reinterpret_cast<uintptr_t>(out_device) & 15u) == 0, MNR_ERR_INVALID_ARGUMENTS,
            "out_device must be a 16-byte aligned device pointer");
    CU(cudaSetDevice(c->device));
    // Overlap with the previous reduction (late dependency wait) only when the caller opted in AND nothing else of this
    // library was launched on the stream since that reduction — an element-wise kernel could be producing this column.
    const bool late = c->reduce_overlap && c->last_reduce_launch == c->launches && c->launches != 0;
    const unsigned long long epoch = x->epoch + 1;   // partials / ticket alternate by epoch parity (overlapped launches)
    CU(launch_reduce_stats_xchg(b->dtype, b->ptr, v ? v->ptr : nullptr, b->len, with_minmax != 0,
                                x->partials + (epoch & 1) * (size_t)reduce_max_grid(), x->ticket + (epoch & 1) * 16,
                                static_cast<AggRaw*>(out_device), nullptr, xchg_dev(x, epoch), late, !x->shares_device,
                                c->stream));
    x->epoch++;   // only a launch that happened consumes an epoch (a failed one would leave this rank ahead of its peers)
    c->launches++;
    c->last_reduce_launch = c->launches;
    return MNR_OK;
}

// Read-and-clear the exchange's error word after the stream has drained.
static int finish_exchange_sync(mnr_ctx* c, mnr_xchg* x) {
    unsigned int* herr = reinterpret_cast<unsigned int*>(static_cast<char*>(c->h_scratch) + 128);
    CU(cudaMemcpyAsync(herr, x->err, 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (*herr) {
        CU(cudaMemsetAsync(x->err, 0, 4, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        return fail(MNR_ERR_CUDA, "fused exchange timed out waiting for a peer's partial (epoch %llu); the result is unusable", x->epoch);
    }
    return MNR_OK;
}

int mnr_reduce_stats_exchange_sync(mnr_ctx* c, mnr_xchg* x, const mnr_buf* b, const mnr_bits* v, int with_minmax, mnr_agg* out_host) {
    REQUIRE(out_host, MNR_ERR_INVALID_ARGUMENTS, "out is NULL");
    int rc = mnr_reduce_stats_exchange(c, x, b, v, with_minmax, c->d_agg);
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->h_scratch, c->d_agg, sizeof(AggRaw), cudaMemcpyDeviceToHost, c->stream));
    rc = finish_exchange_sync(c, x);
    if (rc) return rc;
    memcpy(out_host, c->h_scratch, sizeof(mnr_agg));
    return MNR_OK;
}

// ---- sharded SuperArray / SuperTable reduction: chunk aggregates -> per-column fold -> exchange, one call ---------------
static void agg_identity(mnr_dtype dt, AggRaw* a) {
    a->sum = 0; a->count = 0;
    const uint64_t nan = 0x7ff8000000000000ull;
    switch (dt) {
        case MNR_I8: a->mn = (uint64_t)(int64_t)INT8_MAX; a->mx = (uint64_t)(int64_t)INT8_MIN; break;
        case MNR_I16: a->mn = (uint64_t)(int64_t)INT16_MAX; a->mx = (uint64_t)(int64_t)INT16_MIN; break;
        case MNR_I32: a->mn = (uint64_t)(int64_t)INT32_MAX; a->mx = (uint64_t)(int64_t)INT32_MIN; break;
        case MNR_I64: a->mn = (uint64_t)INT64_MAX; a->mx = (uint64_t)INT64_MIN; break;
        case MNR_U8: a->mn = UINT8_MAX; a->mx = 0; break;
        case MNR_U16: a->mn = UINT16_MAX; a->mx = 0; break;
        case MNR_U32: a->mn = UINT32_MAX; a->mx = 0; break;
        case MNR_U64: a->mn = UINT64_MAX; a->mx = 0; break;
        default: a->mn = nan; a->mx = nan; break;
    }
}
static int dtype_kind(mnr_dtype dt);

static size_t fold_desc_bytes(size_t n, size_t n_cols, size_t* off_idx, size_t* off_id, size_t* off_kind) {
    // fold descriptor: [grp_off (n_cols+1) u32][grp_idx n u32][pad to 16][identity n_cols x 32 B][kind n_cols u8]
    *off_idx = (n_cols + 1) * 4;
    *off_id = (*off_idx + n * 4 + 15) & ~(size_t)15;
    *off_kind = *off_id + n_cols * 32;
    return (*off_kind + n_cols + 15) & ~(size_t)15;
}

// Every allocation a sharded reduction of this shape needs, made up front.  cudaMalloc / cudaFree synchronise the
// device: a rank that allocated lazily after a co-located peer's kernel started spinning on its flag would deadlock
// against it, so mnr_group_reduce_stats reserves on ALL ranks before the first launch.
static int reduce_exchange_reserve(mnr_ctx* c, const ReducePlan& p, size_t n, size_t n_cols) {
    CU(cudaSetDevice(c->device));
    int rc = ensure_batch_scratch(c, p.flat.size(), p.flat.size(), p.pblk_total);
    if (rc) return rc;
    const size_t need_aggs = n ? n : 1;
    if (c->chunk_aggs_cap < need_aggs) {
        if (c->chunk_aggs) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->chunk_aggs)); }
        c->chunk_aggs = nullptr; c->chunk_aggs_cap = 0;
        CU(cudaMalloc(&c->chunk_aggs, sizeof(AggRaw) * need_aggs));
        c->chunk_aggs_cap = need_aggs;
    }
    size_t o1, o2, o3;
    const size_t desc_bytes = fold_desc_bytes(n, n_cols, &o1, &o2, &o3);
    if (c->fold_desc_bytes < desc_bytes * 2) {
        if (c->fold_desc) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->fold_desc)); }
        c->fold_desc = nullptr; c->fold_desc_bytes = 0;
        CU(cudaMalloc(&c->fold_desc, desc_bytes * 2));
        c->fold_desc_bytes = desc_bytes * 2;
    }
    return MNR_OK;
}

static int check_batch_exchange(mnr_ctx* c, mnr_xchg* x, size_t n, const mnr_buf* const* bufs, const mnr_bits* const* validities,
                                size_t n_cols, const uint32_t* col_of_chunk, const mnr_dtype* col_dtypes) {
    REQUIRE(c && (bufs || n == 0) && (col_of_chunk || n == 0) && col_dtypes, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    REQUIRE(n_cols >= 1 && n_cols <= MNR_XCHG_MAX_AGGS, MNR_ERR_INVALID_ARGUMENTS, "n_cols %zu out of range (1..%d)", n_cols, MNR_XCHG_MAX_AGGS);
    REQUIRE(n <= 0xffffffffull, MNR_ERR_INVALID_ARGUMENTS, "too many chunks");
    if (x) { int rc = check_xchg(c, x); if (rc) return rc; }
    for (size_t g = 0; g < n_cols; ++g) REQUIRE(valid_dtype(col_dtypes[g]), MNR_ERR_UNSUPPORTED_TYPE, "column %zu: unknown dtype %d", g, (int)col_dtypes[g]);
    for (size_t i = 0; i < n; ++i) {
        int rc = check_reduce(c, bufs[i], validities ? validities[i] : nullptr);
        if (rc) return rc;
        REQUIRE(col_of_chunk[i] < n_cols, MNR_ERR_OUT_OF_BOUNDS, "chunk %zu: column %u of %zu", i, col_of_chunk[i], n_cols);
        REQUIRE(bufs[i]->dtype == col_dtypes[col_of_chunk[i]], MNR_ERR_TYPE_MISMATCH, "chunk %zu has dtype %d, its column %u has %d", i,
                (int)bufs[i]->dtype, col_of_chunk[i], (int)col_dtypes[col_of_chunk[i]]);
    }
    return MNR_OK;
}

}  // extern "C"
namespace mnr {
int reduce_stats_batch_exchange_reserve(mnr_ctx* c, mnr_xchg* x, size_t n, const mnr_buf* const* bufs,
                                        const mnr_bits* const* validities, int with_minmax, size_t n_cols,
                                        const uint32_t* col_of_chunk, const mnr_dtype* col_dtypes) {
    int rc = check_batch_exchange(c, x, n, bufs, validities, n_cols, col_of_chunk, col_dtypes);
    if (rc) return rc;
    return reduce_exchange_reserve(c, plan_reduce_batch(n, bufs, validities, with_minmax != 0), n, n_cols);
}
}  // namespace mnr
extern "C" {

int mnr_reduce_stats_batch_exchange(mnr_ctx* c, mnr_xchg* x, size_t n, const mnr_buf* const* bufs,
                                    const mnr_bits* const* validities, int with_minmax, size_t n_cols,
                                    const uint32_t* col_of_chunk, const mnr_dtype* col_dtypes, void* out_device) {
    int rc = check_batch_exchange(c, x, n, bufs, validities, n_cols, col_of_chunk, col_dtypes);
    if (rc) return rc;
    REQUIRE(out_device && (reinterpret_cast<uintptr_t>(out_device) & 15u) == 0, MNR_ERR_INVALID_ARGUMENTS,
            "out_device must be a 16-byte aligned device pointer");
    const ReducePlan plan = plan_reduce_batch(n, bufs, validities, with_minmax != 0);
    rc = reduce_exchange_reserve(c, plan, n, n_cols);
    if (rc) return rc;
    size_t off_idx, off_id, off_kind;
    const size_t desc_bytes = fold_desc_bytes(n, n_cols, &off_idx, &off_id, &off_kind);
    std::vector<unsigned char> desc(desc_bytes, 0);
    uint32_t* grp_off = reinterpret_cast<uint32_t*>(desc.data());
    uint32_t* grp_idx = reinterpret_cast<uint32_t*>(desc.data() + off_idx);
    for (size_t i = 0; i < n; ++i) grp_off[col_of_chunk[i] + 1]++;
    for (size_t g = 0; g < n_cols; ++g) grp_off[g + 1] += grp_off[g];
    {
        std::vector<uint32_t> fill(grp_off, grp_off + n_cols);
        for (size_t i = 0; i < n; ++i) grp_idx[fill[col_of_chunk[i]]++] = (uint32_t)i;   // chunk order inside a column
    }
    for (size_t g = 0; g < n_cols; ++g) {
        agg_identity(col_dtypes[g], reinterpret_cast<AggRaw*>(desc.data() + off_id) + g);
        desc[off_kind + g] = (unsigned char)dtype_kind(col_dtypes[g]);
    }
    char* dd = static_cast<char*>(c->fold_desc) + (c->fold_flip ? c->fold_desc_bytes / 2 : 0);
    c->fold_flip ^= 1;
    CU(cudaMemcpyAsync(dd, desc.data(), desc_bytes, cudaMemcpyHostToDevice, c->stream));
    FoldArgs f{};
    f.gticket = c->ticket[3] + 12;
    f.total_segs = (uint32_t)n; f.n_groups = (uint32_t)n_cols;
    f.grp_off = reinterpret_cast<const uint32_t*>(dd);
    f.grp_idx = reinterpret_cast<const uint32_t*>(dd + off_idx);
    f.identity = reinterpret_cast<const AggRaw*>(dd + off_id);
    f.kind = reinterpret_cast<const uint8_t*>(dd + off_kind);
    f.local = c->fold_local;
    f.result = static_cast<AggRaw*>(out_device);
    f.result_host = nullptr;
    const XchgDev xd = x ? xchg_dev(x, x->epoch + 1) : XchgDev{};
    if (n == 0) {
        CU(launch_fold_exchange(c->chunk_aggs, f, xd, c->stream));
        c->launches++;
    } else {
        rc = reduce_batch_launch(c, plan, with_minmax != 0, c->chunk_aggs, f, xd);
        if (rc) { cudaMemsetAsync(f.gticket, 0, 4, c->stream); return rc; }   // the fold only fires after the LAST launch: no epoch was consumed
    }
    if (x) x->epoch++;
    return MNR_OK;
}

int mnr_reduce_stats_batch_exchange_sync(mnr_ctx* c, mnr_xchg* x, size_t n, const mnr_buf* const* bufs,
                                         const mnr_bits* const* validities, int with_minmax, size_t n_cols,
                                         const uint32_t* col_of_chunk, const mnr_dtype* col_dtypes, mnr_agg* out_host) {
    REQUIRE(c && out_host, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    int rc = mnr_reduce_stats_batch_exchange(c, x, n, bufs, validities, with_minmax, n_cols, col_of_chunk, col_dtypes, c->fold_result);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out_host, c->fold_result, sizeof(mnr_agg) * n_cols, cudaMemcpyDeviceToHost, c->stream));
    if (x) return finish_exchange_sync(c, x);
    CU(cudaStreamSynchronize(c->stream));
    return MNR_OK;
}

static int dtype_kind(mnr_dtype dt) {   // 0 signed, 1 unsigned, 2 float
    switch (dt) {
        case MNR_F32: case MNR_F64: return 2;
        case MNR_U8: case MNR_U16: case MNR_U32: case MNR_U64: return 1;
        default: return 0;
    }
}

double mnr_agg_mean(mnr_dtype dtype, const mnr_agg* a) {
    if (!a || a->count == 0) return std::numeric_limits<double>::quiet_NaN();
    const int k = dtype_kind(dtype);
    const double s = k == 2 ? a->sum.f64 : k == 1 ? (double)a->sum.u64 : (double)a->sum.i64;
    return s / (double)a->count;
}

static double fmin_skip(double a, double b) {
    if (b != b) return a;
    if (a != a) return b;
    if (b < a) return b;
    if (b == a && std::signbit(b)) return b;
    return a;
}
static double fmax_skip(double a, double b) {
    if (b != b) return a;
    if (a != a) return b;
    if (b > a) return b;
    if (b == a && !std::signbit(b)) return b;
    return a;
}

int mnr_agg_combine(mnr_dtype dtype, const mnr_agg* p, size_t n, mnr_agg* out) {
    REQUIRE(out && (p || n == 0), MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    REQUIRE(n >= 1, MNR_ERR_INVALID_ARGUMENTS, "need at least one partial (identities are dtype-specific)");
    const int k = dtype_kind(dtype);
    mnr_agg r = p[0];
    for (size_t i = 1; i < n; ++i) {
        r.count += p[i].count;
        if (k == 2) {
            r.sum.f64 = r.sum.f64 + p[i].sum.f64;   // index order: the documented rank-order add
            r.min.f64 = fmin_skip(r.min.f64, p[i].min.f64);
            r.max.f64 = fmax_skip(r.max.f64, p[i].max.f64);
        } else if (k == 1) {
            r.sum.u64 += p[i].sum.u64;
            if (p[i].min.u64 < r.min.u64) r.min.u64 = p[i].min.u64;
            if (p[i].max.u64 > r.max.u64) r.max.u64 = p[i].max.u64;
        } else {
            r.sum.u64 += p[i].sum.u64;   // wrapping
            if (p[i].min.i64 < r.min.i64) r.min.i64 = p[i].min.i64;
            if (p[i].max.i64 > r.max.i64) r.max.i64 = p[i].max.i64;
        }
    }
    *out = r;
    return MNR_OK;
}

// ---- host-slice drop-ins ---------------------------------------------------------------------------------------------
// Chunks of `host_chunk_rows` rows cycle through 3 device staging slots, each with its own stream:
// H2D(inputs) -> kernel -> D2H(outputs) are ordered inside a slot and overlap across slots, so the PCIe link
// is busy in both directions while the kernels run.
static int ensure_stage(mnr_ctx* c, size_t bytes_per_array) {
    if (c->stage_bytes >= bytes_per_array) return MNR_OK;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 6; ++j) {
            if (c->stage[i][j]) { CU(cudaFree(c->stage[i][j])); c->stage[i][j] = nullptr; }
            // 0 lhs, 1 rhs, 2 acc, 3 out: full width; 4 mask, 5 out_mask: 1 bit per row (<= bytes/8 for 8-bit types)
            CU(cudaMalloc(&c->stage[i][j], pad256(j < 4 ? bytes_per_array : bytes_per_array / 8 + 16)));
        }
    c->stage_bytes = bytes_per_array;
    return MNR_OK;
}

static int sync_slots(mnr_ctx* c) {
    for (int i = 0; i < 3; ++i) CU(cudaStreamSynchronize(c->slot_stream[i]));
    return MNR_OK;
}

static int apply_host_impl(mnr_ctx* c, mnr_dtype dtype, int op, bool is_fma, const void* lhs, size_t lhs_len,
                           const void* rhs, size_t rhs_len, const void* acc, size_t acc_len, const uint8_t* mask,
                           void* out, uint8_t* out_mask) {
    REQUIRE(c, MNR_ERR_INVALID_ARGUMENTS, "ctx is NULL");
    REQUIRE(valid_dtype(dtype), MNR_ERR_UNSUPPORTED_TYPE, "unknown dtype %d", (int)dtype);
    REQUIRE(lhs_len == rhs_len, MNR_ERR_LENGTH_MISMATCH, "apply numeric: length mismatch (lhs: %zu, rhs: %zu)", lhs_len, rhs_len);
    if (is_fma) {
        REQUIRE(lhs_len == acc_len, MNR_ERR_LENGTH_MISMATCH, "acc length mismatch (lhs: %zu, rhs: %zu)", lhs_len, acc_len);
        REQUIRE(is_float_dtype(dtype), MNR_ERR_UNSUPPORTED_TYPE, "fma needs F32 or F64");
    } else {
        REQUIRE(op >= MNR_ADD && op <= MNR_FLOORDIV, MNR_ERR_OPERATOR_MISMATCH, "unknown operator %d", op);
    }
    const size_t n = lhs_len;
    if (n == 0) return MNR_OK;
    REQUIRE(lhs && rhs && out && (!is_fma || acc) && (!mask || out_mask), MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    CU(cudaSetDevice(c->device));
    const size_t es = dtype_size(dtype), chunk = c->host_chunk_rows;
    int rc = ensure_stage(c, chunk * 8);
    if (rc) return rc;
    const bool dense_int_div = !mask && !is_fma && !is_float_dtype(dtype) && (op == MNR_DIV || op == MNR_REM || op == MNR_FLOORDIV);
    unsigned int* flag = c->ticket[0] + 8;
    if (dense_int_div) { CU(cudaMemsetAsync(flag, 0, 4, c->slot_stream[0])); CU(cudaStreamSynchronize(c->slot_stream[0])); }
    size_t k = 0;
    for (size_t r0 = 0; r0 < n; r0 += chunk, ++k) {
        const int sl = (int)(k % 3);
        cudaStream_t s = c->slot_stream[sl];
        const size_t rows = n - r0 < chunk ? n - r0 : chunk;
        void** st = c->stage[sl];
        CU(cudaMemcpyAsync(st[0], static_cast<const char*>(lhs) + r0 * es, rows * es, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(st[1], static_cast<const char*>(rhs) + r0 * es, rows * es, cudaMemcpyHostToDevice, s));
        if (is_fma) CU(cudaMemcpyAsync(st[2], static_cast<const char*>(acc) + r0 * es, rows * es, cudaMemcpyHostToDevice, s));
        if (mask) CU(cudaMemcpyAsync(st[4], mask + r0 / 8, mask_bytes(rows), cudaMemcpyHostToDevice, s));
        if (is_fma) {
            CU(launch_ew_fma(dtype, st[0], st[1], st[2], mask ? static_cast<uint8_t*>(st[4]) : nullptr, st[3],
                             mask ? static_cast<uint8_t*>(st[5]) : nullptr, rows, s));
        } else {
            EwArgs a{};
            a.dtype = dtype; a.op = op; a.lhs = st[0]; a.rhs = st[1];
            a.lmask = mask ? static_cast<uint8_t*>(st[4]) : nullptr;
            a.out = st[3]; a.out_mask = mask ? static_cast<uint8_t*>(st[5]) : nullptr; a.n = rows; a.div0_flag = flag; a.k = c->knobs;
            CU(launch_ew_binary(a, s));
        }
        c->launches++;
        CU(cudaMemcpyAsync(static_cast<char*>(out) + r0 * es, st[3], rows * es, cudaMemcpyDeviceToHost, s));
        if (mask) CU(cudaMemcpyAsync(out_mask + r0 / 8, st[5], mask_bytes(rows), cudaMemcpyDeviceToHost, s));
    }
    rc = sync_slots(c);
    if (rc) return rc;
    if (dense_int_div) {
        unsigned int* h = static_cast<unsigned int*>(c->h_scratch);
        CU(cudaMemcpy(h, flag, 4, cudaMemcpyDeviceToHost));
        if (*h) return fail(MNR_ERR_DIVIDE_BY_ZERO, "%s by zero in dense integer kernel (the reference panics here)",
                            op == MNR_DIV ? "Division" : op == MNR_REM ? "Remainder" : "Floor division");
    }
    return MNR_OK;
}

int mnr_apply_host(mnr_ctx* c, mnr_dtype dtype, mnr_op op, const void* lhs, size_t lhs_len, const void* rhs, size_t rhs_len,
                   const uint8_t* mask, void* out, uint8_t* out_mask) {
    return apply_host_impl(c, dtype, (int)op, false, lhs, lhs_len, rhs, rhs_len, nullptr, 0, mask, out, out_mask);
}
int mnr_apply_fma_host(mnr_ctx* c, mnr_dtype dtype, const void* lhs, size_t lhs_len, const void* rhs, size_t rhs_len,
                       const void* acc, size_t acc_len, const uint8_t* mask, void* out, uint8_t* out_mask) {
    return apply_host_impl(c, dtype, 0, true, lhs, lhs_len, rhs, rhs_len, acc, acc_len, mask, out, out_mask);
}

#define MNR_TYPED_APPLY(NAME, CT, DT)                                                                                  \
    int NAME(mnr_ctx* c, const CT* lhs, size_t lhs_len, const CT* rhs, size_t rhs_len, mnr_op op, const uint8_t* mask, \
             CT* out, uint8_t* out_mask) {                                                                             \
        return mnr_apply_host(c, DT, op, lhs, lhs_len, rhs, rhs_len, mask, out, out_mask);                             \
    }
MNR_TYPED_APPLY(mnr_apply_int_i32, int32_t, MNR_I32)
MNR_TYPED_APPLY(mnr_apply_int_u32, uint32_t, MNR_U32)
MNR_TYPED_APPLY(mnr_apply_int_i64, int64_t, MNR_I64)
MNR_TYPED_APPLY(mnr_apply_int_u64, uint64_t, MNR_U64)
MNR_TYPED_APPLY(mnr_apply_float_f32, float, MNR_F32)
MNR_TYPED_APPLY(mnr_apply_float_f64, double, MNR_F64)
#undef MNR_TYPED_APPLY

int mnr_stats_host(mnr_ctx* c, mnr_dtype dtype, const void* data, size_t len, const uint8_t* validity, int with_minmax,
                   mnr_agg* out) {
    REQUIRE(c && out, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    REQUIRE(valid_dtype(dtype), MNR_ERR_UNSUPPORTED_TYPE, "unknown dtype %d", (int)dtype);
    REQUIRE(data || len == 0, MNR_ERR_INVALID_ARGUMENTS, "data is NULL");
    CU(cudaSetDevice(c->device));
    const size_t es = dtype_size(dtype), chunk = c->host_chunk_rows;
    const size_t nchunks = len ? (len + chunk - 1) / chunk : 1;
    int rc = ensure_stage(c, chunk * 8);
    if (rc) return rc;
    if (c->chunk_aggs_cap < nchunks) {
        if (c->chunk_aggs) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->chunk_aggs)); }
        c->chunk_aggs = nullptr; c->chunk_aggs_cap = 0;   // never leave a dangling pointer behind a failed cudaMalloc
        CU(cudaMalloc(&c->chunk_aggs, sizeof(AggRaw) * nchunks));
        c->chunk_aggs_cap = nchunks;
    }
    if (nchunks == 1) {
        // One chunk (every column up to host_chunk_rows rows): copy in, one launch, the aggregate polled out of mapped pinned
        // host memory as in reduce_sync — no stream synchronisation and no device-to-host copy on the way back.
        cudaStream_t s = c->slot_stream[0];
        void** st = c->stage[0];
        if (len) {
            CU(cudaMemcpyAsync(st[0], data, len * es, cudaMemcpyHostToDevice, s));
            if (validity) CU(cudaMemcpyAsync(st[4], validity, mask_bytes(len), cudaMemcpyHostToDevice, s));
        }
        const uint32_t seq = next_host_seq(c);
        CU(launch_reduce_stats(dtype, st[0], validity ? static_cast<uint8_t*>(st[4]) : nullptr, len, with_minmax != 0,
                               c->partials[0], c->ticket[0], c->chunk_aggs,
                               reinterpret_cast<AggRaw*>(const_cast<uint64_t*>(host_result_slot(c))), s, seq));
        c->launches++;
        return await_host_result(c, s, seq, out);
    }
    for (size_t k = 0; k < nchunks; ++k) {
        const int sl = (int)(k % 3);
        cudaStream_t s = c->slot_stream[sl];
        const size_t r0 = k * chunk;
        const size_t rows = len - r0 < chunk ? len - r0 : chunk;
        void** st = c->stage[sl];
        if (rows) {
            CU(cudaMemcpyAsync(st[0], static_cast<const char*>(data) + r0 * es, rows * es, cudaMemcpyHostToDevice, s));
            if (validity) CU(cudaMemcpyAsync(st[4], validity + r0 / 8, mask_bytes(rows), cudaMemcpyHostToDevice, s));
        }
        CU(launch_reduce_stats(dtype, st[0], validity ? static_cast<uint8_t*>(st[4]) : nullptr, rows, with_minmax != 0,
                               c->partials[sl], c->ticket[sl], c->chunk_aggs + k, nullptr, s));
        c->launches++;
    }
    rc = sync_slots(c);
    if (rc) return rc;
    std::vector<mnr_agg> parts(nchunks);
    CU(cudaMemcpy(parts.data(), c->chunk_aggs, sizeof(mnr_agg) * nchunks, cudaMemcpyDeviceToHost));
    return mnr_agg_combine(dtype, parts.data(), nchunks, out);
}

int mnr_bitmask_binop_host(mnr_ctx* c, mnr_logical_op op, const uint8_t* lhs, size_t lo, const uint8_t* rhs, size_t ro,
                           size_t len, uint8_t* out) {
    REQUIRE(c, MNR_ERR_INVALID_ARGUMENTS, "ctx is NULL");
    REQUIRE(op >= MNR_AND && op <= MNR_XOR, MNR_ERR_OPERATOR_MISMATCH, "unknown logical operator %d", (int)op);
    if (len == 0) return MNR_OK;
    REQUIRE(lhs && rhs && out, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    CU(cudaSetDevice(c->device));
    const uint8_t *lp = lhs + lo / 8, *rp = rhs + ro / 8;   // byte-floored window starts
    const size_t chunk_bits = c->host_chunk_rows * 8;        // multiple of 8192 bits
    int rc = ensure_stage(c, c->host_chunk_rows * 8);
    if (rc) return rc;
    size_t k = 0;
    for (size_t b0 = 0; b0 < len; b0 += chunk_bits, ++k) {
        const int sl = (int)(k % 3);
        cudaStream_t s = c->slot_stream[sl];
        const size_t bits = len - b0 < chunk_bits ? len - b0 : chunk_bits;
        void** st = c->stage[sl];
        CU(cudaMemcpyAsync(st[0], lp + b0 / 8, mask_bytes(bits), cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(st[1], rp + b0 / 8, mask_bytes(bits), cudaMemcpyHostToDevice, s));
        CU(launch_bits_op((int)op, static_cast<uint8_t*>(st[0]), 0, bits, static_cast<uint8_t*>(st[1]), 0, bits, bits,
                          static_cast<uint8_t*>(st[3]), s));
        c->launches++;
        CU(cudaMemcpyAsync(out + b0 / 8, st[3], mask_bytes(bits), cudaMemcpyDeviceToHost, s));
    }
    return sync_slots(c);
}

int mnr_host_register(void* ptr, size_t bytes) {
    REQUIRE(ptr && bytes, MNR_ERR_INVALID_ARGUMENTS, "NULL/empty range");
    CU(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return MNR_OK;
}
int mnr_host_unregister(void* ptr) {
    REQUIRE(ptr, MNR_ERR_INVALID_ARGUMENTS, "NULL pointer");
    CU(cudaHostUnregister(ptr));
    return MNR_OK;
}
int mnr_host_alloc(size_t bytes, void** out) {
    REQUIRE(out, MNR_ERR_INVALID_ARGUMENTS, "out is NULL");
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return MNR_OK;
}
void mnr_host_free(void* ptr) {
    if (ptr) cudaFreeHost(ptr);
}

}  // extern "C"
