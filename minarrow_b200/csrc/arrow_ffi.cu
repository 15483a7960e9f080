// Arrow C Data Interface <-> device buffers.
//
// The reference's only existing C ABI is the Arrow C Data Interface (src/ffi/arrow_c_ffi.rs:87-98 ArrowArray, :121-133
// ArrowSchema; export_to_c :432-470 writes buffers = [validity | NULL, values], n_buffers = 2, :481-490; import_from_c
// :640).  These entry points let any Arrow producer (PyArrow, arrow-rs, polars, Minarrow itself) feed the GPU path
// without a Minarrow-specific copy: a numeric or boolean array in HOST memory is uploaded honouring `offset` — an
// element offset for values, an arbitrary BIT offset for validity / boolean data (shifted on the device; the reference's
// bitmask_binop floors sub-byte offsets, src/kernels/bitmask/mod.rs:124-128, simd_mask handles them, src/utils.rs:230-239)
// — and device results are exported back as a host ArrowArray with a release callback.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "internal.h"

using namespace mnr;

namespace mnr {
int fail_public(int code, const char* msg);   // api.cu: sets the thread-local error string
}

namespace {

struct FmtMap { const char* fmt; mnr_dtype dt; };
const FmtMap kFmt[] = {{"c", MNR_I8}, {"C", MNR_U8}, {"s", MNR_I16}, {"S", MNR_U16}, {"i", MNR_I32},
                       {"I", MNR_U32}, {"l", MNR_I64}, {"L", MNR_U64}, {"f", MNR_F32}, {"g", MNR_F64}};

const char* fmt_of(mnr_dtype dt) {
    for (const auto& f : kFmt) if (f.dt == dt) return f.fmt;
    return nullptr;
}

// Exported arrays own one host block: [ArrowArray buffers[2]] + 64-byte aligned validity + values.
struct ExportPriv {
    const void* buffers[2];
    void* validity;
    void* values;
};

void release_array(struct ArrowArray* a) {
    if (!a || !a->release) return;
    ExportPriv* p = static_cast<ExportPriv*>(a->private_data);
    if (p) { free(p->validity); free(p->values); delete p; }
    a->release = nullptr;
    a->private_data = nullptr;
}

void release_schema(struct ArrowSchema* s) {
    if (!s || !s->release) return;
    s->release = nullptr;
}

void* alloc64(size_t bytes) {
    void* p = nullptr;
    if (posix_memalign(&p, 64, (bytes + 63) / 64 * 64 + 64) != 0) return nullptr;   // Vec64-style 64-byte alignment
    return p;
}

#define AFAIL(code, msg) return mnr::fail_public(code, msg)

// Upload bits [bit_offset, bit_offset + len) of a host bitmap as a fresh device mask starting at bit 0.
int upload_bits_window(mnr_ctx* c, const uint8_t* host, int64_t bit_offset, int64_t len, mnr_bits** out) {
    int rc = mnr_bits_alloc(c, (size_t)len, out);
    if (rc || len == 0) return rc;
    const size_t first = (size_t)bit_offset >> 3, shift = (size_t)bit_offset & 7;
    const size_t nbytes = (shift + (size_t)len + 7) >> 3;
    void* tmp = nullptr;
    const char* err = nullptr;
    int code = MNR_ERR_CUDA;
    if (shift == 0) {
        // byte-aligned window: plain copy; the slack bits of the last byte (Bitmask::mask_trailing_bits, bitmask.rs:83-90) are
        // cleared on the host copy of that one byte — a foreign producer may leave anything there
        const size_t nb = (size_t)(len + 7) >> 3;
        uint8_t last = host[first + nb - 1];
        if (len & 7) last &= (uint8_t)((1u << (len & 7)) - 1u);
        bool ok = true;
        if (nb > 1) ok &= cudaMemcpyAsync((*out)->ptr, host + first, nb - 1, cudaMemcpyHostToDevice, c->stream) == cudaSuccess;
        ok &= cudaMemcpyAsync((*out)->ptr + nb - 1, &last, 1, cudaMemcpyHostToDevice, c->stream) == cudaSuccess;
        if (!ok) err = "arrow import: validity upload failed";
    } else if (cudaMallocAsync(&tmp, nbytes + 16, c->stream) != cudaSuccess) {
        tmp = nullptr;
        code = MNR_ERR_OUT_OF_MEMORY;
        err = "arrow import: staging allocation failed";
    } else if (cudaMemcpyAsync(tmp, host + first, nbytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) {
        err = "arrow import: validity upload failed";
    } else if (launch_bits_op(5, static_cast<const uint8_t*>(tmp), shift, nbytes * 8, nullptr, 0, 0, (uint64_t)len, (*out)->ptr, c->stream) != cudaSuccess) {
        err = "arrow import: bit shift failed";
    } else {
        c->launches++;
    }
    if (tmp) cudaFreeAsync(tmp, c->stream);
    if (!err && cudaStreamSynchronize(c->stream) != cudaSuccess) err = "arrow import: synchronize failed";
    if (err) {   // nothing leaks on a failure path: the staging buffer above, the fresh mask here
        mnr_bits_free(*out);
        *out = nullptr;
        AFAIL(code, err);
    }
    return MNR_OK;
}

}  // namespace

extern "C" {

int mnr_arrow_import(mnr_ctx* c, const struct ArrowArray* a, const struct ArrowSchema* s, mnr_buf** values, mnr_bits** data_bits,
                     mnr_bits** validity) {
    if (!c || !a || !s || !validity) AFAIL(MNR_ERR_INVALID_ARGUMENTS, "arrow import: NULL argument");
    if (values) *values = nullptr;
    if (data_bits) *data_bits = nullptr;
    *validity = nullptr;
    if (!s->format) AFAIL(MNR_ERR_INVALID_ARGUMENTS, "arrow import: schema has no format string");
    if (a->n_buffers != 2 || !a->buffers) AFAIL(MNR_ERR_UNSUPPORTED_TYPE, "arrow import: only fixed-width primitive / boolean layouts (n_buffers == 2)");
    if (a->length < 0 || a->offset < 0) AFAIL(MNR_ERR_INVALID_ARGUMENTS, "arrow import: negative length / offset");
    if (a->dictionary || a->n_children != 0) AFAIL(MNR_ERR_UNSUPPORTED_TYPE, "arrow import: nested / dictionary arrays are outside this path");
    if (cudaSetDevice(c->device) != cudaSuccess) AFAIL(MNR_ERR_CUDA, "arrow import: cudaSetDevice failed");
    const uint8_t* vbuf = static_cast<const uint8_t*>(a->buffers[0]);
    const uint8_t* dbuf = static_cast<const uint8_t*>(a->buffers[1]);
    if (!dbuf && a->length > 0) AFAIL(MNR_ERR_INVALID_ARGUMENTS, "arrow import: values buffer is NULL");
    int rc = MNR_OK;
    if (!strcmp(s->format, "b")) {
        if (!data_bits) AFAIL(MNR_ERR_INVALID_ARGUMENTS, "arrow import: boolean array needs data_bits");
        rc = upload_bits_window(c, dbuf, a->offset, a->length, data_bits);
    } else {
        const FmtMap* f = nullptr;
        for (const auto& k : kFmt) if (!strcmp(k.fmt, s->format)) f = &k;
        if (!f) AFAIL(MNR_ERR_UNSUPPORTED_TYPE, "arrow import: format is not a Minarrow numeric type (c C s S i I l L f g b)");
        if (!values) AFAIL(MNR_ERR_INVALID_ARGUMENTS, "arrow import: numeric array needs values");
        rc = mnr_buf_upload(c, f->dt, dbuf ? dbuf + (size_t)a->offset * dtype_size(f->dt) : nullptr, (size_t)a->length, values);
    }
    if (rc) return rc;
    if (vbuf && a->null_count != 0) {   // null_count == 0 => the producer promises all-valid; -1 = unknown
        rc = upload_bits_window(c, vbuf, a->offset, a->length, validity);
        if (rc) {
            if (values) { mnr_buf_free(*values); *values = nullptr; }
            if (data_bits) { mnr_bits_free(*data_bits); *data_bits = nullptr; }
        }
    }
    return rc;
}

static int export_common(mnr_ctx* c, const char* fmt, size_t len, const void* dev_values, size_t value_bytes, const mnr_bits* validity,
                         struct ArrowArray* out_array, struct ArrowSchema* out_schema) {
    if (cudaSetDevice(c->device) != cudaSuccess) AFAIL(MNR_ERR_CUDA, "arrow export: cudaSetDevice failed");
    ExportPriv* p = new ExportPriv();
    p->values = alloc64(value_bytes ? value_bytes : 1);
    p->validity = validity ? alloc64((len + 7) / 8 ? (len + 7) / 8 : 1) : nullptr;
    if (!p->values || (validity && !p->validity)) { free(p->values); free(p->validity); delete p; AFAIL(MNR_ERR_OUT_OF_MEMORY, "arrow export: host allocation failed"); }
    bool ok = true;
    if (value_bytes) ok &= cudaMemcpyAsync(p->values, dev_values, value_bytes, cudaMemcpyDeviceToHost, c->stream) == cudaSuccess;
    if (validity && len) ok &= cudaMemcpyAsync(p->validity, validity->ptr, (len + 7) / 8, cudaMemcpyDeviceToHost, c->stream) == cudaSuccess;
    uint64_t ones = len;
    ok &= cudaStreamSynchronize(c->stream) == cudaSuccess;
    if (!ok) { free(p->values); free(p->validity); delete p; AFAIL(MNR_ERR_CUDA, "arrow export: download failed"); }
    if (validity && len) {
        // `validity` may be longer than the values (a wrapped mask, or one taken before a slice): bits past `len` in the
        // last byte are not part of this array — clear them in the exported bitmap and keep them out of null_count
        uint8_t* vb = static_cast<uint8_t*>(p->validity);
        if (len & 7) vb[(len - 1) >> 3] &= (uint8_t)((1u << (len & 7)) - 1u);
        ones = 0;
        for (size_t i = 0; i < (len + 7) / 8; ++i) ones += (uint64_t)__builtin_popcount(vb[i]);
    }
    p->buffers[0] = p->validity;
    p->buffers[1] = p->values;
    memset(out_array, 0, sizeof *out_array);
    out_array->length = (int64_t)len;
    out_array->null_count = (int64_t)(len - ones);
    out_array->offset = 0;
    out_array->n_buffers = 2;
    out_array->buffers = p->buffers;
    out_array->release = release_array;
    out_array->private_data = p;
    if (out_schema) {
        memset(out_schema, 0, sizeof *out_schema);
        out_schema->format = fmt;          // static strings
        out_schema->name = "";
        out_schema->flags = validity ? 2 : 0;   // ARROW_FLAG_NULLABLE
        out_schema->release = release_schema;
    }
    return MNR_OK;
}

int mnr_arrow_export(mnr_ctx* c, const mnr_buf* values, const mnr_bits* validity, struct ArrowArray* out_array,
                     struct ArrowSchema* out_schema) {
    if (!c || !values || !out_array) AFAIL(MNR_ERR_INVALID_ARGUMENTS, "arrow export: NULL argument");
    if (validity && validity->len < values->len) AFAIL(MNR_ERR_INVALID_ARGUMENTS, "arrow export: validity shorter than values");
    const char* fmt = fmt_of(values->dtype);
    if (!fmt) AFAIL(MNR_ERR_UNSUPPORTED_TYPE, "arrow export: unknown dtype");
    return export_common(c, fmt, values->len, values->ptr, values->len * dtype_size(values->dtype), validity, out_array, out_schema);
}

int mnr_arrow_stream_import(mnr_ctx* c, struct ArrowArrayStream* st, size_t chunk_lo, size_t chunk_hi, size_t capacity,
                            mnr_buf** values, mnr_bits** validity, size_t* n_imported, size_t* n_seen) {
    if (!c || !st || !n_imported) AFAIL(MNR_ERR_INVALID_ARGUMENTS, "arrow stream import: NULL argument");
    if (!st->get_schema || !st->get_next) AFAIL(MNR_ERR_INVALID_ARGUMENTS, "arrow stream import: released or empty stream");
    if (capacity && (!values || !validity)) AFAIL(MNR_ERR_INVALID_ARGUMENTS, "arrow stream import: NULL output arrays");
    *n_imported = 0;
    if (n_seen) *n_seen = 0;
    struct ArrowSchema schema;
    memset(&schema, 0, sizeof schema);
    if (st->get_schema(st, &schema) != 0) {
        const char* e = st->get_last_error ? st->get_last_error(st) : nullptr;
        AFAIL(MNR_ERR_INVALID_ARGUMENTS, e ? e : "arrow stream import: get_schema failed");
    }
    int rc = MNR_OK;
    size_t seen = 0, got = 0;
    for (;;) {
        struct ArrowArray arr;
        memset(&arr, 0, sizeof arr);
        if (st->get_next(st, &arr) != 0) {
            const char* e = st->get_last_error ? st->get_last_error(st) : nullptr;
            rc = mnr::fail_public(MNR_ERR_INVALID_ARGUMENTS, e ? e : "arrow stream import: get_next failed");
            break;
        }
        if (!arr.release) break;   // end of stream
        if (rc == MNR_OK && seen >= chunk_lo && seen < chunk_hi) {
            if (got >= capacity) rc = mnr::fail_public(MNR_ERR_INVALID_ARGUMENTS, "arrow stream import: more chunks in range than capacity");
            else {
                rc = mnr_arrow_import(c, &arr, &schema, &values[got], nullptr, &validity[got]);
                if (rc == MNR_OK) ++got;
            }
        }
        arr.release(&arr);
        ++seen;   // keep draining after an error so the producer is left in a defined state
    }
    if (schema.release) schema.release(&schema);
    if (rc != MNR_OK) {
        for (size_t k = 0; k < got; ++k) { mnr_buf_free(values[k]); mnr_bits_free(validity[k]); values[k] = nullptr; validity[k] = nullptr; }
        return rc;
    }
    *n_imported = got;
    if (n_seen) *n_seen = seen;
    return MNR_OK;
}

int mnr_arrow_export_bool(mnr_ctx* c, const mnr_bits* data_bits, const mnr_bits* validity, struct ArrowArray* out_array,
                          struct ArrowSchema* out_schema) {
    if (!c || !data_bits || !out_array) AFAIL(MNR_ERR_INVALID_ARGUMENTS, "arrow export: NULL argument");
    if (validity && validity->len < data_bits->len) AFAIL(MNR_ERR_INVALID_ARGUMENTS, "arrow export: validity shorter than data");
    return export_common(c, "b", data_bits->len, data_bits->ptr, (data_bits->len + 7) / 8, validity, out_array, out_schema);
}

}  // extern "C"
