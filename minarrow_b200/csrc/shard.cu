// SuperArray / SuperTable chunks as shards over the GPUs of one box, driven from ONE process through the C ABI
// (include/minarrow_b200.h "sharding").  The reference's chunked containers are independent equal-schema units that its
// own route walks chunk by chunk (src/kernels/broadcast/super_array.rs:180-249, super_table.rs:38-73); here chunk i of n
// lives on rank floor(i * G / n), element-wise work is shard-local (one batched launch per device), and a reduction is
// one batched kernel per device whose last block folds the device's chunks per column and exchanges the per-column
// partials with every peer over NVLink peer memory (reduce_kernels.cuh) — no NCCL, no host round trip between devices.
// One-process-per-GPU callers use the same kernels through mnr_xchg_* + CUDA IPC (minarrow_b200/sharded.py).
#include <cstdarg>
#include <cstdio>
#include <vector>

#include "internal.h"

namespace mnr {
int fail_public(int code, const char* msg);
// api.cu: validation + every allocation of mnr_reduce_stats_batch_exchange, without launching anything
int reduce_stats_batch_exchange_reserve(mnr_ctx* c, mnr_xchg* x, size_t n, const mnr_buf* const* bufs,
                                        const mnr_bits* const* validities, int with_minmax, size_t n_cols,
                                        const uint32_t* col_of_chunk, const mnr_dtype* col_dtypes);
}

static int failf(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    return mnr::fail_public(code, buf);
}
#define REQUIRE(cond, code, ...) \
    do {                         \
        if (!(cond)) return failf(code, __VA_ARGS__); \
    } while (0)

struct mnr_group {
    std::vector<mnr_ctx*> ctx;
    std::vector<mnr_xchg*> xchg;
};

extern "C" {

// ---- chunk -> rank map (pure host arithmetic) -----------------------------------------------------------------------------
int mnr_shard_owner(size_t chunk, size_t n_chunks, int world) {
    REQUIRE(world >= 1 && chunk < n_chunks, MNR_ERR_OUT_OF_BOUNDS, "chunk %zu of %zu over %d ranks", chunk, n_chunks, world);
    return (int)(((unsigned __int128)chunk * (unsigned)world) / n_chunks);
}

int mnr_shard_chunk_range(size_t n_chunks, int world, int rank, size_t* lo, size_t* hi) {
    REQUIRE(lo && hi, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    REQUIRE(world >= 1 && rank >= 0 && rank < world, MNR_ERR_INVALID_ARGUMENTS, "rank %d of %d", rank, world);
    // chunks i with floor(i * world / n) == rank  <=>  ceil(rank * n / world) <= i < ceil((rank + 1) * n / world)
    auto first = [&](int r) { return (size_t)((((unsigned __int128)r * n_chunks) + (unsigned)world - 1) / (unsigned)world); };
    *lo = first(rank);
    *hi = first(rank + 1);
    return MNR_OK;
}

int mnr_shard_row_range(size_t n_rows, int world, int rank, size_t align, size_t* offset, size_t* len) {
    REQUIRE(offset && len, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    REQUIRE(world >= 1 && rank >= 0 && rank < world && align >= 1, MNR_ERR_INVALID_ARGUMENTS, "rank %d of %d, align %zu", rank, world, align);
    const size_t units = (n_rows + align - 1) / align;
    size_t start = 0, length = 0;
    for (int r = 0; r <= rank; ++r) {
        start += length;
        const size_t u = units / (size_t)world + ((size_t)r < units % (size_t)world ? 1 : 0);
        length = u * align;
        if (length > n_rows - start) length = n_rows - start;
    }
    *offset = start;
    *len = length;
    return MNR_OK;
}

// ---- one process, many GPUs -------------------------------------------------------------------------------------------------
void mnr_group_destroy(mnr_group* g) {
    if (!g) return;
    for (mnr_xchg* x : g->xchg) mnr_xchg_destroy(x);
    for (mnr_ctx* c : g->ctx) mnr_ctx_destroy(c);
    delete g;
}

int mnr_group_create(int world, const int* devices, mnr_group** out) {
    REQUIRE(out, MNR_ERR_INVALID_ARGUMENTS, "out is NULL");
    *out = nullptr;
    REQUIRE(world >= 1 && world <= 16, MNR_ERR_INVALID_ARGUMENTS, "world %d out of range (1..16)", world);
    mnr_group* g = new mnr_group();
    int rc = MNR_OK;
    for (int r = 0; r < world && !rc; ++r) {
        mnr_ctx* c = nullptr;
        rc = mnr_ctx_create(devices ? devices[r] : r, &c);
        if (!rc) g->ctx.push_back(c);
    }
    for (int r = 0; r < world && !rc; ++r) {
        mnr_xchg* x = nullptr;
        rc = mnr_xchg_create(g->ctx[r], world, r, &x);
        if (!rc) g->xchg.push_back(x);
    }
    for (int r = 0; r < world && !rc; ++r) rc = mnr_xchg_connect_local(g->xchg[r], g->xchg.data());
    if (rc) { mnr_group_destroy(g); return rc; }
    *out = g;
    return MNR_OK;
}

int mnr_group_world(const mnr_group* g) { return g ? (int)g->ctx.size() : 0; }
mnr_ctx* mnr_group_ctx(mnr_group* g, int rank) { return (g && rank >= 0 && rank < (int)g->ctx.size()) ? g->ctx[rank] : nullptr; }
mnr_xchg* mnr_group_xchg(mnr_group* g, int rank) { return (g && rank >= 0 && rank < (int)g->xchg.size()) ? g->xchg[rank] : nullptr; }

int mnr_group_synchronize(mnr_group* g) {
    REQUIRE(g, MNR_ERR_INVALID_ARGUMENTS, "group is NULL");
    for (mnr_ctx* c : g->ctx) {
        int rc = mnr_ctx_synchronize(c);
        if (rc) return rc;
    }
    return MNR_OK;
}

static int rank_of(const mnr_group* g, const mnr_ctx* c) {
    for (size_t r = 0; r < g->ctx.size(); ++r) if (g->ctx[r] == c) return (int)r;
    return -1;
}

int mnr_group_upload(mnr_group* g, mnr_dtype dtype, size_t n_chunks, const void* const* host_chunks, const size_t* lens,
                     const uint8_t* const* host_validity, mnr_buf** out_bufs, mnr_bits** out_validity) {
    REQUIRE(g && (n_chunks == 0 || (host_chunks && lens && out_bufs)), MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    REQUIRE(!host_validity || out_validity, MNR_ERR_INVALID_ARGUMENTS, "validity given but out_validity is NULL");
    const int world = (int)g->ctx.size();
    for (size_t i = 0; i < n_chunks; ++i) { out_bufs[i] = nullptr; if (out_validity) out_validity[i] = nullptr; }
    int rc = MNR_OK;
    for (size_t i = 0; i < n_chunks && !rc; ++i) {
        mnr_ctx* c = g->ctx[mnr_shard_owner(i, n_chunks, world)];
        // asynchronous on the owner's stream: with pinned host memory the G links copy at the same time
        rc = mnr_buf_upload_async(c, dtype, host_chunks[i], lens[i], &out_bufs[i]);
        if (!rc && host_validity && host_validity[i]) rc = mnr_bits_upload_async(c, host_validity[i], lens[i], &out_validity[i]);
    }
    const int rs = mnr_group_synchronize(g);   // the caller may reuse the host chunks on return
    if (!rc) rc = rs;
    if (rc)
        for (size_t i = 0; i < n_chunks; ++i) {
            mnr_buf_free(out_bufs[i]); out_bufs[i] = nullptr;
            if (out_validity) { mnr_bits_free(out_validity[i]); out_validity[i] = nullptr; }
        }
    return rc;
}

// Split a chunk list by owning rank (order inside a rank = caller order).
static int split_by_rank(const mnr_group* g, size_t n, const mnr_buf* const* bufs, std::vector<std::vector<size_t>>& idx) {
    idx.assign(g->ctx.size(), {});
    for (size_t i = 0; i < n; ++i) {
        REQUIRE(bufs[i], MNR_ERR_INVALID_ARGUMENTS, "chunk %zu is NULL", i);
        const int r = rank_of(g, bufs[i]->ctx);
        REQUIRE(r >= 0, MNR_ERR_INVALID_ARGUMENTS, "chunk %zu lives on a context outside the group", i);
        idx[r].push_back(i);
    }
    return MNR_OK;
}

int mnr_group_ew_binary(mnr_group* g, mnr_op op, size_t n, const mnr_buf* const* lhs, const mnr_buf* const* rhs,
                        const mnr_bits* const* lhs_mask, const mnr_bits* const* rhs_mask, mnr_mask_mode mode, mnr_buf** out,
                        mnr_bits** out_mask) {
    REQUIRE(g && (n == 0 || (lhs && rhs && out && out_mask)), MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    for (size_t i = 0; i < n; ++i) { out[i] = nullptr; out_mask[i] = nullptr; }
    std::vector<std::vector<size_t>> idx;
    int rc = split_by_rank(g, n, lhs, idx);
    if (rc) return rc;
    for (size_t i = 0; i < n; ++i) {
        REQUIRE(rhs[i], MNR_ERR_INVALID_ARGUMENTS, "chunk %zu: rhs is NULL", i);
        REQUIRE(rhs[i]->ctx == lhs[i]->ctx, MNR_ERR_INVALID_ARGUMENTS,
                "chunk %zu: operands live on different devices (element-wise work is shard-local: re-shard one side first)", i);
    }
    for (size_t r = 0; r < idx.size() && !rc; ++r) {
        const size_t m = idx[r].size();
        if (!m) continue;
        std::vector<const mnr_buf*> l(m), rr(m);
        std::vector<const mnr_bits*> lm(m, nullptr), rm(m, nullptr);
        std::vector<mnr_buf*> ob(m, nullptr);
        std::vector<mnr_bits*> om(m, nullptr);
        for (size_t k = 0; k < m; ++k) {
            const size_t i = idx[r][k];
            l[k] = lhs[i]; rr[k] = rhs[i];
            if (lhs_mask) lm[k] = lhs_mask[i];
            if (rhs_mask) rm[k] = rhs_mask[i];
        }
        rc = mnr_ew_binary_batch(g->ctx[r], op, m, l.data(), rr.data(), lm.data(), rm.data(), mode, ob.data(), om.data());
        if (!rc) for (size_t k = 0; k < m; ++k) { out[idx[r][k]] = ob[k]; out_mask[idx[r][k]] = om[k]; }
    }
    if (rc) for (size_t i = 0; i < n; ++i) { mnr_buf_free(out[i]); mnr_bits_free(out_mask[i]); out[i] = nullptr; out_mask[i] = nullptr; }
    return rc;
}

int mnr_group_ew_scalar(mnr_group* g, mnr_op op, size_t n, const mnr_buf* const* arrs, const void* const* scalars,
                        int scalar_is_lhs, const mnr_bits* const* masks, mnr_buf** out, mnr_bits** out_mask) {
    REQUIRE(g && (n == 0 || (arrs && scalars && out && out_mask)), MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    for (size_t i = 0; i < n; ++i) { out[i] = nullptr; out_mask[i] = nullptr; }
    std::vector<std::vector<size_t>> idx;
    int rc = split_by_rank(g, n, arrs, idx);
    if (rc) return rc;
    for (size_t r = 0; r < idx.size() && !rc; ++r) {
        const size_t m = idx[r].size();
        if (!m) continue;
        std::vector<const mnr_buf*> a(m);
        std::vector<const void*> sc(m);
        std::vector<const mnr_bits*> mk(m, nullptr);
        std::vector<mnr_buf*> ob(m, nullptr);
        std::vector<mnr_bits*> om(m, nullptr);
        for (size_t k = 0; k < m && !rc; ++k) {
            const size_t i = idx[r][k];
            a[k] = arrs[i]; sc[k] = scalars[i];
            if (masks) mk[k] = masks[i];
            rc = mnr_buf_alloc(g->ctx[r], arrs[i]->dtype, arrs[i]->len, &ob[k]);
            if (!rc && mk[k]) rc = mnr_bits_alloc(g->ctx[r], arrs[i]->len, &om[k]);
            out[i] = ob[k]; out_mask[i] = om[k];
        }
        if (!rc) rc = mnr_ew_scalar_batch_into(g->ctx[r], op, m, a.data(), sc.data(), scalar_is_lhs, mk.data(), ob.data(), om.data());
    }
    if (rc) for (size_t i = 0; i < n; ++i) { mnr_buf_free(out[i]); mnr_bits_free(out_mask[i]); out[i] = nullptr; out_mask[i] = nullptr; }
    return rc;
}

int mnr_group_reduce_stats(mnr_group* g, size_t n, const mnr_buf* const* bufs, const mnr_bits* const* validities,
                           int with_minmax, size_t n_cols, const uint32_t* col_of_chunk, const mnr_dtype* col_dtypes,
                           mnr_agg* out_host) {
    REQUIRE(g && out_host && (n == 0 || (bufs && col_of_chunk)) && col_dtypes, MNR_ERR_INVALID_ARGUMENTS, "NULL argument");
    std::vector<std::vector<size_t>> idx;
    int rc = split_by_rank(g, n, bufs, idx);
    if (rc) return rc;
    const size_t world = g->ctx.size();
    // Validate everything on every rank BEFORE the first launch: a rank that fails after its peers have launched would
    // leave them spinning on its flag until the timeout.
    for (size_t i = 0; i < n; ++i) {
        REQUIRE(col_of_chunk[i] < n_cols, MNR_ERR_OUT_OF_BOUNDS, "chunk %zu: column %u of %zu", i, col_of_chunk[i], n_cols);
        REQUIRE(bufs[i]->dtype == col_dtypes[col_of_chunk[i]], MNR_ERR_TYPE_MISMATCH, "chunk %zu has dtype %d, its column %u has %d", i,
                (int)bufs[i]->dtype, col_of_chunk[i], (int)col_dtypes[col_of_chunk[i]]);
        const mnr_bits* v = validities ? validities[i] : nullptr;
        REQUIRE(!v || v->len >= bufs[i]->len, MNR_ERR_INVALID_ARGUMENTS, "chunk %zu: validity has %zu bits, need %zu", i, v->len, bufs[i]->len);
    }
    REQUIRE(n_cols >= 1 && n_cols <= MNR_XCHG_MAX_AGGS, MNR_ERR_INVALID_ARGUMENTS, "n_cols %zu out of range (1..%d)", n_cols, MNR_XCHG_MAX_AGGS);
    // Pass 0 validates and allocates on every rank, pass 1 launches: cudaMalloc synchronises its device, so a rank that
    // allocated lazily after a co-located peer (virtual ranks) had started spinning on its flag would deadlock against it;
    // and a rank failing validation after its peers launched would leave them waiting until the timeout.
    for (int pass = 0; pass < 2; ++pass)
        for (size_t r = 0; r < world; ++r) {
            const size_t m = idx[r].size();
            std::vector<const mnr_buf*> b(m);
            std::vector<const mnr_bits*> v(m, nullptr);
            std::vector<uint32_t> col(m);
            for (size_t k = 0; k < m; ++k) {
                const size_t i = idx[r][k];
                b[k] = bufs[i]; col[k] = col_of_chunk[i];
                if (validities) v[k] = validities[i];
            }
            mnr_ctx* c = g->ctx[r];
            rc = pass == 0 ? mnr::reduce_stats_batch_exchange_reserve(c, g->xchg[r], m, b.data(), v.data(), with_minmax, n_cols, col.data(), col_dtypes)
                           : mnr_reduce_stats_batch_exchange(c, g->xchg[r], m, b.data(), v.data(), with_minmax, n_cols, col.data(), col_dtypes, c->fold_result);
            if (rc) return rc;
        }
    // Every rank holds the same bits; rank 0's copy goes to the caller.  All ranks are drained so the group is idle on return.
    mnr_ctx* c0 = g->ctx[0];
    cudaSetDevice(c0->device);
    cudaError_t e = cudaMemcpyAsync(out_host, c0->fold_result, sizeof(mnr_agg) * n_cols, cudaMemcpyDeviceToHost, c0->stream);
    if (e != cudaSuccess) return failf(MNR_ERR_CUDA, "CUDA error %d (%s) copying the group result", (int)e, cudaGetErrorString(e));
    rc = mnr_group_synchronize(g);
    if (rc) return rc;
    for (size_t r = 0; r < world; ++r) {
        int bad = 0;
        rc = mnr_xchg_status(g->xchg[r], 1, &bad);
        if (rc) return rc;
        if (bad) return failf(MNR_ERR_CUDA, "fused exchange timed out on rank %zu waiting for a peer's partial; the result is unusable", r);
    }
    return MNR_OK;
}

}  // extern "C"
