// Internal declarations shared by the translation units of libminarrow_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

#include "../../include/minarrow_b200.h"

namespace mnr {

struct AggRaw;

// ---- kernel launchers (one per .cu) --------------------------------------------------------------------
int reduce_max_grid();
// out_host: optional mapped pinned host address that receives a second copy of the aggregate (NULL = none).
// host_seq != 0 (24 bits): out_host is a 64-byte slot that receives eight {sequence, 32-bit word} pairs the host can poll.
cudaError_t launch_reduce_stats(mnr_dtype dt, const void* data, const uint8_t* mask, uint64_t n, bool minmax,
                                AggRaw* partials, unsigned int* ticket, AggRaw* out, AggRaw* out_host, cudaStream_t s,
                                uint32_t host_seq = 0);
// Fused reduction + cross-GPU exchange over peer memory (reduce_kernels.cuh "fused cross-GPU finish").
struct XchgDev;
// late_wait: take the programmatic-dependency wait after the streaming phase (column known to be at rest).
// pdl: launch with the programmatic-stream-serialization attribute (see mnr_xchg::shares_device for when not to).
cudaError_t launch_reduce_stats_xchg(mnr_dtype dt, const void* data, const uint8_t* mask, uint64_t n, bool minmax,
                                     AggRaw* partials, unsigned int* ticket, AggRaw* out, AggRaw* out_host,
                                     const XchgDev& x, bool late_wait, bool pdl, cudaStream_t s);
// Batched form: one launch for `nseg` columns/chunks of one (dtype, alignment tier, masked) class.
struct ReduceSeg;
int reduce_tier(const void* data, bool minmax);
uint32_t reduce_nblk(mnr_dtype dt, uint64_t n, int tier, bool minmax);
// f.gticket != NULL adds the second stage: per-column fold of the chunk aggregates + cross-GPU exchange (FoldArgs).
struct FoldArgs;
cudaError_t launch_reduce_stats_batch(mnr_dtype dt, int tier, bool masked, bool minmax, const ReduceSeg* segs,
                                      uint32_t nseg, uint32_t max_blk, AggRaw* partials, unsigned int* tickets,
                                      AggRaw* outs, const FoldArgs& f, const XchgDev& x, cudaStream_t s);
cudaError_t launch_fold_exchange(const AggRaw* outs, const FoldArgs& f, const XchgDev& x, cudaStream_t s);
// Load every reduction kernel on the current device now (see reduce.cu: lazy loading vs. kernels that wait on peers).
cudaError_t reduce_preload_all();

// Element-wise binary op.  lhs/rhs: device pointers, or NULL for the side held in `scalar_bits`
// (at most one).  lmask/rmask: NULL or validity bytes indexed from bit 0.  out_mask: required iff a mask is
// given.  div0_flag: device word set to 1 when a dense integer Div/Rem/FloorDiv meets a zero divisor.
// Launch-geometry knobs of the element-wise kernels (mnr_ctx_set_option; per context, carried by value in EwArgs).
struct EwKnobs {
    int grid_cap = 0;     // > 0: cap every element-wise grid at this many blocks
    int max_tier = 2;     // 1: never use the 256-bit tier
    int sdiv64_cfg = 0;   // geometry of 64-bit column / scalar (elementwise.cu CfgSdiv64)
    int fdiv_cfg = 0;     // geometry of float Div / FloorDiv (CfgFdiv2 / CfgFdiv3)
    int heavy_cfg = 0;    // geometry of integer Div/Rem/FloorDiv, float Rem, Power
    int cheap8_cfg = 0;   // geometry of masked add / sub / mul on 1-byte columns
};
struct EwArgs {
    mnr_dtype dtype;
    int op;
    const void* lhs;
    const void* rhs;
    uint64_t scalar_bits;
    const uint8_t* lmask;
    const uint8_t* rmask;
    int mask_or;            // 0 = AND, 1 = OR (only when both masks are present)
    void* out;
    uint8_t* out_mask;
    uint64_t n;
    unsigned int* div0_flag;
    // integer Div/Rem/FloorDiv by a non-zero scalar divisor: host-computed multiplicative inverse (divmagic.h);
    // filled in by prepare_scalar_division() in api.cu, sdiv = 0 otherwise.
    int sdiv;
    uint64_t magic_m;
    uint32_t magic_s1, magic_s2;
    int nonzero_divisor;    // host-side: the divisor is a broadcast scalar known to be non-zero (no divide-by-zero check needed)
    EwKnobs k;
};
cudaError_t launch_ew_binary(const EwArgs& a, cudaStream_t s);
// I32 operand promoted on load against an F32/F64 operand (routing/arithmetic.rs:244-269).
cudaError_t launch_ew_promote(const EwArgs& a, mnr_dtype lhs_dtype, mnr_dtype rhs_dtype, cudaStream_t s);
// Batched element-wise launch: `segs` = device array of EwDev descriptors of one (dtype, op class, masked, tier) class.
struct EwDev;
int ew_batch_tier(mnr_dtype dt, int op, bool sdiv, const void* lhs, const void* rhs, const void* out, int max_tier);
cudaError_t launch_ew_batch(mnr_dtype dt, int op, int tier, bool masked, bool sdiv, const EwDev* segs, uint32_t nseg,
                            uint64_t max_n, int grid_cap, cudaStream_t s);
cudaError_t launch_ew_fma(mnr_dtype dt, const void* a, const void* b, const void* c, const uint8_t* mask, void* out,
                          uint8_t* out_mask, uint64_t n, cudaStream_t s);

// Bitmask logic over windows.  op: 0 AND, 1 OR, 2 XOR, 3 XNOR, 4 NOT(a), 5 COPY(a).  Bit positions are the
// exact window starts (the API layer floors them where the reference does).  Writes ceil(len/8) bytes,
// slack bits of the last byte zero; reads stay inside [0, ceil(x_bits_total/8)).
cudaError_t launch_bits_op(int op, const uint8_t* a, uint64_t a_bitpos, uint64_t a_total_bits, const uint8_t* b,
                           uint64_t b_bitpos, uint64_t b_total_bits, uint64_t len, uint8_t* out, cudaStream_t s);
// Device consolidate (concat.cu): chunks in row order -> one contiguous column (+ validity gathered at bit granularity).
struct ConcatSeg {
    const void* data;
    const uint8_t* mask;   // NULL = all valid
    uint64_t rows;
    uint64_t row0;         // destination row offset (prefix sum of rows)
};
cudaError_t launch_concat(int elem_bytes, const ConcatSeg* segs, uint32_t nseg, uint64_t max_rows, uint64_t total_rows, void* out,
                          uint8_t* out_mask, cudaStream_t s);
// Typed compare -> bitmask (compare.cu): bit i = ((data[i] & field_mask) == target), elements of 1/2/4/8 bytes.
cudaError_t launch_eq_mask(int elem_bytes, const void* data, uint64_t n, uint64_t field_mask, uint64_t target, uint8_t* out,
                           cudaStream_t s);
// Popcount of (a [xor b]) over len bits from the given bit positions -> *result (device) and, if non-NULL,
// *result_host (mapped pinned host memory).  partials: >= popcount_max_grid() words; ticket: zeroed, re-armed by the kernel.
int popcount_max_grid();
cudaError_t launch_bits_popcount(const uint8_t* a, uint64_t a_bitpos, uint64_t a_total_bits, const uint8_t* b,
                                 uint64_t b_bitpos, uint64_t b_total_bits, uint64_t len, unsigned long long* partials,
                                 unsigned int* ticket, unsigned long long* result, unsigned long long* result_host,
                                 cudaStream_t s);

}  // namespace mnr

// ---- handle types ----------------------------------------------------------------------------------------
struct mnr_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint64_t launches = 0;
    // scratch (device).  Index 0..2: the host-pipeline slots, 3: the context stream.
    mnr::AggRaw* partials[4] = {};
    unsigned int* ticket[4] = {};          // [i][0] reduction ticket; ticket[0][8] dense-int divide-by-zero flag
    mnr::AggRaw* d_agg = nullptr;          // result slot for synchronous reductions
    unsigned long long* d_count = nullptr; // popcount result
    void* h_scratch = nullptr;             // 256 bytes, pinned
    uint32_t host_seq = 0;                 // sequence number of the last polled host result (reduce_sync)
    // host drop-in pipeline: 3 staging slots, one stream each
    cudaStream_t slot_stream[3] = {};
    void* stage[3][6] = {};                // lhs, rhs, acc, out, mask, out_mask
    size_t stage_bytes = 0;
    size_t host_chunk_rows = (size_t)1 << 22;
    mnr::AggRaw* chunk_aggs = nullptr;     // one aggregate per chunk of mnr_stats_host
    size_t chunk_aggs_cap = 0;
    // batched reductions: per-segment partials / descriptors (double-buffered) / tickets
    void* batch_partials = nullptr;
    size_t batch_partials_bytes = 0;
    void* batch_segs = nullptr;
    size_t batch_segs_bytes = 0;
    void* batch_tickets = nullptr;
    size_t batch_tickets_bytes = 0;
    int batch_flip = 0;
    void* ew_segs = nullptr;               // batched element-wise descriptors (double-buffered)
    size_t ew_segs_bytes = 0;
    int ew_flip = 0;
    mnr::EwKnobs knobs;                    // per-context launch-geometry knobs (mnr_ctx_set_option)
    // sharded reductions: per-column fold descriptors (double-buffered), this rank's column partials, results
    void* fold_desc = nullptr;
    size_t fold_desc_bytes = 0;
    int fold_flip = 0;
    mnr::AggRaw* fold_local = nullptr;     // MNR_XCHG_MAX_AGGS aggregates
    mnr::AggRaw* fold_result = nullptr;    // MNR_XCHG_MAX_AGGS aggregates
    // consecutive reductions may overlap (programmatic dependent launch) when the caller opted in and no other kernel of
    // this library was launched on the stream in between
    int reduce_overlap = 0;
    uint64_t last_reduce_launch = 0;       // value of `launches` right after the last exchange reduction
};

// Cross-GPU mailbox set of one rank (reduce_kernels.cuh "fused cross-GPU finish").
struct mnr_xchg {
    mnr_ctx* ctx = nullptr;
    int world = 0, rank = 0;
    unsigned long long epoch = 0;
    char* mailbox = nullptr;                 // own mailbox (cudaMalloc: IPC-exportable)
    char* peers[16] = {};                    // every rank's mailbox as mapped here (peers[rank] == mailbox)
    bool opened[16] = {};                    // mapped through CUDA IPC (to be closed)
    unsigned int* err = nullptr;             // device word: a peer's flag never arrived
    unsigned long long* done = nullptr;      // device word: last epoch finished on this rank
    mnr::AggRaw* partials = nullptr;         // 2 x reduce_max_grid(): block partials, alternating by epoch parity
    unsigned int* ticket = nullptr;          // 2 x 16 words: finish tickets, alternating by epoch parity
    bool connected = false;
    // Another rank of this exchange lives on the SAME device (virtual ranks).  Then reductions are launched without the
    // programmatic-launch attribute: an early-scheduled successor kernel parks resident blocks at its dependency wait,
    // and on a shared device those blocks can starve the co-located peer whose flag the predecessor is waiting for.
    bool shares_device = false;
};

struct mnr_buf {
    mnr_ctx* ctx;
    mnr_dtype dtype;
    void* ptr;
    size_t len;
    bool owned;
};

struct mnr_bits {
    mnr_ctx* ctx;
    uint8_t* ptr;
    size_t len;     // bits
    bool owned;
};
