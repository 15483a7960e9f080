// TMA bulk-copy (cp.async.bulk + mbarrier) variants of the two headline kernels, timed against the library's LDG kernels
// (tuning tool, not part of the library).  The north-star asks for "TMA bulk copies where staging pays off"; this
// harness is the measurement behind the decision (results: profiles/*_sweep_tma.md).
//   nvcc <lib flags> tools/sweep_tma.cu -o tools/sweep_tma
//
// Variants
//   add.tma_ld      f64 masked add (two masks, fused AND): operands + validity staged into shared memory by 1-D bulk
//                   copies through a STAGES-deep mbarrier ring; results stored straight from registers (st.global.cs).
//   add.tma_ldst    same, results staged in shared memory and written with bulk stores (cp.async.bulk.global.shared).
//   sum.tma_ld      i64 masked sum + count: column + validity staged by bulk copies, per-block partial, atomic finish
//                   (the finish is NOT the library's deterministic one; only the streaming rate is of interest here).
// Every variant's output is checksummed against the library kernel on the same inputs.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../minarrow_b200/csrc/ew_kernels.cuh"
#include "../minarrow_b200/csrc/reduce_kernels.cuh"

using namespace mnr;

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- f64 masked add, two masks ----------------------------------------------------------------------------
// One tile = ROWS rows: x[ROWS], y[ROWS] (8 B each), mx, my (ROWS/8 B each).  ROWS = BLOCK * 4: a thread owns one
// 256-bit vector (4 rows) per tile; lanes 2k/2k+1 share an output validity byte (shuffle gather, as the library).
template <int BLOCK, int STAGES, bool BULK_STORE>
__global__ void __launch_bounds__(BLOCK) add_tma_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                        const uint8_t* __restrict__ mx, const uint8_t* __restrict__ my,
                                                        double* __restrict__ out, uint8_t* __restrict__ om, uint64_t n) {
    constexpr int ROWS = BLOCK * 4;
    constexpr int VB = ROWS * 8, MB = ROWS / 8;
    extern __shared__ __align__(128) unsigned char smem[];
    // layout per stage: x | y | (out) | mx | my | (omask)
    constexpr int STAGE_BYTES = VB * (BULK_STORE ? 3 : 2) + MB * (BULK_STORE ? 3 : 2);
    __shared__ uint64_t full[STAGES];
    const uint64_t ntiles = n / ROWS;   // harness sizes are multiples of ROWS
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto stage_ptr = [&](int s) { return smem + (size_t)s * STAGE_BYTES; };
    auto issue = [&](uint64_t tile, int s) {
        unsigned char* b = stage_ptr(s);
        mbar_expect_tx(&full[s], 2 * VB + 2 * MB);
        bulk_g2s(b, x + tile * ROWS, VB, &full[s]);
        bulk_g2s(b + VB, y + tile * ROWS, VB, &full[s]);
        bulk_g2s(b + VB * (BULK_STORE ? 3 : 2), mx + tile * MB, MB, &full[s]);
        bulk_g2s(b + VB * (BULK_STORE ? 3 : 2) + MB, my + tile * MB, MB, &full[s]);
    };
    uint64_t tile = blockIdx.x;
    if (threadIdx.x == 0) {
        uint64_t t = tile;
        for (int s = 0; s < STAGES - 1 && t < ntiles; ++s, t += gridDim.x) issue(t, s);
    }
    uint32_t it = 0;
    for (; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % STAGES;
        if (threadIdx.x == 0) {
            // refill the stage consumed in the previous iteration (all threads passed its trailing barrier)
            const uint64_t nt = tile + (uint64_t)(STAGES - 1) * gridDim.x;
            if (nt < ntiles) issue(nt, (it + STAGES - 1) % STAGES);
        }
        mbar_wait(&full[s], (it / STAGES) & 1);
        unsigned char* b = stage_ptr(s);
        const double4 a = reinterpret_cast<const double4*>(b)[threadIdx.x];
        const double4 c = reinterpret_cast<const double4*>(b + VB)[threadIdx.x];
        const unsigned char* mp = b + VB * (BULK_STORE ? 3 : 2);
        const uint32_t sh = (threadIdx.x & 1) * 4;
        const uint32_t m = ((mp[threadIdx.x >> 1] & mp[MB + (threadIdx.x >> 1)]) >> sh) & 15u;
        double4 r;
        r.x = (m & 1) ? a.x + c.x : 0.0; r.y = (m & 2) ? a.y + c.y : 0.0;
        r.z = (m & 4) ? a.z + c.z : 0.0; r.w = (m & 8) ? a.w + c.w : 0.0;
        const uint32_t pair = m | (__shfl_down_sync(0xffffffffu, m, 1) << 4);
        if constexpr (BULK_STORE) {
            reinterpret_cast<double4*>(b + 2 * VB)[threadIdx.x] = r;
            if (!(threadIdx.x & 1)) (b + 3 * VB + 2 * MB)[threadIdx.x >> 1] = (unsigned char)pair;
            fence_async_smem();
            // the out buffer the NEXT iteration writes was last read by the bulk store of iteration it+1-STAGES:
            // all but the newest STAGES-2 store groups must have finished reading shared memory
            if (threadIdx.x == 0) bulk_wait_read<STAGES - 2>();
            __syncthreads();
            if (threadIdx.x == 0) {
                bulk_s2g(out + tile * ROWS, b + 2 * VB, VB);
                bulk_s2g(om + tile * MB, b + 3 * VB + 2 * MB, MB);
                bulk_commit();
            }
        } else {
            V32 o; memcpy(&o, &r, 32);
            stg_stream(reinterpret_cast<V32*>(out + tile * ROWS) + threadIdx.x, o);
            if (!(threadIdx.x & 1)) om[tile * MB + (threadIdx.x >> 1)] = (unsigned char)pair;
            __syncthreads();   // everyone is done reading stage s before thread 0 refills it next iteration
        }
    }
    if constexpr (BULK_STORE) { if (threadIdx.x == 0) bulk_wait_read<0>(); }
}

// One-shot form: the grid covers the column (no persistent loop, like the library's full-grid LDG kernel); a CTA issues
// the bulk copies of its K sub-tiles up front and consumes them in order.
template <int BLOCK, int K>
__global__ void __launch_bounds__(BLOCK) add_tma_oneshot_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                                const uint8_t* __restrict__ mx, const uint8_t* __restrict__ my,
                                                                double* __restrict__ out, uint8_t* __restrict__ om, uint64_t n) {
    constexpr int ROWS = BLOCK * 4;
    constexpr int VB = ROWS * 8, MB = ROWS / 8;
    constexpr int STAGE_BYTES = VB * 2 + MB * 2;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[K];
    if (threadIdx.x == 0) {
        for (int s = 0; s < K; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < K; ++s) {
            const uint64_t tile = (uint64_t)blockIdx.x * K + s;
            unsigned char* b = smem + (size_t)s * STAGE_BYTES;
            mbar_expect_tx(&full[s], 2 * VB + 2 * MB);
            bulk_g2s(b, x + tile * ROWS, VB, &full[s]);
            bulk_g2s(b + VB, y + tile * ROWS, VB, &full[s]);
            bulk_g2s(b + 2 * VB, mx + tile * MB, MB, &full[s]);
            bulk_g2s(b + 2 * VB + MB, my + tile * MB, MB, &full[s]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < K; ++s) {
        const uint64_t tile = (uint64_t)blockIdx.x * K + s;
        mbar_wait(&full[s], 0);
        unsigned char* b = smem + (size_t)s * STAGE_BYTES;
        const double4 a = reinterpret_cast<const double4*>(b)[threadIdx.x];
        const double4 c = reinterpret_cast<const double4*>(b + VB)[threadIdx.x];
        const unsigned char* mp = b + 2 * VB;
        const uint32_t sh = (threadIdx.x & 1) * 4;
        const uint32_t m = ((mp[threadIdx.x >> 1] & mp[MB + (threadIdx.x >> 1)]) >> sh) & 15u;
        double4 r;
        r.x = (m & 1) ? a.x + c.x : 0.0; r.y = (m & 2) ? a.y + c.y : 0.0;
        r.z = (m & 4) ? a.z + c.z : 0.0; r.w = (m & 8) ? a.w + c.w : 0.0;
        const uint32_t pair = m | (__shfl_down_sync(0xffffffffu, m, 1) << 4);
        V32 o; memcpy(&o, &r, 32);
        stg_stream(reinterpret_cast<V32*>(out + tile * ROWS) + threadIdx.x, o);
        if (!(threadIdx.x & 1)) om[tile * MB + (threadIdx.x >> 1)] = (unsigned char)pair;
    }
}

// ---- i64 masked sum + count -------------------------------------------------------------------------------
template <int BLOCK, int STAGES, int VPT>   // VPT 128-bit vectors per thread per tile
__global__ void __launch_bounds__(BLOCK) sum_tma_kernel(const int64_t* __restrict__ d, const uint8_t* __restrict__ mk, uint64_t n,
                                                        unsigned long long* __restrict__ out) {
    constexpr int ROWS = BLOCK * 2 * VPT;
    constexpr int VB = ROWS * 8, MB = ROWS / 8;
    constexpr int STAGE_BYTES = VB + ((MB + 127) / 128) * 128;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[STAGES];
    const uint64_t ntiles = n / ROWS;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](uint64_t tile, int s) {
        unsigned char* b = smem + (size_t)s * STAGE_BYTES;
        mbar_expect_tx(&full[s], VB + MB);
        bulk_g2s(b, d + tile * ROWS, VB, &full[s]);
        bulk_g2s(b + VB, mk + tile * MB, MB, &full[s]);
    };
    uint64_t tile = blockIdx.x;
    if (threadIdx.x == 0) {
        uint64_t t = tile;
        for (int s = 0; s < STAGES - 1 && t < ntiles; ++s, t += gridDim.x) issue(t, s);
    }
    uint64_t acc = 0, cnt = 0;
    uint32_t it = 0;
    for (; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % STAGES;
        if (threadIdx.x == 0) {
            const uint64_t nt = tile + (uint64_t)(STAGES - 1) * gridDim.x;
            if (nt < ntiles) issue(nt, (it + STAGES - 1) % STAGES);
        }
        mbar_wait(&full[s], (it / STAGES) & 1);
        const unsigned char* b = smem + (size_t)s * STAGE_BYTES;
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
            const int vi = v * BLOCK + threadIdx.x;                    // vector index in the tile (2 rows each)
            const longlong2 e = reinterpret_cast<const longlong2*>(b)[vi];
            const uint32_t m = ((b + VB)[vi >> 2] >> ((vi & 3) * 2)) & 3u;
            acc += (m & 1 ? (uint64_t)e.x : 0ull) + (m & 2 ? (uint64_t)e.y : 0ull);
            cnt += __popc(m);
        }
        __syncthreads();
    }
    for (int off = 16; off > 0; off >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, off); cnt += __shfl_xor_sync(0xffffffffu, cnt, off); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, (unsigned long long)acc); atomicAdd(out + 1, (unsigned long long)cnt); }
}

// ---- harness ------------------------------------------------------------------------------------------------
__global__ void fill_kernel(uint64_t* p, uint64_t n, uint64_t seed, int as_double) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t z = (i + seed) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        if (as_double) { double dd = (double)(int64_t)(z >> 11) * (1.0 / 9007199254740992.0) * 200.0 - 100.0; p[i] = (uint64_t)__double_as_longlong(dd); }
        else p[i] = z;
    }
}
__global__ void checksum_kernel(const uint64_t* p, uint64_t n, unsigned long long* out) {
    unsigned long long acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) acc += p[i] * (2 * i + 1);
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}
static unsigned long long checksum(const void* p, size_t bytes) {
    unsigned long long* d; unsigned long long h = 0;
    CK(cudaMalloc(&d, 8)); CK(cudaMemset(d, 0, 8));
    checksum_kernel<<<148 * 8, 256>>>((const uint64_t*)p, bytes / 8, d);
    CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost)); CK(cudaFree(d));
    return h;
}
template <class F> static float time_ms(F f, int iters = 15) {
    for (int i = 0; i < 3; ++i) f();
    CK(cudaDeviceSynchronize());
    std::vector<float> ts;
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < iters; ++i) {
        CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); ts.push_back(ms);
    }
    CK(cudaGetLastError());
    std::sort(ts.begin(), ts.end());
    return ts[ts.size() / 2];
}

struct AddBufs { double *x, *y, *o; uint8_t *mx, *my, *om; uint64_t n; };
static unsigned long long g_ref;

template <int BLOCK, int STAGES, bool BULK_STORE> static void add_variant(const AddBufs& b, int ctas_per_sm) {
    constexpr int ROWS = BLOCK * 4;
    constexpr int STAGE_BYTES = ROWS * 8 * (BULK_STORE ? 3 : 2) + ROWS / 8 * (BULK_STORE ? 3 : 2);
    const int smem = STAGE_BYTES * STAGES;
    auto kern = add_tma_kernel<BLOCK, STAGES, BULK_STORE>;
    if (smem > 227 * 1024) { printf("add.%s block=%d stages=%d: %d B of shared memory does not fit\n", BULK_STORE ? "tma_ldst" : "tma_ld", BLOCK, STAGES, smem); return; }
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BLOCK, smem));
    if (occ < 1) { printf("add.%s block=%d stages=%d: does not fit\n", BULK_STORE ? "tma_ldst" : "tma_ld", BLOCK, STAGES); return; }
    const int per_sm = ctas_per_sm > 0 ? std::min(ctas_per_sm, occ) : occ;
    const unsigned grid = 148u * per_sm;
    CK(cudaMemset(b.o, 0, b.n * 8)); CK(cudaMemset(b.om, 0, b.n / 8));
    float ms = time_ms([&] { kern<<<grid, BLOCK, smem>>>(b.x, b.y, b.mx, b.my, b.o, b.om, b.n); });
    unsigned long long cs = checksum(b.o, b.n * 8) ^ checksum(b.om, b.n / 8);
    printf("f64_masked_add_2masks  %-12s block=%3d stages=%d smem=%6d B ctas/SM=%d (occ %d) grid=%5u  %8.4f ms  %8.1f GB/s  %s\n",
           BULK_STORE ? "add.tma_ldst" : "add.tma_ld", BLOCK, STAGES, smem, per_sm, occ, grid, ms, (double)b.n * 24.375 / ms / 1e6,
           cs == g_ref ? "ok" : "CHECKSUM-MISMATCH");
    fflush(stdout);
}

static AggRaw *g_partials, *g_agg; static unsigned int* g_ticket;
template <int BLOCK, int K> static void add_oneshot_variant(const AddBufs& b) {
    constexpr int ROWS = BLOCK * 4;
    const int smem = (ROWS * 8 * 2 + ROWS / 8 * 2) * K;
    auto kern = add_tma_oneshot_kernel<BLOCK, K>;
    if (smem > 227 * 1024) { printf("add.tma_1shot block=%d K=%d: does not fit\n", BLOCK, K); return; }
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BLOCK, smem));
    const unsigned grid = (unsigned)(b.n / ((uint64_t)ROWS * K));
    CK(cudaMemset(b.o, 0, b.n * 8)); CK(cudaMemset(b.om, 0, b.n / 8));
    float ms = time_ms([&] { kern<<<grid, BLOCK, smem>>>(b.x, b.y, b.mx, b.my, b.o, b.om, b.n); });
    unsigned long long cs = checksum(b.o, b.n * 8) ^ checksum(b.om, b.n / 8);
    printf("f64_masked_add_2masks  %-12s block=%3d K=%d smem=%6d B occ=%d grid=%7u (full)  %8.4f ms  %8.1f GB/s  %s\n", "add.tma_1shot", BLOCK, K, smem,
           occ, grid, ms, (double)b.n * 24.375 / ms / 1e6, cs == g_ref ? "ok" : "CHECKSUM-MISMATCH");
    fflush(stdout);
}

template <int BLOCK, int STAGES, int VPT> static void sum_variant(const int64_t* d, const uint8_t* m, uint64_t n, unsigned long long* out,
                                                                  int ctas_per_sm) {
    constexpr int ROWS = BLOCK * 2 * VPT;
    constexpr int STAGE_BYTES = ROWS * 8 + ((ROWS / 8 + 127) / 128) * 128;
    const int smem = STAGE_BYTES * STAGES;
    auto kern = sum_tma_kernel<BLOCK, STAGES, VPT>;
    if (smem > 227 * 1024) { printf("sum.tma_ld block=%d stages=%d vpt=%d: %d B of shared memory does not fit\n", BLOCK, STAGES, VPT, smem); return; }
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BLOCK, smem));
    if (occ < 1) { printf("sum.tma_ld block=%d stages=%d vpt=%d: does not fit\n", BLOCK, STAGES, VPT); return; }
    const int per_sm = ctas_per_sm > 0 ? std::min(ctas_per_sm, occ) : occ;
    const unsigned grid = 148u * per_sm;
    const uint64_t nn = n - n % ROWS;   // whole tiles only (the reference value below is computed over the same rows)
    reduce_stats_kernel<int64_t, V16, true, false, 256, 4, 4><<<kSMs * 4, 256>>>(d, m, nn, g_partials, g_ticket, g_agg, nullptr, XchgDev{});
    AggRaw ref; CK(cudaMemcpy(&ref, g_agg, sizeof ref, cudaMemcpyDeviceToHost));
    float ms = time_ms([&] { cudaMemsetAsync(out, 0, 16); kern<<<grid, BLOCK, smem>>>(d, m, nn, out); });
    unsigned long long h[2]; CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
    printf("i64_masked_sum         %-12s block=%3d stages=%d vpt=%d tile=%5d rows smem=%6d B ctas/SM=%d (occ %d)  %8.4f ms  %8.1f GB/s  %s\n",
           "sum.tma_ld", BLOCK, STAGES, VPT, ROWS, smem, per_sm, occ, ms, (double)nn * 8.125 / ms / 1e6,
           (h[0] == ref.sum && h[1] == ref.count) ? "ok" : "MISMATCH");
    fflush(stdout);
}

int main(int argc, char** argv) {
    const char* only = argc > 1 ? argv[1] : "";
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("# sweep_tma: device %s, %d SMs, sm_%d%d\n", p.name, p.multiProcessorCount, p.major, p.minor);
    if (!*only || strstr(only, "add")) {
        AddBufs b{}; b.n = 1ull << 28;
        CK(cudaMalloc(&b.x, b.n * 8)); CK(cudaMalloc(&b.y, b.n * 8)); CK(cudaMalloc(&b.o, b.n * 8));
        CK(cudaMalloc(&b.mx, b.n / 8)); CK(cudaMalloc(&b.my, b.n / 8)); CK(cudaMalloc(&b.om, b.n / 8));
        fill_kernel<<<148 * 8, 256>>>((uint64_t*)b.x, b.n, 1, 1); fill_kernel<<<148 * 8, 256>>>((uint64_t*)b.y, b.n, 2, 1);
        fill_kernel<<<148 * 8, 256>>>((uint64_t*)b.mx, b.n / 64, 3, 0); fill_kernel<<<148 * 8, 256>>>((uint64_t*)b.my, b.n / 64, 4, 0);
        CK(cudaDeviceSynchronize());
        {   // library kernel (the shipped configuration: 256-bit, U=4, 128 threads, full grid) = reference + baseline time
            EwDev a{}; a.lhs = b.x; a.rhs = b.y; a.lmask = b.mx; a.rmask = b.my; a.mask_or = 0; a.out = b.o; a.out_mask = b.om; a.n = b.n;
            a.op = MNR_ADD; unsigned int* flag; CK(cudaMalloc(&flag, 64)); a.div0_flag = flag;
            auto kern = ew_binary_kernel<double, double, double, V32, CLS_CHEAP, true, 128, 4, 1>;
            const uint64_t nvec = b.n / 4, tiles = nvec / (32 * 4), full = (tiles + 3) / 4;
            float ms = time_ms([&] { kern<<<(unsigned)full, 128>>>(a); });
            g_ref = checksum(b.o, b.n * 8) ^ checksum(b.om, b.n / 8);
            printf("f64_masked_add_2masks  %-12s block=128 U=4 vec=32B grid=full(%llu)  %8.4f ms  %8.1f GB/s  (library kernel)\n", "add.ldg",
                   (unsigned long long)full, ms, (double)b.n * 24.375 / ms / 1e6);
        }
        add_variant<256, 3, false>(b, 0); add_variant<256, 4, false>(b, 0); add_variant<256, 6, false>(b, 0); add_variant<256, 8, false>(b, 0);
        add_variant<256, 4, false>(b, 2); add_variant<256, 4, false>(b, 4); add_variant<256, 3, false>(b, 4);
        add_variant<512, 3, false>(b, 0); add_variant<512, 4, false>(b, 0); add_variant<512, 4, false>(b, 2);
        add_variant<128, 4, false>(b, 0); add_variant<128, 8, false>(b, 0); add_variant<128, 6, false>(b, 8);
        add_variant<1024, 3, false>(b, 0);
        add_oneshot_variant<128, 1>(b); add_oneshot_variant<128, 2>(b); add_oneshot_variant<128, 4>(b); add_oneshot_variant<256, 1>(b);
        add_oneshot_variant<256, 2>(b); add_oneshot_variant<256, 4>(b); add_oneshot_variant<512, 2>(b); add_oneshot_variant<64, 4>(b);
        add_variant<256, 3, true>(b, 0); add_variant<256, 4, true>(b, 0); add_variant<256, 4, true>(b, 2); add_variant<256, 6, true>(b, 0);
        add_variant<512, 3, true>(b, 0); add_variant<512, 4, true>(b, 0); add_variant<128, 4, true>(b, 0); add_variant<128, 6, true>(b, 0);
        add_variant<1024, 3, true>(b, 0);
        cudaFree(b.x); cudaFree(b.y); cudaFree(b.o); cudaFree(b.mx); cudaFree(b.my); cudaFree(b.om);
    }
    if (!*only || strstr(only, "sum")) {
        const uint64_t n = 1000000000ull;
        int64_t* d; uint8_t* m; unsigned long long* out;
        CK(cudaMalloc(&d, n * 8)); CK(cudaMalloc(&m, n / 8 + 64)); CK(cudaMalloc(&out, 64));
        CK(cudaMalloc(&g_partials, sizeof(AggRaw) * kSMs * 32)); CK(cudaMalloc(&g_ticket, 64)); CK(cudaMalloc(&g_agg, sizeof(AggRaw)));
        fill_kernel<<<148 * 8, 256>>>((uint64_t*)d, n, 7, 0); fill_kernel<<<148 * 8, 256>>>((uint64_t*)m, n / 64 + 1, 8, 0);
        CK(cudaMemset(g_ticket, 0, 64)); CK(cudaDeviceSynchronize());
        {
            auto kern = reduce_stats_kernel<int64_t, V16, true, false, 256, 4, 4>;
            float ms = time_ms([&] { kern<<<kSMs * 4, 256>>>(d, m, n, g_partials, g_ticket, g_agg, nullptr, XchgDev{}); });
            AggRaw h; CK(cudaMemcpy(&h, g_agg, sizeof h, cudaMemcpyDeviceToHost));
            printf("i64_masked_sum         %-12s block=256 U=4 vec=16B grid=592  %8.4f ms  %8.1f GB/s  rows=%llu sum=%llu cnt=%llu (library kernel)\n", "sum.ldg", ms,
                   (double)n * 8.125 / ms / 1e6, (unsigned long long)n, (unsigned long long)h.sum, (unsigned long long)h.count);
        }
        // tile sizes that divide 1e9 exactly where possible are not required: the kernel sums whole tiles only and prints rows.
        sum_variant<256, 4, 4>(d, m, n, out, 0); sum_variant<256, 4, 8>(d, m, n, out, 0); sum_variant<256, 6, 4>(d, m, n, out, 0);
        sum_variant<256, 8, 4>(d, m, n, out, 0); sum_variant<256, 3, 8>(d, m, n, out, 0); sum_variant<512, 4, 4>(d, m, n, out, 0);
        sum_variant<256, 4, 4>(d, m, n, out, 2); sum_variant<256, 4, 4>(d, m, n, out, 4); sum_variant<128, 8, 4>(d, m, n, out, 0);
        sum_variant<128, 4, 8>(d, m, n, out, 4); sum_variant<256, 12, 2>(d, m, n, out, 0); sum_variant<1024, 3, 4>(d, m, n, out, 0);
    }
    return 0;
}
