// Kernel-variant sweep on the device (tuning tool, not part of the library).
//   nvcc <lib flags> [-DMNR_LD_POLICY=k -DMNR_ST_POLICY=k] tools/sweep.cu -o tools/sweep_ld<k>_st<k>
// Times the hot kernels at the BASELINE shapes over (block size, loads in flight, vector width, launch bounds,
// grid = full | resident wave), CUDA events, median of 15 after 3 warm-ups, inputs >> L2.  Every variant's output is
// checksummed against the first variant of its group, so a fast-but-wrong variant is flagged.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../minarrow_b200/csrc/ew_kernels.cuh"
#include "../minarrow_b200/csrc/reduce_kernels.cuh"

using namespace mnr;

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void fill_kernel(uint64_t* p, uint64_t n, uint64_t seed, int as_double) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t z = (i + seed) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        if (as_double) { double d = (double)(int64_t)(z >> 11) * (1.0 / 9007199254740992.0) * 200.0 - 100.0; p[i] = (uint64_t)__double_as_longlong(d); }
        else p[i] = z;
    }
}
__global__ void checksum_kernel(const uint64_t* p, uint64_t n, unsigned long long* out) {
    unsigned long long acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        acc += p[i] * (2 * i + 1);
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

static const char* kGroup = "";
static unsigned long long g_ref = 0; static bool g_have_ref = false;
static void report(const char* name, int block, int u, int vecbytes, int minb, const char* gridmode, unsigned grid, int regs,
                   int occ, double bytes, float ms, unsigned long long cs) {
    if (!g_have_ref) { g_ref = cs; g_have_ref = true; }
    printf("%-22s %-14s block=%3d U=%d vec=%2dB minb=%d grid=%-8s(%7u) regs=%3d occ=%2d  %8.4f ms  %8.1f GB/s  %s\n", kGroup, name, block,
           u, vecbytes, minb, gridmode, grid, regs, occ, ms, bytes / ms / 1e6, cs == g_ref ? "ok" : "CHECKSUM-MISMATCH");
    fflush(stdout);
}

struct EwBufs { void *x, *y, *o; uint8_t *mx, *my, *om; uint64_t n; unsigned int* flag; };

template <typename VecT, int CLS, int BLOCK, int U, int MINB>
static void ew_variant(const char* name, const EwBufs& b, int op, bool scalar, bool two_masks, double bytes) {
    using K = void (*)(EwDev);
    K kern = ew_binary_kernel<double, double, double, VecT, CLS, true, BLOCK, U, MINB>;
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BLOCK, 0));
    constexpr int VEC = sizeof(VecT) / 8;
    EwDev a{}; a.lhs = b.x; a.rhs = scalar ? nullptr : b.y; double s = 2.5; memcpy(&a.scalar_bits, &s, 8);
    a.lmask = b.mx; a.rmask = two_masks ? b.my : nullptr; a.mask_or = 0; a.out = b.o; a.out_mask = b.om; a.n = b.n;
    a.div0_flag = b.flag; a.op = op;
    const uint64_t nvec = (b.n + VEC - 1) / VEC, tiles = (nvec + 32ull * U - 1) / (32ull * U);
    const uint64_t full = (tiles + BLOCK / 32 - 1) / (BLOCK / 32);
    for (int mode = 0; mode < 2; ++mode) {
        const unsigned grid = mode == 0 ? (unsigned)full : (unsigned)std::min<uint64_t>(full, (uint64_t)kSMs * occ);
        CK(cudaMemset(b.o, 0, b.n * 8)); CK(cudaMemset(b.om, 0, b.n / 8));
        float ms = time_ms([&] { kern<<<grid, BLOCK>>>(a); });
        unsigned long long cs = checksum(b.o, b.n * 8) ^ checksum(b.om, b.n / 8);
        report(name, BLOCK, U, (int)sizeof(VecT), MINB, mode == 0 ? "full" : "resident", grid, fa.numRegs, occ, bytes, ms, cs);
    }
}

struct RedBufs { const int64_t* d; const uint8_t* m; uint64_t n; AggRaw* partials; unsigned int* ticket; AggRaw* out; };

template <typename VecT, int BLOCK, int MINB, int U, bool MINMAX>
static void red_variant(const char* name, const RedBufs& b, int blocks_per_sm) {
    auto kern = reduce_stats_kernel<int64_t, VecT, true, MINMAX, BLOCK, MINB, U>;
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BLOCK, 0));
    const unsigned grid = (unsigned)kSMs * (blocks_per_sm > 0 ? blocks_per_sm : occ);
    CK(cudaMemset(b.ticket, 0, 64));
    float ms = time_ms([&] { kern<<<grid, BLOCK>>>(b.d, b.m, b.n, b.partials, b.ticket, b.out); });
    AggRaw h; CK(cudaMemcpy(&h, b.out, sizeof h, cudaMemcpyDeviceToHost));
    report(name, BLOCK, U, (int)sizeof(VecT), MINB, blocks_per_sm > 0 ? "fixed" : "resident", grid, fa.numRegs, occ, (double)b.n * 8.125, ms,
           h.sum * 31 + h.count + (MINMAX ? h.mn * 7 + h.mx * 3 : 0));
}

int main(int argc, char** argv) {
    const char* only = argc > 1 ? argv[1] : "";
    printf("# sweep: MNR_LD_POLICY=%d (%s) MNR_ST_POLICY=%d (%s)\n", MNR_LD_POLICY, MNR_LD, MNR_ST_POLICY, MNR_ST);
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("# device: %s, %d SMs, sm_%d%d\n", p.name, p.multiProcessorCount, p.major, p.minor);

    if (!*only || strstr(only, "ew")) {
        EwBufs b{}; b.n = 1ull << 28;
        CK(cudaMalloc(&b.x, b.n * 8)); CK(cudaMalloc(&b.y, b.n * 8)); CK(cudaMalloc(&b.o, b.n * 8));
        CK(cudaMalloc(&b.mx, b.n / 8)); CK(cudaMalloc(&b.my, b.n / 8)); CK(cudaMalloc(&b.om, b.n / 8)); CK(cudaMalloc(&b.flag, 64));
        fill_kernel<<<148 * 8, 256>>>((uint64_t*)b.x, b.n, 1, 1); fill_kernel<<<148 * 8, 256>>>((uint64_t*)b.y, b.n, 2, 1);
        fill_kernel<<<148 * 8, 256>>>((uint64_t*)b.mx, b.n / 64, 3, 0); fill_kernel<<<148 * 8, 256>>>((uint64_t*)b.my, b.n / 64, 4, 0);
        CK(cudaDeviceSynchronize());
        const double B2 = (double)b.n * 24.375, BS = (double)b.n * 16.25;
#define EW(V, BL, U, MB) ew_variant<V, CLS_CHEAP, BL, U, MB>("add2m", b, MNR_ADD, false, true, B2)
        kGroup = "f64_masked_add_2masks"; g_have_ref = false;
        EW(V16, 256, 4, 1); EW(V16, 256, 4, 4); EW(V16, 256, 2, 1); EW(V16, 256, 2, 4); EW(V16, 256, 8, 1); EW(V16, 256, 8, 2);
        EW(V16, 128, 4, 1); EW(V16, 128, 4, 8); EW(V16, 128, 8, 1); EW(V16, 512, 4, 1); EW(V16, 512, 4, 2); EW(V16, 512, 2, 2);
        EW(V32, 256, 2, 1); EW(V32, 256, 2, 4); EW(V32, 256, 4, 1); EW(V32, 256, 4, 2); EW(V32, 128, 2, 1); EW(V32, 128, 4, 1);
        EW(V32, 512, 2, 1); EW(V32, 512, 2, 2); EW(V32, 256, 1, 4); EW(V32, 512, 1, 2);
#undef EW
#define EWS(V, BL, U, MB) ew_variant<V, CLS_CHEAP, BL, U, MB>("scalar_mul", b, MNR_MUL, true, false, BS)
        kGroup = "f64_masked_scalar_mul"; g_have_ref = false;
        EWS(V16, 256, 4, 1); EWS(V16, 256, 8, 1); EWS(V16, 256, 8, 4); EWS(V16, 256, 16, 1); EWS(V16, 128, 8, 1); EWS(V16, 512, 8, 1);
        EWS(V32, 256, 2, 1); EWS(V32, 256, 4, 1); EWS(V32, 256, 4, 4); EWS(V32, 256, 8, 1); EWS(V32, 512, 4, 1); EWS(V32, 128, 4, 1);
#undef EWS
#define EWD(V, BL, U, MB) ew_variant<V, CLS_DIV, BL, U, MB>("div2m", b, MNR_DIV, false, true, B2)
        kGroup = "f64_masked_div_2masks"; g_have_ref = false;
        EWD(V16, 256, 4, 1); EWD(V16, 256, 2, 4); EWD(V16, 256, 2, 1); EWD(V16, 512, 2, 2); EWD(V32, 256, 2, 1); EWD(V32, 256, 1, 4); EWD(V16, 128, 2, 8);
#undef EWD
        cudaFree(b.x); cudaFree(b.y); cudaFree(b.o); cudaFree(b.mx); cudaFree(b.my); cudaFree(b.om);
    }
    if (!*only || strstr(only, "red")) {
        RedBufs r{}; r.n = 1000000000ull;
        int64_t* d; uint8_t* m;
        CK(cudaMalloc(&d, r.n * 8)); CK(cudaMalloc(&m, r.n / 8 + 64)); CK(cudaMalloc(&r.partials, sizeof(AggRaw) * kSMs * 32));
        CK(cudaMalloc(&r.ticket, 64)); CK(cudaMalloc(&r.out, sizeof(AggRaw)));
        fill_kernel<<<148 * 8, 256>>>((uint64_t*)d, r.n, 7, 0); fill_kernel<<<148 * 8, 256>>>((uint64_t*)m, r.n / 64 + 1, 8, 0);
        CK(cudaDeviceSynchronize());
        r.d = d; r.m = m;
        kGroup = "i64_masked_sum_1e9"; g_have_ref = false;
#define RD(V, BL, MB, U, BPS) red_variant<V, BL, MB, U, false>("sum", r, BPS)
        RD(V16, 256, 4, 4, 4); RD(V16, 256, 4, 4, 0); RD(V16, 256, 4, 8, 0); RD(V16, 256, 2, 8, 0); RD(V16, 256, 8, 2, 0); RD(V16, 256, 6, 4, 0);
        RD(V16, 512, 2, 4, 0); RD(V16, 512, 4, 2, 0); RD(V16, 128, 8, 4, 0); RD(V16, 128, 12, 4, 0);
        RD(V32, 256, 4, 2, 0); RD(V32, 256, 4, 4, 0); RD(V32, 256, 2, 4, 0); RD(V32, 512, 2, 2, 0); RD(V32, 128, 8, 2, 0); RD(V32, 256, 8, 1, 0);
        RD(V16, 256, 4, 4, 8); RD(V16, 256, 4, 4, 16);
#undef RD
        kGroup = "i64_masked_stats_1e9"; g_have_ref = false;
#define RM(V, BL, MB, U, BPS) red_variant<V, BL, MB, U, true>("sum+minmax", r, BPS)
        RM(V16, 256, 4, 4, 4); RM(V16, 256, 4, 4, 0); RM(V16, 256, 2, 8, 0); RM(V32, 256, 4, 2, 0); RM(V32, 256, 2, 4, 0);
#undef RM
    }
    return 0;
}
