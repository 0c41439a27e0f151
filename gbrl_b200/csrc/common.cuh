// common.cuh -- shared definitions of the B200 fit/predict engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>
#include <atomic>

namespace gb {

// ---------------------------------------------------------------- constants
constexpr int FT = 32;            // features per histogram tile (one per lane: bank == feature)
constexpr int NB = 256;           // histogram bins per feature (codes 1..256; code 0 never goes right)
constexpr int ITEM_ROWS = 8192;   // max rows per histogram work item (bounds the int32 smem partial sums)
constexpr int LO_BITS = 18;       // fixed-point split: q = hi * 2^18 + lo, lo in [0, 2^18)
constexpr int Q_BITS = 34;        // |q| < 2^32: lo < 2^18 (8192 rows * 2^18 = 2^31 per fold), |hi| < 2^14 (8 items * 8192 rows * 2^14 = 2^30 per flush)
constexpr int FLUSH_ITEMS = 8;    // a CTA flushes its shared-memory histogram to HBM at the latest after this many items of one (node, tile)
constexpr int MAX_DEPTH_SUPPORTED = 12;
constexpr int MAX_OPTS = 64;
constexpr int HIST_THREADS = 512;
constexpr int CODE_SHIFT = 7;       // the code matrix stores code << 7 = byte offset of the code's row in a shared histogram plane
constexpr int SCAN_THREADS = 256;

// ---------------------------------------------------------------- errors
struct Error : public std::runtime_error {
    explicit Error(const std::string &m) : std::runtime_error(m) {}
};

#define GB_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            throw gb::Error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + __FILE__ + \
                            ":" + std::to_string(__LINE__) + " (" #expr ")");                           \
    } while (0)

#define GB_CHECK(cond, msg)                                  \
    do {                                                     \
        if (!(cond)) throw gb::Error(std::string(msg));      \
    } while (0)

// every kernel launch goes through this so that bench.py can report `gpu_launches`
extern std::atomic<long long> g_kernel_launches;
#define GB_LAUNCH(kernel, grid, block, smem, stream, ...)                 \
    do {                                                                  \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);       \
        gb::g_kernel_launches.fetch_add(1, std::memory_order_relaxed);    \
        GB_CUDA(cudaGetLastError());                                      \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device attribute: remembered per (kernel, device), so a
// second model on another GPU of the same process gets its opt-in too (capi.cu)
void ensure_dyn_smem_impl(const void *func, size_t bytes);
template <typename K> inline void ensure_dyn_smem(K kernel, size_t bytes) { ensure_dyn_smem_impl(reinterpret_cast<const void *>(kernel), bytes); }

// ---------------------------------------------------------------- device helpers
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div64(long long a, long long b) { return (a + b - 1) / b; }

// heap indexing of tree nodes: root 0, children of h are 2h+1 (left, x<=thr) and 2h+2 (right, x>thr)
__host__ __device__ inline int level_of(int h) {
    int l = 0;
    while (h >= (2 << l) - 1) ++l;
    return l;
}
__host__ __device__ inline int level_base(int l) { return (1 << l) - 1; }

// CTAs per SM of the grid-stride replay kernels (gather / side bits / summaries).  GBRL_B200_SIDE_GRID overrides it for the
// launches on a side stream (speculative levels): stream priorities act when a CTA is dispatched, so shorter-lived side CTAs
// could give way to the main stream sooner -- measured, it makes no difference (the histogram kernel shares issue slots, not
// dispatch order, with the replay).
int replay_grid_mult(bool side_stream);

#if defined(__CUDACC__)
__device__ __forceinline__ float warp_bcast(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ int warp_bcast(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }

__device__ __forceinline__ long long shfl_down_ll(long long v, int d) {
    int lo = __shfl_down_sync(0xffffffffu, (int)(v & 0xffffffffll), d);
    int hi = __shfl_down_sync(0xffffffffu, (int)(v >> 32), d);
    return ((long long)hi << 32) | (unsigned int)lo;
}

// 64-bit no-return add to global memory (REDG.E.ADD.64)
__device__ __forceinline__ void red_add64(long long *addr, long long v) {
    atomicAdd(reinterpret_cast<unsigned long long *>(addr), static_cast<unsigned long long>(v));
}

// streaming 128-bit / 64-bit loads that do not pollute L1 (rows are touched once per level)
__device__ __forceinline__ uint2 ld_nc_u2(const uint2 *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
#endif

}  // namespace gb
