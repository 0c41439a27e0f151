// preprocess.cu -- gradient preprocessing, fixed-point scale selection, MultiRMSE.
//
// Reference semantics restated:
//   fitter.cpp:57-64 / :204-214   L2: build_grads = (g - mean) / (std + 1e-8), std with 1/(n-1); Cosine: raw g
//   math_ops.cpp:255-300          calculate_mean: T thread-partial sums over contiguous ELEMENT ranges of the
//                                 row-major matrix, merged in thread order, then * (1/n)
//   math_ops.cpp:407-513          calculate_var/std_and_center: same partition for sum (x-mean)^2, centers in place
//   math_ops.cpp:79-105           divide_mat_by_vec_inplace
//   utils.h:64-81                 T = calculate_num_threads(n_elements, par_th)  (emulated max = cfg.ref_threads)
//   loss.cpp:34-62                MultiRMSE gradients; each thread covers exactly n_elements/T elements, so the
//                                 trailing n_elements % T gradients are never written
//
// The reference's float sums are sequential chains; to reproduce their bits each (thread-partition, column)
// chain is evaluated by one CUDA thread in the same order (no FMA: the engine is compiled with -fmad=false).
#include "engine.cuh"
#include "chain.cuh"

namespace gb {

__host__ __device__ inline int calc_threads(long long total, int min_per_thread, int max_threads) {
    long long n = total / (min_per_thread > 0 ? min_per_thread : 1);
    if (n > total) n = total;
    if (n <= 1) return 1;
    if (n > max_threads) return max_threads;
    return (int)n;
}

// chains (t, col): partial[t*D + col] = sequential sum over i in [t*ept, (t+1)*ept or end), i % D == col
// mode 0: sum x          mode 1: sum (x-mean)^2 and center x in place
// One WARP per reference thread t: the 32 lanes load 32 consecutive elements (coalesced), then the elements are
// consumed strictly in memory order -- value j is broadcast with a shuffle and added by the lane that owns its
// column (lane = col % 32), so every per-column chain sees its elements in the reference's order while the loads
// stay coalesced.
__global__ void __launch_bounds__(128)
ref_chain_kernel(float *mat, const float *mean, float *partial, long long n_elements, int D, int T, int mode) {
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= T) return;
    const int lane = threadIdx.x & 31;
    const long long ept = n_elements / T;
    const long long s = (long long)t * ept, e = (t == T - 1) ? n_elements : s + ept;
    float acc0 = 0.0f, acc1 = 0.0f;              // columns lane and lane + 32
    int c0 = (int)(s % D);                       // column of element i0
    float raw_next = (s + lane < e) ? mat[s + lane] : 0.0f;
    for (long long i0 = s; i0 < e; i0 += 32) {
        const long long i = i0 + lane;
        const float raw = raw_next;
        raw_next = (i + 32 < e) ? mat[i + 32] : 0.0f;          // next group is in flight while this one is consumed
        float v = 0.0f;
        if (i < e) {
            v = raw;
            if (mode == 1) {
                const int col = (int)((c0 + lane) % D);
                const float c = raw - mean[col];
                mat[i] = c;
                v = c * c;
            }
        }
        const int cnt = (e - i0) < 32 ? (int)(e - i0) : 32;
        int col = c0;
        if (D == 1) {
#pragma unroll 8
            for (int j = 0; j < cnt; ++j) acc0 = acc0 + __shfl_sync(0xffffffffu, v, j);     // every lane runs the chain
        } else {
#pragma unroll 8
            for (int j = 0; j < cnt; ++j) {
                const float vj = __shfl_sync(0xffffffffu, v, j);
                if ((col & 31) == lane) { if (col < 32) acc0 = acc0 + vj; else acc1 = acc1 + vj; }
                ++col; if (col == D) col = 0;
            }
        }
        c0 = (int)((c0 + 32) % D);
    }
    if (lane < D) partial[(size_t)t * D + lane] = acc0;
    if (lane + 32 < D) partial[(size_t)t * D + lane + 32] = acc1;
}


// ---------------------------------------------------------------- parallel evaluation of the same chains (D <= 4)
// One 512-thread CTA per reference thread t.  The element range of t is streamed through shared memory in stages of
// 512*R rows (R = 8 / 4 / 2 rows per lane for D = 1 / 2 / 3-4; elements outside [s, e) are stored as +0, which a
// float chain ignores).  Per stage: phase A -- every warp summarises its own 32*R-row sub-block for every column
// chain in the chain's current binade (chain.cuh); phase B -- warp d walks the 16 summaries of column d in order,
// applying each one if it is valid for the actual running sum and otherwise running that sub-block as a plain
// sequential float chain.  The result is bit-identical to ref_chain_kernel (and to the reference's loop).
constexpr int DC_THREADS = 512, DC_WARPS = DC_THREADS / 32;

template <int D>
__global__ void __launch_bounds__(DC_THREADS, 1)
dense_chain_kernel(float *mat, const float *mean, float *partial, long long n_elements, int T, int mode, long long *stats) {
    constexpr int R = D == 1 ? 8 : D == 2 ? 4 : 2;
    constexpr int STAGE_ROWS = DC_THREADS * R, SUB_ROWS = 32 * R;
    constexpr int SE = STAGE_ROWS * D, EPT = R * D;      // elements per stage / per thread
    extern __shared__ __align__(16) float sg[];          // 3 stage buffers of DC_THREADS * 8 floats
    __shared__ seq::StageShared sh;
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long ept = n_elements / T;
    const long long s = (long long)t * ept, e = (t == T - 1) ? n_elements : s + ept;
    const long long base0 = (s / D) * D;                 // stage 0 starts at the row that holds element s
    const int n_stages = (int)((e - base0 + SE - 1) / SE);
    float nx1[8], nx2[8];                                // two stages in flight
    auto load = [&](int st, float (&dst)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < EPT) {
                const int o = j * DC_THREADS + tid;
                const long long idx = base0 + (long long)st * SE + o;
                float v = 0.0f;
                if (st < n_stages && idx >= s && idx < e) {
                    const float raw = mat[idx];
                    if (mode == 1) {
                        const float c = raw - mean[o % D];
                        mat[idx] = c;
                        v = c * c;
                    } else v = raw;
                }
                dst[j] = v;
            }
        }
    };
    auto commit = [&](int b, const float (&src)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < EPT) sg[b * (DC_THREADS * 8) + j * DC_THREADS + tid] = src[j];
    };
    // the lane's R consecutive elements of column chain c in sub-block w of stage buffer b
    auto load_x = [&](int b, int c, int w, float (&x)[R]) {
        const float *p = &sg[b * (DC_THREADS * 8) + (w * SUB_ROWS + lane * R) * D];
        float v[8];
        if (EPT == 8) {
            const float4 a = *reinterpret_cast<const float4 *>(p), q = *reinterpret_cast<const float4 *>(p + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = q.x; v[5] = q.y; v[6] = q.z; v[7] = q.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (j < EPT) ? p[j] : 0.0f;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float val = v[r * D];
#pragma unroll
            for (int dd = 1; dd < D; ++dd) val = (c == dd) ? v[r * D + dd] : val;
            x[r] = val;
        }
    };
    int n_fast = 0, n_adv = 0, n_seq = 0;
    seq::pipe_init(sh);
    load(0, nx1); commit(0, nx1);
    load(1, nx1); commit(1, nx1);
    load(2, nx1); load(3, nx2);
    __syncthreads();
    int bc = 0;                                          // buffer of the current stage (st % 3)
    for (int st = 0; st < n_stages; ++st) {
        const int bn = bc == 2 ? 0 : bc + 1, bf = bn == 2 ? 0 : bn + 1;
        auto lc = [&](int c, int w, float (&x)[R]) { load_x(bc, c, w, x); };
        auto ln = [&](int c, int w, float (&x)[R]) { load_x(bn, c, w, x); };
        seq::run_stage_pipe<R>(sh, D, st & 1, DC_WARPS, lc, st + 1 < n_stages ? DC_WARPS : 0, ln, n_fast, n_adv, n_seq);
        commit(bf, nx1);                                 // stage st + 2 (its buffer held stage st - 1)
#pragma unroll
        for (int j = 0; j < 8; ++j) nx1[j] = nx2[j];
        load(st + 4, nx2);
        bc = bn;
        __syncthreads();
    }
    if (tid < D) partial[(size_t)t * D + tid] = sh.state[tid];
    if (stats && lane == 0 && warp < D) {
        atomicAdd((unsigned long long *)&stats[0], (unsigned long long)n_fast);
        atomicAdd((unsigned long long *)&stats[1], (unsigned long long)n_adv);
        atomicAdd((unsigned long long *)&stats[2], (unsigned long long)n_seq);
    }
}

// One WARP per (reference thread, column) chain: 256 chain elements per step (8 consecutive per lane), summarised in the
// binade of the running sum and applied in one integer add, or advanced piecewise where the sum leaves its binade
// (chain.cuh warp_advance).  No barriers: for chains up to a few hundred thousand elements this beats the CTA pipeline
// above, in particular when the sum hovers around a power of two (zero-mean gradients).
template <int MODE>
__global__ void __launch_bounds__(32)
warp_chain_kernel(float *mat, const float *mean, float *partial, long long n_elements, int D, int T, long long *stats) {
    const int t = blockIdx.x / D, col = blockIdx.x - t * D, lane = threadIdx.x;
    const long long ept = n_elements / T;
    const long long s = (long long)t * ept, e = (t == T - 1) ? n_elements : s + ept;
    const long long first = s + ((col - (s % D)) + D) % D;       // first element of column `col` at or after s
    const long long cnt = first < e ? (e - first + D - 1) / D : 0;
    const float mu = MODE == 1 ? mean[col] : 0.0f;
    auto load = [&](long long blk, float (&x)[8]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long j = blk * 256 + lane * 8 + i;
            float v = 0.0f;
            if (j < cnt) {
                const long long idx = first + j * D;
                const float raw = mat[idx];
                if (MODE == 1) {
                    const float c = raw - mu;
                    mat[idx] = c;
                    v = c * c;
                } else v = raw;
            }
            x[i] = v;
        }
    };
    float acc = 0.0f;
    int n_seq = 0;
    const long long n_blk = (cnt + 255) / 256;
    // four blocks in flight (a lone warp per SM has nothing else to hide the load latency behind)
    float xa[8], xb[8], xc[8], xd[8];
    load(0, xa); load(1, xb); load(2, xc); load(3, xd);
    for (long long b = 0; b < n_blk; b += 4) {
        acc = seq::warp_advance<8>(acc, xa, n_seq); load(b + 4, xa);
        if (b + 1 < n_blk) { acc = seq::warp_advance<8>(acc, xb, n_seq); load(b + 5, xb); }
        if (b + 2 < n_blk) { acc = seq::warp_advance<8>(acc, xc, n_seq); load(b + 6, xc); }
        if (b + 3 < n_blk) { acc = seq::warp_advance<8>(acc, xd, n_seq); load(b + 7, xd); }
    }
    if (lane == 0) {
        partial[(size_t)t * D + col] = acc;
        if (stats) { atomicAdd((unsigned long long *)&stats[1], (unsigned long long)n_blk); atomicAdd((unsigned long long *)&stats[2], (unsigned long long)n_seq); }
    }
}

// ---------------------------------------------------------------- the same chains, summaries by the whole GPU
// (the scheme of replay_wide.cu for dense chains)  chain = (reference thread t, column), 256 chain elements per group:
//   wide_dense_sums_kernel  per-group fp64 sums (prediction only); mode 1 also centres the matrix in place
//   wide_dense_tabs_kernel  (after a per-chain prefix of those sums) group summaries for the predicted binade, tagged
//   wide_dense_walk_kernel  one warp per chain composes 32 summaries at a time while tag and range check hold for the
//                           ACTUAL running sum, any other group is advanced piecewise (warp_advance)
struct DenseWide {
    float *mat; const float *mean; float *partial;
    long long n_elements; int D, T, mode;
    double *bsum; float *pred; int4 *tab; float *tag;      // [chains][gmax]
    int gmax;
    // memoised sequential evaluation of the RISKY groups (see wide_dense_cand_kernel)
    int *ridx;             // [chains][gmax] row of the group in `cand`, -1: none
    int *rlist;            // [rcap] group (chain * gmax + g) of every row
    int *rcount;           // rows handed out
    float *cand;           // [rcap][2 * DW_K] chain value after the group for the start  pred + (k - DW_K) * ulp(pred)
    int rcap;
};
constexpr float DW_EMPTY = -1.0f;
constexpr int DW_K = 1024;             // candidate starts on each side of the predicted running sum

// the k-th candidate start of a group whose predicted running sum is `pred` (ulp `u` of its binade); used by the
// simulation and by the walk, which only accepts a candidate that equals its running sum BIT FOR BIT
__device__ __forceinline__ float dw_candidate(float pred, float u, int k) { return pred + (float)k * u; }

struct DenseChain { long long first, cnt; int ng; };
__device__ __forceinline__ DenseChain dense_chain_of(const DenseWide &P, int chain) {
    const int t = chain / P.D, col = chain - t * P.D;
    const long long ept = P.n_elements / P.T;
    const long long s = (long long)t * ept, e = (t == P.T - 1) ? P.n_elements : s + ept;
    DenseChain c;
    c.first = s + ((col - (s % P.D)) + P.D) % P.D;
    c.cnt = c.first < e ? (e - c.first + P.D - 1) / P.D : 0;
    c.ng = (int)((c.cnt + 255) / 256);
    return c;
}
// the lane's 8 consecutive chain elements of group g (mode 1: squares of the already centred values)
__device__ __forceinline__ void dense_group(const DenseWide &P, const DenseChain &c, int g, float (&x)[8]) {
    const int lane = threadIdx.x & 31;
    if (P.D == 1 && (long long)(g + 1) * 256 <= c.cnt) {
        // interior group of a contiguous chain: one address, eight loads (two LDG.128 when the chain start allows it)
        const float *p = P.mat + c.first + (long long)g * 256 + lane * 8;
        if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
            const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 4);
            x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = p[i];
        }
        if (P.mode == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = x[i] * x[i];
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long j = (long long)g * 256 + lane * 8 + i;
        float v = 0.0f;
        if (j < c.cnt) { v = P.mat[c.first + j * P.D]; if (P.mode == 1) v = v * v; }
        x[i] = v;
    }
}

__global__ void __launch_bounds__(256) wide_dense_sums_kernel(DenseWide P) {
    const int lane = threadIdx.x & 31;
    const int n_chains = P.T * P.D;
    const long long total = (long long)n_chains * P.gmax;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long u = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < total; u += warps) {
        const int chain = (int)(u / P.gmax), g = (int)(u - (long long)chain * P.gmax);
        const DenseChain c = dense_chain_of(P, chain);
        if (g >= c.ng) continue;
        const float mu = P.mode == 1 ? P.mean[chain % P.D] : 0.0f;
        double sum = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long j = (long long)g * 256 + i * 32 + lane;           // coalesced order; only the sum matters here
            if (j < c.cnt) {
                const long long idx = c.first + j * P.D;
                float v = P.mat[idx];
                if (P.mode == 1) { v = v - mu; P.mat[idx] = v; v = v * v; }    // math_ops.cpp:480-500: centre, then square
                sum += (double)v;
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            int lo = __double2loint(sum), hi = __double2hiint(sum);
            lo = __shfl_xor_sync(0xffffffffu, lo, o); hi = __shfl_xor_sync(0xffffffffu, hi, o);
            sum += __hiloint2double(hi, lo);
        }
        if (lane == 0) P.bsum[(size_t)chain * P.gmax + g] = sum;
    }
}

// one warp per chain: exclusive prefix of the group sums
__global__ void __launch_bounds__(32) wide_dense_prefix_kernel(DenseWide P) {
    const int chain = blockIdx.x, lane = threadIdx.x;
    const DenseChain c = dense_chain_of(P, chain);
    double carry = 0.0;
    for (int g0 = 0; g0 < c.ng; g0 += 32) {
        const int g = g0 + lane;
        const double v = g < c.ng ? P.bsum[(size_t)chain * P.gmax + g] : 0.0;
        double inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            int lo = __double2loint(inc), hi = __double2hiint(inc);
            lo = __shfl_up_sync(0xffffffffu, lo, o); hi = __shfl_up_sync(0xffffffffu, hi, o);
            if (lane >= o) inc += __hiloint2double(hi, lo);
        }
        if (g < c.ng) P.pred[(size_t)chain * P.gmax + g] = (float)(carry + inc - v);
        int lo = __double2loint(inc), hi = __double2hiint(inc);
        lo = __shfl_sync(0xffffffffu, lo, 31); hi = __shfl_sync(0xffffffffu, hi, 31);
        carry += __hiloint2double(hi, lo);
    }
}

__global__ void __launch_bounds__(256) wide_dense_tabs_kernel(DenseWide P) {
    const int lane = threadIdx.x & 31;
    const int n_chains = P.T * P.D;
    const long long total = (long long)n_chains * P.gmax;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long u = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < total; u += warps) {
        const int chain = (int)(u / P.gmax), g = (int)(u - (long long)chain * P.gmax);
        const DenseChain c = dense_chain_of(P, chain);
        if (g >= c.ng) continue;
        float x[8];
        dense_group(P, c, g, x);
        bool nz = false;
#pragma unroll
        for (int i = 0; i < 8; ++i) nz |= (x[i] != 0.0f) || (x[i] != x[i]);
        float inv_u, uu;
        const bool ok = seq::epoch_of(P.pred[(size_t)chain * P.gmax + g], inv_u, uu);
        float tagv = ok ? inv_u : 0.0f;
        int row = -1;
        if (!__any_sync(0xffffffffu, nz)) {
            tagv = DW_EMPTY;
            if (lane == 0) P.tab[(size_t)chain * P.gmax + g] = make_int4(0, 0, 0, 0);
        } else if (ok) {
            const seq::Tab tb = seq::warp_summarize<8>(x, inv_u);
            if (lane == 0) {
                P.tab[(size_t)chain * P.gmax + g] = make_int4(tb.a0, tb.a1, tb.mn, tb.mx);
                // risky: the summary does not hold for every running sum within DW_K ulps of the prediction (the sum passes a
                // power of two inside the group, or close to it) -> its outcome is memoised for the candidate starts instead
                const float pr = P.pred[(size_t)chain * P.gmax + g];
                const long long mm = (long long)(pr * inv_u);
                const long long lo = (1 << 23) + seq::MARGIN + DW_K, hi = (1 << 24) - seq::MARGIN - DW_K;
                const bool safe = mm > 0 ? (mm + tb.mn > lo && mm + tb.mx < hi) : (mm + tb.mx < -lo && mm + tb.mn > -hi);
                if (!safe && P.rcap > 0) {
                    const int r = atomicAdd(P.rcount, 1);
                    if (r < P.rcap) { row = r; P.rlist[r] = (int)((size_t)chain * P.gmax + g); }
                }
            }
        }
        if (lane == 0) { P.tag[(size_t)chain * P.gmax + g] = tagv; P.ridx[(size_t)chain * P.gmax + g] = row; }
    }
}

// Memoised sequential evaluation: for every risky group, the plain float chain over its 256 elements from each of the 2 * DW_K
// candidate starts around the predicted running sum -- embarrassingly parallel (one lane per candidate, elements broadcast from
// shared memory), ~0.5 M dependent adds per group.  The walk then replaces the 256 dependent adds of such a group by ONE lookup
// when its running sum is one of the candidates; by construction the stored value is what the sequential chain yields from
// exactly that start, so nothing about binades or rounding has to hold for it to be right.
__global__ void __launch_bounds__(256) wide_dense_cand_kernel(DenseWide P) {
    __shared__ __align__(16) float s_x[256];
    const int n_rows = min(*P.rcount, P.rcap);
    for (int r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const int cg = P.rlist[r];
        const int chain = cg / P.gmax, g = cg - chain * P.gmax;
        const DenseChain c = dense_chain_of(P, chain);
        {
            const long long j = (long long)g * 256 + threadIdx.x;
            float v = 0.0f;
            if (j < c.cnt) { v = P.mat[c.first + j * P.D]; if (P.mode == 1) v = v * v; }
            __syncthreads();
            s_x[threadIdx.x] = v;
            __syncthreads();
        }
        const float pr = P.pred[cg];
        float inv_u, u;
        seq::epoch_of(pr, inv_u, u);               // risky groups always have a binade (wide_dense_tabs_kernel)
        float acc[2 * DW_K / 256];
#pragma unroll
        for (int q = 0; q < 2 * DW_K / 256; ++q) acc[q] = dw_candidate(pr, u, (int)threadIdx.x + q * 256 - DW_K);
        const float4 *xs = reinterpret_cast<const float4 *>(s_x);
#pragma unroll 4
        for (int i = 0; i < 64; ++i) {
            const float4 e = xs[i];
#pragma unroll
            for (int q = 0; q < 2 * DW_K / 256; ++q) { acc[q] = acc[q] + e.x; acc[q] = acc[q] + e.y; acc[q] = acc[q] + e.z; acc[q] = acc[q] + e.w; }
        }
#pragma unroll
        for (int q = 0; q < 2 * DW_K / 256; ++q) P.cand[(size_t)r * 2 * DW_K + threadIdx.x + q * 256] = acc[q];
    }
}

__global__ void __launch_bounds__(32) wide_dense_walk_kernel(DenseWide P, long long *stats) {
    __shared__ __align__(16) float s_wbuf[256];
    const unsigned int full = 0xffffffffu;
    const int chain = blockIdx.x, lane = threadIdx.x;
    const DenseChain c = dense_chain_of(P, chain);
    const size_t base = (size_t)chain * P.gmax;
    float acc = 0.0f;
    int n_fast = 0, n_slow = 0, n_seq = 0;
    int4 qnx = make_int4(0, 0, 0, 0);
    float tgnx = 0.0f, prnx = 0.0f;
    int rinx = -1;
    if (lane < c.ng) { qnx = P.tab[base + lane]; tgnx = P.tag[base + lane]; rinx = P.ridx[base + lane]; prnx = P.pred[base + lane]; }
    float pa[8], pb[8];                                   // rows of two groups fetched ahead of need
    int ha = -1, hb = -1;                                 // which groups they are
#pragma unroll 1
    for (int w0 = 0; w0 < c.ng; w0 += 32) {
        const bool in_range = w0 + lane < c.ng;
        const int wn = min(32, c.ng - w0);
        const int4 q = qnx;
        const float tg = tgnx, prw = prnx;
        const int riw = rinx;
        rinx = -1;
        if (w0 + 32 + lane < c.ng) {
            qnx = P.tab[base + w0 + 32 + lane]; tgnx = P.tag[base + w0 + 32 + lane];
            rinx = P.ridx[base + w0 + 32 + lane]; prnx = P.pred[base + w0 + 32 + lane];
        }
        int first = 0;
#pragma unroll 1
        while (first < wn) {
            // longest applicable prefix of the window from `first` (same logic as replay_wide.cu compose_window)
            float inv_u, u;
            int take = 0;
            const bool live = lane >= first;
            if (!seq::epoch_of(acc, inv_u, u)) {
                const unsigned int stop = __ballot_sync(full, live && !(in_range && tg == DW_EMPTY));
                take = (stop ? (__ffs(stop) - 1) : 32) - first;
            } else {
                const bool empty = in_range && tg == DW_EMPTY;
                const bool tag_ok = in_range && (tg == inv_u || empty);
                int i0 = (tag_ok && live) ? q.x : 0, i1 = (tag_ok && live) ? q.y : 0;
                if (!__any_sync(full, i0 != i1)) {
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) { const int g0 = __shfl_up_sync(full, i0, off); if (lane >= off) i0 += g0; }
                    i1 = i0;
                } else {
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const int g0 = __shfl_up_sync(full, i0, off), g1 = __shfl_up_sync(full, i1, off);
                        if (lane >= off) {
                            const int n0 = g0 + ((g0 & 1) ? i1 : i0), n1 = g1 + (((g1 + 1) & 1) ? i1 : i0);
                            i0 = n0; i1 = n1;
                        }
                    }
                }
                int e0 = __shfl_up_sync(full, i0, 1);
                if (lane == 0) e0 = 0;
                const int m = (int)(acc * inv_u);
                const int lo = (1 << 23) + seq::MARGIN, hi = (1 << 24) - seq::MARGIN;
                const int b = m + e0;
                const bool okw = !live || empty || (tag_ok && (m > 0 ? (b + q.z > lo && b + q.w < hi) : (b + q.w < -lo && b + q.z > -hi)));
                const unsigned int bad = __ballot_sync(full, !okw);
                take = (bad ? (__ffs(bad) - 1) : 32) - first;
                if (take > 0) {
                    const int inc0 = __shfl_sync(full, i0, first + take - 1), inc1 = __shfl_sync(full, i1, first + take - 1);
                    acc = (float)(m + ((m & 1) ? inc1 : inc0)) * u;
                }
            }
            if (take > 0) { n_fast += take; first += take; }
            if (first >= wn) break;
            // rows of the failed group; the two groups after it are fetched now (a sum that hovers around a power of two
            // fails group after group, and a lone warp has nothing else to hide the load latency behind)
            const int gq = w0 + first;
            {
                // memoised outcome of this group for the candidate start that IS the running sum, if there is one
                const int row = __shfl_sync(full, riw, first);
                const float pr = __shfl_sync(full, prw, first);
                float pinv, pu;
                if (row >= 0 && seq::epoch_of(pr, pinv, pu)) {
                    const float kf = rintf((acc - pr) * pinv);
                    if (fabsf(kf) < (float)DW_K) {
                        const int k = (int)kf;
                        if (dw_candidate(pr, pu, k) == acc) {
                            acc = __ldcg(P.cand + (size_t)row * 2 * DW_K + (k + DW_K));
                            ++n_fast; ++first;
                            continue;
                        }
                    }
                }
            }
            float x[8];
            if (ha == gq) {
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = pa[i];
            } else if (hb == gq) {
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = pb[i];
            } else dense_group(P, c, gq, x);
            if (hb == gq + 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) pa[i] = pb[i];
                ha = gq + 1;
            } else if (ha != gq + 1 && gq + 1 < c.ng) { dense_group(P, c, gq + 1, pa); ha = gq + 1; }
            if (gq + 2 < c.ng) { dense_group(P, c, gq + 2, pb); hb = gq + 2; }
            acc = seq::warp_seq_block<8>(acc, x, s_wbuf, n_seq);
            ++n_slow; ++first;
        }
    }
    if (lane == 0) {
        P.partial[chain] = acc;
        if (stats) {
            atomicAdd((unsigned long long *)&stats[0], (unsigned long long)n_fast);
            atomicAdd((unsigned long long *)&stats[1], (unsigned long long)n_slow);
            atomicAdd((unsigned long long *)&stats[2], (unsigned long long)n_seq);
        }
    }
}


// ---------------------------------------------------------------- memo walk (mean chains)
// The mean chain of zero-mean gradients is a random walk: |s| ~ sigma sqrt(k), so the running sum touches a power of two in 40 % of
// the 256-element groups of a 62 500-element chain and every summary fails there (chains).  What does NOT depend on
// binades is the plain sequential chain itself, and how far the actual running sum is from its exact-arithmetic prediction is a
// slow random walk of rounding errors (~0.3 ulp per add: a few ulps per group, tens of ulps per chain).  So EVERY group is
// evaluated sequentially for 2 * MW_K candidate starts around its predicted incoming sum -- one lane per candidate, all groups of all
// chains at once: ~1 G independent adds, tens of microseconds for the whole GPU -- and the chain warp only looks its running sum
// up: one group per ~200 cycles instead of 256 dependent adds.  The segment of the table around the current deviation is fetched
// MW_PF groups ahead into registers, so that the lookup is one shuffle.  A start that is not in the table (or not on the
// candidate grid) falls back to the sequential chain, which is right by definition; a candidate is accepted only when it EQUALS the
// running sum bit for bit, so the result is the reference's sequential sum whatever the prediction was.
constexpr int MW_K = 256;             // candidates on each side of the prediction (the deviation is ~0.3 ulp x sqrt(adds): sigma 72 after 62 500)
constexpr int MW_SEG = 64;            // candidates per prefetched segment
constexpr int MW_PF = 6;              // groups of lookahead

// candidate spacing of a group: the ulp of the prediction's binade, or half of it when the candidates reach below that binade
// (sums below the power of two live on the finer grid; on the coarser side every grid point is still a candidate)
__device__ __forceinline__ bool mw_grid(float pred, float &step, float &inv_step) {
    float inv_u, u;
    if (!seq::epoch_of(pred, inv_u, u)) return false;
    const bool fine = fabsf(pred * inv_u) < 8388608.0f + (float)MW_K;
    step = fine ? 0.5f * u : u;
    inv_step = fine ? 2.0f * inv_u : inv_u;
    return true;
}
__device__ __forceinline__ float mw_candidate(float pred, float step, int k) { return pred + (float)k * step; }

// one CTA per group: 2 * MW_K sequential chains over the group's 256 elements (4 per thread, elements broadcast from shared memory)
__global__ void __launch_bounds__(256) dense_memo_sim_kernel(DenseWide P) {
    __shared__ __align__(16) float s_x[256];
    const int n_chains = P.T * P.D;
    const long long total = (long long)n_chains * P.gmax;
    for (long long cg = blockIdx.x; cg < total; cg += gridDim.x) {
        const int chain = (int)(cg / P.gmax), g = (int)(cg - (long long)chain * P.gmax);
        const DenseChain c = dense_chain_of(P, chain);
        if (g >= c.ng) continue;                       // uniform over the CTA
        float step, inv_step;
        const float pr = P.pred[cg];
        if (!mw_grid(pr, step, inv_step)) continue;    // no binade (start of the chain, tiny sums): the walk runs these sequentially
        {
            const long long j = (long long)g * 256 + threadIdx.x;
            float v = 0.0f;
            if (j < c.cnt) v = P.mat[c.first + j * P.D];
            __syncthreads();
            s_x[threadIdx.x] = v;
            __syncthreads();
        }
        constexpr int Q = 2 * MW_K / 256;
        float acc[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) acc[q] = mw_candidate(pr, step, (int)threadIdx.x + q * 256 - MW_K);
        const float4 *xs = reinterpret_cast<const float4 *>(s_x);
#pragma unroll 4
        for (int i = 0; i < 64; ++i) {
            const float4 e = xs[i];
#pragma unroll
            for (int q = 0; q < Q; ++q) { acc[q] = acc[q] + e.x; acc[q] = acc[q] + e.y; acc[q] = acc[q] + e.z; acc[q] = acc[q] + e.w; }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) P.cand[(size_t)cg * 2 * MW_K + threadIdx.x + q * 256] = acc[q];
    }
}

// one warp per chain.  Everything that does not depend on the running sum is prepared ahead: the per-group grid (prediction, candidate
// spacing) sits in shared memory, the table segments of the next MW_PF groups arrive through cp.async (a register ring was measured
// slower: 12 loads in flight share 6 scoreboards, so every lookup waited for the newest load).  What is left on the dependent path
// of a group is ~12 instructions and one shared-memory read.
__device__ __forceinline__ void mw_cp_async_4(unsigned int dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
constexpr int MW_GMAX = 513;          // longest chain the memo walk takes (groups)
__global__ void __launch_bounds__(32) dense_memo_walk_kernel(DenseWide P, long long *stats) {
    __shared__ __align__(16) float s_wbuf[256];
    __shared__ float4 s_g[MW_GMAX + 2 * MW_PF + 1];         // (prediction, step, 1 / step, has a grid) per group; zero beyond the chain
    __shared__ float s_seg[MW_PF][MW_SEG];
    const unsigned int full = 0xffffffffu;
    const int chain = blockIdx.x, lane = threadIdx.x;
    const DenseChain c = dense_chain_of(P, chain);
    const size_t base = (size_t)chain * P.gmax;
    for (int g = lane; g < MW_GMAX + 2 * MW_PF + 1; g += 32) {
        float4 G = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (g < c.ng) {
            G.x = P.pred[base + g];
            G.w = mw_grid(G.x, G.y, G.z) ? 1.0f : 0.0f;
        }
        s_g[g] = G;
    }
    __syncwarp();
    float acc = 0.0f;
    int n_fast = 0, n_slow = 0, n_seq = 0;
    int sk0[MW_PF];
    const unsigned int seg_base = (unsigned int)__cvta_generic_to_shared(&s_seg[0][0]) + lane * 4;
    // table segment of group gp around deviation `dev` (running sum - prediction, MW_PF groups earlier) -> ring slot j
    // (the copies are unconditional: a group without a grid reads the head of its own row, beyond the chain the head of the last row)
    const int g_last = c.ng > 0 ? c.ng - 1 : 0;
    auto prefetch = [&](int j, int &rk0, int gp, float dev) {
        const float4 G = s_g[gp];
        const float kc = rintf(dev * G.z);                   // G.z == 0 where there is no grid
        const int k0 = (int)fminf(fmaxf(kc - (float)(MW_SEG / 2), (float)-MW_K), (float)(MW_K - MW_SEG));
        const float *src = P.cand + (base + min(gp, g_last)) * 2 * MW_K + (k0 + MW_K) + lane;
        mw_cp_async_4(seg_base + j * (MW_SEG * 4), src);
        mw_cp_async_4(seg_base + j * (MW_SEG * 4) + 128, src + 32);
        asm volatile("cp.async.commit_group;" ::: "memory");
        rk0 = k0;
    };
#pragma unroll
    for (int j = 0; j < MW_PF; ++j) prefetch(j, sk0[j], j, 0.0f);
#pragma unroll 1
    for (int g0 = 0; g0 < c.ng; g0 += MW_PF) {
#pragma unroll
        for (int j = 0; j < MW_PF; ++j) {
            const int g = g0 + j;
            const float4 G = s_g[g];                         // beyond the chain: no grid, nothing to do
            asm volatile("cp.async.wait_group %0;" ::"n"(MW_PF - 1) : "memory");
            __syncwarp();
            const float dev = acc - G.x;                     // how far the running sum is from the exact-arithmetic prediction
            bool done = false;
            if (G.w != 0.0f) {
                const float kf = rintf(dev * G.z);
                if (fabsf(kf) < (float)MW_K && G.x + kf * G.y == acc) {      // the candidate IS the running sum, bit for bit
                    const int idx = (int)kf - sk0[j];
                    if ((unsigned int)idx < (unsigned int)MW_SEG) acc = s_seg[j][idx];
                    else acc = __ldcg(P.cand + (base + g) * 2 * MW_K + ((int)kf + MW_K));
                    done = true;
                    ++n_fast;
                }
            }
            if (!done && g < c.ng) {
                float x[8];
                dense_group(P, c, g, x);
                acc = seq::warp_seq_block<8>(acc, x, s_wbuf, n_seq);
                ++n_slow;
            }
            // the deviation drifts by a few ulps per group: centre the segment of group g + MW_PF on today's
            __syncwarp();
            prefetch(j, sk0[j], g + MW_PF, dev);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (lane == 0) {
        P.partial[chain] = acc;
        if (stats) {
            atomicAdd((unsigned long long *)&stats[0], (unsigned long long)n_fast);
            atomicAdd((unsigned long long *)&stats[1], (unsigned long long)n_slow);
            atomicAdd((unsigned long long *)&stats[2], (unsigned long long)n_seq);
        }
    }
}

constexpr size_t DC_SMEM = (size_t)3 * DC_THREADS * 8 * sizeof(float);

template <int D>
static void launch_dense(float *mat, const float *mean, float *partial, long long ne, int T, int mode, cudaStream_t s, long long *stats) {
    ensure_dyn_smem(dense_chain_kernel<D>, DC_SMEM);
    GB_LAUNCH(dense_chain_kernel<D>, T, DC_THREADS, DC_SMEM, s, mat, mean, partial, ne, T, mode, stats);
}

static void launch_ref_chain(Model *m, float *mat, const float *mean, float *partial, long long ne, int D, int T, int mode,
                             cudaStream_t s, long long *stats = nullptr) {
    if (ne <= 0) { GB_CUDA(cudaMemsetAsync(partial, 0, (size_t)T * D * sizeof(float), s)); return; }
    // GPU-wide summaries + one walking warp per chain
    static DevBuf fallback_scratch;                        // diag path (no model workspace)
    DevBuf &scratch = m ? m->ws.dwide : fallback_scratch;
    DenseWide P;
    P.mat = mat; P.mean = mean; P.partial = partial; P.n_elements = ne; P.D = D; P.T = T; P.mode = mode;
    const long long ept = ne / T;
    const long long longest = (ne - (long long)(T - 1) * ept + D - 1) / D + 1;      // the last thread takes the remainder
    P.gmax = (int)((longest + 255) / 256) + 1;
    const size_t per = (size_t)T * D * P.gmax;
    // (chains beyond ~130 000 elements keep the table walk below: it takes 32 clean groups per step, and long chains are mostly clean)
    if (mode == 0 && P.gmax <= MW_GMAX && per * 2 * MW_K * sizeof(float) <= ((size_t)1 << 30)) {
        // mean chains: memoised sequential evaluation of every group + a walk that only looks up (dense_memo_*_kernel)
        scratch.ensure(per * (sizeof(double) + sizeof(float) + 2 * MW_K * sizeof(float)) + 64);
        char *q = scratch.as<char>();
        P.bsum = reinterpret_cast<double *>(q); q += per * sizeof(double);
        P.cand = reinterpret_cast<float *>(q); q += per * 2 * MW_K * sizeof(float);
        P.pred = reinterpret_cast<float *>(q);
        P.tab = nullptr; P.tag = nullptr; P.ridx = nullptr; P.rlist = nullptr; P.rcount = nullptr; P.rcap = 0;
        long long un = (long long)per;
        int gr = (int)((un + 7) / 8);
        if (gr > 148 * 16) gr = 148 * 16;
        if (gr < 1) gr = 1;
        GB_LAUNCH(wide_dense_sums_kernel, gr, 256, 0, s, P);
        GB_LAUNCH(wide_dense_prefix_kernel, T * D, 32, 0, s, P);
        GB_LAUNCH(dense_memo_sim_kernel, (int)(un < 148 * 8 ? un : 148 * 8), 256, 0, s, P);
        GB_LAUNCH(dense_memo_walk_kernel, T * D, 32, 0, s, P, stats);
        return;
    }
    // memo rows: every group may be risky on a short chain; long inputs get a 64 MB budget (groups beyond it run sequentially)
    size_t rcap = per;
    if (rcap > ((size_t)64 << 20) / (2 * DW_K * sizeof(float))) rcap = ((size_t)64 << 20) / (2 * DW_K * sizeof(float));
    scratch.ensure(per * (sizeof(double) + sizeof(int4) + 2 * sizeof(float) + sizeof(int)) + rcap * (sizeof(int) + 2 * DW_K * sizeof(float)) + 64);
    char *p = scratch.as<char>();
    P.tab = reinterpret_cast<int4 *>(p); p += per * sizeof(int4);       // 16-byte entries first (alignment)
    P.bsum = reinterpret_cast<double *>(p); p += per * sizeof(double);
    P.cand = reinterpret_cast<float *>(p); p += rcap * 2 * DW_K * sizeof(float);
    P.pred = reinterpret_cast<float *>(p); p += per * sizeof(float);
    P.tag = reinterpret_cast<float *>(p); p += per * sizeof(float);
    P.ridx = reinterpret_cast<int *>(p); p += per * sizeof(int);
    P.rlist = reinterpret_cast<int *>(p); p += rcap * sizeof(int);
    P.rcount = reinterpret_cast<int *>(p);
    P.rcap = (int)rcap;
    GB_CUDA(cudaMemsetAsync(P.rcount, 0, sizeof(int), s));
    long long units = (long long)per;
    int grid = (int)((units + 7) / 8);
    if (grid > 148 * 16) grid = 148 * 16;
    if (grid < 1) grid = 1;
    GB_LAUNCH(wide_dense_sums_kernel, grid, 256, 0, s, P);
    GB_LAUNCH(wide_dense_prefix_kernel, T * D, 32, 0, s, P);
    GB_LAUNCH(wide_dense_tabs_kernel, grid, 256, 0, s, P);
    GB_LAUNCH(wide_dense_cand_kernel, 148 * 4, 256, 0, s, P);
    GB_LAUNCH(wide_dense_walk_kernel, T * D, 32, 0, s, P, stats);
}

// test hook: thread-partitioned chain sums of a host matrix through launch_ref_chain (impl 0) or the one-thread-per-
// chain reference kernel (impl 1)
void diag_chain_sums(const float *host_mat, long long ne, int D, int T, int mode, const float *host_mean, float *host_partial,
                     float *host_centered, int impl, double *info /* [4]: kernel ms, blocks fast, blocks advanced, lanes sequential */) {
    DevBuf mat, mean, partial, stats;
    mat.ensure((size_t)(ne > 0 ? ne : 1) * sizeof(float)); mean.ensure((size_t)D * sizeof(float)); partial.ensure((size_t)T * D * sizeof(float));
    stats.ensure(4 * sizeof(long long), true);
    GB_CUDA(cudaMemcpy(mat.p, host_mat, (size_t)ne * sizeof(float), cudaMemcpyHostToDevice));
    if (host_mean) GB_CUDA(cudaMemcpy(mean.p, host_mean, (size_t)D * sizeof(float), cudaMemcpyHostToDevice));
    GB_CUDA(cudaMemset(partial.p, 0, (size_t)T * D * sizeof(float)));
    cudaEvent_t e0, e1;
    GB_CUDA(cudaEventCreate(&e0)); GB_CUDA(cudaEventCreate(&e1));
    GB_CUDA(cudaDeviceSynchronize());
    GB_CUDA(cudaEventRecord(e0, 0));
    if (impl == 0) launch_ref_chain(nullptr, mat.as<float>(), mean.as<float>(), partial.as<float>(), ne, D, T, mode, 0, stats.as<long long>());
    else if (impl == 2) {                                  // the one-CTA-per-chain pipeline (chains of output_dim <= 4)
        if (D == 1) launch_dense<1>(mat.as<float>(), mean.as<float>(), partial.as<float>(), ne, T, mode, 0, stats.as<long long>());
        else if (D == 2) launch_dense<2>(mat.as<float>(), mean.as<float>(), partial.as<float>(), ne, T, mode, 0, stats.as<long long>());
        else if (D == 3) launch_dense<3>(mat.as<float>(), mean.as<float>(), partial.as<float>(), ne, T, mode, 0, stats.as<long long>());
        else if (D == 4) launch_dense<4>(mat.as<float>(), mean.as<float>(), partial.as<float>(), ne, T, mode, 0, stats.as<long long>());
        else throw Error("diag_chain_sums: impl 2 needs D <= 4");
    } else if (impl == 3) {                                // one warp per chain, no summaries
        if (mode == 0) GB_LAUNCH(warp_chain_kernel<0>, T * D, 32, 0, 0, mat.as<float>(), mean.as<float>(), partial.as<float>(), ne, D, T, stats.as<long long>());
        else GB_LAUNCH(warp_chain_kernel<1>, T * D, 32, 0, 0, mat.as<float>(), mean.as<float>(), partial.as<float>(), ne, D, T, stats.as<long long>());
    }
    else GB_LAUNCH(ref_chain_kernel, ceil_div(T, 4), 128, 0, 0, mat.as<float>(), mean.as<float>(), partial.as<float>(), ne, D, T, mode);
    GB_CUDA(cudaEventRecord(e1, 0));
    GB_CUDA(cudaDeviceSynchronize());
    float ms = 0.0f;
    GB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    GB_CUDA(cudaMemcpy(host_partial, partial.p, (size_t)T * D * sizeof(float), cudaMemcpyDeviceToHost));
    if (host_centered) GB_CUDA(cudaMemcpy(host_centered, mat.p, (size_t)ne * sizeof(float), cudaMemcpyDeviceToHost));
    if (info) {
        long long h[4];
        GB_CUDA(cudaMemcpy(h, stats.p, sizeof(h), cudaMemcpyDeviceToHost));
        info[0] = ms; info[1] = (double)h[0]; info[2] = (double)h[1]; info[3] = (double)h[2];
    }
}

// merge partials in thread order (d = 0..T*D-1, column d % D), then finish
// mode 0: mean = sum * (1/n)     mode 1: std = sqrtf(sum * 1/(n-1))
__global__ void ref_merge_kernel(const float *partial, float *out, int D, int T, int n_samples, int mode) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= D) return;
    float acc = 0.0f;
    if (T > 1) { for (int t = 0; t < T; ++t) acc = acc + partial[t * D + col]; }
    else acc = partial[col];
    if (mode == 0) out[col] = acc * (1.0f / (float)n_samples);
    else out[col] = sqrtf(acc * (1.0f / ((float)n_samples - 1.0f)));
}

__global__ void __launch_bounds__(256)
divide_and_max_kernel(float *bg, const float *stdv, Ctl *ctl, long long n_elements, int D, int do_divide, int which) {
    float mx = 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_elements; i += (long long)gridDim.x * blockDim.x) {
        float v = bg[i];
        if (do_divide) { v = v / (stdv[i % D] + 1e-8f); bg[i] = v; }
        const float a = fabsf(v);
        if (a > mx && a < INFINITY) mx = a;
        if (!(a < INFINITY) && which == 0) ctl->bg_nonfinite = 1;
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax(which == 0 ? &ctl->max_abs_bg : &ctl->max_abs_raw, __float_as_uint(mx));
}

__global__ void max_only_kernel(const float *g, Ctl *ctl, long long n_elements, int which) {
    float mx = 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_elements; i += (long long)gridDim.x * blockDim.x) {
        const float a = fabsf(g[i]);
        if (a > mx && a < INFINITY) mx = a;
        if (!(a < INFINITY) && which == 0) ctl->bg_nonfinite = 1;
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax(which == 0 ? &ctl->max_abs_bg : &ctl->max_abs_raw, __float_as_uint(mx));
}

__global__ void __launch_bounds__(256) quantize_bg_kernel(const float *__restrict__ bg, int2 *__restrict__ q, const Ctl *__restrict__ ctl, long long ne) {
    const float scale = exp2f((float)ctl->qexp);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ne; i += (long long)gridDim.x * blockDim.x) {
        const long long v = __float2ll_rn(bg[i] * scale);
        q[i] = make_int2((int)(v & ((1ll << LO_BITS) - 1)), (int)(v >> LO_BITS));
    }
}

// q = rint(g * 2^qexp) must satisfy |q| < 2^(Q_BITS-1)
__global__ void qexp_kernel(Ctl *ctl, int which) {
    const float mx = __uint_as_float(which == 0 ? ctl->max_abs_bg : ctl->max_abs_raw);
    int e = 0;
    if (mx > 0.0f) {
        int k;
        frexpf(mx, &k);            // mx = m * 2^k, m in [0.5, 1)  ->  mx < 2^k
        // build_grads go through the int32 histogram planes (Q_BITS); the raw gradients of the leaf values are summed
        // in int64 directly and can keep 40 bits
        e = ((which == 0 ? Q_BITS : 42) - 2) - k;
        if (e > 120) e = 120;
        if (e < -120) e = -120;
    }
    if (which == 0) ctl->qexp = e; else ctl->qexp_raw = e;
}

static void reset_max(Model &m, int which, cudaStream_t s) {
    Ctl *ctl = m.ws.ctl.as<Ctl>();
    GB_CUDA(cudaMemsetAsync(which == 0 ? &ctl->max_abs_bg : &ctl->max_abs_raw, 0, sizeof(unsigned int), s));
    if (which == 0) GB_CUDA(cudaMemsetAsync(&ctl->bg_nonfinite, 0, sizeof(int), s));
}

void column_mean_ref(Model &m, const float *mat, int N, int D, float *out_dev, cudaStream_t s) {
    Workspace &ws = m.ws;
    const long long ne = (long long)N * D;
    const int T = calc_threads(ne, m.cfg.par_th, m.cfg.ref_threads);
    ws.lrs.ensure((size_t)(T * D + 2 * D) * sizeof(float));
    float *partial = ws.lrs.as<float>();
    launch_ref_chain(&m, const_cast<float *>(mat), nullptr, partial, ne, D, T, 0, s);
    GB_LAUNCH(ref_merge_kernel, ceil_div(D, 64), 64, 0, s, partial, out_dev, D, T, N, 0);
}

void build_grads(Model &m, const float *grads, int N, cudaStream_t s) {
    Workspace &ws = m.ws;
    const int D = ws.D;
    const long long ne = (long long)N * D;
    ws.bg.ensure((size_t)(ne > 0 ? ne : 1) * sizeof(float));
    Ctl *ctl = ws.ctl.as<Ctl>();
    reset_max(m, 0, s);
    if (ne > 0) GB_CUDA(cudaMemcpyAsync(ws.bg.p, grads, (size_t)ne * sizeof(float), cudaMemcpyDeviceToDevice, s));
    int grid = (int)((ne + 2047) / 2048);
    if (grid > 1184) grid = 1184;
    if (grid < 1) grid = 1;
    if (m.cfg.split_score_func == GBRL_B200_SCORE_L2 && ne > 0) {
        const int T = calc_threads(ne, m.cfg.par_th, m.cfg.ref_threads);
        ws.lrs.ensure((size_t)(T * D + 2 * D) * sizeof(float));
        float *partial = ws.lrs.as<float>(), *mean = partial + (size_t)T * D, *stdv = mean + D;
        launch_ref_chain(&m, ws.bg.as<float>(), nullptr, partial, ne, D, T, 0, s);
        GB_LAUNCH(ref_merge_kernel, ceil_div(D, 64), 64, 0, s, partial, mean, D, T, N, 0);
        launch_ref_chain(&m, ws.bg.as<float>(), mean, partial, ne, D, T, 1, s);
        GB_LAUNCH(ref_merge_kernel, ceil_div(D, 64), 64, 0, s, partial, stdv, D, T, N, 1);
        GB_LAUNCH(divide_and_max_kernel, grid, 256, 0, s, ws.bg.as<float>(), stdv, ctl, ne, D, 1, 0);
    } else if (ne > 0) {
        GB_LAUNCH(max_only_kernel, grid, 256, 0, s, ws.bg.as<float>(), ctl, ne, 0);
    }
    GB_LAUNCH(qexp_kernel, 1, 1, 0, s, ctl, 0);
    // the fixed-point form the histogram pass adds (q = rint(g * 2^qexp) = hi * 2^18 + lo), converted once per tree
    ws.bgq.ensure((size_t)(ne > 0 ? ne : 1) * sizeof(int2));
    if (ne > 0) GB_LAUNCH(quantize_bg_kernel, grid, 256, 0, s, ws.bg.as<float>(), ws.bgq.as<int2>(), ctl, ne);
}

void raw_grad_scale(Model &m, const float *grads, int N, cudaStream_t s) {
    Workspace &ws = m.ws;
    const long long ne = (long long)N * ws.D;
    Ctl *ctl = ws.ctl.as<Ctl>();
    reset_max(m, 1, s);
    int grid = (int)((ne + 2047) / 2048);
    if (grid > 1184) grid = 1184;
    if (ne > 0) GB_LAUNCH(max_only_kernel, grid, 256, 0, s, grads, ctl, ne, 1);
    GB_LAUNCH(qexp_kernel, 1, 1, 0, s, ctl, 1);
}

// loss.cpp:34-62; only the first T*(n_elements/T) gradients are written
__global__ void multirmse_grads_kernel(const float *preds, const float *targets, float *grads, long long covered) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < covered; i += (long long)gridDim.x * blockDim.x)
        grads[i] = preds[i] - targets[i];
}

void multirmse_grads(Model &m, const float *preds, const float *targets, float *grads, int n, cudaStream_t s) {
    const long long ne = (long long)n * m.cfg.output_dim;
    const int T = calc_threads(ne, m.cfg.par_th, m.cfg.ref_threads);
    const long long covered = (ne / T) * T;
    if (covered <= 0) return;
    int grid = (int)((covered + 1023) / 1024);
    if (grid > 1184) grid = 1184;
    GB_LAUNCH(multirmse_grads_kernel, grid, 256, 0, s, preds, targets, grads, covered);
}

__global__ void __launch_bounds__(256) sq_err_kernel(const float *preds, const float *targets, double *out, long long ne) {
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ne; i += (long long)gridDim.x * blockDim.x) {
        const float gdiff = preds[i] - targets[i];
        acc += (double)(gdiff * gdiff);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// MultiRMSE::get_loss (loss.cpp:64-90): sqrt(0.5 * sum / n).  The sum is accumulated in fp64 (the
// reference's value is an fp32 thread-partitioned sum; agreement is to ~1e-6 relative, not bit level).
void multirmse_loss(Model &m, const float *preds, const float *targets, int n, float *loss_host, cudaStream_t s) {
    Workspace &ws = m.ws;
    const long long ne = (long long)n * m.cfg.output_dim;
    ws.pstage.ensure(sizeof(double));
    GB_CUDA(cudaMemsetAsync(ws.pstage.p, 0, sizeof(double), s));
    int grid = (int)((ne + 1023) / 1024);
    if (grid > 1184) grid = 1184;
    if (ne > 0) GB_LAUNCH(sq_err_kernel, grid, 256, 0, s, preds, targets, ws.pstage.as<double>(), ne);
    double h = 0.0;
    GB_CUDA(cudaMemcpyAsync(&h, ws.pstage.p, sizeof(double), cudaMemcpyDeviceToHost, s));
    GB_CUDA(cudaStreamSynchronize(s));
    *loss_host = sqrtf(0.5f * (float)h * (1.0f / (float)n));
}

}  // namespace gb
