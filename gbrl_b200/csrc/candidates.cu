// candidates.cu -- split-candidate thresholds and the code (candidate-bin) matrix.
//
// Reference semantics restated:
//   * Quantile thresholds: fitter.cpp:77-90 (per-column arg-sort) + split_candidate_generator.cpp:216-249
//     computeQuantiles: n_bins+1 equal-count bins (remainder handed out round-robin), threshold b of
//     feature f = value at sorted rank cum_b-1; the dedup branch at :241 is dead on CPU, so exactly
//     n_bins thresholds per feature are emitted, duplicates included.
//   * Uniform thresholds: split_candidate_generator.cpp:59-76: min + b*(max-min)/n_bins.
//   * A sample goes right iff x > thr (node.cpp:89,339).  With thresholds ascending per feature,
//     code(x) = #{j : thr_j < x} satisfies  x > thr_j  <=>  code(x) > j, so one u16 code per
//     (sample, feature) replaces every later float comparison of the histogram pass.
//
// The exact per-column order statistics come from a hand-written radix multi-select (no sort, no library call).
#include "engine.cuh"
#include <cstdlib>
#include <cfloat>

namespace gb {

// ---------------------------------------------------------------- exact multi-select (quantile thresholds)
// The B order statistics of every column, without sorting: a three-digit MSB radix select (11 + 11 + 10 bits of the
// order-preserving key of the float) run for all 256 ranks of all columns at once.
//   pass A  histogram of digit 1 of every element            (shared-memory counters, lane == column: coalesced rows)
//   scan A  per target rank: its digit-1 bin and the rank inside that bin; targets with the same bin share a "leader"
//   pass B  elements whose digit 1 is some target's bin: histogram of digit 2 in the leader's row
//   scan B  per target: digit 2 and the remaining rank
//   pass C  elements whose 22-bit prefix is some target's prefix: histogram of digit 3
//   scan C  per target: digit 3 -> the key -> the threshold (an exact data value, bit for bit what a sort would deliver)
// X is read three times (coalesced 128-byte row segments); only a few per cent of the elements reach passes B / C.
constexpr int SEL_D1 = 2048, SEL_D2 = 2048, SEL_D3 = 1024;
constexpr int SEL_THREADS = 512;
constexpr int SEL_MAP_STRIDE = SEL_D1 + 2;     // u16 per lane row of the digit-1 map: +1 word rotates the banks from lane to lane (ncu r02:
                                               // 1.2e8 bank conflicts of 1.6e8 wavefronts without it -- all lanes probe similar digits)

__device__ __forceinline__ uint32_t sel_key(float v) {
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);      // ascending unsigned order == ascending float order
}
__device__ __forceinline__ float sel_unkey(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
// 0-based rank of threshold b in the ascending column (split_candidate_generator.cpp:216-240)
__device__ __forceinline__ long long sel_rank(int b, int N, int B) {
    const int actual = B + 1;
    const int spb = N / actual, rem = N % actual;
    long long idx = (long long)(b + 1) * spb + (b + 1 < rem ? b + 1 : rem) - 1;
    if (idx < 0) idx = 0;   // N < n_bins+1: the reference reads index -1 (UB); we clamp
    return idx;
}

struct SelParams {
    const float *X; int N, F, B;
    unsigned int *hist1;        // [F][SEL_D1]
    unsigned int *hist2;        // [F][B][SEL_D2]   row of the leader target
    unsigned int *hist3;        // [F][B][SEL_D3]
    uint16_t *map1;             // [F][SEL_D1]  digit 1 -> leader target (0xffff: no target in this bin)
    unsigned int *key2;         // [F][B]       22-bit prefix of every target (ascending in b)
    int *lead1, *lead2;         // [F][B]       leader target of target b after digit 1 / digit 2
    long long *rank1, *rank2;   // [F][B]       remaining rank inside the bin
    float *thr;                 // [F][B]
};

// PASS 0: digit-1 histogram in shared memory (two columns per 32-bit word); PASS 1 / 2: filtered global histograms
template <int PASS>
__global__ void __launch_bounds__(SEL_THREADS, 1) sel_pass_kernel(SelParams P, int rows_per_cta) {
    extern __shared__ unsigned int sm[];
    const int slab = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = slab * 32 + lane;
    const bool has_col = col < P.F;
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(P.N, r0 + rows_per_cta);
    uint16_t *smap = reinterpret_cast<uint16_t *>(sm);                       // PASS >= 1: [32][SEL_MAP_STRIDE] u16
    unsigned int *skey = sm + 32 * SEL_MAP_STRIDE / 2;                       // PASS 2:    [32][B + 1] sorted 22-bit prefixes
    const int kstride = P.B + 1;
    if (PASS == 0) {
        for (int i = threadIdx.x; i < SEL_D1 * 16; i += SEL_THREADS) sm[i] = 0u;
    } else {
        for (int i = threadIdx.x; i < 32 * SEL_D1; i += SEL_THREADS) {
            const int c = i / SEL_D1, d = i - c * SEL_D1;
            smap[c * SEL_MAP_STRIDE + d] = (slab * 32 + c < P.F) ? P.map1[(size_t)(slab * 32 + c) * SEL_D1 + d] : (uint16_t)0xffff;
        }
        if (PASS == 2) {
            for (int i = threadIdx.x; i < 32 * P.B; i += SEL_THREADS) {
                const int c = i / P.B, b = i - c * P.B;
                skey[c * kstride + b] = (slab * 32 + c < P.F) ? P.key2[(size_t)(slab * 32 + c) * P.B + b] : 0xffffffffu;
            }
        }
    }
    __syncthreads();
    constexpr int NWARP = SEL_THREADS / 32, UNR = 4;
    for (int rowb = r0 + warp; rowb < r1; rowb += NWARP * UNR) {
        uint32_t kk[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {                      // independent loads first: the pass is a latency-bound stream
            const int row = rowb + u * NWARP;
            kk[u] = (has_col && row < r1) ? sel_key(P.X[(size_t)row * P.F + col]) : 0u;
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            if (!has_col || rowb + u * NWARP >= r1) continue;
            const uint32_t k = kk[u];
            const uint32_t d1 = k >> 21;
            if (PASS == 0) {
                atomicAdd(&sm[d1 * 16 + (lane >> 1)], (lane & 1) ? 0x10000u : 1u);
            } else {
                const unsigned int ld = smap[lane * SEL_MAP_STRIDE + d1];
                if (ld == 0xffffu) continue;
                if (PASS == 1) {
                    atomicAdd(&P.hist2[((size_t)col * P.B + ld) * SEL_D2 + ((k >> 10) & 0x7ffu)], 1u);
                } else {
                    // first target whose 22-bit prefix equals this element's (targets ascending): lower bound from the digit-1 leader on
                    const uint32_t pre = k >> 10;
                    const unsigned int *kc = skey + lane * kstride;
                    int lo = (int)ld, hi = P.B;
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (kc[mid] < pre) lo = mid + 1; else hi = mid; }
                    if (lo < P.B && kc[lo] == pre) atomicAdd(&P.hist3[((size_t)col * P.B + lo) * SEL_D3 + (k & 0x3ffu)], 1u);
                }
            }
        }
    }
    if (PASS == 0) {
        __syncthreads();
        for (int i = threadIdx.x; i < SEL_D1 * 16; i += SEL_THREADS) {
            const unsigned int w = sm[i];
            if (!w) continue;
            const int d = i >> 4, c = slab * 32 + (i & 15) * 2;
            if ((w & 0xffffu) && c < P.F) atomicAdd(&P.hist1[(size_t)c * SEL_D1 + d], w & 0xffffu);
            if ((w >> 16) && c + 1 < P.F) atomicAdd(&P.hist1[(size_t)(c + 1) * SEL_D1 + d], w >> 16);
        }
    }
}

// exclusive prefix of `bins` counters (<= 8 per thread of a 256-thread CTA) into shared memory; returns nothing
template <int BINS>
__device__ __forceinline__ void block_prefix_256(const unsigned int *__restrict__ h, long long *s_pre /* [BINS + 1] */) {
    __shared__ long long s_w[8];
    constexpr int PER = BINS / 256;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    unsigned int v[PER];
    long long loc = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) { v[j] = h[t * PER + j]; loc += v[j]; }
    long long inc = loc;
    for (int o = 1; o < 32; o <<= 1) {
        const int hi = __shfl_up_sync(0xffffffffu, (int)(inc >> 32), o);
        const unsigned int lo = (unsigned int)__shfl_up_sync(0xffffffffu, (int)(inc & 0xffffffffll), o);
        if (lane >= o) inc += ((long long)hi << 32) | lo;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    long long base = 0;
    for (int w = 0; w < warp; ++w) base += s_w[w];
    long long run = base + inc - loc;
#pragma unroll
    for (int j = 0; j < PER; ++j) { s_pre[t * PER + j] = run; run += v[j]; }
    if (t == 255) s_pre[BINS] = run;
    __syncthreads();
}

// bin whose [pre[d], pre[d+1]) holds rank r
template <int BINS>
__device__ __forceinline__ int find_bin(const long long *s_pre, long long r) {
    int lo = 0, hi = BINS - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (s_pre[mid] <= r) lo = mid; else hi = mid - 1; }
    return lo;
}

// scan A: one CTA per column
__global__ void __launch_bounds__(256) sel_scan1_kernel(SelParams P) {
    __shared__ long long s_pre[SEL_D1 + 1];
    __shared__ int s_d1[NB];
    const int col = blockIdx.x;
    block_prefix_256<SEL_D1>(P.hist1 + (size_t)col * SEL_D1, s_pre);
    for (int i = threadIdx.x; i < SEL_D1; i += 256) P.map1[(size_t)col * SEL_D1 + i] = 0xffff;
    const int b = threadIdx.x;
    int d1 = 0;
    if (b < P.B) {
        const long long r = sel_rank(b, P.N, P.B);
        d1 = find_bin<SEL_D1>(s_pre, r);
        P.rank1[(size_t)col * P.B + b] = r - s_pre[d1];
        s_d1[b] = d1;
    }
    __syncthreads();
    if (b < P.B) {
        int ld = b;
        while (ld > 0 && s_d1[ld - 1] == d1) --ld;           // ranks ascend, so equal bins are adjacent
        P.lead1[(size_t)col * P.B + b] = ld;
        P.key2[(size_t)col * P.B + b] = (unsigned int)d1 << 11;
        if (ld == b) P.map1[(size_t)col * SEL_D1 + d1] = (uint16_t)b;
    }
}

// scan B: one CTA per (column, target)
__global__ void __launch_bounds__(256) sel_scan2_kernel(SelParams P) {
    __shared__ long long s_pre[SEL_D2 + 1];
    const int col = blockIdx.x / P.B, b = blockIdx.x - col * P.B;
    const size_t o = (size_t)col * P.B + b;
    block_prefix_256<SEL_D2>(P.hist2 + ((size_t)col * P.B + P.lead1[o]) * SEL_D2, s_pre);
    if (threadIdx.x == 0) {
        const long long r = P.rank1[o];
        const int d2 = find_bin<SEL_D2>(s_pre, r);
        P.rank2[o] = r - s_pre[d2];
        P.key2[o] |= (unsigned int)d2;
    }
}

// leaders after digit 2 (equal 22-bit prefixes are adjacent); one thread per (column, target)
__global__ void sel_lead2_kernel(SelParams P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.F * P.B) return;
    const int col = i / P.B, b = i - col * P.B;
    const unsigned int *k = P.key2 + (size_t)col * P.B;
    int ld = b;
    while (ld > 0 && k[ld - 1] == k[b]) --ld;
    P.lead2[i] = ld;
}

// scan C: one CTA per (column, target) -> the threshold
__global__ void __launch_bounds__(256) sel_scan3_kernel(SelParams P) {
    __shared__ long long s_pre[SEL_D3 + 1];
    const int col = blockIdx.x / P.B, b = blockIdx.x - col * P.B;
    const size_t o = (size_t)col * P.B + b;
    block_prefix_256<SEL_D3>(P.hist3 + ((size_t)col * P.B + P.lead2[o]) * SEL_D3, s_pre);
    if (threadIdx.x == 0) {
        const int d3 = find_bin<SEL_D3>(s_pre, P.rank2[o]);
        P.thr[o] = sel_unkey((P.key2[o] << 10) | (unsigned int)d3);
    }
}

// split_candidate_generator.cpp:59-76 (uniform): one block per feature
__global__ void uniform_thresholds_kernel(const float *__restrict__ X, float *__restrict__ thr, int N, int F, int B) {
    int f = blockIdx.x;
    float mx = -INFINITY, mn = INFINITY;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        float v = X[(size_t)i * F + f];
        mx = fmaxf(mx, v);   // NaN-free data assumed (reference: v > max / v < min comparisons)
        mn = fminf(mn, v);
    }
    __shared__ float smx[32], smn[32];
    for (int o = 16; o; o >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if ((threadIdx.x & 31) == 0) { smx[threadIdx.x >> 5] = mx; smn[threadIdx.x >> 5] = mn; }
    __syncthreads();
    if (threadIdx.x < 32) {
        int nw = blockDim.x >> 5;
        mx = threadIdx.x < nw ? smx[threadIdx.x] : -INFINITY;
        mn = threadIdx.x < nw ? smn[threadIdx.x] : INFINITY;
        for (int o = 16; o; o >>= 1) {
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        }
        if (threadIdx.x == 0) { smx[0] = mx; smn[0] = mn; }
    }
    __syncthreads();
    mx = smx[0]; mn = smn[0];
    float step = (mx - mn) / (float)B;
    for (int b = threadIdx.x; b < B; b += blockDim.x) thr[(size_t)f * B + b] = mn + (float)b * step;
}

// thr[f][b] -> thrT[tile][b(256, +inf padded)][32]   (bin-major so that lane == feature == smem bank)
__global__ void tile_thresholds_kernel(const float *__restrict__ thr, float *__restrict__ thrT, int F, int B, int nT) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int total = nT * NB * FT;
    if (idx >= total) return;
    int fs = idx % FT, b = (idx / FT) % NB, t = idx / (FT * NB);
    int f = t * FT + fs;
    thrT[idx] = (f < F && b < B) ? thr[(size_t)f * B + b] : INFINITY;
}

void compute_thresholds(Model &m, const float *X, int N, int F, cudaStream_t s) {
    Workspace &ws = m.ws;
    int B = m.cfg.n_bins;
    ws.thr.ensure((size_t)F * B * sizeof(float));
    ws.thrT.ensure((size_t)ws.nT * NB * FT * sizeof(float));
    if (m.cfg.generator_type == GBRL_B200_GEN_UNIFORM) {
        GB_LAUNCH(uniform_thresholds_kernel, F, 256, 0, s, X, ws.thr.as<float>(), N, F, B);
    } else {
        // exact order statistics of every column by a three-pass radix multi-select (no sort, no library call)
        const size_t FB = (size_t)F * B;
        const size_t bytes = (size_t)F * SEL_D1 * 4 + FB * SEL_D2 * 4 + FB * SEL_D3 * 4 + (size_t)F * SEL_D1 * 2 + FB * 4 * 3 + FB * 8 * 2 + 256;
        ws.sort_tmp.ensure(bytes);
        char *p = ws.sort_tmp.as<char>();
        SelParams P;
        P.X = X; P.N = N; P.F = F; P.B = B; P.thr = ws.thr.as<float>();
        P.rank1 = reinterpret_cast<long long *>(p); p += FB * 8;
        P.rank2 = reinterpret_cast<long long *>(p); p += FB * 8;
        P.hist1 = reinterpret_cast<unsigned int *>(p); p += (size_t)F * SEL_D1 * 4;
        P.hist2 = reinterpret_cast<unsigned int *>(p); p += FB * SEL_D2 * 4;
        P.hist3 = reinterpret_cast<unsigned int *>(p); p += FB * SEL_D3 * 4;
        P.key2 = reinterpret_cast<unsigned int *>(p); p += FB * 4;
        P.lead1 = reinterpret_cast<int *>(p); p += FB * 4;
        P.lead2 = reinterpret_cast<int *>(p); p += FB * 4;
        P.map1 = reinterpret_cast<uint16_t *>(p);
        GB_CUDA(cudaMemsetAsync(P.hist1, 0, (size_t)F * SEL_D1 * 4 + FB * SEL_D2 * 4 + FB * SEL_D3 * 4, s));
        const int slabs = ceil_div(F, 32);
        // rows per CTA: <= 65535 (16-bit shared counters of pass A), enough CTAs to fill the GPU
        int ctas = ceil_div(ws.n_sms * 2, slabs);
        if (ctas < ceil_div(N, 60000)) ctas = ceil_div(N, 60000);
        if (ctas > ceil_div(N, 64)) ctas = ceil_div(N, 64);
        if (ctas < 1) ctas = 1;
        const int rows_per_cta = ceil_div(N, ctas);
        dim3 grid(ceil_div(N, rows_per_cta), slabs);
        const size_t sm0 = (size_t)SEL_D1 * 16 * 4, sm1 = (size_t)32 * SEL_MAP_STRIDE * 2, sm2 = sm1 + (size_t)32 * (B + 1) * 4;
        ensure_dyn_smem(sel_pass_kernel<0>, sm0); ensure_dyn_smem(sel_pass_kernel<1>, sm1); ensure_dyn_smem(sel_pass_kernel<2>, sm2);
        GB_LAUNCH(sel_pass_kernel<0>, grid, SEL_THREADS, sm0, s, P, rows_per_cta);
        GB_LAUNCH(sel_scan1_kernel, F, 256, 0, s, P);
        GB_LAUNCH(sel_pass_kernel<1>, grid, SEL_THREADS, sm1, s, P, rows_per_cta);
        GB_LAUNCH(sel_scan2_kernel, (int)FB, 256, 0, s, P);
        GB_LAUNCH(sel_lead2_kernel, ceil_div((int)FB, 256), 256, 0, s, P);
        GB_LAUNCH(sel_pass_kernel<2>, grid, SEL_THREADS, sm2, s, P, rows_per_cta);
        GB_LAUNCH(sel_scan3_kernel, (int)FB, 256, 0, s, P);
    }
    int total = ws.nT * NB * FT;
    GB_LAUNCH(tile_thresholds_kernel, ceil_div(total, 256), 256, 0, s, ws.thr.as<float>(), ws.thrT.as<float>(), F, B, ws.nT);
    m.have_candidates = true;
}

// ---------------------------------------------------------------- binning
// One CTA = 32 rows x one 32-feature tile per iteration; thread (r = tid>>3, g = tid&7) owns features
// 4g..4g+3 of row r (one float4 / four scalar loads) and looks each of them up in the tile's 256
// ascending thresholds held in shared memory as thrT[b][32]: lane <-> feature is a bijection inside a
// warp (rows are rotated by (k + r) & 3), so every LDS of the search is bank-conflict free.
__global__ void __launch_bounds__(256)
bin_kernel(const float *__restrict__ X, const float *__restrict__ thrT, uint16_t *__restrict__ codes, int N, int F,
           int tile_lo, int rows_per_cta) {
    __shared__ float sthr[NB * FT];
    const int tile = tile_lo + blockIdx.y;
    const float *tt = thrT + (size_t)tile * NB * FT;
    for (int i = threadIdx.x; i < NB * FT; i += blockDim.x) sthr[i] = tt[i];
    __syncthreads();
    const int r = threadIdx.x >> 3, g = threadIdx.x & 7, rl = r & 3;
    const int fbase = tile * FT + g * 4;
    const bool vec = ((F & 3) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
    const int row0 = blockIdx.x * rows_per_cta;
    const int row1 = min(N, row0 + rows_per_cta);
    uint16_t *out = codes + (size_t)blockIdx.y * N * FT;   // local tile index
    for (int row = row0 + r; row < row1; row += 32) {
        float x[4];
        if (vec && fbase + 3 < F) {
            float4 v = *reinterpret_cast<const float4 *>(X + (size_t)row * F + fbase);
            x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) x[k] = (fbase + k < F) ? X[(size_t)row * F + fbase + k] : -INFINITY;
        }
        unsigned int c[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int mslot = (k + rl) & 3;
            const float xv = mslot == 0 ? x[0] : mslot == 1 ? x[1] : mslot == 2 ? x[2] : x[3];
            const int fs = g * 4 + mslot;
            int pos = 0;
#pragma unroll
            for (int st = 128; st >= 1; st >>= 1)
                if (sthr[(pos + st - 1) * FT + fs] < xv) pos += st;
            if (sthr[pos * FT + fs] < xv) pos += 1;   // only possible when pos == 255
            if (mslot == 0) c[0] = pos; else if (mslot == 1) c[1] = pos; else if (mslot == 2) c[2] = pos; else c[3] = pos;
        }
        uint2 w;
        w.x = (c[0] << CODE_SHIFT) | (c[1] << (16 + CODE_SHIFT));      // stored pre-scaled: see hist_kernel
        w.y = (c[2] << CODE_SHIFT) | (c[3] << (16 + CODE_SHIFT));
        *reinterpret_cast<uint2 *>(out + (size_t)row * FT + g * 4) = w;
    }
}

// Feature-major copy of the codes for ALL features (every rank), codesT[f][row] = code (not pre-scaled), row stride
// codesT_stride: the consumers that need ONE feature of many rows -- the stable partition (node.cpp:86-96) and the
// side bits of the near-tie replay (node.cpp:339) -- read 2 coalesced bytes per row instead of a 32-byte sector of
// the row-major fp32 matrix.  x > thr[f][j]  <=>  code(x) > j, so the comparisons are the reference's.
__global__ void __launch_bounds__(256)
bin_featmajor_kernel(const float *__restrict__ X, const float *__restrict__ thrT, uint16_t *__restrict__ codesT, int N, int F,
                     long long stride, int rows_per_cta, uint16_t *__restrict__ codes, int tile_lo, int tile_hi) {
    __shared__ float sthr[NB * FT];
    __shared__ uint16_t s_c[FT][36];
    const int tile = blockIdx.y;
    const float *tt = thrT + (size_t)tile * NB * FT;
    for (int i = threadIdx.x; i < NB * FT; i += blockDim.x) sthr[i] = tt[i];
    __syncthreads();
    const int r = threadIdx.x >> 3, g = threadIdx.x & 7, rl = r & 3;
    const int fbase = tile * FT + g * 4;
    const bool vec = ((F & 3) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
    const int row0 = blockIdx.x * rows_per_cta;
    const int row1 = min(N, row0 + rows_per_cta);
    for (int rb = row0; rb < row1; rb += 32) {
        const int row = rb + r;
        unsigned int c[4] = {0u, 0u, 0u, 0u};
        if (row < row1) {
            float x[4];
            if (vec && fbase + 3 < F) {
                float4 v = *reinterpret_cast<const float4 *>(X + (size_t)row * F + fbase);
                x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) x[k] = (fbase + k < F) ? X[(size_t)row * F + fbase + k] : -INFINITY;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int mslot = (k + rl) & 3;
                const float xv = mslot == 0 ? x[0] : mslot == 1 ? x[1] : mslot == 2 ? x[2] : x[3];
                const int fs = g * 4 + mslot;
                int pos = 0;
#pragma unroll
                for (int st = 128; st >= 1; st >>= 1)
                    if (sthr[(pos + st - 1) * FT + fs] < xv) pos += st;
                if (sthr[pos * FT + fs] < xv) pos += 1;   // only possible when pos == 255
                if (mslot == 0) c[0] = pos; else if (mslot == 1) c[1] = pos; else if (mslot == 2) c[2] = pos; else c[3] = pos;
            }
        }
        // the same codes in the tile-major layout of the histogram pass (pre-scaled), for the tiles this rank owns: X is read once
        if (codes != nullptr && tile >= tile_lo && tile < tile_hi && row < row1) {
            uint2 w;
            w.x = (c[0] << CODE_SHIFT) | (c[1] << (16 + CODE_SHIFT));
            w.y = (c[2] << CODE_SHIFT) | (c[3] << (16 + CODE_SHIFT));
            *reinterpret_cast<uint2 *>(codes + ((size_t)(tile - tile_lo) * N + row) * FT + g * 4) = w;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) s_c[g * 4 + k][r] = (uint16_t)c[k];
        __syncthreads();
        {   // thread -> (feature, 4 consecutive rows): one 8-byte store
            const int ft = threadIdx.x >> 3, rg = (threadIdx.x & 7) * 4;
            if (tile * FT + ft < F && rb + rg < row1) {
                const uint2 w = *reinterpret_cast<const uint2 *>(&s_c[ft][rg]);
                *reinterpret_cast<uint2 *>(codesT + (size_t)(tile * FT + ft) * stride + rb + rg) = w;
            }
        }
        __syncthreads();
    }
}

void bin_features(Model &m, const float *X, int N, int F, cudaStream_t s) {
    Workspace &ws = m.ws;
    int ntl = ws.tile_hi - ws.tile_lo;
    ws.codes.ensure((size_t)ntl * N * FT * sizeof(uint16_t));
    if (ntl <= 0 || N == 0) return;
    int rows_per_cta = 1024;
    // the feature-major copy pays off once the fp32 matrix no longer sits in L2 (a PPO minibatch does)
    // (GBRL_B200_FEATMAJOR=1 / 0 forces it on / off: the parity tests run both paths on the same small inputs)
    const char *force = getenv("GBRL_B200_FEATMAJOR");
    ws.use_codesT = force ? (force[0] == '1') : ((size_t)N * F * sizeof(float) > ((size_t)48 << 20));
    if (ws.use_codesT) {
        // one pass over X writes both layouts
        ws.codesT_stride = ((long long)N + 7) & ~7ll;
        ws.codesT.ensure((size_t)F * ws.codesT_stride * sizeof(uint16_t) + 64);
        dim3 grid_t(ceil_div(N, rows_per_cta), ws.nT);
        GB_LAUNCH(bin_featmajor_kernel, grid_t, 256, 0, s, X, ws.thrT.as<float>(), ws.codesT.as<uint16_t>(), N, F, ws.codesT_stride, rows_per_cta,
                  ws.codes.as<uint16_t>(), ws.tile_lo, ws.tile_hi);
    } else {
        dim3 grid(ceil_div(N, rows_per_cta), ntl);
        GB_LAUNCH(bin_kernel, grid, 256, 0, s, X, ws.thrT.as<float>(), ws.codes.as<uint16_t>(), N, F, ws.tile_lo, rows_per_cta);
    }
}

}  // namespace gb
