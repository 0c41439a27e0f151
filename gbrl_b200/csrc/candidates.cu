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
// The exact per-column order statistics use cub::DeviceRadixSort per column (library call on a row that
// SURVEY 8f lists as "next"); everything else here is hand-written.
#include "engine.cuh"
#include <cstdlib>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_segmented_sort.cuh>
#include <cfloat>

namespace gb {

// transpose a 32-column slab of row-major X into column-major colbuf[c][i]
__global__ void transpose_slab_kernel(const float *__restrict__ X, float *__restrict__ cols, int N, int F, int f0, int nf) {
    __shared__ float tile[32][33];
    int i0 = blockIdx.x * 32;
    int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        int i = i0 + r, f = f0 + tx;
        tile[r][tx] = (i < N && tx < nf) ? X[(size_t)i * F + f] : 0.0f;
    }
    __syncthreads();
    for (int c = ty; c < 32; c += 8) {
        int i = i0 + tx;
        if (i < N && c < nf) cols[(size_t)c * N + i] = tile[tx][c];
    }
}

// split_candidate_generator.cpp:216-240: thr[f][b] = sorted[cum_b - 1]; blockIdx.y = column inside the sorted slab
__global__ void pick_quantiles_kernel(const float *__restrict__ sorted_cols, float *__restrict__ thr, int N, int B, int f0) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float *sorted_col = sorted_cols + (size_t)blockIdx.y * N;
    const int f = f0 + blockIdx.y;
    int actual = B + 1;
    int spb = N / actual, rem = N % actual;
    long long cum = (long long)(b + 1) * spb + (b + 1 < rem ? b + 1 : rem);
    long long idx = cum - 1;
    if (idx < 0) idx = 0;   // N < n_bins+1: the reference reads index -1 (UB); we clamp
    thr[(size_t)f * B + b] = sorted_col[idx];
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
        // exact order statistics of every column: sort a transposed 32-column slab.
        //   N <= 256K  : one cub::DeviceSegmentedSort call per slab (32 segments) -- a handful of launches per
        //                slab, which is what matters at RL batch sizes;
        //   larger N   : one device-wide cub::DeviceRadixSort per column.
        const bool segmented = N <= 262144;
        ws.colbuf[0].ensure((size_t)32 * N * sizeof(float));
        ws.colbuf[1].ensure((size_t)(segmented ? 32 : 1) * N * sizeof(float));
        size_t tmp = 0;
        if (segmented) {
            ws.sort_offsets.ensure(33 * sizeof(int));
            int h_off[33];
            for (int i = 0; i <= 32; ++i) h_off[i] = i * N;
            GB_CUDA(cudaMemcpyAsync(ws.sort_offsets.p, h_off, sizeof(h_off), cudaMemcpyHostToDevice, s));
            GB_CUDA(cudaStreamSynchronize(s));      // h_off is a stack buffer
            cub::DeviceSegmentedSort::SortKeys(nullptr, tmp, ws.colbuf[0].as<float>(), ws.colbuf[1].as<float>(), 32 * N, 32,
                                               ws.sort_offsets.as<int>(), ws.sort_offsets.as<int>() + 1, s);
        } else {
            cub::DeviceRadixSort::SortKeys(nullptr, tmp, ws.colbuf[0].as<float>(), ws.colbuf[1].as<float>(), N, 0, 32, s);
        }
        ws.sort_tmp.ensure(tmp);
        for (int f0 = 0; f0 < F; f0 += 32) {
            int nf = F - f0 < 32 ? F - f0 : 32;
            GB_LAUNCH(transpose_slab_kernel, ceil_div(N, 32), 256, 0, s, X, ws.colbuf[0].as<float>(), N, F, f0, nf);
            if (segmented) {
                size_t t2 = tmp;
                GB_CUDA(cub::DeviceSegmentedSort::SortKeys(ws.sort_tmp.p, t2, ws.colbuf[0].as<float>(), ws.colbuf[1].as<float>(), nf * N, nf,
                                                           ws.sort_offsets.as<int>(), ws.sort_offsets.as<int>() + 1, s));
                g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
                GB_LAUNCH(pick_quantiles_kernel, dim3(ceil_div(B, 256), nf), 256, 0, s, ws.colbuf[1].as<float>(), ws.thr.as<float>(), N, B, f0);
            } else {
                for (int c = 0; c < nf; ++c) {
                    size_t t2 = tmp;
                    GB_CUDA(cub::DeviceRadixSort::SortKeys(ws.sort_tmp.p, t2, ws.colbuf[0].as<float>() + (size_t)c * N,
                                                           ws.colbuf[1].as<float>(), N, 0, 32, s));
                    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
                    GB_LAUNCH(pick_quantiles_kernel, dim3(ceil_div(B, 256), 1), 256, 0, s, ws.colbuf[1].as<float>(), ws.thr.as<float>(), N, B, f0 + c);
                }
            }
        }
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
                     long long stride, int rows_per_cta) {
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
    dim3 grid(ceil_div(N, rows_per_cta), ntl);
    GB_LAUNCH(bin_kernel, grid, 256, 0, s, X, ws.thrT.as<float>(), ws.codes.as<uint16_t>(), N, F, ws.tile_lo, rows_per_cta);
    // the feature-major copy pays off once the fp32 matrix no longer sits in L2 (a PPO minibatch does)
    // (GBRL_B200_FEATMAJOR=1 / 0 forces it on / off: the parity tests run both paths on the same small inputs)
    const char *force = getenv("GBRL_B200_FEATMAJOR");
    ws.use_codesT = force ? (force[0] == '1') : ((size_t)N * F * sizeof(float) > ((size_t)48 << 20));
    if (ws.use_codesT) {
        ws.codesT_stride = ((long long)N + 7) & ~7ll;
        ws.codesT.ensure((size_t)F * ws.codesT_stride * sizeof(uint16_t) + 64);
        dim3 grid_t(ceil_div(N, rows_per_cta), ws.nT);
        GB_LAUNCH(bin_featmajor_kernel, grid_t, 256, 0, s, X, ws.thrT.as<float>(), ws.codesT.as<uint16_t>(), N, F, ws.codesT_stride, rows_per_cta);
    }
}

}  // namespace gb
