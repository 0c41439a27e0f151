// replay_wide.cu -- near-tie replay with the float chains spread over the WHOLE GPU (output_dim <= 2).
//
// replay_stream_kernel (split.cu) gives every (node, candidate) item one CTA; a chain over a million rows is then
// bounded by what one SM can summarise.  Here the summaries (chain.cuh) of all 256-row groups of all items are
// produced by all SMs at once, for a PREDICTED binade of the running sum, and one warp per chain only walks them:
//   wide_bits_kernel    side bits of every (item, row) + per-group sums of every chain (fp64; prediction only)
//   wide_prefix_kernel  exclusive prefix of the group sums per item and chain = predicted running sum at group start
//   wide_tabs_kernel    group summaries (Tab) for the binade of that prediction, tagged with the binade
//   wide_walk_kernel    per chain: composes 32 summaries at a time while their tag equals the binade of the ACTUAL
//                       running sum and their range check holds; any other group is advanced piecewise from the
//                       streams (warp_advance).  The prediction never enters the result -- a wrong tag only costs
//                       time -- so the sums are bit-identical to the reference's sequential chain (node.cpp:336-352).
// Cosine needs a second round (tabs + walk) for the mat_vec_dot_sum chains once the side means are known.
#include "replay.cuh"
#include "chain.cuh"
#include "spec_chain.cuh"

namespace gb {

constexpr int GROUP_ROWS = 256;

__device__ __forceinline__ double shfl_xor_d(double v, int m) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, m); hi = __shfl_xor_sync(0xffffffffu, hi, m);
    return __hiloint2double(hi, lo);
}

// A warp owns a contiguous range of 8-word groups (256 rows each) of the planes: the item of the first group is
// found by one binary search, after that the item index only moves forward.
template <int D>
__global__ void __launch_bounds__(256) wide_bits_kernel(ReplayParams P, NodeArrays na, StreamParams S, WideParams Wd) {
    const int n_items = min(P.ctl->n_replay, S.replay_cap);
    if (n_items <= 0) return;
    const int lane = threadIdx.x & 31;
    const int total_groups = S.woff[n_items] >> 3;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int per = (total_groups + warps - 1) / warps;
    const int g_begin = min(total_groups, wid * per), g_end = min(total_groups, g_begin + per);
    if (g_begin >= g_end) return;
    int it;
    {
        const int gw = g_begin << 3;
        int lo = 0, hi = n_items;                 // last item whose offset is <= gw
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (S.woff[mid] <= gw) lo = mid + 1; else hi = mid;
        }
        it = lo - 1;
    }
    int it_loaded = -1, s0 = 0, n = 0, jb = 0, f = 0, woff_it = 0, md = 1;
    bool is_cand = false;
    float tv = INFINITY;
    const uint16_t *col = P.codesT;
    for (int g = g_begin; g < g_end; ++g) {
        const int gw = g << 3;
        while (it + 1 < n_items && S.woff[it + 1] <= gw) ++it;      // items without words share the offset of their successor
        if (it != it_loaded) {
            const ReplayItem item = P.items[it];
            s0 = na.seg_start[item.node]; n = na.seg_len[item.node];
            is_cand = item.cand >= 0;
            f = is_cand ? item.cand / P.B : 0;
            jb = is_cand ? item.cand - f * P.B : 0;
            tv = is_cand ? P.thr[item.cand] : INFINITY;
            if (P.codesT != nullptr) col = P.codesT + (size_t)f * P.codesT_stride + P.row_offset;
            woff_it = S.woff[it]; md = S.mode[it];
            it_loaded = it;
        }
        if (lane == 0) Wd.gitem[g] = md == 0 ? it : -1;
        if (md != 0) continue;
        const int k0 = (gw - woff_it) << 5;
        int cnt = 0;
        bool rt[8];
        if (P.codesT != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = k0 + j * 32 + lane;
                rt[j] = (k < n && is_cand) && (int)col[P.order[s0 + k]] > jb;           // x > thr[f][jb] <=> code > jb
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = k0 + j * 32 + lane;
                rt[j] = (k < n && is_cand) && P.X[(size_t)P.order[s0 + k] * P.F + f] > tv;
            }
        }
        // group sums in fp32 (tree order); they only steer the binade prediction
        float sl[D], sr[D];
#pragma unroll
        for (int d = 0; d < D; ++d) { sl[d] = 0.0f; sr[d] = 0.0f; }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = k0 + j * 32 + lane;
            const bool right = rt[j];                                          // node.cpp:339
            const unsigned int m = __ballot_sync(0xffffffffu, right);
            if (lane == j) S.bits[gw + j] = m;
            cnt += __popc(m);
            if (k < n) {
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const float gv = S.G[(size_t)(s0 + k) * D + d];
                    if (right) sr[d] += gv; else sl[d] += gv;
                }
            }
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { sl[d] += __shfl_xor_sync(0xffffffffu, sl[d], o); sr[d] += __shfl_xor_sync(0xffffffffu, sr[d], o); }
        }
        if (lane == 0) {
#pragma unroll
            for (int d = 0; d < D; ++d) { Wd.bsum[(size_t)g * 2 * D + d] = (double)sl[d]; Wd.bsum[(size_t)g * 2 * D + D + d] = (double)sr[d]; }
            if (cnt) atomicAdd(&S.nright[it], cnt);
        }
    }
}

// one CTA per item: pred[group][chain] = sum of bsum over the earlier groups of the item
template <int D>
__global__ void __launch_bounds__(256) wide_prefix_kernel(ReplayParams P, NodeArrays na, StreamParams S, WideParams Wd) {
    __shared__ double s_tot[256];
    const int n_items = min(P.ctl->n_replay, S.replay_cap);
    const int t = threadIdx.x;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        if (S.mode[it] != 0) continue;
        const int n = na.seg_len[P.items[it].node];
        const int ng = (n + GROUP_ROWS - 1) / GROUP_ROWS;
        const size_t base = (size_t)(S.woff[it] >> 3);
        const int per = (ng + 255) / 256;
        const int g0 = min(ng, t * per), g1 = min(ng, g0 + per);
        for (int c = 0; c < 2 * D; ++c) {
            double loc = 0.0;
            for (int g = g0; g < g1; ++g) loc += Wd.bsum[(base + g) * 2 * D + c];
            s_tot[t] = loc;
            __syncthreads();
            for (int o = 1; o < 256; o <<= 1) {
                const double v = t >= o ? s_tot[t - o] : 0.0;
                __syncthreads();
                s_tot[t] += v;
                __syncthreads();
            }
            double acc = s_tot[t] - loc;
            __syncthreads();
            for (int g = g0; g < g1; ++g) {
                Wd.pred[(base + g) * 2 * D + c] = (float)acc;
                acc += Wd.bsum[(base + g) * 2 * D + c];
            }
        }
    }
}

// chain elements of one 256-row group of an item, in the reference's order (node.cpp:341-350; math_ops.h:432-449 for the
// Cosine dot chains): PASS 0 -> chain (side, d): g[row][d] of the rows on that side, +0 for the others;
// PASS 1 -> chain side: g[row][d] * mean[side][d] over (row, d), rows of the other side +0.
template <int D, int PASS>
struct SideElems {
    const float *g;            // G + (s0 + first row of the group) * D
    const unsigned int *w;     // side-bit words of the group
    int side, d;
    float m[D];                // PASS 1: the side's means
    __device__ __forceinline__ float operator()(int k) const {
        if (PASS == 0) {
            const bool right = (w[k >> 5] >> (k & 31)) & 1u;
            return (right == (side != 0)) ? g[(size_t)k * D + d] : 0.0f;
        } else {
            const int row = k / D, dd = k - row * D;
            const bool right = (w[row >> 5] >> (row & 31)) & 1u;
            float mv = m[0];
#pragma unroll
            for (int q = 1; q < D; ++q) mv = (dd == q) ? m[q] : mv;
            return (right == (side != 0)) ? g[k] * mv : 0.0f;
        }
    }
};

// one LANE per (group, chain): simulates the group's float chain from candidate starts next to the predicted running sum
template <int D, int PASS>
__global__ void __launch_bounds__(128) wide_sim_kernel(ReplayParams P, NodeArrays na, StreamParams S, WideParams Wd) {
    constexpr int NCH = PASS == 0 ? 2 * D : 2;
    const int n_items = min(P.ctl->n_replay, S.replay_cap);
    if (n_items <= 0) return;
    if (PASS == 1 && P.score_func == GBRL_B200_SCORE_L2) return;
    const long long total = (long long)(S.woff[n_items] >> 3) * NCH;
    for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(u / NCH), c = (int)(u - (long long)g * NCH);
        const int it = Wd.gitem[g];
        if (it < 0) continue;
        const ReplayItem item = P.items[it];
        const int s0 = na.seg_start[item.node], n = na.seg_len[item.node];
        const int lg = g - (S.woff[it] >> 3);
        const int rows = min(GROUP_ROWS, n - lg * GROUP_ROWS);
        if (rows <= 0) continue;
        SideElems<D, PASS> el;
        el.g = S.G + ((size_t)s0 + (size_t)lg * GROUP_ROWS) * D;
        el.w = S.bits + S.woff[it] + lg * 8;
        float pr;
        if (PASS == 0) {
            el.side = c / D; el.d = c - el.side * D;
            pr = Wd.pred[(size_t)g * 2 * D + c];
        } else {
            el.side = c; el.d = 0;
            pr = 0.0f;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                el.m[d] = Wd.fin[(size_t)it * 8 + c * D + d];
                pr += el.m[d] * Wd.pred[(size_t)g * 2 * D + c * D + d];
            }
        }
        spec::Head hd;
        spec::Cand cd[spec::J];
        spec::sim_group(pr, PASS == 0 ? rows : rows * D, el, hd, cd);
        const size_t o = (size_t)g * 2 * D + c;
        Wd.head[o] = hd;
        if (spec::head_flags(hd) & spec::F_CANDS) {
#pragma unroll
            for (int j = 0; j < spec::J; ++j) Wd.cand[o * spec::J + j] = cd[j];
        }
    }
}

// one CTA per item, warp c walks chain c; then the item's score (L2, or Cosine after PASS 1) / the side means (Cosine, PASS 0)
template <int D, int PASS>
__global__ void __launch_bounds__(32 * 2 * D) wide_walk_kernel(ReplayParams P, NodeArrays na, StreamParams S, WideParams Wd, Ctl *ctl_stats) {
    constexpr int NCH = PASS == 0 ? 2 * D : 2;
    constexpr int EPR = PASS == 0 ? 1 : D;                      // chain elements per row
    __shared__ float s_sum[2 * D];
    __shared__ float s_mean[2 * D];
    __shared__ __align__(16) float s_wbuf[NCH][256 * EPR];      // per chain warp: staging of a group that is run sequentially
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_items = min(P.ctl->n_replay, S.replay_cap);
    if (PASS == 1 && P.score_func == GBRL_B200_SCORE_L2) return;
    int n_groups = 0, n_err = 0, n_seq = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        if (S.mode[it] != 0) continue;
        const ReplayItem item = P.items[it];
        const int h = item.node, cand = item.cand;
        const int s0 = na.seg_start[h], n = na.seg_len[h];
        const int ng = (n + GROUP_ROWS - 1) / GROUP_ROWS;
        const size_t base = (size_t)(S.woff[it] >> 3);
        if (PASS == 1) {
            if (threadIdx.x < 2 * D) s_mean[threadIdx.x] = Wd.fin[(size_t)it * 8 + threadIdx.x];
            __syncthreads();
        }
        if (warp < NCH) {
            const int c = warp;
            SideElems<D, PASS> el;
            if (PASS == 0) { el.side = c / D; el.d = c - el.side * D; }
            else {
                el.side = c; el.d = 0;
#pragma unroll
                for (int d = 0; d < D; ++d) el.m[d] = s_mean[c * D + d];
            }
            auto load_head = [&](int g) { return Wd.head[(base + g) * 2 * D + c]; };
            auto load_cand = [&](int g, int j) { return Wd.cand[((base + g) * 2 * D + c) * spec::J + j]; };
            auto seq_group = [&](int g, float a) -> float {
                // the reference's plain sequential chain over the group (node.cpp:341-350 / math_ops.h:432-449)
                SideElems<D, PASS> e2 = el;
                e2.g = S.G + ((size_t)s0 + (size_t)g * GROUP_ROWS) * D;
                e2.w = S.bits + S.woff[it] + g * 8;
                const int cnt = min(GROUP_ROWS, n - g * GROUP_ROWS) * EPR;
                float x[8 * EPR];                                 // the lane's 8 * EPR consecutive chain elements
#pragma unroll
                for (int i = 0; i < 8 * EPR; ++i) { const int k = lane * 8 * EPR + i; x[i] = k < cnt ? e2(k) : 0.0f; }
                int dummy = 0;
                a = seq::warp_seq_block<8 * EPR>(a, x, s_wbuf[c], dummy);      // broadcast LDS.128 feed: the speed of the FADD chain
                return a;
            };
            const float acc = spec::walk_chain(ng, 0.0f, load_head, load_cand, seq_group, n_err, n_seq);
            n_groups += ng;
            if (lane == 0) s_sum[c] = acc;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const int nR = S.nright[it], nL = n - nR;
            const bool invalid = cand >= 0 && (nL < P.min_data || nR < P.min_data);
            const float lcf = (float)nL, rcf = (float)nR;
            float result = 0.0f;
            if (PASS == 0) {
                float lrec, rrec;
                if (cand >= 0) { lrec = nL > 0 ? 1.0f / lcf : 0.0f; rrec = nR > 0 ? 1.0f / rcf : 0.0f; }
                else { lrec = 1.0f / lcf; rrec = 0.0f; }   // parent: n_samples_recip = 1/n (split_candidate_generator.cpp:265,296)
                float ln = 0.0f, rn = 0.0f, mean[2 * D];
#pragma unroll
                for (int d = 0; d < D; ++d) { mean[d] = s_sum[d] * lrec; mean[D + d] = s_sum[D + d] * rrec; }
#pragma unroll
                for (int d = 0; d < D; ++d) { ln = ln + mean[d] * mean[d]; rn = rn + mean[D + d] * mean[D + d]; }   // squared_norm
                if (P.score_func == GBRL_B200_SCORE_L2) {
                    result = cand >= 0 ? (lcf * ln + rcf * rn) : (ln * lcf);
                    P.out[it] = invalid ? -INFINITY : result;
                } else {
#pragma unroll
                    for (int i = 0; i < 2 * D; ++i) Wd.fin[(size_t)it * 8 + i] = mean[i];
                    Wd.fin[(size_t)it * 8 + 4] = ln; Wd.fin[(size_t)it * 8 + 5] = rn;
                }
            } else {
                const float ln = Wd.fin[(size_t)it * 8 + 4], rn = Wd.fin[(size_t)it * 8 + 5];
                const float fnum = s_sum[0], tnum = s_sum[1];
                if (cand >= 0) {
                    const float num = tnum + fnum;
                    const float den = rn * rcf + ln * lcf;
                    result = (den == 0.0f) ? 0.0f : num / sqrtf(den);
                } else {
                    const float den = ln * lcf;
                    result = (n == 0 || den == 0.0f) ? 0.0f : fnum / sqrtf(den);
                }
                P.out[it] = invalid ? -INFINITY : result;
            }
        }
        __syncthreads();
    }
    if (lane == 0 && n_groups) {
        atomicAdd((unsigned long long *)&ctl_stats->stat_chain_fast, (unsigned long long)(n_groups - n_seq));
        atomicAdd((unsigned long long *)&ctl_stats->stat_chain_slow, (unsigned long long)n_seq);
        atomicAdd((unsigned long long *)&ctl_stats->stat_chain_seq, (unsigned long long)n_seq * 8);
        if (n_err) atomicAdd((unsigned long long *)&ctl_stats->stat_chain_err, (unsigned long long)n_err);
    }
}

template <int D>
static void launch_wide_d(Model &m, const ReplayParams &R, const StreamParams &S, const WideParams &Wd, cudaStream_t s) {
    Workspace &ws = m.ws;
    Ctl *ctl = ws.ctl.as<Ctl>();
    const int grid = ws.n_sms * 8;
    GB_LAUNCH((wide_bits_kernel<D>), grid, 256, 0, s, R, ws.na, S, Wd);
    GB_LAUNCH((wide_prefix_kernel<D>), ws.n_sms * 4, 256, 0, s, R, ws.na, S, Wd);
    GB_LAUNCH((wide_sim_kernel<D, 0>), ws.n_sms * 16, 128, 0, s, R, ws.na, S, Wd);
    GB_LAUNCH((wide_walk_kernel<D, 0>), ws.n_sms * 8, 32 * 2 * D, 0, s, R, ws.na, S, Wd, ctl);
    if (m.cfg.split_score_func != GBRL_B200_SCORE_L2) {
        GB_LAUNCH((wide_sim_kernel<D, 1>), ws.n_sms * 16, 128, 0, s, R, ws.na, S, Wd);
        GB_LAUNCH((wide_walk_kernel<D, 1>), ws.n_sms * 8, 32 * 2 * D, 0, s, R, ws.na, S, Wd, ctl);
    }
}

void launch_replay_wide(Model &m, const ReplayParams &R, const StreamParams &S, cudaStream_t s) {
    Workspace &ws = m.ws;
    WideParams Wd;
    const size_t ng = (size_t)ws.rwide_groups, D2 = (size_t)2 * ws.D;
    char *p = ws.rwide.as<char>();
    Wd.head = reinterpret_cast<spec::Head *>(p); p += ng * D2 * sizeof(spec::Head);          // 16-byte entries first (alignment)
    Wd.cand = reinterpret_cast<spec::Cand *>(p); p += ng * D2 * spec::J * sizeof(spec::Cand);
    Wd.bsum = reinterpret_cast<double *>(p); p += ng * D2 * sizeof(double);
    Wd.pred = reinterpret_cast<float *>(p); p += ng * D2 * sizeof(float);
    Wd.gitem = reinterpret_cast<int *>(p); p += ng * sizeof(int);
    Wd.fin = reinterpret_cast<float *>(p);
    Wd.cap_groups = (long long)ng;
    if (ws.D == 1) launch_wide_d<1>(m, R, S, Wd, s);
    else launch_wide_d<2>(m, R, S, Wd, s);
}

}  // namespace gb
