// split.cu -- split-score scan, arg-max, near-tie replay and split decision.
//
// Reference semantics restated (file:line are gbrl/src/cpp/):
//   L2 score        node.cpp:354-375     n_L*|S_L/n_L|^2 + n_R*|S_R/n_R|^2, -inf if a side < min_data_in_leaf
//   Cosine score    node.cpp:225-250 + math_ops.h:538-575  (S_R.mu_R + S_L.mu_L) / sqrt(|mu_R|^2 n_R + |mu_L|^2 n_L)
//   path guard      node.cpp:151-166     a (feature, value) already on the node's path scores -inf
//   greedy          fitter.cpp:292-371   gain = score*w[f] - parent, strict '>' (lowest index wins ties), split iff >= 0
//   parent score    split_candidate_generator.cpp:262-320, root forced to 0 (fitter.cpp:315)
//   oblivious       fitter.cpp:411-477   score_c = (sum over nodes in node order) * w[rev_map[f]], stop iff best == -inf
//
// Two-tier evaluation:
//   1. exact tier  -- per-side sums come from the integer histograms (suffix scan over codes), are rounded
//      once to fp32 and pushed through the reference's formula in the reference's operation order.
//      Identical partitions therefore get bit-identical scores and the lowest-index rule reproduces the
//      reference on every exact tie (duplicate thresholds, small nodes).
//   2. replay tier -- the reference's own sums are sequential fp32 accumulations in ascending sample
//      order, i.e. they carry O(u*sqrt(n)) rounding noise.  Whenever another (non-identical) candidate is
//      within that noise band of the exact best, both are re-scored with exactly the reference's chain of
//      float operations, so the chosen split is the reference's, bit for bit.  The chains are evaluated by the
//      bit-exact parallel evaluator of chain.cuh:
//        output_dim <= 2   replay_plan / replay_gather here, then replay_wide.cu (summaries by the whole GPU,
//                          one walking warp per chain)
//        output_dim 3..4   replay_stream_kernel (one CTA per item over the gathered streams)
//        output_dim  > 4   replay_kernel (one lane per output dimension runs its chain sequentially)
//        items whose side-bit plane does not fit the stream buffer: replay_par_kernel (gathers its rows itself)
#include "engine.cuh"
#include "chain.cuh"
#include "replay.cuh"
#include "plan.cuh"
#include <cfloat>
#include <cstdlib>
#include <climits>

namespace gb {

constexpr float U24 = 5.9604644775390625e-08f;   // 2^-24

struct ScanParams {
    int level, nT, F, B, D, C;                // C = F*B candidates
    int score_func, oblivious, min_data, max_depth, use_subtraction;
    const long long *hist_par;                // parent level buffer
    long long *hist_cur;                      // this level's buffer
    const float *thr;                         // [F*B]
    const float *fw;                          // feature weights [input_dim]
    float *scores;                            // [slot][C]
    uint8_t *cand_flags;                      // greedy: [slot][C] bit0 = dominated; oblivious: [C] bit1 = empty code below
    float2 *tile_best;                        // [slot][F] per-(node, feature) best (gain, idx as float bits)
    const Ctl *ctl;
};

// reference formula on per-side sums (already rounded to fp32), templated on the max dimension
template <int DM>
__device__ __forceinline__ float side_score(int func, int D, int nL, int nR, const float *SL, const float *SR, int min_data) {
    if (nL < min_data || nR < min_data) return -INFINITY;
    const float lcf = (float)nL, rcf = (float)nR;
    const float lrec = nL > 0 ? 1.0f / lcf : 0.0f;
    const float rrec = nR > 0 ? 1.0f / rcf : 0.0f;
    float ln = 0.0f, rn = 0.0f, tnum = 0.0f, fnum = 0.0f;
#pragma unroll
    for (int d = 0; d < DM; ++d) {
        if (d < D) {
            const float lm = SL[d] * lrec, rm = SR[d] * rrec;
            ln = ln + lm * lm;
            rn = rn + rm * rm;
            // sum_i g_i . mu  ==  S . mu  (the exact tier's algebraic form of mat_vec_dot_sum)
            tnum = tnum + SR[d] * rm;
            fnum = fnum + SL[d] * lm;
        }
    }
    if (func == GBRL_B200_SCORE_L2) return lcf * ln + rcf * rn;
    if (nR == 0) tnum = 0.0f;
    if (nL == 0) fnum = 0.0f;
    const float num = tnum + fnum;
    const float den = rn * rcf + ln * lcf;
    if (den == 0.0f) return 0.0f;
    return num / sqrtf(den);
}

template <int DM>
__device__ __forceinline__ float node_parent_score(int func, int D, int n, const long long *tot, double inv_scale) {
    const float nf = (float)n;
    const float rec = 1.0f / nf;
    float norm = 0.0f, dot = 0.0f;
#pragma unroll
    for (int d = 0; d < DM; ++d) {
        if (d < D) {
            const float S = (float)((double)tot[d] * inv_scale);
            const float mu = S * rec;
            norm = norm + mu * mu;
            dot = dot + S * mu;
        }
    }
    if (func == GBRL_B200_SCORE_L2) return norm * nf;
    if (n == 0) return 0.0f;
    const float den = norm * nf;
    if (den == 0.0f) return 0.0f;
    return dot / sqrtf(den);
}

__device__ __forceinline__ bool better(float ga, int ia, float gb_, int ib) {
    return (ga > gb_) || (ga == gb_ && ia < ib);
}

// One CTA per (node slot, feature): thread t owns code-bin b = 255 - t, so an inclusive prefix scan over t is the
// suffix sum over bins that turns the histogram into per-candidate right-side sums: candidate j = b of the feature
// has right = sum of bins >= b.  All 256 candidates of a feature are scored in parallel; (1+D) int64 scans by
// warp shuffles + one shared-memory hop.
__device__ __forceinline__ long long shfl_up_ll(long long v, int d) {
    int lo = __shfl_up_sync(0xffffffffu, (int)(v & 0xffffffffll), d);
    int hi = __shfl_up_sync(0xffffffffu, (int)(v >> 32), d);
    return ((long long)hi << 32) | (unsigned int)lo;
}

template <int DM>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_kernel(ScanParams P, NodeArrays na) {
    __shared__ long long s_wtot[8][1 + DM];
    __shared__ int s_cnt[NB];
    __shared__ float s_gain[8];
    __shared__ int s_idx[8];
    __shared__ int s_path_f[MAX_DEPTH_SUPPORTED];
    __shared__ float s_path_v[MAX_DEPTH_SUPPORTED];
    __shared__ float s_parent;
    const int p = blockIdx.x / P.F, f = blockIdx.x % P.F;
    const int h = level_base(P.level) + p;
    if (na.state[h] != NODE_OPEN) return;
    const int tile = f / FT, fl = f % FT;
    const int D = P.D, HS = 1 + D;
    const int n = na.seg_len[h];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int b = NB - 1 - t;
    const bool derived = (P.level > 0) && P.use_subtraction && (na.direct[h] == 0);
    const double inv_scale = exp2((double)(-P.ctl->qexp));
    const size_t tile_words = (size_t)NB * FT * HS;
    long long *Hc = P.hist_cur + ((size_t)p * P.nT + tile) * tile_words;
    const long long *Hs = nullptr, *Hp = nullptr;
    if (derived) {
        const int sib = (h & 1) ? h + 1 : h - 1;
        const int par = (h - 1) >> 1;
        Hs = P.hist_cur + ((size_t)(sib - level_base(P.level)) * P.nT + tile) * tile_words;
        Hp = P.hist_par + ((size_t)(par - level_base(P.level - 1)) * P.nT + tile) * tile_words;
    }
    // ancestors' conditions (path guard) and the parent score, once per CTA
    if (t == 0) {
        int a = h, k = 0;
        while (a > 0) {
            const int par = (a - 1) >> 1;
            s_path_f[k] = na.split_f[par];
            s_path_v[k] = na.split_thr[par];
            ++k;
            a = par;
        }
        for (; k < MAX_DEPTH_SUPPORTED; ++k) s_path_f[k] = -1;
        float ps = 0.0f;
        if (!P.oblivious && P.level > 0) ps = node_parent_score<DM>(P.score_func, D, n, na.tot_sum + (size_t)h * D, inv_scale);
        s_parent = ps;
        if (f == 0) na.parent_score[h] = ps;
    }
    // load this bin (derived nodes: parent - sibling, written back for the next level)
    const size_t o = ((size_t)b * FT + fl) * HS;
    long long v[1 + DM];
#pragma unroll
    for (int d = 0; d <= DM; ++d) {
        v[d] = 0;
        if (d <= D) {
            v[d] = derived ? (Hp[o + d] - Hs[o + d]) : Hc[o + d];
            if (derived) Hc[o + d] = v[d];
        }
    }
    s_cnt[b] = (int)v[0];
    // inclusive scan over t (== suffix sum over bins)
#pragma unroll
    for (int d = 0; d <= DM; ++d) {
        if (d <= D) {
            long long x = v[d];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const long long y = shfl_up_ll(x, off);
                if (lane >= off) x += y;
            }
            v[d] = x;
            if (lane == 31) s_wtot[warp][d] = x;
        }
    }
    __syncthreads();
#pragma unroll
    for (int d = 0; d <= DM; ++d) {
        if (d <= D) {
            long long add = 0;
            for (int w = 0; w < warp; ++w) add += s_wtot[w][d];
            v[d] += add;
        }
    }
    const float parent = s_parent;
    const bool poisoned = P.ctl->bg_nonfinite != 0;
    const float w = P.fw[f];
    const int depth = P.level;
    float gain = -INFINITY;
    int idx = INT_MAX;
    if (b < P.B) {
        const int nR = (int)v[0];
        const int nL = n - nR;
        float SL[DM], SR[DM];
#pragma unroll
        for (int d = 0; d < DM; ++d) {
            if (d < D) {
                const long long r = v[1 + d];
                SR[d] = (float)((double)r * inv_scale);
                SL[d] = (float)((double)(na.tot_sum[(size_t)h * D + d] - r) * inv_scale);
            } else { SR[d] = 0.0f; SL[d] = 0.0f; }
        }
        const float tv = P.thr[(size_t)f * P.B + b];
        bool reused = false;
        for (int k = 0; k < depth; ++k) reused |= (s_path_f[k] == f && s_path_v[k] == tv);
        float sc = reused ? -INFINITY : side_score<DM>(P.score_func, D, nL, nR, SL, SR, P.min_data);
        // a NaN build_grad poisons every candidate's sequential sum in the reference (e.g. n_samples == 1 with L2:
        // std = sqrt(0 * 1/0)); NaN scores never win (fitter.cpp:338) -> the node stays a leaf
        if (poisoned && !reused) sc = NAN;
        idx = f * P.B + b;
        gain = P.oblivious ? sc : sc * w - parent;
        P.scores[(size_t)p * P.C + idx] = gain;
        // dominated: same partition as the previous threshold of this feature (no sample has code == b), which is
        // itself not excluded by the path guard -> equal score, higher index: can never win
        uint8_t fl8 = 0;
        if (b > 0 && s_cnt[b - 1] == 0) {
            const float tvp = P.thr[(size_t)f * P.B + b - 1];
            bool rp = false;
            for (int k = 0; k < depth; ++k) rp |= (s_path_f[k] == f && s_path_v[k] == tvp);
            if (!rp) fl8 = 1;
            if (P.oblivious && P.level == 0) fl8 |= 2;   // no sample of the whole set has code == b
        }
        if (!P.oblivious) P.cand_flags[(size_t)p * P.C + idx] = fl8;
        else if (P.level == 0) P.cand_flags[idx] = fl8 & 2;
        if (!(gain > -INFINITY)) { gain = -INFINITY; idx = INT_MAX; }
    }
    // block arg-max (lowest index on ties)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float og = __shfl_down_sync(0xffffffffu, gain, off);
        const int oi = __shfl_down_sync(0xffffffffu, idx, off);
        if (better(og, oi, gain, idx)) { gain = og; idx = oi; }
    }
    if (lane == 0) { s_gain[warp] = gain; s_idx[warp] = idx; }
    __syncthreads();
    if (t == 0) {
        for (int w2 = 1; w2 < 8; ++w2)
            if (better(s_gain[w2], s_idx[w2], gain, idx)) { gain = s_gain[w2]; idx = s_idx[w2]; }
        P.tile_best[(size_t)p * P.F + f] = make_float2(gain, __int_as_float(idx));
    }
}

// ---------------------------------------------------------------- greedy: per-node best + replay list
struct SelectParams {
    int level, nT, F, B, D, C, tie_replay, replay_cap;
    float kappa;
    const float *scores;
    const uint8_t *cand_flags;
    const float2 *tile_best;
    ReplayItem *replay;
    float *replay_exact;      // exact-tier score of every replay item (0 for parent items): calibration statistic of the band
    int *state_snap;          // speculative trees: node states before the level's decision (nullptr otherwise)
    Ctl *ctl;
};

__global__ void __launch_bounds__(1024) select_greedy_kernel(SelectParams P, NodeArrays na) {
    __shared__ float s_best;
    __shared__ int s_besti, s_count, s_begin, s_w;
    const int p = blockIdx.x, h = level_base(P.level) + p;
    if (P.state_snap != nullptr && threadIdx.x == 0) P.state_snap[h] = na.state[h];
    if (na.state[h] != NODE_OPEN) return;
    {
        // arg-max over the per-feature bests (lowest candidate index on ties)
        __shared__ float r_g[32];
        __shared__ int r_i[32];
        float g = -INFINITY; int bi = INT_MAX;
        for (int t = threadIdx.x; t < P.F; t += blockDim.x) {
            const float2 v = P.tile_best[(size_t)p * P.F + t];
            const int i = __float_as_int(v.y);
            if (v.x > -INFINITY && better(v.x, i, g, bi)) { g = v.x; bi = i; }
        }
        for (int off = 16; off > 0; off >>= 1) {
            const float og = __shfl_down_sync(0xffffffffu, g, off);
            const int oi = __shfl_down_sync(0xffffffffu, bi, off);
            if (better(og, oi, g, bi)) { g = og; bi = oi; }
        }
        if ((threadIdx.x & 31) == 0) { r_g[threadIdx.x >> 5] = g; r_i[threadIdx.x >> 5] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w2 = 1; w2 < (int)(blockDim.x >> 5); ++w2)
                if (better(r_g[w2], r_i[w2], g, bi)) { g = r_g[w2]; bi = r_i[w2]; }
            s_best = g; s_besti = bi; s_count = 0; s_w = 0;
            na.best_gain[h] = g;
            na.best_idx[h] = (bi == INT_MAX) ? -1 : bi;
            na.rep_begin[h] = 0; na.rep_count[h] = 0;
            atomicAdd((unsigned long long *)&P.ctl->stat_nodes_evaluated, 1ull);
        }
    }
    __syncthreads();
    const float g = s_best;
    if (!P.tie_replay || !(g > -INFINITY)) return;
    const int n = na.seg_len[h];
    const float parent = na.parent_score[h];
    // candidates are compared on score*w (the parent score is common to all of them), so the comparison band
    // only carries the rounding noise of the candidate sums; the parent's noise matters for the sign test gain >= 0
    const float unit = P.kappa * U24 * sqrtf((float)n);
    const float band_c = unit * fabsf(g + parent) + FLT_MIN;
    const float band = unit * (fabsf(g + parent) + fabsf(parent)) + FLT_MIN;
    const float lim = g - band_c;
    const float *sc = P.scores + (size_t)p * P.C;
    const uint8_t *fl = P.cand_flags + (size_t)p * P.C;
    int mine = 0;
    for (int i = threadIdx.x; i < P.C; i += blockDim.x) mine += (sc[i] >= lim && !(fl[i] & 1)) ? 1 : 0;
    if (mine) atomicAdd(&s_count, mine);
    __syncthreads();
    const bool need = (s_count > 1) || (fabsf(g) <= band && P.level > 0);
    if (!need) return;
    const int n_items = s_count + (P.level > 0 ? 1 : 0);
    if (threadIdx.x == 0) {
        const int beg = atomicAdd(&P.ctl->n_replay, n_items);
        if (beg + n_items > P.replay_cap) {
            atomicAdd(&P.ctl->replay_overflow, 1);
            s_begin = -1;
            // the range was handed out all the same: make its in-bounds part harmless (parent items of this node)
            for (int i = beg; i < P.replay_cap && i < beg + n_items; ++i) { P.replay[i].node = h; P.replay[i].cand = -1; }
        } else {
            s_begin = beg;
            na.rep_begin[h] = beg; na.rep_count[h] = n_items; na.band[h] = band;
            if (P.level > 0) { P.replay[beg].node = h; P.replay[beg].cand = -1; P.replay_exact[beg] = 0.0f; s_w = 1; }
            atomicAdd((unsigned long long *)&P.ctl->stat_replay_nodes, 1ull);
            atomicAdd((unsigned long long *)&P.ctl->stat_replay_items, (unsigned long long)n_items);
        }
    }
    __syncthreads();
    if (s_begin < 0) return;
    for (int i = threadIdx.x; i < P.C; i += blockDim.x) {
        if (sc[i] >= lim && !(fl[i] & 1)) {
            const int w = atomicAdd(&s_w, 1);
            P.replay[s_begin + w].node = h;
            P.replay[s_begin + w].cand = i;
            P.replay_exact[s_begin + w] = sc[i];
        }
    }
}

// ---------------------------------------------------------------- oblivious: level-wide reduction
struct OblParams {
    int level, F, B, D, C, nn, tie_replay, replay_cap, nblocks;
    float kappa;
    int N;
    const float *scores;       // [nn][C] raw node scores
    const uint8_t *cand_flags; // [C] bit1
    const float *fw;
    const int *rev_map;
    float *obl_tot;            // [C]
    float *obl_nb;             // [C] rounding-noise scale of the total: w * sqrt(sum_nodes n_node * score^2)
    float2 *blk_best;          // [nblocks]
    ReplayItem *replay;
    int *obl_cands;            // candidate list of replay (stored after the items, see launcher)
    int *state_snap;           // speculative trees: node states before the level's decision (nullptr otherwise)
    Ctl *ctl;
};

__global__ void __launch_bounds__(256) obl_reduce_kernel(OblParams P, NodeArrays na) {
    __shared__ float s_gain[256];
    __shared__ int s_idx[256];
    const int c = blockIdx.x * 256 + threadIdx.x;
    float tot = -INFINITY;
    int idx = INT_MAX;
    if (c < P.C) {
        float s = 0.0f, nb = 0.0f;
        const int base = level_base(P.level);
        for (int p = 0; p < P.nn; ++p) {
            const float sp = P.scores[(size_t)p * P.C + c];
            s += sp;                                                         // fitter.cpp:427-430, node order
            // rounding noise of the node's sequential sums: independent from node to node (disjoint samples), so the
            // noise scale of the total is the root of the sum of squares of the per-node scales 2^-24 sqrt(n) |score|
            if (sp > -INFINITY) nb += (float)na.seg_len[base + p] * sp * sp;
        }
        const int f = c / P.B;
        const float wf = P.fw[P.rev_map[f]];
        s = s * wf;                                                          // fitter.cpp:432-435
        P.obl_tot[c] = s;
        P.obl_nb[c] = sqrtf(nb) * fabsf(wf);
        if (s > -INFINITY) { tot = s; idx = c; }
    }
    s_gain[threadIdx.x] = tot; s_idx[threadIdx.x] = idx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o && better(s_gain[threadIdx.x + o], s_idx[threadIdx.x + o], s_gain[threadIdx.x], s_idx[threadIdx.x])) {
            s_gain[threadIdx.x] = s_gain[threadIdx.x + o]; s_idx[threadIdx.x] = s_idx[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) P.blk_best[blockIdx.x] = make_float2(s_gain[0], __int_as_float(s_idx[0]));
}

__global__ void __launch_bounds__(1024) obl_select_kernel(OblParams P, NodeArrays na) {
    __shared__ float s_best, s_band;
    __shared__ int s_besti, s_count, s_w, s_ok;
    __shared__ float r_g[32];
    __shared__ int r_i[32];
    const int base = level_base(P.level);
    if (P.state_snap != nullptr)
        for (int p = threadIdx.x; p < P.nn; p += blockDim.x) P.state_snap[base + p] = na.state[base + p];
    if (na.state[base] != NODE_OPEN) return;    // tree already stopped
    {
        // arg-max over the per-block bests (lowest candidate index on ties)
        float g = -INFINITY; int bi = INT_MAX;
        for (int b = threadIdx.x; b < P.nblocks; b += blockDim.x) {
            const float2 v = P.blk_best[b];
            const int i = __float_as_int(v.y);
            if (v.x > -INFINITY && better(v.x, i, g, bi)) { g = v.x; bi = i; }
        }
        for (int off = 16; off > 0; off >>= 1) {
            const float og = __shfl_down_sync(0xffffffffu, g, off);
            const int oi = __shfl_down_sync(0xffffffffu, bi, off);
            if (better(og, oi, g, bi)) { g = og; bi = oi; }
        }
        if ((threadIdx.x & 31) == 0) { r_g[threadIdx.x >> 5] = g; r_i[threadIdx.x >> 5] = bi; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float g = r_g[0]; int bi = r_i[0];
        for (int w2 = 1; w2 < (int)(blockDim.x >> 5); ++w2)
            if (better(r_g[w2], r_i[w2], g, bi)) { g = r_g[w2]; bi = r_i[w2]; }
        s_best = g; s_besti = (bi == INT_MAX) ? -1 : bi; s_count = 0; s_w = 0; s_ok = 0;
        P.ctl->obl_best = g; P.ctl->obl_best_idx = s_besti; P.ctl->obl_has_replay = 0;
        atomicAdd((unsigned long long *)&P.ctl->stat_nodes_evaluated, (unsigned long long)P.nn);
        // per-node sequential sums carry noise ~ 2^-24 sqrt(n_node) |score|; the total's noise scale is their quadrature sum
        s_band = 0.5f * P.kappa * U24 * (s_besti >= 0 ? P.obl_nb[s_besti] : 0.0f) + FLT_MIN;
    }
    __syncthreads();
    const float g = s_best;
    if (!P.tie_replay || s_besti < 0) return;
    const float lim = g - s_band;
    const float half = 0.5f * P.kappa * U24;
    auto in_band = [&](int i) -> bool {
        if (!(P.obl_tot[i] + half * P.obl_nb[i] >= lim)) return false;
        const int j = i % P.B;
        if (j > 0 && (P.cand_flags[i] & 2) && P.obl_tot[i - 1] > -INFINITY) return false;   // dominated by j-1
        return true;
    };
    int mine = 0;
    for (int i = threadIdx.x; i < P.C; i += blockDim.x) mine += in_band(i) ? 1 : 0;
    if (mine) atomicAdd(&s_count, mine);
    __syncthreads();
    if (s_count <= 1) return;
    if (threadIdx.x == 0) {
        const long long n_items = (long long)s_count * P.nn;
        if (n_items > P.replay_cap) { atomicAdd(&P.ctl->replay_overflow, 1); s_ok = 0; }
        else {
            s_ok = 1;
            P.ctl->n_replay = (int)n_items;
            P.ctl->obl_has_replay = s_count;
            atomicAdd((unsigned long long *)&P.ctl->stat_replay_nodes, (unsigned long long)P.nn);
            atomicAdd((unsigned long long *)&P.ctl->stat_replay_items, (unsigned long long)n_items);
        }
    }
    __syncthreads();
    if (!s_ok) return;
    for (int i = threadIdx.x; i < P.C; i += blockDim.x) {
        if (in_band(i)) {
            const int w = atomicAdd(&s_w, 1);
            P.obl_cands[w] = i;
            for (int p = 0; p < P.nn; ++p) {
                P.replay[(size_t)w * P.nn + p].node = base + p;
                P.replay[(size_t)w * P.nn + p].cand = i;
            }
        }
    }
}

// ---------------------------------------------------------------- replay: the reference's own arithmetic

// One 128-thread CTA per item.  Rows of the node are visited in ascending sample order (== reference
// sample_indices).  The chain itself is inherently sequential (every float add depends on the previous one), so
// the kernel is organised around keeping that one chain fed: all four warps gather the next stage of rows
// (order -> feature value -> gradients; the loads are issued before the chain of the current stage starts and
// land in registers while it runs), warp 0 runs the chain out of shared memory, lane d owning output dimension d.
// The Cosine numerators are one chain over (row, col) (mat_vec_dot_sum) and are run by every lane redundantly.
// CTA size: 512 threads (16 warps gather a stage of 4096 rows for D == 1 while warp 0 runs the chain) for narrow
// outputs, 128 threads for wide ones (the stage has to fit in shared memory twice)

template <int R, int PASS, int RP_THREADS>
__device__ __forceinline__ void replay_pass(const ReplayParams &P, int s0, int n, int f, float tv, bool is_cand, float *sg /*[2][STAGE*D]*/,
                                            unsigned int *smask /*[2][STAGE/32]*/, const float *smean, float *accL, float *accR,
                                            int *s_nright, float &tnum, float &fnum) {
    constexpr int STAGE = RP_THREADS * R;
    constexpr int DC = (R == 8) ? 1 : (R == 4) ? 2 : 0;     // compile-time output_dim on the two fast paths
    const int D = DC > 0 ? DC : P.D, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_stages = (n + STAGE - 1) / STAGE;
    constexpr int DG = R >= 8 ? 1 : R >= 4 ? 2 : R >= 2 ? 4 : 0;   // gradient values prefetched per row (== D on the fast paths)
    int rows[R];
    float xv[R];
    float gpre[R][DG > 0 ? DG : 1];
    const bool pre = (DG > 0) && (D <= DG);
    // Loads are split in two dependent steps that are issued one stage apart, so that no warp ever waits for a
    // row id before it can issue the gathers that depend on it:
    //   load_rows(st+2) -> registers      (coalesced read of `order`)
    //   gather(st+1)    -> registers      (feature value + gradients of rows whose ids arrived a stage ago)
    //   chain(st)       from shared memory
    //   commit(st+1)    registers -> shared memory
    int rows_n[R];
    auto load_rows = [&](int st, int *dst) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int k = st * STAGE + r * RP_THREADS + tid;
            dst[r] = (st < n_stages && k < n) ? P.order[s0 + k] : -1;
        }
    };
    auto gather = [&](const int *src) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            rows[r] = src[r];
            xv[r] = (src[r] >= 0 && is_cand) ? P.X[(size_t)src[r] * P.F + f] : -INFINITY;
            if (pre && src[r] >= 0) {
#pragma unroll
                for (int d = 0; d < (DG > 0 ? DG : 1); ++d)
                    if (d < D) gpre[r][d] = P.bg[(size_t)src[r] * D + d];
            }
        }
    };
    auto commit = [&](int buf) {
        float *g = sg + (size_t)buf * STAGE * D;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int slot = r * RP_THREADS + tid;
            const bool right = rows[r] >= 0 && (xv[r] > tv);                     // node.cpp:339
            const unsigned int m = __ballot_sync(0xffffffffu, right);
            if (lane == 0) { smask[buf * (STAGE / 32) + slot / 32] = m; if (PASS == 0 && m) atomicAdd(s_nright, __popc(m)); }
            if (rows[r] >= 0) {
                if (pre) {
#pragma unroll
                    for (int d = 0; d < (DG > 0 ? DG : 1); ++d)
                        if (d < D) g[(size_t)slot * D + d] = gpre[r][d];
                } else {
                    for (int d = 0; d < D; ++d) g[(size_t)slot * D + d] = P.bg[(size_t)rows[r] * D + d];
                }
            }
        }
    };
    {
        int r0[R];
        load_rows(0, r0);
        load_rows(1, rows_n);
        gather(r0);
        commit(0);
    }
    __syncthreads();
    for (int st = 0; st < n_stages; ++st) {
        const int buf = st & 1;
        int rows_nn[R];
        if (st + 1 < n_stages) gather(rows_n);
        load_rows(st + 2, rows_nn);
        if (warp == 0) {
            const float *g = sg + (size_t)buf * STAGE * D;
            const int cnt = min(STAGE, n - st * STAGE);
            int w_begin = 0;
            if (DC == 1 && PASS == 0) {
                // D == 1: the whole stage as one software-pipelined chain.  The 32 values and the side mask of word
                // w+1 are loaded (LDS.128 x 8) while the 32 predicated adds of word w run; the only dependency left on
                // the critical path is the 4-cycle FADD chain per side (node.cpp:341-350).
                const int full = cnt / 32;
                if (full > 0) {
                    const unsigned int gaddr = (unsigned int)__cvta_generic_to_shared(g);
                    const unsigned int *mk = smask + buf * (STAGE / 32);
                    float4 cur[8], nxt[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(cur[j].x), "=f"(cur[j].y), "=f"(cur[j].z), "=f"(cur[j].w) : "r"(gaddr + 16u * j));
                    unsigned int mcur = mk[0], mnxt = 0;
                    float aL = accL[0], aR = accR[0];
                    for (int w = 0; w < full; ++w) {
                        if (w + 1 < full) {
                            const unsigned int a2 = gaddr + 128u * (w + 1);
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(nxt[j].x), "=f"(nxt[j].y), "=f"(nxt[j].z), "=f"(nxt[j].w) : "r"(a2 + 16u * j));
                            mnxt = mk[w + 1];
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const unsigned int m4 = mcur >> (4 * j);
                            if (m4 & 1u) aR = aR + cur[j].x; else aL = aL + cur[j].x;
                            if (m4 & 2u) aR = aR + cur[j].y; else aL = aL + cur[j].y;
                            if (m4 & 4u) aR = aR + cur[j].z; else aL = aL + cur[j].z;
                            if (m4 & 8u) aR = aR + cur[j].w; else aL = aL + cur[j].w;
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
                        mcur = mnxt;
                    }
                    accL[0] = aL; accR[0] = aR;
                    w_begin = full;
                }
            }
            for (int w = w_begin; w * 32 < cnt; ++w) {
                const unsigned int mask = smask[buf * (STAGE / 32) + w];
                const int c32 = min(32, cnt - w * 32);
                if (PASS == 0) {
                    if (DC == 1 && c32 == 32) {
                        // D == 1, full word: 8 x LDS.128 with immediate offsets, then 32 predicated adds; the only
                        // true dependency is the 4-cycle FADD chain of the side a row falls on   (node.cpp:341-350)
                        const float4 *g4 = reinterpret_cast<const float4 *>(g + (size_t)w * 32);
                        float4 q4[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) q4[j] = g4[j];
                        float aL = accL[0], aR = accR[0];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const unsigned int m4 = mask >> (4 * j);
                            if (m4 & 1u) aR = aR + q4[j].x; else aL = aL + q4[j].x;
                            if (m4 & 2u) aR = aR + q4[j].y; else aL = aL + q4[j].y;
                            if (m4 & 4u) aR = aR + q4[j].z; else aL = aL + q4[j].z;
                            if (m4 & 8u) aR = aR + q4[j].w; else aL = aL + q4[j].w;
                        }
                        accL[0] = aL; accR[0] = aR;     // every lane runs the same chain; lane 0 is the owner
                    } else if (DC == 2 && c32 == 32) {
                        // D == 2: lanes 0 / 1 own the two columns (stride-2 floats, immediate offsets)
                        const float *gc = g + (size_t)w * 64 + (lane & 1);
                        float aL = accL[0], aR = accR[0];
#pragma unroll
                        for (int b8 = 0; b8 < 4; ++b8) {
                            float v[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = gc[(8 * b8 + j) * 2];
                            const unsigned int m8 = mask >> (8 * b8);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                if (m8 & (1u << j)) aR = aR + v[j]; else aL = aL + v[j];
                            }
                        }
                        if (lane < 2) { accL[0] = aL; accR[0] = aR; }
                    } else if (D <= 32) {
                        // generic D <= 32: batch 8 shared loads, then 8 predicated adds
                        const bool mine = lane < D;
                        const float *gcol = g + (size_t)(w * 32) * D + (mine ? lane : 0);
                        float aL = accL[0], aR = accR[0];
                        int t = 0;
                        for (; t + 8 <= c32; t += 8) {
                            float v[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = gcol[(size_t)(t + j) * D];
                            const unsigned int m8 = mask >> t;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                if (m8 & (1u << j)) aR = aR + v[j]; else aL = aL + v[j];
                            }
                        }
                        for (; t < c32; ++t) {
                            const float v = gcol[(size_t)t * D];
                            if ((mask >> t) & 1u) aR = aR + v; else aL = aL + v;
                        }
                        if (mine) { accL[0] = aL; accR[0] = aR; }
                    } else {
                        for (int t = 0; t < c32; ++t) {
                            const bool r = (mask >> t) & 1u;
#pragma unroll
                            for (int q = 0; q < 2; ++q) {
                                const int d = lane + 32 * q;
                                if (d < D) {
                                    const float v = g[(size_t)(w * 32 + t) * D + d];
                                    if (r) accR[q] = accR[q] + v; else accL[q] = accL[q] + v;
                                }
                            }
                        }
                    }
                } else {
                    // mat_vec_dot_sum chains (math_ops.h:432-449): rows in order, columns inner
                    for (int t = 0; t < c32; ++t) {
                        const bool r = (mask >> t) & 1u;
                        const float *gr = g + (size_t)(w * 32 + t) * D;
                        if (r) { for (int d = 0; d < D; ++d) tnum = tnum + gr[d] * smean[D + d]; }
                        else   { for (int d = 0; d < D; ++d) fnum = fnum + gr[d] * smean[d]; }
                    }
                }
            }
        }
        if (st + 1 < n_stages) commit(buf ^ 1);
#pragma unroll
        for (int r = 0; r < R; ++r) rows_n[r] = rows_nn[r];
        __syncthreads();
    }
}

template <int R, int RP_THREADS>
__global__ void __launch_bounds__(RP_THREADS) replay_kernel(ReplayParams P, NodeArrays na) {
    extern __shared__ float s_dyn[];
    constexpr int STAGE = RP_THREADS * R;
    const int D = P.D, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *sg = s_dyn;                                           // [2][STAGE][D]
    float *smean = sg + (size_t)2 * STAGE * D;                   // [2][D] left / right means
    unsigned int *smask = reinterpret_cast<unsigned int *>(smean + 2 * D);   // [2][STAGE/32]
    const int n_items = P.ctl->n_replay;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const ReplayItem item = P.items[it];
        const int h = item.node, cand = item.cand;
        const int s0 = na.seg_start[h], n = na.seg_len[h];
        const int f = cand >= 0 ? cand / P.B : 0;
        const float tv = cand >= 0 ? P.thr[cand] : INFINITY;
        float accL[2] = {0.0f, 0.0f}, accR[2] = {0.0f, 0.0f};    // lane d and lane d+32 (D <= 64)
        __shared__ int s_nright;
        if (threadIdx.x == 0) s_nright = 0;
        __syncthreads();
        float tnum = 0.0f, fnum = 0.0f;
        replay_pass<R, 0, RP_THREADS>(P, s0, n, f, tv, cand >= 0, sg, smask, smean, accL, accR, &s_nright, tnum, fnum);
        const int nR = s_nright;   // all ballots are committed before the last barrier of the pass
        const int nL = n - nR;
        const bool invalid = cand >= 0 && (nL < P.min_data || nR < P.min_data);
        const float lcf = (float)nL, rcf = (float)nR;
        float ln = 0.0f, rn = 0.0f;
        if (warp == 0) {
            float lrec, rrec;
            if (cand >= 0) { lrec = nL > 0 ? 1.0f / lcf : 0.0f; rrec = nR > 0 ? 1.0f / rcf : 0.0f; }
            else { lrec = 1.0f / lcf; rrec = 0.0f; }   // parent: n_samples_recip = 1/n (split_candidate_generator.cpp:265,296)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int d = lane + 32 * q;
                if (d < D) { smean[d] = accL[q] * lrec; smean[D + d] = accR[q] * rrec; }
            }
            __syncwarp();
            for (int d = 0; d < D; ++d) { ln = ln + smean[d] * smean[d]; rn = rn + smean[D + d] * smean[D + d]; }  // squared_norm
        }
        __syncthreads();
        float result = 0.0f;
        if (P.score_func == GBRL_B200_SCORE_L2) {
            result = cand >= 0 ? (lcf * ln + rcf * rn) : (ln * lcf);
        } else {
            replay_pass<R, 1, RP_THREADS>(P, s0, n, f, tv, cand >= 0, sg, smask, smean, accL, accR, &s_nright, tnum, fnum);
            if (cand >= 0) {
                const float num = tnum + fnum;
                const float den = rn * rcf + ln * lcf;
                result = (den == 0.0f) ? 0.0f : num / sqrtf(den);
            } else {
                const float den = ln * lcf;
                result = (n == 0 || den == 0.0f) ? 0.0f : fnum / sqrtf(den);
            }
        }
        if (invalid) result = -INFINITY;
        if (threadIdx.x == 0) P.out[it] = result;
        __syncthreads();
    }
}


// ---------------------------------------------------------------- replay, parallel chains (output_dim <= 4)
// Same result as replay_kernel, bit for bit, but the float chains are evaluated by the parallel chain evaluator of
// chain.cuh instead of one dependent FADD per row.  CTA = 512 threads per item, stages of 512*R rows (R = 8 / 4 / 2
// rows per lane for D = 1 / 2 / 3-4).  Loads are software-pipelined exactly as in replay_kernel (row ids two stages
// ahead, gathers one stage ahead).  Per stage:
//   phase A  every warp summarises its own sub-block (32*R rows) for every chain, in the chain's current binade;
//   phase B  warp c walks the 16 summaries of chain c in order: valid for the actual running sum -> one integer add,
//            otherwise the sub-block is run as the plain sequential float chain (node.cpp:341-350 order).
// Chains: PASS 0 -> (side, output dim): per-side sums of build_grads; PASS 1 -> side: the mat_vec_dot_sum chain over
// (row, col) of g * mean (math_ops.h:432-449).
template <int DCT> struct ParCfg {
    static constexpr int R = DCT == 1 ? 8 : DCT == 2 ? 4 : 2;
    static constexpr int T = 512, NW = 16, STAGE = T * R, SUB = 32 * R, EPL = R * DCT;
};

template <int DCT, int PASS>
__device__ __forceinline__ void replay_pass_par(const ReplayParams &P, int s0, int n, int f, float tv, bool is_cand, float *sg,
                                                unsigned int *smask, const float *smean, int *s_nright, seq::StageShared &sh,
                                                int &n_fast, int &n_slow, int &n_seq) {
    using C = ParCfg<DCT>;
    constexpr int R = C::R, T = C::T, NW = C::NW, STAGE = C::STAGE, SUB = C::SUB, EPL = C::EPL, D = DCT;
    constexpr int NCH = PASS == 0 ? 2 * D : 2;
    constexpr int KE = PASS == 0 ? R : 8;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_stages = (n + STAGE - 1) / STAGE;
    int rows[R], rows_n[R];
    float xv[R], gpre[R][D];
    auto load_rows = [&](int st, int *dst) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int k = st * STAGE + r * T + tid;
            dst[r] = (st < n_stages && k < n) ? P.order[s0 + k] : -1;
        }
    };
    auto gather = [&](const int *src) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            rows[r] = src[r];
            xv[r] = (src[r] >= 0 && is_cand) ? P.X[(size_t)src[r] * P.F + f] : -INFINITY;
            if (src[r] >= 0) {
#pragma unroll
                for (int d = 0; d < D; ++d) gpre[r][d] = P.bg[(size_t)src[r] * D + d];
            }
        }
    };
    auto commit = [&](int buf) {
        float *g = sg + (size_t)buf * STAGE * D;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int slot = r * T + tid;
            const bool right = rows[r] >= 0 && (xv[r] > tv);                     // node.cpp:339
            const unsigned int m = __ballot_sync(0xffffffffu, right);
            if (lane == 0) { smask[buf * (STAGE / 32) + slot / 32] = m; if (PASS == 0 && m) atomicAdd(s_nright, __popc(m)); }
            if (rows[r] >= 0) {
#pragma unroll
                for (int d = 0; d < D; ++d) g[(size_t)slot * D + d] = gpre[r][d];
            }
        }
    };
    // the lane's R rows x D values of sub-block w (chain order), their side bits and validity bits
    auto lane_vals = [&](int buf, int w, int cnt, float (&v)[8], unsigned int &mb, unsigned int &vb) {
        const int rb = w * SUB + lane * R;
        const float *p = sg + (size_t)buf * STAGE * D + (size_t)rb * D;
        if (EPL == 8) {
            const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                if (j < EPL) { const float2 a = *reinterpret_cast<const float2 *>(p + j); v[j] = a.x; v[j + 1] = a.y; }
                else { v[j] = 0.0f; v[j + 1] = 0.0f; }
            }
        }
        mb = (smask[buf * (STAGE / 32) + (rb >> 5)] >> (rb & 31)) & ((1u << R) - 1u);
        const int nv = min(max(cnt - rb, 0), R);
        vb = (1u << nv) - 1u;
    };
    // the lane's KE consecutive elements of chain c in sub-block w (0 for rows of the other side / past the end)
    auto load_x = [&](int buf, int w, int c, int cnt, float (&x)[KE]) {
        float v[8];
        unsigned int mb, vb;
        lane_vals(buf, w, cnt, v, mb, vb);
        if (PASS == 0) {
            const int side = c / D, d = c - side * D;
            const unsigned int sel = (side ? mb : ~mb) & vb;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float val = v[r * D];
#pragma unroll
                for (int dd = 1; dd < D; ++dd) val = (d == dd) ? v[r * D + dd] : val;
                x[r < KE ? r : 0] = ((sel >> r) & 1u) ? val : 0.0f;
            }
        } else {
            const unsigned int sel = (c ? mb : ~mb) & vb;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r = j / D, d = j - r * D;
                x[j < KE ? j : 0] = (j < EPL && ((sel >> r) & 1u)) ? v[j] * smean[c * D + d] : 0.0f;
            }
        }
    };

    if (tid < 8) sh.state[tid] = 0.0f;
    {
        int r0[R];
        load_rows(0, r0);
        load_rows(1, rows_n);
        gather(r0);
        commit(0);
    }
    __syncthreads();
    for (int st = 0; st < n_stages; ++st) {
        const int buf = st & 1;
        const int cnt = min(STAGE, n - st * STAGE);
        int rows_nn[R];
        if (st + 1 < n_stages) gather(rows_n);
        load_rows(st + 2, rows_nn);
        auto lx = [&](int c, int w, float (&x)[KE]) { load_x(buf, w, c, cnt, x); };
        seq::run_stage<KE>(sh, NCH, (cnt + SUB - 1) / SUB, lx, n_fast, n_slow, n_seq);   // ends with a barrier
        if (st + 1 < n_stages) commit(buf ^ 1);
#pragma unroll
        for (int r = 0; r < R; ++r) rows_n[r] = rows_nn[r];
        __syncthreads();
    }
}

template <int DCT>
__global__ void __launch_bounds__(512, 1) replay_par_kernel(ReplayParams P, NodeArrays na, Ctl *ctl_stats, const int *mode, int replay_cap) {
    using C = ParCfg<DCT>;
    constexpr int D = DCT, STAGE = C::STAGE;
    __shared__ __align__(16) float sg[2 * STAGE * D];
    __shared__ unsigned int smask[2 * (STAGE / 32)];
    __shared__ float smean[2 * D];
    __shared__ seq::StageShared sh;
    __shared__ int s_nright;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_items = min(P.ctl->n_replay, replay_cap);
    int n_fast = 0, n_slow = 0, n_seq = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        if (mode != nullptr && mode[it] != 2) continue;          // streamed by replay_stream_kernel
        const ReplayItem item = P.items[it];
        const int h = item.node, cand = item.cand;
        const int s0 = na.seg_start[h], n = na.seg_len[h];
        const int f = cand >= 0 ? cand / P.B : 0;
        const float tv = cand >= 0 ? P.thr[cand] : INFINITY;
        if (threadIdx.x == 0) s_nright = 0;
        __syncthreads();
        replay_pass_par<DCT, 0>(P, s0, n, f, tv, cand >= 0, sg, smask, smean, &s_nright, sh, n_fast, n_slow, n_seq);
        const int nR = s_nright;   // all ballots are committed before the last barrier of the pass
        const int nL = n - nR;
        const bool invalid = cand >= 0 && (nL < P.min_data || nR < P.min_data);
        const float lcf = (float)nL, rcf = (float)nR;
        float ln = 0.0f, rn = 0.0f;
        __syncthreads();
        if (warp == 0) {
            float lrec, rrec;
            if (cand >= 0) { lrec = nL > 0 ? 1.0f / lcf : 0.0f; rrec = nR > 0 ? 1.0f / rcf : 0.0f; }
            else { lrec = 1.0f / lcf; rrec = 0.0f; }   // parent: n_samples_recip = 1/n (split_candidate_generator.cpp:265,296)
            if (lane < D) { smean[lane] = sh.state[lane] * lrec; smean[D + lane] = sh.state[D + lane] * rrec; }
            __syncwarp();
            for (int d = 0; d < D; ++d) { ln = ln + smean[d] * smean[d]; rn = rn + smean[D + d] * smean[D + d]; }  // squared_norm
        }
        __syncthreads();
        float result = 0.0f;
        if (P.score_func == GBRL_B200_SCORE_L2) {
            result = cand >= 0 ? (lcf * ln + rcf * rn) : (ln * lcf);
        } else {
            replay_pass_par<DCT, 1>(P, s0, n, f, tv, cand >= 0, sg, smask, smean, &s_nright, sh, n_fast, n_slow, n_seq);
            __syncthreads();
            const float fnum = sh.state[0], tnum = sh.state[1];
            if (cand >= 0) {
                const float num = tnum + fnum;
                const float den = rn * rcf + ln * lcf;
                result = (den == 0.0f) ? 0.0f : num / sqrtf(den);
            } else {
                const float den = ln * lcf;
                result = (n == 0 || den == 0.0f) ? 0.0f : fnum / sqrtf(den);
            }
        }
        if (invalid) result = -INFINITY;
        if (threadIdx.x == 0) P.out[it] = result;
        __syncthreads();
    }
    if (lane == 0 && warp < 8 && (n_fast | n_slow)) {
        atomicAdd((unsigned long long *)&ctl_stats->stat_chain_fast, (unsigned long long)n_fast);
        atomicAdd((unsigned long long *)&ctl_stats->stat_chain_slow, (unsigned long long)n_slow);
        atomicAdd((unsigned long long *)&ctl_stats->stat_chain_seq, (unsigned long long)n_seq);
    }
}


// ---------------------------------------------------------------- replay, streamed (the default path for D <= 4)
// A single CTA cannot gather a large node fast enough (one 32-byte sector per row for the feature value, one for the
// gradients): the whole GPU prepares contiguous streams first, the chain CTA of an item then only streams them.
//   replay_plan_kernel    word offset of every item's side-bit plane (8-word groups); items that do not fit -> direct
//   replay_gather_kernel  order-space copy of build_grads for the rows of every node that has replay items
//   replay_bits_kernel    side bit (x > threshold, node.cpp:339) of every (item, row), 32 rows per word, + right counts
//   replay_stream_kernel  the chains (chain.cuh) over the streams; same arithmetic as replay_par_kernel

__global__ void __launch_bounds__(1024) replay_plan_kernel(ReplayParams P, NodeArrays na, StreamParams S) { replay_plan_body<1024>(P, na, S); }

__global__ void __launch_bounds__(256) replay_gather_kernel(ReplayParams P, NodeArrays na, StreamParams S) { replay_gather_body(P, na, S); }

// one warp per 8-word group (256 rows) of a plane
__global__ void __launch_bounds__(256) replay_bits_kernel(ReplayParams P, NodeArrays na, StreamParams S) {
    const int n_items = min(P.ctl->n_replay, S.replay_cap);
    if (n_items <= 0) return;
    const int lane = threadIdx.x & 31;
    const int total_groups = S.woff[n_items] >> 3;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < total_groups; g += warps) {
        const int gw = g << 3;
        // last item whose offset is <= gw (upper bound - 1): items without words share the offset of their successor
        int lo = 0, hi = n_items;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (S.woff[mid] <= gw) lo = mid + 1; else hi = mid;
        }
        const int it = lo - 1;
        if (S.mode[it] != 0) continue;
        const ReplayItem item = P.items[it];
        const int s0 = na.seg_start[item.node], n = na.seg_len[item.node];
        const int f = item.cand / P.B, jb = item.cand - f * P.B;
        const float tv = P.thr[item.cand];
        const int k0 = (gw - S.woff[it]) << 5;
        int cnt = 0;
        bool rt[8];
        if (P.codesT != nullptr) {
            const uint16_t *col = P.codesT + (size_t)f * P.codesT_stride + P.row_offset;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = k0 + j * 32 + lane;
                rt[j] = k < n && (int)col[P.order[s0 + k]] > jb;              // x > thr[f][jb] <=> code > jb
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = k0 + j * 32 + lane;
                rt[j] = k < n && P.X[(size_t)P.order[s0 + k] * P.F + f] > tv;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const unsigned int m = __ballot_sync(0xffffffffu, rt[j]);           // node.cpp:339
            if (lane == j) S.bits[gw + j] = m;
            cnt += __popc(m);
        }
        if (lane == 0 && cnt) atomicAdd(&S.nright[it], cnt);
    }
}

template <int DCT, int PASS>
__device__ __forceinline__ void stream_pass(const float *G, const unsigned int *W, int n, float *sg /*[3][STAGE*D]*/, unsigned int *smask /*[3][MW]*/,
                                            const float *smean, seq::StageShared &sh, int &n_fast, int &n_slow, int &n_seq) {
    using C = ParCfg<DCT>;
    constexpr int R = C::R, T = C::T, STAGE = C::STAGE, SUB = C::SUB, EPL = C::EPL, D = DCT, MW = STAGE / 32;
    constexpr int NCH = PASS == 0 ? 2 * D : 2;
    constexpr int KE = PASS == 0 ? R : 8;
    const int tid = threadIdx.x, lane = tid & 31;
    const int n_stages = (n + STAGE - 1) / STAGE;
    const long long ne = (long long)n * D;
    float nx1[8], nx2[8];
    unsigned int mw1 = 0, mw2 = 0;
    auto load = [&](int st, float (&dst)[8], unsigned int &mw) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j < EPL) {
                const long long e = (long long)st * (STAGE * D) + j * T + tid;
                dst[j] = (st < n_stages && e < ne) ? G[e] : 0.0f;
            }
        }
        mw = 0u;
        if (tid < MW && W != nullptr && st < n_stages) {
            const int wi = st * MW + tid;
            if ((long long)wi * 32 < n) mw = W[wi];
        }
    };
    auto commit = [&](int b, const float (&src)[8], unsigned int mw) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < EPL) sg[b * (STAGE * D) + j * T + tid] = src[j];
        if (tid < MW) smask[b * MW + tid] = mw;
    };
    // the lane's KE consecutive elements of chain c in sub-block w of stage buffer b; rows past the end hold +0 and
    // count as "left"
    auto load_x = [&](int b, int c, int w, float (&x)[KE]) {
        const int rb = w * SUB + lane * R;
        const float *p = sg + b * (STAGE * D) + rb * D;
        float v[8];
        if (EPL == 8) {
            const float4 a = *reinterpret_cast<const float4 *>(p), q = *reinterpret_cast<const float4 *>(p + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = q.x; v[5] = q.y; v[6] = q.z; v[7] = q.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                if (j < EPL) { const float2 a = *reinterpret_cast<const float2 *>(p + j); v[j] = a.x; v[j + 1] = a.y; }
                else { v[j] = 0.0f; v[j + 1] = 0.0f; }
            }
        }
        const unsigned int mb = (smask[b * MW + (rb >> 5)] >> (rb & 31)) & ((1u << R) - 1u);
        if (PASS == 0) {
            const int side = c / D, d = c - side * D;
            const unsigned int sel = side ? mb : ~mb;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float val = v[r * D];
#pragma unroll
                for (int dd = 1; dd < D; ++dd) val = (d == dd) ? v[r * D + dd] : val;
                x[r < KE ? r : 0] = ((sel >> r) & 1u) ? val : 0.0f;
            }
        } else {
            const unsigned int sel = c ? mb : ~mb;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r = j / D, d = j - r * D;
                x[j < KE ? j : 0] = (j < EPL && ((sel >> r) & 1u)) ? v[j] * smean[c * D + d] : 0.0f;
            }
        }
    };
    seq::pipe_init(sh);
    load(0, nx1, mw1); commit(0, nx1, mw1);
    load(1, nx1, mw1); commit(1, nx1, mw1);
    load(2, nx1, mw1); load(3, nx2, mw2);
    __syncthreads();
    int bc = 0;
    for (int st = 0; st < n_stages; ++st) {
        const int bn = bc == 2 ? 0 : bc + 1, bf = bn == 2 ? 0 : bn + 1;
        const int cnt = min(STAGE, n - st * STAGE);
        const int cnt_next = st + 1 < n_stages ? min(STAGE, n - (st + 1) * STAGE) : 0;
        auto lc = [&](int c, int w, float (&x)[KE]) { load_x(bc, c, w, x); };
        auto ln = [&](int c, int w, float (&x)[KE]) { load_x(bn, c, w, x); };
        seq::run_stage_pipe<KE>(sh, NCH, st & 1, (cnt + SUB - 1) / SUB, lc, (cnt_next + SUB - 1) / SUB, ln, n_fast, n_slow, n_seq);
        commit(bf, nx1, mw1);
#pragma unroll
        for (int j = 0; j < 8; ++j) nx1[j] = nx2[j];
        mw1 = mw2;
        load(st + 4, nx2, mw2);
        bc = bn;
        __syncthreads();
    }
}

template <int DCT>
__global__ void __launch_bounds__(512, 1) replay_stream_kernel(ReplayParams P, NodeArrays na, StreamParams S, Ctl *ctl_stats) {
    using C = ParCfg<DCT>;
    constexpr int D = DCT, STAGE = C::STAGE;
    extern __shared__ __align__(16) float s_dyn_stream[];         // [3][STAGE*D] floats, then [3][STAGE/32] mask words
    float *sg = s_dyn_stream;
    unsigned int *smask = reinterpret_cast<unsigned int *>(sg + 3 * STAGE * D);
    __shared__ float smean[2 * D];
    __shared__ seq::StageShared sh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_items = min(P.ctl->n_replay, S.replay_cap);
    int n_fast = 0, n_slow = 0, n_seq = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int md = S.mode[it];
        if (md == 2) continue;                                   // replay_par_kernel gathers this one itself
        const ReplayItem item = P.items[it];
        const int h = item.node, cand = item.cand;
        const int s0 = na.seg_start[h], n = na.seg_len[h];
        const float *G = S.G + (size_t)s0 * D;
        const unsigned int *W = md == 0 ? S.bits + S.woff[it] : nullptr;
        stream_pass<DCT, 0>(G, W, n, sg, smask, smean, sh, n_fast, n_slow, n_seq);
        const int nR = md == 0 ? S.nright[it] : 0;
        const int nL = n - nR;
        const bool invalid = cand >= 0 && (nL < P.min_data || nR < P.min_data);
        const float lcf = (float)nL, rcf = (float)nR;
        float ln = 0.0f, rn = 0.0f;
        if (warp == 0) {
            float lrec, rrec;
            if (cand >= 0) { lrec = nL > 0 ? 1.0f / lcf : 0.0f; rrec = nR > 0 ? 1.0f / rcf : 0.0f; }
            else { lrec = 1.0f / lcf; rrec = 0.0f; }   // parent: n_samples_recip = 1/n (split_candidate_generator.cpp:265,296)
            if (lane < D) { smean[lane] = sh.state[lane] * lrec; smean[D + lane] = sh.state[D + lane] * rrec; }
            __syncwarp();
            for (int d = 0; d < D; ++d) { ln = ln + smean[d] * smean[d]; rn = rn + smean[D + d] * smean[D + d]; }  // squared_norm
        }
        __syncthreads();
        float result = 0.0f;
        if (P.score_func == GBRL_B200_SCORE_L2) {
            result = cand >= 0 ? (lcf * ln + rcf * rn) : (ln * lcf);
        } else {
            stream_pass<DCT, 1>(G, W, n, sg, smask, smean, sh, n_fast, n_slow, n_seq);
            const float fnum = sh.state[0], tnum = sh.state[1];
            if (cand >= 0) {
                const float num = tnum + fnum;
                const float den = rn * rcf + ln * lcf;
                result = (den == 0.0f) ? 0.0f : num / sqrtf(den);
            } else {
                const float den = ln * lcf;
                result = (n == 0 || den == 0.0f) ? 0.0f : fnum / sqrtf(den);
            }
        }
        if (invalid) result = -INFINITY;
        if (threadIdx.x == 0) P.out[it] = result;
        __syncthreads();
    }
    if (lane == 0 && warp < 8 && (n_fast | n_slow)) {
        atomicAdd((unsigned long long *)&ctl_stats->stat_chain_fast, (unsigned long long)n_fast);
        atomicAdd((unsigned long long *)&ctl_stats->stat_chain_slow, (unsigned long long)n_slow);
        atomicAdd((unsigned long long *)&ctl_stats->stat_chain_seq, (unsigned long long)n_seq);
    }
}

// ---------------------------------------------------------------- decisions
struct DecideParams {
    int level, nT, F, B, D, C, max_depth, nn;
    const long long *hist_cur;
    const float *thr, *fw;
    const int *rev_map;
    const ReplayItem *replay;
    const float *replay_scores;
    const int *obl_cands;
    const float *replay_exact;   // exact-tier scores of the replay items (select_greedy_kernel)
    Ctl *ctl;                    // decisions: the live control block; verification of a speculative level: its snapshot
    Ctl *ctl_stats;              // the live control block (statistics)
    const int *state_snap;       // verification: node states before the level's decision
    int use_replay;              // 0: speculative decision on the exact-tier winner (the replay of this level is still running)
    int count_stats;             // 0: this level was counted already (second decision of a rolled-back level)
    int force_flip;              // tests (GBRL_B200_SPEC_FORCE_FLIP=1): speculate on a WRONG candidate wherever there is a choice
    const unsigned int *abort_flag;   // speculative tree: lowest level known to need a second decision (0xffffffff: none)
};

// suffix sum over codes > j of feature f in the node's histogram (one warp)
__device__ void warp_right_sums(const long long *Hn, int nT, int D, int f, int j, long long *out /*[1+D], lane 0*/) {
    const int lane = threadIdx.x & 31;
    const int tile = f / FT, fl = f % FT, HS = 1 + D;
    const long long *Ht = Hn + (size_t)tile * NB * FT * HS;
    for (int d = 0; d <= D; ++d) {
        long long s = 0;
        for (int b = lane; b < NB; b += 32)
            if (b >= j) s += Ht[((size_t)b * FT + fl) * HS + d];
        for (int o = 16; o > 0; o >>= 1) s += shfl_down_ll(s, o);
        if (lane == 0) out[d] = s;
    }
}

__device__ void apply_split(NodeArrays na, const DecideParams &P, int h, int cand, const long long *Hn) {
    // called by one warp; writes the split and initialises both children
    const int lane = threadIdx.x & 31;
    const int f = cand / P.B, j = cand % P.B;
    __shared__ long long s_r[32][65];
    long long *r = s_r[(threadIdx.x >> 5) & 31];
    warp_right_sums(Hn, P.nT, P.D, f, j, r);
    __syncwarp();
    if (lane == 0) {
        const int n = na.seg_len[h], s0 = na.seg_start[h];
        const int nR = (int)r[0], nL = n - nR;
        na.state[h] = NODE_SPLIT;
        na.split_f[h] = f; na.split_j[h] = j; na.split_thr[h] = P.thr[cand];
        const int lc = 2 * h + 1, rc = 2 * h + 2;
        na.seg_start[lc] = s0; na.seg_len[lc] = nL;
        na.seg_start[rc] = s0 + nL; na.seg_len[rc] = nR;
        for (int d = 0; d < P.D; ++d) {
            na.tot_sum[(size_t)rc * P.D + d] = r[1 + d];
            na.tot_sum[(size_t)lc * P.D + d] = na.tot_sum[(size_t)h * P.D + d] - r[1 + d];
        }
        const bool last = (P.level + 1 >= P.max_depth);
        na.state[lc] = last ? NODE_LEAF : NODE_OPEN;
        na.state[rc] = last ? NODE_LEAF : NODE_OPEN;
    }
    __syncwarp();
}

// greedy, one warp: the decision of node h in the reference's arithmetic -- arg-max over the replayed candidates where the node
// has replay items, the exact-tier winner otherwise.  Same value in every lane.
__device__ void final_choice_greedy(const DecideParams &P, NodeArrays na, int p, int h, float &g, int &c) {
    const int lane = threadIdx.x & 31;
    g = na.best_gain[h];
    c = na.best_idx[h];
    const int rc = na.rep_count[h];
    if (rc > 0) {
        // final arg-max over the replayed candidates, in the reference's arithmetic
        const int rb = na.rep_begin[h];
        float parent = 0.0f;
        float bg_ = -INFINITY; int bi = INT_MAX;
        if (P.level > 0) parent = P.replay_scores[rb];   // first item is the parent
        for (int i = rb + lane; i < rb + rc; i += 32) {
            const int cand = P.replay[i].cand;
            if (cand < 0) continue;
            const float sw = P.replay_scores[i] * P.fw[cand / P.B];
            const float gain = sw - parent;
            if (gain > -INFINITY && better(gain, cand, bg_, bi)) { bg_ = gain; bi = cand; }
            // calibration statistic of the band: observed rounding noise of the reference's sum, in units of
            // 2^-24 * sqrt(n) * |score*w|  (exact-tier score*w = stored gain + exact parent)
            const float sw_exact = P.replay_exact[i] + na.parent_score[h];
            const float unit = U24 * sqrtf((float)na.seg_len[h]) * fabsf(sw_exact);
            if (unit > 0.0f && gain > -INFINITY && P.count_stats) atomicMax(&P.ctl_stats->stat_max_noise, __float_as_uint(fabsf(sw - sw_exact) / unit));
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float og = __shfl_down_sync(0xffffffffu, bg_, o);
            const int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (better(og, oi, bg_, bi)) { bg_ = og; bi = oi; }
        }
        bg_ = __shfl_sync(0xffffffffu, bg_, 0); bi = __shfl_sync(0xffffffffu, bi, 0);
        if (bi != INT_MAX) {
            if (lane == 0 && bi != c && P.count_stats) atomicAdd((unsigned long long *)&P.ctl_stats->stat_replay_flips, 1ull);
            g = bg_; c = bi;
        }
    }
    (void)p;
}

// greedy: one warp per node of the level
__device__ void decide_greedy_node(const DecideParams &P, NodeArrays na, int p) {
    const int h = level_base(P.level) + p, lane = threadIdx.x & 31;
    if (na.state[h] != NODE_OPEN) return;
    float g; int c;
    if (P.use_replay) final_choice_greedy(P, na, p, h, g, c);
    else {
        g = na.best_gain[h]; c = na.best_idx[h];
        if (P.force_flip && na.rep_count[h] > 0) {
            // tests: take the first other candidate of the replay list, so that the verification has to roll this level back
            const int rb = na.rep_begin[h], rc = na.rep_count[h];
            for (int i = rb; i < rb + rc; ++i) {
                const int cand = P.replay[i].cand;
                if (cand >= 0 && cand != c) { c = cand; g = 0.0f; break; }
            }
        }
    }
    // fitter.cpp:357: split iff best_score >= 0 (and the node could be split at all)
    const bool split = (c >= 0) && (g >= 0.0f);
    if (!split) {
        if (lane == 0) na.state[h] = NODE_LEAF;
        return;
    }
    const long long *Hn = P.hist_cur + (size_t)p * P.nT * NB * FT * (1 + P.D);
    apply_split(na, P, h, c, Hn);
    // fitter.cpp:300: a child that is empty (or at max depth) is never evaluated -> leaf
    if (lane == 0) {
        const int lc = 2 * h + 1, rcn = 2 * h + 2;
        if (na.seg_len[lc] == 0) na.state[lc] = NODE_LEAF;
        if (na.seg_len[rcn] == 0) na.state[rcn] = NODE_LEAF;
    }
}

// oblivious, one thread: the level's candidate in the reference's arithmetic (fitter.cpp:427-445)
__device__ int final_choice_oblivious(const DecideParams &P) {
    int c = P.ctl->obl_best_idx;
    const int nrep = P.ctl->obl_has_replay;
    if (c >= 0 && nrep > 1) {
        float bg_ = -INFINITY; int bi = INT_MAX;
        for (int w = 0; w < nrep; ++w) {
            const int cand = P.obl_cands[w];
            float s = 0.0f;
            for (int p = 0; p < P.nn; ++p) s += P.replay_scores[(size_t)w * P.nn + p];
            s = s * P.fw[P.rev_map[cand / P.B]];
            if (s > -INFINITY && better(s, cand, bg_, bi)) { bg_ = s; bi = cand; }
        }
        if (bi != INT_MAX) {
            if (bi != c && P.count_stats) atomicAdd((unsigned long long *)&P.ctl_stats->stat_replay_flips, 1ull);
            c = bi;
        }
    }
    return c;
}

// oblivious: one CTA, warp w handles nodes w, w + #warps, ...
__device__ void decide_oblivious_body(const DecideParams &P, NodeArrays na) {
    __shared__ int s_cand;
    const int base = level_base(P.level);
    if (na.state[base] != NODE_OPEN) return;
    if (threadIdx.x == 0) {
        int c;
        if (P.use_replay) c = final_choice_oblivious(P);
        else {
            c = P.ctl->obl_best_idx;
            if (P.force_flip && c >= 0 && P.ctl->obl_has_replay > 1) {
                for (int w = 0; w < P.ctl->obl_has_replay; ++w)
                    if (P.obl_cands[w] != c) { c = P.obl_cands[w]; break; }
            }
        }
        s_cand = c;
        P.ctl->obl_depth = (c >= 0) ? P.level + 1 : P.level;
    }
    __syncthreads();
    const int c = s_cand;
    const int warp = threadIdx.x >> 5;
    for (int p = warp; p < P.nn; p += (int)(blockDim.x >> 5)) {
        const int h = base + p;
        if (c < 0) {
            if ((threadIdx.x & 31) == 0) na.state[h] = NODE_LEAF;   // fitter.cpp:458 break
        } else {
            const long long *Hn = P.hist_cur + (size_t)p * P.nT * NB * FT * (1 + P.D);
            apply_split(na, P, h, c, Hn);
        }
    }
}

// One CTA: the split decisions of level `level` (fitter.cpp:338-371 / 448-477), then -- the children's segments and states
// being known -- the histogram work items of level + 1 (plan.cuh).  One launch instead of three per level.
__global__ void __launch_bounds__(1024) decide_plan_kernel(DecideParams P, NodeArrays na, PlanParams Q, int oblivious) {
    // A speculative tree that is already known to return to a shallower level stops deciding: the nodes of this level stay
    // open, no child is created, and every later kernel of the tree finds nothing to do.
    __shared__ int s_abort;
    if (threadIdx.x == 0) s_abort = (P.abort_flag != nullptr && *(volatile const unsigned int *)P.abort_flag < (unsigned int)P.level) ? 1 : 0;
    __syncthreads();
    if (s_abort) { /* nothing */ }
    else if (oblivious) decide_oblivious_body(P, na);
    else {
        for (int p = threadIdx.x >> 5; p < P.nn; p += (int)(blockDim.x >> 5)) decide_greedy_node(P, na, p);
    }
    __threadfence_block();
    __syncthreads();
    if (threadIdx.x == 0) P.ctl->n_replay = 0;      // the replay list of the next level starts empty (select_*_kernel appends)
    if (P.level + 1 < P.max_depth)
        plan_level_body(na, P.ctl, Q.items, Q.items_cap, P.level + 1, Q.max_depth, Q.nT_local, Q.use_subtraction, Q.oblivious, Q.row_groups,
                        Q.row_group, Q.item_rows_max);
}

// Verification of a speculative level (side stream, after the level's replay): the decision in the reference's arithmetic
// against the one the tree was grown on.  The lowest level with a difference is where grow_tree returns to.
__global__ void __launch_bounds__(1024) verify_kernel(DecideParams P, NodeArrays na, int oblivious, unsigned int *flip_level) {
    const int base = level_base(P.level);
    if (oblivious) {
        if (threadIdx.x != 0 || P.state_snap[base] != NODE_OPEN) return;      // the tree had stopped before this level
        const int c = final_choice_oblivious(P);
        const int taken = na.state[base] == NODE_SPLIT ? na.split_f[base] * P.B + na.split_j[base] : -1;
        if (c != taken) atomicMin(flip_level, (unsigned int)P.level);
        return;
    }
    const int lane = threadIdx.x & 31;
    for (int p = threadIdx.x >> 5; p < P.nn; p += (int)(blockDim.x >> 5)) {
        const int h = base + p;
        if (P.state_snap[h] != NODE_OPEN) continue;
        float g; int c;
        final_choice_greedy(P, na, p, h, g, c);
        const bool split = (c >= 0) && (g >= 0.0f);
        const bool split_taken = na.state[h] == NODE_SPLIT;
        const int taken = na.split_f[h] * P.B + na.split_j[h];
        if (lane == 0 && (split != split_taken || (split && c != taken))) atomicMin(flip_level, (unsigned int)P.level);
    }
}

// Return to the state before the decision of level L: node states of that level from their snapshot, deeper nodes gone (the row
// order, the node of every position and the histograms of level L are still in their per-level buffers).
__global__ void __launch_bounds__(256) rollback_kernel(NodeArrays na, const int *state_snap, int MAXN, int L, unsigned int *flip_level) {
    const int lo = level_base(L), hi = level_base(L + 1);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *flip_level = 0xffffffffu;
    if (i < MAXN) {
        if (i >= hi) na.state[i] = NODE_NONE;
        else if (i >= lo) na.state[i] = state_snap[i];
    }
}

// ---------------------------------------------------------------- launchers
template <int DM>
static void launch_scan_dm(const ScanParams &P, const NodeArrays &na, int grid, cudaStream_t s) {
    GB_LAUNCH(scan_kernel<DM>, grid, SCAN_THREADS, 0, s, P, na);
}

void launch_scan(Model &m, int level, cudaStream_t s) {
    Workspace &ws = m.ws;
    ScanParams P;
    P.level = level; P.nT = ws.nT; P.F = ws.F; P.B = ws.B; P.D = ws.D; P.C = ws.F * ws.B;
    P.score_func = m.cfg.split_score_func; P.oblivious = m.cfg.grow_policy == GBRL_B200_GROW_OBLIVIOUS;
    P.min_data = m.cfg.min_data_in_leaf; P.max_depth = m.cfg.max_depth; P.use_subtraction = m.cfg.use_subtraction;
    P.hist_cur = ws.hist_p[level & 1];
    P.hist_par = ws.hist_p[(level + 1) & 1];
    P.thr = ws.thr.as<float>(); P.fw = m.feature_weights.as<float>();
    P.scores = ws.scores.as<float>(); P.cand_flags = ws.cand_flags.as<uint8_t>();
    P.tile_best = ws.tile_best.as<float2>(); P.ctl = ws.ctl.as<Ctl>();
    const int grid = (1 << level) * ws.F;
    const int D = ws.D;
    if (D <= 1) launch_scan_dm<1>(P, ws.na, grid, s);
    else if (D <= 2) launch_scan_dm<2>(P, ws.na, grid, s);
    else if (D <= 4) launch_scan_dm<4>(P, ws.na, grid, s);
    else if (D <= 8) launch_scan_dm<8>(P, ws.na, grid, s);
    else if (D <= 16) launch_scan_dm<16>(P, ws.na, grid, s);
    else if (D <= 32) launch_scan_dm<32>(P, ws.na, grid, s);
    else launch_scan_dm<64>(P, ws.na, grid, s);      // gbrl_b200_create caps output_dim at 64
}

template <int DCT>
static void launch_stream(const ReplayParams &R, const NodeArrays &na, const StreamParams &S, Ctl *ctl, int n_sms, cudaStream_t s) {
    using C = ParCfg<DCT>;
    const size_t smem = (size_t)3 * C::STAGE * DCT * sizeof(float) + (size_t)3 * (C::STAGE / 32) * sizeof(unsigned int);
    ensure_dyn_smem(replay_stream_kernel<DCT>, smem);
    GB_LAUNCH((replay_stream_kernel<DCT>), n_sms * 2, 512, smem, s, R, na, S, ctl);
}

// slot != nullptr: speculative level.  The caller has bound the slot's buffers into the workspace (tree.cu SlotBind); the selection
// runs on `s`, the replay of the selected items on the slot's side stream, against a snapshot of the control block.
void launch_select_and_replay(Model &m, const float *X, int level, cudaStream_t s, ReplaySlot *slot) {
    Workspace &ws = m.ws;
    const bool obl = m.cfg.grow_policy == GBRL_B200_GROW_OBLIVIOUS;
    const int C = ws.F * ws.B, nn = 1 << level;
    Ctl *ctl = ws.ctl.as<Ctl>();
    // (ctl->n_replay is zero here: init_nodes_kernel for level 0, the decide + plan kernel of the level before otherwise)
    // default band: 6 noise units.  The largest noise ever observed on a replayed candidate is 1.35 units (max_noise_ratio), so two
    // candidates can swap when they are <= 2.7 units apart; the observed spread (sigma ~ 0.35 units) puts 6 units at > 12 sigma.
    const float kappa = m.cfg.band_kappa > 0 ? m.cfg.band_kappa : 6.0f;
    int *obl_cands = reinterpret_cast<int *>(ws.replay.as<ReplayItem>() + ws.replay_cap);
    if (!obl) {
        SelectParams P;
        P.level = level; P.nT = ws.nT; P.F = ws.F; P.B = ws.B; P.D = ws.D; P.C = C; P.tie_replay = m.cfg.tie_replay;
        P.replay_cap = ws.replay_cap; P.kappa = kappa; P.scores = ws.scores.as<float>();
        P.cand_flags = ws.cand_flags.as<uint8_t>(); P.tile_best = ws.tile_best.as<float2>();
        P.replay = ws.replay.as<ReplayItem>(); P.replay_exact = ws.replay_scores.as<float>() + ws.replay_cap; P.ctl = ctl;
        P.state_snap = ws.spec ? ws.state_snap.as<int>() : nullptr;
        GB_LAUNCH(select_greedy_kernel, nn, 1024, 0, s, P, ws.na);
    } else {
        OblParams P;
        P.level = level; P.F = ws.F; P.B = ws.B; P.D = ws.D; P.C = C; P.nn = nn; P.tie_replay = m.cfg.tie_replay;
        P.replay_cap = ws.replay_cap; P.nblocks = ceil_div(C, 256); P.kappa = kappa; P.N = ws.N;
        P.scores = ws.scores.as<float>(); P.cand_flags = ws.cand_flags.as<uint8_t>(); P.fw = m.feature_weights.as<float>();
        P.rev_map = m.rev_num_map.as<int>(); P.obl_tot = ws.obl_tot.as<float>(); P.obl_nb = ws.obl_tot.as<float>() + C;
        P.blk_best = ws.tile_best.as<float2>(); P.replay = ws.replay.as<ReplayItem>(); P.obl_cands = obl_cands; P.ctl = ctl;
        P.state_snap = ws.spec ? ws.state_snap.as<int>() : nullptr;
        GB_LAUNCH(obl_reduce_kernel, P.nblocks, 256, 0, s, P, ws.na);
        GB_LAUNCH(obl_select_kernel, 1, 1024, 0, s, P, ws.na);
    }
    if (!m.cfg.tie_replay) return;
    const Ctl *rctl = ctl;         // what the replay kernels read the item count from
    cudaStream_t rs = s;           // the stream they run on
    if (slot) {
        GB_CUDA(cudaMemcpyAsync(slot->ctl_snap.p, ctl, sizeof(Ctl), cudaMemcpyDeviceToDevice, s));
        GB_CUDA(cudaEventRecord(slot->ev_sel, s));
        GB_CUDA(cudaStreamWaitEvent(slot->stream, slot->ev_sel, 0));
        rctl = slot->ctl_snap.as<Ctl>();
        rs = slot->stream;
        slot->pending = true;
    }
    ReplayParams R;
    R.F = ws.F; R.B = ws.B; R.D = ws.D; R.score_func = m.cfg.split_score_func; R.min_data = m.cfg.min_data_in_leaf;
    R.X = X; R.codesT = ws.use_codesT ? ws.codesT.as<uint16_t>() : nullptr; R.codesT_stride = ws.codesT_stride; R.row_offset = ws.row_offset;
    R.bg = ws.bg.as<float>(); R.order = ws.order_p[0]; R.thr = ws.thr.as<float>();
    R.items = ws.replay.as<ReplayItem>(); R.out = ws.replay_scores.as<float>(); R.ctl = rctl;
    // rows per thread per stage: 8 (1024-row stages) for D == 1 down to 1 for wide outputs, so that a stage's
    // gradients fit in registers while they are in flight and two stages fit in shared memory
    const int D = ws.D;
    if (D <= 4) {
        // bit-exact parallel chains (chain.cuh) over streams prepared by the whole GPU
        StreamParams S;
        S.G = ws.rgrad.as<float>(); S.bits = ws.rbits.as<unsigned int>(); S.woff = ws.rmeta.as<int>();
        S.mode = S.woff + ws.replay_cap + 1; S.nright = S.mode + ws.replay_cap;
        // GBRL_B200_REPLAY_DIRECT=1 (tests): no plane fits, every item takes the direct-gather kernel
        static const bool force_direct = getenv("GBRL_B200_REPLAY_DIRECT") != nullptr && getenv("GBRL_B200_REPLAY_DIRECT")[0] == '1';
        S.cap_words = force_direct ? 0 : ws.rbits_words; S.replay_cap = ws.replay_cap; S.N = ws.N;
        // a speculative level gathers every row: `nid` moves on with the deeper levels while this level is replayed
        S.oblivious = (obl || slot) ? 1 : 0;
        S.pnode = ws.pnode_p[0];
        S.wide = (D <= 2 && (m.cfg.replay_variant & 1) == 0) ? 1 : 0;
        GB_LAUNCH(replay_plan_kernel, 1, 1024, 0, rs, R, ws.na, S);
        GB_LAUNCH(replay_gather_kernel, ws.n_sms * replay_grid_mult(slot != nullptr), 256, 0, rs, R, ws.na, S);
        if (S.wide) {
            // chains spread over the whole GPU (replay_wide.cu); items whose plane did not fit are gathered directly
            launch_replay_wide(m, R, S, rs);
            if (D <= 1) GB_LAUNCH((replay_par_kernel<1>), ws.n_sms * 2, 512, 0, rs, R, ws.na, ctl, S.mode, ws.replay_cap);
            else GB_LAUNCH((replay_par_kernel<2>), ws.n_sms * 2, 512, 0, rs, R, ws.na, ctl, S.mode, ws.replay_cap);
            return;
        }
        GB_LAUNCH(replay_bits_kernel, ws.n_sms * replay_grid_mult(slot != nullptr), 256, 0, rs, R, ws.na, S);
        if (D <= 1) {
            launch_stream<1>(R, ws.na, S, ctl, ws.n_sms, rs);
            GB_LAUNCH((replay_par_kernel<1>), ws.n_sms * 2, 512, 0, rs, R, ws.na, ctl, S.mode, ws.replay_cap);
        } else if (D == 2) {
            launch_stream<2>(R, ws.na, S, ctl, ws.n_sms, rs);
            GB_LAUNCH((replay_par_kernel<2>), ws.n_sms * 2, 512, 0, rs, R, ws.na, ctl, S.mode, ws.replay_cap);
        } else if (D == 3) {
            launch_stream<3>(R, ws.na, S, ctl, ws.n_sms, rs);
            GB_LAUNCH((replay_par_kernel<3>), ws.n_sms * 2, 512, 0, rs, R, ws.na, ctl, S.mode, ws.replay_cap);
        } else {
            launch_stream<4>(R, ws.na, S, ctl, ws.n_sms, rs);
            GB_LAUNCH((replay_par_kernel<4>), ws.n_sms * 2, 512, 0, rs, R, ws.na, ctl, S.mode, ws.replay_cap);
        }
    } else {
        // wide outputs: one lane per output dimension runs its chain sequentially (parallel over D instead of over rows)
        const int T = 128;
        const size_t smem = ((size_t)2 * T * D + 2 * D) * sizeof(float) + (size_t)2 * (T / 32) * sizeof(unsigned int);
        if (smem > 48 * 1024) ensure_dyn_smem(replay_kernel<1, 128>, smem);
        GB_LAUNCH((replay_kernel<1, 128>), ws.n_sms * 4, 128, smem, rs, R, ws.na);
    }
}

static DecideParams decide_params(Model &m, int level) {
    Workspace &ws = m.ws;
    DecideParams P;
    P.level = level; P.nT = ws.nT; P.F = ws.F; P.B = ws.B; P.D = ws.D; P.C = ws.F * ws.B; P.max_depth = m.cfg.max_depth;
    P.nn = 1 << level;
    P.hist_cur = ws.hist_p[level & 1]; P.thr = ws.thr.as<float>(); P.fw = m.feature_weights.as<float>();
    P.rev_map = m.rev_num_map.as<int>(); P.replay = ws.replay.as<ReplayItem>(); P.replay_scores = ws.replay_scores.as<float>();
    P.replay_exact = ws.replay_scores.as<float>() + ws.replay_cap;
    P.obl_cands = reinterpret_cast<int *>(ws.replay.as<ReplayItem>() + ws.replay_cap);
    P.ctl = ws.ctl.as<Ctl>(); P.ctl_stats = ws.ctl.as<Ctl>();
    P.state_snap = ws.state_snap.as<int>();
    P.use_replay = 1; P.count_stats = ws.count_stats ? 1 : 0; P.force_flip = 0;
    P.abort_flag = nullptr;
    return P;
}

// use_replay == false: speculative decision (the slot of the level is bound into the workspace)
void launch_decide(Model &m, int level, cudaStream_t s, bool use_replay) {
    static const bool force_flip = getenv("GBRL_B200_SPEC_FORCE_FLIP") != nullptr && getenv("GBRL_B200_SPEC_FORCE_FLIP")[0] == '1';
    DecideParams P = decide_params(m, level);
    P.use_replay = use_replay ? 1 : 0;
    P.force_flip = (!use_replay && force_flip) ? 1 : 0;
    if (!use_replay) P.abort_flag = m.ws.spec_flag.as<unsigned int>();
    GB_LAUNCH(decide_plan_kernel, 1, 1024, 0, s, P, m.ws.na, plan_params(m), m.cfg.grow_policy == GBRL_B200_GROW_OBLIVIOUS ? 1 : 0);
}

// after launch_decide of a speculative level (slot bound): the side stream waits for the decision, then compares
void launch_verify(Model &m, int level, cudaStream_t s, ReplaySlot &slot) {
    DecideParams P = decide_params(m, level);
    P.ctl = slot.ctl_snap.as<Ctl>();
    GB_CUDA(cudaEventRecord(slot.ev_dec, s));
    GB_CUDA(cudaStreamWaitEvent(slot.stream, slot.ev_dec, 0));
    GB_LAUNCH(verify_kernel, 1, 1024, 0, slot.stream, P, m.ws.na, m.cfg.grow_policy == GBRL_B200_GROW_OBLIVIOUS ? 1 : 0,
              m.ws.spec_flag.as<unsigned int>());
    GB_CUDA(cudaEventRecord(slot.ev_done, slot.stream));
}

void launch_rollback(Model &m, int level, cudaStream_t s) {
    Workspace &ws = m.ws;
    GB_LAUNCH(rollback_kernel, ceil_div(ws.MAXN, 256), 256, 0, s, ws.na, ws.state_snap.as<int>(), ws.MAXN, level,
              ws.spec_flag.as<unsigned int>());
}

}  // namespace gb
