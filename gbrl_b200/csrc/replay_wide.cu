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

namespace gb {

constexpr int GROUP_ROWS = 256;
constexpr int WIN_MIN_GROUPS = 64;     // chains of at least this many groups are walked through window tables first
constexpr float TAG_EMPTY = -1.0f;     // tag of a group in which the chain has no element: applicable to any running sum

__device__ __forceinline__ double shfl_xor_d(double v, int m) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, m); hi = __shfl_xor_sync(0xffffffffu, hi, m);
    return __hiloint2double(hi, lo);
}

// A warp owns a contiguous range of 8-word groups (256 rows each) of the planes: the item of the first group is
// found by one binary search, after that the item index only moves forward.
template <int D>
__device__ __forceinline__ void wide_bits_body(const ReplayParams &P, NodeArrays na, const StreamParams &S, const WideParams &Wd) {
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
__device__ __forceinline__ void wide_prefix_body(const ReplayParams &P, NodeArrays na, const StreamParams &S, const WideParams &Wd) {
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

// the lane's 8 rows of a group: values (row-major, D per row), side bits, members of chain side `side`
template <int D>
__device__ __forceinline__ void group_rows(const float *G, const unsigned int *W, int n, int lg, float (&v)[8 * D], unsigned int &mb) {
    const int lane = threadIdx.x & 31;
    const int kb = lg * GROUP_ROWS + lane * 8;
    const float *p = G + (size_t)kb * D;
    if ((lg + 1) * GROUP_ROWS <= n) {
        // interior group: one address, 8 * D loads (LDG.128 when the node's first row allows it)
        if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 8 * D; j += 4) {
                const float4 a = *reinterpret_cast<const float4 *>(p + j);
                v[j] = a.x; v[j + 1] = a.y; v[j + 2] = a.z; v[j + 3] = a.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8 * D; ++j) v[j] = p[j];
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8 * D; ++j) {
            const int k = kb + j / D;
            v[j] = k < n ? p[j] : 0.0f;
        }
    }
    mb = (W[lg * 8 + (lane >> 2)] >> ((lane & 3) * 8)) & 0xffu;
}

template <int D, int PASS>
__device__ __forceinline__ void chain_elems(const float (&v)[8 * D], unsigned int mb, int c, const float *smean,
                                            float (&x)[PASS == 0 ? 8 : 8 * D]) {
    if (PASS == 0) {
        const int side = c / D, d = c - side * D;
        const unsigned int sel = side ? mb : ~mb;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float val = v[r * D];
#pragma unroll
            for (int dd = 1; dd < D; ++dd) val = (d == dd) ? v[r * D + dd] : val;
            x[r] = ((sel >> r) & 1u) ? val : 0.0f;
        }
    } else {
        const unsigned int sel = c ? mb : ~mb;
#pragma unroll
        for (int j = 0; j < 8 * D; ++j) {
            const int r = j / D, d = j - r * D;
            x[j < (PASS == 0 ? 8 : 8 * D) ? j : 0] = ((sel >> r) & 1u) ? v[j] * smean[c * D + d] : 0.0f;   // math_ops.h:432-449
        }
    }
}

// summaries of all chains of group g for the predicted binade (one warp)
template <int D, int PASS>
__device__ __forceinline__ void tabs_group(const ReplayParams &P, NodeArrays na, const StreamParams &S, const WideParams &Wd, int g) {
    constexpr int NCH = PASS == 0 ? 2 * D : 2;
    constexpr int KE = PASS == 0 ? 8 : 8 * D;
    const int lane = threadIdx.x & 31;
    const int it = Wd.gitem[g];
    if (it < 0) return;
    const ReplayItem item = P.items[it];
    const int s0 = na.seg_start[item.node], n = na.seg_len[item.node];
    const int lg = g - (S.woff[it] >> 3);
    float v[8 * D];
    unsigned int mb;
    group_rows<D>(S.G + (size_t)s0 * D, S.bits + S.woff[it], n, lg, v, mb);
    float smean[2 * D];
    if (PASS == 1) {
#pragma unroll
        for (int i = 0; i < 2 * D; ++i) smean[i] = Wd.fin[(size_t)it * 8 + i];
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        float pr;
        if (PASS == 0) pr = Wd.pred[(size_t)g * 2 * D + c];
        else {
            pr = 0.0f;
#pragma unroll
            for (int d = 0; d < D; ++d) pr += smean[c * D + d] * Wd.pred[(size_t)g * 2 * D + c * D + d];
        }
        float inv_u, u;
        const bool ok = seq::epoch_of(pr, inv_u, u);
        float x[KE];
        chain_elems<D, PASS>(v, mb, c, smean, x);
        bool nz = false;
#pragma unroll
        for (int i = 0; i < KE; ++i) nz |= (x[i] != 0.0f) || (x[i] != x[i]);
        const bool empty = !__any_sync(0xffffffffu, nz);          // the chain has no (non-zero) element in this group
        float tagv = ok ? inv_u : 0.0f;
        if (empty) {
            tagv = TAG_EMPTY;
            if (lane == 0) Wd.tab[(size_t)g * 2 * D + c] = make_int4(0, 0, 0, 0);
        } else if (ok) {
            const seq::Tab tb = seq::warp_summarize<KE>(x, inv_u);
            if (lane == 0) Wd.tab[(size_t)g * 2 * D + c] = make_int4(tb.a0, tb.a1, tb.mn, tb.mx);
        }
        if (lane == 0) Wd.tag[(size_t)g * 2 * D + c] = tagv;
    }
}

// Second level of the summaries: one warp per (window of 32 groups, chain) composes the window's group tables into ONE table
// (tables compose associatively, chain.cuh) when all its groups were summarised for the same binade.  The chain warp
// then walks window tables 32 at a time (1024 groups = 262 144 rows per scan) and descends into a window only where the
// composite is not applicable to the actual running sum.  Items are laid out on 32-group boundaries (replay_plan_body).
template <int D, int PASS>
__device__ __forceinline__ void wtabs_window(const ReplayParams &P, NodeArrays na, const StreamParams &S, const WideParams &Wd, int gw, int c) {
    const unsigned int full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    {
        const int it = Wd.gitem[(size_t)gw * 32];
        if (it < 0) return;
        const int n = na.seg_len[P.items[it].node];
        const int ng = (n + GROUP_ROWS - 1) / GROUP_ROWS;
        if (ng < WIN_MIN_GROUPS) return;                                    // short chains are walked at group level only
        const int lw = gw - ((S.woff[it] >> 3) >> 5);
        const int cnt = min(32, ng - lw * 32);
        if (cnt <= 0) return;
        int4 q = make_int4(0, 0, 0, 0);
        float tg = TAG_EMPTY;
        if (lane < cnt) { q = Wd.tab[((size_t)gw * 32 + lane) * 2 * D + c]; tg = Wd.tag[((size_t)gw * 32 + lane) * 2 * D + c]; }
        // common binade of the non-empty groups (0: mixed / none -> the window has no composite)
        const unsigned int tb = __float_as_uint(tg);
        const bool nonempty = tg != TAG_EMPTY;
        const unsigned int ne_mask = __ballot_sync(full, nonempty);
        float wtag = TAG_EMPTY;
        int4 out = make_int4(0, 0, 0, 0);
        if (ne_mask) {
            const unsigned int t0 = __shfl_sync(full, tb, __ffs(ne_mask) - 1);
            const bool same = !nonempty || tb == t0;
            wtag = (__all_sync(full, same) && __uint_as_float(t0) != 0.0f) ? __uint_as_float(t0) : 0.0f;
            if (wtag != 0.0f) {
                int i0 = nonempty ? q.x : 0, i1 = nonempty ? q.y : 0;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int g0 = __shfl_up_sync(full, i0, off), g1 = __shfl_up_sync(full, i1, off);
                    if (lane >= off) {
                        const int n0 = g0 + ((g0 & 1) ? i1 : i0);
                        const int n1 = g1 + (((g1 + 1) & 1) ? i1 : i0);
                        i0 = n0; i1 = n1;
                    }
                }
                int e0 = __shfl_up_sync(full, i0, 1);
                if (lane == 0) e0 = 0;
                // min / max prefix along the even-parity path (the odd path differs by at most 1, as inside a group)
                long long mn = nonempty ? (long long)e0 + q.z : 0, mx = nonempty ? (long long)e0 + q.w : 0;
                mn = max(mn, (long long)-seq::BAD); mx = min(mx, (long long)seq::BAD);
                int mni = (int)mn, mxi = (int)mx;
                mni = __reduce_min_sync(full, mni); mxi = __reduce_max_sync(full, mxi);
                // increments beyond +-2^29 cannot be applied inside a binade anyway: refuse the composite instead of overflowing
                const bool small = abs(q.x) < (1 << 24) && abs(q.y) < (1 << 24);
                if (!__all_sync(full, small)) wtag = 0.0f;
                out = make_int4(__shfl_sync(full, i0, 31), __shfl_sync(full, i1, 31), mni - 1, mxi + 1);
            }
        }
        if (lane == 0) { Wd.wtab[(size_t)gw * 2 * D + c] = out; Wd.wtag[(size_t)gw * 2 * D + c] = wtag; }
    }
}

// composes the longest applicable prefix of a window of 32 groups (lane = group); returns how many were consumed
__device__ __forceinline__ int compose_window(float &acc, const int4 q, float tg, bool in_range, int first) {
    const unsigned int full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    float inv_u, u;
    const bool live = lane >= first;                      // lanes before `first` are consumed already
    if (!seq::epoch_of(acc, inv_u, u)) {
        // no binade yet (acc == 0): only groups without elements can be skipped
        const unsigned int stop = __ballot_sync(full, live && !(in_range && tg == TAG_EMPTY));
        return (stop ? (__ffs(stop) - 1) : 32) - first;
    }
    const bool empty = in_range && tg == TAG_EMPTY;
    const bool tag_ok = in_range && (tg == inv_u || empty);
    int i0 = (tag_ok && live) ? q.x : 0, i1 = (tag_ok && live) ? q.y : 0;
    if (!__any_sync(full, i0 != i1)) {
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int g0 = __shfl_up_sync(full, i0, off);
            if (lane >= off) i0 += g0;
        }
        i1 = i0;
    } else {
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int g0 = __shfl_up_sync(full, i0, off), g1 = __shfl_up_sync(full, i1, off);
            if (lane >= off) {
                const int n0 = g0 + ((g0 & 1) ? i1 : i0);
                const int n1 = g1 + (((g1 + 1) & 1) ? i1 : i0);
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
    const int take = (bad ? (__ffs(bad) - 1) : 32) - first;   // groups first .. first + take - 1 of the window are applied
    if (take <= 0) return 0;
    const int inc0 = __shfl_sync(full, i0, first + take - 1), inc1 = __shfl_sync(full, i1, first + take - 1);
    acc = (float)(m + ((m & 1) ? inc1 : inc0)) * u;
    return take;
}

// one CTA per item, warp c walks chain c; then the item's score (L2, or Cosine after PASS 1) / the side means (Cosine, PASS 0)
template <int D, int PASS>
__device__ __forceinline__ void wide_walk_body(const ReplayParams &P, NodeArrays na, const StreamParams &S, const WideParams &Wd, Ctl *ctl_stats) {
    constexpr int NCH = PASS == 0 ? 2 * D : 2;
    constexpr int KE = PASS == 0 ? 8 : 8 * D;
    __shared__ float s_sum[2 * D];
    __shared__ float s_mean[2 * D];
    __shared__ __align__(16) float s_wbuf[NCH][32 * KE];       // per chain warp: staging of a block that is run sequentially
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_items = min(P.ctl->n_replay, S.replay_cap);
    if (PASS == 1 && P.score_func == GBRL_B200_SCORE_L2) return;
    int n_fast = 0, n_slow = 0, n_seq = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        if (S.mode[it] != 0) continue;
        const ReplayItem item = P.items[it];
        const int h = item.node, cand = item.cand;
        const int s0 = na.seg_start[h], n = na.seg_len[h];
        const int ng = (n + GROUP_ROWS - 1) / GROUP_ROWS;
        const size_t base = (size_t)(S.woff[it] >> 3);
        const float *G = S.G + (size_t)s0 * D;
        const unsigned int *W = S.bits + S.woff[it];
        if (PASS == 1) {
            if (threadIdx.x < 2 * D) s_mean[threadIdx.x] = Wd.fin[(size_t)it * 8 + threadIdx.x];
            __syncthreads();
        }
        if (warp < NCH) {
            const int c = warp;
            float acc = 0.0f;
            // group-level walk of the windows [wa, wb) of this chain (window = 32 groups, lane = group)
            auto walk_groups = [&](int wa, int wb) {
                int4 qnx = make_int4(0, 0, 0, 0);          // the next window's summaries / tags, loaded a window ahead
                float tgnx = 0.0f;
                if (wa * 32 + lane < ng) { qnx = Wd.tab[(base + wa * 32 + lane) * 2 * D + c]; tgnx = Wd.tag[(base + wa * 32 + lane) * 2 * D + c]; }
#pragma unroll 1
                for (int w0 = wa * 32; w0 < ng && w0 < wb * 32; w0 += 32) {
                    const bool in_range = w0 + lane < ng;
                    const int wn = min(32, ng - w0);
                    const int4 q = qnx;
                    const float tg = tgnx;
                    if (w0 + 32 + lane < ng && w0 + 32 < wb * 32) { qnx = Wd.tab[(base + w0 + 32 + lane) * 2 * D + c]; tgnx = Wd.tag[(base + w0 + 32 + lane) * 2 * D + c]; }
                    int first = 0;
                    float vn[8 * D];                       // rows of the group after a failed one, fetched while that one is advanced
                    unsigned int mbn = 0u;
                    int have = -1;                         // group (window-relative) whose rows vn holds
#pragma unroll 1
                    while (first < wn) {
                        const int take = compose_window(acc, q, tg, in_range, first);
                        n_fast += take;
                        first += take;
                        if (first >= wn) break;
                        // this group is in another binade than predicted, or the sum leaves its binade inside it
                        float v[8 * D], x[KE];
                        unsigned int mb;
                        if (have == first) {
#pragma unroll
                            for (int j = 0; j < 8 * D; ++j) v[j] = vn[j];
                            mb = mbn;
                        } else group_rows<D>(G, W, n, w0 + first, v, mb);
                        if (w0 + first + 1 < ng) { group_rows<D>(G, W, n, w0 + first + 1, vn, mbn); have = first + 1; }
                        chain_elems<D, PASS>(v, mb, c, s_mean, x);
                        acc = seq::warp_seq_block<KE>(acc, x, s_wbuf[c], n_seq);
                        ++n_slow; ++first;
                    }
                }
            };
            const int nw = (ng + 31) >> 5;                 // windows of this chain
            if (ng < WIN_MIN_GROUPS) walk_groups(0, nw);
            else {
                // window tables first: 32 windows (1024 groups) per scan; a window whose composite does not apply is walked at group level
                const size_t wbase = base >> 5;
#pragma unroll 1
                for (int s0w = 0; s0w < nw; s0w += 32) {   // super-window, lane = window
                    const bool in_range = s0w + lane < nw;
                    const int sn = min(32, nw - s0w);
                    int4 q = make_int4(0, 0, 0, 0);
                    float tg = 0.0f;
                    if (in_range) { q = Wd.wtab[(wbase + s0w + lane) * 2 * D + c]; tg = Wd.wtag[(wbase + s0w + lane) * 2 * D + c]; }
                    int first = 0;
#pragma unroll 1
                    while (first < sn) {
                        const int take = compose_window(acc, q, tg, in_range, first);
                        n_fast += take * 32;
                        first += take;
                        if (first >= sn) break;
                        walk_groups(s0w + first, s0w + first + 1);
                        ++first;
                    }
                }
            }
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
    if (lane == 0 && (n_fast | n_slow)) {
        atomicAdd((unsigned long long *)&ctl_stats->stat_chain_fast, (unsigned long long)n_fast);
        atomicAdd((unsigned long long *)&ctl_stats->stat_chain_slow, (unsigned long long)n_slow);
        atomicAdd((unsigned long long *)&ctl_stats->stat_chain_seq, (unsigned long long)n_seq);
    }
}

// ---------------------------------------------------------------- kernels
// (One cooperative launch with grid-wide barriers between the stages was measured too: C2 select_replay 0.59 -> 0.96 ms --
//  a barrier over n_sms x occupancy CTAs costs more than the back-to-back launches it replaces.  profiles/README.md.)
template <int D>
__global__ void __launch_bounds__(256) wide_bits_kernel(ReplayParams P, NodeArrays na, StreamParams S, WideParams Wd) { wide_bits_body<D>(P, na, S, Wd); }
template <int D>
__global__ void __launch_bounds__(256) wide_prefix_kernel(ReplayParams P, NodeArrays na, StreamParams S, WideParams Wd) { wide_prefix_body<D>(P, na, S, Wd); }

// one CTA per window of 32 groups: 8 warps x 4 groups of summaries, then one warp per chain composes the window's table
template <int D, int PASS>
__global__ void __launch_bounds__(256) wide_tabs_kernel(ReplayParams P, NodeArrays na, StreamParams S, WideParams Wd) {
    constexpr int NCH = PASS == 0 ? 2 * D : 2;
    const int n_items = min(P.ctl->n_replay, S.replay_cap);
    if (n_items <= 0) return;
    if (PASS == 1 && P.score_func == GBRL_B200_SCORE_L2) return;
    const int warp = threadIdx.x >> 5;
    const int total_windows = (S.woff[n_items] >> 3) >> 5;      // planes are padded to whole windows (replay_plan_body)
    for (int gw = blockIdx.x; gw < total_windows; gw += gridDim.x) {
#pragma unroll 1
        for (int gi = warp; gi < 32; gi += 8) tabs_group<D, PASS>(P, na, S, Wd, gw * 32 + gi);
        __syncthreads();                                        // the window's group tables are visible to the CTA
        if (warp < NCH) wtabs_window<D, PASS>(P, na, S, Wd, gw, warp);
    }
}

template <int D, int PASS>
__global__ void __launch_bounds__(256) wide_walk_kernel(ReplayParams P, NodeArrays na, StreamParams S, WideParams Wd, Ctl *ctl_stats) {
    wide_walk_body<D, PASS>(P, na, S, Wd, ctl_stats);
}

template <int D>
static void launch_wide_d(Model &m, const ReplayParams &R, const StreamParams &S, const WideParams &Wd, cudaStream_t s) {
    Workspace &ws = m.ws;
    Ctl *ctl = ws.ctl.as<Ctl>();
    const int grid = ws.n_sms * replay_grid_mult(R.ctl != ws.ctl.as<Ctl>());      // side stream <=> the control block is a snapshot
    GB_LAUNCH((wide_bits_kernel<D>), grid, 256, 0, s, R, ws.na, S, Wd);
    GB_LAUNCH((wide_prefix_kernel<D>), ws.n_sms, 256, 0, s, R, ws.na, S, Wd);
    GB_LAUNCH((wide_tabs_kernel<D, 0>), grid, 256, 0, s, R, ws.na, S, Wd);
    GB_LAUNCH((wide_walk_kernel<D, 0>), ws.n_sms * 2, 256, 0, s, R, ws.na, S, Wd, ctl);
    if (m.cfg.split_score_func != GBRL_B200_SCORE_L2) {
        GB_LAUNCH((wide_tabs_kernel<D, 1>), grid, 256, 0, s, R, ws.na, S, Wd);
        GB_LAUNCH((wide_walk_kernel<D, 1>), ws.n_sms * 2, 256, 0, s, R, ws.na, S, Wd, ctl);
    }
}

void launch_replay_wide(Model &m, const ReplayParams &R, const StreamParams &S, cudaStream_t s) {
    Workspace &ws = m.ws;
    WideParams Wd;
    const size_t ng = (size_t)ws.rwide_groups, D2 = (size_t)2 * ws.D;
    char *p = ws.rwide.as<char>();
    Wd.bsum = reinterpret_cast<double *>(p); p += ng * D2 * sizeof(double);
    Wd.tab = reinterpret_cast<int4 *>(p); p += ng * D2 * sizeof(int4);
    Wd.pred = reinterpret_cast<float *>(p); p += ng * D2 * sizeof(float);
    Wd.tag = reinterpret_cast<float *>(p); p += ng * D2 * sizeof(float);
    Wd.gitem = reinterpret_cast<int *>(p); p += ng * sizeof(int);
    Wd.fin = reinterpret_cast<float *>(p); p += (size_t)ws.replay_cap * 8 * sizeof(float);
    p = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15);
    Wd.wtab = reinterpret_cast<int4 *>(p); p += (ng / 32 + 2) * D2 * sizeof(int4);
    Wd.wtag = reinterpret_cast<float *>(p);
    Wd.cap_groups = (long long)ng;
    if (ws.D == 1) launch_wide_d<1>(m, R, S, Wd, s);
    else launch_wide_d<2>(m, R, S, Wd, s);
}

}  // namespace gb
