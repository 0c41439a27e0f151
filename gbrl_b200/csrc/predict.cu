// predict.cu -- batched ensemble predict.
//
// Reference semantics restated:
//   predictor.cpp:122-185  preds = bias, then for every tree (ascending) the matching leaf's value is handed to each
//                          optimizer: theta[i] -= lr(t) * value[i] for i in [start_idx, stop_idx)  (optimizer.cpp:110-118)
//   predictor.cpp:231-265  oblivious: leaf_idx |= (x[f_k] > thr_k) << (depth-1-k)
//   predictor.cpp:188-229  greedy: first leaf of the tree whose path conditions all hold.  A tree is a binary
//                          partition, so this equals walking the tree from the root; we walk the per-tree heap
//                          topology stored next to the reference-layout arrays (O(depth) instead of O(leaves*depth)).
//                          A depth-0 tree never matches in the reference (passed=false) -> contributes nothing.
//   scheduler.h:124-135,182-185  Linear / Const learning rate
// Per sample the trees are applied sequentially in ascending order with mul-then-sub (no FMA), i.e. the
// float result equals the reference's sample-parallel / serial mode bit for bit.
#include "engine.cuh"

namespace gb {

struct DevOpt { int sched, start_idx, stop_idx, T; float init_lr, stop_lr; };

__device__ __forceinline__ float sched_lr(const DevOpt &o, int t) {
    if (o.sched == GBRL_B200_SCHED_CONST) return o.init_lr;
    const float T_ = (float)o.T;
    const float t_ = (float)t + 1;
    const float progress_remaining = (T_ - t_) / T_;
    const float lr = o.init_lr + (1.0f - progress_remaining) * (o.stop_lr - o.init_lr);
    return lr < o.stop_lr ? o.stop_lr : lr;
}

struct PredictParams {
    const float *X; float *preds; const float *bias;
    const int *tree_indices, *depths, *feature_indices, *heap_feat, *heap_leaf;
    const float *values, *feature_values, *heap_thr;
    const DevOpt *opts;
    int n_opts, N, F, D, md, start_tree, stop_tree, add_bias, oblivious;
};

template <int DM>
__global__ void __launch_bounds__(128) predict_kernel(PredictParams P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.N) return;
    const float *x = P.X + (size_t)i * P.F;
    float theta[DM];
#pragma unroll
    for (int d = 0; d < DM; ++d) theta[d] = (d < P.D) ? (P.add_bias ? P.bias[d] : P.preds[(size_t)i * P.D + d]) : 0.0f;
    const int md = P.md;
    for (int t = P.start_tree; t < P.stop_tree; ++t) {
        int leaf = -1;
        if (P.oblivious) {
            const int dep = P.depths[t];
            int li = 0;
            for (int k = 0; k < dep; ++k) {
                const int f = P.feature_indices[(size_t)t * md + k];
                const float thr = P.feature_values[(size_t)t * md + k];
                li |= (x[f] > thr ? 1 : 0) << (dep - 1 - k);
            }
            leaf = P.tree_indices[t] + li;
        } else {
            const int *hf = P.heap_feat + (size_t)t * (1 << md);
            const float *ht = P.heap_thr + (size_t)t * (1 << md);
            int h = 0, f = hf[0];
            if (f < 0) continue;                 // depth-0 tree: never matches in the reference
            while (f >= 0) {
                h = 2 * h + 1 + (x[f] > ht[h] ? 1 : 0);
                f = (h < (1 << md) - 1) ? hf[h] : -1;
            }
            leaf = P.tree_indices[t] + P.heap_leaf[(size_t)t * (2 << md) + h];
        }
        const float *v = P.values + (size_t)leaf * P.D;
        for (int o = 0; o < P.n_opts; ++o) {
            const DevOpt op = P.opts[o];
            const float lr = sched_lr(op, t);
#pragma unroll
            for (int d = 0; d < DM; ++d)
                if (d >= op.start_idx && d < op.stop_idx) theta[d] = theta[d] - lr * v[d];
        }
    }
#pragma unroll
    for (int d = 0; d < DM; ++d)
        if (d < P.D) P.preds[(size_t)i * P.D + d] = theta[d];
}

// ---------------------------------------------------------------- rollout-shape predict (BASELINE config 4)
// 100k trees x 8192 observations: one thread per observation cannot fill 148 SMs, and cutting the tree range into
// independently summed chunks changes the summation order -- at 100k trees the reference's own sequential fp32 rounding is
// already 2e-5, so only the reference's ORDER gives predictions within 1e-5 of it.  What is sequential in
// theta <- fl(theta - fl(lr * v_t)) (optimizer.cpp:110-118) is one FADD per (tree, output) -- 4 cycles; everything else (the
// walk, the value gather, lr * v) is order-free.  So a CTA owns 64 observations and splits its warps:
//   producer warps   walk the trees of the current chunk for a 32-observation sub-tile (lane = observation, features
//                    transposed in shared memory: lane == bank) and store x = lr * v[leaf] into a shared element ring;
//   consumer warps   (one per sub-tile) apply the PREVIOUS chunk's elements in tree order: theta = theta - x, bit for bit the
//                    reference's sample-parallel / serial mode (predictor.cpp:167-178).
// The split parameters and leaf values of a chunk are contiguous in the reference's SoA layout (types.h:279-304), so one
// elected thread brings them into shared memory with TMA bulk copies (cp.async.bulk + mbarrier complete_tx), double
// buffered one chunk ahead: the ensemble is read from L2 once per 64 observations instead of once per warp and tree.
constexpr int PT_SAMPLES = 64, PT_SUB = 2, PT_WARPS = 32, PT_CONS = PT_SUB, PT_PROD = PT_WARPS - PT_CONS;
constexpr int PT_XS = PT_SAMPLES + 1;          // row stride of the transposed observation tile (bank == lane)

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tPT_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra PT_DONE;\n\tbra PT_WAIT;\n\tPT_DONE:\n\t}"
                 ::"r"(bar), "r"(parity) : "memory");
}

struct TileLayout {          // byte offsets inside the dynamic shared memory of predict_tiles_kernel
    int xT, stage[2], elems[2], bars;
    int fi, fv, dep, ti, val;            // inside a stage
    int val_bytes, stage_bytes, total;
};
static TileLayout tile_layout(int F, int md, int DM, int TC, int val_bytes) {
    auto up = [](int v) { return (v + 127) & ~127; };
    TileLayout L;
    int o = 0;
    L.xT = o; o += up(F * PT_XS * 4);
    L.fi = 0; int so = up(TC * md * 4);
    L.fv = so; so += up(TC * md * 4);
    L.dep = so; so += up(TC * 4);
    L.ti = so; so += up((TC + 4) * 4);
    L.val = so; so += up(val_bytes + 16);
    L.val_bytes = val_bytes; L.stage_bytes = so;
    L.stage[0] = o; o += so; L.stage[1] = o; o += so;
    const int eb = up(TC * PT_SUB * DM * 32 * 4);
    L.elems[0] = o; o += eb; L.elems[1] = o; o += eb;
    L.bars = o; o += 128;
    L.total = o;
    return L;
}

template <int DM>
__global__ void __launch_bounds__(PT_WARPS * 32, 1)
predict_tiles_kernel(PredictParams P, TileLayout L, int TC, int n_trees_total, int n_leaves_total, long long val_capacity_floats) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ int s_opt_of_dim[DM];
    __shared__ float s_lr[DM];          // per output: the learning rate of its optimizer when every scheduler is constant
    __shared__ int s_all_const;
    __shared__ int s_voff[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s0 = blockIdx.x * PT_SAMPLES;
    const int md = P.md, D = P.D;
    float *xT = reinterpret_cast<float *>(smem + L.xT);
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t bar[2] = {sbase + (uint32_t)L.bars, sbase + (uint32_t)L.bars + 8u};
    // chunks are aligned to absolute multiples of TC so that every bulk copy starts on a 16-byte boundary
    const int c_first = P.start_tree / TC, c_last = (P.stop_tree - 1) / TC;      // stop_tree > start_tree (checked by the launcher)
    const int n_chunks = c_last - c_first + 1;

    if (threadIdx.x == 0) { mbar_init(bar[0], 1); mbar_init(bar[1], 1); }
    if (threadIdx.x < DM) {
        int o_of = -1;
        for (int o = 0; o < P.n_opts; ++o)
            if ((int)threadIdx.x >= P.opts[o].start_idx && (int)threadIdx.x < P.opts[o].stop_idx) o_of = o;
        s_opt_of_dim[threadIdx.x] = o_of;
        s_lr[threadIdx.x] = o_of >= 0 ? P.opts[o_of].init_lr : 0.0f;       // an output without optimizer is never updated: x = 0
    }
    if (threadIdx.x == 0) {
        int ac = 1;
        for (int o = 0; o < P.n_opts; ++o) ac &= (P.opts[o].sched == GBRL_B200_SCHED_CONST);
        s_all_const = ac;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // observations, transposed: xT[f][sample]
    for (int idx = threadIdx.x; idx < P.F * PT_SAMPLES; idx += PT_WARPS * 32) {
        const int sm = idx / P.F, f = idx - sm * P.F;
        xT[f * PT_XS + sm] = (s0 + sm < P.N) ? P.X[(size_t)(s0 + sm) * P.F + f] : 0.0f;
    }
    __syncthreads();

    // one elected thread: TMA bulk copies of a chunk's split parameters and leaf values into stage `b`
    auto issue_chunk = [&](int c, int b) {
        const int t0 = (c_first + c) * TC;
        const int tcnt = min(TC, n_trees_total - t0);
        const uint32_t st = sbase + (uint32_t)L.stage[b];
        const int l0 = P.tree_indices[t0];
        const int l1 = (t0 + tcnt < n_trees_total) ? P.tree_indices[t0 + tcnt] : n_leaves_total;
        const long long v0 = (long long)l0 * D, v1 = (long long)l1 * D;
        const long long a0 = v0 & ~3ll;                                   // align the start down to 16 bytes
        long long a1 = (v1 + 3) & ~3ll;
        if (a1 > val_capacity_floats) a1 = val_capacity_floats & ~3ll;    // never read past the allocation
        s_voff[b] = (int)(v0 - a0);
        const uint32_t b_fi = (uint32_t)(((tcnt * md * 4) + 15) & ~15), b_dep = (uint32_t)(((tcnt * 4) + 15) & ~15);
        const uint32_t b_ti = (uint32_t)((((tcnt + 1) * 4) + 15) & ~15), b_val = (uint32_t)((a1 - a0) * 4);
        mbar_expect_tx(bar[b], 2 * b_fi + b_dep + b_ti + b_val);
        tma_bulk_g2s(st + L.fi, P.feature_indices + (size_t)t0 * md, b_fi, bar[b]);
        tma_bulk_g2s(st + L.fv, P.feature_values + (size_t)t0 * md, b_fi, bar[b]);
        tma_bulk_g2s(st + L.dep, P.depths + t0, b_dep, bar[b]);
        tma_bulk_g2s(st + L.ti, P.tree_indices + t0, b_ti, bar[b]);
        if (b_val) tma_bulk_g2s(st + L.val, P.values + a0, b_val, bar[b]);
    };
    if (threadIdx.x == 0) issue_chunk(0, 0);

    float theta[DM];
    const int csub = warp;                                   // consumer warps 0 .. PT_CONS-1 own sub-tile `warp`
    if (warp < PT_CONS) {
        const int i = s0 + csub * 32 + lane;
#pragma unroll
        for (int d = 0; d < DM; ++d) theta[d] = (d < D && i < P.N) ? (P.add_bias ? P.bias[d] : P.preds[(size_t)i * D + d]) : 0.0f;
    }
    uint32_t phase[2] = {0u, 0u};
    const bool all_const = s_all_const != 0;          // published by the __syncthreads() above
#pragma unroll 1
    for (int c = 0; c <= n_chunks; ++c) {
        const int b = c & 1;
        if (c < n_chunks) {
            if (threadIdx.x == 0 && c + 1 < n_chunks) issue_chunk(c + 1, b ^ 1);      // stage b^1 was last read in iteration c-1
            mbar_wait(bar[b], phase[b]);
            phase[b] ^= 1u;
        }
        if (warp >= PT_CONS) {
            if (c < n_chunks) {
                // ---- producers: elements of chunk c
                const int t0 = (c_first + c) * TC;
                const int tcnt = min(TC, n_trees_total - t0);
                const unsigned char *st = smem + L.stage[b];
                const int *fi = reinterpret_cast<const int *>(st + L.fi);
                const float *fv = reinterpret_cast<const float *>(st + L.fv);
                const int *dep = reinterpret_cast<const int *>(st + L.dep);
                const int *ti = reinterpret_cast<const int *>(st + L.ti);
                const float *val = reinterpret_cast<const float *>(st + L.val) + s_voff[b];
                float *E = reinterpret_cast<float *>(smem + L.elems[b]);
                const int lbase = ti[0];
                // a producer warp takes whole trees: the split parameters are read once for both sub-tiles
                for (int j = warp - PT_CONS; j < tcnt; j += PT_PROD) {
                    const int t = t0 + j;
                    if (t < P.start_tree || t >= P.stop_tree) continue;
                    const int dj = dep[j];
                    int li0 = 0, li1 = 0;
                    if ((md & 1) == 0) {
                        // even max_depth: the tree's (feature, threshold) rows are 8-byte aligned -> two levels per broadcast LDS.64
                        const int2 *fi2 = reinterpret_cast<const int2 *>(fi + j * md);
                        const float2 *fv2 = reinterpret_cast<const float2 *>(fv + j * md);
                        for (int k = 0; k < dj; k += 2) {                                    // predictor.cpp:248-252, most significant bit first
                            const int2 f2 = fi2[k >> 1];
                            const float2 t2 = fv2[k >> 1];
                            const float *xa = xT + f2.x * PT_XS + lane;
                            li0 = (li0 << 1) | (xa[0] > t2.x ? 1 : 0);
                            li1 = (li1 << 1) | (xa[32] > t2.x ? 1 : 0);
                            if (k + 1 < dj) {
                                const float *xb = xT + f2.y * PT_XS + lane;
                                li0 = (li0 << 1) | (xb[0] > t2.y ? 1 : 0);
                                li1 = (li1 << 1) | (xb[32] > t2.y ? 1 : 0);
                            }
                        }
                    } else {
                        for (int k = 0; k < dj; ++k) {
                            const float *xr = xT + fi[j * md + k] * PT_XS + lane;
                            const float th = fv[j * md + k];
                            li0 = (li0 << 1) | (xr[0] > th ? 1 : 0);
                            li1 = (li1 << 1) | (xr[32] > th ? 1 : 0);
                        }
                    }
                    const float *vb = val + (size_t)(ti[j] - lbase) * D;
                    float *e = E + ((size_t)(j * PT_SUB) * DM) * 32 + lane;
                    if (DM == 2 && D == 2 && all_const && (reinterpret_cast<uintptr_t>(vb) & 7) == 0) {
                        // both outputs of a leaf with one 8-byte gather per sub-tile
                        const float2 v0 = *reinterpret_cast<const float2 *>(vb + (size_t)li0 * 2);
                        const float2 v1 = *reinterpret_cast<const float2 *>(vb + (size_t)li1 * 2);
                        e[0] = s_lr[0] * v0.x; e[32] = s_lr[1] * v0.y;
                        e[2 * 32] = s_lr[0] * v1.x; e[3 * 32] = s_lr[1] * v1.y;
                        continue;
                    }
#pragma unroll
                    for (int d = 0; d < DM; ++d) {
                        float x0 = 0.0f, x1 = 0.0f;
                        if (d < D) {
                            float lr = s_lr[d];                                              // optimizer.cpp:110-118: lr * value ...
                            if (!all_const) { const int o = s_opt_of_dim[d]; lr = o >= 0 ? sched_lr(P.opts[o], t) : 0.0f; }
                            x0 = lr * vb[(size_t)li0 * D + d];
                            x1 = lr * vb[(size_t)li1 * D + d];
                        }
                        e[d * 32] = x0;
                        e[(DM + d) * 32] = x1;
                    }
                }
            }
        } else if (c > 0) {
            // ---- consumers: chunk c-1, in tree order:  theta = theta - lr * value
            const int t0 = (c_first + c - 1) * TC;
            const int tcnt = min(TC, n_trees_total - t0);
            const float *E = reinterpret_cast<const float *>(smem + L.elems[b ^ 1]);
            const int j0 = max(0, P.start_tree - t0), j1 = min(tcnt, P.stop_tree - t0);
            for (int j = j0; j < j1; ++j) {
                const float *e = E + ((size_t)(j * PT_SUB + csub) * DM) * 32 + lane;
#pragma unroll
                for (int d = 0; d < DM; ++d) theta[d] = theta[d] - e[d * 32];
            }
        }
        __syncthreads();
    }
    if (warp < PT_CONS) {
        const int i = s0 + csub * 32 + lane;
        if (i < P.N) {
#pragma unroll
            for (int d = 0; d < DM; ++d)
                if (d < D) P.preds[(size_t)i * D + d] = theta[d];
        }
    }
}

template <int DM>
static bool launch_predict_tiles(Model &m, const PredictParams &P, cudaStream_t s) {
    const Ensemble &e = m.ens;
    const int md = P.md > 0 ? P.md : 1;
    // trees per chunk: the leaf values of a chunk (<= TC * 2^md * D floats) must fit a 32 KB stage
    int TC = 60;                                      // 2 trees per producer warp and chunk; a multiple of 4 (16-byte aligned bulk copies)
    while (TC > 4 && (long long)TC * (1ll << md) * P.D * 4 > 32768) TC = (TC / 2) & ~3;
    if ((long long)TC * (1ll << md) * P.D * 4 > 32768) return false;
    const int val_bytes = TC * (1 << md) * P.D * 4 + 32;
    const TileLayout L = tile_layout(P.F, md, DM, TC, val_bytes);
    if (L.total > 220 * 1024) return false;
    ensure_dyn_smem(predict_tiles_kernel<DM>, (size_t)L.total);
    GB_LAUNCH(predict_tiles_kernel<DM>, ceil_div(P.N, PT_SAMPLES), PT_WARPS * 32, (size_t)L.total, s, P, L, TC, e.n_trees, e.n_leaves,
              (long long)(e.values.bytes / sizeof(float)));
    return true;
}

void upload_optimizers(Model &m, cudaStream_t s) {
    std::vector<DevOpt> h(m.opts.size());
    for (size_t i = 0; i < m.opts.size(); ++i) {
        h[i].sched = m.opts[i].sched; h[i].start_idx = m.opts[i].start_idx; h[i].stop_idx = m.opts[i].stop_idx;
        h[i].T = m.opts[i].T; h[i].init_lr = m.opts[i].init_lr; h[i].stop_lr = m.opts[i].stop_lr;
    }
    m.d_opts.ensure((h.size() > 0 ? h.size() : 1) * sizeof(DevOpt));
    if (!h.empty()) {
        GB_CUDA(cudaMemcpyAsync(m.d_opts.p, h.data(), h.size() * sizeof(DevOpt), cudaMemcpyHostToDevice, s));
        GB_CUDA(cudaStreamSynchronize(s));
    }
}

template <int DM>
static void launch_predict_dm(const PredictParams &P, cudaStream_t s) {
    GB_LAUNCH(predict_kernel<DM>, ceil_div(P.N, 128), 128, 0, s, P);
}

void launch_predict(Model &m, const float *X, int N, int F, int start_tree, int stop_tree, float *preds, bool add_bias,
                    cudaStream_t s) {
    if (N <= 0) return;
    Ensemble &e = m.ens;
    PredictParams P;
    P.X = X; P.preds = preds; P.bias = m.bias.as<float>();
    P.tree_indices = e.tree_indices.as<int>(); P.depths = e.depths.as<int>(); P.feature_indices = e.feature_indices.as<int>();
    P.heap_feat = e.heap_feat.as<int>(); P.heap_leaf = e.heap_leaf.as<int>(); P.values = e.values.as<float>();
    P.feature_values = e.feature_values.as<float>(); P.heap_thr = e.heap_thr.as<float>();
    P.opts = m.d_opts.as<DevOpt>(); P.n_opts = (int)m.opts.size();
    P.N = N; P.F = F; P.D = m.cfg.output_dim; P.md = m.cfg.max_depth; P.start_tree = start_tree; P.stop_tree = stop_tree;
    P.add_bias = add_bias ? 1 : 0; P.oblivious = m.cfg.grow_policy == GBRL_B200_GROW_OBLIVIOUS;
    const int D = P.D;
    // rollout shape (many trees): parallel walks + ordered accumulation (predict_tiles_kernel); oblivious trees, disjoint
    // optimizer ranges (every output belongs to at most one optimizer: the learners' layout, actor_critic_learner.py:88)
    const int n_t = stop_tree - start_tree;
    bool disjoint = true;
    for (size_t a = 0; a < m.opts.size(); ++a)
        for (size_t b2 = a + 1; b2 < m.opts.size(); ++b2)
            if (m.opts[a].start_idx < m.opts[b2].stop_idx && m.opts[b2].start_idx < m.opts[a].stop_idx) disjoint = false;
    if (P.oblivious && n_t >= 256 && D <= 4 && disjoint && e.n_trees > 0) {
        bool done = false;
        if (D <= 1) done = launch_predict_tiles<1>(m, P, s);
        else if (D <= 2) done = launch_predict_tiles<2>(m, P, s);
        else done = launch_predict_tiles<4>(m, P, s);
        if (done) return;
    }
    if (D <= 1) launch_predict_dm<1>(P, s);
    else if (D <= 2) launch_predict_dm<2>(P, s);
    else if (D <= 4) launch_predict_dm<4>(P, s);
    else if (D <= 8) launch_predict_dm<8>(P, s);
    else if (D <= 16) launch_predict_dm<16>(P, s);
    else if (D <= 32) launch_predict_dm<32>(P, s);
    else launch_predict_dm<64>(P, s);
}

// the newest tree only, applied on top of existing predictions (fit loop)
void launch_update_preds_last_tree(Model &m, const float *X, int N, int F, float *preds, cudaStream_t s) {
    launch_predict(m, X, N, F, m.ens.n_trees - 1, m.ens.n_trees, preds, false, s);
}

// Same update for the rows the newest tree was just grown on: every row still carries the heap id of the node it ended
// in (`nid`, the same x > threshold comparisons the walk would repeat) and tree.cu numbered that node's leaf, so the
// tree does not have to be walked again: theta -= lr * value[leaf]  (optimizer.cpp:110-118), one coalesced pass.
__global__ void __launch_bounds__(256)
update_preds_from_nodes_kernel(float *__restrict__ preds, const int *__restrict__ nid, const int *__restrict__ leaf_index,
                               const int *__restrict__ tree_indices, const float *__restrict__ values, const int *__restrict__ heap_feat_t,
                               const DevOpt *__restrict__ opts, int n_opts, int N, int D, int t, int oblivious) {
    if (!oblivious && heap_feat_t[0] < 0) return;          // depth-0 tree: never matches in the reference (predictor.cpp:211-217)
    const int first_leaf = tree_indices[t];
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)N * D; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / D), d = (int)(e - (long long)i * D);
        const int li = leaf_index[nid[i]];
        if (li < 0) continue;
        const float v = values[(size_t)(first_leaf + li) * D + d];
        float th = preds[e];
        for (int o = 0; o < n_opts; ++o) {
            const DevOpt op = opts[o];
            if (d >= op.start_idx && d < op.stop_idx) th = th - sched_lr(op, t) * v;
        }
        preds[e] = th;
    }
}

void launch_update_preds_from_nodes(Model &m, int N, float *preds, cudaStream_t s) {
    if (N <= 0) return;
    Ensemble &e = m.ens;
    const int t = e.n_trees - 1, D = m.cfg.output_dim, md = m.cfg.max_depth;
    const bool obl = m.cfg.grow_policy == GBRL_B200_GROW_OBLIVIOUS;
    long long ne = (long long)N * D;
    int grid = (int)((ne + 1023) / 1024);
    if (grid > 148 * 16) grid = 148 * 16;
    GB_LAUNCH(update_preds_from_nodes_kernel, grid, 256, 0, s, preds, m.ws.nid.as<int>(), m.ws.na.leaf_index, e.tree_indices.as<int>(),
              e.values.as<float>(), obl ? nullptr : e.heap_feat.as<int>() + (size_t)t * (1 << md), m.d_opts.as<DevOpt>(),
              (int)m.opts.size(), N, D, t, obl ? 1 : 0);
}

// ---------------------------------------------------------------- rebuild heap topology from leaf paths
// (ensembles loaded in the reference layout: gbrl_b200_set_ensemble)
__global__ void rebuild_heap_kernel(const int *tree_indices, const int *depths, const int *feature_indices,
                                    const float *feature_values, const uint8_t *ineq, int *heap_feat, float *heap_thr,
                                    int *heap_leaf, int n_trees, int n_leaves, int md) {
    const int t = blockIdx.x;
    if (t >= n_trees) return;
    const int l0 = tree_indices[t], l1 = (t + 1 < n_trees) ? tree_indices[t + 1] : n_leaves;
    for (int h = threadIdx.x; h < (1 << md); h += blockDim.x) { heap_feat[(size_t)t * (1 << md) + h] = -1; heap_thr[(size_t)t * (1 << md) + h] = 0.0f; }
    for (int h = threadIdx.x; h < (2 << md); h += blockDim.x) heap_leaf[(size_t)t * (2 << md) + h] = -1;
    __syncthreads();
    for (int leaf = l0 + threadIdx.x; leaf < l1; leaf += blockDim.x) {
        const int dep = depths[leaf];
        int h = 0;
        for (int k = 0; k < dep; ++k) {
            heap_feat[(size_t)t * (1 << md) + h] = feature_indices[(size_t)leaf * md + k];
            heap_thr[(size_t)t * (1 << md) + h] = feature_values[(size_t)leaf * md + k];
            h = 2 * h + 1 + (ineq[(size_t)leaf * md + k] ? 1 : 0);
        }
        heap_leaf[(size_t)t * (2 << md) + h] = leaf - l0;
    }
}

void rebuild_heap_topology(Model &m, cudaStream_t s) {
    Ensemble &e = m.ens;
    if (m.cfg.grow_policy == GBRL_B200_GROW_OBLIVIOUS || e.n_trees == 0) return;
    GB_LAUNCH(rebuild_heap_kernel, e.n_trees, 64, 0, s, e.tree_indices.as<int>(), e.depths.as<int>(), e.feature_indices.as<int>(),
              e.feature_values.as<float>(), e.ineq.as<uint8_t>(), e.heap_feat.as<int>(), e.heap_thr.as<float>(),
              e.heap_leaf.as<int>(), e.n_trees, e.n_leaves, m.cfg.max_depth);
}

}  // namespace gb
