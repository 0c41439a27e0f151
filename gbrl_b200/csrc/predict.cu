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

// ---------------------------------------------------------------- tree-chunked predict (rollout shape)
// BASELINE config 4 (100k trees x 8192 observations): 8192 samples cannot fill 148 SMs with one thread per sample, so
// the tree range is cut into chunks: a CTA owns 32 samples (lane = sample) x one chunk of trees, its 4 warps split
// the chunk.  The 32 observations are staged transposed in shared memory (xT[f][33]: lane == bank, conflict-free
// for the uniform feature index of a tree level); tree parameters are warp-uniform loads.  Partial sums per chunk
// are written out and combined in chunk order by a second kernel, so the result is deterministic (the reference's
// own tree-parallel mode also sums per-thread partial buffers, predictor.cpp:147-165); agreement with the
// sequential order is within float rounding (<= 1e-5 is tested).
template <int DM>
__global__ void __launch_bounds__(128) predict_chunk_kernel(PredictParams P, float *__restrict__ partials, int trees_per_chunk) {
    extern __shared__ float xT[];                      // [F][33]
    __shared__ float red[4][32][DM];
    const int tile = blockIdx.x, chunk = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s0 = tile * 32;
    for (int idx = threadIdx.x; idx < P.F * 32; idx += 128) {
        const int s = idx / P.F, f = idx % P.F;
        xT[f * 33 + s] = (s0 + s < P.N) ? P.X[(size_t)(s0 + s) * P.F + f] : 0.0f;
    }
    __syncthreads();
    const int c0 = P.start_tree + chunk * trees_per_chunk;
    const int c1 = min(P.stop_tree, c0 + trees_per_chunk);
    const int per_warp = (c1 - c0 + 3) / 4;
    const int t0 = c0 + warp * per_warp, t1 = min(c1, t0 + per_warp);
    float acc[DM];
#pragma unroll
    for (int d = 0; d < DM; ++d) acc[d] = 0.0f;
    const int md = P.md;
    for (int t = t0; t < t1; ++t) {
        int leaf;
        if (P.oblivious) {
            const int dep = P.depths[t];
            int li = 0;
            for (int k = 0; k < dep; ++k) {
                const int f = P.feature_indices[(size_t)t * md + k];
                const float thr = P.feature_values[(size_t)t * md + k];
                li |= (xT[f * 33 + lane] > thr ? 1 : 0) << (dep - 1 - k);
            }
            leaf = P.tree_indices[t] + li;
        } else {
            const int *hf = P.heap_feat + (size_t)t * (1 << md);
            const float *ht = P.heap_thr + (size_t)t * (1 << md);
            int h = 0, f = hf[0];
            if (f < 0) continue;
            while (f >= 0) {
                h = 2 * h + 1 + (xT[f * 33 + lane] > ht[h] ? 1 : 0);
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
                if (d >= op.start_idx && d < op.stop_idx) acc[d] = acc[d] - lr * v[d];
        }
    }
#pragma unroll
    for (int d = 0; d < DM; ++d) red[warp][lane][d] = acc[d];
    __syncthreads();
    if (warp == 0 && s0 + lane < P.N) {
#pragma unroll
        for (int d = 0; d < DM; ++d) {
            if (d < P.D) {
                float sum = red[0][lane][d];
                sum = sum + red[1][lane][d]; sum = sum + red[2][lane][d]; sum = sum + red[3][lane][d];
                partials[((size_t)chunk * P.N + s0 + lane) * P.D + d] = sum;
            }
        }
    }
}

__global__ void predict_combine_kernel(const float *__restrict__ partials, const float *__restrict__ bias, float *__restrict__ preds,
                                       int N, int D, int n_chunks, int add_bias) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * D) return;
    float acc = add_bias ? bias[i % D] : preds[i];
    for (int c = 0; c < n_chunks; ++c) acc = acc + partials[(size_t)c * N * D + i];
    preds[i] = acc;
}

template <int DM>
static void launch_predict_chunked(Model &m, const PredictParams &P, int n_chunks, int tpc, cudaStream_t s) {
    const size_t smem = (size_t)P.F * 33 * sizeof(float);
    if (smem > 48 * 1024) ensure_dyn_smem(predict_chunk_kernel<DM>, smem);
    m.ws.pred_partials.ensure((size_t)n_chunks * P.N * P.D * sizeof(float));
    dim3 grid(ceil_div(P.N, 32), n_chunks);
    GB_LAUNCH(predict_chunk_kernel<DM>, grid, 128, smem, s, P, m.ws.pred_partials.as<float>(), tpc);
    GB_LAUNCH(predict_combine_kernel, ceil_div(P.N * P.D, 256), 256, 0, s, m.ws.pred_partials.as<float>(), P.bias, P.preds, P.N, P.D,
              n_chunks, P.add_bias);
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
    // rollout shape: few samples, many trees -> cut the tree range into chunks (see predict_chunk_kernel)
    const int n_t = stop_tree - start_tree;
    const int tiles = ceil_div(N, 32);
    if (n_t >= 1024 && tiles * 4 < 148 * 24 && D <= 8 && (size_t)F * 33 * 4 <= 200 * 1024) {
        int n_chunks = ceil_div(148 * 8, tiles);
        if (n_chunks > n_t / 128) n_chunks = n_t / 128;
        if (n_chunks < 1) n_chunks = 1;
        const int tpc = ceil_div(n_t, n_chunks);
        n_chunks = ceil_div(n_t, tpc);
        if (D <= 1) launch_predict_chunked<1>(m, P, n_chunks, tpc, s);
        else if (D <= 2) launch_predict_chunked<2>(m, P, n_chunks, tpc, s);
        else if (D <= 4) launch_predict_chunked<4>(m, P, n_chunks, tpc, s);
        else launch_predict_chunked<8>(m, P, n_chunks, tpc, s);
        return;
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
