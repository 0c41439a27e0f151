// tree.cu -- per-tree orchestration: init, level loop, leaf emission in the reference's ensemble layout,
// leaf values.
//
// Reference semantics restated:
//   leaf order, greedy     fitter.cpp:292-371,364-365  DFS, left child popped first; one ensemble row per leaf with
//                          its full path (update_ensemble_per_leaf, fitter.cpp:493-515)
//   leaf order, oblivious  fitter.cpp:461-471,517-542  2^depth leaves in heap order; split arrays per TREE,
//                          inequality_directions / edge_weights per LEAF
//   edge weight            node.cpp:131,141            n_child / n_parent (0 if the parent is empty)
//   leaf value             fitter.cpp:545-582          mean of the RAW gradients of the leaf's samples, 0 if none;
//                          a depth-0 leaf never matches (passed=false, :559-564) and keeps value 0
#include "engine.cuh"
#include <utility>
#include <cstddef>

namespace gb {

// ---------------------------------------------------------------- init
__global__ void init_rows_kernel(int *order, int *pnode, int N) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < N) { order[k] = k; pnode[k] = 0; }
}

__global__ void init_nodes_kernel(NodeArrays na, Ctl *ctl, int MAXN, int N, int D, int max_depth, int oblivious) {
    for (int h = threadIdx.x; h < MAXN; h += blockDim.x) {
        na.state[h] = NODE_NONE; na.seg_start[h] = 0; na.seg_len[h] = 0; na.split_f[h] = -1; na.split_j[h] = 0;
        na.split_thr[h] = 0.0f; na.direct[h] = 1; na.rep_begin[h] = 0; na.rep_count[h] = 0; na.best_idx[h] = -1;
        na.best_gain[h] = -INFINITY; na.parent_score[h] = 0.0f; na.leaf_index[h] = -1;
        for (int d = 0; d < D; ++d) na.tot_sum[(size_t)h * D + d] = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        na.seg_len[0] = N;
        // fitter.cpp:300 (greedy): depth == max_depth or empty node -> leaf; oblivious loops while depth < max_depth
        const bool open = (max_depth > 0) && (oblivious || N > 0);
        na.state[0] = open ? NODE_OPEN : NODE_LEAF;
        ctl->n_items = 0; ctl->n_replay = 0; ctl->obl_depth = 0; ctl->obl_best_idx = -1; ctl->obl_has_replay = 0;
        ctl->tree_leaves = 0;
    }
}

// root totals of the fixed-point build_grads (children get theirs from the chosen split's suffix sums)
__global__ void __launch_bounds__(256) root_totals_kernel(const float *__restrict__ bg, long long *tot, const Ctl *ctl, int N, int D) {
    const float scale = exp2f((float)ctl->qexp);
    const long long total = (long long)N * D;
    for (int d = 0; d < D; ++d) {
        long long acc = 0;
        for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < N; r += (long long)gridDim.x * blockDim.x)
            acc += __float2ll_rn(bg[r * D + d] * scale);
        for (int o = 16; o > 0; o >>= 1) acc += shfl_down_ll(acc, o);
        if ((threadIdx.x & 31) == 0 && acc != 0) red_add64(tot + d, acc);
    }
    (void)total;
}

void launch_init_tree(Model &m, int N, cudaStream_t s) {
    Workspace &ws = m.ws;
    if (N > 0) GB_LAUNCH(init_rows_kernel, ceil_div(N, 256), 256, 0, s, ws.order_p[0], ws.pnode_p[0], N);
    GB_LAUNCH(init_nodes_kernel, 1, 256, 0, s, ws.na, ws.ctl.as<Ctl>(), ws.MAXN, N, ws.D, m.cfg.max_depth,
              m.cfg.grow_policy == GBRL_B200_GROW_OBLIVIOUS);
    if (N > 0) {
        int grid = ceil_div(N, 256 * 8);
        if (grid > 1184) grid = 1184;
        GB_LAUNCH(root_totals_kernel, grid, 256, 0, s, ws.bg.as<float>(), ws.na.tot_sum, ws.ctl.as<Ctl>(), N, ws.D);
    }
}

// ---------------------------------------------------------------- ensemble storage
static void grow_buf(DevBuf &b, size_t old_bytes, size_t new_bytes, cudaStream_t s) {
    (void)old_bytes;
    b.ensure(new_bytes, /*keep=*/true, s);
}

void ensure_ensemble_capacity(Model &m, int extra_trees, cudaStream_t s) {
    Ensemble &e = m.ens;
    const int md = m.cfg.max_depth, D = m.cfg.output_dim;
    const int leaves_per_tree = 1 << md;
    const bool obl = m.cfg.grow_policy == GBRL_B200_GROW_OBLIVIOUS;
    const long long need_trees = (long long)e.n_trees + extra_trees;
    const long long have_leaves = e.n_leaves_ub > e.n_leaves ? e.n_leaves_ub : (long long)e.n_leaves;
    const long long need_leaves = have_leaves + (long long)extra_trees * leaves_per_tree;
    if (need_trees > e.cap_trees || need_leaves > e.cap_leaves) {
        long long ct = e.cap_trees, cl = e.cap_leaves;
        if (need_trees > ct) ct = need_trees * 2 + 64;
        if (need_leaves > cl) cl = need_leaves * 2 + 64 * leaves_per_tree;
        const size_t S = obl ? (size_t)ct : (size_t)cl;
        const int mdz = md > 0 ? md : 1;
        grow_buf(e.tree_indices, 0, (size_t)ct * sizeof(int), s);
        grow_buf(e.depths, 0, S * sizeof(int), s);
        grow_buf(e.values, 0, (size_t)cl * D * sizeof(float), s);
        grow_buf(e.feature_indices, 0, S * mdz * sizeof(int), s);
        grow_buf(e.feature_values, 0, S * mdz * sizeof(float), s);
        grow_buf(e.edge_weights, 0, (size_t)cl * mdz * sizeof(float), s);
        grow_buf(e.ineq, 0, (size_t)cl * mdz, s);
        if (!obl) {
            grow_buf(e.heap_feat, 0, (size_t)ct * ((1 << md)) * sizeof(int), s);
            grow_buf(e.heap_thr, 0, (size_t)ct * ((1 << md)) * sizeof(float), s);
            grow_buf(e.heap_leaf, 0, (size_t)ct * ((2 << md)) * sizeof(int), s);
        }
        e.cap_trees = (int)ct; e.cap_leaves = (int)cl;
    }
}

struct EnsPtrs {
    int *tree_indices, *depths, *feature_indices, *heap_feat, *heap_leaf;
    float *values, *feature_values, *edge_weights, *heap_thr;
    uint8_t *ineq;
};
static EnsPtrs ens_ptrs(Ensemble &e) {
    EnsPtrs p;
    p.tree_indices = e.tree_indices.as<int>(); p.depths = e.depths.as<int>(); p.feature_indices = e.feature_indices.as<int>();
    p.heap_feat = e.heap_feat.as<int>(); p.heap_leaf = e.heap_leaf.as<int>(); p.values = e.values.as<float>();
    p.feature_values = e.feature_values.as<float>(); p.edge_weights = e.edge_weights.as<float>();
    p.heap_thr = e.heap_thr.as<float>(); p.ineq = e.ineq.as<uint8_t>();
    return p;
}

// ---------------------------------------------------------------- finalize: greedy
// single CTA; thread 0 numbers the leaves in DFS-left-first order, then all threads write the paths
__global__ void __launch_bounds__(256) finalize_greedy_kernel(NodeArrays na, Ctl *ctl, EnsPtrs E, int md, int D) {
    __shared__ int s_leaf_nodes[4096];
    __shared__ int s_n;
    const int tree = ctl->n_trees, base_leaf = ctl->n_leaves;
    if (threadIdx.x == 0) {
        int stack[2 * MAX_DEPTH_SUPPORTED + 4], sp = 0, n = 0;
        stack[sp++] = 0;
        while (sp > 0) {
            const int h = stack[--sp];
            if (na.state[h] == NODE_SPLIT) { stack[sp++] = 2 * h + 2; stack[sp++] = 2 * h + 1; }   // fitter.cpp:364-365
            else { na.leaf_index[h] = n; if (n < 4096) s_leaf_nodes[n] = h; ++n; }
        }
        s_n = n;
        E.tree_indices[tree] = base_leaf;
        ctl->tree_leaves = n;
    }
    __syncthreads();
    const int n = s_n;
    const int n_all = (2 << md) - 1;
    // heap topology for the O(depth) predict walk
    for (int h = threadIdx.x; h < n_all; h += blockDim.x) {
        const int st = na.state[h];
        if (h < (1 << md)) {
            E.heap_feat[(size_t)tree * (1 << md) + h] = (st == NODE_SPLIT) ? na.split_f[h] : -1;
            E.heap_thr[(size_t)tree * (1 << md) + h] = (st == NODE_SPLIT) ? na.split_thr[h] : 0.0f;
        }
        E.heap_leaf[(size_t)tree * (2 << md) + h] = (st == NODE_LEAF) ? na.leaf_index[h] : -1;
    }
    for (int li = threadIdx.x; li < n; li += blockDim.x) {
        const int h = s_leaf_nodes[li];
        const int depth = level_of(h);
        const size_t row = (size_t)(base_leaf + li);
        E.depths[row] = depth;
        int a = h;
        for (int k = depth - 1; k >= 0; --k) {
            const int par = (a - 1) >> 1;
            const bool right = (a == 2 * par + 2);
            E.feature_indices[row * md + k] = na.split_f[par];
            E.feature_values[row * md + k] = na.split_thr[par];
            E.ineq[row * md + k] = right ? 1 : 0;
            const int np = na.seg_len[par], nc = na.seg_len[a];
            E.edge_weights[row * md + k] = np > 0 ? (float)nc / (float)np : 0.0f;
            a = par;
        }
        for (int k = depth; k < md; ++k) {   // zero the unused tail like the reference's zeroed allocation
            E.feature_indices[row * md + k] = 0; E.feature_values[row * md + k] = 0.0f;
            E.ineq[row * md + k] = 0; E.edge_weights[row * md + k] = 0.0f;
        }
        for (int d = 0; d < D; ++d) E.values[row * D + d] = 0.0f;
    }
}

// ---------------------------------------------------------------- finalize: oblivious
__global__ void __launch_bounds__(256) finalize_oblivious_kernel(NodeArrays na, Ctl *ctl, EnsPtrs E, int md, int D) {
    const int tree = ctl->n_trees, base_leaf = ctl->n_leaves;
    const int depth = ctl->obl_depth, n = 1 << depth, lb = level_base(depth);
    if (threadIdx.x == 0) {
        E.tree_indices[tree] = base_leaf;
        E.depths[tree] = depth;
        ctl->tree_leaves = n;
        for (int k = 0; k < md; ++k) {
            const int h = level_base(k);      // every node of level k carries the same split
            E.feature_indices[(size_t)tree * md + k] = k < depth ? na.split_f[h] : 0;
            E.feature_values[(size_t)tree * md + k] = k < depth ? na.split_thr[h] : 0.0f;
        }
    }
    for (int li = threadIdx.x; li < n; li += blockDim.x) {
        const int h = lb + li;
        na.leaf_index[h] = li;
        const size_t row = (size_t)(base_leaf + li);
        int a = h;
        for (int k = depth - 1; k >= 0; --k) {
            const int par = (a - 1) >> 1;
            const bool right = (a == 2 * par + 2);
            E.ineq[row * md + k] = right ? 1 : 0;
            const int np = na.seg_len[par], nc = na.seg_len[a];
            E.edge_weights[row * md + k] = np > 0 ? (float)nc / (float)np : 0.0f;
            a = par;
        }
        for (int k = depth; k < md; ++k) { E.ineq[row * md + k] = 0; E.edge_weights[row * md + k] = 0.0f; }
        for (int d = 0; d < D; ++d) E.values[row * D + d] = 0.0f;
    }
}

// ---------------------------------------------------------------- leaf values
// rows are grouped by node in `order`, so a warp usually sees a single leaf: warp-reduce then one REDG per dim
__global__ void __launch_bounds__(256)
leaf_sums_kernel(const float *__restrict__ raw, const int *__restrict__ order, const int *__restrict__ pnode, int *__restrict__ nid, NodeArrays na,
                 const Ctl *ctl, long long *leaf_acc /*[leaves][1+D]*/, int N, int D) {
    const float scale = exp2f((float)ctl->qexp_raw);
    const int lane = threadIdx.x & 31;
    for (long long k0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) - lane; k0 < N; k0 += (long long)gridDim.x * blockDim.x) {
        const long long k = k0 + lane;
        int li = -1, i = 0;
        if (k < N) {
            i = order[k];
            const int h = pnode[k];
            nid[i] = h;                      // by row, once per tree: launch_update_preds_from_nodes reads it
            li = na.leaf_index[h];
        }
        const int li0 = __shfl_sync(0xffffffffu, li, 0);
        const bool uniform = __all_sync(0xffffffffu, li == li0);
        if (uniform) {
            if (li0 < 0) continue;
            for (int d = 0; d < D; ++d) {
                long long q = __float2ll_rn(raw[(size_t)i * D + d] * scale);
                for (int o = 16; o > 0; o >>= 1) q += shfl_down_ll(q, o);
                if (lane == 0) red_add64(leaf_acc + (size_t)li0 * (1 + D) + 1 + d, q);
            }
            if (lane == 0) red_add64(leaf_acc + (size_t)li0 * (1 + D), 32);
        } else if (li >= 0) {
            for (int d = 0; d < D; ++d)
                red_add64(leaf_acc + (size_t)li * (1 + D) + 1 + d, __float2ll_rn(raw[(size_t)i * D + d] * scale));
            red_add64(leaf_acc + (size_t)li * (1 + D), 1);
        }
    }
}

__global__ void leaf_values_kernel(const long long *leaf_acc, Ctl *ctl, EnsPtrs E, NodeArrays na, int D, int md, int oblivious, int MAXN) {
    const int n = ctl->tree_leaves, base_leaf = ctl->n_leaves, tree = ctl->n_trees;
    const double inv = exp2((double)(-ctl->qexp_raw));
    // depth-0 tree: the reference never matches the single leaf, its value stays 0
    const bool depth0 = oblivious ? (E.depths[tree] == 0) : (na.state[0] != NODE_SPLIT);
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n * D; t += gridDim.x * blockDim.x) {
        const int li = t / D, d = t % D;
        const long long cnt = leaf_acc[(size_t)li * (1 + D)];
        float v = 0.0f;
        if (cnt > 0 && !depth0) v = (float)(((double)leaf_acc[(size_t)li * (1 + D) + 1 + d] * inv) / (double)cnt);
        E.values[(size_t)(base_leaf + li) * D + d] = v;
    }
    (void)md; (void)MAXN;
}

__global__ void commit_tree_kernel(Ctl *ctl) {
    ctl->n_leaves += ctl->tree_leaves;
    ctl->n_trees += 1;
}

void launch_finalize_tree(Model &m, const float *raw_grads, int N, int cur, cudaStream_t s) {
    (void)cur;
    Workspace &ws = m.ws;
    const int md = m.cfg.max_depth, D = ws.D;
    const bool obl = m.cfg.grow_policy == GBRL_B200_GROW_OBLIVIOUS;
    EnsPtrs E = ens_ptrs(m.ens);
    Ctl *ctl = ws.ctl.as<Ctl>();
    if (obl) GB_LAUNCH(finalize_oblivious_kernel, 1, 256, 0, s, ws.na, ctl, E, md, D);
    else GB_LAUNCH(finalize_greedy_kernel, 1, 256, 0, s, ws.na, ctl, E, md, D);
    // leaf accumulators live in the (now free) replay score buffer region: use chunk_sums' sibling buffer
    const size_t acc_bytes = (size_t)(1 << md) * (1 + D) * sizeof(long long);
    ws.loss_parts.ensure(acc_bytes);
    GB_CUDA(cudaMemsetAsync(ws.loss_parts.p, 0, acc_bytes, s));
    if (N > 0) {
        int grid = ceil_div(N, 256);
        if (grid > 2368) grid = 2368;
        GB_LAUNCH(leaf_sums_kernel, grid, 256, 0, s, raw_grads, ws.order_p[0], ws.pnode_p[0], ws.nid.as<int>(), ws.na, ctl,
                  ws.loss_parts.as<long long>(), N, D);
    }
    GB_LAUNCH(leaf_values_kernel, ceil_div((1 << md) * D, 256), 256, 0, s, ws.loss_parts.as<long long>(), ctl, E, ws.na, D, md, obl ? 1 : 0, ws.MAXN);
    GB_LAUNCH(commit_tree_kernel, 1, 1, 0, s, ctl);
}

// ---------------------------------------------------------------- one tree
// Speculative levels (ws.spec).  The near-tie replay (split.cu) re-scores a handful of candidates per level in the reference's
// sequential fp32 arithmetic; its chains are latency-bound walks of single warps, and they change the exact-tier decision
// rarely (C2: 0 of 85 replayed nodes, C3: 6 of 104 levels, C5: 3 of 313 nodes; profiles/README.md).  So the tree keeps growing
// on the exact-tier winners while every level's replay runs on its own low-priority side stream, against per-level buffers
// (replay items, planes, summaries; the level's row order and histograms are kept too).  A verification kernel behind each
// replay compares the decision in the reference's arithmetic with the one taken and records the lowest level that differs.
// At the end of the tree the host reads that one word: none -> the tree IS the reference's tree; level L -> node states and
// row -> node ids are returned to level L (rollback_kernel; order, histograms, scan results and replayed scores of L are still
// there), L is decided again -- this time from the replayed scores -- and the levels below L are grown speculatively again.  L strictly increases, so
// a tree costs at most max_depth returns; the result never depends on speculation.
namespace {
struct SlotBind {            // swaps a slot's replay buffers into the workspace for the launches of one level
    Workspace &ws; ReplaySlot *sl;
    static void sw(DevBuf &a, DevBuf &b) { std::swap(a.p, b.p); std::swap(a.bytes, b.bytes); }
    void swap_all() {
        sw(ws.replay, sl->replay); sw(ws.replay_scores, sl->replay_scores); sw(ws.rgrad, sl->rgrad);
        sw(ws.rbits, sl->rbits); sw(ws.rmeta, sl->rmeta); sw(ws.rwide, sl->rwide);
    }
    SlotBind(Workspace &w, ReplaySlot *s) : ws(w), sl(s) { if (sl) swap_all(); }
    ~SlotBind() { if (sl) swap_all(); }
};
}  // namespace

void grow_tree(Model &m, const float *X, const float *raw_grads, int N, int F, cudaStream_t s) {
    (void)F;
    Workspace &ws = m.ws;
    const int md = m.cfg.max_depth;
    ensure_ensemble_capacity(m, 1, s);
    const bool spec = ws.spec && md > 0 && N > 0;
    if (spec) {
        ws.order_p[0] = ws.order_lv[0].as<int>(); ws.order_p[1] = ws.order_lv[1].as<int>();
        ws.pnode_p[0] = ws.pnode_lv[0].as<int>(); ws.pnode_p[1] = ws.pnode_lv[1].as<int>();
        GB_CUDA(cudaMemsetAsync(ws.spec_flag.p, 0xff, sizeof(unsigned int), s));
        m.spec_trees += 1;
    } else {
        ws.order_p[0] = ws.order[0].as<int>(); ws.order_p[1] = ws.order[1].as<int>();
        ws.pnode_p[0] = ws.pnode[0].as<int>(); ws.pnode_p[1] = ws.pnode[1].as<int>();
        ws.hist_p[0] = ws.hist[0].as<long long>(); ws.hist_p[1] = ws.hist[1].as<long long>();
    }
    ws.count_stats = true;
    { ProfScope ps(m, P_PRE, s); raw_grad_scale(m, raw_grads, N, s); launch_init_tree(m, N, s); }
    const size_t slot_bytes = (size_t)ws.nT * NB * FT * (1 + ws.D) * sizeof(long long);
    if (md > 0) { ProfScope ps(m, P_DECIDE, s); launch_plan_level(m, 0, s); }
    int start = 0;
    for (;;) {
        for (int level = start; level < md; ++level) {
            if (spec) {
                ws.hist_p[level & 1] = ws.hist_lv[level].as<long long>();
                ws.order_p[0] = ws.order_lv[level].as<int>(); ws.order_p[1] = ws.order_lv[level + 1].as<int>();
                ws.pnode_p[0] = ws.pnode_lv[level].as<int>(); ws.pnode_p[1] = ws.pnode_lv[level + 1].as<int>();
            }
            { ProfScope ps(m, P_DECIDE, s); GB_CUDA(cudaMemsetAsync(ws.hist_p[level & 1], 0, slot_bytes << level, s)); }
            { ProfScope ps(m, P_HIST, s); launch_histogram(m, level, s); }
            if (m.world > 1) { ProfScope ps(m, P_ALLREDUCE, s);
                dist_allreduce_hist(m, ws.hist_p[level & 1], (slot_bytes << level) / sizeof(long long), s); }
            { ProfScope ps(m, P_SCAN, s); launch_scan(m, level, s); }
            ReplaySlot *slot = spec ? &ws.slots[level] : nullptr;
            SlotBind bind(ws, slot);
            { ProfScope ps(m, P_SELECT, s); launch_select_and_replay(m, X, level, s, slot); }
            { ProfScope ps(m, P_DECIDE, s);      // (the node states of the level were snapshotted by the selection kernel)
              launch_decide(m, level, s, slot == nullptr);
              if (slot) launch_verify(m, level, s, *slot); }
            { ProfScope ps(m, P_PART, s); launch_partition(m, X, level, 0, s); }
        }
        if (!spec) break;
        unsigned int flip;
        { ProfScope ps(m, P_SPEC, s);
          for (int l = 0; l < md; ++l)
              if (ws.slots[l].pending) { GB_CUDA(cudaStreamWaitEvent(s, ws.slots[l].ev_done, 0)); ws.slots[l].pending = false; }
          GB_CUDA(cudaMemcpyAsync(ws.h_spec_flag, ws.spec_flag.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, s)); }
        GB_CUDA(cudaStreamSynchronize(s));
        flip = *ws.h_spec_flag;
        if (flip >= (unsigned int)md) break;
        // The replay changed the decision of level L: return to it.  Its scan results (best candidate, replay list: per-node arrays)
        // and its replayed scores (the level's slot) are all still in place, so the level is simply DECIDED again, this time with
        // the replay, then partitioned; the levels below are grown speculatively again.
        const int L = (int)flip;
        m.spec_rollbacks += 1;
        { ProfScope ps(m, P_SPEC, s); launch_rollback(m, L, s); }
        ws.hist_p[L & 1] = ws.hist_lv[L].as<long long>();
        ws.order_p[0] = ws.order_lv[L].as<int>(); ws.order_p[1] = ws.order_lv[L + 1].as<int>();
        ws.pnode_p[0] = ws.pnode_lv[L].as<int>(); ws.pnode_p[1] = ws.pnode_lv[L + 1].as<int>();
        {
            ReplaySlot &sl = ws.slots[L];
            SlotBind bind(ws, &sl);
            ProfScope ps(m, P_DECIDE, s);
            // the oblivious decision reads the level's exact-tier winner and candidate count from the control block: the snapshot's
            constexpr size_t o0 = offsetof(Ctl, obl_best_idx), o1 = offsetof(Ctl, obl_band) + sizeof(float);
            GB_CUDA(cudaMemcpyAsync(ws.ctl.as<char>() + o0, sl.ctl_snap.as<char>() + o0, o1 - o0, cudaMemcpyDeviceToDevice, s));
            ws.count_stats = false;          // verify_kernel counted this level already
            launch_decide(m, L, s, true);
            ws.count_stats = true;
        }
        { ProfScope ps(m, P_PART, s); launch_partition(m, X, L, 0, s); }
        start = L + 1;
    }
    ws.count_stats = true;
    { ProfScope ps(m, P_FIN, s); launch_finalize_tree(m, raw_grads, N, 0, s); }
    m.ens.n_trees += 1;   // host mirror; the exact n_leaves is read back by the caller (sync_ctl)
    m.ens.n_leaves_ub = (m.ens.n_leaves_ub > m.ens.n_leaves ? m.ens.n_leaves_ub : m.ens.n_leaves) + (1ll << md);
}

}  // namespace gb
