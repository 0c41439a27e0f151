// engine.cuh -- internal data structures of the B200 fit/predict engine.
//
// HBM layout (all device-resident for the lifetime of a step()/fit() call):
//   X        [N x F]  fp32 row-major      caller's matrix (borrowed or staged copy)      node.cpp:339
//   codes    [nT][N][32] u16              per-feature-tile candidate-bin codes of X:
//                                         code(x) = #{j : thr[f][j] < x} in [0, n_bins]
//                                         (x > thr[f][j]  <=>  code > j); stored as code << 6;
//                                         64 B per row and tile
//   bg       [N x D]  fp32                build_grads (fitter.cpp:57-64)
//   order    [N] int32 (ping-pong)        rows grouped by tree node, ascending inside a node
//                                         (== the reference's per-node sample_indices, node.cpp:86-96)
//   pnode    [N] int32 (ping-pong)        heap id of the node the row at each POSITION of `order` sits in (coalesced for the
//                                         partition and the leaf sums; `nid`, the same by ROW, is written once per tree)
//   hist     [slot][nT][256][32][1+D] i64 per-node (count, sum of fixed-point build_grads) per
//                                         (feature, code-1); two level buffers (parent / current)
//   scores   [slot][F*n_bins] fp32        per-(node,candidate) score (exact-arithmetic path)
#pragma once
#include "common.cuh"
#include "../../include/gbrl_b200.h"
#include <vector>

namespace gb {

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    void ensure(size_t n, bool keep = false, cudaStream_t s = 0);
    void release();
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

enum NodeState : int { NODE_NONE = 0, NODE_OPEN = 1, NODE_LEAF = 2, NODE_SPLIT = 3 };

struct Item {          // one histogram work item: rows [k0,k1) of `order` (all in one node) x one tile
    int slot, tile, k0, k1;
};
struct ReplayItem {    // re-score (node, candidate) in the reference's sequential order; cand -1 = parent
    int node, cand;
};

// device-side control block (one per model workspace); everything the level pipeline needs to decide
// without a host round trip
struct Ctl {
    int n_items;             // histogram items of the current level
    int n_replay;            // replay items of the current level
    int n_partials;          // histogram partials handed out at the current level (pool index)
    int replay_overflow;
    int qexp;                // fixed-point exponent of build_grads: q = rint(g * 2^qexp)
    int qexp_raw;            // same for the raw gradients (leaf values)
    unsigned int max_abs_bg; // float bits of max |bg|
    unsigned int max_abs_raw;
    int n_leaves;            // running leaf count of the ensemble (device copy)
    int n_trees;
    int tree_leaves;         // leaves of the tree being finalised
    int obl_best_idx;        // oblivious: chosen candidate of the level (-1: stop)
    float obl_best;
    int obl_depth;           // oblivious: depth reached
    int obl_has_replay;
    float obl_band;
    int bg_nonfinite;        // some build_grad is NaN/inf: the reference's scores are all NaN -> no split
    unsigned int stat_max_noise;   // float bits: max over replayed candidates of |replayed - exact| / (2^-24 sqrt(n) |score|)
    long long stat_replay_items, stat_replay_nodes, stat_nodes_evaluated, stat_replay_overflow, stat_hist_rows;
    long long stat_chain_fast, stat_chain_slow, stat_chain_seq;   // chain groups applied from their record / run sequentially; lanes run sequentially
    long long stat_replay_flips;   // split decisions in which the replayed arg-max differs from the exact-tier arg-max
};

// per-node arrays, heap indexed, MAXN = 2^(max_depth+1)-1 entries each
struct NodeArrays {
    int *seg_start, *seg_len, *state, *split_f, *split_j, *direct, *rep_begin, *rep_count, *best_idx;
    float *split_thr, *best_gain, *parent_score, *band;
    long long *tot_sum;      // [MAXN x D] fixed-point sum of build_grads of the node
    int *leaf_index;         // leaf number inside the tree (DFS-left-first), -1 for internal nodes
};

struct Optimizer {
    int sched, start_idx, stop_idx, T;
    float init_lr, stop_lr;
};

struct Ensemble {            // reference layout, types.h:279-304 (numerical features only)
    DevBuf tree_indices, depths, values, feature_indices, feature_values, edge_weights, ineq;
    // auxiliary per-tree heap topology for the O(depth) walk used by the greedy predict kernel
    DevBuf heap_feat, heap_thr, heap_leaf;
    int cap_trees = 0, cap_leaves = 0;
    int n_trees = 0, n_leaves = 0;
    long long n_leaves_ub = 0;   // host-side upper bound of n_leaves while trees are grown without a host sync
};

// Per-level buffers of a SPECULATIVE near-tie replay (tree.cu grow_tree): the level's replay items, side-bit planes, group
// summaries and replayed scores live here while the chains are walked on `stream`, concurrently with the deeper levels
// of the tree that the main stream grows on the exact-tier winners.
struct ReplaySlot {
    DevBuf replay, replay_scores, rgrad, rbits, rmeta, rwide, ctl_snap;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_sel = nullptr, ev_dec = nullptr, ev_done = nullptr;
    bool pending = false;
};

struct Workspace {           // sized for (N, F, D, depth); reused across calls with the same shape
    int N = 0, F = 0, nT = 0, D = 0, depth = 0, MAXN = 0, B = 0;
    int tile_lo = 0, tile_hi = 0;   // feature tiles owned by this rank [tile_lo, tile_hi)
    int row_groups = 1, row_group = 0;   // 2-D sharding: ranks beyond the tile count split each node's row chunks round-robin
    DevBuf codesT;                  // feature-major copy of the codes, all features: [F][codesT_stride] u16
    long long codesT_stride = 0;
    bool use_codesT = false;
    int codes_rows = 0;             // rows of the code matrix (tile stride); >= N when a tree is grown on a mini-batch
    int row_offset = 0;             // first row of the current mini-batch inside the code matrix
    DevBuf bgq;                          // build_grads as fixed point, split (lo 18 bits, hi): what the histogram atomics add
    DevBuf codes, thr, thrT, bg, order[2], pnode[2], nid, rflag, chunk_sums, hist[2], scores, cand_flags;
    int chunk_cap = 0;                   // capacity of chunk_sums (entries); entry [chunk_cap] is the partition's done counter
    DevBuf items, replay, replay_scores, nodes, ctl, tile_best, obl_tot, sort_tmp, colbuf[2], lrs;
    DevBuf rgrad, rbits, rmeta;          // replay streams: order-space build_grads, side-bit planes, per-item offsets / modes / counts
    long long rbits_words = 0;
    DevBuf dwide;                        // GPU-wide evaluation of the mean / std chains (preprocess.cu)
    DevBuf rwide;                        // GPU-wide replay: per-group sums / predictions / summaries / tags, per-item hand-over
    long long rwide_groups = 0;
    int n_sms = 0;
    DevBuf sort_offsets;
    DevBuf xstage, gstage, tstage, preds_full, grads_fit, loss_parts, pstage, pred_partials;
    NodeArrays na{};
    // current / next row order and the histogram buffer of (level & 1): ws.order[] / ws.hist[], or -- while a tree is grown
    // speculatively -- the per-level buffers below (a rollback needs the order and the histograms of the level it returns to)
    int *order_p[2] = {nullptr, nullptr};
    int *pnode_p[2] = {nullptr, nullptr};
    long long *hist_p[2] = {nullptr, nullptr};
    bool spec = false;                   // speculative replay enabled for this workspace shape
    bool count_stats = true;             // false while a rolled-back level is decided a second time
    DevBuf order_lv[MAX_DEPTH_SUPPORTED + 1], pnode_lv[MAX_DEPTH_SUPPORTED + 1], hist_lv[MAX_DEPTH_SUPPORTED];
    ReplaySlot slots[MAX_DEPTH_SUPPORTED];
    DevBuf state_snap, spec_flag;        // node states of every level before its decision; lowest level whose decision the replay changed
    unsigned int *h_spec_flag = nullptr; // pinned host copy of spec_flag
    size_t sort_tmp_bytes = 0;
    int replay_cap = 0, items_cap = 0;
};

// per-kernel-class device timing (CUDA events on the launching stream), enabled by gbrl_b200_profile()
enum ProfCat : int { P_CAND = 0, P_BIN, P_PRE, P_HIST, P_ALLREDUCE, P_SCAN, P_SELECT, P_DECIDE, P_PART, P_FIN, P_PRED, P_SPEC, P_NCAT };

struct FitSession {          // state of fit_begin / fit_iterate / fit_end (fitter.cpp:117-261)
    bool active = false, incremental = true;
    const float *X = nullptr, *T = nullptr;
    int N = 0, F = 0, it = 0, batch_start = 0, batch_n = 0, n_trees0 = 0;
    float *g_regular = nullptr, *g_last = nullptr;
};

struct Model {
    gbrl_b200_config cfg{};
    int device = 0;
    int n_num_features = 0, n_cat_features = 0, iteration = 0;
    Ensemble ens;
    std::vector<Optimizer> opts;
    DevBuf bias, feature_weights, rev_num_map, d_opts;
    std::vector<float> h_bias, h_fw;
    std::vector<int> h_mapping, h_rev_num, h_rev_cat;
    std::vector<uint8_t> h_numerics;
    Workspace ws;
    // multi-GPU
    void *nccl_comm = nullptr;
    int rank = 0, world = 1;
    // statistics
    long long replay_items = 0, replay_nodes = 0, replay_overflow = 0, nodes_evaluated = 0, chain_fast = 0, chain_slow = 0, chain_seq = 0, replay_flips = 0;
    long long spec_trees = 0, spec_rollbacks = 0;   // trees grown speculatively / levels that had to be rolled back
    bool have_candidates = false;
    long long hist_rows = 0;          // rows scanned by the histogram kernel (read back from Ctl)
    float max_noise = 0.0f;           // see Ctl::stat_max_noise
    FitSession fs;
    DevBuf fit_x, fit_t, fit_batch_preds;
    // profiling
    bool profile = false;
    std::vector<cudaEvent_t> prof_events;      // pairs
    std::vector<int> prof_cats;
    std::vector<long long> prof_launches;
    size_t prof_used = 0;
    double prof_ms[P_NCAT] = {0};
    long long prof_n[P_NCAT] = {0};
};

struct ProfScope {
    Model &m; int cat; cudaStream_t s; size_t idx = 0; long long l0 = 0; bool on;
    ProfScope(Model &m_, int cat_, cudaStream_t s_);
    ~ProfScope();
};
void prof_collect(Model &m, cudaStream_t s);

// ---------------------------------------------------------------- kernel launchers (one per .cu)
// candidates.cu
void compute_thresholds(Model &m, const float *X, int N, int F, cudaStream_t s);
void bin_features(Model &m, const float *X, int N, int F, cudaStream_t s);
// preprocess.cu
void build_grads(Model &m, const float *grads, int N, cudaStream_t s);          // -> ws.bg, ctl.qexp
void column_mean_ref(Model &m, const float *mat, int N, int D, float *out_dev, cudaStream_t s);
void multirmse_grads(Model &m, const float *preds, const float *targets, float *grads, int n, cudaStream_t s);
void multirmse_loss(Model &m, const float *preds, const float *targets, int n, float *loss_host, cudaStream_t s);
void raw_grad_scale(Model &m, const float *grads, int N, cudaStream_t s);       // -> ctl.qexp_raw
// tree growth
void grow_tree(Model &m, const float *X, const float *raw_grads, int N, int F, cudaStream_t s);
// histogram.cu
void launch_plan_level(Model &m, int level, cudaStream_t s);      // level 0; later levels: launch_decide plans level + 1
struct PlanParams;
PlanParams plan_params(const Model &m);
int hist_item_rows(const Model &m);
void launch_histogram(Model &m, int level, cudaStream_t s);
// split.cu
void launch_scan(Model &m, int level, cudaStream_t s);
void launch_select_and_replay(Model &m, const float *X, int level, cudaStream_t s, ReplaySlot *slot = nullptr);
void launch_decide(Model &m, int level, cudaStream_t s, bool use_replay = true);
void launch_verify(Model &m, int level, cudaStream_t s, ReplaySlot &slot);      // speculative level: replayed decision vs the one taken
void launch_rollback(Model &m, int level, cudaStream_t s);
// partition.cu
void launch_partition(Model &m, const float *X, int level, int cur, cudaStream_t s);
// tree.cu
void launch_init_tree(Model &m, int N, cudaStream_t s);
void launch_finalize_tree(Model &m, const float *raw_grads, int N, int cur, cudaStream_t s);
void ensure_ensemble_capacity(Model &m, int extra_trees, cudaStream_t s);
// predict.cu
void launch_predict(Model &m, const float *X, int N, int F, int start_tree, int stop_tree, float *preds, bool add_bias,
                    cudaStream_t s);
void launch_update_preds_last_tree(Model &m, const float *X, int N, int F, float *preds, cudaStream_t s);
void launch_update_preds_from_nodes(Model &m, int N, float *preds, cudaStream_t s);
void upload_optimizers(Model &m, cudaStream_t s);
void rebuild_heap_topology(Model &m, cudaStream_t s);
// dist.cu
void dist_allreduce_hist(Model &m, long long *buf, size_t count, cudaStream_t s);

}  // namespace gb
