/*
 * gbrl_b200.h -- C-ABI of the B200-native fit/predict engine (libgbrl_b200.so).
 *
 * This is the drop-in seam for the reference's `class GBRL` (gbrl/src/cpp/gbrl.h:56-518), i.e. exactly the
 * calls that the reference's pybind11 layer (gbrl/src/cpp/binding.cpp:421-1134) forwards to: plain
 * pointers + sizes + a host/device flag per buffer (the reference's `dataHolder<T>{T* data; deviceType
 * device}`, gbrl/src/cpp/types.h:262-270).  No torch / pybind types appear here.  INTEGRATION.md shows
 * the binding a gbrl maintainer would add on top of this header.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on error; gbrl_b200_last_error() gives the message
 *    (the reference throws std::runtime_error -> Python RuntimeError; the host shim re-raises the same).
 *  - `*_dev` arguments: 0 = pointer is host memory, 1 = pointer is CUDA device memory on the model's
 *    device.  Inputs are borrowed for the duration of the call only (binding.cpp:452-530 contract).
 *  - matrices are C-contiguous row-major float32: obs [n_samples x n_features], grads/targets/preds
 *    [n_samples x output_dim]  (node.cpp:339, binding.cpp:102-199).
 *  - `stream` is a cudaStream_t (NULL = legacy default stream).  Calls are synchronous with respect to
 *    the host on return, like the reference.
 *  - there is NO CPU fallback: creating a model without a usable CUDA device fails.
 */
#ifndef GBRL_B200_H
#define GBRL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gbrl_b200_model gbrl_b200_model;

enum { GBRL_B200_SCORE_L2 = 0, GBRL_B200_SCORE_COSINE = 1 };       /* scoreFunc,     types.h:120 */
enum { GBRL_B200_GEN_UNIFORM = 0, GBRL_B200_GEN_QUANTILE = 1 };    /* generatorType, types.h:129 */
enum { GBRL_B200_GROW_GREEDY = 0, GBRL_B200_GROW_OBLIVIOUS = 1 };  /* growPolicy,    types.h:138 */
enum { GBRL_B200_SCHED_CONST = 0, GBRL_B200_SCHED_LINEAR = 1 };    /* schedulerFunc, types.h:100 */

/* constructor arguments of GBRL::GBRL (gbrl.cpp:76-114; kwargs and defaults binding.cpp:423-440) plus
 * the engine's own parity knobs. */
typedef struct {
    int input_dim, output_dim, policy_dim;
    int max_depth;          /* default 4   */
    int min_data_in_leaf;   /* default 0   */
    int n_bins;             /* default 256; this engine supports 1..256 */
    int par_th;             /* default 10; only used to mirror the reference's thread partitioning */
    int batch_size;         /* default 5000 (fit() mini-batch) */
    int split_score_func;   /* GBRL_B200_SCORE_*  */
    int generator_type;     /* GBRL_B200_GEN_*    */
    int grow_policy;        /* GBRL_B200_GROW_*   */
    int verbose;
    int device_ordinal;     /* CUDA device the model lives on */
    /* parity knobs (no reference equivalent) */
    int ref_threads;        /* emulated omp_get_max_threads() of the reference host: fixes the summation
                               partition of its mean/std/bias reductions (math_ops.cpp:255-513). >=1 */
    int tie_replay;         /* 1: candidates whose exact-arithmetic score is within the rounding band of
                               the best are re-scored in the reference's sequential fp32 order so that
                               the chosen split is bit-identical to the reference's; 0: exact-arithmetic
                               arg-max only */
    float band_kappa;       /* band = kappa * 2^-24 * sqrt(n_node) * |score|; <=0 -> default 6 */
    int use_subtraction;    /* 1: histogram only the smaller child, derive the sibling from the parent */
    int hist_variant;       /* 0: streaming histogram kernel (cp.async row ring, carried shared histogram); 1: per-item kernel;
                               2: streaming kernel with 32-warp CTAs for output_dim == 1 (measured slower than the 24-warp default) */
    int replay_variant;     /* bit 0: 0 = replay chains spread over the whole GPU where output_dim <= 2, 1 = one CTA per replay item;
                               bit 1: 0 = speculative levels (the tree keeps growing on the exact-tier winners while the chains of a
                               level are walked on a side stream; a level whose decision the replay changes is rolled back),
                               1 = every level waits for its replay */
} gbrl_b200_config;

typedef struct {           /* mirrors binding.cpp:309-328 get_metadata + engine statistics */
    int input_dim, output_dim, policy_dim, max_depth, min_data_in_leaf, n_bins, par_th, batch_size;
    int split_score_func, generator_type, grow_policy, verbose;
    int n_num_features, n_cat_features, n_trees, n_leaves, iteration;
    /* statistics since creation */
    long long kernel_launches;   /* number of engine kernels launched */
    long long replay_items;      /* (node,candidate) pairs re-scored in reference order */
    long long replay_nodes;      /* nodes for which the near-tie replay was triggered */
    long long replay_overflow;   /* replay requests dropped because the list was full (should be 0) */
    long long nodes_evaluated;   /* nodes whose candidates were scored */
    float max_noise_ratio;       /* max over replayed candidates of |reference-order score - exact score| in units of
                                    2^-24*sqrt(n)*|score|: the quantity band_kappa has to cover twice */
    long long chain_blocks_fast; /* replay sub-blocks evaluated by the parallel chain evaluator (csrc/chain.cuh) */
    long long chain_blocks_slow; /* replay sub-blocks advanced piecewise (a binade change inside the block) */
    long long chain_lanes_seq;   /* lanes (R rows each) of those blocks that were run as a plain sequential float chain */
    long long replay_flips;      /* split decisions (greedy: nodes, oblivious: levels) in which the near-tie replay changed the exact-tier winner */
    long long spec_trees;        /* trees grown with the replay chains off the critical path (speculative levels, see replay_variant) */
    long long spec_rollbacks;    /* levels of those trees that were decided a second time because the replay changed their decision */
} gbrl_b200_metadata;

const char *gbrl_b200_last_error(void);
int gbrl_b200_cuda_available(void);                       /* GBRL::cuda_available, gbrl.cpp:541-547 */

int  gbrl_b200_create(const gbrl_b200_config *cfg, gbrl_b200_model **out);   /* GBRL::GBRL, gbrl.cpp:76 */
void gbrl_b200_destroy(gbrl_b200_model *m);                                   /* GBRL::~GBRL, gbrl.cpp:141 */

/* GBRL::set_bias gbrl.cpp:213, set_feature_weights :241, set_feature_mapping :269 */
int gbrl_b200_set_bias(gbrl_b200_model *m, const float *bias, int n, int bias_dev);
int gbrl_b200_set_feature_weights(gbrl_b200_model *m, const float *w, int n, int w_dev);
int gbrl_b200_set_feature_mapping(gbrl_b200_model *m, const int *feature_mapping, const uint8_t *mapping_numerics, int n);
int gbrl_b200_get_bias(gbrl_b200_model *m, float *out);                       /* gbrl.cpp:318 */
int gbrl_b200_get_feature_weights(gbrl_b200_model *m, float *out);            /* gbrl.cpp:334 */
int gbrl_b200_get_feature_mapping(gbrl_b200_model *m, int *mapping, uint8_t *numerics, int *rev_num, int *rev_cat);

/* GBRL::set_optimizer gbrl.cpp:452-525 (SGD only; Adam is CPU-only in the reference too, :476) */
int gbrl_b200_set_optimizer(gbrl_b200_model *m, int scheduler, float init_lr, int start_idx, int stop_idx, float stop_lr, int T);
int gbrl_b200_n_optimizers(gbrl_b200_model *m);
int gbrl_b200_get_optimizer(gbrl_b200_model *m, int i, int *scheduler, float *init_lr, int *start_idx, int *stop_idx, float *stop_lr, int *T);
int gbrl_b200_get_scheduler_lrs(gbrl_b200_model *m, float *out);              /* gbrl.cpp:527-539 */

/* GBRL::step gbrl.cpp:939-981 -> one boosting iteration on caller-supplied gradients */
int gbrl_b200_step(gbrl_b200_model *m, const float *obs, int obs_dev, const float *grads, int grads_dev,
                   int n_samples, int n_features, void *stream);
/* GBRL::fit gbrl.cpp:983-1104 -> supervised MultiRMSE loop; returns the final full-data loss */
int gbrl_b200_fit(gbrl_b200_model *m, const float *obs, int obs_dev, const float *targets, int targets_dev,
                  int iterations, int n_samples, int n_features, int shuffle, float *loss_out, void *stream);
/* fit() as three calls (fit == fit_begin + fit_iterate(iterations) + fit_end): lets a caller time exactly K
 * boosting iterations of fitter.cpp:176-231 with the inputs already resident in HBM (bench.py). */
int gbrl_b200_fit_begin(gbrl_b200_model *m, const float *obs, int obs_dev, const float *targets, int targets_dev,
                        int n_samples, int n_features, int shuffle, void *stream);
int gbrl_b200_fit_iterate(gbrl_b200_model *m, int iterations, int sync, void *stream);
int gbrl_b200_fit_end(gbrl_b200_model *m, float *loss_out, void *stream);
/* GBRL::predict gbrl.cpp:369-422; `preds` receives n_samples*output_dim floats */
int gbrl_b200_predict(gbrl_b200_model *m, const float *obs, int obs_dev, int n_samples, int n_features,
                      int start_tree_idx, int stop_tree_idx, float *preds, int preds_dev, void *stream);

int gbrl_b200_get_metadata(gbrl_b200_model *m, gbrl_b200_metadata *out);      /* binding.cpp:309-328 */
/* binding.cpp:330-390 get_ensemble_data: host copies in the reference layout.  S = n_trees (oblivious)
 * or n_leaves (greedy).  Sizes: tree_indices[n_trees], depths[S], values[n_leaves*output_dim],
 * feature_indices/feature_values[S*max_depth], edge_weights/inequality_directions[n_leaves*max_depth]. */
int gbrl_b200_get_ensemble(gbrl_b200_model *m, int *tree_indices, int *depths, float *values,
                           int *feature_indices, float *feature_values, float *edge_weights,
                           uint8_t *inequality_directions);
/* load an ensemble given in the reference layout (used by GBRL.load / copy-ctor / tests) */
int gbrl_b200_set_ensemble(gbrl_b200_model *m, int n_trees, int n_leaves, const int *tree_indices, const int *depths,
                           const float *values, const int *feature_indices, const float *feature_values,
                           const float *edge_weights, const uint8_t *inequality_directions, int n_num_features);
/* restores ensembleMetaData::iteration of a loaded model (GBRL::loadFromFile, gbrl.cpp:1175-1252) */
int gbrl_b200_set_iteration(gbrl_b200_model *m, int iteration);

/* ---- building blocks exposed for tests / profiling (same kernels the calls above use) ---- */
/* thresholds[f*n_bins + b] of the last step/fit (fitter.cpp:77-90 candidates), host copy */
int gbrl_b200_get_candidates(gbrl_b200_model *m, float *thresholds, int *n_candidates);
/* per-candidate scores of the root node of the last grown tree (exact-arithmetic path), host copy */
int gbrl_b200_get_root_scores(gbrl_b200_model *m, float *scores, int *n_candidates);

/* ---- multi-GPU: feature-block sharded histograms + one NCCL all-reduce per level ---- */
int gbrl_b200_dist_unique_id(uint8_t id[128]);
int gbrl_b200_dist_init(gbrl_b200_model *m, const uint8_t id[128], int rank, int world_size);
int gbrl_b200_dist_shutdown(gbrl_b200_model *m);

/* per-kernel-class device timing with CUDA events on the launching stream (no reference equivalent).
 * classes: 0 candidates, 1 binning, 2 gradient preprocessing, 3 histogram, 4 all-reduce, 5 split scan,
 * 6 select+replay, 7 plan+decide, 8 partition, 9 leaf emission/values, 10 predict.  ms[i] / launches[i]
 * accumulate since gbrl_b200_profile(m, 1); hist_rows = rows scanned by the histogram kernel since creation. */
int gbrl_b200_profile(gbrl_b200_model *m, int enable);
int gbrl_b200_get_profile(gbrl_b200_model *m, double *ms, long long *launches, int n, long long *hist_rows);

/* microbenchmarks used by DESIGN.md's kernel budgets (not part of the reference surface) */
int gbrl_b200_microbench(int which, int iters, double *result);

/* test hook for the bit-exact parallel float-chain evaluator (csrc/chain.cuh): the reference's thread-partitioned
 * column sums (math_ops.cpp:255-300, mode 0) or centred sums of squares (math_ops.cpp:461-513, mode 1) of a HOST
 * matrix [n_elements / D x D] for T emulated reference threads; partial[T*D] and (mode 1) the centred matrix are
 * copied back.  impl 0 = the parallel evaluator, impl 1 = the plain sequential chain kernel.  info (optional, 4 doubles):
 * kernel ms, sub-blocks applied from their summary, sub-blocks advanced piecewise, lanes run sequentially. */
int gbrl_b200_diag_chain_sums(const float *mat, long long n_elements, int D, int T, int mode, const float *mean,
                              float *partial, float *centered, int impl, double *info);

#ifdef __cplusplus
}
#endif
#endif /* GBRL_B200_H */
