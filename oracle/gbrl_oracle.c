/*
 * gbrl_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference (NVlabs/gbrl v1.1.6) CPU fit/predict hot path.  It is the
 * checker for the CUDA engine: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load it.  The product path (gbrl_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file bit-for-bit against the
 * reference itself (oracle/_ref, compiled from /root/reference by oracle/Makefile) and against the
 * committed golden vectors under tests/golden/ that were generated from the reference.
 *
 * Every function cites the reference file:line it restates (paths relative to gbrl/src/cpp/).
 * Arithmetic contract: IEEE fp32, source-order evaluation, no FMA contraction (-ffp-contract=off);
 * the reference build in oracle/Makefile uses the same contract.
 *
 * Thread emulation: several reference reductions split the element range over
 * T = calculate_num_threads(n_elements, par_th) OpenMP threads (utils.h:64-81) and merge partials in
 * thread order, so their float bits depend on T.  `ref_threads` is the emulated omp_get_max_threads().
 * The arithmetic here is executed serially in exactly that partial/merge order; OpenMP in this file
 * is used only across candidates (independent) to make the oracle finish sooner.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { SCORE_L2 = 0, SCORE_COSINE = 1 };
enum { GEN_UNIFORM = 0, GEN_QUANTILE = 1 };
enum { GROW_GREEDY = 0, GROW_OBLIVIOUS = 1 };
enum { SCHED_CONST = 0, SCHED_LINEAR = 1 };
#define MAX_OPTS 64

typedef struct {
    int start_idx, stop_idx, sched, T;
    float init_lr, stop_lr;
} OracleOpt;

typedef struct {
    int input_dim, output_dim, max_depth, min_data_in_leaf, n_bins, par_th, batch_size;
    int split_score_func, generator_type, grow_policy, ref_threads;
    int n_trees, n_leaves, iteration;
    int cap_trees, cap_leaves;
    float *bias, *feature_weights;
    int *reverse_num_feature_mapping;
    /* ensemble SoA, types.h:279-304 */
    int *tree_indices, *depths, *feature_indices;
    float *values, *feature_values, *edge_weights;
    uint8_t *inequality_directions;
    int n_opts;
    OracleOpt opts[MAX_OPTS];
} Oracle;

typedef struct { int feature_idx; float feature_value; } Cand;
typedef struct { int feature_idx; float feature_value; uint8_t dir; float edge_weight; } Cond;

/* utils.h:64-81 calculate_num_threads with an explicit max_threads */
static int calc_threads(int total, int min_per_thread, int max_threads) {
    int n = total / min_per_thread;
    if (n > total) n = total;
    if (n <= 1) return 1;
    if (n > max_threads) return max_threads;
    return n;
}

/* ---------------------------------------------------------------- model lifetime */
Oracle *oracle_create(int input_dim, int output_dim, int max_depth, int min_data_in_leaf, int n_bins,
                      int par_th, int batch_size, int split_score_func, int generator_type,
                      int grow_policy, int ref_threads) {
    Oracle *o = (Oracle *)calloc(1, sizeof(Oracle));
    o->input_dim = input_dim; o->output_dim = output_dim; o->max_depth = max_depth;
    o->min_data_in_leaf = min_data_in_leaf; o->n_bins = n_bins; o->par_th = par_th;
    o->batch_size = batch_size; o->split_score_func = split_score_func;
    o->generator_type = generator_type; o->grow_policy = grow_policy;
    o->ref_threads = ref_threads < 1 ? 1 : ref_threads;
    o->bias = (float *)calloc(output_dim, sizeof(float));
    o->feature_weights = (float *)calloc(input_dim, sizeof(float));
    /* types.cpp:232-234: zero-initialised until set_feature_mapping is called */
    o->reverse_num_feature_mapping = (int *)calloc(input_dim, sizeof(int));
    return o;
}

static void ensure_capacity(Oracle *o, int extra_trees, int extra_leaves) {
    int d = o->max_depth, D = o->output_dim;
    if (o->n_trees + extra_trees > o->cap_trees) {
        int nc = (o->n_trees + extra_trees) * 2 + 16;
        o->tree_indices = (int *)realloc(o->tree_indices, nc * sizeof(int));
        if (o->grow_policy == GROW_OBLIVIOUS) {
            o->depths = (int *)realloc(o->depths, nc * sizeof(int));
            o->feature_indices = (int *)realloc(o->feature_indices, (size_t)nc * d * sizeof(int));
            o->feature_values = (float *)realloc(o->feature_values, (size_t)nc * d * sizeof(float));
            for (int i = o->cap_trees; i < nc; ++i) o->depths[i] = 0;
            memset(o->feature_indices + (size_t)o->cap_trees * d, 0, (size_t)(nc - o->cap_trees) * d * sizeof(int));
            memset(o->feature_values + (size_t)o->cap_trees * d, 0, (size_t)(nc - o->cap_trees) * d * sizeof(float));
        }
        o->cap_trees = nc;
    }
    if (o->n_leaves + extra_leaves > o->cap_leaves) {
        int nc = (o->n_leaves + extra_leaves) * 2 + 64;
        int oc = o->cap_leaves;
        o->values = (float *)realloc(o->values, (size_t)nc * D * sizeof(float));
        o->edge_weights = (float *)realloc(o->edge_weights, (size_t)nc * d * sizeof(float));
        o->inequality_directions = (uint8_t *)realloc(o->inequality_directions, (size_t)nc * d);
        memset(o->values + (size_t)oc * D, 0, (size_t)(nc - oc) * D * sizeof(float));
        memset(o->edge_weights + (size_t)oc * d, 0, (size_t)(nc - oc) * d * sizeof(float));
        memset(o->inequality_directions + (size_t)oc * d, 0, (size_t)(nc - oc) * d);
        if (o->grow_policy == GROW_GREEDY) {
            o->depths = (int *)realloc(o->depths, nc * sizeof(int));
            o->feature_indices = (int *)realloc(o->feature_indices, (size_t)nc * d * sizeof(int));
            o->feature_values = (float *)realloc(o->feature_values, (size_t)nc * d * sizeof(float));
            for (int i = oc; i < nc; ++i) o->depths[i] = 0;
            memset(o->feature_indices + (size_t)oc * d, 0, (size_t)(nc - oc) * d * sizeof(int));
            memset(o->feature_values + (size_t)oc * d, 0, (size_t)(nc - oc) * d * sizeof(float));
        }
        o->cap_leaves = nc;
    }
}

void oracle_destroy(Oracle *o) {
    if (!o) return;
    free(o->bias); free(o->feature_weights); free(o->reverse_num_feature_mapping);
    free(o->tree_indices); free(o->depths); free(o->feature_indices); free(o->values);
    free(o->feature_values); free(o->edge_weights); free(o->inequality_directions);
    free(o);
}

void oracle_set_bias(Oracle *o, const float *b) { memcpy(o->bias, b, o->output_dim * sizeof(float)); }
void oracle_set_feature_weights(Oracle *o, const float *w) { memcpy(o->feature_weights, w, o->input_dim * sizeof(float)); }

/* gbrl.cpp:271-316 set_feature_mapping (numerical part) */
void oracle_set_feature_mapping(Oracle *o, const int *feature_mapping, const uint8_t *mapping_numerics) {
    (void)feature_mapping;
    int j = 0;
    for (int i = 0; i < o->input_dim; ++i) o->reverse_num_feature_mapping[i] = -1;
    for (int i = 0; i < o->input_dim; ++i)
        if (mapping_numerics[i]) o->reverse_num_feature_mapping[j++] = i;
}

/* gbrl.cpp:452-525 set_optimizer (SGD only on this path) */
int oracle_set_optimizer(Oracle *o, int sched, float init_lr, int start_idx, int stop_idx, float stop_lr, int T) {
    if (o->n_opts >= o->output_dim || o->n_opts >= MAX_OPTS) return -1;
    if (start_idx >= stop_idx || start_idx < 0 || stop_idx > o->output_dim) return -2;
    OracleOpt *p = &o->opts[o->n_opts++];
    p->start_idx = start_idx; p->stop_idx = stop_idx; p->sched = sched; p->init_lr = init_lr;
    p->stop_lr = stop_lr; p->T = T;
    return 0;
}

int oracle_n_trees(const Oracle *o) { return o->n_trees; }
int oracle_n_leaves(const Oracle *o) { return o->n_leaves; }
int oracle_iteration(const Oracle *o) { return o->iteration; }

/* copy the ensemble out in the layout of binding.cpp:330-390 get_ensemble_data */
void oracle_get_ensemble(const Oracle *o, int *tree_indices, int *depths, float *values, int *feature_indices,
                         float *feature_values, float *edge_weights, uint8_t *inequality_directions) {
    int d = o->max_depth, D = o->output_dim;
    int S = o->grow_policy == GROW_OBLIVIOUS ? o->n_trees : o->n_leaves;
    memcpy(tree_indices, o->tree_indices, o->n_trees * sizeof(int));
    memcpy(depths, o->depths, S * sizeof(int));
    memcpy(values, o->values, (size_t)o->n_leaves * D * sizeof(float));
    memcpy(feature_indices, o->feature_indices, (size_t)S * d * sizeof(int));
    memcpy(feature_values, o->feature_values, (size_t)S * d * sizeof(float));
    memcpy(edge_weights, o->edge_weights, (size_t)o->n_leaves * d * sizeof(float));
    memcpy(inequality_directions, o->inequality_directions, (size_t)o->n_leaves * d);
}

/* scheduler.h:124-135 (Linear), :182-185 (Const) */
static float get_lr(const OracleOpt *p, int t) {
    if (p->sched == SCHED_CONST) return p->init_lr;
    float T_ = (float)p->T;
    float t_ = (float)t + 1;
    float progress_remaining = (T_ - t_) / T_;
    float lr = p->init_lr + (1.0f - progress_remaining) * (p->stop_lr - p->init_lr);
    if (lr < p->stop_lr) return p->stop_lr;
    return lr;
}

/* ---------------------------------------------------------------- gradient preprocessing */
/* math_ops.cpp:255-300 calculate_mean: per-thread partial sums over contiguous ELEMENT ranges of the
 * row-major matrix, merged in thread order, then * (1/n_samples). */
static void ref_mean(const float *mat, int n_samples, int n_cols, int par_th, int max_threads, float *mean) {
    int n_elements = n_samples * n_cols;
    float recip = 1.0f / (float)n_samples;
    for (int d = 0; d < n_cols; ++d) mean[d] = 0.0f;
    int T = calc_threads(n_elements, par_th, max_threads);
    if (T > 1) {
        int ept = n_elements / T;
        float *tm = (float *)calloc((size_t)T * n_cols, sizeof(float));
        for (int t = 0; t < T; ++t) {
            int s = t * ept, e = (t == T - 1) ? n_elements : s + ept;
            for (int i = s; i < e; ++i) tm[t * n_cols + i % n_cols] += mat[i];
        }
        for (int d = 0; d < T * n_cols; ++d) mean[d % n_cols] += tm[d];
        free(tm);
    } else {
        for (int i = 0; i < n_elements; ++i) mean[i % n_cols] += mat[i];
    }
    for (int d = 0; d < n_cols; ++d) mean[d] *= recip;
}

/* math_ops.cpp:407-459 calculate_var_and_center / :461-513 calculate_std_and_center: centers `mat`
 * in place and returns sqrt(sum((x-mean)^2) * 1/(n-1)) per column (the two reference variants apply
 * the same float ops: var*recip then sqrtf). */
static void ref_std_and_center(float *mat, const float *mean, int n_samples, int n_cols, int par_th,
                               int max_threads, float *std) {
    int n_elements = n_samples * n_cols;
    float recip = 1.0f / ((float)n_samples - 1.0f);
    for (int d = 0; d < n_cols; ++d) std[d] = 0.0f;
    int T = calc_threads(n_elements, par_th, max_threads);
    if (T > 1) {
        int ept = n_elements / T;
        float *tv = (float *)calloc((size_t)T * n_cols, sizeof(float));
        for (int t = 0; t < T; ++t) {
            int s = t * ept, e = (t == T - 1) ? n_elements : s + ept;
            for (int i = s; i < e; ++i) {
                int col = i % n_cols;
                float value = mat[i] - mean[col];
                tv[t * n_cols + col] += value * value;
                mat[i] -= mean[col];
            }
        }
        for (int d = 0; d < T * n_cols; ++d) std[d % n_cols] += tv[d];
        free(tv);
    } else {
        for (int i = 0; i < n_elements; ++i) {
            int col = i % n_cols;
            float value = mat[i] - mean[col];
            std[col] += value * value;
            mat[i] -= mean[col];
        }
    }
    for (int d = 0; d < n_cols; ++d) std[d] = sqrtf(std[d] * recip);
}

/* fitter.cpp:57-64 (step_cpu) and :204-214 (fit_cpu): build_grads for the L2 score;
 * math_ops.cpp:79-105 divide_mat_by_vec_inplace. Cosine uses the raw gradients. */
void oracle_build_grads(const Oracle *o, const float *grads, int n_samples, float *build_grads) {
    int D = o->output_dim;
    memcpy(build_grads, grads, (size_t)n_samples * D * sizeof(float));
    if (o->split_score_func != SCORE_L2) return;
    float *mean = (float *)malloc(D * sizeof(float)), *std = (float *)malloc(D * sizeof(float));
    ref_mean(build_grads, n_samples, D, o->par_th, o->ref_threads, mean);
    ref_std_and_center(build_grads, mean, n_samples, D, o->par_th, o->ref_threads, std);
    for (int i = 0; i < n_samples * D; ++i) build_grads[i] /= (std[i % D] + 1e-8f);
    free(mean); free(std);
}

/* ---------------------------------------------------------------- split candidates */
static int cmp_float(const void *a, const void *b) {
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

/* fitter.cpp:77-90 + split_candidate_generator.cpp:79-115,216-249 (Quantile; the dedup test at :241 reads
 * the member n_candidates which is still 0, so every one of the n_bins thresholds per feature is
 * emitted, duplicates included) and :59-76 (Uniform).  Candidate order: feature-major, bin-minor.
 * Only the VALUE at each sorted rank is used, so sorting values directly is equivalent to the
 * reference's index arg-sort. */
int oracle_candidates(const Oracle *o, const float *obs, int n_samples, int n_features, Cand *out) {
    int B = o->n_bins;
    if (o->generator_type == GEN_UNIFORM) {
        for (int f = 0; f < n_features; ++f) {
            float mx = -INFINITY, mn = INFINITY;
            for (int i = 0; i < n_samples; ++i) {
                float v = obs[(size_t)i * n_features + f];
                if (v > mx) mx = v;
                if (v < mn) mn = v;
            }
            float step = (mx - mn) / (float)B;
            for (int b = 0; b < B; ++b) {
                out[f * B + b].feature_idx = f;
                out[f * B + b].feature_value = mn + (float)b * step;
            }
        }
        return B * n_features;
    }
    int actual_bins = B + 1;
    int spb = n_samples / actual_bins, rem = n_samples % actual_bins;
    int *bin_counts = (int *)malloc(actual_bins * sizeof(int));
    for (int i = 0; i < actual_bins; ++i) bin_counts[i] = spb;
    while (rem > 0) {
        for (int i = 0; i < actual_bins; ++i) { bin_counts[i] += 1; rem -= 1; if (rem == 0) break; }
    }
#pragma omp parallel
    {
        float *col = (float *)malloc((size_t)n_samples * sizeof(float));
#pragma omp for
        for (int f = 0; f < n_features; ++f) {
            for (int i = 0; i < n_samples; ++i) col[i] = obs[(size_t)i * n_features + f];
            qsort(col, n_samples, sizeof(float), cmp_float);
            int cum = 0;
            for (int b = 0; b < B; ++b) {
                cum += bin_counts[b];
                out[f * B + b].feature_idx = f;
                /* cum-1 can be -1 when n_samples < n_bins+1; the reference then reads out of bounds
                 * (split_candidate_generator.cpp:237); we clamp and document it as undefined there. */
                out[f * B + b].feature_value = col[cum > 0 ? cum - 1 : 0];
            }
        }
        free(col);
    }
    free(bin_counts);
    return B * n_features;
}

/* ---------------------------------------------------------------- split scores */
/* math_ops.h:432-449 mat_vec_dot_sum */
static float mat_vec_dot_sum(const int *idx, const float *g, const float *vec, int n, int D) {
    float sum = 0.0f;
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < D; ++c) sum += g[(size_t)idx[r] * D + c] * vec[c];
    return sum;
}
/* math_ops.h:476-486 squared_norm */
static float squared_norm(const float *v, int n) {
    float s = 0.0f;
    for (int i = 0; i < n; ++i) s += v[i] * v[i];
    return s;
}

/* node.cpp:151-166 path-reuse guard of getSplitScore */
static int path_reuses(const Cond *path, int depth, const Cand *c) {
    for (int i = 0; i < depth; ++i)
        if (path[i].feature_value == c->feature_value && path[i].feature_idx == c->feature_idx) return 1;
    return 0;
}

/* node.cpp:321-376 splitScoreL2 / node.cpp:187-251 splitScoreCosine + math_ops.h:538-575 cosine_score.
 * scratch: 2*D floats + (cosine) 2*n ints. */
static float split_score(int func, const float *obs, const float *grads, const int *samples, int n, int F,
                         int D, const Cand *c, int min_data_in_leaf, float *lm, float *rm, int *li, int *ri) {
    int lc = 0, rc = 0;
    for (int d = 0; d < D; ++d) { lm[d] = 0; rm[d] = 0; }
    for (int k = 0; k < n; ++k) {
        int s = samples[k];
        const float *g = grads + (size_t)s * D;
        if (obs[(size_t)s * F + c->feature_idx] > c->feature_value) {
            for (int d = 0; d < D; ++d) rm[d] += g[d];
            if (ri) ri[rc] = s;
            ++rc;
        } else {
            for (int d = 0; d < D; ++d) lm[d] += g[d];
            if (li) li[lc] = s;
            ++lc;
        }
    }
    if (lc < min_data_in_leaf || rc < min_data_in_leaf) return -INFINITY;
    float lcf = (float)lc, rcf = (float)rc;
    float lrecip = (lc > 0) ? 1.0f / lcf : 0.0f;
    float rrecip = (rc > 0) ? 1.0f / rcf : 0.0f;
    for (int d = 0; d < D; ++d) { lm[d] *= lrecip; rm[d] *= rrecip; }
    if (func == SCORE_L2) {
        float ln = squared_norm(lm, D), rn = squared_norm(rm, D);
        return lcf * ln + rcf * rn;
    }
    /* cosine_score(true=right, false=left) */
    float tnum = 0.0f, fnum = 0.0f;
    if (rc > 0) tnum = mat_vec_dot_sum(ri, grads, rm, rc, D);
    if (lc > 0) fnum = mat_vec_dot_sum(li, grads, lm, lc, D);
    float tden = squared_norm(rm, D) * rcf;
    float fden = squared_norm(lm, D) * lcf;
    float num = tnum + fnum, den = tden + fden;
    if (den == 0.0f) return 0.0f;
    return num / sqrtf(den);
}

/* split_candidate_generator.cpp:293-320 scoreL2 / :262-290 scoreCosine + math_ops.h:500-520 cosine_dist */
static float parent_score(int func, const float *grads, const int *samples, int n, int D, float *mean) {
    float nf = (float)n;
    float recip = 1.0f / nf;
    for (int d = 0; d < D; ++d) mean[d] = 0.0f;
    for (int k = 0; k < n; ++k)
        for (int d = 0; d < D; ++d) mean[d] += grads[(size_t)samples[k] * D + d];
    for (int d = 0; d < D; ++d) mean[d] *= recip;
    if (func == SCORE_L2) return squared_norm(mean, D) * nf;
    if (n == 0) return 0.0f;
    float dot = mat_vec_dot_sum(samples, grads, mean, n, D);
    float den = squared_norm(mean, D) * nf;
    if (den == 0.0f) return 0.0f;
    /* math_ops.h:519 calls sqrt() on a float; with libstdc++ that resolves to the float overload */
    return dot / sqrtf(den);
}

/* ---------------------------------------------------------------- tree nodes */
typedef struct Node {
    int *samples; int n; int depth; Cond *path;
} Node;

static Node *node_new(int *samples, int n, int depth, int max_depth) {
    Node *nd = (Node *)malloc(sizeof(Node));
    nd->samples = samples; nd->n = n; nd->depth = depth;
    nd->path = (Cond *)calloc(max_depth > 0 ? max_depth : 1, sizeof(Cond));
    return nd;
}
static void node_free(Node *nd) { if (nd) { free(nd->samples); free(nd->path); free(nd); } }

/* node.cpp:64-149 splitNode: stable partition (x > thr -> right), child path = parent path + condition */
static void split_node(const Node *p, const float *obs, int F, const Cand *c, int max_depth, Node **left, Node **right) {
    int *pl = (int *)malloc((p->n > 0 ? p->n : 1) * sizeof(int)), *pr = (int *)malloc((p->n > 0 ? p->n : 1) * sizeof(int));
    int lc = 0, rc = 0;
    for (int k = 0; k < p->n; ++k) {
        int s = p->samples[k];
        if (obs[(size_t)s * F + c->feature_idx] > c->feature_value) pr[rc++] = s; else pl[lc++] = s;
    }
    Node *l = node_new(pl, lc, p->depth + 1, max_depth), *r = node_new(pr, rc, p->depth + 1, max_depth);
    memcpy(l->path, p->path, p->depth * sizeof(Cond));
    memcpy(r->path, p->path, p->depth * sizeof(Cond));
    Cond cl = { c->feature_idx, c->feature_value, 0, (p->n > 0) ? (float)lc / (float)p->n : 0.0f };
    Cond cr = { c->feature_idx, c->feature_value, 1, (p->n > 0) ? (float)rc / (float)p->n : 0.0f };
    l->path[p->depth] = cl; r->path[p->depth] = cr;
    *left = l; *right = r;
}

/* score one candidate for one node, incl. the guard (node.cpp:151-185) */
static float node_cand_score(const Oracle *o, const Node *nd, const float *obs, const float *bg, int F,
                             const Cand *c, float *lm, float *rm, int *li, int *ri) {
    if (nd->depth > 0 && path_reuses(nd->path, nd->depth, c)) return -INFINITY;
    return split_score(o->split_score_func, obs, bg, nd->samples, nd->n, F, o->output_dim, c,
                       o->min_data_in_leaf, lm, rm, li, ri);
}

/* fitter.cpp:545-582 calc_leaf_value: re-test ALL samples against the leaf path, mean of RAW grads */
static void calc_leaf_value(Oracle *o, const float *obs, const float *grads, int n_samples, int F, int leaf_idx, int tree_idx) {
    int D = o->output_dim, md = o->max_depth;
    int obl = o->grow_policy == GROW_OBLIVIOUS;
    int depth = obl ? o->depths[tree_idx] : o->depths[leaf_idx];
    int cond = obl ? tree_idx * md : leaf_idx * md;
    int ineq = leaf_idx * md;
    float count = 0;
    for (int i = 0; i < n_samples; ++i) {
        int passed = 0;
        for (int k = depth - 1; k >= 0; --k) {
            passed = (obs[(size_t)i * F + o->feature_indices[cond + k]] > o->feature_values[cond + k]) == (o->inequality_directions[ineq + k] != 0);
            if (!passed) break;
        }
        if (passed) {
            for (int d = 0; d < D; ++d) o->values[(size_t)leaf_idx * D + d] += grads[(size_t)i * D + d];
            count += 1;
        }
    }
    if (count > 0)
        for (int d = 0; d < D; ++d) o->values[(size_t)leaf_idx * D + d] /= count;
}

/* fitter.cpp:493-515 update_ensemble_per_leaf */
static void emit_leaf(Oracle *o, const Node *nd) {
    ensure_capacity(o, 0, 1);
    int idx = o->n_leaves, md = o->max_depth;
    o->depths[idx] = nd->depth;
    for (int i = 0; i < nd->depth; ++i) {
        o->feature_indices[idx * md + i] = nd->path[i].feature_idx;
        o->feature_values[idx * md + i] = nd->path[i].feature_value;
        o->inequality_directions[idx * md + i] = nd->path[i].dir;
        o->edge_weights[idx * md + i] = nd->path[i].edge_weight;
    }
    for (int d = 0; d < o->output_dim; ++d) o->values[(size_t)idx * o->output_dim + d] = 0.0f;
    o->n_leaves += 1;
}

/* fitter.cpp:263-375 fit_greedy_tree.  DFS stack, right pushed before left (left popped first);
 * gain = score*feature_weights[feature_idx] - parent_score (root parent forced to 0); strict '>' with
 * lowest candidate index winning ties (:338-353); split iff best >= 0 (:357). */
static int fit_greedy(Oracle *o, const float *obs, const float *bg, int n_samples, int F, const Cand *cands, int n_cands) {
    int D = o->output_dim, md = o->max_depth;
    ensure_capacity(o, 1, 0);
    o->tree_indices[o->n_trees] = o->n_leaves;
    int *root_idx = (int *)malloc((n_samples > 0 ? n_samples : 1) * sizeof(int));
    for (int i = 0; i < n_samples; ++i) root_idx[i] = i;
    int cap = 2 * md + 4, sp = 0;
    Node **stack = (Node **)malloc(cap * sizeof(Node *));
    stack[sp++] = node_new(root_idx, n_samples, 0, md);
    float *scores = (float *)malloc((n_cands > 0 ? n_cands : 1) * sizeof(float));
    int added = 0;
    while (sp > 0) {
        Node *nd = stack[--sp];
        int to_split = !(nd->depth == md || nd->n == 0 || n_cands == 0);
        float best = -INFINITY; int chosen = 0;
        if (to_split) {
            float *pm = (float *)malloc(D * sizeof(float));
            float parent = parent_score(o->split_score_func, bg, nd->samples, nd->n, D, pm);
            free(pm);
            if (nd->depth == 0) parent = 0.0f;
#pragma omp parallel
            {
                float *lm = (float *)malloc(2 * D * sizeof(float)), *rm = lm + D;
                int *li = NULL, *ri = NULL;
                if (o->split_score_func == SCORE_COSINE) { li = (int *)malloc((nd->n + 1) * sizeof(int)); ri = (int *)malloc((nd->n + 1) * sizeof(int)); }
#pragma omp for schedule(dynamic, 64)
                for (int j = 0; j < n_cands; ++j) {
                    float s = node_cand_score(o, nd, obs, bg, F, &cands[j], lm, rm, li, ri);
                    scores[j] = s * o->feature_weights[cands[j].feature_idx] - parent;
                }
                free(lm); free(li); free(ri);
            }
            for (int j = 0; j < n_cands; ++j) if (scores[j] > best) { best = scores[j]; chosen = j; }
        }
        if (best >= 0 && to_split) {
            Node *l, *r;
            split_node(nd, obs, F, &cands[chosen], md, &l, &r);
            stack[sp++] = r; stack[sp++] = l;
        } else {
            emit_leaf(o, nd);
            added += 1;
        }
        node_free(nd);
    }
    free(stack); free(scores);
    o->n_trees += 1;
    return added;
}

/* fitter.cpp:377-484 fit_oblivious_tree + :517-542 update_ensemble_per_tree.  Per depth:
 * score_c = (sum over nodes, in node order, of node score) * feature_weights[reverse_num_map[f]];
 * stop only when the best is -inf; all 2^depth nodes split on the winner; children [2i]=left,[2i+1]=right. */
static int fit_oblivious(Oracle *o, const float *obs, const float *bg, int n_samples, int F, const Cand *cands, int n_cands) {
    int D = o->output_dim, md = o->max_depth;
    ensure_capacity(o, 1, 1 << md);
    o->tree_indices[o->n_trees] = o->n_leaves;
    int maxl = 1 << md;
    Node **nodes = (Node **)calloc(maxl, sizeof(Node *)), **child = (Node **)calloc(maxl, sizeof(Node *));
    int *root_idx = (int *)malloc((n_samples > 0 ? n_samples : 1) * sizeof(int));
    for (int i = 0; i < n_samples; ++i) root_idx[i] = i;
    nodes[0] = node_new(root_idx, n_samples, 0, md);
    float *scores = (float *)malloc((n_cands > 0 ? n_cands : 1) * sizeof(float));
    int depth = 0;
    while (depth < md) {
        float best = -INFINITY; int chosen = 0;
        int nn = 1 << depth;
#pragma omp parallel
        {
            float *lm = (float *)malloc(2 * D * sizeof(float)), *rm = lm + D;
            int *li = NULL, *ri = NULL;
            if (o->split_score_func == SCORE_COSINE) { li = (int *)malloc((n_samples + 1) * sizeof(int)); ri = (int *)malloc((n_samples + 1) * sizeof(int)); }
#pragma omp for schedule(dynamic, 64)
            for (int j = 0; j < n_cands; ++j) {
                float s = 0.0f;
                for (int k = 0; k < nn; ++k) s += node_cand_score(o, nodes[k], obs, bg, F, &cands[j], lm, rm, li, ri);
                scores[j] = s * o->feature_weights[o->reverse_num_feature_mapping[cands[j].feature_idx]];
            }
            free(lm); free(li); free(ri);
        }
        for (int j = 0; j < n_cands; ++j) if (scores[j] > best) { best = scores[j]; chosen = j; }
        if (best == -INFINITY) break;
        for (int k = 0; k < nn; ++k) {
            split_node(nodes[k], obs, F, &cands[chosen], md, &child[2 * k], &child[2 * k + 1]);
            node_free(nodes[k]);
        }
        depth += 1;
        for (int k = 0; k < (1 << depth); ++k) { nodes[k] = child[k]; child[k] = NULL; }
    }
    /* update_ensemble_per_tree: split arrays per TREE, directions/edge weights per LEAF */
    int t = o->n_trees, nl = 1 << depth;
    for (int k = 0; k < nl; ++k) {
        Node *nd = nodes[k];
        o->depths[t] = nd->depth;
        for (int i = 0; i < nd->depth; ++i) {
            o->feature_indices[t * md + i] = nd->path[i].feature_idx;
            o->feature_values[t * md + i] = nd->path[i].feature_value;
            o->inequality_directions[o->n_leaves * md + i] = nd->path[i].dir;
            o->edge_weights[o->n_leaves * md + i] = nd->path[i].edge_weight;
        }
        for (int d = 0; d < D; ++d) o->values[(size_t)o->n_leaves * D + d] = 0.0f;
        o->n_leaves += 1;
        node_free(nd);
    }
    free(nodes); free(child); free(scores);
    o->n_trees += 1;
    return nl;
}

/* shared by step and fit: grow one tree on (obs, raw grads, build_grads) and fill its leaves
 * (fitter.cpp:98-102 / :220-225, fit_leaves :487-491) */
static void grow_tree(Oracle *o, const float *obs, const float *grads, const float *bg, int n, int F, const Cand *cands, int nc) {
    int added = (o->grow_policy == GROW_GREEDY) ? fit_greedy(o, obs, bg, n, F, cands, nc) : fit_oblivious(o, obs, bg, n, F, cands, nc);
    int t = o->n_trees - 1;
    for (int l = 0; l < added; ++l) calc_leaf_value(o, obs, grads, n, F, o->tree_indices[t] + l, t);
}

/* fitter.cpp:50-115 step_cpu (numerical features, no control variates) */
int oracle_step(Oracle *o, const float *obs, const float *grads, int n_samples, int n_features) {
    int D = o->output_dim;
    float *bg = (float *)malloc((size_t)(n_samples > 0 ? n_samples : 1) * D * sizeof(float));
    oracle_build_grads(o, grads, n_samples, bg);
    Cand *cands = (Cand *)malloc((size_t)o->n_bins * n_features * sizeof(Cand));
    int nc = oracle_candidates(o, obs, n_samples, n_features, cands);
    grow_tree(o, obs, grads, bg, n_samples, n_features, cands, nc);
    free(bg); free(cands);
    o->iteration++;
    return 0;
}

/* ---------------------------------------------------------------- predict */
/* predictor.cpp:188-229 predict_over_leaves / :231-265 predict_over_trees + optimizer.cpp:110-118 */
static void predict_sample(const Oracle *o, const float *x, float *theta, int start_tree, int stop_tree) {
    int md = o->max_depth, D = o->output_dim;
    if (o->grow_policy == GROW_OBLIVIOUS) {
        for (int t = start_tree; t < stop_tree; ++t) {
            int cond = t * md, leaf = 0, dep = o->depths[t];
            for (int k = 0; k < dep; ++k) {
                int passed = x[o->feature_indices[cond + k]] > o->feature_values[cond + k];
                leaf |= passed << (dep - 1 - k);
            }
            const float *v = o->values + (size_t)(o->tree_indices[t] + leaf) * D;
            for (int p = 0; p < o->n_opts; ++p) {
                float lr = get_lr(&o->opts[p], t);
                for (int i = o->opts[p].start_idx; i < o->opts[p].stop_idx; ++i) theta[i] -= lr * v[i];
            }
        }
        return;
    }
    int t = start_tree;
    if (t >= stop_tree) return;
    int leaf = o->tree_indices[t];
    while (leaf < o->n_leaves && t < stop_tree) {
        int dep = o->depths[leaf], cond = leaf * md, passed = 0;
        for (int k = dep - 1; k >= 0; --k) {
            passed = (x[o->feature_indices[cond + k]] > o->feature_values[cond + k]) == (o->inequality_directions[cond + k] != 0);
            if (!passed) break;
        }
        if (passed) {
            const float *v = o->values + (size_t)leaf * D;
            for (int p = 0; p < o->n_opts; ++p) {
                float lr = get_lr(&o->opts[p], t);
                for (int i = o->opts[p].start_idx; i < o->opts[p].stop_idx; ++i) theta[i] -= lr * v[i];
            }
            ++t;
            if (t < stop_tree) leaf = o->tree_indices[t];
        } else {
            ++leaf;
        }
    }
}

/* predictor.cpp:122-185 predict_cpu, sample-parallel / serial form (per sample, sequential over trees).
 * preds must be zero-initialised by the caller (gbrl.cpp:418). */
int oracle_predict(const Oracle *o, const float *obs, int n_samples, int n_features, int start_tree, int stop_tree, float *preds) {
    int D = o->output_dim;
    for (int i = 0; i < n_samples; ++i)
        for (int d = 0; d < D; ++d) preds[(size_t)i * D + d] += o->bias[d];
    if (o->n_trees == 0) return 0;
    if (stop_tree > o->n_trees) return -1;
    if (stop_tree == 0) stop_tree = o->n_trees;
    if (o->n_opts == 0) return -2;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n_samples; ++i)
        predict_sample(o, obs + (size_t)i * n_features, preds + (size_t)i * D, start_tree, stop_tree);
    return 0;
}

/* ---------------------------------------------------------------- fit (supervised loop) */
/* loss.cpp:34-62 MultiRMSE::get_loss_and_gradients.  NOTE the reference quirk: each of the T threads
 * handles exactly n_elements/T elements, so the trailing n_elements%T gradients are never written
 * (they keep the previous buffer contents). */
static float multirmse(const Oracle *o, const float *preds, const float *targets, float *grads, int n_samples, int write_grads) {
    int D = o->output_dim, n_elements = n_samples * D;
    float recip = 1.0f / (float)n_samples;
    int T = calc_threads(n_elements, o->par_th, o->ref_threads);
    int ept = n_elements / T;
    float loss = 0.0f;
    for (int t = 0; t < T; ++t) {
        int s = t * ept, e = (s + ept > n_elements) ? n_elements : s + ept;
        float lt = 0.0f;
        for (int i = s; i < e; ++i) {
            float gv = preds[i] - targets[i];
            if (write_grads) grads[i] = gv;
            lt += gv * gv;
        }
        loss += lt;
    }
    return sqrtf(0.5f * loss * recip);
}

/* gbrl.cpp:983-1104 fit (shuffle=False) + fitter.cpp:117-261 fit_cpu: bias := column mean of targets,
 * candidates once on the full data, then per iteration a cyclic mini-batch of batch_size rows:
 * predict(trees 0..i) -> MultiRMSE grads -> (L2) standardise -> grow -> leaves. */
float oracle_fit(Oracle *o, const float *obs, const float *targets, int iterations, int n_samples, int n_features) {
    int D = o->output_dim, F = n_features, bs = o->batch_size;
    float *bias = (float *)malloc(D * sizeof(float));
    ref_mean(targets, n_samples, D, o->par_th, o->ref_threads, bias);
    oracle_set_bias(o, bias);
    free(bias);
    Cand *cands = (Cand *)malloc((size_t)o->n_bins * F * sizeof(Cand));
    int nc = oracle_candidates(o, obs, n_samples, F, cands);
    int batch_start = 0;
    int batch_n = batch_start + bs < n_samples ? bs : n_samples - batch_start;
    int last_sz = (n_samples % bs) * D;
    float *bp = (float *)calloc((size_t)bs * D + 1, sizeof(float)), *bg_ = (float *)calloc((size_t)bs * D + 1, sizeof(float));
    float *bb = (float *)calloc((size_t)bs * D + 1, sizeof(float));
    float *lp = (float *)calloc(last_sz + 1, sizeof(float)), *lg = (float *)calloc(last_sz + 1, sizeof(float)), *lb = (float *)calloc(last_sz + 1, sizeof(float));
    for (int it = 0; it < iterations; ++it) {
        const float *bobs = obs + (size_t)batch_start * F;
        const float *btargets = targets + (size_t)batch_start * D;
        int is_last = batch_start + bs > n_samples;
        float *preds = is_last ? lp : bp, *grads = is_last ? lg : bg_, *build = is_last ? lb : bb;
        memset(preds, 0, (size_t)(is_last ? last_sz : bs * D) * sizeof(float));
        oracle_predict(o, bobs, batch_n, F, 0, it, preds);
        multirmse(o, preds, btargets, grads, batch_n, 1);
        oracle_build_grads(o, grads, batch_n, build);
        grow_tree(o, bobs, grads, build, batch_n, F, cands, nc);
        batch_start += batch_n;
        if (batch_start >= n_samples) batch_start = 0;
        batch_n = batch_start + bs < n_samples ? bs : n_samples - batch_start;
        o->iteration++;
    }
    float *full = (float *)calloc((size_t)n_samples * D, sizeof(float));
    oracle_predict(o, obs, n_samples, F, 0, iterations, full);
    float loss = multirmse(o, full, targets, NULL, n_samples, 0);
    free(full); free(cands); free(bp); free(bg_); free(bb); free(lp); free(lg); free(lb);
    return loss;
}

/* ---------------------------------------------------------------- diagnostics for band calibration */
/* root-node scores of every candidate, as the reference would compute them (used by the tests that
 * calibrate the CUDA engine's near-tie band; not part of the reference API) */
int oracle_root_scores(const Oracle *o, const float *obs, const float *grads, int n_samples, int n_features, float *scores_out, float *thresholds_out) {
    int D = o->output_dim;
    float *bg = (float *)malloc((size_t)n_samples * D * sizeof(float));
    oracle_build_grads(o, grads, n_samples, bg);
    Cand *cands = (Cand *)malloc((size_t)o->n_bins * n_features * sizeof(Cand));
    int nc = oracle_candidates(o, obs, n_samples, n_features, cands);
    int *idx = (int *)malloc((size_t)n_samples * sizeof(int));
    for (int i = 0; i < n_samples; ++i) idx[i] = i;
#pragma omp parallel
    {
        float *lm = (float *)malloc(2 * D * sizeof(float)), *rm = lm + D;
        int *li = (int *)malloc((n_samples + 1) * sizeof(int)), *ri = (int *)malloc((n_samples + 1) * sizeof(int));
#pragma omp for schedule(dynamic, 64)
        for (int j = 0; j < nc; ++j) {
            scores_out[j] = split_score(o->split_score_func, obs, bg, idx, n_samples, n_features, D, &cands[j], o->min_data_in_leaf, lm, rm, li, ri);
            thresholds_out[j] = cands[j].feature_value;
        }
        free(lm); free(li); free(ri);
    }
    free(bg); free(cands); free(idx);
    return nc;
}
