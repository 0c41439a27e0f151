// b200_backend.h -- the forwarding class a gbrl maintainer would add next to gbrl/src/cpp/gbrl.h to put libgbrl_b200.so
// under the reference's own pybind11 binding.  It uses the reference's types (types.h:
// dataHolder<T>, scoreFunc, generatorType, growPolicy, schedulerFunc, deviceType) and nothing of this repo but the C-ABI
// header.  oracle/Makefile target `backend_check` compiles it against the reference's headers where they are mounted and
// links it against gbrl_b200/lib/libgbrl_b200.so, so that this file is code, not prose.
#pragma once
#include <omp.h>
#include <stdexcept>

#include "gbrl_b200.h"
#include "types.h"       // the reference's gbrl/src/cpp/types.h

class GBRL_B200 {
    gbrl_b200_model *h_ = nullptr;
    static void ck(int rc) { if (rc) throw std::runtime_error(gbrl_b200_last_error()); }   // -> Python RuntimeError
 public:
    GBRL_B200(int input_dim, int output_dim, int policy_dim, int max_depth, int min_data_in_leaf, int n_bins, int par_th,
              float /*cv_beta*/, scoreFunc score, generatorType gen, bool use_cv, int batch_size, growPolicy grow, int verbose,
              deviceType device) {
        if (device != gpu || use_cv) throw std::runtime_error("B200 backend: CUDA only, no control variates");
        gbrl_b200_config c{};                                  // gbrl.cpp:76-114 arguments, one to one
        c.input_dim = input_dim; c.output_dim = output_dim; c.policy_dim = policy_dim; c.max_depth = max_depth;
        c.min_data_in_leaf = min_data_in_leaf; c.n_bins = n_bins; c.par_th = par_th; c.batch_size = batch_size;
        c.split_score_func = (score == L2) ? GBRL_B200_SCORE_L2 : GBRL_B200_SCORE_COSINE;
        c.generator_type = (gen == Uniform) ? GBRL_B200_GEN_UNIFORM : GBRL_B200_GEN_QUANTILE;
        c.grow_policy = (grow == GREEDY) ? GBRL_B200_GROW_GREEDY : GBRL_B200_GROW_OBLIVIOUS;
        c.verbose = verbose; c.device_ordinal = 0;
        c.ref_threads = omp_get_max_threads();                 // reproduce this host's reduction partition
        c.tie_replay = 1; c.band_kappa = 0.f; c.use_subtraction = 1;   // engine defaults (near-tie replay on, band = 6 noise units)
        c.hist_variant = 0; c.replay_variant = 0;                      // streaming histogram kernel; GPU-wide replay chains on side streams (speculative levels)
        ck(gbrl_b200_create(&c, &h_));
    }
    ~GBRL_B200() { gbrl_b200_destroy(h_); }
    GBRL_B200(const GBRL_B200 &) = delete;
    GBRL_B200 &operator=(const GBRL_B200 &) = delete;
    // GBRL::step (gbrl.cpp:939)            dataHolder<T>{data, device} -> (pointer, device flag)
    void step(dataHolder<const float> *obs, dataHolder<float> *grads, int n, int n_num, void *stream) {
        ck(gbrl_b200_step(h_, obs->data, obs->device == gpu, grads->data, grads->device == gpu, n, n_num, stream));
    }
    // GBRL::fit (gbrl.cpp:983)
    float fit(dataHolder<const float> *obs, dataHolder<const float> *targets, int iters, int n, int n_num, bool shuffle, void *stream) {
        float loss = 0.f;
        ck(gbrl_b200_fit(h_, obs->data, obs->device == gpu, targets->data, targets->device == gpu, iters, n, n_num, shuffle, &loss, stream));
        return loss;
    }
    // GBRL::predict (gbrl.cpp:369): the caller allocates n * output_dim floats (cudaMalloc for the DLPack path, binding.cpp:208-261)
    void predict(dataHolder<const float> *obs, int n, int n_num, int start, int stop, float *out, bool out_on_gpu, void *stream) {
        ck(gbrl_b200_predict(h_, obs->data, obs->device == gpu, n, n_num, start, stop, out, out_on_gpu, stream));
    }
    void set_bias(dataHolder<const float> *b, int n) { ck(gbrl_b200_set_bias(h_, b->data, n, b->device == gpu)); }                 // gbrl.cpp:213
    void set_feature_weights(dataHolder<const float> *w, int n) { ck(gbrl_b200_set_feature_weights(h_, w->data, n, w->device == gpu)); }   // gbrl.cpp:241
    void set_feature_mapping(const int *m, const bool *num, int n) {                                                              // gbrl.cpp:269
        ck(gbrl_b200_set_feature_mapping(h_, m, reinterpret_cast<const uint8_t *>(num), n));
    }
    void set_optimizer(schedulerFunc f, float lr, int a, int b, float stop_lr, int T) {                                          // gbrl.cpp:452
        ck(gbrl_b200_set_optimizer(h_, f == Const ? GBRL_B200_SCHED_CONST : GBRL_B200_SCHED_LINEAR, lr, a, b, stop_lr, T));
    }
    int get_num_trees() { gbrl_b200_metadata md; ck(gbrl_b200_get_metadata(h_, &md)); return md.n_trees; }
    int get_iteration() { gbrl_b200_metadata md; ck(gbrl_b200_get_metadata(h_, &md)); return md.iteration; }
};
