// replay.cuh -- parameter blocks shared by the near-tie replay kernels (split.cu, replay_wide.cu).
#pragma once
#include "engine.cuh"

namespace gb {

struct ReplayParams {
    int F, B, D, score_func, min_data;
    const float *X;            // raw features, row-major
    const uint16_t *codesT;    // feature-major codes [F][codesT_stride] (x > thr[f][j] <=> code > j), rows offset by row_offset
    long long codesT_stride;
    int row_offset;
    const float *bg;           // build_grads
    const int *order;
    const float *thr;
    const ReplayItem *items;
    float *out;
    const Ctl *ctl;
};

struct StreamParams {
    float *G;                  // [N x D] build_grads in `order` space
    unsigned int *bits;        // side-bit planes
    int *woff;                 // [n_items + 1] word offset of the item's plane (prefix; parents / direct items have 0 words)
    int *mode;                 // [n_items] 0 = streamed, 1 = parent (no plane), 2 = direct (plane did not fit)
    int *nright;               // [n_items]
    long long cap_words;
    int replay_cap, N, oblivious;
    const int *nid;
    int wide;                  // 1: parents get a (zero) plane too and every streamed item is mode 0 (replay_wide.cu)
};

// GPU-wide replay (replay_wide.cu): per 256-row group of every item
struct WideParams {
    double *bsum;              // [groups][2D] exact-ish sum of the group's elements per chain (side * D + d)
    float *pred;               // [groups][2D] predicted running sum of the chain at the start of the group
    int4 *tab;                 // [groups][2D] summary of the group for the predicted binade (chain.cuh Tab)
    float *tag;                // [groups][2D] the binade (inv_u) the summary was computed for, 0 = none
    int *gitem;                // [groups] item of the group
    float *fin;                // [items][8] pass 0 -> pass 1: means (left D, right D), ln, rn
    long long cap_groups;
};

void launch_replay_wide(Model &m, const ReplayParams &R, const StreamParams &S, cudaStream_t s);

}  // namespace gb
