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
    const int *pnode;          // node of the row at each position of `order`
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
    int4 *wtab;                // [windows][2D] composite table of a window of 32 groups (replay_wide.cu wide_wtabs_body)
    float *wtag;               // [windows][2D] its binade, TAG_EMPTY, or 0 = no composite
    long long cap_groups;
};

void launch_replay_wide(Model &m, const ReplayParams &R, const StreamParams &S, cudaStream_t s);

#if defined(__CUDACC__)
// word offset of every item's side-bit plane (8-word groups); items that do not fit -> direct.  One CTA of BT threads.
template <int BT>
__device__ __forceinline__ void replay_plan_body(const ReplayParams &P, NodeArrays na, const StreamParams &S) {
    __shared__ int s_scan[BT];
    __shared__ long long s_carry;
    const int n_items = min(P.ctl->n_replay, S.replay_cap);
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n_items; i0 += BT) {
        const int it = i0 + threadIdx.x;
        int words = 0, md = 1;
        if (it < n_items) {
            const ReplayItem item = P.items[it];
            if (item.cand >= 0 || S.wide) { words = ((na.seg_len[item.node] + 255) >> 8) << 3; md = 0; }   // 8-word groups
            if (S.wide) words = (words + 255) & ~255;     // planes start on 32-group boundaries (window tables, replay_wide.cu)
        }
        s_scan[threadIdx.x] = words;
        __syncthreads();
        for (int o = 1; o < BT; o <<= 1) {
            const int v = threadIdx.x >= o ? s_scan[threadIdx.x - o] : 0;
            __syncthreads();
            s_scan[threadIdx.x] += v;
            __syncthreads();
        }
        const long long carry = s_carry;
        const long long excl = carry + s_scan[threadIdx.x] - words;
        if (it < n_items) {
            if (md == 0 && excl + words > S.cap_words) md = 2;
            // direct items keep their (unused) slot in the prefix so that the prefix stays monotone
            S.woff[it] = (int)min(excl, S.cap_words);
            S.mode[it] = md;
            S.nright[it] = 0;
        }
        __syncthreads();
        if (threadIdx.x == BT - 1) s_carry = carry + s_scan[BT - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) S.woff[n_items] = (int)min(s_carry, S.cap_words);
}

// order-space copy of build_grads for the rows of every node that has replay items (grid-stride)
__device__ __forceinline__ void replay_gather_body(const ReplayParams &P, NodeArrays na, const StreamParams &S) {
    if (P.ctl->n_replay <= 0) return;
    const int D = P.D;
    const bool all = S.oblivious != 0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < S.N; k += gridDim.x * blockDim.x) {
        const int row = P.order[k];
        if (!all && na.rep_count[S.pnode[k]] <= 0) continue;
        for (int d = 0; d < D; ++d) S.G[(size_t)k * D + d] = P.bg[(size_t)row * D + d];
    }
}
#endif


}  // namespace gb
