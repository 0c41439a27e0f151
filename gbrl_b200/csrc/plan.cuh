// plan.cuh -- level planning of the histogram pass (shared by histogram.cu and the fused decide + plan kernel of split.cu)
#pragma once
#include "engine.cuh"

namespace gb {

struct PlanParams {
    Item *items; int items_cap, max_depth, nT_local, use_subtraction, oblivious, row_groups, row_group, item_rows_max;
};

// ---------------------------------------------------------------- level planning
// Decides, for every node of the level, whether its histogram is built directly or derived as
// parent - sibling (only the smaller child is histogrammed), and emits the work items.
// (a block-level routine: every thread of a CTA of up to 1024 threads has to call it)
__device__ __forceinline__ void plan_level_body(NodeArrays na, Ctl *ctl, Item *items, int items_cap, int level, int max_depth,
                                                int nT_local, int use_subtraction, int oblivious, int row_groups, int row_group, int item_rows_max) {
    __shared__ int s_cnt[1024], s_off[1024], s_len[1024], s_start[1024];
    __shared__ int s_total;
    __shared__ unsigned long long s_rows;
    const int base = level_base(level), nn = 1 << level;
    __shared__ unsigned long long s_direct_rows;
    __shared__ int s_item_rows;
    if (threadIdx.x == 0) { s_total = 0; s_rows = 0; s_direct_rows = 0; }
    __syncthreads();
    // pass 0: rows that will be scanned at this level -> item size.  Items are at most ITEM_ROWS rows (int32 overflow
    // bound of the shared-memory partial sums) and shrink (down to 2048) when the level has too few rows to give
    // every SM a few items, which is what limits latency hiding on the deep levels.
    for (int n0 = 0; n0 < nn; n0 += blockDim.x) {
        const int p = n0 + threadIdx.x;
        if (p < nn) {
            const int h = base + p;
            if (na.state[h] == NODE_OPEN) {
                const int len = na.seg_len[h];
                int direct = 1;
                if (level > 0 && use_subtraction) {
                    const int sib = (h & 1) ? h + 1 : h - 1;
                    const int slen = na.seg_len[sib];
                    direct = (len < slen) || (len == slen && (h & 1));
                }
                if (direct && len > 0) atomicAdd(&s_direct_rows, (unsigned long long)len);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long want = (long long)(s_direct_rows / (unsigned long long)(row_groups > 0 ? row_groups : 1)) * nT_local / (148 * 6);
        int ir = item_rows_max;
        // (measured: smaller items do not pay off while every item ends with a full 8192-entry REDG flush)
        (void)want;
        s_item_rows = ir;
    }
    __syncthreads();
    const int item_rows = s_item_rows;
    for (int n0 = 0; n0 < nn; n0 += blockDim.x) {
        const int p = n0 + threadIdx.x;
        int my = 0, len = 0, start = 0, slot = p;
        if (p < nn) {
            const int h = base + p;
            int st = na.state[h];
            len = na.seg_len[h];
            start = na.seg_start[h];
            int direct = 0;
            if (st == NODE_OPEN) {
                if (level == 0 || !use_subtraction) direct = 1;
                else {
                    const int sib = (h & 1) ? h + 1 : h - 1;       // left children are odd
                    const int slen = na.seg_len[sib];
                    const bool left = (h & 1);
                    // the smaller child is direct (ties: the left one); its sibling is derived
                    direct = (len < slen) || (len == slen && left);
                }
                if (!oblivious && len == 0) direct = 1;            // nothing to add, histogram stays zero
                na.direct[h] = direct;
                if (direct && len > 0) {
                    const int chunks = ceil_div(len, item_rows);
                    // chunks c with c % row_groups == row_group belong to this rank
                    const int mine = chunks > row_group ? (chunks - row_group + row_groups - 1) / row_groups : 0;
                    my = mine * nT_local;
                    long long rows_mine = 0;
                    for (int c = row_group; c < chunks; c += row_groups) rows_mine += min(len, (c + 1) * item_rows) - c * item_rows;
                    atomicAdd(&s_rows, (unsigned long long)rows_mine);
                }
            }
        }
        // block exclusive scan of `my`
        s_cnt[threadIdx.x] = my;
        __syncthreads();
        for (int o = 1; o < blockDim.x; o <<= 1) {
            int v = (threadIdx.x >= o) ? s_cnt[threadIdx.x - o] : 0;
            __syncthreads();
            s_cnt[threadIdx.x] += v;
            __syncthreads();
        }
        const int incl = s_cnt[threadIdx.x];
        const int off = s_total + incl - my;
        // the items of every node of this pass are written cooperatively (a large node has thousands of them)
        __syncthreads();
        s_cnt[threadIdx.x] = my;
        s_off[threadIdx.x] = off; s_len[threadIdx.x] = len; s_start[threadIdx.x] = start;
        __syncthreads();
        const int in_pass = min((int)blockDim.x, nn - n0);
        for (int q = 0; q < in_pass; ++q) {
            const int qmy = s_cnt[q];
            if (qmy <= 0) continue;
            const int qlen = s_len[q], qstart = s_start[q], qoff = s_off[q];
            const int chunks = ceil_div(qlen, item_rows);
            const int mine = chunks > row_group ? (chunks - row_group + row_groups - 1) / row_groups : 0;   // chunks of this rank
            for (int j = threadIdx.x; j < qmy; j += blockDim.x) {
                // pair-major order: tile outer, row chunk inner
                const int t = j / mine, c = row_group + (j - t * mine) * row_groups;
                if (qoff + j < items_cap) {
                    Item it;
                    it.slot = n0 + q; it.tile = t;
                    it.k0 = qstart + c * item_rows;
                    it.k1 = min(qstart + qlen, it.k0 + item_rows);
                    items[qoff + j] = it;
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_total += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        ctl->n_items = min(s_total, items_cap);
        ctl->stat_hist_rows += s_rows;
    }
}


}  // namespace gb
