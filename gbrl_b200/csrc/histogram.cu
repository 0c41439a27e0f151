// histogram.cu -- per-level candidate-bin gradient histograms (the dominant kernel of the fit path).
//
// What it replaces in the reference: the O(C * n_node * D) brute-force rescan of every node sample for
// every candidate (TreeNode::splitScoreL2 / splitScoreCosine, node.cpp:187-251, 321-376).  Because a
// sample is right of candidate (f, j) iff code(x_f) > j, the per-side (count, sum g) of ALL n_bins
// candidates of a feature follow from one histogram over codes: right(j) = sum_{c > j} H[f][c].
//
// Kernel shape (sm_100a):
//   * work item = (node, 32-feature tile, <= 8192 consecutive rows of the node's segment of `order`)
//   * lane <-> feature: a warp handles 4 rows x 8 lanes x 4 features; in each of its 4 rounds the 32
//     lanes address 32 different features, and the shared-memory histogram is laid out
//     plane[w][code-1][feature], so bank == feature: every ATOMS.ADD is bank-conflict free by
//     construction, whatever the codes are.
//   * integer accumulation (north_star: "int32 atomics into shared-memory histograms"): build_grads
//     are converted to 36-bit fixed point q = hi*2^18 + lo; count, lo and hi are accumulated in three
//     int32 planes (8192 rows * 2^18 < 2^31), then flushed once per item to the global int64 histogram
//     with REDG.ADD.64.  Integer sums are associative, so the histogram (and everything derived from
//     it, including the parent - sibling subtraction and the multi-GPU all-reduce) is bit-reproducible.
//   * rows are gathered through `order` (64 B per row and tile, one DRAM burst), 32 rows in flight per
//     warp before the atomics start.
// Algorithmic bytes per level: N * (4F + 4D + 4)  (SURVEY 8d); DRAM traffic is lower because the
// fp32 feature matrix was quantised to u16 codes once per tree.
#include "engine.cuh"

namespace gb {

// ---------------------------------------------------------------- level planning
// Decides, for every node of the level, whether its histogram is built directly or derived as
// parent - sibling (only the smaller child is histogrammed), and emits the work items.
__global__ void plan_level_kernel(NodeArrays na, Ctl *ctl, Item *items, int items_cap, int level, int max_depth,
                                  int nT_local, int use_subtraction, int oblivious, int row_groups, int row_group) {
    __shared__ int s_cnt[1024];
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
        int ir = ITEM_ROWS;
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
        if (my > 0) {
            int w = off;
            const int chunks = ceil_div(len, item_rows);
            for (int c = row_group; c < chunks; c += row_groups)
                for (int t = 0; t < nT_local; ++t) {
                    if (w < items_cap) {
                        Item it;
                        it.slot = slot; it.tile = t;
                        it.k0 = start + c * item_rows;
                        it.k1 = min(start + len, it.k0 + item_rows);
                        items[w] = it;
                    }
                    ++w;
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

void launch_plan_level(Model &m, int level, cudaStream_t s) {
    Workspace &ws = m.ws;
    GB_LAUNCH(plan_level_kernel, 1, 1024, 0, s, ws.na, ws.ctl.as<Ctl>(), ws.items.as<Item>(), ws.items_cap, level,
              m.cfg.max_depth, ws.tile_hi - ws.tile_lo, m.cfg.use_subtraction, m.cfg.grow_policy == GBRL_B200_GROW_OBLIVIOUS,
              ws.row_groups, ws.row_group);
}

// ---------------------------------------------------------------- the histogram kernel
// dynamic shared memory: (1 + 2*ND) planes of (NB+1)*FT int32; row 0 of every plane is a dump row for code 0
// (x <= every threshold: right of no candidate), which keeps the inner loop branch-free.
// The code matrix stores code*64 (u16), so byte offset of (code, feature fs) inside a plane = stored*2 + fs*4.
constexpr int HPLANE = (NB + 1) * FT;          // ints per plane

__device__ __forceinline__ void red_shared(unsigned int addr, int v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
template <int OFF>
__device__ __forceinline__ void red_shared_off(unsigned int addr, int v) {
    asm volatile("red.shared.add.u32 [%0+%2], %1;" ::"r"(addr), "r"(v), "n"(OFF) : "memory");
}

template <int ND>
__global__ void __launch_bounds__(HIST_THREADS, ND == 1 ? 2 : 1)
hist_kernel(const uint16_t *__restrict__ codes, const float *__restrict__ bg, const int *__restrict__ order,
            const Item *__restrict__ items, const Ctl *__restrict__ ctl, long long *__restrict__ hist, int codes_rows,
            int row_offset, int D, int d0, int nT_local, int tile_lo, int nT_total, int write_count) {
    extern __shared__ int sh[];
    constexpr int W = 1 + 2 * ND;
    constexpr int PB = HPLANE * 4;               // plane size in bytes
    const int n_items = ctl->n_items;
    const float scale = exp2f((float)ctl->qexp);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rl = (lane >> 3) & 3, g = lane & 7;
    const int HS = 1 + D;   // int64 words per (bin, feature)

    for (int i = threadIdx.x; i < W * HPLANE; i += HIST_THREADS) sh[i] = 0;
    __syncthreads();

    // per-lane constants of the 4 rounds: which halfword of the 8-byte code quad, and its column offset.
    // round k handles feature slot ms = (k + rl) & 3 of the lane's quad, so the 32 lanes of a warp always address
    // 32 different features (= 32 different banks).
    const unsigned int sbase = (unsigned int)__cvta_generic_to_shared(sh);
    unsigned int lbase[4], psel[4];
    bool upper[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int ms = (k + rl) & 3;
        lbase[k] = sbase + (unsigned int)(g * 4 + ms) * 4u;
        upper[k] = (ms & 2) != 0;                              // halfword lives in .y
        const unsigned int lo = 2u * (ms & 1), hi = lo + 1u;   // bytes of the halfword inside its 32-bit word
        psel[k] = lo | (hi << 4) | (4u << 8) | (4u << 12);     // upper two bytes from the zero operand
    }

    for (int itx = blockIdx.x; itx < n_items; itx += gridDim.x) {
        const Item it = items[itx];
        const uint16_t *ctile = codes + ((size_t)it.tile * codes_rows + row_offset) * FT;
        constexpr int STEP = (HIST_THREADS / 32) * 32;
        // Software pipeline over half-blocks of 16 rows (4 per lane group): while the shared-memory atomics of one
        // half are issued, the code quads / gradients of the next half are already in flight, and the row ids of
        // the block after that are prefetched.  Register budget is the same as a non-pipelined 32-row block.
        uint2 bA[4], bB[4];
        float gA[4][ND], gB[4][ND];
        auto load_half = [&](int rows32, int half, uint2 *b, float (*gv)[ND]) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int row = __shfl_sync(0xffffffffu, rows32, (half * 4 + s) * 4 + rl);
                if (row >= 0) {
                    b[s] = ld_nc_u2(reinterpret_cast<const uint2 *>(ctile + (size_t)row * FT + g * 4));
#pragma unroll
                    for (int dd = 0; dd < ND; ++dd) gv[s][dd] = __ldg(bg + (size_t)row * D + d0 + dd);
                } else {
                    b[s] = make_uint2(0u, 0u);                  // code 0 -> dump row
#pragma unroll
                    for (int dd = 0; dd < ND; ++dd) gv[s][dd] = 0.0f;
                }
            }
        };
        auto add_half = [&](const uint2 *b, const float (*gv)[ND]) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                int lo[ND], hi[ND];
#pragma unroll
                for (int dd = 0; dd < ND; ++dd) {
                    const long long q = __float2ll_rn(gv[s][dd] * scale);
                    lo[dd] = (int)(q & ((1ll << LO_BITS) - 1));
                    hi[dd] = (int)(q >> LO_BITS);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned int cs = __byte_perm(upper[k] ? b[s].y : b[s].x, 0u, psel[k]);   // code * 64, zero-extended
                    const unsigned int addr = lbase[k] + cs * 2u;
                    red_shared(addr, 1);
                    red_shared_off<PB>(addr, lo[0]);
                    red_shared_off<2 * PB>(addr, hi[0]);
                    if (ND > 1) { red_shared_off<3 * PB>(addr, lo[ND > 1 ? 1 : 0]); red_shared_off<4 * PB>(addr, hi[ND > 1 ? 1 : 0]); }
                    if (ND > 2) { red_shared_off<5 * PB>(addr, lo[ND > 2 ? 2 : 0]); red_shared_off<6 * PB>(addr, hi[ND > 2 ? 2 : 0]); }
                }
            }
        };
        int kb = it.k0 + warp * 32;
        int cur_row = (kb + lane < it.k1) ? order[kb + lane] : -1;
        int next_row = (kb + STEP + lane < it.k1) ? order[kb + STEP + lane] : -1;
        load_half(cur_row, 0, bA, gA);
        for (; kb < it.k1; kb += STEP) {
            const int kn2 = kb + 2 * STEP + lane;
            const int next2_row = (kn2 < it.k1) ? order[kn2] : -1;  // row ids two blocks ahead
            // pull the code rows of the next block into L2 while this block is processed (no registers needed)
            if (next_row >= 0) {
                const char *pf = reinterpret_cast<const char *>(ctile + (size_t)next_row * FT);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
            }
            load_half(cur_row, 1, bB, gB);
            add_half(bA, gA);
            load_half(next_row, 0, bA, gA);
            add_half(bB, gB);
            cur_row = next_row;
            next_row = next2_row;
        }
        __syncthreads();
        // flush: (bin, feature) e -> global [slot][tile][bin][feature][1+D]; shared row = bin + 1
        long long *hb = hist + ((size_t)it.slot * nT_total + (tile_lo + it.tile)) * (size_t)(NB * FT) * HS;
        for (int e = threadIdx.x; e < NB * FT; e += HIST_THREADS) {
            const int se = e + FT;
            const int cnt = sh[se];
            if (cnt != 0) {
                if (write_count) red_add64(hb + (size_t)e * HS, (long long)cnt);
                sh[se] = 0;
#pragma unroll
                for (int dd = 0; dd < ND; ++dd) {
                    const unsigned int l = (unsigned int)sh[(1 + 2 * dd) * HPLANE + se];
                    const int h = sh[(2 + 2 * dd) * HPLANE + se];
                    const long long tot = ((long long)h << LO_BITS) + (long long)l;
                    if (tot != 0) red_add64(hb + (size_t)e * HS + 1 + d0 + dd, tot);
                    sh[(1 + 2 * dd) * HPLANE + se] = 0;
                    sh[(2 + 2 * dd) * HPLANE + se] = 0;
                }
            }
        }
        __syncthreads();
    }
}

template <int ND>
static void launch_hist_nd(Model &m, int d0, int write_count, long long *hist, int n_sms, cudaStream_t s) {
    Workspace &ws = m.ws;
    const size_t smem = (size_t)(1 + 2 * ND) * HPLANE * sizeof(int);
    static bool attr_set[4] = {false, false, false, false};
    if (!attr_set[ND]) {
        GB_CUDA(cudaFuncSetAttribute(hist_kernel<ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[ND] = true;
    }
    const int ctas_per_sm = ND == 1 ? 2 : 1;
    GB_LAUNCH(hist_kernel<ND>, n_sms * ctas_per_sm, HIST_THREADS, smem, s, ws.codes.as<uint16_t>(), ws.bg.as<float>(),
              ws.order[0].as<int>(), ws.items.as<Item>(), ws.ctl.as<Ctl>(), hist, ws.codes_rows, ws.row_offset, ws.D, d0,
              ws.tile_hi - ws.tile_lo, ws.tile_lo, ws.nT, write_count);
}

// `order` ping-pong: launch_partition swaps the two DevBufs, so ws.order[0] is always the current one.
void launch_histogram(Model &m, int level, cudaStream_t s) {
    Workspace &ws = m.ws;
    static int n_sms = 0;
    if (!n_sms) {
        cudaDeviceProp p;
        GB_CUDA(cudaGetDeviceProperties(&p, m.device));
        n_sms = p.multiProcessorCount;
    }
    long long *hist = ws.hist[level & 1].as<long long>();
    const int D = ws.D;
    int d0 = 0;
    while (d0 < D) {
        const int nd = (D - d0 >= 3) ? 3 : (D - d0);
        const int wc = (d0 == 0);
        if (nd == 1) launch_hist_nd<1>(m, d0, wc, hist, n_sms, s);
        else if (nd == 2) launch_hist_nd<2>(m, d0, wc, hist, n_sms, s);
        else launch_hist_nd<3>(m, d0, wc, hist, n_sms, s);
        d0 += nd;
    }
}

}  // namespace gb
