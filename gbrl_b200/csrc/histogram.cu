// histogram.cu -- per-level candidate-bin gradient histograms (the dominant kernel of the fit path).
//
// What it replaces in the reference: the O(C * n_node * D) brute-force rescan of every node sample for
// every candidate (TreeNode::splitScoreL2 / splitScoreCosine, node.cpp:187-251, 321-376).  Because a
// sample is right of candidate (f, j) iff code(x_f) > j, the per-side (count, sum g) of ALL n_bins
// candidates of a feature follow from one histogram over codes: right(j) = sum_{c > j} H[f][c].
//
// Kernel shape (sm_100a):
//   * work item = (node, 32-feature tile, <= 8192 consecutive rows of the node's segment of `order`);
//     items are sorted by (node, tile), and each of the 148 persistent CTAs (1024 threads, one per SM) takes
//     a contiguous run of them, so it usually stays on one (node, tile) "pair" for many items.
//   * lane <-> feature: a warp handles 4 rows x 8 lanes x 4 features; in each of its 4 rounds the 32
//     lanes address 32 different features, and the shared-memory histogram is laid out
//     plane[w][code][feature], so bank == feature: every ATOMS.ADD is bank-conflict free by
//     construction, whatever the codes are (ncu: 1.05 wavefronts per atomic).
//   * integer accumulation (north_star: "int32 atomics into shared-memory histograms"): build_grads
//     are converted to 30-bit fixed point q = hi*2^9 + lo and accumulated with TWO atomics per (sample, feature):
//     plane A += (1 << 20) + lo  (count in the top 12 bits, sum of lo below), plane B += hi; every 2048 rows the
//     planes are folded into per-CTA int64 / int32 shared accumulators (2048 * 2^9 = 2^20, 2048 * 2^20 = 2^31, so
//     nothing can overflow), which are written out only when the CTA leaves a pair:
//       - pair owned by this CTA alone  -> plain stores into the global int64 histogram
//       - pair shared by several CTAs   -> plain stores of a partial (96 KB) + hist_reduce_kernel
//     so the global histogram costs no atomics at all (the first version flushed 16K REDG.ADD.64 per item and
//     was bound by L2 atomic throughput, 193 G/s measured).  Integer sums are associative, so the histogram
//     (and everything derived from it: the parent - sibling subtraction, the multi-GPU all-reduce) is
//     bit-reproducible.
//   * rows are gathered through `order` (64 B per row and tile, one DRAM burst); loads of the next half-block
//     are in flight while the atomics of the current one are issued.
// Algorithmic bytes per level: N * (4F + 4D + 4)  (SURVEY 8d); DRAM traffic is lower because the
// fp32 feature matrix was quantised to u16 codes once per tree.
#include "engine.cuh"

namespace gb {

constexpr int HIST2_THREADS = 1024;
constexpr int HPLANE = (NB + 1) * FT;          // ints per plane; row 0 is the dump row of code 0
constexpr int HENT = NB * FT;                  // histogram entries per (node, tile)
constexpr int PARTIAL_WORDS = HENT * 3;        // int32 words of one partial: cnt[HENT] then sum[HENT] (int64)

// ---------------------------------------------------------------- level planning
// Decides, for every node of the level, whether its histogram is built directly or derived as
// parent - sibling (only the smaller child is histogrammed), and emits the work items sorted by
// (node, tile, chunk) together with the per-pair tables the histogram kernel needs.
__global__ void plan_level_kernel(NodeArrays na, Ctl *ctl, Item *items, int items_cap, int *pair_first, int *pair_nitems,
                                  int *pl_count, int level, int max_depth, int nT_local, int use_subtraction, int oblivious,
                                  int row_groups, int row_group) {
    __shared__ int s_cnt[1024];
    __shared__ int s_total;
    __shared__ unsigned long long s_rows;
    const int base = level_base(level), nn = 1 << level;
    if (threadIdx.x == 0) { s_total = 0; s_rows = 0; }
    __syncthreads();
    for (int n0 = 0; n0 < nn; n0 += blockDim.x) {
        const int p = n0 + threadIdx.x;
        int my = 0, len = 0, start = 0, slot = p, mine = 0;
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
                    const int chunks = ceil_div(len, ITEM_ROWS);
                    // chunks c with c % row_groups == row_group belong to this rank (2-D sharding)
                    mine = chunks > row_group ? (chunks - row_group + row_groups - 1) / row_groups : 0;
                    my = mine * nT_local;
                    long long rows_mine = 0;
                    for (int c = row_group; c < chunks; c += row_groups) rows_mine += min(len, (c + 1) * ITEM_ROWS) - c * ITEM_ROWS;
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
        if (p < nn) {
            for (int t = 0; t < nT_local; ++t) {
                const int pair = slot * nT_local + t;
                pair_first[pair] = off + t * mine;
                pair_nitems[pair] = mine;
                pl_count[pair] = 0;
            }
        }
        if (my > 0) {
            int w = off;
            const int chunks = ceil_div(len, ITEM_ROWS);
            for (int t = 0; t < nT_local; ++t)
                for (int c = row_group; c < chunks; c += row_groups) {
                    if (w < items_cap) {
                        Item it;
                        it.slot = slot; it.tile = t;
                        it.k0 = start + c * ITEM_ROWS;
                        it.k1 = min(start + len, it.k0 + ITEM_ROWS);
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
        ctl->n_partials = 0;
        ctl->stat_hist_rows += s_rows;
    }
}

void launch_plan_level(Model &m, int level, cudaStream_t s) {
    Workspace &ws = m.ws;
    GB_LAUNCH(plan_level_kernel, 1, 1024, 0, s, ws.na, ws.ctl.as<Ctl>(), ws.items.as<Item>(), ws.items_cap, ws.pair_first.as<int>(),
              ws.pair_nitems.as<int>(), ws.pl_count.as<int>(), level, m.cfg.max_depth, ws.tile_hi - ws.tile_lo,
              m.cfg.use_subtraction, m.cfg.grow_policy == GBRL_B200_GROW_OBLIVIOUS, ws.row_groups, ws.row_group);
}

// ---------------------------------------------------------------- the histogram kernel
// dynamic shared memory:
//   2 int32 planes of (NB+1)*FT   A = count<<20 | sum lo, B = sum hi of the current <=2048 rows (row 0 = dump row for
//                                 code 0, which keeps the inner loop branch-free)                         65,792 B
//   acc_sum int64[HENT], acc_cnt int32[HENT]   the CTA's running totals of the current pair              98,304 B
// The code matrix stores code*64 (u16), so the byte offset of (code, feature fs) inside a plane = stored*2 + fs*4.
__device__ __forceinline__ void red_shared(unsigned int addr, int v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
template <int OFF>
__device__ __forceinline__ void red_shared_off(unsigned int addr, int v) {
    asm volatile("red.shared.add.u32 [%0+%2], %1;" ::"r"(addr), "r"(v), "n"(OFF) : "memory");
}

struct HistParams {
    const uint16_t *codes; const float *bg; const int *order; const Item *items; Ctl *ctl;
    long long *hist;            // level buffer [slot][nT_total][NB][FT][1+D]
    int *partials;              // pool of partials, PARTIAL_WORDS int32 each
    const int *pair_first, *pair_nitems;
    int *pl_count, *pl_ids;     // per pair: number of partials and their pool ids
    int codes_rows, row_offset, D, d0, nT_local, tile_lo, nT_total, write_count, max_partials, pl_stride;
};

__global__ void __launch_bounds__(HIST2_THREADS, 1) hist_kernel(HistParams P) {
    extern __shared__ int sh[];
    constexpr int PB = HPLANE * 4;               // plane size in bytes
    long long *acc_sum = reinterpret_cast<long long *>(sh + 2 * HPLANE);
    int *acc_cnt = reinterpret_cast<int *>(acc_sum + HENT);
    __shared__ int s_pid;
    const int n_items = P.ctl->n_items;
    const float scale = exp2f((float)P.ctl->qexp);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rl = (lane >> 3) & 3, g = lane & 7;
    const int HS = 1 + P.D;

    for (int i = threadIdx.x; i < 2 * HPLANE; i += HIST2_THREADS) sh[i] = 0;
    for (int i = threadIdx.x; i < HENT; i += HIST2_THREADS) { acc_sum[i] = 0; acc_cnt[i] = 0; }
    __syncthreads();

    // per-lane constants of the 4 rounds: round k handles feature slot ms = (k + rl) & 3 of the lane's quad, so the
    // 32 lanes of a warp always address 32 different features (= 32 different banks).
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

    // contiguous run of items for this CTA
    const int i0 = (int)((long long)n_items * blockIdx.x / gridDim.x);
    const int i1 = (int)((long long)n_items * (blockIdx.x + 1) / gridDim.x);
    int cur_pair = -1, cur_slot = 0, cur_tile = 0;

    // writes the CTA's running totals of `pair` out and clears them
    auto emit = [&](int pair, int slot, int tile) {
        const int pf = P.pair_first[pair], pn = P.pair_nitems[pair];
        const bool exclusive = (pf >= i0) && (pf + pn <= i1);
        long long *hb = P.hist + ((size_t)slot * P.nT_total + (P.tile_lo + tile)) * (size_t)HENT * HS;
        if (exclusive) {
            for (int e = threadIdx.x; e < HENT; e += HIST2_THREADS) {
                const int c = acc_cnt[e];
                if (c != 0) {
                    if (P.write_count) hb[(size_t)e * HS] = (long long)c;
                    hb[(size_t)e * HS + 1 + P.d0] = acc_sum[e];
                    acc_cnt[e] = 0; acc_sum[e] = 0;
                }
            }
        } else {
            if (threadIdx.x == 0) s_pid = atomicAdd(&P.ctl->n_partials, 1);
            __syncthreads();
            const int pid = s_pid;
            if (pid < P.max_partials) {
                int *pc = P.partials + (size_t)pid * PARTIAL_WORDS;
                long long *ps = reinterpret_cast<long long *>(pc + HENT);
                for (int e = threadIdx.x; e < HENT; e += HIST2_THREADS) {
                    pc[e] = acc_cnt[e]; ps[e] = acc_sum[e];
                    acc_cnt[e] = 0; acc_sum[e] = 0;
                }
                __threadfence();
                __syncthreads();
                if (threadIdx.x == 0) {
                    const int k = atomicAdd(&P.pl_count[pair], 1);
                    P.pl_ids[(size_t)pair * P.pl_stride + k] = pid;
                }
            } else {
                // pool exhausted (cannot happen with max_partials = 2*grid + 2): fall back to global atomics
                for (int e = threadIdx.x; e < HENT; e += HIST2_THREADS) {
                    const int c = acc_cnt[e];
                    if (c != 0) {
                        if (P.write_count) red_add64(hb + (size_t)e * HS, (long long)c);
                        red_add64(hb + (size_t)e * HS + 1 + P.d0, acc_sum[e]);
                        acc_cnt[e] = 0; acc_sum[e] = 0;
                    }
                }
            }
        }
        __syncthreads();
    };

    for (int itx = i0; itx < i1; ++itx) {
        const Item it = P.items[itx];
        const int pair = it.slot * P.nT_local + it.tile;
        if (pair != cur_pair) {
            if (cur_pair >= 0) emit(cur_pair, cur_slot, cur_tile);
            cur_pair = pair; cur_slot = it.slot; cur_tile = it.tile;
        }
        const uint16_t *ctile = P.codes + ((size_t)it.tile * P.codes_rows + P.row_offset) * FT;
        constexpr int STEP = (HIST2_THREADS / 32) * 32;
        // Software pipeline over half-blocks of 16 rows (4 per lane group): while the shared-memory atomics of one
        // half are issued, the code quads / gradients of the next half are already in flight, and the row ids of
        // the block after that are prefetched.
        uint2 bA[4], bB[4];
        float gA[4], gB[4];
        auto load_half = [&](int rows32, int half, uint2 *b, float *gv) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int row = __shfl_sync(0xffffffffu, rows32, (half * 4 + s) * 4 + rl);
                if (row >= 0) {
                    b[s] = ld_nc_u2(reinterpret_cast<const uint2 *>(ctile + (size_t)row * FT + g * 4));
                    gv[s] = __ldg(P.bg + (size_t)row * P.D + P.d0);
                } else {
                    b[s] = make_uint2(0u, 0u);                  // code 0 -> dump row
                    gv[s] = 0.0f;
                }
            }
        };
        auto add_half = [&](const uint2 *b, const float *gv) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int q = __float2int_rn(gv[s] * scale);                    // |q| <= 2^28
                const int a = (1 << CNT_SHIFT) + (q & ((1 << LO_BITS) - 1));      // count unit + lo
                const int hi = q >> LO_BITS;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned int cs = __byte_perm(upper[k] ? b[s].y : b[s].x, 0u, psel[k]);   // code * 64, zero-extended
                    const unsigned int addr = lbase[k] + cs * 2u;
                    red_shared(addr, a);
                    red_shared_off<PB>(addr, hi);
                }
            }
        };
        for (int c0 = it.k0; c0 < it.k1; c0 += FOLD_ROWS) {
            const int c1 = min(it.k1, c0 + FOLD_ROWS);
            int kb = c0 + warp * 32;
            int cur_row = (kb + lane < c1) ? P.order[kb + lane] : -1;
            load_half(cur_row, 0, bA, gA);
            for (; kb < c1; kb += STEP) {
                const int kn = kb + STEP + lane;
                const int next_row = (kn < c1) ? P.order[kn] : -1;   // row ids of the next block
                load_half(cur_row, 1, bB, gB);
                add_half(bA, gA);
                load_half(next_row, 0, bA, gA);
                add_half(bB, gB);
                cur_row = next_row;
            }
            __syncthreads();
            // fold the planes of these <= 2048 rows into the running totals (8 entries per thread), clear the planes
            for (int e = threadIdx.x; e < HENT; e += HIST2_THREADS) {
                const int se = e + FT;                              // shared row = bin + 1
                const unsigned int a = (unsigned int)sh[se];
                if (a != 0u) {
                    const int h = sh[HPLANE + se];
                    acc_cnt[e] += (int)(a >> CNT_SHIFT);
                    acc_sum[e] += ((long long)h << LO_BITS) + (long long)(a & ((1u << CNT_SHIFT) - 1u));
                    sh[se] = 0; sh[HPLANE + se] = 0;
                }
            }
            __syncthreads();
        }
    }
    if (cur_pair >= 0) emit(cur_pair, cur_slot, cur_tile);
}

// sums the partials of every pair that was shared by several CTAs into the level histogram (plain stores:
// the slots were zero-filled and nobody else writes these entries)
__global__ void __launch_bounds__(256)
hist_reduce_kernel(const int *__restrict__ partials, const int *__restrict__ pl_count, const int *__restrict__ pl_ids,
                   long long *__restrict__ hist, int nT_local, int tile_lo, int nT_total, int D, int d0, int write_count,
                   int pl_stride, int max_partials) {
    const int pair = blockIdx.x;
    const int np = pl_count[pair];
    if (np == 0) return;
    const int slot = pair / nT_local, tile = pair % nT_local;
    const int HS = 1 + D;
    long long *hb = hist + ((size_t)slot * nT_total + (tile_lo + tile)) * (size_t)HENT * HS;
    for (int e = blockIdx.y * 256 + threadIdx.x; e < HENT; e += gridDim.y * 256) {
        long long c = 0, sm = 0;
#pragma unroll 4
        for (int k = 0; k < np; ++k) {
            const int pid = pl_ids[(size_t)pair * pl_stride + k];
            if (pid >= max_partials) continue;
            const int *pc = partials + (size_t)pid * PARTIAL_WORDS;
            c += pc[e];
            sm += reinterpret_cast<const long long *>(pc + HENT)[e];
        }
        if (c != 0) {
            if (write_count) hb[(size_t)e * HS] = c;
            hb[(size_t)e * HS + 1 + d0] = sm;
        }
    }
}

// `order` ping-pong: launch_partition swaps the two DevBufs, so ws.order[0] is always the current one.
void launch_histogram(Model &m, int level, cudaStream_t s) {
    Workspace &ws = m.ws;
    const int n_sms = ws.n_sms;
    static bool attr = false;
    const size_t smem = (size_t)2 * HPLANE * sizeof(int) + (size_t)HENT * (sizeof(long long) + sizeof(int));
    if (!attr) {
        GB_CUDA(cudaFuncSetAttribute(hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    const int nT_local = ws.tile_hi - ws.tile_lo;
    if (nT_local <= 0) return;
    HistParams P;
    P.codes = ws.codes.as<uint16_t>(); P.bg = ws.bg.as<float>(); P.order = ws.order[0].as<int>(); P.items = ws.items.as<Item>();
    P.ctl = ws.ctl.as<Ctl>(); P.hist = ws.hist[level & 1].as<long long>();
    P.partials = ws.partials.as<int>(); P.pair_first = ws.pair_first.as<int>(); P.pair_nitems = ws.pair_nitems.as<int>();
    P.pl_count = ws.pl_count.as<int>(); P.pl_ids = ws.pl_ids.as<int>();
    P.codes_rows = ws.codes_rows; P.row_offset = ws.row_offset; P.D = ws.D; P.nT_local = nT_local; P.tile_lo = ws.tile_lo;
    P.nT_total = ws.nT; P.max_partials = ws.max_partials; P.pl_stride = ws.pl_stride;
    const int n_pairs = (1 << level) * nT_local;
    Ctl *ctl = ws.ctl.as<Ctl>();
    // one pass per output dimension (the count plane is written by the first pass only)
    for (int d0 = 0; d0 < ws.D; ++d0) {
        P.d0 = d0; P.write_count = (d0 == 0);
        if (d0 > 0) {
            GB_CUDA(cudaMemsetAsync(&ctl->n_partials, 0, sizeof(int), s));
            GB_CUDA(cudaMemsetAsync(ws.pl_count.p, 0, (size_t)n_pairs * sizeof(int), s));
        }
        GB_LAUNCH(hist_kernel, n_sms, HIST2_THREADS, smem, s, P);
        GB_LAUNCH(hist_reduce_kernel, dim3(n_pairs, 32), 256, 0, s, ws.partials.as<int>(), ws.pl_count.as<int>(), ws.pl_ids.as<int>(),
                  P.hist, nT_local, ws.tile_lo, ws.nT, ws.D, d0, P.write_count, ws.pl_stride, ws.max_partials);
    }
}

}  // namespace gb
