// histogram.cu -- per-level candidate-bin gradient histograms (the dominant kernel of the fit path).
//
// What it replaces in the reference: the O(C * n_node * D) brute-force rescan of every node sample for
// every candidate (TreeNode::splitScoreL2 / splitScoreCosine, node.cpp:187-251, 321-376).  Because a
// sample is right of candidate (f, j) iff code(x_f) > j, the per-side (count, sum g) of ALL n_bins
// candidates of a feature follow from one histogram over codes: right(j) = sum_{c > j} H[f][c].
//
// Kernel shape (sm_100a), common to both variants:
//   * lane <-> feature: a warp handles 4 rows x 8 lanes x 4 features; in each of its 4 rounds the 32
//     lanes address 32 different features, and the shared-memory histogram is laid out
//     plane[w][code-1][feature], so bank == feature: every ATOMS.ADD is bank-conflict free by
//     construction, whatever the codes are.
//   * integer accumulation (north_star: "int32 atomics into shared-memory histograms"): build_grads
//     are converted to 34-bit fixed point q = hi*2^18 + lo; count, lo and hi are accumulated in three
//     int32 planes, then flushed to the global int64 histogram with REDG.ADD.64.  Integer sums are
//     associative, so the histogram (and everything derived from it, including the parent - sibling
//     subtraction and the multi-GPU all-reduce) is bit-reproducible.
//   * hist_stream_kernel (default): one CTA per SM, rows reach the atomics through a per-warp cp.async ring,
//     the shared histogram is carried across the items of a (node, tile) pair -- see the comment above it.
//   * hist_kernel (hist_variant = 1): one work item = (node, tile, <= 8192 rows) per CTA visit, 2 CTAs per SM,
//     rows prefetched into registers, one flush per item.
// Algorithmic bytes per level: N * (4F + 4D + 4)  (SURVEY 8d); DRAM traffic is lower because the
// fp32 feature matrix was quantised to u16 codes once per tree.
#include "engine.cuh"
#include "plan.cuh"

namespace gb {

// ---------------------------------------------------------------- level planning (plan.cuh)
__global__ void __launch_bounds__(1024) plan_level_kernel(NodeArrays na, Ctl *ctl, PlanParams Q, int level) {
    plan_level_body(na, ctl, Q.items, Q.items_cap, level, Q.max_depth, Q.nT_local, Q.use_subtraction, Q.oblivious, Q.row_groups, Q.row_group,
                    Q.item_rows_max);
}

// rows per work item: the streaming kernel wants NWARPS * 64 (two 32-row blocks per warp), the per-item kernel 8192
int hist_item_rows(const Model &m) {
    if (m.cfg.hist_variant == 1) return ITEM_ROWS;
    return (m.cfg.output_dim == 1 && m.cfg.hist_variant == 2) ? 32 * 64 : 24 * 64;
}

PlanParams plan_params(const Model &m) {
    const Workspace &ws = m.ws;
    PlanParams Q;
    Q.items = ws.items.as<Item>(); Q.items_cap = ws.items_cap; Q.max_depth = m.cfg.max_depth; Q.nT_local = ws.tile_hi - ws.tile_lo;
    Q.use_subtraction = m.cfg.use_subtraction; Q.oblivious = m.cfg.grow_policy == GBRL_B200_GROW_OBLIVIOUS;
    Q.row_groups = ws.row_groups; Q.row_group = ws.row_group; Q.item_rows_max = hist_item_rows(m);
    return Q;
}

// level 0 only: the plan of every later level is made by the fused decide + plan kernel of the level before (split.cu)
void launch_plan_level(Model &m, int level, cudaStream_t s) {
    Workspace &ws = m.ws;
    GB_LAUNCH(plan_level_kernel, 1, 1024, 0, s, ws.na, ws.ctl.as<Ctl>(), plan_params(m), level);
}

// ---------------------------------------------------------------- the histogram kernel
// dynamic shared memory: (1 + 2*ND) planes of (NB+1)*FT int32; row 0 of every plane is a dump row for code 0
// (x <= every threshold: right of no candidate), which keeps the inner loop branch-free.
// The code matrix stores code*128 (u16) = the byte offset of the code's row inside a plane; + fs*4 for the feature.
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
                    const unsigned int addr = lbase[k] + cs;
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

// ---------------------------------------------------------------- streaming variant (1 CTA per SM)
// Same arithmetic and the same global histogram as hist_kernel; what changes is how rows reach the atomics:
//   * one CTA of HS_WARPS warps per SM and ONE shared-memory histogram, which leaves room for a per-warp ring of
//     HS_STAGES x 16 rows of codes (64 B per row and tile) + gradients filled by cp.async (LDGSTS): rows are
//     gathered two stages (32 rows per warp, ~96 KB per SM) ahead of the atomics without costing registers, which
//     is what hides the HBM / L2 latency of the `order` gather on the deeper levels;
//   * items arrive pair-major ((node, tile) outer, row chunk inner) and a CTA owns a contiguous range of them, so
//     the shared histogram is carried from item to item (lo -> hi carry fold after every item) and is flushed to
//     HBM only when the (node, tile) pair changes or after FLUSH_ITEMS items: ~n_sms + n_pairs flushes per level
//     instead of one per item.
constexpr int HS_STAGE_ROWS = 16;
constexpr int HS_FOLD_ITEMS = 4;     // lo planes: 4 items x 2048 rows x 2^18 = 2^31
constexpr int HS_FLUSH_ITEMS = 32;   // hi planes: 32 items x 2048 rows x 2^14 = 2^30

__device__ __forceinline__ void cp_async_16(unsigned int dst, const void *src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_8(unsigned int dst, const void *src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int ND, int NWARPS, int NST>
__global__ void __launch_bounds__(NWARPS * 32, 1)
hist_stream_kernel(const uint16_t *__restrict__ codes, const int2 *__restrict__ bgq, const int *__restrict__ order,
                   const Item *__restrict__ items, const Ctl *__restrict__ ctl, long long *__restrict__ hist, int codes_rows,
                   int row_offset, int D, int d0, int nT_local, int tile_lo, int nT_total, int write_count) {
    extern __shared__ int sh[];
    constexpr int W = 1 + 2 * ND;
    constexpr int PB = HPLANE * 4;               // plane size in bytes
    constexpr int NTHREADS = NWARPS * 32;
    constexpr int RING_GRAD = HS_STAGE_ROWS * 64;              // a ring stage: 16 rows x 64 B of codes, then 16 x ND fixed-point gradients (lo, hi)
    constexpr int RING_STAGE = RING_GRAD + HS_STAGE_ROWS * 8 * ND;
    const int n_items = ctl->n_items;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rl = (lane >> 3) & 3, g = lane & 7;
    const int HS = 1 + D;
    // contiguous range of items of this CTA
    const int i0 = (int)((long long)n_items * blockIdx.x / gridDim.x), i1 = (int)((long long)n_items * (blockIdx.x + 1) / gridDim.x);
    if (i0 >= i1) return;
    for (int i = threadIdx.x; i < W * HPLANE; i += NTHREADS) sh[i] = 0;
    const unsigned int sbase = (unsigned int)__cvta_generic_to_shared(sh);
    const unsigned int ring = sbase + (unsigned int)(W * PB) + (unsigned int)(warp * NST * RING_STAGE);
    unsigned int lbase[4], psel[4];
    bool upper[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int ms = (k + rl) & 3;
        lbase[k] = sbase + (unsigned int)(g * 4 + ms) * 4u;
        upper[k] = (ms & 2) != 0;
        const unsigned int lo = 2u * (ms & 1), hi = lo + 1u;
        psel[k] = lo | (hi << 4) | (4u << 8) | (4u << 12);
    }
    __syncthreads();

    // flush: (bin, feature) e -> global [slot][tile][bin][feature][1+D]; shared row = bin + 1.  The level buffer is zeroed
    // before the launch, so a (node, tile) pair whose items all belong to this CTA is written with plain stores; pairs
    // that are shared with a neighbouring CTA (or flushed more than once) are added with REDG.64.
    auto flush = [&](const Item &it, bool excl) {
        long long *hb = hist + ((size_t)it.slot * nT_total + (tile_lo + it.tile)) * (size_t)(NB * FT) * HS;
        for (int e = threadIdx.x; e < NB * FT; e += NTHREADS) {
            const int se = e + FT;
            const int cnt = sh[se];
            if (cnt != 0) {
                if (write_count) { if (excl) hb[(size_t)e * HS] = (long long)cnt; else red_add64(hb + (size_t)e * HS, (long long)cnt); }
                sh[se] = 0;
#pragma unroll
                for (int dd = 0; dd < ND; ++dd) {
                    const unsigned int l = (unsigned int)sh[(1 + 2 * dd) * HPLANE + se];
                    const int h = sh[(2 + 2 * dd) * HPLANE + se];
                    const long long tot = ((long long)h << LO_BITS) + (long long)l;
                    if (tot != 0) { if (excl) hb[(size_t)e * HS + 1 + d0 + dd] = tot; else red_add64(hb + (size_t)e * HS + 1 + d0 + dd, tot); }
                    sh[(1 + 2 * dd) * HPLANE + se] = 0;
                    sh[(2 + 2 * dd) * HPLANE + se] = 0;
                }
            }
        }
        // the dump rows (code 0) are never flushed; keep them from overflowing
        for (int e = threadIdx.x; e < FT; e += NTHREADS)
#pragma unroll
            for (int w = 0; w < W; ++w) sh[w * HPLANE + e] = 0;
    };
    // carry fold: lo keeps 18 bits, the rest moves to hi (lo is read as unsigned: it may have reached 2^31)
    auto fold = [&]() {
        for (int e = threadIdx.x; e < (NB + 1) * FT; e += NTHREADS) {
#pragma unroll
            for (int dd = 0; dd < ND; ++dd) {
                const unsigned int l = (unsigned int)sh[(1 + 2 * dd) * HPLANE + e];
                if (l >> LO_BITS) {
                    sh[(2 + 2 * dd) * HPLANE + e] += (int)(l >> LO_BITS);
                    sh[(1 + 2 * dd) * HPLANE + e] = (int)(l & ((1u << LO_BITS) - 1u));
                }
            }
        }
    };

    // Every item is at most HS_ITEM_ROWS = NWARPS * 64 rows: a warp owns two blocks of 32 consecutive positions of it
    // (block b at k0 + (b * NWARPS + warp) * 32), i.e. exactly four ring stages of 16 rows per item.  The stages of
    // all the CTA's items form one stream per warp, t = 0 .. 4 * n_my - 1 (item t >> 2, block (t >> 1) & 1, half
    // t & 1); gathers run NST - 1 stages ahead of the atomics ACROSS item boundaries.
    const int n_my = i1 - i0, T = 4 * n_my;
    int rows_i = -1, rows_n = -1;                 // row ids of the block being issued / of the block after it (lane = position)
    int tile_i = 0, tile_n = 0;
    unsigned int vbits = 0;                       // bit (t & 31): stage t has at least one valid row
    auto block_rows = [&](int blk, int &tile) -> int {        // blk = global block index of this warp, = t >> 1
        if (blk >= 2 * n_my) { tile = 0; return -1; }
        const Item it = items[i0 + (blk >> 1)];
        tile = it.tile;
        const int k = it.k0 + ((blk & 1) * NWARPS + warp) * 32 + lane;
        return k < it.k1 ? order[k] : -1;
    };
    int slot_issue = 0, slot_use = 0;             // ring slots of the stage being issued / consumed (t % NST, kept incrementally)
    auto issue = [&](int t) {
        if (t < T) {
            if ((t & 1) == 0) {                   // first half of a block: rotate the prefetched row ids
                rows_i = rows_n; tile_i = tile_n;
                rows_n = block_rows((t >> 1) + 1, tile_n);
            }
            const int half = t & 1;
            const int mine = __shfl_sync(0xffffffffu, rows_i, half * 16 + (lane & 15));
            if (__any_sync(0xffffffffu, mine >= 0)) {
                vbits |= 1u << (t & 31);
                const uint16_t *ctile = codes + ((size_t)tile_i * codes_rows + row_offset) * FT;
                const unsigned int dst = ring + (unsigned int)(slot_issue * RING_STAGE);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int r = j * 8 + (lane >> 2), q = lane & 3;            // row in stage, 16-byte quarter
                    const int row = __shfl_sync(0xffffffffu, rows_i, half * 16 + r);
                    cp_async_16(dst + (unsigned int)(r * 64 + q * 16), ctile + (size_t)(row >= 0 ? row : 0) * FT + q * 8, row >= 0 ? 16 : 0);
                }
                if (lane < 16) {
#pragma unroll
                    for (int dd = 0; dd < ND; ++dd)
                        cp_async_8(dst + (unsigned int)(RING_GRAD + (lane * ND + dd) * 8), bgq + (size_t)(mine >= 0 ? mine : 0) * D + d0 + dd,
                                   mine >= 0 ? 8 : 0);
                }
            } else vbits &= ~(1u << (t & 31));
        }
        slot_issue = slot_issue + 1 == NST ? 0 : slot_issue + 1;      // ring slot of the next stage (no modulo in the loop)
        cp_async_commit();
    };
    rows_n = block_rows(0, tile_n);
#pragma unroll 1
    for (int t = 0; t < NST - 1; ++t) issue(t);
    int since_flush = 0, since_fold = 0;
    Item prev = items[i0];
    // is the current (node, tile) pair exclusively this CTA's?  (its first item is inside the CTA's range; cleared by a forced flush)
    bool pair_excl = true;
    if (i0 > 0) { const Item pv = items[i0 - 1]; pair_excl = pv.slot != prev.slot || pv.tile != prev.tile; }
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
        issue(t + NST - 1);
        cp_async_wait<NST - 1>();
        __syncwarp();
        if ((vbits >> (t & 31)) & 1u) {
            const unsigned int src = ring + (unsigned int)(slot_use * RING_STAGE);
#pragma unroll
            for (int ss = 0; ss < 4; ++ss) {
                const int r = ss * 4 + rl;
                uint2 b;
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(b.x), "=r"(b.y) : "r"(src + (unsigned int)(r * 64 + g * 8)));
                int lo[ND], hi[ND];
#pragma unroll
                for (int dd = 0; dd < ND; ++dd)       // (lo, hi) of the fixed-point gradient, converted once per tree (preprocess.cu)
                    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(lo[dd]), "=r"(hi[dd]) : "r"(src + (unsigned int)(RING_GRAD + (r * ND + dd) * 8)));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned int cs = __byte_perm(upper[k] ? b.y : b.x, 0u, psel[k]);   // code * 128, zero-extended
                    const unsigned int addr = lbase[k] + cs;
                    red_shared(addr, 1);
                    red_shared_off<PB>(addr, lo[0]);
                    red_shared_off<2 * PB>(addr, hi[0]);
                    if (ND > 1) { red_shared_off<3 * PB>(addr, lo[ND > 1 ? 1 : 0]); red_shared_off<4 * PB>(addr, hi[ND > 1 ? 1 : 0]); }
                }
            }
        }
        __syncwarp();                             // the stage buffer is refilled NST - 1 iterations later by other lanes
        slot_use = slot_use + 1 == NST ? 0 : slot_use + 1;
        if ((t & 3) == 3) {                       // item boundary (CTA-uniform decisions: they depend on the item list only)
            ++since_flush; ++since_fold;
            const int nx = (t >> 2) + 1;
            bool pair_ends = false, forced = false;
            Item nxt = prev;
            if (nx < n_my) {
                nxt = items[i0 + nx];
                pair_ends = nxt.slot != prev.slot || nxt.tile != prev.tile;
                forced = !pair_ends && since_flush >= HS_FLUSH_ITEMS;
            }
            if (pair_ends || forced) {
                __syncthreads();
                flush(prev, pair_excl && pair_ends);
                __syncthreads();
                since_flush = 0; since_fold = 0;
                pair_excl = pair_ends;            // a new pair starts inside the range; a forced flush leaves the pair shared with itself
            } else if (since_fold >= HS_FOLD_ITEMS && nx < n_my) {
                __syncthreads();
                fold();
                __syncthreads();
                since_fold = 0;
            }
            prev = nxt;
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    {
        bool last_inside = true;                  // does the pair end with this CTA's last item?
        if (i1 < n_items) { const Item nx = items[i1]; last_inside = nx.slot != prev.slot || nx.tile != prev.tile; }
        flush(prev, pair_excl && last_inside);
    }
}

template <int ND, int NWARPS, int NST>
static void launch_hist_stream(Model &m, int d0, int write_count, long long *hist, int n_sms, cudaStream_t s) {
    Workspace &ws = m.ws;
    const size_t smem = (size_t)(1 + 2 * ND) * HPLANE * sizeof(int) + (size_t)NWARPS * NST * HS_STAGE_ROWS * (64 + 8 * ND);   // planes + rings
    ensure_dyn_smem(hist_stream_kernel<ND, NWARPS, NST>, smem);
    GB_LAUNCH((hist_stream_kernel<ND, NWARPS, NST>), n_sms, NWARPS * 32, smem, s, ws.codes.as<uint16_t>(), ws.bgq.as<int2>(),
              ws.order_p[0], ws.items.as<Item>(), ws.ctl.as<Ctl>(), hist, ws.codes_rows, ws.row_offset, ws.D, d0,
              ws.tile_hi - ws.tile_lo, ws.tile_lo, ws.nT, write_count);
}

template <int ND>
static void launch_hist_nd(Model &m, int d0, int write_count, long long *hist, int n_sms, cudaStream_t s) {
    Workspace &ws = m.ws;
    const size_t smem = (size_t)(1 + 2 * ND) * HPLANE * sizeof(int);
    ensure_dyn_smem(hist_kernel<ND>, smem);
    const int ctas_per_sm = ND == 1 ? 2 : 1;
    GB_LAUNCH(hist_kernel<ND>, n_sms * ctas_per_sm, HIST_THREADS, smem, s, ws.codes.as<uint16_t>(), ws.bg.as<float>(),
              ws.order_p[0], ws.items.as<Item>(), ws.ctl.as<Ctl>(), hist, ws.codes_rows, ws.row_offset, ws.D, d0,
              ws.tile_hi - ws.tile_lo, ws.tile_lo, ws.nT, write_count);
}

// `order` ping-pong: launch_partition swaps the two DevBufs, so ws.order[0] is always the current one.
void launch_histogram(Model &m, int level, cudaStream_t s) {
    Workspace &ws = m.ws;
    const int n_sms = ws.n_sms;
    long long *hist = ws.hist_p[level & 1];
    const int D = ws.D;
    int d0 = 0;
    while (d0 < D) {
        const int nd = (D - d0 >= 3) ? 3 : (D - d0);
        const int wc = (d0 == 0);
        const bool stream = m.cfg.hist_variant != 1;      // 0: 24 warps per CTA (more registers per thread; measured faster); 2: 32 warps where output_dim == 1
        if (stream) {
            // two output dimensions per launch at most: the ring needs the shared memory a third pair of planes would take
            if (nd >= 2) { launch_hist_stream<2, 24, 2>(m, d0, wc, hist, n_sms, s); d0 += 2; }
            else if (D == 1 && m.cfg.hist_variant == 2) { launch_hist_stream<1, 32, 3>(m, d0, wc, hist, n_sms, s); d0 += 1; }
            else { launch_hist_stream<1, 24, 3>(m, d0, wc, hist, n_sms, s); d0 += 1; }      // same item size as the ND = 2 launches
            continue;
        }
        if (nd == 1) launch_hist_nd<1>(m, d0, wc, hist, n_sms, s);
        else if (nd == 2) launch_hist_nd<2>(m, d0, wc, hist, n_sms, s);
        else launch_hist_nd<3>(m, d0, wc, hist, n_sms, s);
        d0 += nd;
    }
}

}  // namespace gb
