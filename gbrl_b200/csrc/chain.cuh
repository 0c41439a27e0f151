// chain.cuh -- bit-exact PARALLEL evaluation of a sequential fp32 summation chain.
//
// The reference sums floats one after the other:  s <- fl(s + x_i)   (node.cpp:336-352, math_ops.cpp:255-300).
// A float chain is not associative, but inside one binade it is integer arithmetic in disguise: while
// |s| stays in [2^e, 2^(e+1)) every result is a multiple of u = 2^(e-23), so with m = s/u (an integer,
// 2^23 <= |m| < 2^24) and y = x/u
//        fl(s + x) = u * RN(m + y) = u * (m + RN_parity(y)),
// where RN(y) is round-to-nearest of y alone, except for exact ties (frac(y) == 1/2), which go to the EVEN
// neighbour of m + y, i.e. depend on the parity of m.  After a tie the running sum is even, so the only state a
// block of elements needs from its predecessor is the parity p of the incoming m.  A block is therefore summarised
// by a two-entry table  a[p] = total integer increment when the incoming parity is p,  tables compose
// associatively ((g then f)[p] = g[p] + f[(p + g[p]) & 1]), and integer prefix sums are exact in any order.
// The model is valid only while every intermediate m stays strictly inside (2^23, 2^24) with the sign of the
// incoming m; the summary carries the min / max prefix so that the consumer can check that for its actual m and
// otherwise fall back to the plain sequential float chain for that block (always right by definition).
//
// Used by the near-tie replay (split.cu) and by the thread-partitioned mean / std chains (preprocess.cu).
#pragma once
#include "common.cuh"

namespace gb {
namespace seq {

struct Tab {           // summary of a block of chain elements for one epoch (binade) of the running sum
    int a0, a1;        // total increment (units of u) if the incoming m is even / odd
    int mn, mx;        // min / max inclusive prefix (path p = 0; the other path differs by at most 1)
};
constexpr int MARGIN = 4;
constexpr int BAD = 1 << 30;

// binade of s: inv_u = 2^(23-e), u = 2^(e-23).  False for 0, denormals / tiny sums, inf, NaN.
__device__ __forceinline__ bool epoch_of(float s, float &inv_u, float &u) {
    const unsigned int ex = (__float_as_uint(s) >> 23) & 0xffu;
    if (ex < 27u || ex == 255u) return false;
    inv_u = __uint_as_float((277u - ex) << 23);
    u = __uint_as_float((ex - 23u) << 23);
    return true;
}

// Per-lane view of a summarised block.  One warp; lane l holds KE CONSECUTIVE chain elements x[0..KE) (lanes in chain
// order), 0 for "not a member".
struct Scan {
    int e0, e1;        // exclusive table: increment from the block start to this lane's first element (incoming m even / odd)
    int mn, mx;        // min / max inclusive prefix inside the lane, relative to the lane start (lane path p = 0)
    int t0, t1;        // block totals
    bool big;          // this lane holds an element the integer model does not cover (|x| >= |s| / 2, inf, NaN)
};

template <int KE>
__device__ __forceinline__ Scan warp_scan(const float (&x)[KE], float inv_u) {
    const unsigned int full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    // Common case first (no exact tie, nothing out of range in the whole warp): RN(y) by the magic-number add
    // z = y + 1.5*2^23 (exact round-to-nearest-even of y to an integer for |y| < 2^22), 7 plain FP32/INT ops per element.
    constexpr float MAGIC = 12582912.0f;
    constexpr int MAGIC_I = 0x4B400000;
    Scan sc;
    {
        int a = 0, mn = 0, mx = 0;
        bool anytie = false, anybig = false;
#pragma unroll
        for (int i = 0; i < KE; ++i) {
            const float y = x[i] * inv_u;                // exact (power of two), may overflow to inf -> big
            const float z = y + MAGIC;
            const float fr = y - (z - MAGIC);            // exact rounding error of RN(y) while |y| < 2^22
            anybig |= !(fabsf(y) < 4194304.0f);          // 2^22 (also inf / NaN): never summarised
            anytie |= fabsf(fr) == 0.5f;
            a += __float_as_int(z) - MAGIC_I;
            mn = min(mn, a); mx = max(mx, a);
        }
        sc.big = anybig;
        const bool special = anytie || anybig;
        if (!__any_sync(full, special)) {
            int inc = a;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int v = __shfl_up_sync(full, inc, off);
                if (lane >= off) inc += v;
            }
            sc.e0 = sc.e1 = inc - a;
            sc.mn = mn; sc.mx = mx;
            sc.t0 = sc.t1 = __shfl_sync(full, inc, 31);
            return sc;
        }
    }
    // General case: exact ties resolved per incoming parity, out-of-range elements flagged (their lane is never applied).
    int d0[KE], d1[KE];
#pragma unroll
    for (int i = 0; i < KE; ++i) {
        const float y = x[i] * inv_u;
        const float d = rintf(y);
        const float fr = y - d;                          // exact
        const bool big = !(fabsf(y) < 4194304.0f);
        const bool tie = fabsf(fr) == 0.5f;
        const int di = big ? 0 : (int)d;
        const int k = di - ((tie && fr < 0.0f) ? 1 : 0); // floor(y) of a tie
        d0[i] = tie ? (k + (k & 1)) : di;                // m even: the even one of {m+k, m+k+1}
        d1[i] = tie ? (k + ((k + 1) & 1)) : di;          // m odd
    }
    int a0 = 0, a1 = 0, mn = 0, mx = 0;
#pragma unroll
    for (int i = 0; i < KE; ++i) {
        a0 += (a0 & 1) ? d1[i] : d0[i];
        a1 += ((a1 + 1) & 1) ? d1[i] : d0[i];
        mn = min(mn, a0); mx = max(mx, a0);
    }
    int i0 = a0, i1 = a1;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int g0 = __shfl_up_sync(full, i0, off), g1 = __shfl_up_sync(full, i1, off);
        if (lane >= off) {
            const int n0 = g0 + ((g0 & 1) ? i1 : i0);
            const int n1 = g1 + (((g1 + 1) & 1) ? i1 : i0);
            i0 = n0; i1 = n1;
        }
    }
    sc.e0 = __shfl_up_sync(full, i0, 1);
    sc.e1 = __shfl_up_sync(full, i1, 1);
    if (lane == 0) { sc.e0 = 0; sc.e1 = 0; }
    sc.mn = mn - 1; sc.mx = mx + 1;                      // the lane may be entered on its other parity path
    sc.t0 = __shfl_sync(full, i0, 31);
    sc.t1 = __shfl_sync(full, i1, 31);
    return sc;
}

// Block summary of a scan (same value in every lane).
__device__ __forceinline__ Tab tab_of(const Scan &sc) {
    const unsigned int full = 0xffffffffu;
    Tab t;
    if (__any_sync(full, sc.big)) { t.a0 = t.a1 = 0; t.mn = -BAD; t.mx = BAD; return t; }
    t.a0 = sc.t0; t.a1 = sc.t1;
    t.mn = __reduce_min_sync(full, sc.e0 + sc.mn) - 1;
    t.mx = __reduce_max_sync(full, sc.e0 + sc.mx) + 1;
    return t;
}

template <int KE>
__device__ __forceinline__ Tab warp_summarize(const float (&x)[KE], float inv_u) {
    return tab_of(warp_scan<KE>(x, inv_u));
}

// Runs the block through the chain starting from the exact running sum acc when its summary was not applicable.
// Attempt: summarise the remaining lanes in the binade of acc, apply the longest valid lane prefix in one step, run the
// lanes around the point where the sum leaves its binade as a plain sequential float chain (values broadcast by
// shuffles; SEQ_RUN lanes, because a sum that has just crossed a power of two tends to cross back), try again.  A sum
// that hovers around a power of two defeats every summary, and the plain chain (~5 cycles per element) is then the
// fastest exact evaluation there is: after MAX_RETRY attempts the rest of the block is run sequentially.
// Lanes whose elements are all +0 are skipped (s + 0 == s).  n_seq counts sequentially run lanes.
constexpr int MAX_RETRY = 2;
constexpr int SEQ_RUN = 3;

// sequential chain over the lanes in `lanes` (ascending), shuffles of the next lane overlap the adds of the current one
template <int KE>
__device__ __forceinline__ float warp_seq_lanes(float acc, const float (&x)[KE], unsigned int lanes, int &n_seq) {
    const unsigned int full = 0xffffffffu;
    if (!lanes) return acc;
    int l = __ffs(lanes) - 1;
    lanes &= lanes - 1u;
    float q[KE];
#pragma unroll
    for (int i = 0; i < KE; ++i) q[i] = __shfl_sync(full, x[i], l);
#pragma unroll 1
    for (;;) {
        ++n_seq;
        float qn[KE];
        const int ln = lanes ? (__ffs(lanes) - 1) : -1;
        if (ln >= 0) {
#pragma unroll
            for (int i = 0; i < KE; ++i) qn[i] = __shfl_sync(full, x[i], ln);
        }
#pragma unroll
        for (int i = 0; i < KE; ++i) acc = acc + q[i];
        if (ln < 0) break;
        lanes &= lanes - 1u;
#pragma unroll
        for (int i = 0; i < KE; ++i) q[i] = qn[i];
    }
    return acc;
}

// The plain sequential float chain over a whole block (lanes l0 .. 31), staged through a warp-private shared buffer of
// 32 * KE floats so that every add is fed by broadcast LDS.128 loads issued 16 elements ahead: ~4.5 cycles per element,
// the speed of the dependent FADD chain itself.  A lone warp cannot do better on a block in which the running sum
// keeps crossing a power of two (its instruction latency makes every summarise-and-check attempt cost more than that).
template <int KE>
__device__ __forceinline__ float warp_seq_block(float acc, const float (&x)[KE], float *wbuf, int &n_seq, int l0 = 0) {
    static_assert(KE % 8 == 0, "warp_seq_block: 8 elements per step");
    const int lane = threadIdx.x & 31;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < KE; i += 4)
        *reinterpret_cast<float4 *>(wbuf + lane * KE + i) = make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]);
    __syncwarp();
    const float4 *p = reinterpret_cast<const float4 *>(wbuf + l0 * KE);
    const int n8 = (32 - l0) * (KE / 8);                 // steps of 8 elements
    float4 a0 = p[0], a1 = p[1];
#pragma unroll 4
    for (int j = 0; j < n8; ++j) {
        float4 b0 = a0, b1 = a1;
        if (j + 1 < n8) { b0 = p[2 * j + 2]; b1 = p[2 * j + 3]; }
        acc = acc + a0.x; acc = acc + a0.y; acc = acc + a0.z; acc = acc + a0.w;
        acc = acc + a1.x; acc = acc + a1.y; acc = acc + a1.z; acc = acc + a1.w;
        a0 = b0; a1 = b1;
    }
    n_seq += 32 - l0;
    __syncwarp();
    return acc;
}

template <int KE>
__device__ __forceinline__ float warp_advance(float acc, const float (&x)[KE], int &n_seq) {
    const unsigned int full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    bool nz = false;
#pragma unroll
    for (int i = 0; i < KE; ++i) nz |= (x[i] != 0.0f) || (x[i] != x[i]);
    unsigned int todo = __ballot_sync(full, nz);         // lanes that still have to be consumed
#pragma unroll 1
    for (int iter = 0; iter < MAX_RETRY && todo; ++iter) {
        float inv_u, u;
        if (epoch_of(acc, inv_u, u)) {
            float xm[KE];
#pragma unroll
            for (int i = 0; i < KE; ++i) xm[i] = ((todo >> lane) & 1u) ? x[i] : 0.0f;
            const Scan sc = warp_scan<KE>(xm, inv_u);
            const int m = (int)(acc * inv_u);
            const int lo = (1 << 23) + MARGIN, hi = (1 << 24) - MARGIN;
            const int b = m + sc.e0;
            const bool okl = !sc.big && (m > 0 ? (b + sc.mn > lo && b + sc.mx < hi) : (b + sc.mx < -lo && b + sc.mn > -hi));
            const unsigned int bad = __ballot_sync(full, !okl) & todo;
            const int lstar = bad ? (__ffs(bad) - 1) : 32;
            const int src = lstar < 32 ? lstar : 31;
            int inc0 = __shfl_sync(full, sc.e0, src), inc1 = __shfl_sync(full, sc.e1, src);
            if (lstar == 32) { inc0 = sc.t0; inc1 = sc.t1; }
            const int inc = (m & 1) ? inc1 : inc0;
            if (inc != 0) acc = (float)(m + inc) * u;     // every lane before lstar is valid for this m
            todo = lstar < 32 ? (todo & (0xffffffffu << lstar)) : 0u;
            if (!todo) break;
        }
        // the next SEQ_RUN pending lanes as a plain chain
        unsigned int run = 0u, rest = todo;
#pragma unroll
        for (int k = 0; k < SEQ_RUN; ++k) { const unsigned int low = rest & (0u - rest); run |= low; rest &= ~low; }
        acc = warp_seq_lanes<KE>(acc, x, run, n_seq);
        todo = rest;
    }
    return warp_seq_lanes<KE>(acc, x, todo, n_seq);
}

// Walks the per-sub-block summaries tab[0..nsub) (nsub <= 32, all computed for the binade inv_a) of one chain in
// order, in parallel: lane w composes the tables of sub-blocks 0..w, the longest prefix that is valid for the
// actual running sum is applied in one step.  Returns the number of sub-blocks consumed (== nsub if all were
// valid); the caller advances the next one piecewise (warp_advance) and calls again with `first` moved on.
__device__ __forceinline__ int warp_compose(float &acc, float inv_a, const int4 *tab, int first, int nsub) {
    const unsigned int full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    float inv_u, u;
    if (inv_a == 0.0f || !epoch_of(acc, inv_u, u) || inv_u != inv_a) return first;
    const int w = first + lane;
    int4 q = make_int4(0, 0, 0, 0);
    if (w < nsub) q = tab[w];
    int i0 = q.x, i1 = q.y;
    if (!__any_sync(full, i0 != i1)) {                    // no exact tie anywhere: plain prefix sums
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int g0 = __shfl_up_sync(full, i0, off);
            if (lane >= off) i0 += g0;
        }
        i1 = i0;
    } else {
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int g0 = __shfl_up_sync(full, i0, off), g1 = __shfl_up_sync(full, i1, off);
            if (lane >= off) {
                const int n0 = g0 + ((g0 & 1) ? i1 : i0);
                const int n1 = g1 + (((g1 + 1) & 1) ? i1 : i0);
                i0 = n0; i1 = n1;
            }
        }
    }
    int e0 = __shfl_up_sync(full, i0, 1);
    if (lane == 0) e0 = 0;
    const int m = (int)(acc * inv_u);
    const int lo = (1 << 23) + MARGIN, hi = (1 << 24) - MARGIN;
    const int b = m + e0;                                 // path p = 0; the other path differs by at most 1 (MARGIN)
    const bool okw = (w >= nsub) || (m > 0 ? (b + q.z > lo && b + q.w < hi) : (b + q.w < -lo && b + q.z > -hi));
    const unsigned int bad = __ballot_sync(full, !okw);
    const int k = bad ? (__ffs(bad) - 1) : 32;            // sub-blocks first .. first+k-1 are valid
    const int take = min(k, nsub - first);
    if (take <= 0) return first;
    const int src = take - 1;
    const int inc0 = __shfl_sync(full, i0, src), inc1 = __shfl_sync(full, i1, src);
    const int inc = (m & 1) ? inc1 : inc0;
    acc = (float)(m + inc) * u;
    return first + take;
}

// ---------------------------------------------------------------- one stage of up to NCH chains, whole CTA
// A stage is NW*32*KE chain elements per chain, resident in shared memory; sub-block w (32*KE elements) belongs to
// warp w, and load_x(c, w, x) hands the calling lane its KE consecutive elements of chain c in sub-block w.
// Rounds: (A) every warp summarises its sub-block for every unfinished chain in that chain's current binade;
// (B) warp c walks chain c: summaries valid for the actual running sum are applied (warp_compose), a sub-block in
// which the sum leaves its binade is advanced piecewise (warp_advance); when the binade has changed and at least two
// sub-blocks remain, the chain asks for another round so that the rest is re-summarised by all warps in parallel.
// Must be called by every thread of the CTA (it contains barriers).  state[c] is the exact running sum of chain c.
struct StageShared {
    float state[8];        // exact running sum of every chain
    float invu[2][8];      // binade (inv_u, 0 = none) the summaries in tab[parity] were computed for
    float snap[8];         // binade of the chain when the previous stage ended (used to pre-summarise the next stage)
    int next[8];           // first sub-block of the chain that is not consumed yet
    int more[2];           // "another round" flag, double-buffered by round parity
    int unit;              // work counter of the pre-summarisation units
    int4 tab[2][8][16];
};

template <int KE, class LoadX>
__device__ __forceinline__ void run_stage(StageShared &sh, int nch, int nsub, LoadX load_x, int &n_fast, int &n_slow, int &n_seq) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < nch) sh.next[tid] = 0;
    __syncthreads();
#pragma unroll 1
    for (int round = 0;; ++round) {
        // ---- A
        if (tid == 0) sh.more[round & 1] = 0;
        if (warp < nsub) {
#pragma unroll 1
            for (int c = 0; c < nch; ++c) {
                if (warp < sh.next[c]) continue;
                float inv_u, u;
                const bool ok = epoch_of(sh.state[c], inv_u, u);
                if (warp == sh.next[c] && lane == 0) sh.invu[0][c] = ok ? inv_u : 0.0f;
                if (ok) {
                    float x[KE];
                    load_x(c, warp, x);
                    const Tab tb = warp_summarize<KE>(x, inv_u);
                    if (lane == 0) sh.tab[0][c][warp] = make_int4(tb.a0, tb.a1, tb.mn, tb.mx);
                }
            }
        }
        __syncthreads();
        // ---- B
        if (warp < nch) {
            const int c = warp;
            int w = sh.next[c];
            if (w < nsub) {
                float acc = sh.state[c];
                const float inv_a = sh.invu[0][c];
#pragma unroll 1
                while (w < nsub) {
                    const int w2 = warp_compose(acc, inv_a, sh.tab[0][c], w, nsub);
                    n_fast += w2 - w;
                    w = w2;
                    if (w >= nsub) break;
                    float inv_u, u;
                    const bool stale = !epoch_of(acc, inv_u, u) || inv_u != inv_a;
                    if (stale && inv_a != 0.0f && nsub - w >= 2 && epoch_of(acc, inv_u, u)) break;   // new binade: re-summarise in parallel
                    float x[KE];
                    load_x(c, w, x);
                    acc = warp_advance<KE>(acc, x, n_seq);
                    ++n_slow; ++w;
                    if (inv_a == 0.0f && nsub - w >= 2 && epoch_of(acc, inv_u, u)) break;            // the chain has a binade now
                }
                if (lane == 0) { sh.state[c] = acc; sh.next[c] = w; if (w < nsub) sh.more[round & 1] = 1; }
            }
        }
        __syncthreads();
        if (!sh.more[round & 1]) break;
    }
}

// Software-pipelined variant: while warp c walks chain c through the CURRENT stage (summaries in tab[par]), all warps
// pre-summarise the NEXT stage into tab[par ^ 1] for the binade each chain had when the previous stage ended
// (snap[]; one stage stale -- if the chain has moved to another binade since, the walk notices the tag mismatch and
// asks for a re-summarisation round, exactly as run_stage does).  Call protocol, all threads of the CTA:
//     pipe_init(sh);  __syncthreads();
//     for every stage st:  run_stage_pipe(sh, nch, st & 1, nsub(st), load_x(st), nsub(st + 1) or 0, load_x(st + 1));
//                          ... (anything that does not touch sh) ...;  __syncthreads();
__device__ __forceinline__ void pipe_init(StageShared &sh) {
    const int tid = threadIdx.x;
    if (tid < 8) { sh.state[tid] = 0.0f; sh.invu[0][tid] = 0.0f; sh.invu[1][tid] = 0.0f; sh.snap[tid] = 0.0f; sh.next[tid] = 0; }
    if (tid == 0) { sh.more[0] = 0; sh.more[1] = 0; sh.unit = 0; }
}

template <int KE, class LoadCur, class LoadNext>
__device__ __forceinline__ void run_stage_pipe(StageShared &sh, int nch, int par, int nsub, LoadCur load_cur, int nsub_next,
                                               LoadNext load_next, int &n_fast, int &n_slow, int &n_seq) {
    const unsigned int full = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // the walk of chain c over the current stage (warp c); leaves next[c] < nsub and raises *more when it wants the
    // rest of the stage re-summarised for the chain's new binade
    auto walk = [&](int *more) {
        const int c = warp;
        int w = sh.next[c];
        if (w >= nsub) return;
        float acc = sh.state[c];
        const float inv_a = sh.invu[par][c];
        int mism = 0;                                                   // consecutive sub-blocks met in another binade
#pragma unroll 1
        while (w < nsub) {
            const int w2 = warp_compose(acc, inv_a, sh.tab[par][c], w, nsub);
            n_fast += w2 - w;
            if (w2 > w) mism = 0;
            w = w2;
            if (w >= nsub) break;
            float inv_u, u;
            const bool has = epoch_of(acc, inv_u, u);
            // In another binade than the summaries.  A sum that hovers around a power of two flips back and forth, and a
            // re-summarisation round costs more than advancing a sub-block piecewise: ask for the round only when the
            // new binade has persisted (or when there are no summaries at all yet).
            if (has && inv_u != inv_a) {
                ++mism;
                if ((mism >= 2 || inv_a == 0.0f) && nsub - w >= 3) break;
            } else mism = 0;
            float x[KE];
            load_cur(c, w, x);
            acc = warp_advance<KE>(acc, x, n_seq);
            ++n_slow; ++w;
        }
        if (lane == 0) { sh.state[c] = acc; sh.next[c] = w; if (w < nsub) *more = 1; }
    };
    // ---- phase 1: walk (chain warps) || pre-summarise the next stage (everybody, dynamic units)
    if (warp < nch) walk(&sh.more[0]);
    {
        const int nunits = nch * nsub_next;
#pragma unroll 1
        for (;;) {
            int uidx = 0;
            if (lane == 0) uidx = atomicAdd(&sh.unit, 1);
            uidx = __shfl_sync(full, uidx, 0);
            if (uidx >= nunits) break;
            const int w = uidx / nch, c = uidx - w * nch;
            const float inv_u = sh.snap[c];
            if (inv_u != 0.0f) {
                float x[KE];
                load_next(c, w, x);
                const Tab tb = warp_summarize<KE>(x, inv_u);
                if (lane == 0) sh.tab[par ^ 1][c][w] = make_int4(tb.a0, tb.a1, tb.mn, tb.mx);
            }
        }
    }
    __syncthreads();
    // ---- phase 2: re-summarisation rounds for the current stage
#pragma unroll 1
    for (int round = 0; sh.more[round & 1]; ++round) {
        if (tid == 0) sh.more[(round + 1) & 1] = 0;
        if (warp < nsub) {
#pragma unroll 1
            for (int c = 0; c < nch; ++c) {
                const int nx = sh.next[c];
                if (warp < nx || nx >= nsub) continue;
                float inv_u, u;
                const bool ok = epoch_of(sh.state[c], inv_u, u);
                if (warp == nx && lane == 0) sh.invu[par][c] = ok ? inv_u : 0.0f;
                if (ok) {
                    float x[KE];
                    load_cur(c, warp, x);
                    const Tab tb = warp_summarize<KE>(x, inv_u);
                    if (lane == 0) sh.tab[par][c][warp] = make_int4(tb.a0, tb.a1, tb.mn, tb.mx);
                }
            }
        }
        __syncthreads();
        if (warp < nch) walk(&sh.more[(round + 1) & 1]);
        __syncthreads();
    }
    // ---- hand-over to the next stage (the caller's barrier publishes it)
    if (tid < nch) {
        float inv_u, u;
        sh.invu[par ^ 1][tid] = sh.snap[tid];
        sh.snap[tid] = epoch_of(sh.state[tid], inv_u, u) ? inv_u : 0.0f;
        sh.next[tid] = 0;
    }
    if (tid == 0) { sh.unit = 0; sh.more[0] = 0; }
}

// Applies a block summary to the running sum if the block stays inside the epoch for this s; false = not applied.
__device__ __forceinline__ bool apply(const Tab &t, float &s, float inv_u, float u) {
    const int m = (int)(s * inv_u);                      // exact integer, 2^23 <= |m| < 2^24
    const int lo = (1 << 23) + MARGIN, hi = (1 << 24) - MARGIN;
    const bool ok = m > 0 ? (m + t.mn > lo && m + t.mx < hi) : (m + t.mx < -lo && m + t.mn > -hi);
    if (!ok) return false;
    const int m2 = m + ((m & 1) ? t.a1 : t.a0);
    s = (float)m2 * u;
    return true;
}

}  // namespace seq
}  // namespace gb
