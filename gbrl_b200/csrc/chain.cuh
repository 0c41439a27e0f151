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

// One warp; lane l holds KE CONSECUTIVE chain elements x[0..KE) (lanes in chain order), 0 for "not a member".
template <int KE>
__device__ __forceinline__ Tab warp_summarize(const float (&x)[KE], float inv_u) {
    const unsigned int full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int d0[KE], d1[KE];
    bool anytie = false, anybig = false;
#pragma unroll
    for (int i = 0; i < KE; ++i) {
        const float y = x[i] * inv_u;                    // exact (power of two), may overflow to inf -> big
        const float d = rintf(y);
        const float fr = y - d;                          // exact
        const bool big = !(fabsf(y) < 4194304.0f);       // 2^22 (also inf / NaN): never summarised
        const bool tie = fabsf(fr) == 0.5f;
        const int di = big ? 0 : (int)d;
        const int k = di - ((tie && fr < 0.0f) ? 1 : 0); // floor(y) of a tie
        d0[i] = tie ? (k + (k & 1)) : di;                // m even: the even one of {m+k, m+k+1}
        d1[i] = tie ? (k + ((k + 1) & 1)) : di;          // m odd
        anytie |= tie; anybig |= big;
    }
    Tab t;
    if (__any_sync(full, anybig)) { t.a0 = t.a1 = 0; t.mn = -BAD; t.mx = BAD; return t; }
    if (!__any_sync(full, anytie)) {
        int a = 0, mn = 0, mx = 0;
#pragma unroll
        for (int i = 0; i < KE; ++i) { a += d0[i]; mn = min(mn, a); mx = max(mx, a); }
        int inc = a;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(full, inc, off);
            if (lane >= off) inc += v;
        }
        const int exc = inc - a;
        t.mn = __reduce_min_sync(full, exc + mn);
        t.mx = __reduce_max_sync(full, exc + mx);
        t.a0 = t.a1 = __shfl_sync(full, inc, 31);
        return t;
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
    int e0 = __shfl_up_sync(full, i0, 1);
    if (lane == 0) e0 = 0;
    t.mn = __reduce_min_sync(full, e0 + mn) - 1;
    t.mx = __reduce_max_sync(full, e0 + mx) + 1;
    t.a0 = __shfl_sync(full, i0, 31);
    t.a1 = __shfl_sync(full, i1, 31);
    return t;
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
