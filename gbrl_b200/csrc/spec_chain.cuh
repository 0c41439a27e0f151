// spec_chain.cuh -- bit-exact evaluation of a sequential fp32 summation chain by SPECULATIVE GROUP SIMULATION.
//
// The reference sums floats one after the other, s <- fl(s + x_i) (node.cpp:336-352, math_ops.cpp:255-300,
// math_ops.h:432-449).  The chain is cut into groups of G elements.  For every group an (approximate) start value p is
// predicted from exact prefix sums, and one lane SIMULATES the group's chain from a candidate start c near p with plain
// float adds -- 32 groups per warp, all groups of all chains at once.  The simulated end value is reusable for the ACTUAL
// start a = c + delta because rounding to a lattice commutes with translations of that lattice:
//
//     if every intermediate result s_k of the simulated chain stays strictly inside its binade by more than |delta|
//     (margin), and delta is a multiple of the coarsest ulp any result was rounded to (lattice; twice that where a
//     rounding was an exact tie, because round-half-even looks at the parity), then the actual chain is the simulated
//     one shifted by delta at every step:  end_actual = end_sim + delta  (exactly).
//
// delta is only known up to ~sqrt(n) ulps in advance, but only its residue matters: candidates c_j = c_0 + j*ulp(p),
// j = 0..J-1, are simulated for the groups whose lattice is coarser than ulp(p) (the sum rises into a higher binade or
// meets a tie); the walk picks j = delta/ulp mod J, and the remainder -- a multiple of J*ulp(p) -- is on every lattice up
// to log2(J) binades above p.  One warp per chain then WALKS the group records: an exclusive prefix sum of the chosen
// increments (fp64: differences of floats, exact) gives every group's actual start, all 32 groups of a window are
// verified at once, and the (rare) group whose conditions fail for the actual start is run as the plain sequential
// chain, which is right by definition.  Nothing that is predicted ever enters the result: a wrong prediction costs time,
// never bits.  tests/test_gpu_chain.py checks this against a sequential accumulation bit for bit.
#pragma once
#include "common.cuh"
#include <cfloat>

namespace gb {
namespace spec {

constexpr int J = 8;                       // candidate starts per group (power of two)
constexpr uint32_t F_NOSIM = 1u << 8;      // no usable prediction (zero / tiny / non-finite): always run sequentially
constexpr uint32_t F_CANDS = 1u << 9;      // candidates 1..J-1 were simulated as well
constexpr float MARGIN_INF = 1.0e38f;

// group record: head (candidate 0) + J-1 further candidates
struct Head {
    float c0;          // candidate 0: p with the low log2(J) mantissa bits cleared
    float end0;        // simulated end value from c0
    float margin0;     // |delta| must be < margin0  (0 = only delta == 0 is usable)
    uint32_t info;     // bits 0-7: lattice exponent el0 (delta must be a multiple of 2^(el0-150)); F_* flags
};
struct Cand { float end; float margin_el; };      // margin with its low 8 mantissa bits replaced by the lattice exponent

__device__ __forceinline__ float pack_margin(float margin, int el) {
    if (!(margin > 0.0f)) margin = 0.0f;
    if (margin > MARGIN_INF) margin = MARGIN_INF;
    return __uint_as_float((__float_as_uint(margin) & 0xffffff00u) | (uint32_t)(el & 0xff));     // rounds the margin DOWN
}
__device__ __forceinline__ float unpack_margin(float packed, int &el) {
    const uint32_t b = __float_as_uint(packed);
    el = (int)(b & 0xffu);
    return __uint_as_float(b & 0xffffff00u);
}

// candidate 0 and the signed ulp of the prediction p; false: no simulation possible
__device__ __forceinline__ bool cand_base(float p, float &c0, float &us, int &pe) {
    const uint32_t pb = __float_as_uint(p);
    pe = (int)((pb >> 23) & 0xffu);
    if (pe < 27 || pe == 255) return false;
    c0 = __uint_as_float(pb & ~(uint32_t)(J - 1));
    us = __uint_as_float((pb & 0x80000000u) | ((uint32_t)(pe - 23) << 23));
    return true;
}

// running state of one simulated chain
struct Sim {
    float s;           // running sum
    uint32_t mx;       // max |s_k| bits over the results (-> max exponent)
    float margin;      // min over results of (distance to the nearer binade edge - 2 ulp)
    int etie;          // max over exact-tie steps of (exponent + 1)
    __device__ __forceinline__ void init(float start) { s = start; mx = 0u; margin = MARGIN_INF; etie = 0; }
    __device__ __forceinline__ void step(float x) {
        const float a = s;
        const float r = a + x;
        const float bb = r - a;
        const float err = (a - (r - bb)) + (x - bb);                 // TwoSum: a + x == r + err exactly
        const uint32_t ar = __float_as_uint(r) & 0x7fffffffu;
        mx = max(mx, ar);
        const uint32_t ex = ar >> 23;
        const uint32_t f = ar & 0x7fffffu;
        const float u = __uint_as_float((ex >= 24u ? ex - 23u : 1u) << 23);   // ulp(r)
        const int du = (int)min(f, 0x800000u - f) - 2;
        float m = (float)du * u;
        if (ex < 24u || ex == 255u) m = -1.0f;                          // zero / tiny / inf / NaN result: nothing can be shifted
        margin = fminf(margin, m);
        if (fabsf(err) == 0.5f * u && x != 0.0f) etie = max(etie, (int)ex + 1);
        s = r;
    }
    __device__ __forceinline__ int lattice_exp() const { return max((int)(mx >> 23), etie); }
};

// ---------------------------------------------------------------- simulation of one group (one lane)
// elem(k), k in [0, cnt): the group's chain elements in order (+0 for "not a member").  Writes the head and, when the
// lattice of candidate 0 is coarser than ulp(p), the other J-1 candidates.
template <class Elem>
__device__ __forceinline__ void sim_group(float p, int cnt, const Elem &elem, Head &hd, Cand *cands /* [J] , entry 0 unused */) {
    float c0 = 0.0f, us = 0.0f;
    int pe = 0;
    if (!cand_base(p, c0, us, pe)) { hd.c0 = 0.0f; hd.end0 = 0.0f; hd.margin0 = 0.0f; hd.info = F_NOSIM; return; }
    Sim a;
    a.init(c0);
    bool any = false;
    for (int k = 0; k < cnt; ++k) {
        const float x = elem(k);
        any |= !(x == 0.0f);
        a.step(x);
    }
    hd.c0 = c0;
    if (!any) {                                   // s + 0 == s for every s: the group is the identity on any start
        hd.end0 = c0; hd.margin0 = MARGIN_INF; hd.info = 0u;
        return;
    }
    const int el0 = a.lattice_exp();
    hd.end0 = a.s; hd.margin0 = a.margin > 0.0f ? a.margin : 0.0f; hd.info = (uint32_t)(el0 & 0xff);
    if (el0 <= pe) return;                        // every multiple of ulp(p) is on candidate 0's lattice already
    hd.info |= F_CANDS;
    Sim c[J - 1];
#pragma unroll
    for (int j = 1; j < J; ++j) c[j - 1].init(c0 + (float)j * us);      // exact: same binade as c0 (low bits were cleared)
    for (int k = 0; k < cnt; ++k) {
        const float x = elem(k);
#pragma unroll
        for (int j = 0; j < J - 1; ++j) c[j].step(x);                    // J-1 independent chains: the FADD latency is hidden by ILP
    }
#pragma unroll
    for (int j = 1; j < J; ++j) { cands[j].end = c[j - 1].s; cands[j].margin_el = pack_margin(c[j - 1].margin, c[j - 1].lattice_exp()); }
}

// ---------------------------------------------------------------- helpers of the walk
__device__ __forceinline__ double shfl_d(double v, int src) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(0xffffffffu, lo, src); hi = __shfl_sync(0xffffffffu, hi, src);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_up_d(double v, int d) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_up_sync(0xffffffffu, lo, d); hi = __shfl_up_sync(0xffffffffu, hi, d);
    return __hiloint2double(hi, lo);
}
// 2^k as a double, k in [-1022, 1023]
__device__ __forceinline__ double pow2d(int k) { return __hiloint2double((k + 1023) << 20, 0); }

// is (start = c + delta) a valid shift of the chain simulated from c?   margin / lattice as recorded by the simulation
__device__ __forceinline__ bool shift_ok(double delta, float margin, int el) {
    if (delta == 0.0) return true;
    if (!(fabs(delta) < (double)margin)) return false;
    if (el == 0) return true;                               // a group without elements: the identity on every start
    if (el < 24 || el > 254) return false;
    const double q = delta * pow2d(150 - el);               // delta / 2^(el-150), exact (power of two)
    return q == rint(q) && fabs(q) < 4.0e15;
}

// The walk of ONE chain by one warp.
//   ng                 number of groups
//   load_head(g)       -> Head of group g           (called by lane g - w0 for the 32 groups of a window)
//   load_cand(g, j)    -> Cand j (1..J-1) of group g (only for groups with F_CANDS)
//   seq_group(g, a)    -> runs group g as the plain sequential chain from the exact running sum a; WARP-collective
// Returns the exact final sum (same value in every lane).  Counters: groups run sequentially; internal inconsistencies
// (a running sum that is not a float -- cannot happen while the fp64 prefix is exact; tests assert it stays 0).
template <class LoadHead, class LoadCand, class SeqGroup>
__device__ __forceinline__ float walk_chain(int ng, float a_start, LoadHead load_head, LoadCand load_cand, SeqGroup seq_group,
                                            int &n_err, int &n_seq) {
    const unsigned int full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    double base = (double)a_start;                      // exact running sum at the first live group of the window
#pragma unroll 1
    for (int w0 = 0; w0 < ng; w0 += 32) {
        const int g = w0 + lane;
        const bool in_range = g < ng;
        Head hd;
        hd.c0 = 0.0f; hd.end0 = 0.0f; hd.margin0 = MARGIN_INF; hd.info = 0u;
        if (in_range) hd = load_head(g);
        Cand cd[J];
        const bool has_cands = in_range && (hd.info & F_CANDS);
        if (has_cands) {
#pragma unroll
            for (int j = 1; j < J; ++j) cd[j] = load_cand(g, j);
        }
        // the chosen candidate of this lane's group (0 until the walk says otherwise)
        float c_sel = hd.c0, margin_sel = hd.margin0;
        int el_sel = (int)(hd.info & 0xffu);
        double inc_sel = in_range ? ((double)hd.end0 - (double)hd.c0) : 0.0;
        bool nosim = in_range && (hd.info & F_NOSIM);
        bool settled = false;                          // this lane's candidate choice is final
        int live_from = 0;                             // lanes before it are consumed
        // exactness guard of the fp64 prefix: exponent spread of everything that is added must stay below 2^29
        {
            int emin = 255, emax = 0;
            if (in_range && !nosim) {
                const int e1 = (int)((__float_as_uint(hd.c0) >> 23) & 0xff), e2 = (int)((__float_as_uint(hd.end0) >> 23) & 0xff);
                emin = min(e1, e2 > 0 ? e2 : e1); emax = max(e1, e2);
                if (has_cands) {
#pragma unroll
                    for (int j = 1; j < J; ++j) { const int e3 = (int)((__float_as_uint(cd[j].end) >> 23) & 0xff); emax = max(emax, e3); if (e3 > 0) emin = min(emin, e3); }
                }
            }
            emin = __reduce_min_sync(full, emin); emax = __reduce_max_sync(full, emax);
            const int eb = (int)((__double2hiint(base) >> 20) & 0x7ff) - 1023 + 127;
            if (base != 0.0) { emax = max(emax, eb); emin = min(emin, eb); }
            if (emax - emin > 28) nosim = in_range;     // pathological dynamic range: run the window sequentially
        }
#pragma unroll 1
        for (;;) {
            const bool live = in_range && lane >= live_from;
            // exclusive prefix of the chosen increments over the live lanes
            double incv = live ? inc_sel : 0.0, pre = incv;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const double v = shfl_up_d(pre, off);
                if (lane >= off) pre += v;
            }
            const double tot = shfl_d(pre, 31);
            const double A = base + (pre - incv);       // exact running sum at the start of this lane's group
            bool ok = true;
            if (live && !settled) ok = !nosim && shift_ok(A - (double)c_sel, margin_sel, el_sel) && ((double)(float)A == A);
            const unsigned int bad = __ballot_sync(full, !ok);
            if (!bad) { base += tot; break; }
            const int F = __ffs(bad) - 1;
            // lane F: is one of the other candidates usable for the actual start?
            int found = 0;
            if (lane == F && has_cands && !nosim && ((double)(float)A == A)) {
                int pe; float c0, us;
                cand_base(hd.c0, c0, us, pe);                              // hd.c0 has its low bits cleared already: c0 == hd.c0
                const double D = (A - (double)hd.c0) / (double)us;
                if (D == rint(D) && fabs(D) < 1.0e15) {
                    const int j = (int)(((long long)D) & (long long)(J - 1));
                    if (j != 0) {
                        Cand cj = cd[1];
#pragma unroll
                        for (int t = 2; t < J; ++t) if (j == t) cj = cd[t];
                        int el;
                        const float mg = unpack_margin(cj.margin_el, el);
                        const float cjv = hd.c0 + (float)j * us;
                        if (shift_ok(A - (double)cjv, mg, el)) {
                            c_sel = cjv; margin_sel = mg; el_sel = el; inc_sel = (double)cj.end - (double)cjv;
                            found = 1;
                        }
                    }
                }
            }
            found = __shfl_sync(full, found, F);
            if (found) { if (lane == F) settled = true; continue; }
            // group F is run as the plain sequential chain from its exact start
            const double AF = shfl_d(A, F);
            const float aF = (float)AF;                 // exact: A of the first failing lane is built from verified groups only
            if ((double)aF != AF) ++n_err;
            const float a_next = seq_group(w0 + F, aF);
            ++n_seq;
            base = (double)a_next;
            live_from = F + 1;
            if (w0 + live_from >= ng || live_from >= 32) break;
        }
    }
    if ((double)(float)base != base) ++n_err;
    return (float)base;
}

}  // namespace spec
}  // namespace gb
