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
// flags in the low byte of Head::mflags
constexpr uint32_t F_NOSIM = 1u;           // no usable prediction (zero / tiny / non-finite): always run sequentially
constexpr uint32_t F_PFREE = 2u;           // candidate 0 serves every start on ulp(p)'s lattice (its own lattice is not coarser)
constexpr uint32_t F_SIMPLE = 4u;          // candidates 0 / 1 (even / odd start) serve every start: lattice <= 2 ulp(p)
constexpr uint32_t F_CANDS = 8u;           // the per-candidate records cand[0 .. J) were written
constexpr float MARGIN_INF = 1.0e38f;

// group record.  The walk's fast path only needs the head: with a = c0 + D * ulp(p) the group maps
//     a -> a + inc0            (D even, or F_PFREE)          inc0 = end0 - c0
//     a -> a + inc0 + d1       (D odd)                        inc1 = end1 - (c0 + ulp)
// provided |D - parity| * ulp < margin.  Everything else goes through the per-candidate records.
struct Head {
    float c0;          // candidate 0: p with the low log2(J) mantissa bits cleared
    float end0;        // simulated end value from c0
    float d1;          // inc1 - inc0 (exact)
    float mflags;      // min margin of candidates 0 / 1, low 8 mantissa bits replaced by the F_* flags
};
struct Cand { float end; float margin_el; };      // margin with its low 8 mantissa bits replaced by the lattice exponent
__device__ __forceinline__ uint32_t head_flags(const Head &h) { return __float_as_uint(h.mflags) & 0xffu; }

__device__ __forceinline__ float pack_margin(float margin, int low8) {
    if (!(margin > 0.0f)) margin = 0.0f;
    if (margin > MARGIN_INF) margin = MARGIN_INF;
    return __uint_as_float((__float_as_uint(margin) & 0xffffff00u) | (uint32_t)(low8 & 0xff));     // rounds the margin DOWN
}
__device__ __forceinline__ float unpack_margin(float packed, int &low8) {
    const uint32_t b = __float_as_uint(packed);
    low8 = (int)(b & 0xffu);
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
// lattice of candidate 0 is coarser than ulp(p), the per-candidate records.
template <class Elem>
__device__ __forceinline__ void sim_group(float p, int cnt, const Elem &elem, Head &hd, Cand *cands /* [J] */) {
    float c0 = 0.0f, us = 0.0f;
    int pe = 0;
    if (!cand_base(p, c0, us, pe)) { hd.c0 = 0.0f; hd.end0 = 0.0f; hd.d1 = 0.0f; hd.mflags = pack_margin(0.0f, F_NOSIM); return; }
    Sim a;
    a.init(c0);
    bool any = false;
    for (int k = 0; k < cnt; ++k) {
        const float x = elem(k);
        any |= !(x == 0.0f);
        a.step(x);
    }
    hd.c0 = c0; hd.d1 = 0.0f;
    if (!any) {                                   // s + 0 == s for every s: the group is the identity on any start
        hd.end0 = c0; hd.mflags = pack_margin(MARGIN_INF, F_PFREE);
        return;
    }
    const int el0 = a.lattice_exp();
    hd.end0 = a.s;
    if (el0 <= pe) { hd.mflags = pack_margin(a.margin, F_PFREE); return; }     // every multiple of ulp(p) is on candidate 0's lattice
    Sim c[J - 1];
#pragma unroll
    for (int j = 1; j < J; ++j) c[j - 1].init(c0 + (float)j * us);      // exact: same binade as c0 (low bits were cleared)
    for (int k = 0; k < cnt; ++k) {
        const float x = elem(k);
#pragma unroll
        for (int j = 0; j < J - 1; ++j) c[j].step(x);                    // J-1 independent chains: the FADD latency is hidden by ILP
    }
    cands[0].end = a.s; cands[0].margin_el = pack_margin(a.margin, el0);
#pragma unroll
    for (int j = 1; j < J; ++j) { cands[j].end = c[j - 1].s; cands[j].margin_el = pack_margin(c[j - 1].margin, c[j - 1].lattice_exp()); }
    // even / odd starts through candidates 0 / 1 alone?  (both lattices at most 2 ulp(p); the increment difference a float)
    const int el1 = c[0].lattice_exp();
    const double inc0 = (double)a.s - (double)c0, inc1 = (double)c[0].s - ((double)c0 + (double)us);
    const float d1 = (float)(inc1 - inc0);
    uint32_t fl = F_CANDS;
    if (el0 <= pe + 1 && el1 <= pe + 1 && (double)d1 == inc1 - inc0) { fl |= F_SIMPLE; hd.d1 = d1; }
    hd.mflags = pack_margin(fminf(a.margin, c[0].margin), fl);
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
//   load_cand(g, j)    -> Cand j (0..J-1) of group g (only for groups with F_CANDS)
//   seq_group(g, a)    -> runs group g as the plain sequential chain from the exact running sum a; WARP-collective
// Returns the exact final sum (same value in every lane).  Counters: groups run sequentially; internal inconsistencies
// (a running sum that is not a float -- cannot happen while the arithmetic below is exact; tests assert it stays 0).
//
// A window of 32 groups is integer arithmetic in units of the finest ulp u of anything in the window (starts, ends, the
// running sum): with M = a / u the group of lane g maps  M -> M + I0_g  (start on an even multiple of its own ulp 2^k_g u,
// or F_PFREE)  or  M -> M + I0_g + dI_g  (odd multiple).  c0_g has its low bits cleared, so the parity is bit k_g of M.
//   1. inclusive prefix of I0 over the lanes (int64 warp scan) = every group's incoming M if all parities were even;
//   2. the parity-dependent groups are visited in order by the whole warp (a few integer instructions and three shuffles
//      each): bit k of (own low bits + correction so far) picks the candidate, its dI joins the correction;
//   3. all lanes check margin and lattice for their actual start at once.
// The first group that does not fit (coarser lattice than 2 ulp, margin, no simulation) is resolved on its own -- candidate
// records in fp64, else the sequential chain -- and the rest of the window is redone behind it.
__device__ __forceinline__ long long shfl_ll(long long v, int src) {
    int lo = (int)(v & 0xffffffffll), hi = (int)(v >> 32);
    lo = __shfl_sync(0xffffffffu, lo, src); hi = __shfl_sync(0xffffffffu, hi, src);
    return ((long long)hi << 32) | (unsigned int)lo;
}
__device__ __forceinline__ long long shfl_up_ll(long long v, int d) {
    int lo = (int)(v & 0xffffffffll), hi = (int)(v >> 32);
    lo = __shfl_up_sync(0xffffffffu, lo, d); hi = __shfl_up_sync(0xffffffffu, hi, d);
    return ((long long)hi << 32) | (unsigned int)lo;
}
// exponent of the ulp of v (biased like the float exponent field; 255 for 0: "any lattice")
__device__ __forceinline__ int ulp_exp(float v) {
    const int e = (int)((__float_as_uint(v) >> 23) & 0xffu);
    return (__float_as_uint(v) & 0x7fffffffu) == 0u ? 255 : e;
}
// v / 2^(emin-150) as an integer (v a multiple of that unit, exponent spread <= 36)
__device__ __forceinline__ long long to_units(float v, int emin) {
    const uint32_t b = __float_as_uint(v);
    const int e = (int)((b >> 23) & 0xffu);
    if ((b & 0x7fffffffu) == 0u) return 0;
    const long long m = (long long)((b & 0x7fffffu) | 0x800000u) << (e - emin);
    return (b >> 31) ? -m : m;
}

template <class LoadHead, class LoadCand, class SeqGroup>
__device__ __forceinline__ float walk_chain(int ng, float a_start, LoadHead load_head, LoadCand load_cand, SeqGroup seq_group,
                                            int &n_err, int &n_seq) {
    const unsigned int full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    float base = a_start;                               // exact running sum at the first live group
    Head nx;
    nx.c0 = 0.0f; nx.end0 = 0.0f; nx.d1 = 0.0f; nx.mflags = 0.0f;
    if (lane < ng) nx = load_head(lane);
#pragma unroll 1
    for (int w0 = 0; w0 < ng; w0 += 32) {
        const int g = w0 + lane;
        const bool in_range = g < ng;
        const Head hd = nx;
        if (w0 + 32 + lane < ng) nx = load_head(w0 + 32 + lane);          // next window's heads are in flight during this one
        int flags;
        const float margin = unpack_margin(hd.mflags, flags);
        const int pe = (int)((__float_as_uint(hd.c0) >> 23) & 0xffu);
        const bool usable = in_range && !(flags & F_NOSIM) && (flags & (F_PFREE | F_SIMPLE)) && pe >= 27;
        const int wn = min(32, ng - w0);
        int live_from = 0;
#pragma unroll 1
        while (live_from < wn) {
            const int L = live_from;
            // ---- the unit: finest ulp of the live lanes' starts / ends / increments and of the running sum
            int e_lo = 255, e_hi = 0;
            if (usable && lane >= L) {
                e_lo = min(min(ulp_exp(hd.c0), ulp_exp(hd.end0)), ulp_exp(hd.d1));
                e_hi = max(pe, (int)((__float_as_uint(hd.end0) >> 23) & 0xffu));
            }
            e_lo = min(e_lo, ulp_exp(base));
            e_hi = max(e_hi, (int)((__float_as_uint(base) >> 23) & 0xffu));
            e_lo = __reduce_min_sync(full, e_lo); e_hi = __reduce_max_sync(full, e_hi);
            const bool span_ok = e_lo >= 24 && e_hi - e_lo <= 32;            // 24 mantissa bits + 32 + 5 (32 lanes) < 63
            bool mine = usable && lane >= L && span_ok;
            long long C0 = 0, I0 = 0, dI = 0;
            int k = 0;
            if (mine) {
                C0 = to_units(hd.c0, e_lo);
                I0 = to_units(hd.end0, e_lo) - C0;
                dI = to_units(hd.d1, e_lo);
                k = pe - e_lo;                                               // own ulp = 2^k units
                if (dI > 0x3fffffffll || dI < -0x3fffffffll || k > 30) mine = false;
            }
            // the run ends at the first lane that is not usable
            const unsigned int stop = __ballot_sync(full, !mine && lane >= L);
            const int R = stop ? (__ffs(stop) - 1) : 32;
            int F = L;
            if (R > L && span_ok) {
                const bool in_run = lane >= L && lane < R;
                long long pre = in_run ? I0 : 0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const long long v = shfl_up_ll(pre, o);
                    if (lane >= o) pre += v;
                }
                const long long baseM = to_units(base, e_lo);
                const long long Minc = baseM + pre - (in_run ? I0 : 0);      // incoming M if every parity before was even
                // ---- parity-dependent lanes, in order
                unsigned int todo = __ballot_sync(full, in_run && !(flags & F_PFREE));
                const unsigned int alow = (unsigned int)(Minc & 0xffffffffll);
                long long corr = 0, mycorr = 0;
                int par = 0;
                while (todo) {
                    const int i = __ffs(todo) - 1;
                    todo &= todo - 1u;
                    const unsigned int ai = __shfl_sync(full, alow, i);
                    const int ki = __shfl_sync(full, k, i);
                    const int di = __shfl_sync(full, (int)dI, i);
                    const int pi = (int)(((ai + (unsigned int)(corr & 0xffffffffll)) >> ki) & 1u);
                    if (lane == i) { par = pi; mycorr = corr; }
                    if (pi) corr += di;
                    if (lane > i) mycorr = corr;
                }
                if (flags & F_PFREE) { par = 0; }
                // lanes after the last parity lane took mycorr = corr above; parity-free lanes before any parity lane keep 0
                const long long M = Minc + mycorr;
                const long long dl = M - C0;                                 // a - c0 in units
                const long long step1 = (__float_as_uint(hd.c0) >> 31) ? -(1ll << k) : (1ll << k);      // candidate 1 = c0 + signed ulp
                const long long dr = dl - (par ? step1 : 0ll);               // remainder after the chosen candidate
                bool ok = true;
                if (in_run) {
                    ok = (dl & ((1ll << k) - 1ll)) == 0ll;                   // the start is on the group's own lattice
                    if (ok && dr != 0) {
                        const long long adr = dr < 0 ? -dr : dr;
                        ok = adr < (1ll << 40) && (float)adr * __uint_as_float((uint32_t)(e_lo - 23) << 23) < margin;      // unit = 2^(e_lo - 150)
                    }
                }
                const unsigned int bad = __ballot_sync(full, !ok);
                F = bad ? (__ffs(bad) - 1) : R;
                if (F > L) {
                    // lanes L .. F-1 are applied: the running sum behind lane F-1
                    const long long Mout = M + I0 + (par ? dI : 0);
                    const long long Mo = shfl_ll(Mout, F - 1);
                    const double v = (double)Mo * pow2d(e_lo - 150);
                    base = (float)v;
                    if ((double)base != v) ++n_err;
                }
            }
            live_from = F;
            if (F >= wn) break;
            // ---- group F on its own: candidate records in fp64, else the sequential chain
            float nb = 0.0f;
            int how = 0;                                                           // 1: resolved from a record
            if (lane == F && in_range && !(flags & F_NOSIM)) {
                float c0, us;
                int pe2;
                if (cand_base(hd.c0, c0, us, pe2)) {
                    const double delta = (double)base - (double)hd.c0;
                    if (flags & F_PFREE) {
                        if (shift_ok(delta, margin, pe2)) { const double v = (double)hd.end0 + delta; nb = (float)v; how = ((double)nb == v); }
                    } else if (flags & F_CANDS) {
                        const double D = delta * (1.0 / (double)us);
                        if (D == rint(D) && fabs(D) < 1.0e15) {
                            const int j = (int)(((long long)D) & (long long)(J - 1));
                            const Cand cj = load_cand(w0 + F, j);
                            int el;
                            const float mg = unpack_margin(cj.margin_el, el);
                            const float cjv = hd.c0 + (float)j * us;
                            const double dj = (double)base - (double)cjv;
                            if (shift_ok(dj, mg, el)) { const double v = (double)cj.end + dj; nb = (float)v; how = ((double)nb == v); }
                        }
                    }
                }
            }
            how = __shfl_sync(full, how, F);
            if (how) base = __shfl_sync(full, nb, F);
            else { base = seq_group(w0 + F, base); ++n_seq; }
            live_from = F + 1;
        }
    }
    return base;
}

}  // namespace spec
}  // namespace gb
