"""CPU suite: the arithmetic behind gbrl_b200/csrc/spec_chain.cuh (speculative group simulation of a sequential fp32
chain), restated in numpy float32 and checked against a plain sequential accumulation.

A group of chain elements is simulated from a candidate start c next to a PREDICTED running sum.  For the actual start
a = c + delta the simulated end is reused, end(a) = end(c) + delta, when (i) every simulated result stays more than
|delta| (+ 2 ulp) inside its binade and (ii) delta is a multiple of the coarsest ulp a result was rounded to (twice that
where a rounding was an exact tie) -- rounding to a lattice commutes with translations of the lattice.  Candidates
c_0 + j ulp(p), j < 8, cover the residues of delta for groups whose lattice is coarser than ulp(p).  Anything else is run
sequentially.  tests/test_gpu_chain.py checks the CUDA implementation on the device; this file pins the mathematics."""
import zlib

import numpy as np
import pytest

f32 = np.float32
J = 8


def bits(x):
    return int(np.array([x], np.float32).view(np.uint32)[0])


def from_bits(b):
    return np.array([b & 0xFFFFFFFF], np.uint32).view(np.float32)[0]


def cand_base(p):
    pb = bits(p)
    pe = (pb >> 23) & 0xFF
    if pe < 27 or pe == 255:
        return None
    c0 = from_bits(pb & ~(J - 1))
    us = from_bits((pb & 0x80000000) | ((pe - 23) << 23))
    return c0, us, pe


def sim(start, xs):
    """-> (end, margin, lattice exponent, any element non-zero)"""
    s = f32(start)
    mx, margin, etie, any_nz = 0, np.inf, 0, False
    with np.errstate(all="ignore"):
        for x in xs:
            x = f32(x)
            any_nz |= not (x == 0)
            r = f32(s + x)
            bb = f32(r - s)
            err = f32(f32(s - f32(r - bb)) + f32(x - bb))          # TwoSum: s + x == r + err exactly
            ar = bits(r) & 0x7FFFFFFF
            mx = max(mx, ar)
            ex, fr = ar >> 23, ar & 0x7FFFFF
            if ex < 24 or ex == 255:
                margin = -1.0
            else:
                u = 2.0 ** (ex - 150)
                margin = min(margin, (min(fr, 0x800000 - fr) - 2) * u)
                if abs(float(err)) == 0.5 * u and x != 0:
                    etie = max(etie, ex + 1)
            s = r
    return s, margin, max(mx >> 23, etie), any_nz


def shift_ok(delta, margin, el):
    if delta == 0:
        return True
    if not abs(delta) < margin:
        return False
    if el == 0:
        return True
    if el < 24 or el > 254:
        return False
    q = delta / 2.0 ** (el - 150)
    return q == np.floor(q)


def spec_chain(xs, G):
    """Returns (sum, groups run sequentially); python floats (fp64) hold the exact differences of floats."""
    xs = np.asarray(xs, np.float32)
    n = len(xs)
    ng = (n + G - 1) // G
    gs = [float(np.sum(xs[g * G:(g + 1) * G].astype(np.float64))) for g in range(ng)]
    pred = np.concatenate([[0.0], np.cumsum(gs)[:-1]]) if ng else []
    a, n_seq = f32(0), 0
    for g in range(ng):
        blk = xs[g * G:(g + 1) * G]
        done = False
        with np.errstate(all="ignore"):
            cb = cand_base(f32(pred[g])) if np.isfinite(pred[g]) else None
        if cb is not None:
            c0, us, pe = cb
            end0, m0, el0, any_nz = sim(c0, blk)
            if not any_nz:
                done = True                                         # identity on every start
            else:
                delta = float(a) - float(c0)
                if shift_ok(delta, max(m0, 0.0), el0):
                    a, done = f32(float(end0) + delta), True
                elif el0 > pe:
                    D = delta / float(us)
                    if D == np.floor(D) and abs(D) < 1e15 and int(D) % J != 0:
                        j = int(D) % J
                        cj = f32(float(c0) + j * float(us))
                        endj, mj, elj, _ = sim(cj, blk)
                        mj = float(from_bits(bits(f32(max(mj, 0.0))) & 0xFFFFFF00))   # the packed (rounded-down) margin
                        dr = float(a) - float(cj)
                        if shift_ok(dr, mj, elj):
                            v = float(endj) + dr
                            assert float(f32(v)) == v
                            a, done = f32(v), True
        if not done:
            n_seq += 1
            with np.errstate(all="ignore"):
                for x in blk:
                    a = f32(a + x)
    return a, n_seq


def seq_sum(xs):
    with np.errstate(all="ignore"):
        return np.cumsum(np.asarray(xs, np.float32), dtype=np.float32)[-1] if len(xs) else f32(0)


def same(a, b):
    return bits(a) == bits(b) or (np.isnan(a) and np.isnan(b))


KINDS = ["walk", "drift", "squares", "masked", "range", "ties", "spikes", "tiny"]


def make(kind, n, rng):
    if kind == "walk":
        x = 0.3 * rng.standard_normal(n)
    elif kind == "drift":
        x = rng.standard_normal(n) + 0.3
    elif kind == "squares":
        x = rng.standard_normal(n) ** 2
    elif kind == "masked":
        x = rng.standard_normal(n) + 0.05
        x[rng.random(n) < 0.66] = 0.0
    elif kind == "range":
        x = rng.standard_normal(n) * np.exp2(rng.integers(-20, 20, n))
    elif kind == "ties":
        x = (rng.integers(-4, 5, n)) * 0.5
    elif kind == "spikes":
        x = 1e-3 * rng.standard_normal(n)
        x[::97] = 1e6 * rng.standard_normal(len(x[::97]))
    else:
        x = rng.standard_normal(n) * 1e-36
    return x.astype(np.float32)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("G", [16, 64])
def test_speculative_chain_is_bit_identical_to_sequential(kind, G):
    rng = np.random.default_rng(zlib.crc32(("%s-%d" % (kind, G)).encode()))
    total_seq = total_groups = 0
    for n in (1, 17, 1000, 3000):
        xs = make(kind, n, rng)
        got, n_seq = spec_chain(xs, G)
        assert same(got, seq_sum(xs)), (kind, G, n, got, seq_sum(xs))
        total_seq += n_seq
        total_groups += (n + G - 1) // G
    if kind in ("drift", "squares", "masked"):
        assert total_seq <= 0.25 * total_groups + 4          # the records carry almost every group of a drifting sum


def test_nonfinite_elements_fall_back_to_the_sequential_chain():
    rng = np.random.default_rng(3)
    xs = make("drift", 2000, rng)
    xs[700] = np.inf
    got, _ = spec_chain(xs, 64)
    assert same(got, seq_sum(xs))
    xs[1500] = -np.inf
    got, _ = spec_chain(xs, 64)
    assert same(got, seq_sum(xs))


def test_shift_rule_on_a_binade_crossing():
    """The lattice rule in isolation: a group that rises one binade is exact for even shifts only (candidate 0), and the
    other residues are served by their own candidates."""
    xs = np.full(8, f32(0.30000001), np.float32)
    for start in (f32(7.0), f32(7.0000005), f32(7.000001), f32(7.0000014)):
        c0, us, pe = cand_base(start)
        ref = start
        for x in xs:
            ref = f32(ref + x)
        end0, m0, el0, _ = sim(c0, xs)
        assert el0 > pe                                          # the sum crosses 8.0: coarser lattice than ulp(start)
        delta = float(start) - float(c0)
        j = int(round(delta / float(us))) % J
        cj = f32(float(c0) + j * float(us))
        endj, mj, elj, _ = sim(cj, xs)
        assert shift_ok(float(start) - float(cj), max(mj, 0.0), elj)
        assert same(f32(float(endj) + (float(start) - float(cj))), ref)
