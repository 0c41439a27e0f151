"""CPU suite: the arithmetic behind gbrl_b200/csrc/chain.cuh, restated in numpy and checked against a plain sequential
float32 accumulation.

Inside one binade of the running sum s (ulp u) the chain  s <- fl(s + x)  is integer arithmetic: with m = s/u and
y = x/u,  fl(s + x) = u * (m + RN(y)), except for exact ties (frac(y) == 1/2), which go to the even neighbour of m + y
and hence depend on the parity of m.  A block of elements is summarised by a two-entry table a[p] (increment for
incoming parity p) and the min / max prefix; the summary is applied when the whole block stays strictly inside
(2^23, 2^24) for the actual m, otherwise the block is run sequentially.  The CUDA code implements exactly this
(tests/test_gpu_chain.py checks it on the device); this file pins the mathematics where no GPU is needed."""
import numpy as np
import pytest

f32 = np.float32
MARGIN = 4


def seq_sum(xs, s0=f32(0)):
    s = f32(s0)
    for x in xs:
        s = f32(s + x)
    return s


def epoch_of(s):
    """inv_u = 2^(23-e), u = 2^(e-23) for |s| in [2^e, 2^(e+1)); None for 0 / tiny / non-finite sums."""
    bits = np.array([s], dtype=np.float32).view(np.uint32)[0]
    ex = int((bits >> 23) & 0xFF)
    if ex < 27 or ex == 255:
        return None
    return f32(2.0 ** (23 - (ex - 127))), f32(2.0 ** ((ex - 127) - 23))


def summarize(xs, inv_u):
    """(a0, a1, mn, mx) of a block, or None if an element is outside the integer model (|x| >= |s|/2, inf, NaN)."""
    y = (xs * inv_u).astype(np.float32)                     # exact: power-of-two scaling
    if not np.all(np.abs(y) < f32(2.0 ** 22)):
        return None
    d = np.rint(y).astype(np.float32)
    fr = (y - d).astype(np.float32)                         # exact
    tie = np.abs(fr) == f32(0.5)
    di = d.astype(np.int64)
    k = di - ((fr < 0) & tie)                               # floor(y) of a tie
    d0 = np.where(tie, k + (k & 1), di)                     # m even: the even one of {m + k, m + k + 1}
    d1 = np.where(tie, k + ((k + 1) & 1), di)               # m odd
    a0 = a1 = mn = mx = 0
    for i in range(len(xs)):
        a0 += int(d1[i] if (a0 & 1) else d0[i])
        a1 += int(d1[i] if ((a1 + 1) & 1) else d0[i])
        mn, mx = min(mn, a0), max(mx, a0)
    return a0, a1, mn - 1, mx + 1                           # the other parity path differs by at most 1


def chain(xs, block=256):
    s = f32(0)
    fast = slow = 0
    for i in range(0, len(xs), block):
        blk = xs[i:i + block]
        done = False
        ep = epoch_of(s)
        if ep is not None:
            inv_u, u = ep
            t = summarize(blk, inv_u)
            if t is not None:
                a0, a1, mn, mx = t
                m = int(f32(s * inv_u))
                lo, hi = 2 ** 23 + MARGIN, 2 ** 24 - MARGIN
                ok = (m + mn > lo and m + mx < hi) if m > 0 else (m + mx < -lo and m + mn > -hi)
                if ok:
                    s = f32(f32(m + (a1 if (m & 1) else a0)) * u)
                    done = True
        if done:
            fast += 1
        else:
            s = seq_sum(blk, s)
            slow += 1
    return s, fast, slow


def data(kind, n, rng):
    if kind == "walk":
        x = rng.standard_normal(n)
    elif kind == "drift":
        x = rng.standard_normal(n) + 0.3
    elif kind == "range":
        x = rng.standard_normal(n) * np.exp(5.0 * rng.standard_normal(n))
    elif kind == "ties":
        x = np.round(rng.standard_normal(n) * 64) / 64 + 0.125
    elif kind == "sparse":
        x = rng.standard_normal(n) - 0.2
        x[rng.random(n) < 0.7] = 0.0
    elif kind == "negdrift":
        x = rng.standard_normal(n) * 0.1 - 1.0
    else:
        raise ValueError(kind)
    return x.astype(np.float32)


@pytest.mark.parametrize("kind", ["walk", "drift", "range", "ties", "sparse", "negdrift"])
@pytest.mark.parametrize("block", [32, 256])
def test_block_summaries_reproduce_the_sequential_chain(kind, block):
    rng = np.random.default_rng(abs(hash((kind, block))) % (2 ** 32))
    total_fast = 0
    for _ in range(3):
        xs = data(kind, int(rng.integers(2000, 12000)), rng)
        want = seq_sum(xs)
        got, fast, slow = chain(xs, block)
        assert want.tobytes() == got.tobytes(), (kind, block, want, got)
        total_fast += fast
    if kind in ("drift", "negdrift"):
        assert total_fast > 0          # the summaries are actually used, not only the fallback


def test_tables_compose_associatively():
    """(g then f)[p] = g[p] + f[(p + g[p]) & 1]: composing block tables equals summarising the concatenation."""
    rng = np.random.default_rng(3)
    inv_u = f32(2.0 ** 10)
    for _ in range(200):
        a = (np.round(rng.standard_normal(16) * 8) / 2 ** 11).astype(np.float32)      # many exact ties at this inv_u
        b = (np.round(rng.standard_normal(16) * 8) / 2 ** 11).astype(np.float32)
        ta, tb, tab = summarize(a, inv_u), summarize(b, inv_u), summarize(np.concatenate([a, b]), inv_u)
        for p in (0, 1):
            ga = ta[p]
            assert ga + tb[(p + ga) & 1] == tab[p]
