"""CPU model of the memoised mean chains (gbrl_b200/csrc/preprocess.cu dense_memo_*_kernel), in numpy float32.

The GPU tests check the kernels bit for bit against a sequential sum (tests/test_gpu_chain.py).  This file checks the two
claims the design rests on, independently of any GPU:
  1. exactness needs no argument: a table value is used only when the candidate start EQUALS the running sum, so the walk
     reproduces the sequential float32 sum whatever the prediction was (also with a deliberately bad prediction);
  2. the premise that makes it fast: the running sum of a zero-mean chain stays within a few hundred ulps of its
     exact-arithmetic prediction (accumulated rounding error only), so +-256 candidates per group catch nearly every group.
"""
import numpy as np

K = 256          # MW_K
GROUP = 256


def _grid(pred):
    """mw_grid: candidate spacing = ulp of the prediction's binade, half of it when the candidates reach below the binade."""
    pred = np.float32(pred)
    if pred == 0 or not np.isfinite(pred):
        return None
    ex = (np.frombuffer(np.float32(pred).tobytes(), np.uint32)[0] >> 23) & 0xFF
    if ex < 27 or ex == 255:
        return None
    u = np.float32(2.0) ** np.float32(int(ex) - 127 - 23)
    fine = abs(np.float32(pred) / u) < np.float32(8388608.0 + K)
    return np.float32(0.5) * u if fine else u


def _sequential(x, start=np.float32(0)):
    acc = np.float32(start)
    for v in x:
        acc = np.float32(acc + v)
    return acc


def _memo_walk(x, pred_error=0.0):
    """Returns (sum, groups looked up, groups run sequentially)."""
    n = len(x)
    ng = (n + GROUP - 1) // GROUP
    exact_prefix = np.concatenate([[0.0], np.cumsum(x.astype(np.float64))])
    acc = np.float32(0)
    hits = misses = 0
    for g in range(ng):
        xs = x[g * GROUP:(g + 1) * GROUP]
        pred = np.float32(exact_prefix[g * GROUP] * (1.0 + pred_error))
        step = _grid(pred)
        done = False
        if step is not None:
            kf = np.rint(np.float32(np.float32(acc - pred) / step))
            if abs(kf) < K and np.float32(pred + np.float32(kf) * step) == acc:
                # the table row of the group: sequential chains from all candidate starts (vectorised over the candidates)
                ks = np.arange(-K, K, dtype=np.float32)
                cand = (pred + ks * step).astype(np.float32)
                for v in xs:
                    cand = (cand + v).astype(np.float32)
                acc = cand[int(kf) + K]
                done = True
                hits += 1
        if not done:
            acc = _sequential(xs, acc)
            misses += 1
    return acc, hits, misses


def test_memo_walk_equals_sequential_sum_and_mostly_hits():
    rng = np.random.default_rng(3)
    for scale in (0.4, 3e-3, 250.0):
        x = (rng.standard_normal(20000) * scale).astype(np.float32)
        want = _sequential(x)
        got, hits, misses = _memo_walk(x)
        assert got.tobytes() == want.tobytes()
        assert hits >= 0.85 * (hits + misses), (scale, hits, misses)     # the first group and a few grid misses run sequentially


def test_memo_walk_is_exact_with_a_bad_prediction():
    rng = np.random.default_rng(4)
    x = (rng.standard_normal(6000) * 0.7 + 0.01).astype(np.float32)
    want = _sequential(x)
    for err in (1e-6, 1e-4, -3e-3):
        got, hits, misses = _memo_walk(x, pred_error=err)
        assert got.tobytes() == want.tobytes()


def test_deviation_from_exact_prefix_is_a_few_hundred_ulps():
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(62500) * 0.4).astype(np.float32)      # one reference thread's share of C2's gradients
    acc = np.float32(0)
    exact = 0.0
    worst = 0.0
    for i, v in enumerate(x):
        acc = np.float32(acc + v)
        exact += float(v)
        if i % 256 == 255 and exact != 0.0:
            worst = max(worst, abs(float(acc) - exact) / float(np.spacing(np.float32(exact))))
    assert worst < K, worst
