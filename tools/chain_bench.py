#!/usr/bin/env python
"""Timing + path statistics of the parallel float-chain evaluator (csrc/chain.cuh) on the GPU box."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_chain import _data, _gpu_partials

rng = np.random.default_rng(0)
for kind in ("walk", "drift", "sparse", "squares", "grad"):
    for (n, D, T) in ((1_000_000, 1, 16), (1_000_000, 1, 1), (250_000, 4, 16)):
        if kind == "squares":
            mat = (_data("walk", n, D, rng) * 0.4) ** 2
        elif kind == "grad":
            mat = (_data("walk", n, D, rng) * 0.4 + 0.001).astype(np.float32)
        else:
            mat = _data(kind, n, D, rng)
        info = np.zeros(4)
        for impl in (0, 2, 3, 1):
            _gpu_partials(mat, D, T, 0, None, impl=impl, info=info)      # warm
            _gpu_partials(mat, D, T, 0, None, impl=impl, info=info)
            print("%-8s n=%d D=%d T=%d impl=%d: %.3f ms  fast %d adv %d seq_lanes %d" % (kind, n, D, T, impl, info[0], info[1], info[2], info[3]), flush=True)
