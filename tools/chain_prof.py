#!/usr/bin/env python
"""One drifting 1M-element chain through the parallel evaluator (for ncu)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_chain import _data, _gpu_partials
rng = np.random.default_rng(0)
mat = _data("drift", 1_000_000, 1, rng)
info = np.zeros(4)
for _ in range(2):
    _gpu_partials(mat, 1, 1, 0, None, impl=0, info=info)
print(info)
