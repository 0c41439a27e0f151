#!/usr/bin/env python
"""One chain workload through one evaluator variant (for ncu):  chain_prof.py <kind> <T> <impl>"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_chain import _data, _gpu_partials
kind = sys.argv[1] if len(sys.argv) > 1 else "drift"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1
impl = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rng = np.random.default_rng(0)
mat = (_data("walk", 1_000_000, 1, rng) * 0.4 + 0.001).astype(np.float32) if kind == "grad" else _data(kind, 1_000_000, 1, rng)
info = np.zeros(4)
for _ in range(2):
    _gpu_partials(mat, 1, T, 0, None, impl=impl, info=info)
print(info)
