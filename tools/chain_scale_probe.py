import os, sys
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from test_gpu_chain import _data, _gpu_partials
rng = np.random.default_rng(0)
for n in (1_000_000, 8_000_000):
    for kind in ("grad", "drift"):
        mat = (rng.standard_normal((n, 1)) * 0.4 + (0.001 if kind == "grad" else 0.3)).astype(np.float32)
        info = np.zeros(4)
        for impl in (0,):
            _gpu_partials(mat, 1, 16, 0, None, impl=impl, info=info)
            _gpu_partials(mat, 1, 16, 0, None, impl=impl, info=info)
            print("%s n=%d T=16 impl=%d: %.3f ms fast %d slow %d seq_lanes %d" % (kind, n, impl, info[0], info[1], info[2], info[3]), flush=True)
