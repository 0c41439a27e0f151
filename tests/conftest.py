import os
import sys

# The reference oracle (oracle/_ref) partitions some float reductions over its OpenMP threads; pin it
# before libgomp is loaded so that `ref_threads` in the tests means what it says.
os.environ.setdefault("OMP_NUM_THREADS", "4")
REF_THREADS = int(os.environ["OMP_NUM_THREADS"])

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
