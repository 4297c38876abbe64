import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_helpers as H
from test_gpu_window import _data, _run, _same_chains
X, y = _data(500, 2, 3)
K, C, sweeps = 3, 64, 40
def sse_diff(a, b):
    return int(np.sum(a["st"]["sse"] != b["st"]["sse"])), _same_chains(a, b, rel=0.0)
seq_c = _run(X, y, K, C, sweeps, seed=11, sequential=True)
os.environ["BSR_NO_COL_CACHE"] = "1"
seq_p = _run(X, y, K, C, sweeps, seed=11, sequential=True)
del os.environ["BSR_NO_COL_CACHE"]
w32 = _run(X, y, K, C, sweeps, seed=11)
w1 = _run(X, y, K, C, sweeps, seed=11, window=1)
w5 = _run(X, y, K, C, sweeps, seed=11, window=5, chunks=[7, 33])
print("seq cached vs seq plain", sse_diff(seq_c, seq_p))
print("seq plain vs w32", sse_diff(seq_p, w32))
print("seq cached vs w32", sse_diff(seq_c, w32))
print("w1 vs w32", sse_diff(w1, w32))
print("w5 vs w32", sse_diff(w5, w32))
st = H.replay_window_run_in_oracle(X, y, K=3, n_chains=48, sweeps=25, seed=77, precision="fp32")
print(st)
st = H.replay_gpu_run_in_oracle(X, y, K=3, n_chains=48, sweeps=25, seed=77, precision="fp32")
print(st)
