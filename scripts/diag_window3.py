import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_helpers as H
rng = np.random.default_rng(9)
X = rng.uniform(-2, 2, (2999, 3))
y = np.sin(X[:, 0]) + X[:, 1] ** 2 + 0.05 * rng.normal(size=2999)
for prec in ("fp64", "fp32"):
    st = H.replay_window_run_in_oracle(X, y, K=7, n_chains=24, sweeps=6, seed=21, precision=prec)
    print("window", prec, st)
    st = H.replay_gpu_run_in_oracle(X, y, K=7, n_chains=24, sweeps=6, seed=21, precision=prec)
    print("seq   ", prec, st)
