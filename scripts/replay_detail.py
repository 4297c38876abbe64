#!/usr/bin/env python
"""Details of the hard deviations of one window replay (same arguments as tests/test_gpu_window.py::test_window_run_replays_through_oracle)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
g.build()
import parity_helpers as H
prec = sys.argv[1] if len(sys.argv) > 1 else "fp64"
rng = np.random.default_rng(12)
X = rng.uniform(-3, 3, (300, 2))
y = 2.5 * X[:, 0] ** 4 - 1.3 * X[:, 0] ** 3 + 0.5 * X[:, 1] ** 2 - 1.7 * X[:, 1]
st = H.replay_window_run_in_oracle(X, y, K=3, n_chains=48, sweeps=25, seed=77, precision=prec, run_chunks=(4, None), window=64, detail=True)
for d in st["details"]:
    print(json.dumps(dict((k, v) for k, v in d.items() if k not in ("state", "proposed_enc"))))
print(dict((k, v) for k, v in st.items() if k != "details"))
