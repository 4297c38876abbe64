"""Small runs of the window path for compute-sanitizer (memcheck / racecheck / synccheck), covering what the shared-memory carving of
k_weval / careful_tile depends on: out-of-range columns (second pass, vector list), several row tiles and splits, ragged row counts,
K = 3 / 5 / 10 and the generic-K kernel, a record ring that wraps.  Run under gpurun:
    compute-sanitizer --tool memcheck  python scripts/sanitize_window.py
    compute-sanitizer --tool racecheck python scripts/sanitize_window.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_helpers as H  # noqa: E402

CASES = [  # K, n, d, X scale, chains, sweeps, env
    (3, 333, 2, 30.0, 48, 40, {}),                                              # out-of-range columns, ragged n, one tile
    (3, 1500, 2, 30.0, 24, 30, {"BSR_WIN_TILE": "128", "BSR_WIN_SPLITS": "3"}),  # tiles and splits, second pass on several tiles
    (5, 1201, 8, 3.0, 32, 12, {}),
    (10, 2001, 8, 3.0, 16, 6, {"BSR_WIN_TILE": "512"}),
    (7, 999, 3, 3.0, 16, 6, {}),                                                 # generic-K kernels
]
tot = np.zeros(8)
for K, n, d, scale, C, sweeps, env in CASES:
    for k, v in env.items():
        os.environ[k] = v
    rng = np.random.default_rng(K * 100 + d)
    X = rng.uniform(-scale, scale, (n, d))
    y = np.sin(X[:, 0]) * X[:, 1] + 0.5 * X[:, d - 1] ** 2
    eng = H.default_engine(K, C, d, val=0, plateau=False)
    eng.set_window(64)
    eng.set_data(X, y)
    eng.init_chains(4242)
    for chunk in (1, sweeps):          # (two calls: the ring carries over)
        eng.run(chunk)
    c = eng.get_stats()["counters"].sum(axis=0)
    print("K=%d n=%d d=%d chains=%d: proposals %d accepts %d rank-rejects %d out-of-range %d" % (K, n, d, C, c[0], c[1], c[2], c[4]), flush=True)
    tot[:len(c)] += c[:8]
    eng.close()
    for k in env:
        del os.environ[k]
print("done", tot[:5])
