import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_window import _run
rng = np.random.default_rng(9)
X = rng.uniform(-2, 2, (2999, 3))
y = np.sin(X[:, 0]) + X[:, 1] ** 2 + 0.05 * rng.normal(size=2999)
K, C, sweeps = 7, 64, 6
a = _run(X, y, K, C, sweeps, seed=21)
s = _run(X, y, K, C, sweeps, seed=21, sequential=True)
for c in range(C):
    keys = []
    for nm, (x, z) in zip(["tok", "pa", "pb", "nn"], zip(s["cur"], a["cur"])):
        if not np.array_equal(x[c], z[c]): keys.append("cur." + nm)
    for nm, (x, z) in zip(["tok", "pa", "pb", "nn"], zip(s["rep"], a["rep"])):
        if not np.array_equal(x[c], z[c]): keys.append("rep." + nm)
    for nm in ["sigma", "sa", "sb", "done", "nerr"]:
        if not np.array_equal(s["st"][nm][c], a["st"][nm][c]): keys.append(nm)
    if not np.array_equal(s["st"]["counters"][c][[0, 1, 2, 3, 7]], a["st"]["counters"][c][[0, 1, 2, 3, 7]]): keys.append("counters")
    rb = np.max(np.abs(s["st"]["beta"][c] - a["st"]["beta"][c]) / (np.abs(s["st"]["beta"][c]) + 1e-300))
    rs = abs(s["st"]["sse"][c] - a["st"]["sse"][c]) / abs(s["st"]["sse"][c])
    if keys or rb > 1e-6 or rs > 1e-6:
        print(c, keys, "beta rel", rb, "sse rel", rs, "counters", s["st"]["counters"][c][:4], a["st"]["counters"][c][:4])
        print("   beta seq", s["st"]["beta"][c], "\n   beta win", a["st"]["beta"][c])
