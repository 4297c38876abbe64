"""Diagnostics: window path vs sequential pipeline, first differing chain."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_helpers as H
from test_gpu_window import _data, _run, _same_chains

prec = sys.argv[1] if len(sys.argv) > 1 else "fp64"
X, y = _data(500, 2, 3)
K, C, sweeps = 3, 64, int(sys.argv[2]) if len(sys.argv) > 2 else 10
seq = _run(X, y, K, C, sweeps, seed=11, precision=prec, sequential=True)
win = _run(X, y, K, C, sweeps, seed=11, precision=prec)
print("differ", _same_chains(seq, win))
shown = 0
for c in range(C):
    keys = []
    for nm, (x, z) in zip(["tok", "pa", "pb", "nn"], zip(seq["cur"], win["cur"])):
        if not np.array_equal(x[c], z[c]): keys.append("cur." + nm)
    for nm, (x, z) in zip(["tok", "pa", "pb", "nn"], zip(seq["rep"], win["rep"])):
        if not np.array_equal(x[c], z[c]): keys.append("rep." + nm)
    for nm in ["sigma", "sa", "sb", "done", "nerr", "beta", "sse"]:
        if not np.array_equal(seq["st"][nm][c], win["st"][nm][c]): keys.append(nm)
    if not np.array_equal(seq["st"]["counters"][c], win["st"]["counters"][c]): keys.append("counters")
    if keys and shown < 6:
        shown += 1
        print("chain", c, keys)
        print("  seq counters", seq["st"]["counters"][c], "win", win["st"]["counters"][c])
        print("  seq nn", seq["cur"][3][c], "win nn", win["cur"][3][c], "sigma", seq["st"]["sigma"][c], win["st"]["sigma"][c])
        print("  beta", seq["st"]["beta"][c], win["st"]["beta"][c], "sse", seq["st"]["sse"][c], win["st"]["sse"][c])
