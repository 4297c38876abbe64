import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_helpers as H
from mcmc_symreg_b200 import capi
TR = capi.TR
K, n, d = 2, 333, 3
rng = np.random.default_rng(K * 100 + d)
X = rng.uniform(-3, 3, (n, d))
y = np.sin(X[:, 0]) * X[:, 1] + 0.5 * X[:, d - 1] ** 2
steps = 30 * K

def run(groups):
    eng = H.default_engine(K, 1024, d, val=25, plateau=True)
    eng.set_data(X, y); eng.init_chains(4242); eng.set_launch_geometry(0, groups)
    eng.set_tape(None, steps)
    eng.run(7); eng.run(23)
    tr = eng.get_trace(steps); st = eng.get_stats()
    eng.close()
    return tr, st

inv = {v: k for k, v in TR.items()}
for rep in range(int(os.environ.get("REPS", "25"))):
    a, sa = run(1); b, sb = run(int(os.environ.get("G2", "4")))
    same = (a == b) | ((a != a) & (b != b))
    if same.all():
        continue
    idx = np.argwhere(~same)
    c, s = idx[0][0], idx[0][1]
    print("rep", rep, "chains differing", len(np.unique(idx[:, 0])), "first: chain", c, "step", s)
    for f in range(a.shape[2]):
        if not (a[c, s, f] == b[c, s, f] or (a[c, s, f] != a[c, s, f] and b[c, s, f] != b[c, s, f])):
            print("    field", inv.get(f, f), a[c, s, f], b[c, s, f])
    print("    row A", dict((inv.get(f, f), a[c, s, f]) for f in range(a.shape[2])))
    break
print("done")
