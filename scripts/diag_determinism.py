import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_helpers as H

def run(K, n, d, groups, C=1024, window=32):
    rng = np.random.default_rng(K * 100 + d)
    X = rng.uniform(-3, 3, (n, d))
    y = np.sin(X[:, 0]) * X[:, 1] + 0.5 * X[:, d - 1] ** 2
    eng = H.default_engine(K, C, d, val=25, plateau=True)
    eng.set_window(window)
    eng.set_data(X, y); eng.init_chains(4242); eng.set_launch_geometry(0, groups)
    eng.run(7); eng.run(23)
    out = (eng.get_trees(current=True), eng.get_stats(), eng.get_trees(current=False))
    eng.close()
    return out

for (K, n, d) in [(2, 333, 3), (3, 1000, 2), (5, 700, 8)]:
    ref = run(K, n, d, 1)
    for rep in range(8):
        g = 1 if rep < 4 else 4
        got = run(K, n, d, g)
        bad = [c for c in range(1024) if not all(np.array_equal(a[c], b[c]) for a, b in zip(ref[0], got[0]))]
        st = [k for k in ("sigma", "sse", "beta", "counters", "done", "nerr") if not np.array_equal(ref[1][k], got[1][k], equal_nan=(k not in ("counters", "done", "nerr")))]
        badr = [c for c in range(1024) if not all(np.array_equal(a[c], b[c]) for a, b in zip(ref[2], got[2]))]
        if badr:
            c = badr[0]
            print("   REPORTED trees differ", badr[:8], len(badr), "done", ref[1]["done"][c], got[1]["done"][c], "nerr", ref[1]["nerr"][c], got[1]["nerr"][c],
                  "nn ref", ref[2][3][c], "got", got[2][3][c], "cur nn", ref[0][3][c])
        print((K, n, d), "groups", g, "chains with different trees", bad[:8], len(bad), "stats differing", st)
        if bad:
            c = bad[0]
            print("   counters ref", ref[1]["counters"][c], "got", got[1]["counters"][c], "nn", ref[0][3][c], got[0][3][c])
