import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_helpers as H

def run(K, n, d, groups, sequential):
    rng = np.random.default_rng(K * 100 + d)
    X = rng.uniform(-3, 3, (n, d))
    y = np.sin(X[:, 0]) * X[:, 1] + 0.5 * X[:, d - 1] ** 2
    eng = H.default_engine(K, 96 if sequential else 1024, d, val=25, plateau=True)
    eng.set_pipeline(sequential)
    eng.set_data(X, y); eng.init_chains(4242); eng.set_launch_geometry(0, groups)
    eng.run(7); eng.run(23)
    out = (eng.get_trees(current=True), eng.get_trees(current=False), eng.get_stats(), eng.get_err_trace())
    eng.close()
    return out

names = ["tok", "pa", "pb", "nn"]
G2 = int(os.environ.get('G2', '4'))
for rep in range(int(os.environ.get('REPS', '6'))):
    for (K, n, d) in [(2, 333, 3)]:
        for seq in (False,):
            a = run(K, n, d, 1, seq); b = run(K, n, d, G2, seq)
            msgs = []
            for i in range(4):
                if not np.array_equal(a[0][i], b[0][i]):
                    cs = np.unique(np.argwhere(a[0][i] != b[0][i])[:, 0]); msgs.append("cur." + names[i] + str(cs[:5]) + str(len(cs)))
                if not np.array_equal(a[1][i], b[1][i]):
                    cs = np.unique(np.argwhere(a[1][i] != b[1][i])[:, 0]); msgs.append("rep." + names[i] + str(cs[:5]) + str(len(cs)))
            for k in ("sigma", "sa", "sb", "beta", "sse", "done", "nerr", "counters"):
                if not np.array_equal(a[2][k], b[2][k], equal_nan=a[2][k].dtype.kind == "f"):
                    cs = np.unique(np.argwhere(~((a[2][k] == b[2][k]) | ((a[2][k] != a[2][k]) & (b[2][k] != b[2][k]))))[:, 0]); msgs.append(k + str(cs[:5]) + str(len(cs)))
            if msgs:
                print("rep", rep, (K, n, d), "seq" if seq else "win", msgs)
                c = int(cs[0])
                print("    chain", c, "counters", a[2]["counters"][c], b[2]["counters"][c], "nn", a[0][3][c], b[0][3][c], "sse", a[2]["sse"][c], b[2]["sse"][c])
print("done")
