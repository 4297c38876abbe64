import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_helpers as H
K, n, d = 2, 333, 3
rng = np.random.default_rng(K * 100 + d)
X = rng.uniform(-3, 3, (n, d))
y = np.sin(X[:, 0]) * X[:, 1] + 0.5 * X[:, d - 1] ** 2
eng = H.default_engine(K, int(os.environ.get("CH", "256")), d, val=25, plateau=True)
eng.set_data(X, y); eng.init_chains(4242)
eng.run(int(os.environ.get("SW", "8")))
print(eng.get_stats()["counters"].sum(axis=0))
eng.close()
