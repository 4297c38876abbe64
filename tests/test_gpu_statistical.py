"""Statistical parity of the production path (Philox RNG, fp32 evaluation) with the reference algorithm: the oracle,
pinned to the unmodified reference, runs its own independent chains with numpy's RNG; the GPU runs many more chains
with Philox.  Same model, same stop rules (val consecutive rejections, plateau) => acceptance rate, fitted RMSE and
model size must agree in distribution (two-sample tests, not equality)."""
import numpy as np
import pytest
from scipy import stats

from oracle import bsr_oracle as O

pytestmark = pytest.mark.gpu


def _sim_data(n, seed):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-3, 3, (n, 2))                                                  # codes/simulations.py:66-67
    y = 1.35 * X[:, 0] * X[:, 1] + 5.5 * np.sin((X[:, 0] - 1) * (X[:, 1] - 1))       # codes/simulations.py:71
    return X, y


def test_posterior_quality_matches_reference_algorithm():
    from mcmc_symreg_b200 import BSR
    from mcmc_symreg_b200.trees import getNum
    K, val, n = 3, 60, 100
    X, y = _sim_data(n, 2001)
    Xt, yt = _sim_data(200, 2002)
    # ---- oracle: 160 independent restarts ----
    cfg = O.Config(n_feature=2)
    o_acc, o_rmse, o_test, o_size, o_props = [], [], [], [], []
    for m in range(160):
        r = O.run_chain(X, y, K, cfg, O.GeneratorDraws(10_000 + m), val=val)
        o_acc.append(r.n_accepts); o_props.append(r.n_proposals)
        pred = O.predict(r.trees, r.beta, X).ravel()
        o_rmse.append(float(np.sqrt(np.mean((pred - y) ** 2))))
        o_test.append(float(np.sqrt(np.mean((O.predict(r.trees, r.beta, Xt).ravel() - yt) ** 2))))
        o_size.append(sum(len(t) for t in r.trees))
    # ---- GPU: 3000 restarts through the estimator ----
    est = BSR(K, 3000, val=val, seed=99)
    est.fit(X, y)
    cnt = est.counters_
    g_acc, g_props = cnt[:, 1], cnt[:, 0]
    from mcmc_symreg_b200 import capi
    tok, pa, pb, nn = est._enc_
    g_rmse, g_test = [], []
    for m in range(0, 3000, 5):
        b = est.betas_[m].ravel()
        pr = capi.predict_trees(0, tok[m], pa[m], pb[m], nn[m], b, X)
        g_rmse.append(float(np.sqrt(np.mean((pr - y) ** 2))))
        pt = capi.predict_trees(0, tok[m], pa[m], pb[m], nn[m], b, Xt)
        g_test.append(float(np.sqrt(np.mean((pt - yt) ** 2))))
    g_size = nn.sum(axis=1)
    o_rate, g_rate = np.sum(o_acc) / np.sum(o_props), g_acc.sum() / g_props.sum()
    print("acceptance rate  oracle %.4f  gpu %.4f" % (o_rate, g_rate))
    print("proposals/chain  oracle %.1f  gpu %.1f" % (np.mean(o_props), g_props.mean()))
    print("train RMSE median oracle %.3f gpu %.3f | test RMSE median oracle %.3f gpu %.3f" %
          (np.median(o_rmse), np.median(g_rmse), np.median(o_test), np.median(g_test)))
    print("nodes/model mean oracle %.2f gpu %.2f" % (np.mean(o_size), g_size.mean()))
    fin = lambda a: np.asarray(a)[np.isfinite(a)]
    # acceptance: binomial two-proportion z-test on pooled proposals (over-dispersed across chains => generous bound)
    se = np.sqrt(o_rate * (1 - o_rate) / np.sum(o_props) + g_rate * (1 - g_rate) / g_props.sum())
    assert abs(o_rate - g_rate) < 6 * se + 0.15 * o_rate
    for name, a, b in (("accepts/chain", o_acc, g_acc), ("proposals/chain", o_props, g_props), ("train rmse", fin(o_rmse), fin(g_rmse)),
                       ("test rmse", fin(o_test), fin(g_test)), ("model size", o_size, g_size)):
        p = stats.ks_2samp(a, b).pvalue
        print("KS %-16s p = %.3g" % (name, p))
        assert p > 1e-3, name
