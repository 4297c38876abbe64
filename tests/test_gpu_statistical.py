"""Statistical parity of the production path (Philox RNG, fp32 evaluation) with the reference algorithm on the benchmark suite of
BASELINE.json configs[1]: the paper's f1 ... f6 (bsr_paper.pdf p.5, Eqs. 3-8) and the target of codes/simulations.py:71.  The
oracle, pinned to the unmodified reference, runs its own independent restarts with numpy's RNG (one per host core at a time);
the GPU runs many more restarts with Philox through the estimator.  Same model, same stop rules (val consecutive rejections,
plateau) => acceptance rate, proposals per restart, fitted RMSE (train and held-out) and model size must agree in distribution
(two-sample tests, not equality)."""
import multiprocessing as mp
import os

import numpy as np
import pytest
from scipy import stats

from oracle import bsr_oracle as O

pytestmark = pytest.mark.gpu

TARGETS = {
    "f1": lambda X: 2.5 * X[:, 0] ** 4 - 1.3 * X[:, 0] ** 3 + 0.5 * X[:, 1] ** 2 - 1.7 * X[:, 1],
    "f2": lambda X: 8 * X[:, 0] ** 2 + 8 * X[:, 1] ** 3 - 15,
    "f3": lambda X: 0.2 * X[:, 0] ** 3 + 0.5 * X[:, 1] ** 3 - 1.2 * X[:, 1] - 0.5 * X[:, 0],
    "f4": lambda X: 1.5 * np.exp(X[:, 0]) + 5 * np.cos(X[:, 1]),
    "f5": lambda X: 6.0 * np.sin(X[:, 0]) * np.cos(X[:, 1]),
    "f6": lambda X: 1.35 * X[:, 0] * X[:, 1] + 5.5 * np.sin((X[:, 0] - 1) * (X[:, 1] - 1)),          # = codes/simulations.py:71
}
K, VAL, N_TRAIN, N_ORACLE, N_GPU = 3, 60, 100, 128, 3000


def _data(name, n, seed):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-3, 3, (n, 2))                                                  # codes/simulations.py:66-67
    return X, TARGETS[name](X)


def _oracle_restart(args):
    name, m = args
    X, y = _data(name, N_TRAIN, 2001)
    Xt, yt = _data(name, 200, 2002)
    r = O.run_chain(X, y, K, O.Config(n_feature=2), O.GeneratorDraws(10_000 + m), val=VAL)
    with np.errstate(all="ignore"):
        tr = float(np.sqrt(np.mean((O.predict(r.trees, r.beta, X).ravel() - y) ** 2)))
        te = float(np.sqrt(np.mean((O.predict(r.trees, r.beta, Xt).ravel() - yt) ** 2)))
    return r.n_accepts, r.n_proposals, tr, te, sum(len(t) for t in r.trees)


@pytest.mark.parametrize("name", sorted(TARGETS))
def test_posterior_quality_matches_reference_algorithm(name):
    from mcmc_symreg_b200 import BSR, capi
    X, y = _data(name, N_TRAIN, 2001)
    Xt, yt = _data(name, 200, 2002)
    # ---- oracle: independent restarts, spread over the host cores ----
    with mp.get_context("fork").Pool(min(os.cpu_count() or 1, 16)) as pool:
        res = pool.map(_oracle_restart, [(name, m) for m in range(N_ORACLE)], chunksize=2)
    o_acc, o_props, o_rmse, o_test, o_size = [np.array(v) for v in zip(*res)]
    # ---- GPU: N_GPU restarts through the estimator ----
    est = BSR(K, N_GPU, val=VAL, seed=99)
    est.fit(X, y)
    cnt = est.counters_
    g_acc, g_props = cnt[:, 1], cnt[:, 0]
    sub = np.arange(0, N_GPU, 4)
    tok, pa, pb, nn = est._enc_
    beta = np.stack([est.betas_[m].ravel() for m in sub])
    with np.errstate(all="ignore"):
        g_rmse = np.sqrt(np.mean((capi.predict_many(0, tok[sub], pa[sub], pb[sub], nn[sub], beta, X) - y) ** 2, axis=1))
        g_test = np.sqrt(np.mean((capi.predict_many(0, tok[sub], pa[sub], pb[sub], nn[sub], beta, Xt) - yt) ** 2, axis=1))
    g_size = nn.sum(axis=1)
    o_rate, g_rate = o_acc.sum() / o_props.sum(), g_acc.sum() / g_props.sum()
    print("%s: acceptance rate oracle %.4f gpu %.4f | proposals/restart %.1f %.1f | train RMSE median %.3f %.3f | test RMSE median %.3f %.3f | "
          "nodes/model %.2f %.2f" % (name, o_rate, g_rate, o_props.mean(), g_props.mean(), np.nanmedian(o_rmse), np.nanmedian(g_rmse),
                                     np.nanmedian(o_test), np.nanmedian(g_test), o_size.mean(), g_size.mean()))
    fin = lambda a: np.asarray(a, dtype=float)[np.isfinite(a)]
    # acceptance: binomial two-proportion z-test on pooled proposals (over-dispersed across restarts => generous bound)
    se = np.sqrt(o_rate * (1 - o_rate) / o_props.sum() + g_rate * (1 - g_rate) / g_props.sum())
    assert abs(o_rate - g_rate) < 6 * se + 0.15 * o_rate
    for what, a, b in (("accepts/restart", o_acc, g_acc), ("proposals/restart", o_props, g_props), ("train rmse", fin(o_rmse), fin(g_rmse)),
                       ("test rmse", fin(o_test), fin(g_test)), ("model size", o_size, g_size)):
        p = stats.ks_2samp(a, b).pvalue
        print("   KS %-18s p = %.3g" % (what, p))
        assert p > 1e-3, (name, what, p)
