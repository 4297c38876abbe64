"""Experiment drivers: what the reference's scripts do around ``BSR``.

``codes/simulations.py:60-107`` draws a training set and two test sets (the training range and twice that range) for one
benchmark function, fits ``BSR(K, MM)``, predicts and shows the trees; ``codes/BSR.py`` is the same recipe as a usage
template; ``archive/data_generate_funcs.py`` holds the generators of the paper's benchmark functions.  Here the same
recipe is one call -- ``run_experiment`` -- and a command line (``python -m mcmc_symreg_b200.experiments --func f6``),
with the restarts running as chains on the GPU and the summaries the scripts produce by hand (RMSE on the training
range and under extrapolation, model size, expressions, best restart, across-restart diagnostics) returned as a dict.

The genetic-programming comparison of the script (gplearn, ``codes/simulations.py:150-175``) is not rebuilt.
"""
import argparse
import json
import time

import numpy as np

# The benchmark suite: the paper's f1 ... f6 (bsr_paper.pdf p.5, Eqs. 3-8); "sim" is the function the script itself uses
# (codes/simulations.py:71, = f6).  All take an (n, 2) matrix.
TARGETS = {
    "f1": (lambda X: 2.5 * X[:, 0] ** 4 - 1.3 * X[:, 0] ** 3 + 0.5 * X[:, 1] ** 2 - 1.7 * X[:, 1], "2.5 x0^4 - 1.3 x0^3 + 0.5 x1^2 - 1.7 x1"),
    "f2": (lambda X: 8 * X[:, 0] ** 2 + 8 * X[:, 1] ** 3 - 15, "8 x0^2 + 8 x1^3 - 15"),
    "f3": (lambda X: 0.2 * X[:, 0] ** 3 + 0.5 * X[:, 1] ** 3 - 1.2 * X[:, 1] - 0.5 * X[:, 0], "0.2 x0^3 + 0.5 x1^3 - 1.2 x1 - 0.5 x0"),
    "f4": (lambda X: 1.5 * np.exp(X[:, 0]) + 5 * np.cos(X[:, 1]), "1.5 exp(x0) + 5 cos(x1)"),
    "f5": (lambda X: 6.0 * np.sin(X[:, 0]) * np.cos(X[:, 1]), "6 sin(x0) cos(x1)"),
    "f6": (lambda X: 1.35 * X[:, 0] * X[:, 1] + 5.5 * np.sin((X[:, 0] - 1) * (X[:, 1] - 1)), "1.35 x0 x1 + 5.5 sin((x0 - 1)(x1 - 1))"),
}
TARGETS["sim"] = TARGETS["f6"]


def make_dataset(func, n_train=100, n_test=30, lo=-3.0, hi=3.0, extrapolate=2.0, seed=None):
    """Training set on U(lo, hi)^2, a test set on the same range and one on the range stretched by ``extrapolate``
    (codes/simulations.py:64-85: n = 100 on (-3, 3), 30 + 30 test points on (-3, 3) and (-6, 6)).  ``seed=None`` draws from
    numpy's global state like the script does."""
    f = TARGETS[func][0]
    rng = np.random.default_rng(seed) if seed is not None else np.random
    X = rng.uniform(lo, hi, (n_train, 2))
    Xt = rng.uniform(lo, hi, (n_test, 2))
    Xe = rng.uniform(lo * extrapolate, hi * extrapolate, (n_test, 2))
    return dict(func=func, formula=TARGETS[func][1], X=X, y=f(X), X_test=Xt, y_test=f(Xt), X_extra=Xe, y_extra=f(Xe))


def rmse(pred, y):
    with np.errstate(all="ignore"):
        return float(np.sqrt(np.mean((np.asarray(pred).ravel() - np.asarray(y).ravel()) ** 2)))


def summarize(est, data):
    """What the script reads off a fitted estimator (codes/simulations.py:98-107, 127-128 and the commented diagnostics)."""
    best = est.best_chain()
    out = dict(func=data["func"], formula=data["formula"], restarts=len(est.betas_), K=est.treeNum)
    for tag, Xk, yk in (("train", "X", "y"), ("test", "X_test", "y_test"), ("extrapolation", "X_extra", "y_extra")):
        out["rmse_%s_last" % tag] = rmse(est.predict(data[Xk]), data[yk])             # the reference's own answer: the last restart
        out["rmse_%s_best" % tag] = rmse(est.predict_best(data[Xk]), data[yk])
        out["rmse_%s_mean" % tag] = rmse(est.predict_mean(data[Xk]), data[yk])         # posterior-predictive mean over restarts
    out["model_last"] = est.model()
    out["complexity_last"] = est.complexity()
    out["model_best"] = est.model(last_ind=len(est.betas_) - best)
    out["beta_best"] = [float(v) for v in np.asarray(est.betas_[best]).ravel()]
    nn = est._packed_.nn
    out["complexity_best"] = int(nn[best].sum())
    out["complexity_mean"] = float(nn.sum(axis=1).mean())
    c = est.counters_
    out["proposals"] = int(c[:, 0].sum())
    out["accept_rate"] = float(c[:, 1].sum() / max(1, c[:, 0].sum()))
    out["accepts_per_restart"] = float(np.mean([len(e) for e in est.train_err_]))
    out["diagnostics"] = est.chain_diagnostics()
    return out


def run_experiment(func="sim", K=3, MM=50, n_train=100, n_test=30, val=100, seed=0, data_seed=None, **bsr_kwargs):
    """One benchmark function end to end: data, ``BSR(K, MM).fit``, predictions on both test sets, summaries."""
    from .bsr_class import BSR
    data = make_dataset(func, n_train, n_test, seed=seed if data_seed is None else data_seed)
    est = BSR(K, MM, val=val, seed=seed, **bsr_kwargs)
    t0 = time.perf_counter()
    est.fit(data["X"], data["y"])
    out = summarize(est, data)
    out["fit_seconds"] = time.perf_counter() - t0
    return out, est, data


def run_suite(funcs=("f1", "f2", "f3", "f4", "f5", "f6"), repeats=1, **kw):
    """The whole benchmark suite (the paper's Table 1 recipe: every function, ``repeats`` independent data sets)."""
    rows = []
    seed0 = kw.pop("seed", 0)
    for f in funcs:
        for r in range(repeats):
            out, _, _ = run_experiment(f, seed=seed0 + r, **kw)
            out["repeat"] = r
            rows.append(out)
    return rows


def format_row(o):
    return ("%-4s restarts %5d | RMSE train / test / extrapolation: last %.3g / %.3g / %.3g  best %.3g / %.3g / %.3g  mean %.3g / %.3g / %.3g | "
            "nodes last %d best %d mean %.1f | accept %.2f %% | %.2f s" %
            (o["func"], o["restarts"], o["rmse_train_last"], o["rmse_test_last"], o["rmse_extrapolation_last"], o["rmse_train_best"],
             o["rmse_test_best"], o["rmse_extrapolation_best"], o["rmse_train_mean"], o["rmse_test_mean"], o["rmse_extrapolation_mean"],
             o["complexity_last"], o["complexity_best"], o["complexity_mean"], 100 * o["accept_rate"], o["fit_seconds"]))


def main(argv=None):
    ap = argparse.ArgumentParser(description="BSR benchmark-function experiments on the GPU (codes/simulations.py recipe)")
    ap.add_argument("--func", default="sim", help="one of %s, or 'all'" % ", ".join(sorted(TARGETS)))
    ap.add_argument("--K", type=int, default=3)
    ap.add_argument("--MM", type=int, default=50, help="restarts (chains on the GPU)")
    ap.add_argument("--n-train", type=int, default=100)
    ap.add_argument("--n-test", type=int, default=30)
    ap.add_argument("--val", type=int, default=100)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--repeats", type=int, default=1)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--json", action="store_true", help="one JSON object per experiment instead of the table")
    a = ap.parse_args(argv)
    funcs = ("f1", "f2", "f3", "f4", "f5", "f6") if a.func == "all" else (a.func,)
    rows = run_suite(funcs, a.repeats, K=a.K, MM=a.MM, n_train=a.n_train, n_test=a.n_test, val=a.val, seed=a.seed, precision=a.precision)
    for o in rows:
        if a.json:
            print(json.dumps(o))
        else:
            print(format_row(o))
            print("     target     y = %s" % o["formula"])
            print("     best model y = %.4g + %s" % (o["beta_best"][0], " + ".join("%.4g * [%s]" % (b, e) for b, e in zip(o["beta_best"][1:], o["model_best"]))))
    return rows


if __name__ == "__main__":
    main()
