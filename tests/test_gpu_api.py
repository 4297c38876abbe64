"""GPU tests of the drop-in surface: the BSR estimator (codes/bsr_class.py:26-278) over the C-ABI, whole-fit
replays of the reference's own runs (stop rules included), and properties at BASELINE.json sizes."""
import pickle

import numpy as np
import pytest

from oracle import bsr_oracle as O
import parity_helpers as H

pytestmark = pytest.mark.gpu

import os
_G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIT_FILES = [f for f in ["fits_f1_k3.json.gz", "fits_f6_k2.json.gz", "fits_plateau.json.gz"] if os.path.exists(os.path.join(_G, f))]


def node_to_oracle(nd):
    """Node-shaped tree (what BSR.roots_ holds) -> oracle Tree."""
    from mcmc_symreg_b200.trees import encode_tree
    tok, pa, pb, n = encode_tree(nd)
    return H.dec_tree(tok, pa, pb, n)


@pytest.mark.parametrize("pipeline", ["window", "sequential"])
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("fname", FIT_FILES)
def test_reference_fit_replayed_on_gpu(golden, fname, precision, pipeline):
    """A whole BSR.fit of the unmodified reference (MM restarts, `val` consecutive-rejection stop, plateau stop and
    the Q16 snapshot) replayed on the GPU: every restart becomes one chain fed the reference's own draws."""
    from mcmc_symreg_b200 import capi
    from mcmc_symreg_b200.trees import Express, decode_tree, getNum
    g = golden(fname)
    X, y = np.array(g["X"]), np.array(g["y"])
    K, MM, d = g["K"], g["MM"], g["d"]
    cfg = O.Config(n_feature=d, beta=g["beta"])
    # the pinned oracle segments the reference tape into per-restart init draws + per-proposal draws
    dr = O.TapeDraws(g["tape"])
    inits, tapes, results = [], [], []
    for m in range(MM):
        p0 = dr.pos
        sigma = dr.invgamma(1.0)
        trees, sa, sb = [], [], []
        for _ in range(K):
            a_, b_ = dr.invgamma(1.0), dr.invgamma(1.0)
            trees.append(O.grow(0, cfg, a_, b_, dr))
            sa.append(a_); sb.append(b_)
        inits.append(dict(sigma=sigma, trees=trees, sigma_a=sa, sigma_b=sb))
        rec = _SegmentingDraws(dr)
        r = O.run_chain(X, y, K, cfg, rec, val=g["val"], init=inits[-1], on_step=rec.cut, keep_traces=True)
        tapes.append(rec.segments)
        results.append(r)
    assert dr.pos == len(g["tape"])
    soft, unresolved = _replay_and_compare(X, y, K, d, g["beta"], g["val"], inits, tapes, results, precision,
                                           roots=g["roots"], betas=g["betas"], errs=g["train_err"], pipeline=pipeline)
    # the two short fits replay completely; the 10k-proposal plateau fit may meet a proposal the type does not resolve
    assert soft <= (1 if "plateau" in fname else 0)
    assert [O.express(t) for t in results[-1].trees] == g["model"]


def _replay_and_compare(X, y, K, d, beta, val, inits, tapes, results, precision, roots=None, betas=None, errs=None, ops=None, weights=None,
                        pipeline="window"):
    """Feed per-restart tapes to the GPU (one chain per restart; ``pipeline``: the production window kernels of bsr_run or
    the proposal-by-proposal pipeline), compare with the oracle's chain results.

    Decisions: the first proposal at which a chain takes another decision than the reference is classified by
    parity_helpers.judge_step; only a soft deviation (the evaluation type does not resolve the proposal, or the uniform sits
    on the threshold) is tolerated, and such a restart is counted and not followed further.  Numbers: every RMSE-at-accept
    entry, the final beta and the predictions are held to the tolerance wherever the yardstick arithmetic of `precision`
    itself reproduces the float64 value within a quarter of it (parity_helpers.state_fits / resolves); the entries it does
    not resolve are counted.  Returns (restarts with a soft decision deviation, restarts with an unresolved number)."""
    from mcmc_symreg_b200 import capi
    TR = capi.TR
    MM = len(inits)
    steps = max(r.n_proposals for r in results)
    steps += (-steps) % K
    if ops is None:
        eng = H.default_engine(K, MM, d, precision=precision, val=val, plateau=True, beta=beta, err_cap=1024)
        cfg = O.Config(n_feature=d, beta=beta)
    else:
        eng = capi.Engine(K, MM, ops, weights, beta=beta, val=val, plateau_rule=True, precision=precision, err_cap=1024)
        cfg = O.Config(n_feature=d, beta=beta, ops=ops, weights=weights)
    eng.set_pipeline(pipeline == "sequential")
    eng.set_data(X, y)
    tok, pa, pb, nn = H.pack_state([i["trees"] for i in inits], K)
    eng.set_state(tok, pa, pb, nn, [i["sigma"] for i in inits], [i["sigma_a"] for i in inits], [i["sigma_b"] for i in inits])
    eng.set_tape([t + [[]] * (steps - len(t)) for t in tapes], steps)
    eng.run(steps // K)
    tr = eng.get_trace(steps)
    st = eng.get_stats()
    tok, pa, pb, nn = eng.get_trees(current=False)
    err = eng.get_err_trace()
    tol = 1e-6 if precision == "fp64" else 2e-3
    tol_err = 1e-7 if precision == "fp64" else 1e-4
    tol_logr = 1e-6 if precision == "fp64" else 1e-3
    soft_restarts = unresolved_restarts = 0
    n_err = n_err_cmp = 0
    for m, r in enumerate(results):
        div = None
        state = list(inits[m]["trees"])
        sigma, sa, sb = inits[m]["sigma"], list(inits[m]["sigma_a"]), list(inits[m]["sigma_b"])
        acc_states = []
        for s_i, ot in enumerate(r.traces):
            t = tr[m, s_i]
            k = s_i % K
            if bool(t[TR["accepted"]]) != ot.accepted or bool(t[TR["rank_reject"]]) != ot.rank_deficient:
                gpu = dict(rank_reject=bool(t[TR["rank_reject"]]), accepted=bool(t[TR["accepted"]]), logR=float(t[TR["logR"]]))
                v = H.judge_step(state, k, sigma, sa[k], sb[k], y, X, cfg, tapes[m][s_i], gpu, precision, tol_logr)
                assert not v.hard, ("restart %d step %d" % (m, s_i), v.hard, v.cls, O.express(ot.proposed), [O.express(x) for x in state],
                                    ot.logR, t[TR["logR"]])
                div = s_i
                break
            if ot.accepted:
                state[k] = ot.proposed
                sigma, sa[k], sb[k] = ot.new_sigma, ot.new_sa2, ot.new_sb2
                acc_states.append(list(state))
            # (on a reject the reference keeps sigma, sigma_a, sigma_b: funcs.py:1298-1306)
        if div is not None:
            soft_restarts += 1
            continue
        assert st["done"][m]
        assert int(st["counters"][m, 0]) == r.n_proposals and int(st["counters"][m, capi.CNT["accepts"]]) == r.n_accepts
        exp_roots = roots[m] if roots is not None else [H.enc_golden(t) for t in r.trees]
        for k in range(K):
            t = H.dec_tree(tok[m, k], pa[m, k], pb[m, k], nn[m, k])
            assert H.trees_equal(t, H.tree_from_golden(exp_roots[k])), ("reported roots", m, k)
        exp_beta = np.asarray(betas[m] if betas is not None else r.beta).ravel()
        exp_err = errs[m] if errs is not None else r.err_list
        assert len(exp_err) == len(acc_states) == int(st["nerr"][m])
        unresolved = False
        for j, stt in enumerate(acc_states):
            n_err += 1
            f = H.state_fits(stt, X, y, precision)
            if not H.resolves(f["rmse"], tol_err, margin=H.MARGIN[precision]):
                unresolved = True
                continue
            n_err_cmp += 1
            assert abs(err[m, j] - exp_err[j]) <= tol_err * abs(exp_err[j]), ("rmse at accept", m, j, err[m, j], exp_err[j])
        final_state = acc_states[-1] if acc_states else list(inits[m]["trees"])
        f = H.state_fits(final_state, X, y, precision)
        if H.resolves(f["beta"], tol, floor=1.0, margin=H.MARGIN[precision]):
            np.testing.assert_allclose(st["beta"][m], exp_beta, rtol=tol, atol=tol * max(1.0, np.max(np.abs(exp_beta))))
            roots_t = [H.tree_from_golden(e) for e in exp_roots]
            ref = O.predict(roots_t, exp_beta.reshape(-1, 1), X).ravel()
            if all(H.column_comparable(t, X, "fp64", 1e-10) for t in roots_t):   # bsr_predict evaluates in float64
                pred = eng.predict(m, X, reported=True)
                np.testing.assert_allclose(pred, ref, rtol=tol, atol=tol * max(1.0, np.max(np.abs(ref))))
        else:
            unresolved = True
        unresolved_restarts += unresolved
    eng.close()
    print("replay %s %s: %d restarts, %d with a soft decision deviation, %d with a number the type does not resolve; rmse entries "
          "compared %d of %d" % (precision, pipeline, MM, soft_restarts, unresolved_restarts, n_err_cmp, n_err))
    assert n_err == 0 or n_err_cmp >= 0.5 * n_err
    return soft_restarts, unresolved_restarts


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_readme_usage_fit_replayed_on_gpu(golden, precision):
    """BASELINE.json configs[0]: the README usage ``BSR(3, 50)`` (50 restarts, val = 100) on the paper's f1 with n = 100,
    recorded from the unmodified reference (tests/golden/fits_c1_readme.json.gz, 82.6 k draws).  Every restart becomes
    one chain of the production window kernels fed the reference's own draws, in the default precision (fp32) and in fp64.

    Rule for what is compared (parity_helpers): a decision, an RMSE-at-accept entry or a beta is held to its tolerance
    wherever the yardstick arithmetic of the precision (numpy float32 / 80-bit) reproduces the reference's float64 value
    within a quarter of that tolerance; where it does not, the deviation belongs to the type (restart 18 of this fit holds
    ``sin(((-2.0689*(x[1])+-1.2423)^3)^3)``, arguments up to 6.6e7: numpy float32 columns give RMSE 53.095 against the
    reference's 53.144) and the restart is COUNTED: at most 5 of the 50 restarts may contain a decision the type does not
    resolve, at most 10 a number it does not resolve."""
    if not os.path.exists(os.path.join(_G, "fits_c1_readme.json.gz")):
        pytest.skip("fixture not present")
    g = golden("fits_c1_readme.json.gz")
    X, y = np.array(g["X"]), np.array(g["y"])
    K, MM, d = g["K"], g["MM"], g["d"]
    cfg = O.Config(n_feature=d, beta=g["beta"])
    dr = O.TapeDraws(g["tape"])
    inits, tapes, results = [], [], []
    for m in range(MM):
        sigma = dr.invgamma(1.0)
        trees, sa, sb = [], [], []
        for _ in range(K):
            a_, b_ = dr.invgamma(1.0), dr.invgamma(1.0)
            trees.append(O.grow(0, cfg, a_, b_, dr))
            sa.append(a_); sb.append(b_)
        inits.append(dict(sigma=sigma, trees=trees, sigma_a=sa, sigma_b=sb))
        rec = _SegmentingDraws(dr)
        results.append(O.run_chain(X, y, K, cfg, rec, val=g["val"], init=inits[-1], on_step=rec.cut, keep_traces=True))
        tapes.append(rec.segments)
    assert dr.pos == len(g["tape"])
    soft, unresolved = _replay_and_compare(X, y, K, d, g["beta"], g["val"], inits, tapes, results, precision,
                                           roots=g["roots"], betas=g["betas"], errs=g["train_err"])
    print("c1 readme fit (%s): %d of %d restarts with a soft decision deviation, %d with an unresolved number" % (precision, soft, MM, unresolved))
    assert soft <= 5 and unresolved <= 10


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_plateau_stop_and_q16_snapshot(precision):
    """The RMSE plateau stop (bsr_class.py:248-252) and its side effect Q16 (roots_ keeps the pre-accept trees while
    betas_ is post-accept).  The oracle (pinned to the reference on fits_plateau) drives chains on tiny noisy data
    with an operator set without sin/cos/exp, so no proposal is numerically chaotic and the replay must be exact."""
    ops = [O.OP_LT, O.OP_NEG, O.OP_SQUARE, O.OP_CUBIC, O.OP_ADD, O.OP_MUL]
    w = [0.25, 0.15, 0.15, 0.15, 0.15, 0.15]
    K, d, val = 2, 2, 1500
    inits, tapes, results = [], [], []
    seed = 100
    while len(results) < 3 and seed < 140:
        rng = np.random.default_rng(seed)
        if not results:
            X = rng.uniform(-2, 2, (6, d)); y = rng.normal(0, 3, 6)
        cfg = O.Config(n_feature=d, ops=ops, weights=w)
        dr = O.GeneratorDraws(seed, record=True)
        sigma = dr.invgamma(1.0)
        trees, sa, sb = [], [], []
        for _ in range(K):
            a_, b_ = dr.invgamma(1.0), dr.invgamma(1.0)
            trees.append(O.grow(0, cfg, a_, b_, dr)); sa.append(a_); sb.append(b_)
        init = dict(sigma=sigma, trees=trees, sigma_a=sa, sigma_b=sb)
        segs = []
        mark = [len(dr.tape)]

        def cut():
            segs.append(list(dr.tape[mark[0]:]))
            mark[0] = len(dr.tape)

        r = O.run_chain(X, y, K, cfg, dr, val=val, init=init, on_step=cut, keep_traces=True, max_sweeps=6000)
        seed += 1
        fired = len(r.err_list) > 100 and r.n_proposals < 12000 and [t.key() for t in r.trees] != [t.key() for t in r.final_state]
        if fired:
            inits.append(init); tapes.append(segs); results.append(r)
    assert results and len(results[0].err_list) > 100, "no plateau stop found"
    soft, unresolved = _replay_and_compare(X, y, K, d, -1.0, val, inits, tapes, results, precision, ops=ops, weights=w)
    assert soft == 0


class _SegmentingDraws:
    """Pass-through over a TapeDraws that remembers where each proposal's draws start."""

    def __init__(self, inner):
        self.inner, self.segments, self.start = inner, [], inner.pos

    def cut(self):
        self.segments.append(list(self.inner.tape[self.start:self.inner.pos]))
        self.start = self.inner.pos

    def __getattr__(self, name):
        return getattr(self.inner, name)


def test_bsr_estimator_api():
    """Same call sequence as the reference README (README.md:22-38) and the same fitted attributes."""
    import pandas as pd
    from mcmc_symreg_b200 import BSR, Express, getNum
    rng = np.random.default_rng(0)
    X = pd.DataFrame(rng.uniform(-3, 3, (100, 2)))
    y = 2.5 * X[0] ** 4 - 1.3 * X[0] ** 3 + 0.5 * X[1] ** 2 - 1.7 * X[1]
    K, MM = 3, 50
    est = BSR(K, MM, seed=17)
    assert est.get_params()["treeNum"] == 3 and est.get_params()["val"] == 100
    assert est.fit(X, y) is None
    assert len(est.roots_) == MM and len(est.betas_) == MM and len(est.train_err_) == MM
    assert all(len(r) == K for r in est.roots_) and all(b.shape == (K + 1, 1) for b in est.betas_)
    model = est.model()
    assert isinstance(model, list) and len(model) == K and all(isinstance(s, str) and "x[" in s for s in model)
    assert model == [Express(r) for r in est.roots_[-1]]
    assert est.model(last_ind=MM) == [Express(r) for r in est.roots_[0]]
    assert est.complexity() == sum(getNum(r) for r in est.roots_[-1])
    Xt = rng.uniform(-3, 3, (30, 2))
    pred = est.predict(Xt)
    assert pred.shape == (30, 1)
    # predictions agree with a float64 host evaluation of the reported trees and betas
    for li in (1, 7):
        ref = O.predict([node_to_oracle(r) for r in est.roots_[-li]], est.betas_[-li], Xt)
        got = est.predict(pd.DataFrame(Xt), last_ind=li)
        np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(ref).max()))
    with pytest.raises(UnboundLocalError):
        est.predict(Xt, method="best")
    # train_err_ is the RMSE at each accept: the last entry is the RMSE of the reported model when no Q16 snapshot
    for m in range(MM):
        if est.train_err_[m] and est.counters_[m, 1] == len(est.train_err_[m]):
            rmse = float(np.sqrt(np.mean((est.predict(X.values, last_ind=MM - m).ravel() - y.values) ** 2)))
            assert abs(rmse - est.train_err_[m][-1]) <= 1e-3 * max(1.0, rmse)
    # reproducible from the seed, independent chains differ
    est2 = BSR(K, MM, seed=17)
    est2.fit(X.values, y.values)
    assert est2.model() == model and [Express(r) for r in est2.roots_[3]] == [Express(r) for r in est.roots_[3]]
    assert len({tuple(Express(t) for t in r) for r in est.roots_}) > MM // 2
    # sklearn plumbing the reference inherits (bsr_class.py:26)
    assert np.isfinite(est.score(X, y))
    est3 = pickle.loads(pickle.dumps(est))
    np.testing.assert_array_equal(est3.predict(Xt), pred)
    # the fit improves on the prior draw: median train RMSE across restarts beats predicting the mean
    final = [e[-1] for e in est.train_err_ if e]
    assert len(final) > MM // 4 and np.median(final) < np.std(y.values)


def test_packed_results_posterior_mean_best_chain_and_untruncated_trace():
    """Results leave the device as node-count-long prefixes (bsr_pack_trees); roots_ decodes lazily; the RMSE-at-accept trace is not
    truncated under fit() however small the initial capacity; predict_mean / predict_best / chain_diagnostics
    agree with a float64 host evaluation of the reported models."""
    from mcmc_symreg_b200 import BSR
    rng = np.random.default_rng(3)
    X = rng.uniform(-3, 3, (150, 2))
    y = 1.35 * X[:, 0] * X[:, 1] + 5.5 * np.sin((X[:, 0] - 1) * (X[:, 1] - 1))
    # packed == dense
    eng = H.default_engine(3, 200, 2)
    eng.set_data(X, y); eng.init_chains(4); eng.run(40)
    dense = eng.get_trees(current=False)
    packed = eng.get_trees_packed(current=False)
    assert eng.last_tree_bytes < 0.2 * sum(a.nbytes for a in dense)
    for a, b in zip(dense, packed.dense()):
        assert np.array_equal(a, b)
    t = packed.tree(17, 2)
    assert np.array_equal(t[0], dense[0][17, 2]) and np.array_equal(t[1], dense[1][17, 2]) and t[3] == dense[3][17, 2]
    eng.close()
    # estimator: tiny trace capacity, long fit
    MM = 300
    est = BSR(3, MM, val=150, seed=5, err_cap=4)
    est.fit(X, y)
    assert est.done_.all() and not est.train_err_truncated_
    acc = est.counters_[:, 1]
    assert [len(e) for e in est.train_err_] == [int(a) for a in acc] and acc.max() > 4
    # lazy roots behave like the reference's list of lists
    assert len(est.roots_) == MM and len(est.roots_[-1]) == 3 and len(est.roots_._cache) == 1
    assert [len(r) for r in est.roots_[2:5]] == [3, 3, 3]
    Xt = rng.uniform(-3, 3, (40, 2))
    preds = np.stack([O.predict([node_to_oracle(r) for r in est.roots_[m]], est.betas_[m], Xt).ravel() for m in range(MM)])
    ok = np.all(np.isfinite(preds), axis=1)
    mean, std = est.predict_mean(Xt, return_std=True)
    fin = np.isfinite(preds)
    ref_mean = np.array([preds[fin[:, j], j].mean() for j in range(preds.shape[1])])
    ref_std = np.array([preds[fin[:, j], j].std(ddof=1) for j in range(preds.shape[1])])
    np.testing.assert_allclose(mean.ravel(), ref_mean, rtol=1e-8, atol=1e-8 * np.abs(ref_mean).max())
    np.testing.assert_allclose(std.ravel(), ref_std, rtol=1e-6, atol=1e-8 * np.abs(ref_std).max())
    assert est.n_used_ >= ok.sum() - 1
    b = est.best_chain()
    assert est.final_rmse_[b] == np.nanmin(est.final_rmse_)
    np.testing.assert_allclose(est.predict_best(Xt).ravel(), preds[b], rtol=1e-9, atol=1e-9 * np.abs(preds[b]).max())
    np.testing.assert_allclose(est.predict(Xt, last_ind=3).ravel(), preds[MM - 3], rtol=1e-9, atol=1e-9 * np.abs(preds[MM - 3]).max())
    dg = est.chain_diagnostics()
    assert dg["best"] == b and dg["final_rmse_best"] <= dg["final_rmse_median"] and (np.isnan(dg["rhat"]) or dg["rhat"] >= 0.9)
    # a subset of restarts (here: the better half by training RMSE) averages the same way
    half = np.argsort(est.final_rmse_)[:MM // 2]
    sub = est.predict_mean(Xt, chains=half)
    fin_h = np.isfinite(preds[half])
    np.testing.assert_allclose(sub.ravel(), [preds[half][fin_h[:, j], j].mean() for j in range(preds.shape[1])], rtol=1e-8, atol=1e-8 * np.abs(ref_mean).max())
    est2 = pickle.loads(pickle.dumps(est))
    assert est2.model() == est.model() and len(est2.roots_) == MM


def test_error_behaviour():
    from mcmc_symreg_b200 import BSR, capi
    with pytest.raises(capi.BsrError):
        capi.Engine(0, 4, [1, 2], [0.5, 0.5])
    with pytest.raises(capi.BsrError):
        capi.Engine(3, 4, [1, 99], [0.5, 0.5])
    eng = capi.Engine(2, 4, [1, 2, 9], [0.3, 0.3, 0.4])
    with pytest.raises(capi.BsrError):
        eng.init_chains(1)                 # no data yet
    with pytest.raises(capi.BsrError):
        eng.run(1)
    eng.close()
    with pytest.raises(ValueError):
        BSR(2, 2).fit(np.zeros((5, 2)), np.zeros(4))


@pytest.mark.parametrize("K,C,n,d,sweeps", [(3, 4096, 1000, 2, 30), (5, 2048, 5000, 8, 6), (10, 96, 10000, 8, 3)])
def test_properties_at_baseline_sizes(K, C, n, d, sweeps):
    """BASELINE.json configs C2 / C4 / C3 shapes: size-independent invariants + spot checks against the oracle."""
    from mcmc_symreg_b200 import capi
    rng = np.random.default_rng(K * 7 + d)
    X = rng.uniform(-3, 3, (n, d))
    y = np.exp(0.4 * X[:, 0]) + 2 * np.cos(X[:, 1]) + 0.3 * X[:, d - 1] * X[:, 0] + rng.normal(0, 0.1, n)
    eng = H.default_engine(K, C, d)
    eng.set_data(X, y)
    eng.init_chains(5)
    eng.run(sweeps)
    st = eng.get_stats()
    tok, pa, pb, nn = eng.get_trees(current=True)
    cnt = st["counters"]
    assert (cnt[:, capi.CNT["proposals"]] == sweeps * K).all() and (cnt[:, capi.CNT["sweeps"]] == sweeps).all()
    assert (cnt[:, capi.CNT["accepts"]] + cnt[:, capi.CNT["rank_rejects"]] + cnt[:, capi.CNT["capacity_rejects"]] <= sweeps * K).all()
    assert cnt[:, capi.CNT["accepts"]].sum() > 0
    assert (nn >= 2).all() and (nn <= H.MAX_NODES).all()
    assert np.isfinite(st["sigma"]).all() and (st["sigma"] > 0).all() and (st["sa"] > 0).all() and (st["sb"] > 0).all()
    ok = np.isfinite(st["sse"])
    assert ok.mean() > 0.9 and (st["sse"][ok] >= 0).all() and (st["sse"][ok] <= 1.0001 * float(y @ y)).all()
    # every tree decodes (well-formed pre-order) and spot-checked chains match the oracle's SSE and intercept fit
    n_checked = 0
    for c in rng.choice(C, 24, replace=False):
        trees = [H.dec_tree(tok[c, k], pa[c, k], pb[c, k], nn[c, k]) for k in range(K)]
        for t in trees:
            assert O.subtree_sizes(t.op)[0] == len(t)
        if not ok[c]:
            continue
        f = H.state_fits(trees, X, y, "fp32")
        if not H.resolves(f["sse"], 2e-3) or not H.resolves(f["beta"], 5e-3, floor=1.0):
            continue                        # float32 itself does not resolve this state's fit
        n_checked += 1
        cols = [O.eval_tree(t, X) for t in trees]
        sse = O.sse_no_intercept(y, np.stack(cols, axis=1))
        assert abs(sse - st["sse"][c]) <= 2e-3 * max(sse, 1e-9 * float(y @ y)), (c, sse, st["sse"][c])
        if cnt[c, capi.CNT["accepts"]] == 0:
            beta, _ = O.intercept_fit(cols, y)
            np.testing.assert_allclose(st["beta"][c], beta.ravel(), rtol=5e-3, atol=5e-3 * np.abs(beta).max())
    assert n_checked >= 8, n_checked
    eng.close()
