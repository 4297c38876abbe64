"""Pins oracle/bsr_oracle.py against fixtures recorded from the UNMODIFIED reference
(tests/golden/gen_golden.py).  Integer bookkeeping must be bit-exact; floats agree to 1e-9 rel."""
import math

import numpy as np
import pytest

from oracle import bsr_oracle as O

# steps_w_d3_k3: non-uniform Op_weights (quirk Q7: op_ind is stale after reassignOperator, which only shows with unequal weights)
STEP_FILES = ["steps_f1_d2_k3.json.gz", "steps_mix_d8_k5.json.gz", "steps_deep_d3_k2.json.gz", "steps_w_d3_k3.json.gz"]
# fits_c1_readme: BASELINE.json configs[0], the README usage BSR(3, 50) on the paper's f1 with n = 100 (50 restarts, val = 100)
FIT_FILES = ["fits_f1_k3.json.gz", "fits_f6_k2.json.gz", "fits_plateau.json.gz", "fits_c1_readme.json.gz"]


def tree_from(enc):
    return O.Tree(enc["op"], enc["oi"], enc["ft"], enc["a"], enc["b"])


def assert_tree_equal(t, enc, what=""):
    assert t.op == enc["op"], what
    assert t.ft == enc["ft"], what
    assert t.oi == enc["oi"], what
    for i, o in enumerate(t.op):
        if o == O.OP_LT:
            assert t.a[i] == enc["a"][i] and t.b[i] == enc["b"][i], (what, i, t.a[i], enc["a"][i])


def close(a, b, rel=1e-9, abs_=1e-12):
    if a is None or b is None:
        return a is None and b is None
    if math.isnan(a) or math.isnan(b):
        return math.isnan(a) and math.isnan(b)
    if math.isinf(a) or math.isinf(b):
        return a == b
    return abs(a - b) <= abs_ + rel * max(abs(a), abs(b))


@pytest.mark.parametrize("fname", STEP_FILES)
def test_init_replay(golden, fname):
    g = golden(fname)
    cfg = O.Config(n_feature=g["d"], beta=g["beta"], weights=g["weights"])
    for ch in g["chains"]:
        dr = O.TapeDraws(ch["init_tape"])
        sigma = dr.invgamma(1.0)
        assert sigma == ch["init"]["sigma"]
        for k in range(g["K"]):
            sa, sb = dr.invgamma(1.0), dr.invgamma(1.0)
            t = O.grow(0, cfg, sa, sb, dr)
            assert_tree_equal(t, ch["init"]["trees"][k], "init tree")
        assert dr.pos == len(ch["init_tape"])


@pytest.mark.parametrize("fname", STEP_FILES)
def test_newprop_replay(golden, fname):
    g = golden(fname)
    X, y = np.array(g["X"]), np.array(g["y"])
    cfg = O.Config(n_feature=g["d"], beta=g["beta"], weights=g["weights"])
    K = g["K"]
    n_steps = n_acc = n_rank = 0
    moves = set()
    for ci, ch in enumerate(g["chains"]):
        trees = [tree_from(e) for e in ch["init"]["trees"]]
        sigma, sa, sb = ch["init"]["sigma"], list(ch["init"]["sa"]), list(ch["init"]["sb"])
        for si, st in enumerate(ch["steps"]):
            what = "%s chain %d step %d" % (fname, ci, si)
            c = st["count"]
            dr = O.TapeDraws(st["tape"])
            acc, sigma, newt, sa[c], sb[c], tr = O.new_prop(trees, c, sigma, y, X, cfg, sa[c], sb[c], dr)
            assert dr.pos == len(st["tape"]), what + " draw count"
            assert tr.change == st["change"], what
            assert_tree_equal(tr.proposed, st["proposed"], what + " proposed")
            assert close(tr.Q, st["Q"]) and close(tr.Qinv, st["Qinv"]), (what, tr.Q, st["Q"], tr.Qinv, st["Qinv"])
            if st["change"] != 0:
                assert close(tr.hratio, st["aux"][0], rel=1e-8) and close(tr.detjacob, st["aux"][1]), what
                assert tr.new_sa2 == st["aux"][2] and tr.new_sb2 == st["aux"][3]
            else:
                assert tr.new_sa2 == st["aux"][0] and tr.new_sb2 == st["aux"][1]
            col = O.eval_tree(tr.proposed, X)
            np.testing.assert_allclose(col[:len(st["pcol"])], st["pcol"], rtol=1e-12, atol=0, err_msg=what)
            assert tr.rank_deficient == st["rank_reject"], what
            if not st["rank_reject"]:
                assert close(tr.logR, st["logR"], rel=1e-7, abs_=1e-7), (what, tr.logR, st["logR"])
            assert acc == st["accepted"], what
            assert sigma == st["sigma"] and sa[c] == st["sa"] and sb[c] == st["sb"], what
            if acc:
                assert_tree_equal(newt, st["tree"], what + " accepted tree")
                assert O.get_num(newt) == st["num"] and O.get_height(newt) == st["height"]
                assert O.num_lt(newt.op) == st["numlt"] and O.express(newt) == st["expr"], what
                trees = list(trees)
                trees[c] = newt
            n_steps += 1; n_acc += acc; n_rank += tr.rank_deficient
            moves.add((tr.move, tr.change))
        for k in range(K):
            assert_tree_equal(trees[k], ch["final"][k], "final")
            np.testing.assert_allclose(O.eval_tree(trees[k], X), ch["final_cols"][k], rtol=1e-12)
    assert n_steps > 500 and n_acc > 5
    assert {m for m, _ in moves} == set(range(7)), moves     # every move type was exercised
    assert {c for _, c in moves} == {0, 1, 2}


@pytest.mark.parametrize("fname", FIT_FILES)
def test_fit_replay(golden, fname):
    """Whole BSR.fit runs (MM restarts, val stop rule) replayed from the reference's tape."""
    g = golden(fname)
    X, y = np.array(g["X"]), np.array(g["y"])
    cfg = O.Config(n_feature=g["d"], beta=g["beta"])
    dr = O.TapeDraws(g["tape"])
    results = [O.run_chain(X, y, g["K"], cfg, dr, val=g["val"]) for _ in range(g["MM"])]
    assert dr.pos == len(g["tape"])
    for m, r in enumerate(results):
        for k in range(g["K"]):
            assert_tree_equal(r.trees[k], g["roots"][m][k], "restart %d tree %d" % (m, k))
        np.testing.assert_allclose(r.beta.ravel(), g["betas"][m], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(r.err_list, g["train_err"][m], rtol=1e-8)
    last = results[-1]
    assert [O.express(t) for t in last.trees] == g["model"]
    assert [O.express(t) for t in results[0].trees] == g["model_first"]
    assert sum(O.get_num(t) for t in last.trees) == g["complexity"]
    np.testing.assert_allclose(O.predict(last.trees, last.beta, np.array(g["Xtest"])).ravel(), g["predict"], rtol=1e-6, atol=1e-9)


def test_generator_draws_self_consistent():
    """The oracle's own stream, recorded and replayed, reproduces the same chain."""
    rng = np.random.default_rng(3)
    X = rng.uniform(-3, 3, (50, 2)); y = X[:, 0] ** 2 + np.sin(X[:, 1])
    cfg = O.Config(n_feature=2)
    d1 = O.GeneratorDraws(5, record=True)
    r1 = O.run_chain(X, y, 3, cfg, d1, val=40)
    r2 = O.run_chain(X, y, 3, cfg, O.TapeDraws(d1.tape), val=40)
    assert [t.key() for t in r1.trees] == [t.key() for t in r2.trees]
    assert r1.err_list == r2.err_list and r1.n_proposals == r2.n_proposals
