"""GPU parity tests (run on the B200 box with ``-m gpu``): the CUDA path, driven through the C-ABI, against the
oracle and against the fixtures recorded from the unmodified reference."""
import numpy as np
import pytest

from oracle import bsr_oracle as O
import parity_helpers as H

pytestmark = pytest.mark.gpu

STEP_FILES = ["steps_f1_d2_k3.json.gz", "steps_mix_d8_k5.json.gz", "steps_deep_d3_k2.json.gz"]
# tolerances (north_star): tree outputs rel 1e-4, log-likelihood / acceptance ratio rel 1e-3 in fp32;
# fp64 evaluation mode must agree with the float64 reference to rounding.
TOL_COL = {"fp32": 1e-4, "fp64": 1e-10}
TOL_LOGR = {"fp32": 1e-3, "fp64": 1e-6}


def _capi():
    from mcmc_symreg_b200 import capi
    return capi


def random_trees(n_trees, d, seed, beta=-0.6):
    cfg = O.Config(n_feature=d, beta=beta)
    dr = O.GeneratorDraws(seed)
    out = []
    while len(out) < n_trees:
        t = O.grow(0, cfg, 0.5, 0.5, dr)
        if len(t) <= H.MAX_NODES:
            out.append(t)
    return out


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_eval_trees_vs_oracle(precision):
    """allcal (funcs.py:175-220): every operator incl. the exp / inv guards, deep trees, ragged n."""
    rng = np.random.default_rng(5)
    n, d = 1003, 4
    X = rng.uniform(-3, 3, (n, d))
    X[0, 0] = 0.0                      # inv guard
    X[1, 1] = 250.0                    # exp guard
    y = rng.normal(size=n)
    trees = random_trees(300, d, seed=9)
    hand = [O.Tree([O.OP_INV, 0], [0, 0], [0, 0], [0, 0], [0, 0]), O.Tree([O.OP_EXP, 0], [5, 0], [0, 1], [0, 0], [0, 0]),
            O.Tree([O.OP_EXP, O.OP_EXP, 0], [5, 5, 0], [0, 0, 1], [0] * 3, [0] * 3),
            O.Tree([O.OP_LT, O.OP_CUBIC, 0], [1, 7, 0], [0, 0, 2], [1.5, 0, 0], [-0.25, 0, 0]),
            O.Tree([O.OP_MUL, O.OP_ADD, 0, 0, O.OP_SIN, 0], [9, 8, 0, 0, 3, 0], [0, 0, 0, 1, 0, 2], [0] * 6, [0] * 6)]
    trees = hand + trees
    eng = H.default_engine(3, 1, d, precision=precision)
    eng.set_data(X, y)
    tok, pa, pb, nn = H.pack_state([trees], len(trees))
    got = eng.eval_trees(tok[0], pa[0], pb[0], nn[0], precision=precision)
    eng.close()
    worst, n_cmp = 0.0, 0
    for i, t in enumerate(trees):
        ref = O.eval_tree(t, X)
        if not np.all(np.isfinite(ref)) or np.max(np.abs(ref)) > 1e30:
            continue
        # trees whose intermediate values are huge are ill-conditioned in fp32 (sin/cos of 1e6): fp64 only
        scale = np.max(np.abs(ref)) + 1e-300
        err = np.max(np.abs(got[i] - ref)) / scale
        if not H.column_comparable(t, X, precision, TOL_COL[precision]):   # the type's own rounding is amplified beyond the
            continue                                                         # tolerance (1/sin(1/cos^2), cos(exp(x^6)) ...)
        worst = max(worst, err)
        n_cmp += 1
        assert err <= TOL_COL[precision], (i, O.express(t), err)
    assert n_cmp > 150
    print("eval parity", precision, "trees", n_cmp, "worst normalised error", worst)


@pytest.mark.parametrize("pipeline", ["window", "sequential"])
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
@pytest.mark.parametrize("fname", STEP_FILES + ["steps_w_d3_k3.json.gz"])
def test_golden_newprop_replay(golden, fname, precision, pipeline):
    """Reference newProp sequences (tapes recorded from the unmodified reference, codes/funcs.py:1184-1306 call by call)
    replayed on the GPU -- through the production kernels of bsr_run (k_wclassify / k_wpropose / k_weval / k_wresolve fed
    the tape, ``pipeline == "window"``) and through the proposal-by-proposal pipeline of the phase API.
    Bit-exact: move bookkeeping, proposed trees, change flag, sigma draws, draw counts, rank verdicts.  Tolerance: Q, Qinv,
    hratio, logR.  Every proposal is classified by parity_helpers.judge_step; no hard deviation is tolerated, and the
    share of proposals whose logR was held to the tolerance is asserted."""
    capi = _capi()
    TR = capi.TR
    import os
    if not os.path.exists(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", fname)):
        pytest.skip("fixture not present")
    g = golden(fname)
    K, d = g["K"], g["d"]
    X, y = np.array(g["X"]), np.array(g["y"])
    cfg = O.Config(n_feature=d, beta=g["beta"], weights=g["weights"])
    chains = [ch for ch in g["chains"] if len(ch["steps"]) % K == 0 and len(ch["steps"]) > 0]
    steps = min(len(ch["steps"]) for ch in chains)
    steps -= steps % K
    C = len(chains)
    eng = H.default_engine(K, C, d, precision=precision, beta=g["beta"], weights=g["weights"])
    eng.set_pipeline(pipeline == "sequential")
    eng.set_data(X, y)
    tok, pa, pb, nn = H.pack_state([[H.tree_from_golden(e) for e in ch["init"]["trees"]] for ch in chains], K)
    eng.set_state(tok, pa, pb, nn, [ch["init"]["sigma"] for ch in chains], [ch["init"]["sa"] for ch in chains],
                  [ch["init"]["sb"] for ch in chains])
    eng.set_tape([[ch["steps"][s]["tape"] for s in range(steps)] for ch in chains], steps)
    if pipeline == "window":
        eng.trace_trees()
        eng.run(steps // K)
        ptok, ppa, ppb, pnn = eng.get_trace_trees(steps)
    else:
        ptok = np.zeros((C, steps, H.MAX_NODES), dtype=np.uint32); ppa = np.zeros((C, steps, H.MAX_NODES))
        ppb = np.zeros((C, steps, H.MAX_NODES)); pnn = np.zeros((C, steps), dtype=np.int32)
        for s in range(steps // K):
            eng.sweep_propose()
            a_, b_, c_, n_ = eng.get_proposals()
            ptok[:, s * K:(s + 1) * K], ppa[:, s * K:(s + 1) * K], ppb[:, s * K:(s + 1) * K], pnn[:, s * K:(s + 1) * K] = a_, b_, c_, n_
            eng.sweep_eval()
            eng.sweep_resolve()
    trace = eng.get_trace(steps)
    tokf, paf, pbf, nnf = eng.get_trees(current=True)
    stf = eng.get_stats()
    eng.close()

    n_cmp = n_soft = 0
    cls = dict(compared=0, rank_both=0, type_limited=0, nonfinite=0)
    worst_logr = 0.0
    for c, ch in enumerate(chains):
        state = [H.tree_from_golden(e) for e in ch["init"]["trees"]]
        sigma, sa, sb = ch["init"]["sigma"], list(ch["init"]["sa"]), list(ch["init"]["sb"])
        alive = True
        for s in range(steps):
            st, t, k = ch["steps"][s], trace[c, s], s % K
            what = "%s chain %d step %d" % (fname, c, s)
            if not alive:
                break           # after a (soft) decision flip the chain states differ: the rest of this chain's tape is foreign
            assert int(t[TR["flags"]]) == 0, what + " flags"
            # ---- bit-exact bookkeeping ----
            gp = H.dec_tree(ptok[c, s], ppa[c, s], ppb[c, s], pnn[c, s])
            assert H.trees_equal(gp, H.tree_from_golden(st["proposed"])), what + " proposed tree"
            assert int(t[TR["change"]]) == st["change"], what
            aux = st["aux"]
            sa2, sb2 = (aux[2], aux[3]) if st["change"] else (aux[0], aux[1])
            assert t[TR["new_sa2"]] == sa2 and t[TR["new_sb2"]] == sb2, what
            # ---- tolerance-bounded scalars ----
            assert H.close(t[TR["Q"]], st["Q"], 1e-9) and H.close(t[TR["Qinv"]], st["Qinv"], 1e-9), what
            if st["change"]:
                assert H.close(t[TR["hratio"]], aux[0], 1e-7, 1e-300) and H.close(t[TR["detjacob"]], aux[1], 1e-12), what
            gpu = dict(rank_reject=bool(t[TR["rank_reject"]]), accepted=bool(t[TR["accepted"]]), logR=float(t[TR["logR"]]))
            v = H.judge_step(state, k, sigma, sa[k], sb[k], y, X, cfg, st["tape"], gpu, precision, TOL_LOGR[precision])
            # the oracle is pinned to the reference on these very fixtures (tests/test_oracle_golden.py)
            assert v.tr.rank_deficient == st["rank_reject"] and v.acc == st["accepted"], what
            assert not v.hard, (what, v.hard, v.cls, O.express(v.tr.proposed), [O.express(x) for x in state], gpu, v.tr.logR)
            if v.cls != "type_limited" or not v.soft:
                assert int(t[TR["ndraws"]]) == len(st["tape"]), (what, t[TR["ndraws"]], len(st["tape"]))
            cls[v.cls] += 1
            worst_logr = max(worst_logr, v.err)
            n_cmp += 1
            n_soft += bool(v.soft)
            if gpu["accepted"] != st["accepted"] or (gpu["rank_reject"] and not st["rank_reject"]):
                alive = False
                continue
            if st["accepted"]:
                state[k] = H.tree_from_golden(st["tree"])
            sigma, sa[k], sb[k] = st["sigma"], st["sa"], st["sb"]
        if alive:
            for k in range(K):
                gt = H.dec_tree(tokf[c, k], paf[c, k], pbf[c, k], nnf[c, k])
                assert H.trees_equal(gt, H.tree_from_golden(ch["final"][k])), "final state chain %d tree %d" % (c, k)
            assert stf["sigma"][c] == ch["steps"][steps - 1]["sigma"]
    print(fname, precision, pipeline, "steps", n_cmp, "classes", cls, "worst logR rel err", worst_logr, "soft deviations", n_soft)
    assert n_cmp > 300
    assert cls["compared"] + cls["rank_both"] >= (0.9 if precision == "fp64" else 0.6) * n_cmp, cls


@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_philox_run_replayed_in_oracle(precision):
    """The phase API with its own RNG: GPU draws (Philox) are recorded and the oracle must reach the same trees/decisions."""
    rng = np.random.default_rng(11)
    X = rng.uniform(-3, 3, (300, 3))
    y = 1.35 * X[:, 0] * X[:, 1] + 5.5 * np.sin((X[:, 0] - 1) * (X[:, 1] - 1))
    st = H.replay_gpu_run_in_oracle(X, y, K=3, n_chains=48, sweeps=25, seed=123, precision=precision)
    print("philox replay", precision, st)
    assert st["proposals"] >= 48 * 3 * 25 * 0.9
    assert st["tree_mismatch"] == 0 and st["scalar_mismatch"] == 0 and st["state_mismatch"] == 0
    assert st["rank_mismatch"] == 0 and st["logr_mismatch"] == 0 and st["decision_mismatch"] == 0 and st["nonfinite_mismatch"] == 0
    assert st["compared_share"] + st["rank_both_share"] > (0.9 if precision == "fp64" else 0.6), st


def test_chain_results_do_not_depend_on_sharding():
    """Chains are keyed by global id: running ids [0,64) in one engine or as [0,32)+[32,64) gives identical bits."""
    rng = np.random.default_rng(2)
    X = rng.uniform(-3, 3, (500, 2))
    y = X[:, 0] ** 2 - X[:, 1]
    outs = []
    for lo, hi in ((0, 64), (0, 32), (32, 64)):
        eng = H.default_engine(3, hi - lo, 2, chain_offset=lo)
        eng.set_data(X, y)
        eng.init_chains(99)
        eng.run(20)
        outs.append((eng.get_trees(current=True), eng.get_stats()))
        eng.close()
    (tok, pa, pb, nn), st = outs[0]
    (tok1, pa1, pb1, nn1), st1 = outs[1]
    (tok2, pa2, pb2, nn2), st2 = outs[2]
    assert np.array_equal(tok, np.concatenate([tok1, tok2])) and np.array_equal(nn, np.concatenate([nn1, nn2]))
    assert np.array_equal(pa, np.concatenate([pa1, pa2])) and np.array_equal(pb, np.concatenate([pb1, pb2]))
    assert np.array_equal(st["sigma"], np.concatenate([st1["sigma"], st2["sigma"]]))
    assert np.array_equal(st["beta"], np.concatenate([st1["beta"], st2["beta"]]))
    assert st["counters"][:, 1].sum() > 0


@pytest.mark.parametrize("sequential", [True, False])
@pytest.mark.parametrize("K,n,d", [(3, 1000, 2), (2, 333, 3), (5, 700, 8)])
def test_stream_groups_do_not_change_results(K, n, d, sequential):
    """bsr_run can pipeline chain groups on separate streams; chains are independent, so any grouping is bit-identical
    (both for the proposal-by-proposal pipeline and for the window path, tests/test_gpu_window.py has more of the latter)."""
    rng = np.random.default_rng(K * 100 + d)
    X = rng.uniform(-3, 3, (n, d))
    y = np.sin(X[:, 0]) * X[:, 1] + 0.5 * X[:, d - 1] ** 2
    res = {}
    for groups in (1, 4):
        eng = H.default_engine(K, 96 if sequential else 1024, d, val=25, plateau=True)     # stop rules active: done-masks must agree too
        eng.set_pipeline(sequential)
        eng.set_data(X, y)
        eng.init_chains(4242)
        eng.set_launch_geometry(0, groups)
        eng.run(7)
        eng.run(23)
        res[groups] = (eng.get_trees(current=True), eng.get_trees(current=False), eng.get_stats(), eng.get_err_trace(), eng.launch_count())
        eng.close()
    a, b = res[1], res[4]
    for i in range(4):
        assert np.array_equal(a[0][i], b[0][i]) and np.array_equal(a[1][i], b[1][i])
    for key in ("sigma", "sa", "sb", "beta", "sse", "done", "nerr"):
        assert np.array_equal(a[2][key], b[2][key], equal_nan=True), key
    assert np.array_equal(a[2]["counters"], b[2]["counters"])
    assert np.array_equal(a[3], b[3])
    assert a[2]["counters"][:, 1].sum() > 0 and a[2]["done"].sum() > 0
    if sequential:
        assert a[4] == 30 * 4 and b[4] == 30 * 4 * 4      # propose, eval fp32, eval fp64 (flagged chains), resolve
    else:
        assert 0 < a[4] < b[4]
