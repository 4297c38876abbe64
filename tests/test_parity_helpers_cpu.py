"""CPU tests of the comparison rules the GPU parity tests rely on (tests/parity_helpers.py): the classification of a
proposal by judge_step, with the oracle's own float64 result -- or a perturbed one -- standing in for the device."""
import numpy as np

from oracle import bsr_oracle as O
import parity_helpers as H


def _case(seed, n=200, d=2, K=3):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-3, 3, (n, d))
    y = 1.35 * X[:, 0] * X[:, 1] + 5.5 * np.sin((X[:, 0] - 1) * (X[:, 1] - 1))
    cfg = O.Config(n_feature=d)
    dr = O.GeneratorDraws(seed)
    sigma = dr.invgamma(1.0)
    trees, sa, sb = [], [], []
    for _ in range(K):
        a, b = dr.invgamma(1.0), dr.invgamma(1.0)
        trees.append(O.grow(0, cfg, a, b, dr)); sa.append(a); sb.append(b)
    return X, y, cfg, trees, sigma, sa, sb


def test_helpers_present():
    for name in ("default_engine", "replay_window_run_in_oracle", "replay_gpu_run_in_oracle", "replay_chain_in_oracle", "judge_step",
                 "yardstick", "state_fits", "resolves", "column_comparable", "pack_state", "dec_tree", "enc_tree", "trees_equal"):
        assert callable(getattr(H, name)), name


def test_judge_step_classes():
    """The oracle's own numbers are never a hard failure; a logR off by ten tolerances on a resolvable proposal is; classes are
    exclusive and every proposal gets one."""
    seen = set()
    n_hard_detected = 0
    for seed in range(40):
        X, y, cfg, trees, sigma, sa, sb = _case(seed)
        for k in range(3):
            rec = O.GeneratorDraws(1000 * seed + k, record=True)
            try:
                acc, s2, newt, a2, b2, tr = O.new_prop(trees, k, sigma, y, X, cfg, sa[k], sb[k], rec)
            except np.linalg.LinAlgError:
                continue
            gpu = dict(rank_reject=tr.rank_deficient, accepted=acc, logR=tr.logR)
            v = H.judge_step(trees, k, sigma, sa[k], sb[k], y, X, cfg, rec.tape, gpu, "fp32", 1e-3)
            assert v.cls in ("compared", "rank_both", "type_limited", "nonfinite")
            assert not v.hard, (seed, k, v.hard, v.cls)
            seen.add(v.cls)
            if v.cls == "compared":
                bad = dict(gpu, logR=tr.logR + 1e-2 * v.scale)
                w = H.judge_step(trees, k, sigma, sa[k], sb[k], y, X, cfg, rec.tape, bad, "fp32", 1e-3)
                assert "logR" in w.hard
                flipped = dict(gpu, accepted=not acc)
                w = H.judge_step(trees, k, sigma, sa[k], sb[k], y, X, cfg, rec.tape, flipped, "fp32", 1e-3)
                if abs(tr.log_u - min(tr.logR, 0.0)) > 1e-3 * v.scale:
                    assert "decision" in w.hard
                    n_hard_detected += 1
    assert "compared" in seen and n_hard_detected > 10


def test_yardstick_marks_what_float32_cannot_resolve():
    rng = np.random.default_rng(0)
    X = rng.uniform(-3, 3, (300, 2))
    chaotic = O.Tree([O.OP_SIN, O.OP_CUBIC, O.OP_CUBIC, O.OP_LT, 0], [0] * 5, [0, 0, 0, 0, 1], [0, 0, 0, -2.07, 0], [0, 0, 0, -1.24, 0])
    tame = O.Tree([O.OP_MUL, O.OP_SIN, 0, 0], [0] * 4, [0, 0, 0, 1], [0] * 4, [0] * 4)
    assert not H.column_comparable(chaotic, X, "fp32", 1e-4)      # sin of arguments up to 6e7
    assert H.column_comparable(tame, X, "fp32", 1e-4)
    assert H.column_comparable(tame, X, "fp64", 1e-10)
    f = H.state_fits([tame, chaotic], X, X[:, 0] ** 2, "fp32")
    assert not H.resolves(f["rmse"], 1e-4)
    f = H.state_fits([tame], X, X[:, 0] ** 2, "fp32")
    assert H.resolves(f["rmse"], 1e-4) and H.resolves(f["beta"], 2e-3, floor=1.0)


def test_fp32_yardstick_follows_the_value_rule_per_vector():
    """oracle.eval_tree_sfu models the device's fp32 mode including its value rule: a vector of four
    consecutive rows with a non-finite float32 value comes from the float64 evaluation, every other vector keeps its float32
    values -- also where float32 underflowed.  Tree: 1 / exp(1 / -(x^3)); rows are chosen so that float32 is exact-ish, overflows
    (1 / denormal), or underflows to 0 (then 1 / 0 -> 0 by the reference's guard)."""
    t = O.Tree()
    for op in (O.OP_INV, O.OP_EXP, O.OP_INV, O.OP_NEG, O.OP_CUBIC):
        t.append_tok(op)
    t.append_tok(O.OP_LEAF, ft=0)
    x = np.full(12, 1.5)
    x[5] = 100.0 ** (-1.0 / 3.0)      # 1 / -(x^3) = -100: exp is a float32 denormal, its reciprocal overflows float32 -> vector 1 (rows 4..7) is widened
    x[9] = 300.0 ** (-1.0 / 3.0)      # 1 / -(x^3) = -300: exp underflows to 0 in float32, the guard gives 1 / 0 -> 0, a finite value -> vector 2 stays float32
    X = x.reshape(-1, 1)
    got = O.eval_tree_sfu(t, X, 0)
    ref = O.eval_tree(t, X)
    assert np.all(np.isfinite(got))
    # vector 0: plain float32 accuracy
    assert np.allclose(got[:4], ref[:4], rtol=1e-5)
    # vector 1: all four rows come from the float64 evaluation -- the overflowing one (e^100, beyond float32) and its neighbours
    assert got[5] > 3.4e38 and np.isclose(got[5], ref[5], rtol=1e-4)
    assert np.allclose(got[4:8], ref[4:8], rtol=1e-4)          # (the float64 pass carries the SFU error bounds too)
    # vector 2: float32 values, the underflowed row is 0 where float64 has e^300
    assert got[9] == 0.0 and ref[9] > 1e100
    assert np.allclose(got[[8, 10, 11]], ref[[8, 10, 11]], rtol=1e-5)
