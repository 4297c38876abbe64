"""GPU tests of the production path of ``bsr_run``: speculative proposal windows (csrc/bsr_window.cuh).

A window generates W consecutive proposals of a chain from one live state and consumes them up to the first accept,
so the chain it produces must be the one the proposal-by-proposal pipeline produces (codes/bsr_class.py:174-252 run
one newProp at a time) for every window size, every split of the run into calls, and both precisions; and the
recorded draws of every consumed proposal must replay through the oracle."""
import numpy as np
import pytest

import parity_helpers as H

pytestmark = pytest.mark.gpu


def _data(n, d, seed, target="f1"):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-3, 3, (n, d))
    if target == "f1":
        y = 2.5 * X[:, 0] ** 4 - 1.3 * X[:, 0] ** 3 + 0.5 * X[:, 1] ** 2 - 1.7 * X[:, 1]
    else:
        y = 1.35 * X[:, 0] * X[:, 1] + 5.5 * np.sin((X[:, 0] - 1) * (X[:, 1] - 1))
    return X, y


def _run(X, y, K, C, sweeps, seed, precision="fp32", sequential=False, window=32, chunks=None, val=0, plateau=False,
         groups=0, err_cap=512):
    eng = H.default_engine(K, C, X.shape[1], precision=precision, val=val, plateau=plateau, err_cap=err_cap)
    eng.set_pipeline(sequential)
    if not sequential:
        eng.set_window(window)
    if groups:
        eng.set_launch_geometry(n_groups=groups)
    eng.set_data(X, y)
    eng.init_chains(seed)
    for n in (chunks or [sweeps]):
        eng.run(n)
    out = dict(cur=eng.get_trees(current=True), rep=eng.get_trees(current=False), st=eng.get_stats(), err=eng.get_err_trace())
    eng.close()
    return out


def _same_chains(a, b, rel=0.0, strict=True):
    """Number of chains that differ between two runs: live trees, reported trees, sigma's, accept / proposal / sweep
    counters, stop flags (always exact); the rank-reject counter too when strict; SSE within rel, beta within 100 rel.

    Window runs against each other are compared with rel = 0 (bit-identical).  Against the sequential pipeline in fp32
    the numbers may differ at the fp32 evaluation level: when any column of a sweep leaves the fp32 range the sequential
    pipeline re-evaluates ALL columns of that chain in fp64, the window kernels only the column concerned, and the Gram
    systems are ill-conditioned enough to turn 1e-7 in a column into 1e-3 in the SSE."""
    C = a["st"]["sigma"].shape[0]
    cols = [0, 1, 2, 3, 7] if strict else [0, 1, 3, 7]
    diff = 0
    for c in range(C):
        same = all(np.array_equal(x[c], y[c]) for x, y in zip(a["cur"], b["cur"]))
        same = same and all(np.array_equal(x[c], y[c]) for x, y in zip(a["rep"], b["rep"]))
        same = same and a["st"]["sigma"][c] == b["st"]["sigma"][c]
        same = same and np.array_equal(a["st"]["sa"][c], b["st"]["sa"][c]) and np.array_equal(a["st"]["sb"][c], b["st"]["sb"][c])
        same = same and np.array_equal(a["st"]["counters"][c][cols], b["st"]["counters"][c][cols])
        same = same and a["st"]["done"][c] == b["st"]["done"][c] and a["st"]["nerr"][c] == b["st"]["nerr"][c]
        if same:
            same = np.allclose(a["st"]["beta"][c], b["st"]["beta"][c], rtol=100 * rel, atol=1e-300 + rel, equal_nan=True) and \
                np.allclose(a["st"]["sse"][c], b["st"]["sse"][c], rtol=rel, atol=0.0, equal_nan=True)
        diff += (not same)
    return diff


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_window_matches_sequential_pipeline(precision):
    X, y = _data(500, 2, 3)
    K, C, sweeps = 3, 256, 40
    seq = _run(X, y, K, C, sweeps, seed=11, precision=precision, sequential=True)
    win = _run(X, y, K, C, sweeps, seed=11, precision=precision)
    acc = int(seq["st"]["counters"][:, 1].sum())
    assert acc > 50, acc
    d = _same_chains(seq, win, rel=1e-9) if precision == "fp64" else _same_chains(seq, win, rel=1e-2, strict=False)
    print("accepts", acc, "chains that differ", d)
    # fp64: identical chains.  fp32: see _same_chains; a decision sitting at its threshold may flip
    assert d <= (0 if precision == "fp64" else C // 64)


@pytest.mark.parametrize("window,chunks", [(1, None), (4, None), (7, [1, 2, 30, 7]), (32, [13, 27]), (32, [1] * 40), (64, None), (45, [9, 31]),
                                           (64, [1] * 40)])
def test_window_size_and_call_split_do_not_change_chains(window, chunks):
    X, y = _data(333, 2, 4, target="sim")
    K, C, sweeps = 3, 128, 40
    ref = _run(X, y, K, C, sweeps, seed=5, window=32)
    got = _run(X, y, K, C, sweeps, seed=5, window=window, chunks=chunks)
    assert _same_chains(ref, got, rel=0.0) == 0


def test_window_groups_and_k5_d8():
    rng = np.random.default_rng(8)
    X = rng.uniform(-3, 3, (1201, 8))
    y = X[:, 0] * X[:, 3] - np.cos(X[:, 5]) + 0.1 * rng.normal(size=1201)
    K, C, sweeps = 5, 600, 12
    a = _run(X, y, K, C, sweeps, seed=2, groups=1)
    b = _run(X, y, K, C, sweeps, seed=2, groups=3, window=9)
    s = _run(X, y, K, C, sweeps, seed=2, sequential=True)
    assert _same_chains(a, b, rel=0.0) == 0
    assert _same_chains(a, s, rel=1e-2, strict=False) <= 600 // 64


def test_window_generic_k_and_row_tiles(monkeypatch):
    """K = 7 takes the generic (capacity 16) kernels; a small tile and forced row splits exercise the tile loop and the
    per-split partial records."""
    rng = np.random.default_rng(9)
    X = rng.uniform(-2, 2, (2999, 3))
    y = np.sin(X[:, 0]) + X[:, 1] ** 2 + 0.05 * rng.normal(size=2999)
    K, C, sweeps = 7, 64, 6
    a = _run(X, y, K, C, sweeps, seed=21)
    monkeypatch.setenv("BSR_WIN_TILE", "256")
    monkeypatch.setenv("BSR_WIN_SPLITS", "3")
    b = _run(X, y, K, C, sweeps, seed=21)
    monkeypatch.delenv("BSR_WIN_TILE")
    monkeypatch.delenv("BSR_WIN_SPLITS")
    s = _run(X, y, K, C, sweeps, seed=21, sequential=True)
    assert _same_chains(a, b, rel=1e-6, strict=False) <= 1      # other tile / split geometry: same sums in another order
    assert _same_chains(a, s, rel=1e-2, strict=False) <= 1


def test_window_stop_rules_match_sequential():
    """val (consecutive rejections, checked at sweep boundaries, bsr_class.py:174) and the plateau rule with its Q16
    snapshot (bsr_class.py:248-252) must end every chain at the same proposal as the sequential pipeline."""
    X, y = _data(200, 2, 6)
    K, C = 3, 256
    for val, plateau, sweeps in ((7, False, 60), (40, True, 400)):
        seq = _run(X, y, K, C, sweeps, seed=31, sequential=True, val=val, plateau=plateau, err_cap=128)
        win = _run(X, y, K, C, sweeps, seed=31, val=val, plateau=plateau, err_cap=128)
        win64 = _run(X, y, K, C, sweeps, seed=31, val=val, plateau=plateau, err_cap=128, window=64)
        assert _same_chains(win, win64, rel=0.0) == 0
        nd = int(seq["st"]["done"].sum())
        print("val", val, "plateau", plateau, "done chains", nd, "of", C)
        assert nd > C // 2
        assert _same_chains(seq, win, rel=1e-2, strict=False) <= C // 64
        if _same_chains(seq, win, rel=1e-2, strict=False) == 0:
            assert np.allclose(seq["err"], win["err"], rtol=1e-2, equal_nan=True)


@pytest.mark.parametrize("window", [32, 64])
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_window_run_replays_through_oracle(precision, window):
    X, y = _data(300, 2, 12)
    st = H.replay_window_run_in_oracle(X, y, K=3, n_chains=48, sweeps=25, seed=77, precision=precision, run_chunks=(4, None),
                                       window=window)
    print(st)
    assert st["proposals"] >= 48 * 3 * 25 * 0.9
    assert st["accepts"] > 10
    assert st["scalar_mismatch"] == 0 and st["tree_mismatch"] == 0 and st["state_mismatch"] == 0 and st["counter_mismatch"] == 0
    assert st["rank_mismatch"] == 0 and st["decision_mismatch"] == 0 and st["logr_mismatch"] == 0 and st["nonfinite_mismatch"] == 0
    assert st["compared_share"] + st["rank_both_share"] > (0.9 if precision == "fp64" else 0.6), st


def test_window_ragged_rows_repeatable():
    """n % 4 in {1, 2}: the last 4-row vector of the tile has padding rows.  Regression test for a padding slot of the
    shared-memory tile that was never written (0 * stale NaN bits = NaN in p.y, about once per 1e6 proposals): repeated
    runs in one process (so that shared memory holds whatever the previous kernels left) must agree bit for bit."""
    rng = np.random.default_rng(203)
    for n in (333, 334):
        X = rng.uniform(-3, 3, (n, 3))
        y = np.sin(X[:, 0]) * X[:, 1] + 0.5 * X[:, 2] ** 2
        ref = _run(X, y, 2, 1024, 30, seed=4242, val=25, plateau=True)
        for rep in range(5):
            got = _run(X, y, 2, 1024, 30, seed=4242, val=25, plateau=True, groups=1 + 3 * (rep % 2))
            assert _same_chains(ref, got, rel=0.0) == 0


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("tiled", [False, True])
def test_shared_records_of_repeated_trees_do_not_change_chains(monkeypatch, precision, tiled):
    """The proposals of a window start from one live state, so many are the same tree; k_weval interprets a repeated
    tree once and shares its record, and a tree that the chain's previous window already held (same live state: no accept
    in between) takes its record from there (csrc/bsr_window.cuh: dedup_window).  The chains must be bit-identical to a run
    that interprets every proposal (BSR_WIN_NO_DEDUP), in the one-tile geometry (in-block fp64 pass) and with row
    tiles / splits, for any split of the run into calls; the counter of executed node evaluations must drop,
    the reference-equivalent one not."""
    X, y = _data(1000, 2, 5, target="sim")
    K, C, sweeps = 3, 256, 60
    if tiled:
        monkeypatch.setenv("BSR_WIN_TILE", "256")
        monkeypatch.setenv("BSR_WIN_SPLITS", "2")
    a = _run(X, y, K, C, sweeps, seed=33, precision=precision, window=64, chunks=[7, 1, 52])
    monkeypatch.setenv("BSR_WIN_NO_CACHE", "1")                   # repeated trees within a window only
    m = _run(X, y, K, C, sweeps, seed=33, precision=precision, window=64)
    monkeypatch.delenv("BSR_WIN_NO_CACHE")
    monkeypatch.setenv("BSR_WIN_NO_DEDUP", "1")                   # every slot interpreted
    b = _run(X, y, K, C, sweeps, seed=33, precision=precision, window=64)
    assert _same_chains(a, b, rel=0.0) == 0 and _same_chains(m, b, rel=0.0) == 0
    ca, cm, cb = a["st"]["counters"], m["st"]["counters"], b["st"]["counters"]
    assert np.array_equal(ca[:, 4], cb[:, 4]) and np.array_equal(cm[:, 4], cb[:, 4])     # out-of-range proposals counted alike
    assert np.array_equal(ca[:, 5], cb[:, 5]) and np.array_equal(cm[:, 5], cb[:, 5])     # reference-equivalent node evaluations
    # executed node evaluations: fewer with the in-window search, fewer still with the previous window as a record cache
    assert cm[:, 6].sum() < 0.9 * cb[:, 6].sum() and ca[:, 6].sum() < 0.9 * cm[:, 6].sum()
    print("executed node-row evaluations: all slots %d, in-window duplicates shared %d, previous window as cache %d" % (cb[:, 6].sum(), cm[:, 6].sum(), ca[:, 6].sum()))


def test_out_of_range_columns_do_not_depend_on_tiles_or_splits(monkeypatch):
    """The value rule of the fp32 mode (csrc/bsr_window.cuh: vec_finite / eval_tree_wide2): a 4-row vector whose fp32
    interpretation is not finite is interpreted in double range, every other vector stays fp32 -- whatever the row tile, the
    row split or the role of the tree (live or proposed).  With inputs on (-30, 30) a large share of the proposals overflow
    fp32 on some rows (cubes of cubes, products of exponentials); the chains of a run with small tiles and three row splits
    must be the chains of the one-tile run (sums in another order: SSE to 1e-9), including the out-of-range counter."""
    rng = np.random.default_rng(77)
    X = rng.uniform(-30, 30, (1500, 2))
    y = 0.02 * X[:, 0] ** 3 - 3.0 * X[:, 1] + 40 * np.sin(0.3 * X[:, 0])
    K, C, sweeps = 3, 192, 40
    a = _run(X, y, K, C, sweeps, seed=8, window=64)
    monkeypatch.setenv("BSR_WIN_TILE", "128")
    monkeypatch.setenv("BSR_WIN_SPLITS", "3")
    b = _run(X, y, K, C, sweeps, seed=8, window=64)
    monkeypatch.delenv("BSR_WIN_TILE")
    monkeypatch.delenv("BSR_WIN_SPLITS")
    wide = int(a["st"]["counters"][:, 4].sum())
    props = int(a["st"]["counters"][:, 0].sum())
    print("proposals", props, "needed the double range", wide, "accepts", int(a["st"]["counters"][:, 1].sum()))
    assert wide > 0.03 * props
    d = _same_chains(a, b, rel=1e-9)
    assert d <= 1, d                                     # (a decision sitting exactly at its threshold may see the other summation order)
    assert np.array_equal(a["st"]["counters"][:, 4], b["st"]["counters"][:, 4]) or d == 1


def test_window_geometry_rule():
    """Tile and split geometry of the evaluation kernels (csrc/bsr_tu_window.cu: win_geometry): 1024-row tiles where the resident
    blocks still fit (K = 10: two blocks of <= 972 rows), one split when the chains alone fill the GPU or the rows are few, ~28 waves
    of (chain, split) blocks when few chains share many rows -- and whole tiles per split."""
    def geom(K, C, n, d=3):
        eng = H.default_engine(K, C, d)
        rng = np.random.default_rng(1)
        X = rng.uniform(-1, 1, (n, d))
        eng.set_data(X, X[:, 0] + X[:, 1] ** 2)
        g = eng.window_geometry()
        eng.close()
        return g
    g = geom(3, 4096, 1000)
    assert g["splits"] == 1 and g["tile_rows"] == 1000 and g["window"] == 64
    g = geom(5, 2048, 5000)
    assert g["splits"] == 1 and g["tile_rows"] == 1024
    g = geom(10, 512, 10000)
    assert g["splits"] == 1 and 900 <= g["tile_rows"] <= 972 and g["tile_rows"] % 4 == 0
    g = geom(5, 64, 400000)                       # few chains, many rows: 12432 / 64 = 195 blocks wanted per chain, 8192 rows at least per split
    assert g["tile_rows"] == 1024 and g["rows_per_split"] % 1024 == 0 and g["rows_per_split"] >= 8192
    assert 40 <= g["splits"] <= 49 and (g["splits"] - 1) * g["rows_per_split"] < 400000 <= g["splits"] * g["rows_per_split"]


def test_large_trees_checked_first_pass(monkeypatch):
    """K > 5 kernels: for trees of 12+ nodes the first pass of k_weval checks every 4-row vector itself, keeps the sums of the
    finite ones and leaves a bitmap of the others for the second pass (csrc/bsr_window.cuh).  On 31-node transcendental trees
    (the C3 initial state of bench.py) with inputs that overflow fp32 on many rows, the chains must be those of a run with
    the check switched off (where the second pass re-interprets the whole tile in fp32 first): same trees, same counters -- the
    out-of-range counter included --, sums in another order."""
    import bench
    w = dict(bench.WORKLOADS["c3"], n=2500, chains=96)
    rng = np.random.default_rng(5)
    X = rng.uniform(-3, 3, (w["n"], w["d"]))
    y = np.sin(X[:, 0]) * np.exp(0.5 * X[:, 1]) + X[:, 2] * X[:, 3] + 0.1 * rng.normal(size=w["n"])
    state = bench.deep_state(w, w["chains"], 0)

    def run():
        from mcmc_symreg_b200 import capi
        ops = w["ops"]
        eng = capi.Engine(w["K"], w["chains"], ops, [1.0 / len(ops)] * len(ops), beta=-1.0, val=0, plateau_rule=False)
        eng.set_data(X, y)
        eng.set_state(*state, seed=11)
        eng.run(6)
        out = dict(cur=eng.get_trees(current=True), rep=eng.get_trees(current=False), st=eng.get_stats(), err=eng.get_err_trace())
        eng.close()
        return out
    a = run()
    monkeypatch.setenv("BSR_WIN_NO_CHECKED", "1")
    b = run()
    monkeypatch.delenv("BSR_WIN_NO_CHECKED")
    c = a["st"]["counters"]
    print("proposals", int(c[:, 0].sum()), "out of range", int(c[:, 4].sum()), "accepts", int(c[:, 1].sum()), "rank rejects", int(c[:, 2].sum()))
    assert c[:, 4].sum() > 0.1 * c[:, 0].sum()                   # the path under test is taken by a good share of the proposals
    assert np.array_equal(c[:, 4], b["st"]["counters"][:, 4])
    assert _same_chains(a, b, rel=1e-9) <= 1
