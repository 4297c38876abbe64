"""bench.py's reference arm runs on host cores only, so its side of the output contract is checked here on CPU:
exactly one JSON line on stdout (library banners go to stderr), the keys the driver reads, and rank != 0 printing nothing."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0"]
    return subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)


def test_reference_arm_prints_one_json_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mh_proposals_per_sec" and d["unit"] == "proposals/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    # the unmodified reference when oracle/_ref is present (built by __graft_entry__.build() where /root/reference exists), else the port
    from oracle import ref_loader
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_loader.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == dict(value=d["value"], unit=d["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert d["config"]["workload"] == "c1" and d["config"]["K"] == 3 and d["config"]["chains_per_gpu"] == 50


def test_reference_arm_other_ranks_print_nothing():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ""


def test_reference_copy_is_the_reference():
    """oracle/_ref holds byte-for-byte copies (sha256 manifest) and the loader imports the reference's own package from there."""
    sys.path.insert(0, ROOT)
    from oracle import build_ref, ref_loader
    man = build_ref.build(verbose=False)
    if man is None:
        import pytest
        pytest.skip("neither /root/reference nor an earlier copy is present")
    for f, digest in man["files"].items():
        assert build_ref.sha256(os.path.join(build_ref.DEST, f)) == digest
        src = os.path.join(build_ref.REF_ROOT, "codes", f)
        if os.path.exists(src):
            assert build_ref.sha256(src) == digest
    bsr = ref_loader.load()
    assert os.path.abspath(bsr.__file__).startswith(ref_loader.REF_DIR)
    assert callable(bsr.newProp) and callable(bsr.BSR)


def test_c3_deep_initial_trees():
    """The seeded generator of the C3 workload: 31 nodes, height >= 6, operators from the transcendental set."""
    sys.path.insert(0, ROOT)
    import bench
    from oracle import bsr_oracle as O
    rng = np.random.default_rng(5)
    for _ in range(50):
        op, oi, ft, a, b = bench.deep_tree(rng, 8, bench.C3_OPS)
        t = O.Tree(op, oi, ft, a, b)
        assert len(t) == 31 and O.subtree_sizes(t.op)[0] == 31 and O.get_height(t) >= 6
        assert set(o for o in op if o != 0) <= set(bench.C3_OPS) and all(bench.C3_OPS[i] == o for o, i in zip(op, oi) if o != 0)
        assert max(ft) < 8
    w = dict(bench.WORKLOADS["c3"], n=200)
    X, y = bench.make_data(w)
    assert np.all(np.isfinite(y)) and np.std(y) > 0
    tok, pa, pb, nn, sig, sa, sb = bench.deep_state(w, 5, 100)
    assert (nn == 31).all() and tok.shape == (5, 10, 64)
    tok2 = bench.deep_state(w, 3, 102)[0]
    assert np.array_equal(tok[2:], tok2)            # keyed by global chain id: independent of the sharding
    col = bench.eval_enc((op, oi, ft, a, b), X)
    ref = O.eval_tree(t, X)
    assert np.allclose(col, ref, equal_nan=True)
