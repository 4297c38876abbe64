"""How many proposals of a speculative window repeat a tree?  (CPU, oracle only -- the measurement quoted in DESIGN.md
sections 5 and 10 and profiles/README.md r01j.)

For a few oracle chains at C2 (K = 3, n = 1000, simulations.py target) after `--sweeps` sweeps, draw `--windows` windows of
64 proposals each from the unchanged live state (Prop + IG(4) + auxProp, like newProp up to the evaluation) and count,
per window: proposals equal to the live tree they would replace, proposals that repeat an earlier slot of the same
window, and distinct trees that already occurred in the previous window / in any earlier window of the same live state.
A tree is its opcodes, the features of its leaves and the lt parameters (what k_weval's duplicate search compares)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                   # noqa: E402
from oracle import bsr_oracle as O             # noqa: E402


def key(t):
    n = len(t.op)
    return (tuple(t.op), tuple(t.ft[i] if t.op[i] == 0 else 0 for i in range(n)),
            tuple(t.a[i] for i in range(n) if t.op[i] == 2), tuple(t.b[i] for i in range(n) if t.op[i] == 2))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=12)
    ap.add_argument("--sweeps", type=int, default=300)
    ap.add_argument("--windows", type=int, default=6)
    ap.add_argument("--window", type=int, default=64)
    a = ap.parse_args()
    w = bench.WORKLOADS["c2"]
    X, y = bench.make_data(w)
    K = w["K"]
    cfg = O.Config(n_feature=w["d"])
    tot = same_as_live = repeats = distinct = in_prev = in_hist = 0
    for seed in range(a.chains):
        dr = O.GeneratorDraws(seed)
        r = O.run_chain(X, y, K, cfg, dr, val=0, max_sweeps=a.sweeps, fixed_sweeps=True)
        trees, sa, sb = r.final_state, r.sigma_a, r.sigma_b
        live = [key(t) for t in trees]
        prev, hist = set(), set()
        for win in range(a.windows):
            cur = set()
            for p in range(a.window):
                j = p % K
                pr = O.prop(trees[j], cfg, sa[j], sb[j], dr)
                dr.invgamma(4.0)
                O.aux_prop(pr, sa[j], sb[j], dr)
                k = key(pr.new)
                tot += 1
                same_as_live += (k == live[j])
                if k in cur:
                    repeats += 1
                else:
                    distinct += 1
                    if win > 0:
                        in_prev += (k in prev)
                        in_hist += (k in hist)
                cur.add(k)
            prev = cur
            hist |= cur
    later = distinct * (a.windows - 1) / a.windows
    print("proposals %d: equal to the live tree %.1f %%, repeat an earlier slot of their window %.1f %%, distinct %.1f %%"
          % (tot, 100.0 * same_as_live / tot, 100.0 * repeats / tot, 100.0 * distinct / tot))
    print("distinct trees of windows 2..%d: %.1f %% were in the previous window, %.1f %% in an earlier window of the same live state"
          % (a.windows, 100.0 * in_prev / max(later, 1), 100.0 * in_hist / max(later, 1)))


if __name__ == "__main__":
    main()
