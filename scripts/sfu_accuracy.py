#!/usr/bin/env python
"""Measured accuracy of the device's fp32 operator arithmetic (bsr_eval.cuh: OpMath<float>) against float64, per operator
and argument range: what tests/parity_helpers.py's yardstick (oracle.eval_tree_sfu) has to model.  Uses bsr_eval_trees
(allcal for explicit trees, one-node trees over a single feature) on rows that sweep the range.

    python scripts/sfu_accuracy.py            # prints one line per (operator, range)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import __graft_entry__ as g
    g.build()
    import parity_helpers as H
    from oracle import bsr_oracle as O
    rng = np.random.default_rng(0)
    n = 200000
    ranges = [("1e-8..1e-6", 1e-8, 1e-6), ("1e-6..1e-4", 1e-6, 1e-4), ("1e-4..1e-2", 1e-4, 1e-2), ("1e-2..1", 1e-2, 1.0), ("1..10", 1.0, 10.0),
              ("10..1e3", 10.0, 1e3), ("1e3..1e5", 1e3, 1e5)]
    ops = [("sin", O.OP_SIN, np.sin), ("cos", O.OP_COS, np.cos), ("exp", O.OP_EXP, np.exp), ("inv", O.OP_INV, lambda v: 1.0 / v)]
    for rname, lo, hi in ranges:
        x = np.exp(rng.uniform(np.log(lo), np.log(hi), n)) * rng.choice([-1.0, 1.0], n)
        x = x.astype(np.float32).astype(np.float64)          # exactly representable inputs: the error measured is the operator's
        X = x.reshape(-1, 1)
        eng = H.default_engine(1, 1, 1, precision="fp32")
        eng.set_data(X, np.zeros(n))
        trees = [O.Tree([op, 0], [0, 0], [0, 0], [0, 0], [0, 0]) for _, op, _ in ops]
        tok, pa, pb, nn = H.pack_state([trees], len(trees))
        got = eng.eval_trees(tok[0], pa[0], pb[0], nn[0], precision="fp32")
        eng.close()
        for i, (name, op, fn) in enumerate(ops):
            if name == "exp" and hi > 80:
                continue
            with np.errstate(all="ignore"):
                ref = fn(x)
            err = np.abs(got[i] - ref)
            rel = err / np.maximum(np.abs(ref), 1e-300)
            print("%-4s |x| in %-11s max abs err %.3e  max rel err %.3e  (median rel %.2e)" % (name, rname, err.max(), rel.max(), np.median(rel)), flush=True)


if __name__ == "__main__":
    main()
