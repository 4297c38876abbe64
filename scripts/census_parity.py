#!/usr/bin/env python
"""Rank / decision mismatch census of the production path against the oracle.

The GPU sampler (bsr_run: speculative windows, Philox) records every draw of every consumed proposal; each chain is
then replayed proposal by proposal through the oracle (oracle/bsr_oracle.py, pinned to the unmodified reference) and
every proposal is classified (tests/parity_helpers.py: replay_chain_in_oracle).  Prints one JSON object per
configuration: the number of proposals, the share whose logR was compared against the tolerance, the shares excluded
and why, and the rank / decision mismatches with the trees involved.

    python scripts/census_parity.py [--shapes c2,c4] [--proposals 100000] [--precision fp32] [--out gpurun_out/census.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SHAPES = {
    # BASELINE.json configs[0..3] shapes (rows, features, K) with the targets bench.py uses
    "c1": dict(n=100, d=2, K=3, target="f1", seed=1001),
    "c2": dict(n=1000, d=2, K=3, target="sim", seed=2001),
    "c4": dict(n=5000, d=8, K=5, target="mix8", seed=4001),
    "c3": dict(n=10000, d=8, K=10, target="mix8", seed=3001),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="c2,c4")
    ap.add_argument("--proposals", type=int, default=100000)
    ap.add_argument("--sweeps", type=int, default=60)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--window", type=int, default=64)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "census.json"))
    ap.add_argument("--pipeline", default="window", choices=["window", "sequential"])
    args = ap.parse_args()
    import __graft_entry__ as g
    g.build()
    import bench
    import parity_helpers as H
    res = {}
    for name in args.shapes.split(","):
        w = SHAPES[name]
        X, y = bench.make_data(dict(w))
        K = w["K"]
        chains = max(1, -(-args.proposals // (args.sweeps * K)))
        t0 = time.perf_counter()
        if args.pipeline == "window":
            st = H.replay_window_run_in_oracle(X, y, K=K, n_chains=chains, sweeps=args.sweeps, seed=w["seed"] + 7, precision=args.precision,
                                               window=args.window, run_chunks=(7, None), procs=args.procs, detail=True)
        else:
            st = H.replay_gpu_run_in_oracle(X, y, K=K, n_chains=chains, sweeps=args.sweeps, seed=w["seed"] + 7, precision=args.precision,
                                            procs=args.procs, detail=True)
        st["wall_s"] = time.perf_counter() - t0
        st["config"] = dict(shape=name, chains=chains, sweeps=args.sweeps, precision=args.precision, window=args.window, **w)
        p = max(1, st["proposals"])
        st["rates"] = dict((key, st[key] / p) for key in ("logr_compared", "rank_both", "type_limited", "nonfinite", "rank_mismatch",
                                                          "decision_mismatch", "logr_mismatch", "nonfinite_mismatch", "rank_soft",
                                                          "decision_soft"))
        res[name] = st
        short = dict((k, v) for k, v in st.items() if k != "details")
        print(json.dumps(short), flush=True)
        for dd in st["details"][:60]:
            print("   ", json.dumps(dict((k, v) for k, v in dd.items() if k not in ("state", "proposed_enc"))), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
