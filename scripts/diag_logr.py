"""Diagnostic: largest fp32 logR deviations on the golden step fixtures (run on the GPU box)."""
import sys, os, gzip, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g; g.build()
from oracle import bsr_oracle as O
import parity_helpers as H
from mcmc_symreg_b200 import capi
TR = capi.TR
for fname in ["steps_f1_d2_k3.json.gz", "steps_deep_d3_k2.json.gz", "steps_mix_d8_k5.json.gz"]:
    gd = json.load(gzip.open(os.path.join(ROOT, "tests/golden", fname), "rt"))
    K, d = gd["K"], gd["d"]
    X, y = np.array(gd["X"]), np.array(gd["y"])
    chains = gd["chains"]
    steps = min(len(ch["steps"]) for ch in chains); steps -= steps % K
    rows = []
    for prec in ["fp32", "fp64"]:
        eng = H.default_engine(K, len(chains), d, precision=prec, beta=gd["beta"], weights=gd["weights"])
        eng.set_data(X, y)
        tok, pa, pb, nn = H.pack_state([[H.tree_from_golden(e) for e in ch["init"]["trees"]] for ch in chains], K)
        eng.set_state(tok, pa, pb, nn, [ch["init"]["sigma"] for ch in chains], [ch["init"]["sa"] for ch in chains], [ch["init"]["sb"] for ch in chains])
        eng.set_tape([[ch["steps"][s]["tape"] for s in range(steps)] for ch in chains], steps)
        eng.run(steps // K)
        tr = eng.get_trace(steps)
        eng.close()
        for c, ch in enumerate(chains):
            state = [H.tree_from_golden(e) for e in ch["init"]["trees"]]
            for s in range(steps):
                st = ch["steps"][s]
                if not st["rank_reject"] and np.isfinite(st["logR"]):
                    t = tr[c, s]
                    rows.append((abs(t[TR["logR"]] - st["logR"]), prec, c, s, st["logR"], t[TR["logR"]], t[TR["sse_new"]], t[TR["sse_old"]],
                                 t[TR["new_sigma"]], O.express(H.tree_from_golden(st["proposed"])), [O.express(x) for x in state]))
                if st["accepted"]:
                    state[st["count"]] = H.tree_from_golden(st["tree"])
    for prec in ["fp32", "fp64"]:
        rr = sorted([r for r in rows if r[1] == prec], key=lambda r: -r[0])
        errs = np.array([r[0] for r in rr])
        print(fname, prec, "n", len(errs), "abs err quantiles 50/90/99/max", np.quantile(errs, [0.5, 0.9, 0.99, 1.0]))
        for r in rr[:6]:
            print("   abs %.3g  c%d s%d ref %.6g got %.6g sse_new %.6g sse_old %.6g ns %.4g\n      prop %s\n      state %s" % (r[0], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[10]))
