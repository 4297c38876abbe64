"""Diagnostic: first step at which the GPU replay of a golden fit diverges from the oracle (run on the GPU box)."""
import sys, os, gzip, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g; g.build()
from oracle import bsr_oracle as O
import parity_helpers as H
from mcmc_symreg_b200 import capi
TR = capi.TR
fname = sys.argv[1] if len(sys.argv) > 1 else "fits_plateau.json.gz"
prec = sys.argv[2] if len(sys.argv) > 2 else "fp64"
gd = json.load(gzip.open(os.path.join(ROOT, "tests/golden", fname), "rt"))
X, y = np.array(gd["X"]), np.array(gd["y"]); K, d = gd["K"], gd["d"]
cfg = O.Config(n_feature=d, beta=gd["beta"])
dr = O.TapeDraws(gd["tape"])
sigma = dr.invgamma(1.0); trees, sa, sb = [], [], []
for _ in range(K):
    a_, b_ = dr.invgamma(1.0), dr.invgamma(1.0); trees.append(O.grow(0, cfg, a_, b_, dr)); sa.append(a_); sb.append(b_)
init = dict(sigma=sigma, trees=trees, sigma_a=sa, sigma_b=sb)
class Seg:
    def __init__(s, inner): s.inner, s.segments, s.start = inner, [], inner.pos
    def cut(s): s.segments.append(list(s.inner.tape[s.start:s.inner.pos])); s.start = s.inner.pos
    def __getattr__(s, n): return getattr(s.inner, n)
rec = Seg(dr)
r = O.run_chain(X, y, K, cfg, rec, val=gd["val"], init=init, on_step=rec.cut, keep_traces=True)
steps = r.n_proposals + (-r.n_proposals) % K
eng = H.default_engine(K, 1, d, precision=prec, val=gd["val"], plateau=True, beta=gd["beta"], err_cap=1024)
eng.set_data(X, y)
tok, pa, pb, nn = H.pack_state([trees], K)
eng.set_state(tok, pa, pb, nn, [sigma], [sa], [sb])
eng.set_tape([rec.segments + [[]] * (steps - len(rec.segments))], steps)
eng.run(steps // K)
tr = eng.get_trace(steps)[0]
print("oracle proposals", r.n_proposals, "accepts", r.n_accepts, "gpu counters", eng.get_stats()["counters"][0])
for s, ot in enumerate(r.traces):
    t = tr[s]
    bad = bool(t[TR["accepted"]]) != ot.accepted or bool(t[TR["rank_reject"]]) != ot.rank_deficient
    if bad:
        print("first divergence at step", s, "oracle acc", ot.accepted, "rank", ot.rank_deficient, "logR", ot.logR, "log_u", ot.log_u)
        print(" gpu acc", t[TR["accepted"]], "rank", t[TR["rank_reject"]], "logR", t[TR["logR"]], "u", t[TR["u"]], "sse_new", t[TR["sse_new"]], "sse_old", t[TR["sse_old"]], "flags", t[TR["flags"]])
        print(" proposed", O.express(ot.proposed), " yll_new", ot.yll_new, "yll_old", ot.yll_old)
        col = O.eval_tree(ot.proposed, X); print(" col oracle", col)
        tk, a_, b_, n_ = H.enc_tree(ot.proposed)
        print(" col gpu64 ", eng.eval_trees(tk, a_, b_, [n_], precision="fp64")[0])
        # innermost big argument
        for cut in range(1, len(ot.proposed)):
            sub = ot.proposed.slice(cut, cut + O.subtree_sizes(ot.proposed.op)[cut])
            v = O.eval_tree(sub, X)
            if np.max(np.abs(v)) > 1e6: print("  subtree", O.express(sub)[:60], "max|v| %.3g" % np.max(np.abs(v))); break
        break
else:
    print("no divergence")
