"""Shared helpers for the parity tests: oracle <-> device encodings and replay drivers."""
import numpy as np

from oracle import bsr_oracle as O

MAX_NODES = 64


def enc_tree(t):
    """oracle Tree -> (tok, pa, pb, n) device encoding (include/bsr_b200.h)."""
    tok = np.zeros(MAX_NODES, dtype=np.uint32)
    pa = np.zeros(MAX_NODES)
    pb = np.zeros(MAX_NODES)
    n = len(t)
    assert n <= MAX_NODES
    for i in range(n):
        tok[i] = t.op[i] | (t.oi[i] << 8) | (t.ft[i] << 16)
        if t.op[i] == O.OP_LT:
            pa[i], pb[i] = t.a[i], t.b[i]
    return tok, pa, pb, n


def dec_tree(tok, pa, pb, n):
    """device encoding -> oracle Tree"""
    n = int(n)
    op = [int(t) & 0xFF for t in tok[:n]]
    oi = [(int(t) >> 8) & 0xFF for t in tok[:n]]
    ft = [int(t) >> 16 for t in tok[:n]]
    a = [float(pa[i]) if op[i] == O.OP_LT else 0.0 for i in range(n)]
    b = [float(pb[i]) if op[i] == O.OP_LT else 0.0 for i in range(n)]
    return O.Tree(op, oi, ft, a, b)


def enc_golden(t):
    return dict(op=list(t.op), oi=list(t.oi), ft=list(t.ft), a=list(t.a), b=list(t.b))


def tree_from_golden(enc):
    return O.Tree(enc["op"], enc["oi"], enc["ft"], enc["a"], enc["b"])


def pack_state(chains_trees, K):
    """list over chains of K oracle Trees -> arrays for Engine.set_state"""
    C = len(chains_trees)
    tok = np.zeros((C, K, MAX_NODES), dtype=np.uint32)
    pa = np.zeros((C, K, MAX_NODES))
    pb = np.zeros((C, K, MAX_NODES))
    nn = np.zeros((C, K), dtype=np.int32)
    for c in range(C):
        for k in range(K):
            tok[c, k], pa[c, k], pb[c, k], nn[c, k] = enc_tree(chains_trees[c][k])
    return tok, pa, pb, nn


def trees_equal(t1, t2, params_rel=0.0):
    if t1.op != t2.op or t1.ft != t2.ft or t1.oi != t2.oi:
        return False
    for i, o in enumerate(t1.op):
        if o == O.OP_LT:
            for x, y in ((t1.a[i], t2.a[i]), (t1.b[i], t2.b[i])):
                if not (x == y or abs(x - y) <= params_rel * max(abs(x), abs(y))):
                    return False
    return True


def close(a, b, rel, abs_=0.0):
    if np.isnan(a) or np.isnan(b):
        return bool(np.isnan(a) and np.isnan(b))
    if np.isinf(a) or np.isinf(b):
        return a == b
    return abs(a - b) <= abs_ + rel * max(abs(a), abs(b))


def well_conditioned(t, X):
    """fp32 cannot hold the argument of sin/cos/exp to 1e-4 absolute once it exceeds ~1e3."""
    st = []
    ok = True
    with np.errstate(all="ignore"):
        for i in range(len(t) - 1, -1, -1):
            o = t.op[i]
            if o == O.OP_LEAF:
                st.append(X[:, t.ft[i]].astype(float))
                continue
            if o in (O.OP_ADD, O.OP_MUL):
                l, r = st.pop(), st.pop()
                v = l + r if o == O.OP_ADD else l * r
                if o == O.OP_ADD and np.any(np.abs(v) < 1e-3 * (np.abs(l) + np.abs(r))):
                    ok = False            # cancellation
            else:
                a = st.pop()
                if o in (O.OP_SIN, O.OP_COS, O.OP_EXP) and np.max(np.abs(a)) > 50:
                    ok = False
                if o == O.OP_INV and np.min(np.abs(a)) < 1e-3:
                    ok = False
                sub = O.Tree([o, 0], [0, 0], [0, 0], [t.a[i], 0], [t.b[i], 0])
                v = O.eval_tree(sub, a.reshape(-1, 1))
                if o == O.OP_LT and np.any(np.abs(v) < 1e-3 * (np.abs(t.a[i] * a) + abs(t.b[i]))):
                    ok = False
            st.append(v)
    return ok




def default_engine(K, C, d, precision="fp32", val=0, plateau=False, beta=-1.0, weights=None, chain_offset=0, err_cap=512):
    from mcmc_symreg_b200 import capi
    ops = list(range(1, 11))
    w = weights if weights is not None else [0.1] * 10
    return capi.Engine(K, C, ops, w, beta=beta, val=val, plateau_rule=plateau, precision=precision, chain_offset=chain_offset,
                       err_cap=err_cap)


def replay_gpu_run_in_oracle(X, y, K, n_chains, sweeps, seed, precision="fp32", beta=-1.0, logr_rel=None, chain_offset=0):
    """Run the GPU sampler with its own Philox stream while recording every drawn value, then replay each chain
    through the oracle with those values and compare step by step.  Returns mismatch statistics."""
    from mcmc_symreg_b200 import capi
    TR = capi.TR
    d = X.shape[1]
    eng = default_engine(K, n_chains, d, precision=precision, beta=beta, chain_offset=chain_offset)
    eng.set_data(X, y)
    eng.init_chains(seed)
    tok0, pa0, pb0, nn0 = eng.get_trees(current=True)
    st0 = eng.get_stats()
    steps = sweeps * K
    eng.record_draws(steps, 256)
    eng.set_tape(None, steps)           # Philox, but keep a trace
    props = []
    for s in range(sweeps):
        eng.sweep_propose()
        props.append(eng.get_proposals())
        eng.sweep_eval()
        eng.sweep_resolve()
    trace = eng.get_trace(steps)
    rec, cnt = eng.get_recorded_draws()
    tokf, paf, pbf, nnf = eng.get_trees(current=True)
    stf = eng.get_stats()
    eng.close()

    cfg = O.Config(n_feature=d, beta=beta)
    if logr_rel is None:
        logr_rel = 1e-6 if precision == "fp64" else 1e-3
    out = dict(logr_compared=0, proposals=0, accepts=0, tree_mismatch=0, scalar_mismatch=0, logr_mismatch=0, decision_mismatch=0,
               rank_mismatch=0, diverged_chains=0, state_mismatch=0, max_logr_err=0.0)
    for c in range(n_chains):
        trees = [dec_tree(tok0[c, k], pa0[c, k], pb0[c, k], nn0[c, k]) for k in range(K)]
        sigma = float(st0["sigma"][c])
        sa, sb = list(st0["sa"][c]), list(st0["sb"][c])
        diverged = False
        for s in range(steps):
            k = s % K
            t = trace[c, s]
            tape = list(rec[c, s, :cnt[c, s]])
            if not t[TR["rank_reject"]] and not (int(t[TR["flags"]]) & 1):
                tape.append(float(t[TR["u"]]))
            dr = O.TapeDraws(tape)
            acc, sigma, newt, sa[k], sb[k], tr = O.new_prop(trees, k, sigma, y, X, cfg, sa[k], sb[k], dr)
            out["proposals"] += 1
            ptok, ppa, ppb, pnn = props[s // K]
            gp = dec_tree(ptok[c, k], ppa[c, k], ppb[c, k], pnn[c, k])
            if not trees_equal(gp, tr.proposed, params_rel=1e-13):
                out["tree_mismatch"] += 1
            if tr.change != int(t[TR["change"]]) or tr.move != int(t[TR["move"]]) or not close(tr.Q, t[TR["Q"]], 1e-9) \
                    or not close(tr.Qinv, t[TR["Qinv"]], 1e-9):
                out["scalar_mismatch"] += 1
            if tr.change != 0 and (not close(tr.hratio, t[TR["hratio"]], 1e-7, 1e-300) or not close(tr.detjacob, t[TR["detjacob"]], 1e-12)):
                out["scalar_mismatch"] += 1
            if bool(t[TR["rank_reject"]]) != tr.rank_deficient:
                out["rank_mismatch"] += 1
            elif not tr.rank_deficient and np.isfinite(tr.logR) and np.isfinite(t[TR["logR"]]) and \
                    all(well_conditioned(x, X) for x in [tr.proposed] + list(trees)):
                # logR is a difference of two log-likelihoods: bound the error relative to their magnitude
                err = abs(tr.logR - t[TR["logR"]]) / max(1.0, abs(tr.logR), abs(tr.yll_new), abs(tr.yll_old))
                out["logr_compared"] += 1
                out["max_logr_err"] = max(out["max_logr_err"], err)
                if err > logr_rel:
                    out["logr_mismatch"] += 1
            gacc = bool(t[TR["accepted"]])
            if gacc != acc:
                out["decision_mismatch"] += 1
                diverged = True
                break
            if acc:
                out["accepts"] += 1
                trees = list(trees)
                trees[k] = newt
        if diverged:
            out["diverged_chains"] += 1
            continue
        for k in range(K):
            gt = dec_tree(tokf[c, k], paf[c, k], pbf[c, k], nnf[c, k])
            if not trees_equal(gt, trees[k], params_rel=1e-13):
                out["state_mismatch"] += 1
        if not close(sigma, float(stf["sigma"][c]), 1e-12):
            out["state_mismatch"] += 1
    return out


def replay_window_run_in_oracle(X, y, K, n_chains, sweeps, seed, precision="fp32", beta=-1.0, logr_rel=None, window=32,
                                run_chunks=(None,)):
    """Same as replay_gpu_run_in_oracle, for the production path: ``Engine.run`` (speculative windows) records the
    draws and a trace row of every CONSUMED proposal; each chain is then replayed proposal by proposal through the
    oracle.  run_chunks: sizes of the successive run() calls (None = all sweeps in one call)."""
    from mcmc_symreg_b200 import capi
    TR = capi.TR
    d = X.shape[1]
    eng = default_engine(K, n_chains, d, precision=precision, beta=beta)
    eng.set_window(window)
    eng.set_data(X, y)
    eng.init_chains(seed)
    tok0, pa0, pb0, nn0 = eng.get_trees(current=True)
    st0 = eng.get_stats()
    steps = sweeps * K
    eng.record_draws(steps, 256)
    eng.set_tape(None, steps)           # Philox, but keep a trace
    left = sweeps
    for ch in run_chunks:
        n = left if ch is None else min(ch, left)
        if n > 0:
            eng.run(n)
        left -= n
    if left > 0:
        eng.run(left)
    trace = eng.get_trace(steps)
    rec, cnt = eng.get_recorded_draws()
    tokf, paf, pbf, nnf = eng.get_trees(current=True)
    stf = eng.get_stats()
    eng.close()

    cfg = O.Config(n_feature=d, beta=beta)
    if logr_rel is None:
        logr_rel = 1e-6 if precision == "fp64" else 1e-3
    out = dict(logr_compared=0, proposals=0, accepts=0, scalar_mismatch=0, logr_mismatch=0, decision_mismatch=0,
               rank_mismatch=0, diverged_chains=0, state_mismatch=0, max_logr_err=0.0, counter_mismatch=0)
    for c in range(n_chains):
        trees = [dec_tree(tok0[c, k], pa0[c, k], pb0[c, k], nn0[c, k]) for k in range(K)]
        sigma = float(st0["sigma"][c])
        sa, sb = list(st0["sa"][c]), list(st0["sb"][c])
        diverged = False
        n_acc = 0
        for s in range(steps):
            k = s % K
            t = trace[c, s]
            tape = list(rec[c, s, :cnt[c, s]])
            if not t[TR["rank_reject"]] and not (int(t[TR["flags"]]) & 1):
                tape.append(float(t[TR["u"]]))
            dr = O.TapeDraws(tape)
            try:
                acc, sigma, newt, sa[k], sb[k], tr = O.new_prop(trees, k, sigma, y, X, cfg, sa[k], sb[k], dr)
            except IndexError:
                # the device rejected on rank (no accept draw recorded) where the oracle wants to draw: a rank mismatch
                out["rank_mismatch"] += 1
                out.setdefault("rank_mismatch_detail", []).append((c, s, [O.express(x) for x in trees], t[TR["move"]], t[TR["m_new"]]))
                diverged = True
                break
            out["proposals"] += 1
            if tr.change != int(t[TR["change"]]) or tr.move != int(t[TR["move"]]) or not close(tr.Q, t[TR["Q"]], 1e-9) \
                    or not close(tr.Qinv, t[TR["Qinv"]], 1e-9) or len(tr.proposed) != int(t[TR["m_new"]]):
                out["scalar_mismatch"] += 1
            if tr.change != 0 and (not close(tr.hratio, t[TR["hratio"]], 1e-7, 1e-300) or not close(tr.detjacob, t[TR["detjacob"]], 1e-12)):
                out["scalar_mismatch"] += 1
            if bool(t[TR["rank_reject"]]) != tr.rank_deficient:
                out["rank_mismatch"] += 1
            elif not tr.rank_deficient and np.isfinite(tr.logR) and np.isfinite(t[TR["logR"]]) and \
                    all(well_conditioned(x, X) for x in [tr.proposed] + list(trees)):
                err = abs(tr.logR - t[TR["logR"]]) / max(1.0, abs(tr.logR), abs(tr.yll_new), abs(tr.yll_old))
                out["logr_compared"] += 1
                out["max_logr_err"] = max(out["max_logr_err"], err)
                if err > logr_rel:
                    out["logr_mismatch"] += 1
            gacc = bool(t[TR["accepted"]])
            if gacc != acc:
                out["decision_mismatch"] += 1
                diverged = True
                break
            if acc:
                n_acc += 1
                out["accepts"] += 1
                trees = list(trees)
                trees[k] = newt
        if diverged:
            out["diverged_chains"] += 1
            continue
        for k in range(K):
            gt = dec_tree(tokf[c, k], paf[c, k], pbf[c, k], nnf[c, k])
            if not trees_equal(gt, trees[k], params_rel=1e-13):
                out["state_mismatch"] += 1
        if not close(sigma, float(stf["sigma"][c]), 1e-12):
            out["state_mismatch"] += 1
        cn = stf["counters"][c]
        if int(cn[0]) != steps or int(cn[1]) != n_acc or int(cn[7]) != sweeps:
            out["counter_mismatch"] += 1
    return out
