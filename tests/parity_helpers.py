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




def yardstick(precision):
    """The arithmetics a comparison in `precision` is held against, as a list of tree evaluators: for the fp32 device path
    float32 with SFU-accuracy transcendentals, their documented error bounds applied upwards, downwards and with random signs
    (oracle.eval_tree_sfu); for the fp64 path the 80-bit long double (oracle.eval_tree_ld) and float64 with the results the
    device's libm / FMA may round differently moved by those few ulp, again up, down and randomly (oracle.eval_tree_ulp)."""
    if precision == "fp32":
        return [lambda t, X: O.eval_tree_sfu(t, X, 1), lambda t, X: O.eval_tree_sfu(t, X, -1), lambda t, X: O.eval_tree_sfu(t, X, 0)]
    return [O.eval_tree_ld, lambda t, X: O.eval_tree_ulp(t, X, 1), lambda t, X: O.eval_tree_ulp(t, X, -1), lambda t, X: O.eval_tree_ulp(t, X, 0)]


MARGIN = {"fp32": 4.0, "fp64": 4.0}    # a proposal is compared when every yardstick arithmetic stays within tolerance / MARGIN of the float64 value
COL_CHAOS = 1e-3         # a column that a yardstick arithmetic reproduces no better than this (normalised by max|column|; ten times
                         # north_star's tolerance on tree outputs) is chaotic in the type: 1/sin(exp(8.8 ...)) has spikes whose height and
                         # sign depend on the last bit of the argument, and logR comparisons through it are luck either way
RANGE_LIMIT = 1e150      # the device sums squares of column values in float64: beyond this they overflow and the column counts
                         # as non-finite; the reference scales by max|.| first and goes on to 1e308
RANK_ANGLE = 2e-6        # sine of the smallest angle between a column and the span of the others that the device's Gram-based
                         # rank test resolves (pivot_tol = 3e-13 in fp32, bsr_solve.cuh, with a margin); numpy's SVD goes down to ~2e-13


def column_comparable(t, X, precision, tol):
    """A tree's column can be held to `tol` (normalised by max|column|) iff every yardstick arithmetic reproduces the
    float64 column within a quarter of it; else the deviation is the type's, whoever evaluates (cos(x^18), 1/sin(exp(.)))."""
    ref = O.eval_tree(t, X)
    if not np.all(np.isfinite(ref)):
        return False
    scale = np.max(np.abs(ref)) + 1e-300
    for ev in yardstick(precision):
        yd = ev(t, X)
        if not np.all(np.isfinite(yd)) or float(np.max(np.abs(yd - ref)) / scale) > tol / MARGIN[precision]:
            return False
    return True


def state_fits(trees, X, y, precision):
    """K-column SSE (ylogLike), intercept fit and its RMSE of a live state, in the reference's float64 arithmetic and in
    each yardstick arithmetic of `precision`: dict(sse, beta, rmse), each a list [float64 value, yardstick values ...]."""
    out = dict(sse=[], beta=[], rmse=[])
    with np.errstate(all="ignore"):
        for ev in [O.eval_tree] + yardstick(precision):
            cols = [ev(t, X) for t in trees]
            out["sse"].append(O.sse_no_intercept(y, np.stack(cols, axis=1)))
            beta, fit = O.intercept_fit(cols, y)
            out["beta"].append(beta.ravel())
            out["rmse"].append(float(np.sqrt(np.sum(np.square(fit[:, 0] - y)) / len(y))))
    return out


def resolves(vals, tol, floor=0.0, margin=4.0):
    """True when every yardstick arithmetic reproduces the float64 value(s) within tol / margin (relative, or of `floor`)."""
    a = np.asarray(vals[0], dtype=float)
    if not np.all(np.isfinite(a)):
        return False
    for b in vals[1:]:
        b = np.asarray(b, dtype=float)
        if not np.all(np.isfinite(b)) or not np.all(np.abs(a - b) <= tol / margin * np.maximum(np.max(np.abs(a)), floor)):
            return False
    return True


def min_angle_sine(cols):
    """sine of the smallest angle between a column and the span of the others = smallest singular value of the unit-norm columns
    (up to a factor <= sqrt(K))."""
    with np.errstate(all="ignore"):
        nrm = np.sqrt(np.sum(cols * cols, axis=0))
        if not np.all(np.isfinite(nrm)) or np.any(nrm == 0):
            return 0.0
        sv = np.linalg.svd(cols / nrm, compute_uv=False)
    return float(sv[-1])


class StepVerdict:
    """How one proposal of a device run compares with the oracle (judge_step)."""
    __slots__ = ("cls", "err", "hard", "soft", "tr", "acc", "newt", "sigma", "sa", "sb", "scale")


def judge_step(trees, k, sigma, sa_k, sb_k, y, X, cfg, tape, gpu, precision, logr_rel):
    """Replays one proposal through the oracle (float64, the reference's arithmetic) AND through the yardstick arithmetics of
    `precision`, then classifies what the device did.  gpu: dict(rank_reject, accepted, logR).

    cls is exactly one of
      rank_both      both reject on rank / a non-finite column (no logR exists)
      compared       logR held to logr_rel * max(1, |logR|, |ll_new|, |ll_old|): err is the normalised error
      type_limited   the device's evaluation type does not resolve this proposal, whoever computes in it; nothing but the
                     bookkeeping is asserted.  One of: a yardstick arithmetic moves logR by more than logr_rel / MARGIN,
                     changes the rank verdict or misses a column involved by more than COL_CHAOS; a column exceeds RANGE_LIMIT; the oracle finds full rank with a
                     column closer than RANK_ANGLE to the span of the others (the Gram-based rank test cannot tell that
                     from the rounding of two evaluations of one function, which numpy -- in float64 -- calls collinear)
      nonfinite      logR is NaN / inf on both sides (Q14: NaN accepts)
    hard: list of failures (rank, logR, decision, nonfinite) -- never tolerated.
    soft: list of deviations that follow from the class (rank / decision on a type_limited proposal, or a decision with
          log u within the tolerance of the threshold)."""
    v = StepVerdict()
    tape64 = list(tape)

    def run(ev, store=None):
        fn = ev
        if ev is not None and store is not None:
            def fn(t, Xa):
                c = ev(t, Xa)
                store.append(c)
                return c
        try:
            r = O.new_prop(trees, k, sigma, y, X, cfg, sa_k, sb_k, O.TapeDraws(tape64), eval_fn=fn)
            return r, False
        except IndexError:       # the device drew no accept uniform (it rejected on rank) where the oracle wants one
            if store is not None:
                del store[:]
            return O.new_prop(trees, k, sigma, y, X, cfg, sa_k, sb_k, O.TapeDraws(tape64 + [0.5]), eval_fn=fn), True

    (acc, sig2, newt, sa2, sb2, tr), u_missing = run(None)
    v.tr, v.acc, v.newt, v.sigma, v.sa, v.sb = tr, acc, newt, sig2, sa2, sb2
    v.hard, v.soft, v.err = [], [], 0.0
    scale = max(1.0, abs(tr.logR) if np.isfinite(tr.logR) else 1.0, abs(tr.yll_new) if np.isfinite(tr.yll_new) else 1.0,
                abs(tr.yll_old) if np.isfinite(tr.yll_old) else 1.0)
    v.scale = scale
    K = len(trees)
    with np.errstate(all="ignore"):
        cols = np.stack([O.eval_tree(tr.proposed if i == k else trees[i], X) for i in range(K)] + [O.eval_tree(trees[k], X)], axis=1)
    # the columns in the order new_prop evaluates them (funcs.py:1212-1224): slot k gives the proposal then the old tree
    order = []
    for i in range(K):
        order += [i, K] if i == k else [i]
    limited = bool(np.all(np.isfinite(cols)) and np.max(np.abs(cols)) > RANGE_LIMIT)
    if not limited and not tr.rank_deficient and min_angle_sine(cols[:, :K]) < RANK_ANGLE:
        limited = True
    if not limited:
        for ev in yardstick(precision):
            ycols = []
            try:
                ty = run(ev, ycols)[0][5]
            except np.linalg.LinAlgError:
                limited = True
                break
            with np.errstate(all="ignore"):
                for j, yc in zip(order, ycols):
                    ref = cols[:, j]
                    if np.all(np.isfinite(ref)) and (not np.all(np.isfinite(yc)) or
                                                     np.max(np.abs(yc - ref)) > COL_CHAOS * (np.max(np.abs(ref)) + 1e-300)):
                        limited = True
            if limited:
                break
            if tr.rank_deficient != ty.rank_deficient:
                limited = True
            elif not tr.rank_deficient:
                f64, fy = np.isfinite(tr.logR), np.isfinite(ty.logR)
                if f64 != fy or (f64 and abs(tr.logR - ty.logR) / scale > logr_rel / MARGIN[precision]):
                    limited = True
            if limited:
                break
    if limited:
        v.cls = "type_limited"
        if gpu["rank_reject"] != tr.rank_deficient:
            v.soft.append("rank")
        if u_missing or gpu["accepted"] != acc:
            v.soft.append("decision")
        return v
    if tr.rank_deficient:
        v.cls = "rank_both"
        if not gpu["rank_reject"]:
            v.hard.append("rank(gpu full rank, oracle reject)")
        return v
    if gpu["rank_reject"]:
        v.cls = "compared"
        v.hard.append("rank(gpu reject, oracle full rank)")
        return v
    if not np.isfinite(tr.logR):
        v.cls = "nonfinite"
        if np.isfinite(gpu["logR"]):
            v.hard.append("nonfinite")
    else:
        v.cls = "compared"
        if not np.isfinite(gpu["logR"]):
            v.hard.append("nonfinite")
        else:
            v.err = abs(tr.logR - gpu["logR"]) / scale
            if v.err > logr_rel:
                v.hard.append("logR")
    if gpu["accepted"] != acc:
        # a uniform within the tolerance of the threshold may fall on either side
        if np.isfinite(tr.logR) and abs(tr.log_u - min(tr.logR, 0.0)) <= logr_rel * scale:
            v.soft.append("decision")
        else:
            v.hard.append("decision")
    return v


def default_engine(K, C, d, precision="fp32", val=0, plateau=False, beta=-1.0, weights=None, chain_offset=0, err_cap=512):
    from mcmc_symreg_b200 import capi
    ops = list(range(1, 11))
    w = weights if weights is not None else [0.1] * 10
    return capi.Engine(K, C, ops, w, beta=beta, val=val, plateau_rule=plateau, precision=precision, chain_offset=chain_offset,
                       err_cap=err_cap)


def _zero_stats():
    return dict(proposals=0, accepts=0, logr_compared=0, rank_both=0, type_limited=0, nonfinite=0,
                tree_mismatch=0, scalar_mismatch=0, logr_mismatch=0, decision_mismatch=0, rank_mismatch=0, nonfinite_mismatch=0,
                decision_soft=0, rank_soft=0, diverged_chains=0, diverged_soft=0, state_mismatch=0, counter_mismatch=0,
                max_logr_err=0.0, details=[],
                # histogram of log10(smallest pivot of the device's rank test) by the oracle's verdict: bins [-21 .. 0], bin i holds
                # 10^(i-21) <= pivot < 10^(i-20); bin 0 also holds everything below (zero / negative pivots)
                piv_oracle_deficient=[0] * 22, piv_oracle_full=[0] * 22)


def replay_chain_in_oracle(job):
    """Replay one chain of a recorded device run through the oracle, proposal by proposal (a plain function of picklable
    arguments so that a process pool can map it).  Every consumed proposal is classified by judge_step; bookkeeping (move,
    change, Q, Qinv, hratio, detjacob, node count, and the proposed tree itself when the run logged it) is compared for
    every proposal whatever its class.  ``*_mismatch`` are hard failures; ``*_soft`` are deviations on proposals the
    evaluation type does not resolve (or a uniform sitting on the threshold).  A chain whose decision deviates is not
    followed further (``diverged_chains``; ``diverged_soft`` of them for a soft reason)."""
    (X, y, K, beta, c, trees_enc, sigma, sa, sb, trace, rec, cnt, logr_rel, final_enc, final_sigma, counters, sweeps, detail,
     precision, plog) = job
    from mcmc_symreg_b200 import capi
    TR = capi.TR
    d = X.shape[1]
    cfg = O.Config(n_feature=d, beta=beta)
    steps = trace.shape[0]
    out = _zero_stats()
    trees = [dec_tree(*e) for e in trees_enc]
    sa, sb = list(sa), list(sb)
    diverged = False
    n_acc = 0

    def describe(kind, s, v, t):
        if not detail:
            return
        tr = v.tr
        dd = dict(kind=kind, cls=v.cls, chain=int(c), step=int(s), k=int(s % K), live=[O.express(x) for x in trees],
                  proposed=O.express(tr.proposed), move=float(t[TR["move"]]), m_new=float(t[TR["m_new"]]),
                  gpu_logR=float(t[TR["logR"]]), oracle_logR=float(tr.logR), scale=float(v.scale), err=float(v.err),
                  gpu_u=float(t[TR["u"]]), oracle_log_u=float(tr.log_u), gpu_rank_reject=bool(t[TR["rank_reject"]]),
                  oracle_rank_deficient=bool(tr.rank_deficient), gpu_sse_new=float(t[TR["sse_new"]]), gpu_sse_old=float(t[TR["sse_old"]]),
                  pivot_min=float(t[TR["pivot_min"]]), sv_ratio=float(t[TR["sv_ratio"]]), rank_path=float(t[TR["rank_path"]]),
                  wide=float(t[TR["wide"]]), sigma=float(sigma), new_sigma=float(t[TR["new_sigma"]]),
                  state=[enc_golden(x) for x in trees], proposed_enc=enc_golden(tr.proposed))
        with np.errstate(all="ignore"):
            cols = np.stack([O.eval_tree(tr.proposed if i == s % K else trees[i], X) for i in range(K)], axis=1)
            if np.all(np.isfinite(cols)):
                sv = np.linalg.svd(cols, compute_uv=False)
                dd["sv_ratio_numpy"] = float(sv[-1] / sv[0]) if sv[0] > 0 else 0.0
                dd["numpy_tol"] = float(max(cols.shape) * np.finfo(float).eps)
        out["details"].append(dd)

    for s in range(steps):
        k = s % K
        t = trace[s]
        flags = int(t[TR["flags"]])
        gpu = dict(rank_reject=bool(t[TR["rank_reject"]]), accepted=bool(t[TR["accepted"]]), logR=float(t[TR["logR"]]))
        tape = list(rec[s, :cnt[s]])
        if not gpu["rank_reject"] and not (flags & 1):
            tape.append(float(t[TR["u"]]))
        v = judge_step(trees, k, sigma, sa[k], sb[k], y, X, cfg, tape, gpu, precision, logr_rel)
        tr = v.tr
        out["proposals"] += 1
        if tr.change != int(t[TR["change"]]) or tr.move != int(t[TR["move"]]) or not close(tr.Q, t[TR["Q"]], 1e-9) \
                or not close(tr.Qinv, t[TR["Qinv"]], 1e-9) or len(tr.proposed) != int(t[TR["m_new"]]):
            out["scalar_mismatch"] += 1
        if tr.change != 0 and (not close(tr.hratio, t[TR["hratio"]], 1e-7, 1e-300) or not close(tr.detjacob, t[TR["detjacob"]], 1e-12)):
            out["scalar_mismatch"] += 1
        if plog is not None:
            gp = dec_tree(plog[0][s], plog[1][s], plog[2][s], plog[3][s])
            if not trees_equal(gp, tr.proposed, params_rel=1e-13):
                out["tree_mismatch"] += 1
        pm = float(t[TR["pivot_min"]])
        if int(t[TR["rank_path"]]) >= 2 and np.isfinite(pm):
            b = 0 if pm <= 1e-21 else min(21, max(0, int(np.floor(np.log10(pm))) + 21))
            out["piv_oracle_deficient" if tr.rank_deficient else "piv_oracle_full"][b] += 1
        if v.cls == "compared":
            out["logr_compared"] += 1
            out["max_logr_err"] = max(out["max_logr_err"], v.err)
        else:
            out[v.cls] += 1
        for h in v.hard:
            key = "rank_mismatch" if h.startswith("rank") else (h.lower() + "_mismatch" if h != "logR" else "logr_mismatch")
            out[key] += 1
            describe(h, s, v, t)
        for h in v.soft:
            out[h + "_soft"] += 1
            describe("soft " + h, s, v, t)
        # the chain can be followed as long as the device took the oracle's decision
        if gpu["accepted"] != v.acc or (gpu["rank_reject"] != tr.rank_deficient and not tr.rank_deficient):
            diverged = True
            out["diverged_soft"] += int(not v.hard)
            break
        sigma, sa[k], sb[k] = v.sigma, v.sa, v.sb
        if v.acc:
            n_acc += 1
            out["accepts"] += 1
            trees = list(trees)
            trees[k] = v.newt
    if diverged:
        out["diverged_chains"] += 1
        return out
    for k in range(K):
        gt = dec_tree(*final_enc[k])
        if not trees_equal(gt, trees[k], params_rel=1e-13):
            out["state_mismatch"] += 1
    if not close(sigma, float(final_sigma), 1e-12):
        out["state_mismatch"] += 1
    if counters is not None and (int(counters[0]) != steps or int(counters[1]) != n_acc or int(counters[7]) != sweeps):
        out["counter_mismatch"] += 1
    return out


def _merge(parts, detail):
    out = _zero_stats()
    for p in parts:
        for key, v in p.items():
            if key == "details":
                out["details"].extend(v)
            elif key.startswith("piv_"):
                out[key] = [a + b for a, b in zip(out[key], v)]
            elif key == "max_logr_err":
                out[key] = max(out[key], v)
            else:
                out[key] += v
    p = max(1, out["proposals"])
    out["compared_share"] = out["logr_compared"] / p
    out["type_limited_share"] = out["type_limited"] / p
    out["rank_both_share"] = out["rank_both"] / p
    if not detail:
        del out["details"]
    return out


def _replay_jobs(X, y, K, beta, tok0, pa0, pb0, nn0, st0, trace, rec, cnt, tokf, paf, pbf, nnf, stf, sweeps, logr_rel, detail, precision,
                 plog=None, check_counters=True):
    jobs = []
    for c in range(trace.shape[0]):
        jobs.append((X, y, K, beta, c, [(tok0[c, k].copy(), pa0[c, k].copy(), pb0[c, k].copy(), int(nn0[c, k])) for k in range(K)],
                     float(st0["sigma"][c]), list(st0["sa"][c]), list(st0["sb"][c]), trace[c], rec[c], cnt[c], logr_rel,
                     [(tokf[c, k].copy(), paf[c, k].copy(), pbf[c, k].copy(), int(nnf[c, k])) for k in range(K)],
                     float(stf["sigma"][c]), stf["counters"][c].copy() if check_counters else None, sweeps, detail, precision,
                     None if plog is None else tuple(a[c] for a in plog)))
    return jobs


def _run_jobs(jobs, procs, detail):
    if procs > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            parts = pool.map(replay_chain_in_oracle, jobs, chunksize=max(1, len(jobs) // (4 * procs)))
    else:
        parts = [replay_chain_in_oracle(j) for j in jobs]
    return _merge(parts, detail)


def replay_gpu_run_in_oracle(X, y, K, n_chains, sweeps, seed, precision="fp32", beta=-1.0, logr_rel=None, chain_offset=0, procs=1,
                             detail=False):
    """The proposal-by-proposal pipeline (phase API: bsr_sweep_propose / eval / resolve) with its own Philox stream, every
    drawn value recorded; each chain is then replayed through the oracle with those values (replay_chain_in_oracle)."""
    d = X.shape[1]
    eng = default_engine(K, n_chains, d, precision=precision, beta=beta, chain_offset=chain_offset)
    eng.set_data(X, y)
    eng.init_chains(seed)
    tok0, pa0, pb0, nn0 = eng.get_trees(current=True)
    st0 = eng.get_stats()
    steps = sweeps * K
    eng.record_draws(steps, 256)
    eng.set_tape(None, steps)           # Philox, but keep a trace
    ptok = np.zeros((n_chains, steps, MAX_NODES), dtype=np.uint32)
    ppa = np.zeros((n_chains, steps, MAX_NODES))
    ppb = np.zeros((n_chains, steps, MAX_NODES))
    pnn = np.zeros((n_chains, steps), dtype=np.int32)
    for s in range(sweeps):
        eng.sweep_propose()
        a, b, c_, n_ = eng.get_proposals()
        ptok[:, s * K:(s + 1) * K], ppa[:, s * K:(s + 1) * K], ppb[:, s * K:(s + 1) * K], pnn[:, s * K:(s + 1) * K] = a, b, c_, n_
        eng.sweep_eval()
        eng.sweep_resolve()
    trace = eng.get_trace(steps)
    rec, cnt = eng.get_recorded_draws()
    tokf, paf, pbf, nnf = eng.get_trees(current=True)
    stf = eng.get_stats()
    eng.close()
    if logr_rel is None:
        logr_rel = 1e-6 if precision == "fp64" else 1e-3
    jobs = _replay_jobs(X, y, K, beta, tok0, pa0, pb0, nn0, st0, trace, rec, cnt, tokf, paf, pbf, nnf, stf, sweeps, logr_rel, detail,
                        precision, plog=(ptok, ppa, ppb, pnn), check_counters=False)
    return _run_jobs(jobs, procs, detail)


def replay_window_run_in_oracle(X, y, K, n_chains, sweeps, seed, precision="fp32", beta=-1.0, logr_rel=None, window=32,
                                run_chunks=(None,), procs=1, detail=False):
    """Same for the production path: ``Engine.run`` (speculative windows) records the draws, a trace row and the proposed
    tree of every CONSUMED proposal.  run_chunks: sizes of the successive run() calls (None = all remaining sweeps)."""
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
    eng.trace_trees()
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
    plog = eng.get_trace_trees(steps)
    tokf, paf, pbf, nnf = eng.get_trees(current=True)
    stf = eng.get_stats()
    eng.close()
    if logr_rel is None:
        logr_rel = 1e-6 if precision == "fp64" else 1e-3
    jobs = _replay_jobs(X, y, K, beta, tok0, pa0, pb0, nn0, st0, trace, rec, cnt, tokf, paf, pbf, nnf, stf, sweeps, logr_rel, detail,
                        precision, plog=plog)
    return _run_jobs(jobs, procs, detail)
