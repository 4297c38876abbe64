"""CPU oracle for the BSR sampling hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A from-scratch numpy / pure-Python restatement of the reference algorithm
(ying531/MCMC-SymReg, ``codes/funcs.py`` + the chain driver in ``codes/bsr_class.py``) on
*prefix-order token arrays* instead of the reference's pointer trees.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may import
this module; the product path (``mcmc-symreg_b200``) never does and has no CPU fallback.

Parity pinning: the reference ships no tests or golden vectors.  This file is
pinned against the UNMODIFIED reference run in the build container: ``tests/golden/gen_golden.py``
records value-level RNG tapes + results from the reference, ``tests/test_oracle_golden.py`` replays
them through this oracle (integer bookkeeping bit-exact, floats to 1e-9 rel).

Every function cites the reference lines it restates.  Reference quirks are kept on purpose
(SURVEY.md §8a quirk register Q1-Q21); do not "fix" them here.

Tree encoding (shared with the CUDA side): slot i of the arrays is the node with pre-order index i
(= ``genList`` order, funcs.py:127-142).
    op[i]  semantic opcode (OP_*), 0 = terminal
    oi[i]  index into cfg.ops the node was *created* with (reference ``op_ind``; goes stale after
           reassignOperator, funcs.py:812,829,879,900 -- quirk Q7)
    ft[i]  feature index for terminals
    a[i], b[i]  lt() parameters (reference operator name 'ln', funcs.py:180-181)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

# ----------------------------------------------------------------------------------------------
# opcodes (fixed semantic codes; cfg.ops selects/weights a subset)           funcs.py:175-212
# ----------------------------------------------------------------------------------------------
OP_LEAF, OP_INV, OP_LT, OP_NEG, OP_SIN, OP_COS, OP_EXP, OP_SQUARE, OP_CUBIC, OP_ADD, OP_MUL = range(11)
OP_NAMES = {OP_INV: "inv", OP_LT: "ln", OP_NEG: "neg", OP_SIN: "sin", OP_COS: "cos", OP_EXP: "exp",
            OP_SQUARE: "square", OP_CUBIC: "cubic", OP_ADD: "+", OP_MUL: "*"}
NAME_TO_OP = {v: k for k, v in OP_NAMES.items()}
ARITY = [0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2]

CH_NONE, CH_EXPANSION, CH_SHRINKAGE = 0, 1, 2
MOVE_STAY, MOVE_GROW, MOVE_PRUNE, MOVE_DETR, MOVE_TRANS, MOVE_ROP, MOVE_RFEAT = range(7)


@dataclass
class Config:
    """Hard-coded in the reference at bsr_class.py:110-112 (10 ops, uniform weights)."""
    n_feature: int
    ops: Sequence[int] = (OP_INV, OP_LT, OP_NEG, OP_SIN, OP_COS, OP_EXP, OP_SQUARE, OP_CUBIC, OP_ADD, OP_MUL)
    weights: Optional[Sequence[float]] = None
    beta: float = -1.0

    def __post_init__(self):
        if self.weights is None:
            self.weights = [1.0 / len(self.ops)] * len(self.ops)
        self.weights = [float(w) for w in self.weights]
        self.cdf = np.cumsum(np.asarray(self.weights, dtype=np.float64))

    def psplit(self, depth: int) -> float:
        # funcs.py:79   prob = 1 / np.power((1 + depth), -beta)
        return float(1.0 / np.power(float(1 + depth), -float(self.beta)))


class Tree:
    __slots__ = ("op", "oi", "ft", "a", "b")

    def __init__(self, op=None, oi=None, ft=None, a=None, b=None):
        self.op = list(op or [])
        self.oi = list(oi or [])
        self.ft = list(ft or [])
        self.a = list(a or [])
        self.b = list(b or [])

    def __len__(self):
        return len(self.op)

    def copy(self):
        return Tree(self.op, self.oi, self.ft, self.a, self.b)

    def slice(self, lo, hi):
        return Tree(self.op[lo:hi], self.oi[lo:hi], self.ft[lo:hi], self.a[lo:hi], self.b[lo:hi])

    def splice(self, lo, hi, sub: "Tree"):
        """Replace tokens [lo, hi) by the tokens of ``sub`` (returns a new Tree)."""
        return Tree(self.op[:lo] + sub.op + self.op[hi:], self.oi[:lo] + sub.oi + self.oi[hi:],
                    self.ft[:lo] + sub.ft + self.ft[hi:], self.a[:lo] + sub.a + self.a[hi:],
                    self.b[:lo] + sub.b + self.b[hi:])

    def append_tok(self, op, oi=0, ft=0, a=0.0, b=0.0):
        self.op.append(op); self.oi.append(oi); self.ft.append(ft); self.a.append(a); self.b.append(b)

    def key(self):
        return (tuple(self.op), tuple(self.oi), tuple(self.ft))


# ----------------------------------------------------------------------------------------------
# draw sources (value-level RNG tape, SURVEY.md §4.2 / §8a "RNG tape")
# ----------------------------------------------------------------------------------------------
class TapeDraws:
    """Replays a recorded tape; every reference-level draw call consumes exactly one value."""

    def __init__(self, tape, pos=0):
        self.tape = tape
        self.pos = pos

    def _next(self):
        v = self.tape[self.pos]
        self.pos += 1
        return v

    def uniform(self):
        return float(self._next())

    def randint(self, lo, hi):
        v = int(self._next())
        assert lo <= v < hi, ("tape desync: randint", lo, hi, v, self.pos)
        return v

    def choice(self, cfg):
        v = int(self._next())
        assert 0 <= v < len(cfg.ops), ("tape desync: choice", v, self.pos)
        return v

    def normal(self, loc, scale):
        return float(self._next())

    def invgamma(self, shape):
        return float(self._next())


class GeneratorDraws:
    """The oracle's own stream (numpy Generator); optionally records a tape for the CUDA replay."""

    def __init__(self, seed, record=False):
        self.rng = np.random.default_rng(seed)
        self.tape = [] if record else None

    def _rec(self, v):
        if self.tape is not None:
            self.tape.append(float(v))
        return v

    def uniform(self):
        return self._rec(float(self.rng.random()))

    def randint(self, lo, hi):
        return int(self._rec(int(self.rng.integers(lo, hi))))

    def choice(self, cfg):
        # np.random.choice(p=w): searchsorted(cdf, u, side='right')   (funcs.py:86)
        u = self.rng.random()
        return int(self._rec(min(int(np.searchsorted(cfg.cdf, u, side="right")), len(cfg.ops) - 1)))

    def normal(self, loc, scale):
        return self._rec(float(loc + scale * self.rng.standard_normal()))

    def invgamma(self, shape):
        return self._rec(float(1.0 / self.rng.gamma(shape)))


# ----------------------------------------------------------------------------------------------
# integer bookkeeping on prefix arrays                   funcs.py:127-142, 255-307
# ----------------------------------------------------------------------------------------------
def subtree_sizes(op) -> List[int]:
    m = len(op)
    sz = [1] * m
    for i in range(m - 1, -1, -1):
        ar = ARITY[op[i]]
        if ar == 1:
            sz[i] = 1 + sz[i + 1]
        elif ar == 2:
            l = sz[i + 1]
            sz[i] = 1 + l + sz[i + 1 + l]
    return sz


def depths(op, root_depth=0) -> List[int]:
    """upDepth (funcs.py:298-307): depth of every pre-order slot, root = root_depth."""
    m = len(op)
    sz = subtree_sizes(op)
    dp = [0] * m
    if m:
        dp[0] = root_depth
    for i in range(m):
        ar = ARITY[op[i]]
        if ar >= 1:
            dp[i + 1] = dp[i] + 1
        if ar == 2:
            dp[i + 1 + sz[i + 1]] = dp[i] + 1
    return dp


def get_num(t: Tree) -> int:           # getNum  funcs.py:269-277
    return len(t)


def get_height(t: Tree) -> int:        # getHeight funcs.py:255-263 (leaf = 0)
    return max(depths(t.op)) if len(t) else 0


def num_lt(op) -> int:                 # numLT  funcs.py:283-292
    return sum(1 for o in op if o == OP_LT)


def det_candidates(op) -> List[int]:
    """detcd (funcs.py:454-468): non-terminals, except a root all of whose children are terminal."""
    out = []
    sz = subtree_sizes(op)
    for i, o in enumerate(op):
        if o == OP_LEAF:
            continue
        if i == 0:
            if ARITY[o] == 1 and op[1] == OP_LEAF:
                continue
            if ARITY[o] == 2 and op[1] == OP_LEAF and op[1 + sz[1]] == OP_LEAF:
                continue
        out.append(i)
    return out


def express(t: Tree, i=0, _sz=None) -> str:
    """Express (funcs.py:314-342).  Features print as ``x[j]`` because the reference stores them as
    1-element arrays; lt prints ``round(a,4)*(...)+round(b,4)`` via numpy float64 rounding."""
    sz = _sz or subtree_sizes(t.op)
    o = t.op[i]
    if o == OP_LEAF:
        return "x[" + str(int(t.ft[i])) + "]"
    if ARITY[o] == 1:
        s = express(t, i + 1, sz)
        if o == OP_EXP:
            return "exp(" + s + ")"
        if o == OP_LT:
            return str(round(np.float64(t.a[i]), 4)) + "*(" + s + ")+" + str(round(np.float64(t.b[i]), 4))
        if o == OP_INV:
            return "1/[" + s + "]"
        if o == OP_SIN:
            return "sin(" + s + ")"
        if o == OP_COS:
            return "cos(" + s + ")"
        if o == OP_SQUARE:
            return "(" + s + ")^2"
        if o == OP_CUBIC:
            return "(" + s + ")^3"
        return "-(" + s + ")"
    l = express(t, i + 1, sz)
    r = express(t, i + 1 + sz[i + 1], sz)
    if o == OP_ADD:
        return l + "+" + r
    return "(" + l + ")*(" + r + ")"


# ----------------------------------------------------------------------------------------------
# prior sampler                                                        funcs.py:74-119
# ----------------------------------------------------------------------------------------------
def grow(depth: int, cfg: Config, sigma_a: float, sigma_b: float, dr, out: Optional[Tree] = None) -> Tree:
    """Sample a subtree whose root sits at ``depth``; tokens appended in pre-order.
    Draw order per node: depth>0: U, then terminal => RI,RI (second kept, Q6) / op => CH;
    depth 0: CH; 'ln' => N(a), N(b) before descending (funcs.py:104-107)."""
    if out is None:
        out = Tree()
    terminal = False
    oi = 0
    if depth > 0:
        prob = cfg.psplit(depth)
        test = dr.uniform()
        if test > prob:
            dr.randint(0, cfg.n_feature)          # funcs.py:83 (overwritten at :99)
            terminal = True
        else:
            oi = dr.choice(cfg)
    else:
        oi = dr.choice(cfg)
    if terminal:
        out.append_tok(OP_LEAF, 0, dr.randint(0, cfg.n_feature))
        return out
    op = cfg.ops[oi]
    if ARITY[op] == 1:
        a = b = 0.0
        if op == OP_LT:
            a = dr.normal(1.0, math.sqrt(sigma_a))
            b = dr.normal(0.0, math.sqrt(sigma_b))
        out.append_tok(op, oi, 0, a, b)
        grow(depth + 1, cfg, sigma_a, sigma_b, dr, out)
    else:
        out.append_tok(op, oi, 0)
        grow(depth + 1, cfg, sigma_a, sigma_b, dr, out)
        grow(depth + 1, cfg, sigma_a, sigma_b, dr, out)
    return out


# ----------------------------------------------------------------------------------------------
# structure prior                                                      funcs.py:349-398
# ----------------------------------------------------------------------------------------------
def f_struc(t: Tree, dp: Sequence[int], cfg: Config, sigma_a: float, sigma_b: float):
    """[log p(T,M), log p(Theta | T, sigma_a, sigma_b)] of the token span ``t`` whose slots carry
    the (possibly stale, see Prop/detransform) depths ``dp``."""
    ll = 0.0
    lp = 0.0
    for i in range(len(t)):
        o = t.op[i]
        d = dp[i]
        if o == OP_LEAF:
            ll += math.log(1.0 - cfg.psplit(d)) if cfg.psplit(d) < 1.0 else -math.inf
            ll -= math.log(cfg.n_feature)
        else:
            if d == 0:
                ll += math.log(cfg.weights[t.oi[i]])
            else:
                ll += math.log(1 + d) * cfg.beta + math.log(cfg.weights[t.oi[i]])
            if o == OP_LT:
                lp -= (t.a[i] - 1.0) ** 2 / (2.0 * sigma_a)
                lp -= t.b[i] ** 2 / (2.0 * sigma_b)
                lp -= 0.5 * math.log(2.0 * math.pi * sigma_a)
                lp -= 0.5 * math.log(2.0 * math.pi * sigma_b)
    return ll, lp


# ----------------------------------------------------------------------------------------------
# structural proposal                                                  funcs.py:406-923
# ----------------------------------------------------------------------------------------------
@dataclass
class Proposal:
    old: Tree
    new: Tree
    move: int
    change: int
    Q: float
    Qinv: float
    last_a: List[float]
    last_b: List[float]
    still_ln: List[bool]           # per old lt node: does lnPointers[i].operator still read 'ln'?


def prop(t: Tree, cfg: Config, sigma_a: float, sigma_b: float, dr) -> Proposal:
    op = t.op
    m = len(t)
    sz = subtree_sizes(op)
    dp = depths(op)
    term = [i for i in range(m) if op[i] == OP_LEAF]                    # funcs.py:431-438
    nterm = [i for i in range(m) if op[i] != OP_LEAF]
    lts = [i for i in range(m) if op[i] == OP_LT]                       # funcs.py:421-428,441-445
    L, T, Nt = len(lts), len(term), len(nterm)
    last_a = [t.a[i] for i in lts]
    last_b = [t.b[i] for i in lts]
    changed_ln = -1                       # old pre-order slot of an lt node whose operator field is overwritten
    detcd = det_candidates(op)
    D = len(detcd)
    w = cfg.weights
    d_feat = cfg.n_feature

    # funcs.py:475-480
    p_stay = 0.25 * L / (L + 3)
    p_grow = (1 - p_stay) * min(1, 4 / (Nt + 2)) / 3
    p_prune = (1 - p_stay) / 3 - p_grow
    p_detr = (1 - p_stay) * (1 / 3) * D / (3 + D)
    p_trans = (1 - p_stay) / 3 - p_detr
    p_rop = (1 - p_stay) / 6

    test = dr.uniform()                                                 # funcs.py:483
    change = CH_NONE
    Q = Qinv = 1.0
    new = t.copy()

    def counts(tr: Tree):
        Lp = num_lt(tr.op)
        Tp = sum(1 for o in tr.op if o == OP_LEAF)
        return Lp, Tp, len(tr) - Tp, len(tr)

    if test <= p_stay:                                                  # stay  funcs.py:490-500
        move = MOVE_STAY
        Q = Qinv = p_stay
        for i in lts:
            new.a[i] = dr.normal(1.0, math.sqrt(sigma_a))
            new.b[i] = dr.normal(1.0, math.sqrt(sigma_b))               # loc=1 for b: quirk Q5

    elif test <= p_stay + p_grow:                                       # grow  funcs.py:503-536
        move = MOVE_GROW
        pod = dr.randint(0, T)
        i = term[pod]
        sub = grow(dp[i], cfg, sigma_a, sigma_b, dr)
        new = t.splice(i, i + 1, sub)
        if sub.op[0] == OP_LEAF:
            Q = Qinv = 1.0
        else:
            fs = f_struc(sub, depths(sub.op, dp[i]), cfg, sigma_a, sigma_b)[0]
            Q = p_grow * math.exp(fs) / T
            Lp, Tp, Ntp, mp = counts(new)
            new_p = (1 - 0.25 * Lp / (Lp + 3)) * (1 - min(1, 4 / (Ntp + 2))) / 3
            Qinv = new_p / max(1, (mp - Tp - 1))
            if Lp > L:
                change = CH_EXPANSION

    elif test <= p_stay + p_grow + p_prune:                             # prune funcs.py:539-579
        move = MOVE_PRUNE
        pod = dr.randint(1, Nt)
        i = nterm[pod]
        sub = t.slice(i, i + sz[i])
        fs = f_struc(sub, dp[i:i + sz[i]], cfg, sigma_a, sigma_b)[0]
        if num_lt(sub.op) > 0:
            change = CH_SHRINKAGE
        if op[i] == OP_LT:
            changed_ln = i
        leaf = Tree()
        leaf.append_tok(OP_LEAF, 0, dr.randint(0, d_feat))
        new = t.splice(i, i + sz[i], leaf)
        Lp, Tp, Ntp, mp = counts(new)
        Q = p_prune / ((Nt - 1) * d_feat)
        pg = 1 - 0.25 * Lp / (Lp + 3) * 0.75 * min(1, 4 / (Ntp + 2))    # literal precedence: quirk Q8
        Qinv = pg * math.exp(fs) / Tp

    elif test <= p_stay + p_grow + p_prune + p_detr:                    # detransform funcs.py:582-673
        move = MOVE_DETR
        det_od = dr.randint(0, D)
        i = detcd[det_od]
        Q = p_detr / D
        cut = None            # (lo, hi) span of the discarded child in the OLD tree
        if ARITY[op[i]] == 1:
            keep = (i + 1, i + sz[i])
        else:
            l = (i + 1, i + 1 + sz[i + 1])
            r = (l[1], i + sz[i])
            if i == 0 and op[l[0]] == OP_LEAF:                          # funcs.py:597-599
                cut, keep = l, r
            elif i == 0 and op[r[0]] == OP_LEAF:                        # funcs.py:600-602
                cut, keep = r, l
            else:                                                       # funcs.py:603-611 / 623-640
                aa = dr.uniform()
                if aa <= 0.5:
                    cut, keep = r, l
                else:
                    cut, keep = l, r
                Q = Q / 2
        new = t.splice(i, i + sz[i], t.slice(*keep))
        Lp, Tp, Ntp, mp = counts(new)
        if Lp < L:
            change = CH_SHRINKAGE
        new_pstay = 0.25 * Lp / (Lp + 3)
        Dp = len(det_candidates(new.op))
        new_pdetr = (1 - new_pstay) * (1 / 3) * Dp / (Dp + 3)
        new_ptr = (1 - new_pstay) / 3 - new_pdetr
        Qinv = new_ptr * w[t.oi[i]] / mp
        if cut is not None:                                             # cut keeps OLD-tree depths
            fs = f_struc(t.slice(*cut), dp[cut[0]:cut[1]], cfg, sigma_a, sigma_b)[0]
            Qinv = Qinv * math.exp(fs)

    elif test <= p_stay + p_grow + p_prune + p_detr + p_trans:          # transform funcs.py:679-786
        move = MOVE_TRANS
        i = dr.randint(0, m)
        ins_oi = dr.choice(cfg)
        ins_op = cfg.ops[ins_oi]
        node = Tree()
        node.append_tok(ins_op, ins_oi, 0, 0.0, 0.0)                    # lt params not drawn here (:701-704)
        if ARITY[ins_op] == 1:
            if ins_op == OP_LT:
                change = CH_EXPANSION
            new = t.splice(i, i, node)
            Q = p_trans * w[ins_oi] / m
        else:
            right = grow(dp[i] + 1, cfg, sigma_a, sigma_b, dr)
            fs = f_struc(right, depths(right.op, dp[i] + 1), cfg, sigma_a, sigma_b)[0]
            new = t.splice(i + sz[i], i + sz[i], right).splice(i, i, node)
            Q = p_trans * w[ins_oi] * math.exp(fs) / m
        Lp, Tp, Ntp, mp = counts(new)
        if Lp > L:
            change = CH_EXPANSION
        new_pstay = 0.25 * Lp / (Lp + 3)
        Dp = len(det_candidates(new.op))
        new_pdetr = (1 - new_pstay) * (1 / 3) * Dp / (Dp + 3)
        Qinv = new_pdetr / Dp
        if ARITY[ins_op] == 2:
            nsz = subtree_sizes(new.op)
            if new.op[i + 1] != OP_LEAF and new.op[i + 1 + nsz[i + 1]] != OP_LEAF:
                Qinv = Qinv / 2

    elif test <= p_stay + p_grow + p_prune + p_detr + p_trans + p_rop:  # reassignOperator :791-903
        move = MOVE_ROP
        pod = dr.randint(0, Nt)
        i = nterm[pod]
        last_op = op[i]
        last_oi = t.oi[i]                                               # never refreshed: quirk Q7
        new_oi = dr.choice(cfg)
        new_op = cfg.ops[new_oi]
        if ARITY[last_op] == 1:
            if ARITY[new_op] == 1:                                      # u -> u  :810-825
                new.op[i] = new_op
                if last_op == OP_LT:
                    if new_op != OP_LT:
                        new.a[i] = new.b[i] = 0.0
                        change = CH_SHRINKAGE
                        changed_ln = i
                elif new_op == OP_LT:
                    change = CH_EXPANSION
                Q = w[new_oi]
                Qinv = w[last_oi]
            else:                                                       # u -> b  :827-860
                new.op[i] = new_op
                if last_op == OP_LT:
                    new.a[i] = new.b[i] = 0.0
                    changed_ln = i
                right = grow(dp[i] + 1, cfg, sigma_a, sigma_b, dr)
                fs = f_struc(right, depths(right.op, dp[i] + 1), cfg, sigma_a, sigma_b)[0]
                new = new.splice(i + sz[i], i + sz[i], right)
                Q = p_rop * math.exp(fs) * w[new_oi] / Nt
                Lp, Tp, Ntp, mp = counts(new)
                new_p0 = Lp / (4 * (Lp + 3))
                Qinv = 0.125 * (1 - new_p0) * w[last_oi] / (mp - Tp)
                if Lp > L:
                    change = CH_EXPANSION
                elif Lp < L:
                    change = CH_SHRINKAGE
        else:
            if ARITY[new_op] == 1:                                      # b -> u  :867-894
                lo = i + 1 + sz[i + 1]
                hi = i + sz[i]
                cutted = t.slice(lo, hi)
                p_lt = num_lt(cutted.op)
                if p_lt > 1:                                            # '>1': quirk Q12
                    change = CH_SHRINKAGE
                elif new_op == OP_LT and p_lt == 0:
                    change = CH_EXPANSION
                new.op[i] = new_op
                new = new.splice(lo, hi, Tree())
                Q = p_rop * w[new_oi] / Nt
                Lp, Tp, Ntp, mp = counts(new)
                new_p0 = Lp / (4 * (Lp + 3))
                fs = f_struc(cutted, dp[lo:hi], cfg, sigma_a, sigma_b)[0]
                # newTerm is created empty and never filled in this branch (funcs.py:887-894), so the
                # reference divides by the node count m', not by the non-terminal count: quirk Q22
                Qinv = 0.125 * (1 - new_p0) * math.exp(fs) * w[last_oi] / mp
            else:                                                       # b -> b  :898-903
                new.op[i] = new_op
                Q = w[new_oi]
                Qinv = w[last_oi]

    else:                                                               # reassignFeature :907-917
        move = MOVE_RFEAT
        pod = dr.randint(0, T)
        new.ft[term[pod]] = dr.randint(0, d_feat)
        Q = Qinv = 1.0

    still = [(i != changed_ln) for i in lts]
    return Proposal(t, new, move, change, Q, Qinv, last_a, last_b, still)


# ----------------------------------------------------------------------------------------------
# reversible-jump auxiliary step                                        funcs.py:935-1138
# ----------------------------------------------------------------------------------------------
def _log_ig_pdf(x, a):
    # np.log(invgamma.pdf(x, a)),  pdf = x^(-a-1) exp(-1/x) / Gamma(a)
    return -(a + 1.0) * math.log(x) - 1.0 / x - math.lgamma(a)


def _log_norm_pdf(x, loc, scale):
    z = (x - loc) / scale
    return -0.5 * z * z - math.log(scale) - 0.5 * math.log(2.0 * math.pi)


def _norm_pdf(x, loc, scale):
    z = (x - loc) / scale
    return math.exp(-0.5 * z * z) / (scale * math.sqrt(2.0 * math.pi))


def aux_prop(p: Proposal, sigma_a: float, sigma_b: float, dr):
    """Returns (hratio, detjacob, new_sa2, new_sb2); hratio/detjacob are None in the same-dimension
    branch.  Mutates p.new's lt parameters (assignment by pre-order position, funcs.py:1022-1024)."""
    new = p.new
    od = [i for i in range(len(new)) if new.op[i] == OP_LT]
    new_sa2 = dr.invgamma(1.0)                                          # funcs.py:945-946
    new_sb2 = dr.invgamma(1.0)
    old_sa2, old_sb2 = sigma_a, sigma_b
    L = len(p.last_a)

    if p.change == CH_SHRINKAGE:                                        # funcs.py:950-1026
        pa = [p.last_a[i] for i in range(L) if p.still_ln[i]]
        pb = [p.last_b[i] for i in range(L) if p.still_ln[i]]
        ca = [p.last_a[i] for i in range(L) if not p.still_ln[i]]
        cb = [p.last_b[i] for i in range(L) if not p.still_ln[i]]
        for i in range(len(od) - len(pa)):
            pa.append(ca[i]); pb.append(cb[i])
        n0 = len(pa)
        Ua, Ub = [], []
        for i in range(n0):
            Ua.append(dr.normal(0.0, math.sqrt(new_sa2)))
            Ub.append(dr.normal(0.0, math.sqrt(new_sb2)))
        Na = [pa[i] + Ua[i] for i in range(n0)]
        Nb = [pb[i] + Ub[i] for i in range(n0)]
        NUa = [pa[i] - Ua[i] for i in range(n0)] + list(p.last_a)      # all last_a appended: quirk Q10
        NUb = [pb[i] - Ub[i] for i in range(n0)] + list(p.last_b)
        logh = _log_ig_pdf(new_sa2, 1.0) + _log_ig_pdf(new_sb2, 1.0)
        loghstar = _log_ig_pdf(old_sa2, 1.0) + _log_ig_pdf(old_sb2, 1.0)
        for i in range(n0):
            logh += _log_norm_pdf(Ua[i], 0.0, math.sqrt(new_sa2))
            logh += _log_norm_pdf(Ub[i], 0.0, math.sqrt(new_sb2))
        for i in range(len(NUa)):
            loghstar += _log_norm_pdf(NUa[i], 0.0, math.sqrt(old_sa2))
            loghstar += _log_norm_pdf(NUb[i], 0.0, math.sqrt(old_sb2))
        hratio = _safe_exp(loghstar - logh)
        detjacob = float(2.0 ** (2 * n0))
        for k, i in enumerate(od):
            new.a[i] = Na[k]; new.b[i] = Nb[k]
        return hratio, detjacob, new_sa2, new_sb2

    if p.change == CH_EXPANSION:                                        # funcs.py:1030-1110
        new_sa2 = dr.invgamma(1.0)                                      # drawn a second time: quirk Q11
        new_sb2 = dr.invgamma(1.0)
        Ua, Ub = [], []
        for i in range(L):
            Ua.append(dr.normal(0.0, math.sqrt(new_sa2)))
            Ub.append(dr.normal(0.0, math.sqrt(new_sb2)))
        Na = [(p.last_a[i] + Ua[i]) / 2 for i in range(L)]
        Nb = [(p.last_b[i] + Ub[i]) / 2 for i in range(L)]
        NUa = [(p.last_a[i] - Ua[i]) / 2 for i in range(L)]
        NUb = [(p.last_b[i] - Ub[i]) / 2 for i in range(L)]
        nn = len(od) - L
        for i in range(nn):
            Na.append(dr.normal(1.0, math.sqrt(new_sa2)))
            Nb.append(dr.normal(0.0, math.sqrt(new_sb2)))
        logh = _log_ig_pdf(new_sa2, 1.0) + _log_ig_pdf(new_sb2, 1.0)
        loghstar = _log_ig_pdf(old_sa2, 1.0) + _log_ig_pdf(old_sb2, 1.0)
        for i in range(L, nn):                                          # pdf, not log-pdf: quirk Q9
            logh += _norm_pdf(Na[i], 1.0, math.sqrt(new_sa2))
            logh += _norm_pdf(Nb[i], 0.0, math.sqrt(new_sb2))
        for i in range(L):
            logh += _log_norm_pdf(Ua[i], 0.0, math.sqrt(new_sa2))
            logh += _log_norm_pdf(Ub[i], 0.0, math.sqrt(new_sb2))
        for i in range(L):
            loghstar += _log_norm_pdf(NUa[i], 0.0, math.sqrt(old_sa2))
            loghstar += _log_norm_pdf(NUb[i], 0.0, math.sqrt(old_sb2))
        hratio = _safe_exp(loghstar - logh)
        detjacob = 1.0 / float(2.0 ** (2 * L))
        for k, i in enumerate(od):
            new.a[i] = Na[k]; new.b[i] = Nb[k]
        return hratio, detjacob, new_sa2, new_sb2

    # same dimension  funcs.py:1113-1138 (sigma's are drawn again at :1127-1128 and those are kept)
    new_sa2 = dr.invgamma(1.0)
    new_sb2 = dr.invgamma(1.0)
    for i in od:
        new.a[i] = dr.normal(1.0, math.sqrt(new_sa2))
        new.b[i] = dr.normal(0.0, math.sqrt(new_sb2))
    return None, None, new_sa2, new_sb2


def _safe_exp(x):
    if x != x:
        return float("nan")
    if x > 709.0:
        return float("inf")
    return math.exp(x)


# ----------------------------------------------------------------------------------------------
# tree evaluation                                                      funcs.py:175-220
# ----------------------------------------------------------------------------------------------
def eval_tree(t: Tree, X: np.ndarray) -> np.ndarray:
    """allcal on all rows, float64.  X is (n, d).  Scan tokens right-to-left with a value stack."""
    st = []
    with np.errstate(all="ignore"):
        for i in range(len(t) - 1, -1, -1):
            o = t.op[i]
            if o == OP_LEAF:
                st.append(np.array(X[:, t.ft[i]], dtype=np.float64))
            elif o == OP_ADD:
                l = st.pop(); r = st.pop(); st.append(l + r)
            elif o == OP_MUL:
                l = st.pop(); r = st.pop(); st.append(l * r)
            else:
                v = st.pop()
                if o == OP_LT:
                    v = t.a[i] * v + t.b[i]
                elif o == OP_EXP:                       # funcs.py:184-188 (NaN -> 1e10 too)
                    v = np.where(v <= 200, np.exp(np.minimum(v, 200)), 1e10)
                elif o == OP_INV:                       # funcs.py:191-195
                    v = np.where(v == 0, 0.0, 1.0 / np.where(v == 0, 1.0, v))
                elif o == OP_NEG:
                    v = -1 * v
                elif o == OP_SIN:
                    v = np.sin(v)
                elif o == OP_COS:
                    v = np.cos(v)
                elif o == OP_SQUARE:
                    v = np.square(v)
                elif o == OP_CUBIC:
                    v = np.power(v, 3)
                st.append(v)
    assert len(st) == 1
    return st[0]


def eval_tree_as(t: Tree, X: np.ndarray, dtype) -> np.ndarray:
    """allcal in another numpy arithmetic (float32 or longdouble), result widened / rounded to float64 -- NOT the
    reference's arithmetic: a yardstick for what an evaluation TYPE can resolve (tests/parity_helpers.py: a proposal whose
    logR moves by more than a quarter of the tolerance between eval_tree and this function in the type under test is
    limited by that type's rounding, whoever computes it -- float32 for the fp32 device path; for the fp64 device path
    the float64 result is held against the 80-bit one).  Inputs, lt parameters and every intermediate are rounded to
    `dtype`."""
    f = dtype
    Xt = np.asarray(X, dtype=f)
    st = []
    with np.errstate(all="ignore"):
        for i in range(len(t) - 1, -1, -1):
            o = t.op[i]
            if o == OP_LEAF:
                st.append(np.array(Xt[:, t.ft[i]], dtype=f))
            elif o == OP_ADD:
                l = st.pop(); r = st.pop(); st.append(l + r)
            elif o == OP_MUL:
                l = st.pop(); r = st.pop(); st.append(l * r)
            else:
                v = st.pop()
                if o == OP_LT:
                    v = f(t.a[i]) * v + f(t.b[i])
                elif o == OP_EXP:
                    v = np.where(v <= f(200), np.exp(np.minimum(v, f(200))), f(1e10))
                elif o == OP_INV:
                    v = np.where(v == 0, f(0), f(1) / np.where(v == 0, f(1), v))
                elif o == OP_NEG:
                    v = -v
                elif o == OP_SIN:
                    v = np.sin(v)
                elif o == OP_COS:
                    v = np.cos(v)
                elif o == OP_SQUARE:
                    v = v * v
                elif o == OP_CUBIC:
                    v = v * v * v
                st.append(np.asarray(v, dtype=f))
    return st[0].astype(np.float64)


def eval_tree_f32(t: Tree, X: np.ndarray) -> np.ndarray:
    """eval_tree_as(float32); a column that leaves the float32 range (exp beyond 88.7, powers of large values) is evaluated
    again in float64 from the float32-rounded inputs, as the device path re-interprets such columns in double range."""
    out = eval_tree_as(t, X, np.float32)
    if not np.all(np.isfinite(out)):
        return eval_tree(t, np.asarray(X, dtype=np.float32).astype(np.float64))
    return out


_SFU_ABS = 2.0 ** -21.41          # __sinf / __cosf on [-pi, pi]: maximum absolute error (CUDA C Programming Guide, intrinsic functions)
_F32_EPS = 2.0 ** -23
_F64_EPS = 2.0 ** -52


def eval_tree_sfu(t: Tree, X: np.ndarray, sign: int = 1) -> np.ndarray:
    """The yardstick of the fp32 device path: float32 arithmetic whose transcendentals have the ACCURACY OF THE GPU's special
    function unit rather than of a correctly rounded libm (north_star: "transcendentals go through the SFU with an fp32
    accuracy budget").  Each such result is moved by its error bound, with the sign given (+1 / -1: everywhere up / down;
    0: a fixed pseudo-random sign per value):
        sin, cos   absolute 2^-21.41 (MUFU.SIN / COS after the reduction to [-pi, pi]; documented bound, and measured:
                   scripts/sfu_accuracy.py) + |x| 2^-23 (the two-constant reduction k * 2 pi in float32); sin of a reduced
                   argument below 2^-5: relative 2^-23 (series, not MUFU: bsr_eval.cuh sin_reduced)
        exp        relative 2^-22 + |x| 2^-23 (ex2.approx of the float32 product x * log2 e)
        inv        relative 2^-23 (rcp.approx)
        lt         absolute |a x| 2^-24 (a * x + b is one fused multiply-add on the device, two roundings in numpy)
    everything else is correctly rounded float32.  The device's VALUE RULE is part of the type: a vector of
    four consecutive rows (row index a multiple of 4) with a non-finite float32 value is evaluated again in float64 from the
    float32-rounded inputs with the same bounds (sin / cos: absolute 2^-21.41 + |x| 2^-50, the float64 reduction), as the device
    re-interprets such vectors in double range with SFU transcendentals; every other vector keeps its float32 values -- also
    where float32 underflowed (1 / exp(-200) is 1 / 0 -> 0 by the reference's guard, not 7e86).
    tests/parity_helpers.py evaluates a proposal with sign = +1, -1, 0: if any of the three moves logR by more than a quarter
    of the tolerance (or changes the rank verdict, or misses a column grossly), errors the type PERMITS decide the comparison,
    and the proposal is counted as type-limited instead of compared -- e.g. a relative tolerance on 1/sin(.) near a zero of
    the sine."""
    b32 = dict(trig_abs=_SFU_ABS, trig_arg=_F32_EPS, sin_small=0.03125, sin_small_rel=_F32_EPS, trig_rel=0.0, exp_rel=2.0 ** -22,
               exp_arg=_F32_EPS, inv_rel=_F32_EPS, cubic_rel=0.0, lt_abs=2.0 ** -24)
    out = _eval_perturbed(t, np.asarray(X, dtype=np.float32), np.float32, sign, b32).astype(np.float64)
    bad = ~np.isfinite(out)
    if np.any(bad):
        wide = _eval_perturbed(t, np.asarray(X, dtype=np.float32).astype(np.float64), np.float64, sign, dict(b32, trig_arg=2.0 ** -50))
        n = out.shape[0]
        vec = np.zeros((n + 3) // 4 * 4, dtype=bool)
        vec[:n] = bad
        vec = np.repeat(vec.reshape(-1, 4).any(axis=1), 4)[:n]
        out = np.where(vec, wide, out)
    return out


def eval_tree_ulp(t: Tree, X: np.ndarray, sign: int = 1) -> np.ndarray:
    """The yardstick of the fp64 device path besides the 80-bit evaluation: float64 arithmetic with every result that the
    device may round differently from numpy moved by that difference, sign as in eval_tree_sfu: sin / cos / exp relative
    2 ulp (CUDA's libm: 1 - 2 ulp), cubic relative 1 ulp (x * x * x rounds twice, numpy's pow(x, 3) once), lt absolute
    |a x| 2^-53 (fused multiply-add).  sin(exp(x^3)) turns one ulp of x^3 = 27 into 1e-3 rad: such a proposal is not
    something float64 resolves to 1e-6, whoever computes it."""
    b64 = dict(trig_abs=0.0, trig_arg=0.0, sin_small=0.0, sin_small_rel=0.0, trig_rel=2 * _F64_EPS, exp_rel=2 * _F64_EPS, exp_arg=0.0,
               inv_rel=0.0, cubic_rel=_F64_EPS, lt_abs=2.0 ** -53)
    return _eval_perturbed(t, np.asarray(X, dtype=np.float64), np.float64, sign, b64)


def _eval_perturbed(t, Xt, f, sign, b):
    st = []
    rng = np.random.default_rng(12345)
    n = Xt.shape[0]

    def sg():
        if sign != 0:
            return f(sign)
        return (rng.integers(0, 2, n) * 2 - 1).astype(f)

    with np.errstate(all="ignore"):
        for i in range(len(t) - 1, -1, -1):
            o = t.op[i]
            if o == OP_LEAF:
                st.append(np.array(Xt[:, t.ft[i]], dtype=f))
            elif o == OP_ADD:
                l = st.pop(); r = st.pop(); st.append(l + r)
            elif o == OP_MUL:
                l = st.pop(); r = st.pop(); st.append(l * r)
            else:
                v = st.pop()
                if o == OP_LT:
                    av = f(t.a[i]) * v
                    v = av + f(t.b[i]) + sg() * np.where(np.isfinite(av), np.abs(av), f(0)) * f(b["lt_abs"])
                elif o == OP_EXP:
                    vm = np.minimum(v, f(200))
                    e = np.exp(vm) * (1 + sg() * (f(b["exp_rel"]) + np.abs(vm) * f(b["exp_arg"])))
                    v = np.where(v <= f(200), e, f(1e10))
                elif o == OP_INV:
                    v = np.where(v == 0, f(0), (f(1) / np.where(v == 0, f(1), v)) * (1 + sg() * f(b["inv_rel"])))
                elif o == OP_NEG:
                    v = -v
                elif o == OP_SIN or o == OP_COS:
                    w = np.sin(v) if o == OP_SIN else np.cos(v)
                    av = np.where(np.isfinite(v), np.abs(v), f(0))
                    bound = f(b["trig_abs"]) + av * f(b["trig_arg"]) + np.abs(w) * f(b["trig_rel"])
                    if o == OP_SIN and b["sin_small"] > 0:     # a small reduced argument goes through the series: relative accuracy
                        red = np.abs(v - f(2 * np.pi) * np.rint(v / f(2 * np.pi)))
                        bound = np.where(red < f(b["sin_small"]), np.abs(w) * f(b["sin_small_rel"]) + av * f(b["trig_arg"]), bound)
                    v = np.clip(w + sg() * bound, f(-1), f(1))
                elif o == OP_SQUARE:
                    v = v * v
                elif o == OP_CUBIC:
                    v = (v * v * v) * (1 + sg() * f(b["cubic_rel"]))
                st.append(np.asarray(v, dtype=f))
    return st[0]


def eval_tree_ld(t: Tree, X: np.ndarray) -> np.ndarray:
    return eval_tree_as(t, X, np.longdouble)


# ----------------------------------------------------------------------------------------------
# likelihood + OLS                                                     funcs.py:1147-1174
# ----------------------------------------------------------------------------------------------
def ridge_fit(outputs: np.ndarray, y: np.ndarray):
    """Scaled ridge OLS shared by ylogLike and the driver's intercept refit
    (funcs.py:1148-1157, bsr_class.py:155-163).  Returns (beta_scaled, scale, fitted)."""
    with np.errstate(all="ignore"):
        scale = np.max(np.abs(outputs))
        XX = outputs / scale
        eps = np.eye(XX.shape[1]) * 1e-6
        yy = np.asarray(y, dtype=np.float64).reshape(-1, 1)
        Bm = np.linalg.inv(np.matmul(XX.T, XX) + eps)
        beta = np.matmul(Bm, np.matmul(XX.T, yy))
        out = np.matmul(XX, beta)
    return beta, scale, out


def ylog_like(y: np.ndarray, outputs: np.ndarray, sigma: float) -> float:
    beta, scale, out = ridge_fit(outputs, y)
    with np.errstate(all="ignore"):
        error = float(np.sum(np.square(np.asarray(y, dtype=np.float64) - out[:, 0])))
        ll = -error / (2 * sigma * sigma)
        ll -= 0.5 * len(y) * math.log(2 * math.pi * sigma * sigma)
    return ll


def sse_no_intercept(y, outputs):
    beta, scale, out = ridge_fit(outputs, y)
    return float(np.sum(np.square(np.asarray(y, dtype=np.float64) - out[:, 0])))


# ----------------------------------------------------------------------------------------------
# one MH step                                                          funcs.py:1184-1306
# ----------------------------------------------------------------------------------------------
@dataclass
class StepTrace:
    move: int = -1
    change: int = 0
    Q: float = 1.0
    Qinv: float = 1.0
    hratio: float = float("nan")
    detjacob: float = float("nan")
    new_sigma: float = float("nan")
    new_sa2: float = float("nan")
    new_sb2: float = float("nan")
    rank_deficient: bool = False
    yll_new: float = float("nan")
    yll_old: float = float("nan")
    fs_new: float = float("nan")     # prior term of the proposed tree as it enters log_strucratio
    fs_old: float = float("nan")
    logR: float = float("nan")
    log_u: float = float("nan")
    accepted: bool = False
    proposed: Optional[Tree] = None
    n_draws: int = 0


def new_prop(trees: List[Tree], count: int, sigma: float, y: np.ndarray, X: np.ndarray, cfg: Config,
             sigma_a: float, sigma_b: float, dr, cols: Optional[List[np.ndarray]] = None, eval_fn=None):
    """Returns (accepted, sigma, tree, sigma_a, sigma_b, trace).  ``cols`` (optional) caches the
    current trees' outputs -- the reference re-evaluates them every call (funcs.py:1212-1224).
    ``eval_fn`` (tests only): evaluate trees with another arithmetic (eval_tree_f32) instead of eval_tree."""
    if eval_fn is not None:
        _ev = eval_fn
    else:
        _ev = eval_tree
    K = len(trees)
    tr = StepTrace()
    p = prop(trees[count], cfg, sigma_a, sigma_b, dr)
    new_sigma = dr.invgamma(4.0)                                        # funcs.py:1194-1195
    hratio, detjacob, new_sa2, new_sb2 = aux_prop(p, sigma_a, sigma_b, dr)
    tr.move, tr.change, tr.Q, tr.Qinv = p.move, p.change, p.Q, p.Qinv
    tr.new_sigma, tr.new_sa2, tr.new_sb2 = new_sigma, new_sa2, new_sb2
    if hratio is not None:
        tr.hratio, tr.detjacob = hratio, detjacob
    tr.proposed = p.new

    n = len(y)
    new_out = np.zeros((n, K))
    old_out = np.zeros((n, K))
    for i in range(K):
        if i == count:
            new_out[:, i] = _ev(p.new, X)
            old_out[:, i] = cols[i] if cols is not None else _ev(p.old, X)
        else:
            c = cols[i] if cols is not None else _ev(trees[i], X)
            new_out[:, i] = c
            old_out[:, i] = c

    if not np.all(np.isfinite(new_out)):
        # reference: inf -> matrix_rank == 0 -> reject; NaN -> LinAlgError aborts fit (Q15).
        # oracle + CUDA path: any non-finite proposed column => counted reject, no accept draw.
        tr.rank_deficient = True
        return False, sigma, p.old, sigma_a, sigma_b, tr
    if np.linalg.matrix_rank(new_out) < K:                              # funcs.py:1226-1228
        tr.rank_deficient = True
        return False, sigma, p.old, sigma_a, sigma_b, tr

    yll_new = ylog_like(y, new_out, new_sigma)
    yll_old = ylog_like(y, old_out, sigma)
    tr.yll_new, tr.yll_old = yll_new, yll_old
    log_yratio = yll_new - yll_old
    f_new = f_struc(p.new, depths(p.new.op), cfg, new_sa2, new_sb2)
    f_old = f_struc(p.old, depths(p.old.op), cfg, sigma_a, sigma_b)
    with np.errstate(all="ignore"):
        log_q = math.log(max(1e-5, p.Qinv / p.Q))                       # the 1e-5 floor is load-bearing (Q21)
        if p.change != CH_NONE:                                         # funcs.py:1241-1254 / 1265-1276
            sl = f_new[0] + f_new[1]
            slstar = f_old[0] + f_old[1]
            logR = log_yratio + (slstar - sl) + log_q + math.log(max(1e-5, hratio)) + math.log(max(1e-5, detjacob))
        else:                                                           # funcs.py:1287-1296
            sl, slstar = f_new[0], f_old[0]
            logR = log_yratio + (slstar - sl) + log_q
        logR = logR + _log_ig_pdf(new_sigma, 4.0) - _log_ig_pdf(sigma, 4.0)
    tr.fs_new, tr.fs_old = sl, slstar
    tr.logR = logR
    alpha = min(logR, 0)                                                # NaN logR -> NaN -> accept (Q14)
    test = dr.uniform()
    tr.log_u = math.log(test) if test > 0 else -math.inf
    if tr.log_u >= alpha:
        return False, sigma, p.old, sigma_a, sigma_b, tr
    tr.accepted = True
    return True, new_sigma, p.new, new_sa2, new_sb2, tr


# ----------------------------------------------------------------------------------------------
# chain driver (one restart of BSR.fit)                                bsr_class.py:99-273
# ----------------------------------------------------------------------------------------------
@dataclass
class ChainResult:
    trees: List[Tree]                 # what the reference appends to ROOTS (incl. quirk Q16)
    beta: np.ndarray                  # (K+1, 1), intercept first, un-scaled
    err_list: List[float]
    n_proposals: int = 0
    n_accepts: int = 0
    n_rank_rejects: int = 0
    node_evals_ref: int = 0           # reference-equivalent node-row evaluations
    traces: List[StepTrace] = field(default_factory=list)
    final_state: Optional[list] = None     # current trees at exit (differs from .trees only by Q16)
    sigma: float = float("nan")
    sigma_a: Optional[List[float]] = None
    sigma_b: Optional[List[float]] = None


def intercept_fit(cols: List[np.ndarray], y: np.ndarray):
    """(K+1)-coefficient refit with a ones column (bsr_class.py:147-163 / 211-227)."""
    n = len(y)
    XX = np.concatenate([np.ones((n, 1))] + [c.reshape(n, 1) for c in cols], axis=1)
    beta, scale, out = ridge_fit(XX, y)
    return beta / scale, out


def run_chain(X: np.ndarray, y: np.ndarray, K: int, cfg: Config, dr, val: int = 100,
              max_sweeps: Optional[int] = None, keep_traces: bool = False, fixed_sweeps: bool = False,
              init: Optional[dict] = None, on_step=None) -> ChainResult:
    """One restart.  ``fixed_sweeps``: benchmark mode -- run exactly ``max_sweeps`` sweeps, no stop rules."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = len(y)
    if init is None:
        sigma = dr.invgamma(1.0)                                        # bsr_class.py:123
        trees, sa, sb = [], [], []
        for _ in range(K):                                              # bsr_class.py:128-142
            s_a = dr.invgamma(1.0)
            s_b = dr.invgamma(1.0)
            trees.append(grow(0, cfg, s_a, s_b, dr))
            sa.append(s_a); sb.append(s_b)
    else:
        sigma = init["sigma"]; trees = [t.copy() for t in init["trees"]]
        sa = list(init["sigma_a"]); sb = list(init["sigma_b"])
    cols = [eval_tree(t, X) for t in trees]
    beta, _ = intercept_fit(cols, y)
    res = ChainResult(trees=trees, beta=beta, err_list=[])
    total = 0
    sweeps = 0
    roots_snapshot = list(trees)
    stop = False
    while (total < val) if not fixed_sweeps else True:
        if max_sweeps is not None and sweeps >= max_sweeps:
            break
        sweeps += 1
        for count in range(K):
            roots_snapshot = list(trees)                                # bsr_class.py:180-182
            m_others = sum(len(trees[i]) for i in range(K) if i != count)
            acc, sigma, newt, s_a, s_b, tr = new_prop(trees, count, sigma, y, X, cfg, sa[count], sb[count], dr, cols)
            if on_step is not None:
                on_step()
            res.n_proposals += 1
            res.node_evals_ref += n * (len(tr.proposed) + len(trees[count]) + m_others)
            res.n_rank_rejects += int(tr.rank_deficient)
            if keep_traces:
                res.traces.append(tr)
            total += 1
            sa[count] = s_a; sb[count] = s_b                            # bsr_class.py:197-198
            if acc:
                res.n_accepts += 1
                trees = list(trees)
                trees[count] = newt
                cols[count] = eval_tree(newt, X)
                beta, out = intercept_fit(cols, y)                      # bsr_class.py:211-227
                with np.errstate(all="ignore"):
                    rmse = float(np.sqrt(np.sum(np.square(out[:, 0] - y)) / n))   # :229-233
                res.err_list.append(rmse)
                total = 0
            if not fixed_sweeps:
                e = res.err_list                                        # bsr_class.py:248-252
                k10 = min(10, len(e))
                if len(e) > 100 and 1 - np.min(e[-k10:]) / np.mean(e[-k10:]) < 0.05:
                    stop = True
                    break
        if stop:
            break
    res.trees = roots_snapshot if stop else list(trees)                 # quirk Q16
    res.final_state = list(trees)
    res.beta = beta
    res.sigma, res.sigma_a, res.sigma_b = sigma, sa, sb
    return res


def predict(trees: List[Tree], beta: np.ndarray, X: np.ndarray) -> np.ndarray:
    """BSR.predict (bsr_class.py:53-68): returns (n, 1)."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    XX = np.concatenate([np.ones((n, 1))] + [eval_tree(t, X).reshape(n, 1) for t in trees], axis=1)
    return np.matmul(XX, beta)
