"""Host-side tree objects: decode the device's pre-order token arrays into ``Node``-shaped objects.

``BSR.roots_`` must be consumable by the reference's own helpers, which read ``.type, .left, .right,
.operator, .feature`` (a 1-element int array), ``.a, .b, .depth, .parent`` (codes/funcs.py:175-220,
269-277, 314-342).  The helpers below give the same results as the reference functions of the same
name on such trees; they are host utilities for API parity, not part of the sampling path.
"""
import numpy as np

MAX_NODES = 64
OP_LEAF, OP_INV, OP_LT, OP_NEG, OP_SIN, OP_COS, OP_EXP, OP_SQUARE, OP_CUBIC, OP_ADD, OP_MUL = range(11)
# operator names of the reference (codes/bsr_class.py:110); 'ln' is the linear transform a*x+b
OP_NAME = {OP_INV: "inv", OP_LT: "ln", OP_NEG: "neg", OP_SIN: "sin", OP_COS: "cos", OP_EXP: "exp",
           OP_SQUARE: "square", OP_CUBIC: "cubic", OP_ADD: "+", OP_MUL: "*"}
NAME_OP = {v: k for k, v in OP_NAME.items()}
DEFAULT_OPS = ["inv", "ln", "neg", "sin", "cos", "exp", "square", "cubic", "+", "*"]


def arity(op):
    return 0 if op == OP_LEAF else (2 if op >= OP_ADD else 1)


class Node:
    """Attribute-compatible with the reference's Node (codes/funcs.py:30-54)."""

    __slots__ = ("type", "order", "left", "right", "depth", "parent", "operator", "op_ind", "data", "feature", "a", "b")

    def __init__(self, depth=0):
        self.type = -1
        self.order = 0
        self.left = None
        self.right = None
        self.depth = depth
        self.parent = None
        self.operator = None
        self.op_ind = None
        self.data = None
        self.feature = None
        self.a = None
        self.b = None


def decode_tree(tok, pa, pb, n):
    """Pre-order token arrays -> Node tree (slot i becomes the node with ``order == i``)."""
    pos = [0]

    def build(depth, parent):
        i = pos[0]
        if i >= n:
            raise ValueError("malformed tree encoding")
        pos[0] += 1
        t = int(tok[i])
        op, oi, ft = t & 0xFF, (t >> 8) & 0xFF, t >> 16
        nd = Node(depth)
        nd.order = i
        nd.parent = parent
        ar = arity(op)
        nd.type = ar
        if ar == 0:
            nd.feature = np.array([ft])
            return nd
        nd.operator = OP_NAME[op]
        nd.op_ind = oi
        if op == OP_LT:
            nd.a = np.float64(pa[i])
            nd.b = np.float64(pb[i])
        nd.left = build(depth + 1, nd)
        if ar == 2:
            nd.right = build(depth + 1, nd)
        return nd

    root = build(0, None)
    if pos[0] != n:
        raise ValueError("malformed tree encoding: %d tokens used of %d" % (pos[0], n))
    return root


def encode_tree(root, ops=None):
    """Node tree -> (tok[MAX_NODES] uint32, pa, pb float64, n)."""
    tok = np.zeros(MAX_NODES, dtype=np.uint32)
    pa = np.zeros(MAX_NODES)
    pb = np.zeros(MAX_NODES)
    i = [0]

    def rec(nd):
        k = i[0]
        if k >= MAX_NODES:
            raise ValueError("tree exceeds the %d-node capacity" % MAX_NODES)
        i[0] += 1
        if nd.type == 0:
            tok[k] = int(np.asarray(nd.feature).ravel()[0]) << 16
            return
        op = NAME_OP[nd.operator]
        oi = nd.op_ind if nd.op_ind is not None else (ops.index(nd.operator) if ops else 0)
        tok[k] = op | (int(oi) << 8)
        if op == OP_LT:
            pa[k] = float(nd.a)
            pb[k] = float(nd.b)
        rec(nd.left)
        if nd.type == 2:
            rec(nd.right)

    rec(root)
    return tok, pa, pb, i[0]


def genList(node):
    """Pre-order node list, sets ``order`` (codes/funcs.py:127-142)."""
    out = []
    stack = [node]
    while stack:
        nd = stack.pop()
        out.append(nd)
        if nd.right is not None:
            stack.append(nd.right)
        if nd.left is not None:
            stack.append(nd.left)
    for i, nd in enumerate(out):
        nd.order = i
    return out


def getNum(node):
    """codes/funcs.py:269-277"""
    return len(genList(node))


def getHeight(node):
    """codes/funcs.py:255-263 (a terminal has height 0)"""
    if node.type == 0:
        return 0
    if node.type == 1:
        return getHeight(node.left) + 1
    return max(getHeight(node.left), getHeight(node.right)) + 1


def numLT(node):
    """codes/funcs.py:283-292"""
    return sum(1 for nd in genList(node) if nd.type == 1 and nd.operator == "ln")


def Express(node):
    """String form of a tree, byte-identical to the reference's (codes/funcs.py:314-342)."""
    if node.type == 0:
        return "x" + str(node.feature)
    if node.type == 1:
        s = Express(node.left)
        op = node.operator
        if op == "exp":
            return "exp(" + s + ")"
        if op == "ln":
            return str(round(node.a, 4)) + "*(" + s + ")+" + str(round(node.b, 4))
        if op == "inv":
            return "1/[" + s + "]"
        if op == "sin":
            return "sin(" + s + ")"
        if op == "cos":
            return "cos(" + s + ")"
        if op == "square":
            return "(" + s + ")^2"
        if op == "cubic":
            return "(" + s + ")^3"
        return "-(" + s + ")"
    if node.operator == "+":
        return Express(node.left) + "+" + Express(node.right)
    return "(" + Express(node.left) + ")*(" + Express(node.right) + ")"


def allcal(node, indata):
    """Host float64 evaluation of a decoded tree (codes/funcs.py:175-220); returns (n, 1).
    Convenience for inspecting fitted models; the sampler evaluates trees on the GPU."""
    X = np.asarray(indata.values if hasattr(indata, "values") else indata, dtype=np.float64)

    def rec(nd):
        if nd.type == 0:
            return np.array(X[:, int(np.asarray(nd.feature).ravel()[0])], dtype=np.float64)
        v = rec(nd.left)
        op = nd.operator
        with np.errstate(all="ignore"):
            if nd.type == 2:
                r = rec(nd.right)
                return v + r if op == "+" else v * r
            if op == "ln":
                return nd.a * v + nd.b
            if op == "exp":
                return np.where(v <= 200, np.exp(np.minimum(v, 200)), 1e10)
            if op == "inv":
                return np.where(v == 0, 0.0, 1.0 / np.where(v == 0, 1.0, v))
            if op == "neg":
                return -v
            if op == "sin":
                return np.sin(v)
            if op == "cos":
                return np.cos(v)
            if op == "square":
                return np.square(v)
            return np.power(v, 3)

    out = rec(node).reshape(-1, 1)
    node.data = out
    return out
