"""Host-side tree objects (mcmc-symreg_b200/trees.py) against values recorded from the UNMODIFIED reference
(tests/golden/fits_*.json.gz: roots_, model(), complexity(), predict()).

``BSR.roots_`` hands out Node-shaped objects decoded from the device's token arrays; ``model`` / ``complexity`` are the
reference's ``Express`` / ``getNum`` over them (codes/bsr_class.py:37-51, codes/funcs.py:269-277, 314-342).  The fixture
trees go through the same encoding the device uses (op | op_ind << 8 | feature << 16, pre-order) and back."""
import importlib

import numpy as np
import pytest

T = importlib.import_module("mcmc_symreg_b200.trees")

FIT_FILES = ["fits_f1_k3.json.gz", "fits_f6_k2.json.gz", "fits_plateau.json.gz", "fits_c1_readme.json.gz"]


def _slot(enc):
    n = len(enc["op"])
    tok = np.zeros(T.MAX_NODES, np.uint32)
    pa = np.zeros(T.MAX_NODES)
    pb = np.zeros(T.MAX_NODES)
    tok[:n] = [o | (oi << 8) | (ft << 16) for o, oi, ft in zip(enc["op"], enc["oi"], enc["ft"])]
    pa[:n] = enc["a"]
    pb[:n] = enc["b"]
    return tok, pa, pb, n


@pytest.mark.parametrize("fname", FIT_FILES)
def test_decoded_trees_give_the_reference_model_complexity_and_predictions(golden, fname):
    g = golden(fname)
    K = g["K"]
    X, Xt = np.array(g["X"]), np.array(g["Xtest"])
    last = [T.decode_tree(*_slot(enc)) for enc in g["roots"][-1]]
    first = [T.decode_tree(*_slot(enc)) for enc in g["roots"][0]]
    assert [T.Express(r) for r in last] == g["model"]
    assert [T.Express(r) for r in first] == g["model_first"]
    assert sum(T.getNum(r) for r in last) == g["complexity"]
    beta = np.array(g["betas"][-1]).reshape(-1, 1)
    for data, want in ((Xt, g["predict"]), (X, g["predict_train"])):
        cols = np.hstack([np.ones((len(data), 1))] + [T.allcal(r, data) for r in last])
        np.testing.assert_allclose((cols @ beta).ravel(), np.array(want), rtol=1e-9, atol=1e-9)
    assert len(last) == K


@pytest.mark.parametrize("fname", FIT_FILES)
def test_encode_decode_round_trip_and_node_fields(golden, fname):
    g = golden(fname)
    for restart in g["roots"]:
        for enc in restart:
            tok, pa, pb, n = _slot(enc)
            root = T.decode_tree(tok, pa, pb, n)
            nodes = T.genList(root)
            assert len(nodes) == n == T.getNum(root)
            assert [nd.order for nd in nodes] == list(range(n))                    # slot i is genList position i
            assert root.parent is None and root.depth == 0
            for nd, o, ft, a, b in zip(nodes, enc["op"], enc["ft"], enc["a"], enc["b"]):
                assert nd.type == T.arity(o)
                if o == T.OP_LEAF:
                    assert int(np.asarray(nd.feature).ravel()[0]) == ft and nd.left is None and nd.right is None
                else:
                    assert nd.operator == T.OP_NAME[o] and nd.left.parent is nd and nd.left.depth == nd.depth + 1
                    if o == T.OP_LT:
                        assert nd.a == a and nd.b == b
            assert T.numLT(root) == sum(1 for o in enc["op"] if o == T.OP_LT)
            tok2, pa2, pb2, n2 = T.encode_tree(root)
            assert n2 == n
            # op_ind is whatever the caller's operator list says; opcode, feature and lt parameters must survive
            assert [int(t) & 0xFF for t in tok2[:n]] == enc["op"]
            assert [int(t) >> 16 for t in tok2[:n]] == enc["ft"]
            lt = [i for i, o in enumerate(enc["op"]) if o == T.OP_LT]
            assert [pa2[i] for i in lt] == [enc["a"][i] for i in lt] and [pb2[i] for i in lt] == [enc["b"][i] for i in lt]
