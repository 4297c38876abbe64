"""bsr-b200: B200-native Bayesian Symbolic Regression sampler (drop-in for ying531/MCMC-SymReg's ``BSR``).

Host side only: the estimator API of the reference (codes/bsr_class.py:26-278) over the C-ABI of
``libbsr_b200.so`` (include/bsr_b200.h).  All sampling runs in hand-written sm_100a CUDA; there is no
CPU fallback -- importing works anywhere, computing needs the built library and a CUDA device.
"""
from .trees import Node, Express, getNum, getHeight, numLT, genList, allcal, decode_tree, encode_tree  # noqa: F401
from .bsr_class import BSR  # noqa: F401
from . import capi  # noqa: F401

__all__ = ["BSR", "Node", "Express", "getNum", "getHeight", "numLT", "genList", "allcal", "capi"]
