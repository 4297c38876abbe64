"""``BSR`` estimator: the reference's public API (codes/bsr_class.py:26-278) over the B200 engine.

Same constructor arguments, ``fit / predict / model / complexity`` and fitted attributes
(``roots_``, ``betas_``, ``train_err_``) as the reference.  The reference runs its ``itrNum`` independent
restarts one after another (codes/bsr_class.py:99); here they are ``itrNum`` chains stepping in lock-step on
the GPU.  Extra keyword-only arguments expose what the reference hard-codes (operator set / weights,
codes/bsr_class.py:110-112) and what a batched sampler needs (seed, sweep budget, precision, device).
"""
import numpy as np

try:  # sklearn is what the reference derives from (codes/bsr_class.py:19,26); keep get_params/set_params/score
    from sklearn.base import BaseEstimator, RegressorMixin
except Exception:  # pragma: no cover
    class BaseEstimator(object):
        pass

    class RegressorMixin(object):
        pass

from . import capi, parallel
from .trees import DEFAULT_OPS, NAME_OP, Express, decode_tree, getNum


class BSR(BaseEstimator, RegressorMixin):
    def __init__(self, treeNum=3, itrNum=5000, alpha1=0.4, alpha2=0.4, beta=-1, disp=False, val=100, *,
                 seed=None, max_sweeps=100000, fixed_sweeps=None, ops=None, op_weights=None, precision="fp32",
                 device=None, err_cap=512, plateau_rule=True, distributed=None):
        self.treeNum = treeNum
        self.itrNum = itrNum
        self.alpha1 = alpha1      # stored and never used, as in the reference (bsr_class.py:94-95)
        self.alpha2 = alpha2
        self.beta = beta
        self.disp = disp
        self.val = val
        self.seed = seed
        self.max_sweeps = max_sweeps
        self.fixed_sweeps = fixed_sweeps
        self.ops = ops
        self.op_weights = op_weights
        self.precision = precision
        self.device = device
        self.err_cap = err_cap
        self.plateau_rule = plateau_rule
        self.distributed = distributed

    # ---- reference API ------------------------------------------------------------------------------
    def model(self, last_ind=1):
        """codes/bsr_class.py:37-41"""
        return [Express(self.roots_[-last_ind][i]) for i in range(self.treeNum)]

    def complexity(self):
        """codes/bsr_class.py:43-51 (total node count of the last restart)"""
        return int(sum(getNum(self.roots_[-1][i]) for i in range(self.treeNum)))

    def predict(self, test_data, method="last", last_ind=1):
        """codes/bsr_class.py:53-68: (n_test, 1) predictions of restart ``-last_ind`` (evaluated on the GPU)."""
        X = _as_matrix(test_data)
        if method != "last":
            # the reference only implements 'last' and dies on anything else (bsr_class.py:59,68)
            raise UnboundLocalError("predict: only method='last' is implemented (as in the reference)")
        m = len(self.roots_) - last_ind
        tok, pa, pb, nn = self._enc_
        out = capi.predict_trees(self._device_index(), tok[m], pa[m], pb[m], nn[m], np.asarray(self.betas_[m]).ravel(), X)
        return out.reshape(-1, 1)

    def fit(self, train_data, train_y):
        """codes/bsr_class.py:77-278.  ``itrNum`` chains; each stops after ``val`` consecutive rejections or on
        the RMSE plateau rule, exactly as one reference restart does."""
        X = _as_matrix(train_data)
        y = np.asarray(train_y, dtype=np.float64).ravel()
        if X.shape[0] != y.shape[0]:
            raise ValueError("train_data and train_y disagree on the number of rows")
        MM, K = int(self.itrNum), int(self.treeNum)
        ops = list(self.ops) if self.ops is not None else list(DEFAULT_OPS)
        opcodes = [NAME_OP[o] if isinstance(o, str) else int(o) for o in ops]
        w = list(self.op_weights) if self.op_weights is not None else [1.0 / len(ops)] * len(ops)
        seed = self.seed
        if seed is None:          # the reference draws from the global numpy state; so does the default seed
            seed = int(np.random.randint(0, 2 ** 31 - 1))
        dist = parallel.context(self.distributed)
        seed = dist.broadcast_int(seed)
        lo, hi = parallel.shard_range(MM, dist.rank, dist.world)
        fixed = self.fixed_sweeps is not None
        res = None
        if hi > lo:
            eng = capi.Engine(K, hi - lo, opcodes, w, beta=float(self.beta), val=(0 if fixed else int(self.val)),
                              plateau_rule=(self.plateau_rule and not fixed), precision=self.precision,
                              err_cap=int(self.err_cap), device=self._device_index(), chain_offset=lo)
            try:
                eng.set_data(X, y)
                eng.init_chains(seed)
                if self.disp:
                    print("starting training...")
                if fixed:
                    eng.run(int(self.fixed_sweeps))
                    sweeps = int(self.fixed_sweeps)
                else:
                    # 32 sweeps per stop-rule check: K * 32 proposals fill the 32-slot windows of bsr_run exactly
                    sweeps = eng.run_until_done(int(self.max_sweeps), check_every=32)
                res = parallel.collect(eng)
                res["sweeps"] = sweeps
            finally:
                eng.close()
        res = dist.gather_results(res, MM, K)
        self._set_results(res)
        if self.disp:
            c = res["counters"]
            print("chains: %d  proposals: %d  accepts: %d  rank-rejects: %d" % (MM, c[:, 0].sum(), c[:, 1].sum(), c[:, 2].sum()))
        return

    # ---- helpers ------------------------------------------------------------------------------------
    def _device_index(self):
        if self.device is not None:
            return int(self.device)
        return parallel.default_device()

    def _set_results(self, res):
        tok, pa, pb, nn = res["tok"], res["pa"], res["pb"], res["nn"]
        MM, K = nn.shape
        self._enc_ = (tok, pa, pb, nn)
        self.roots_ = [[decode_tree(tok[m, k], pa[m, k], pb[m, k], int(nn[m, k])) for k in range(K)] for m in range(MM)]
        self.betas_ = [res["beta"][m].reshape(K + 1, 1).copy() for m in range(MM)]
        ne = res["nerr"]
        cap = res["err"].shape[1]
        self.train_err_ = [[float(v) for v in res["err"][m, :min(int(ne[m]), cap)]] for m in range(MM)]
        self.counters_ = res["counters"]
        self.n_sweeps_ = res.get("sweeps")

    def __getstate__(self):
        return dict(self.__dict__)


def _as_matrix(data):
    if hasattr(data, "values"):
        data = data.values
    X = np.asarray(data, dtype=np.float64)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    return np.ascontiguousarray(X)
