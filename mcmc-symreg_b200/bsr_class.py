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
        m = len(self.betas_) - last_ind
        return self._predict_chains([m], X)[0].reshape(-1, 1)

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
                    # 32 sweeps per stop-rule check; the RMSE-at-accept trace grows between checks (never truncated)
                    sweeps = eng.run_until_done(int(self.max_sweeps), check_every=32)
                res = parallel.collect(eng)
                res["sweeps"] = sweeps
            finally:
                eng.close()
        res = dist.gather_results(res, MM, K)
        self._set_results(res)
        n_open = int(MM - np.count_nonzero(self.done_))
        if n_open and not fixed:
            # the reference loops until its stop rule fires (bsr_class.py:174); a sweep budget that ends earlier is said out loud
            import warnings
            warnings.warn("BSR.fit: %d of %d restarts had not met their stop rule after max_sweeps = %d sweeps; roots_ / betas_ hold "
                          "their current states (see done_)" % (n_open, MM, int(self.max_sweeps)), RuntimeWarning)
        if self.disp:
            c = res["counters"]
            print("chains: %d  proposals: %d  accepts: %d  rank-rejects: %d" % (MM, c[:, 0].sum(), c[:, 1].sum(), c[:, 2].sum()))
        return

    # ---- beyond the reference: what a batch of restarts allows ----------------------
    def best_chain(self):
        """Index of the restart with the lowest training RMSE at its last accept (codes/simulations.py:127-128 picks restarts
        by their training error by hand).  Restarts that never accepted rank by the RMSE of their initial fit."""
        return int(np.nanargmin(self.final_rmse_))

    def predict_best(self, test_data):
        """predict() of the best restart instead of the last one (the reference reads the last: quirk Q17)."""
        return self._predict_chains([self.best_chain()], _as_matrix(test_data))[0].reshape(-1, 1)

    def predict_mean(self, test_data, chains=None, return_std=False):
        """Posterior-predictive mean over restarts: every restart's model is evaluated on the GPU (float64) and the predictions
        are averaged on the device; restarts whose prediction is not finite are left out.  ``chains``: indices to use
        (default: all).  Returns (n_test, 1) [and the per-row standard deviation across restarts]."""
        X = _as_matrix(test_data)
        idx = np.arange(len(self.betas_)) if chains is None else np.asarray(chains, dtype=np.int64)
        mean, std, used = self._predict_chains(idx, X, reduce=True)
        self.n_used_ = int(used)
        return (mean.reshape(-1, 1), std.reshape(-1, 1)) if return_std else mean.reshape(-1, 1)

    def chain_diagnostics(self, tail=0.5):
        """Across-restart convergence summary: the best restart, the spread of the final training RMSE, and the Gelman-Rubin
        statistic R-hat of log RMSE-at-accept over the last ``tail`` share of the accepts, on the restarts that have at least
        the median number of accepts (traces cut to that common length)."""
        lens = np.array([len(e) for e in self.train_err_])
        out = dict(best=self.best_chain(), final_rmse_median=float(np.nanmedian(self.final_rmse_)),
                   final_rmse_best=float(np.nanmin(self.final_rmse_)), accepts_median=float(np.median(lens)))
        L = int(np.median(lens))
        sel = [np.log(np.asarray(e[:L])) for e in self.train_err_ if len(e) >= L and L >= 4]
        if len(sel) >= 2:
            A = np.stack(sel)[:, int(L * (1 - tail)):]
            A = A[np.all(np.isfinite(A), axis=1)]
        if len(sel) >= 2 and A.shape[0] >= 2 and A.shape[1] >= 2:
            n = A.shape[1]
            W = float(np.mean(np.var(A, axis=1, ddof=1)))
            B = float(n * np.var(np.mean(A, axis=1), ddof=1))
            out["rhat"] = float(np.sqrt(((n - 1) / n * W + B / n) / W)) if W > 0 else float("inf")
            out["rhat_chains"], out["rhat_length"] = int(A.shape[0]), int(n)
        else:
            out["rhat"] = float("nan")
        return out

    # ---- helpers ------------------------------------------------------------------------------------
    def _device_index(self):
        if self.device is not None:
            return int(self.device)
        return parallel.default_device()

    def _predict_chains(self, idx, X, reduce=False):
        pk = self._packed_
        idx = np.asarray(idx, dtype=np.int64)
        K = pk.nn.shape[1]
        tok = np.zeros((len(idx), K, capi.MAX_NODES), dtype=np.uint32); pa = np.zeros((len(idx), K, capi.MAX_NODES)); pb = np.zeros_like(pa)
        nn = np.zeros((len(idx), K), dtype=np.int32)
        for j, m in enumerate(idx):
            for k in range(K):
                tok[j, k], pa[j, k], pb[j, k], nn[j, k] = pk.tree(int(m), k)
        beta = np.stack([np.asarray(self.betas_[int(m)]).ravel() for m in idx])
        return capi.predict_many(self._device_index(), tok, pa, pb, nn, beta, X, reduce=reduce)

    def _set_results(self, res):
        self._packed_ = capi.PackedTrees(res["nn"], res["ptok"], res["pab"])
        MM, K = res["nn"].shape
        self.roots_ = _LazyRoots(self._packed_)
        self.betas_ = [res["beta"][m].reshape(K + 1, 1).copy() for m in range(MM)]
        ne = res["nerr"]
        cap = res["err"].shape[1]
        # the newest min(nerr, capacity) entries; run_until_done grows the capacity, so nothing is lost under fit()
        self.train_err_ = [[float(v) for v in res["err"][m, :min(int(ne[m]), cap)]] for m in range(MM)]
        self.train_err_truncated_ = bool(np.any(ne > cap))
        self.counters_ = res["counters"]
        self.done_ = res["done"].astype(bool)
        self.n_sweeps_ = res.get("sweeps")
        # RMSE of every restart's final state: its last accept, or the initial fit (K-column SSE is not it: use the trace / nan)
        self.final_rmse_ = np.array([e[-1] if e else np.nan for e in self.train_err_])
        if np.all(np.isnan(self.final_rmse_)):
            self.final_rmse_ = np.sqrt(np.maximum(res["sse"], 0) / max(1, res.get("n_rows", 1)))

    @property
    def _enc_(self):
        """dense [MM][K][64] encodings (tok, pa, pb, nn) of roots_"""
        return self._packed_.dense()

    def __getstate__(self):
        return dict(self.__dict__)


class _LazyRoots:
    """``roots_``: a list (over restarts) of lists of K ``Node`` trees, decoded from the packed device result on first access
    (a fit of thousands of restarts does not pay for Python tree objects nobody looks at)."""

    def __init__(self, packed):
        self._pk, self._cache = packed, {}

    def __len__(self):
        return self._pk.nn.shape[0]

    def _get(self, m):
        if m not in self._cache:
            K = self._pk.nn.shape[1]
            self._cache[m] = [decode_tree(*self._pk.tree(m, k)) for k in range(K)]
        return self._cache[m]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._get(m) for m in range(*i.indices(len(self)))]
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError("list index out of range")
        return self._get(i)

    def __iter__(self):
        return (self._get(m) for m in range(len(self)))

    def __getstate__(self):
        return dict(_pk=self._pk, _cache={})


def _as_matrix(data):
    if hasattr(data, "values"):
        data = data.values
    X = np.asarray(data, dtype=np.float64)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    return np.ascontiguousarray(X)
