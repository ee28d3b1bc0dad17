"""GPU principal component analysis with scikit-learn's ``PCA`` semantics for the options
the reference uses (``n_components`` int / variance fraction / None, no whitening):
mean-centred, components signed by ``svd_flip(u_based_decision=False)``, float
``n_components`` resolved as ``searchsorted(cumsum(ratio), thr, side='right') + 1``
(sklearn/decomposition/_pca.py:603-666).  It is the ``dim_red`` class to inject into
``DimRedReshape`` and the ``crossPtDecoder_*`` classes.

Tall inputs (n >= features) go through the feature covariance (fp64-accumulated Gram +
shared-memory Jacobi); wide inputs through the sample Gram (tensor-core / SIMT Gram + block
Jacobi) -- the two routes of the engine's per-patient and pooled PCA stages.
"""
import numpy as np
from sklearn.base import BaseEstimator, TransformerMixin

from .. import ops


def _resolve_k(n_components, evals, n, kmax, total=None):
    if n_components is None:
        return kmax
    if isinstance(n_components, (float, np.floating)) and 0 < n_components < 1:
        return int(ops.select_k(evals[None], float(n_components), 0, n=[n], kmin=1, kmax=kmax,
                                total=None if total is None else [total])[0])
    return int(min(int(n_components), kmax))


def _eig_leading(A, n_components, kmax):
    """Eigen-pairs of a symmetric PSD matrix, leading ones first.  Large matrices go through
    the top-k subspace solver (ops.eig_topk) when the request provably fits its 128-wide block
    (converged Ritz residuals; variance threshold crossed inside the block); otherwise, and
    for small matrices, the full Jacobi solvers.  Returns (evals, evecs, total variance)."""
    n = A.shape[0]
    want_int = n_components is not None and not (isinstance(n_components, (float, np.floating))
                                                 and 0 < n_components < 1)
    if n > 256 and n_components is not None and (not want_int or int(n_components) <= 112):
        out = ops.eig_topk(A[None], m=128, iters=8, rounds=1)
        ev, V, tot = out['evals'][0], out['V'][0], float(out['total'][0])
        if want_int:
            k = min(int(n_components), kmax)
        else:
            k = int(ops.select_k(ev[None], float(n_components), 0, n=[128], kmin=1, kmax=128,
                                 total=[tot])[0])
        if not out['status'][0] and k <= 112 and out['resid'][0, :k].max() <= 5e-6 * max(ev[0], 1e-30):
            return ev, V, tot
    ev, V = ops.eig_sym(A, f64=n <= 128)
    return ev, V, None


class PCA(BaseEstimator, TransformerMixin):
    def __init__(self, n_components=None):
        self.n_components = n_components

    def fit(self, X, y=None):
        X = np.asarray(X)
        n, F = X.shape
        X32 = np.ascontiguousarray(X, dtype=np.float32)
        self.mean_ = ops.colmean(X32).astype(np.float64)
        kmax = min(n, F)
        if n >= F:
            cov = ops.gram_tn(X32, muA=self.mean_.astype(np.float32), alpha=1.0 / max(n - 1, 1),
                              f64=F <= 128)
            ev, V, tot = _eig_leading(cov, self.n_components, kmax)
            ev = np.maximum(ev.astype(np.float64), 0.0)
            k = _resolve_k(self.n_components, ev.astype(np.float32), len(ev), kmax, tot)
            comps = V[:, :k].T.astype(np.float64)
            var, total = ev, tot
        else:
            Xc = X32 - self.mean_.astype(np.float32)
            K = ops.gram_nt(Xc)
            ev, U, tot = _eig_leading(K, self.n_components, kmax)
            ev = np.maximum(ev.astype(np.float64), 0.0)
            k = _resolve_k(self.n_components, ev.astype(np.float32), len(ev), kmax, tot)
            sig = np.sqrt(ev[:k])
            # right singular vectors  V_k = Xc^T U_k / sigma_k   (features x k)
            comps = ops.project(np.ascontiguousarray(Xc.T), U[:, :k]).astype(np.float64)
            comps = (comps / np.where(sig > 0, sig, 1.0)).T
            var = ev / max(n - 1, 1)
            total = None if tot is None else tot / max(n - 1, 1)
        # svd_flip(u_based_decision=False): largest-|entry| of each component is positive
        idx = np.argmax(np.abs(comps), axis=1)
        sgn = np.sign(comps[np.arange(k), idx])
        sgn[sgn == 0] = 1.0
        self.components_ = comps * sgn[:, None]
        self.n_components_ = k
        self.n_samples_, self.n_features_in_ = n, F
        if total is None:
            total = var[:kmax].sum()
        self.explained_variance_ = var[:k]
        self.explained_variance_ratio_ = var[:k] / total if total > 0 else np.zeros(k)
        self.singular_values_ = np.sqrt(var[:k] * max(n - 1, 1))
        return self

    def transform(self, X):
        if not hasattr(self, 'components_'):
            raise ValueError('PCA must be fit before transforming data.')
        return ops.project(np.asarray(X), self.components_.T, self.mean_).astype(np.float64)

    def fit_transform(self, X, y=None):
        return self.fit(X).transform(X)
