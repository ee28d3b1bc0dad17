"""Drop-in for the reference's ``decomposition/NoCenterPCA.py``: PCA without mean-centring
(thin SVD of the raw matrix).  Same quirks as the reference: ``components_`` is
``(n_features, k)``, ``explained_variance_`` holds ALL squared singular values, a float
``n_components`` keeps ``argmax(cum_var >= thr) + 1`` components (NoCenterPCA.py:86-103) and
``None`` / too-large values print a notice and keep ``min(X.shape)``."""
import numpy as np
from sklearn.base import BaseEstimator, TransformerMixin

from .. import ops


class NoCenterPCA(BaseEstimator, TransformerMixin):
    def __init__(self, n_components=None):
        self.n_components = n_components
        self._fit = False

    def fit(self, X, y=None):
        X = np.asarray(X)
        n, F = X.shape
        X32 = np.ascontiguousarray(X, dtype=np.float32)
        if n >= F:
            G = ops.gram_tn(X32, f64=F <= 128)
            ev, V = ops.eig_sym(G, f64=F <= 128)
            S2 = np.maximum(ev.astype(np.float64), 0.0)
            k = self._get_components(X, S2)
            comps = V[:, :k].astype(np.float64)
        else:
            K = ops.gram_nt(X32)
            ev, U = ops.eig_sym(K)
            S2 = np.maximum(ev.astype(np.float64), 0.0)
            k = self._get_components(X, S2)
            sig = np.sqrt(S2[:k])
            comps = ops.project(np.ascontiguousarray(X32.T), U[:, :k]).astype(np.float64)
            comps = comps / np.where(sig > 0, sig, 1.0)
        self.components_ = comps
        self.explained_variance_ = S2[:min(n, F)]
        self._fit = True
        return self

    def transform(self, X):
        self._check_fit()
        return ops.project(np.asarray(X), self.components_).astype(np.float64)

    def fit_transform(self, X, y=None):
        self.fit(X, y)
        return self.transform(X)

    def _get_components(self, X, S2):
        if self.n_components is None or self.n_components >= min(X.shape):
            print("n_components is None or greater than the number of features"
                  "/samples. Using n_components = min(X.shape)")
            return min(X.shape)
        if self.n_components < 1:
            n = min(X.shape)
            return int(ops.select_k(S2[None, :n].astype(np.float32), float(self.n_components), 2,
                                    n=[n], kmin=1, kmax=n)[0])
        return int(self.n_components)

    def _check_fit(self):
        if not self._fit:
            raise ValueError("PCA must be fit before transforming data.")
