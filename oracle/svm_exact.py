"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Float64 numpy solver for the optimum of liblinear's L2-regularised squared-hinge SVM with
regularised bias (sklearn ``LinearSVC(loss='squared_hinge', C, fit_intercept=True,
intercept_scaling=1)``; objective in sklearn/svm/src/liblinear/linear.cpp, class
``l2r_l2_svc_fun``).  Used to certify that the liblinear run the oracle relies on has
actually converged: on this path's unscaled PCA scores ``LinearSVC(dual=True)`` does not
converge even in 1e6 epochs, so the oracle decoder is ``LinearSVC(dual=False, tol=1e-10)``
(same objective, liblinear's own trust-region Newton) cross-checked against this solver.

liblinear's trust-region Newton can itself stall far from the optimum (measured on the noisy
CCA configuration, fold 3, class 5: 100 000 iterations at every tolerance from 1e-4 to 1e-10,
|gradient| 0.22, objective 0.0477 against the optimum's 0.0213 -- labels then depend on where
the solver happened to stop).  ``CertifiedLinearSVC`` is the oracle decoder: liblinear's primal
solver, every one-vs-rest problem certified by its gradient and re-solved by the finite Newton
method below when liblinear did not reach the (unique) optimum.
"""
import warnings

import numpy as np
from sklearn.svm import LinearSVC


def objective(w, X1, y, C):
    m = 1.0 - y * (X1 @ w)
    m[m < 0] = 0
    return 0.5 * w @ w + C * (m @ m)


def gradient(w, X1, y, C):
    z = y * (X1 @ w)
    a = z < 1
    return w - 2 * C * (X1[a].T @ (y[a] * (1 - z[a])))


def solve_binary(X, y, C=1.0, tol=1e-12, max_iter=200):
    """Finite Newton with exact line search.  X (n,k) without bias column, y in {-1,+1}."""
    X1 = np.hstack([X, np.ones((X.shape[0], 1))])
    k = X1.shape[1]
    w = np.zeros(k)
    g0 = np.abs(gradient(w, X1, y, C)).max()
    for _ in range(max_iter):
        z = y * (X1 @ w)
        a = z < 1
        g = w - 2 * C * (X1[a].T @ (y[a] * (1 - z[a])))
        if np.abs(g).max() <= tol * max(1.0, g0):
            break
        H = np.eye(k) + 2 * C * (X1[a].T @ X1[a])
        d = -np.linalg.solve(H, g)
        q = y * (X1 @ d)
        wd, dd = w @ d, d @ d

        def dphi(t):
            m = 1 - z - t * q
            s = m > 0
            return wd + t * dd - 2 * C * (m[s] @ q[s])
        lo, hi = 0.0, 1.0
        while dphi(hi) < 0:
            lo, hi = hi, 2 * hi
        for _ in range(200):
            mid = 0.5 * (lo + hi)
            if dphi(mid) < 0:
                lo = mid
            else:
                hi = mid
        w = w + 0.5 * (lo + hi) * d
    return w


def solve_ovr(X, y, C=1.0):
    """Returns (classes, W (n_classes, k+1)) -- bias in the last column."""
    classes = np.unique(y)
    W = np.array([solve_binary(X, np.where(y == c, 1.0, -1.0), C) for c in classes])
    return classes, W


def predict_ovr(classes, W, X):
    dec = X @ W[:, :-1].T + W[:, -1]
    return classes[np.argmax(dec, axis=1)], dec


class CertifiedLinearSVC(LinearSVC):
    """``LinearSVC(dual=False)`` whose fitted weights are certified per class: a one-vs-rest
    problem whose gradient at liblinear's answer exceeds ``1e-7 |gradient(0)|`` is re-solved
    with ``solve_binary``.  ``refit_`` lists the classes that needed it."""

    def fit(self, X, y, sample_weight=None):
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            super().fit(X, y, sample_weight)
        X = np.asarray(X, dtype=np.float64)
        X1 = np.hstack([X, np.ones((X.shape[0], 1))])
        binary = len(self.classes_) == 2
        self.refit_ = []
        coef, icpt = self.coef_.copy(), np.atleast_1d(self.intercept_).astype(np.float64).copy()
        for r in range(1 if binary else len(self.classes_)):
            pos = self.classes_[1] if binary else self.classes_[r]
            yy = np.where(np.asarray(y) == pos, 1.0, -1.0)
            w = np.r_[coef[r], icpt[r]]
            g0 = np.abs(gradient(np.zeros_like(w), X1, yy, self.C)).max()
            if np.abs(gradient(w, X1, yy, self.C)).max() > 1e-7 * max(1.0, g0):
                w = solve_binary(X, yy, self.C)
                coef[r], icpt[r] = w[:-1], w[-1]
                self.refit_.append(pos)
        self.coef_, self.intercept_ = coef, icpt
        return self


def oracle_linear_svc(C=1.0):
    """The pinned oracle decoder (see the module docstring)."""
    return CertifiedLinearSVC(dual=False, C=C, tol=1e-10, max_iter=100000)
