"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Float64 numpy solver for the optimum of liblinear's L2-regularised squared-hinge SVM with
regularised bias (sklearn ``LinearSVC(loss='squared_hinge', C, fit_intercept=True,
intercept_scaling=1)``; objective in sklearn/svm/src/liblinear/linear.cpp, class
``l2r_l2_svc_fun``).  Used to certify that the liblinear run the oracle relies on has
actually converged: on this path's unscaled PCA scores ``LinearSVC(dual=True)`` does not
converge even in 1e6 epochs, so the oracle decoder is ``LinearSVC(dual=False, tol=1e-10)``
(same objective, liblinear's own trust-region Newton) cross-checked against this solver.
"""
import numpy as np


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
