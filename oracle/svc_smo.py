"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement (numpy, float64 with float32 kernel values) of the algorithm behind the
reference scripts' literal decoders -- ``sklearn.svm.SVC(kernel='rbf', class_weight='balanced')``
(scripts/aligned_decode_svm_ncv.py:313-317) and ``SVC(kernel='linear')``
(scripts/aligned_decode_svm.py:262-263).  The arithmetic lives in a third-party dependency:
libsvm as vendored by scikit-learn (pinned 1.6.1 in the reference's environment.yml:154; 1.9.0
installed here, same solver), whose published algorithm is restated here: C-SVC dual solved by
SMO with second-order working-set selection (Fan, Chen, Lin, JMLR 2005, "WSS 2"), the
``m(alpha) - M(alpha) < eps`` stopping rule, ``rho`` from the free variables, one-vs-one votes
with the first maximum winning, sklearn's ``gamma='scale'`` and ``class_weight='balanced'``.
No shrinking (it changes the iteration path, not the accepted optimum).
Pinned in tests/test_oracle.py against sklearn.svm.SVC itself; csrc/svc.cu follows the same
steps (one warp per class pair).
"""
import numpy as np

TAU = 1e-12


def kernel_matrix(X, Z, kernel, gamma):
    """libsvm's kernel values (computed in double, cached as float: ``Qfloat``)."""
    if kernel == 'linear':
        return X @ Z.T
    sx, sz = (X * X).sum(1), (Z * Z).sum(1)
    return np.exp(-gamma * (sx[:, None] + sz[None, :] - 2.0 * (X @ Z.T)))


def smo(K, y, Cp, Cn, eps=1e-3, max_iter=10000000):
    """Dual of one class pair.  K: (m, m) kernel values, y: +-1.  Returns alpha, rho, iterations."""
    m = len(y)
    K = K.astype(np.float32).astype(np.float64)
    Cb = np.where(y > 0, Cp, Cn)
    alpha = np.zeros(m)
    G = -np.ones(m)
    qd = np.diag(K).copy()
    it = 0
    while it < max_iter:
        up = ((y > 0) & (alpha < Cb)) | ((y < 0) & (alpha > 0))
        low = ((y > 0) & (alpha > 0)) | ((y < 0) & (alpha < Cb))
        if not up.any():
            break
        viol = np.where(up, -y * G, -np.inf)
        gmax = viol.max()
        i = int(np.nonzero(viol == gmax)[0][-1])            # libsvm scans upwards with >=
        gmax2 = np.where(low, y * G, -np.inf).max() if low.any() else -np.inf
        gd = gmax + y * G                                   # > 0 for violating pairs
        quad = qd[i] + qd - 2.0 * K[i]
        quad = np.where(quad > 0, quad, TAU)
        obj = np.where(low & (gd > 0), gd * gd / quad, -np.inf)
        if gmax + gmax2 < eps or not np.isfinite(obj.max()):
            break
        j = int(np.nonzero(obj == obj.max())[0][-1])
        it += 1
        Ci, Cj = Cb[i], Cb[j]
        ai, aj = alpha[i], alpha[j]
        Qij = y[i] * y[j] * K[i, j]
        if y[i] != y[j]:
            q = qd[i] + qd[j] + 2.0 * Qij
            q = q if q > 0 else TAU
            delta = (-G[i] - G[j]) / q
            diff = ai - aj
            ai, aj = ai + delta, aj + delta
            if diff > 0:
                if aj < 0:
                    aj, ai = 0.0, diff
            elif ai < 0:
                ai, aj = 0.0, -diff
            if diff > Ci - Cj:
                if ai > Ci:
                    ai, aj = Ci, Ci - diff
            elif aj > Cj:
                aj, ai = Cj, Cj + diff
        else:
            q = qd[i] + qd[j] - 2.0 * Qij
            q = q if q > 0 else TAU
            delta = (G[i] - G[j]) / q
            tot = ai + aj
            ai, aj = ai - delta, aj + delta
            if tot > Ci:
                if ai > Ci:
                    ai, aj = Ci, tot - Ci
            elif aj < 0:
                aj, ai = 0.0, tot
            if tot > Cj:
                if aj > Cj:
                    aj, ai = Cj, tot - Cj
            elif ai < 0:
                ai, aj = 0.0, tot
        dai, daj = ai - alpha[i], aj - alpha[j]
        G += (y * y[i] * K[i]) * dai + (y * y[j] * K[j]) * daj
        alpha[i], alpha[j] = ai, aj
    free = (alpha > 0) & (alpha < Cb)
    yG = y * G
    if free.any():
        rho = yG[free].mean()
    else:
        ub_set = ((alpha >= Cb) & (y < 0)) | ((alpha <= 0) & (y > 0))
        lb_set = ((alpha >= Cb) & (y > 0)) | ((alpha <= 0) & (y < 0))
        ub = yG[ub_set].min() if ub_set.any() else np.inf
        lb = yG[lb_set].max() if lb_set.any() else -np.inf
        rho = 0.5 * (ub + lb)
    return alpha, rho, it


def fit_ovo(X, y, C=1.0, kernel='rbf', gamma='scale', balanced=False, tol=1e-3):
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y)
    classes, counts = np.unique(y, return_counts=True)
    if gamma == 'scale':
        v = X.var()
        g = 1.0 / (X.shape[1] * v) if v > 0 else 1.0
    elif gamma == 'auto':
        g = 1.0 / X.shape[1]
    else:
        g = float(gamma)
    w = len(y) / (len(classes) * counts) if balanced else np.ones(len(classes))
    K = kernel_matrix(X, X, kernel, g)
    pairs = []
    for a in range(len(classes)):
        for b in range(a + 1, len(classes)):
            idx = np.concatenate([np.nonzero(y == classes[a])[0], np.nonzero(y == classes[b])[0]])
            ys = np.where(y[idx] == classes[a], 1.0, -1.0)
            alpha, rho, it = smo(K[np.ix_(idx, idx)], ys, C * w[a], C * w[b], eps=tol)
            pairs.append(dict(a=a, b=b, idx=idx, coef=alpha * ys, rho=rho, it=it))
    return dict(X=X, classes=classes, kernel=kernel, gamma=g, pairs=pairs)


def decision_ovo(model, Z):
    Kz = kernel_matrix(np.asarray(Z, dtype=np.float64), model['X'], model['kernel'], model['gamma'])
    return np.stack([Kz[:, p['idx']] @ p['coef'] - p['rho'] for p in model['pairs']], axis=1)


def predict_ovo(model, Z):
    dec = decision_ovo(model, Z)
    votes = np.zeros((dec.shape[0], len(model['classes'])), dtype=np.int64)
    for c, p in enumerate(model['pairs']):
        votes[:, p['a']] += dec[:, c] > 0
        votes[:, p['b']] += dec[:, c] <= 0
    return model['classes'][np.argmax(votes, axis=1)]                  # first maximum wins
