"""One-vs-rest linear SVM estimator backed by the batched CUDA solver (cpsd_svm_fit_ovr):
the objective of ``sklearn.svm.LinearSVC(loss='squared_hinge', dual=True, C,
fit_intercept=True, intercept_scaling=1)`` -- the dual-coordinate-descent decoder the
north star names -- solved to its unique optimum (dual CD epochs + Newton polish).  Use it
as the last step of the decoder pipeline injected into the ``crossPtDecoder_*`` classes."""
import numpy as np
from sklearn.base import BaseEstimator, ClassifierMixin

from . import ops


class LinearSVC(BaseEstimator, ClassifierMixin):
    def __init__(self, C=1.0, tol=1e-4, dcd_epochs=2, max_newton=400, tol_newton=1e-9):
        self.C = C
        self.tol = tol
        self.dcd_epochs = dcd_epochs
        self.max_newton = max_newton
        self.tol_newton = tol_newton

    def fit(self, X, y):
        X = np.asarray(X)
        cls, W, info = ops.svm_fit_ovr(X.reshape(X.shape[0], -1), np.asarray(y), C=self.C,
                                       dcd_epochs=self.dcd_epochs, max_newton=self.max_newton,
                                       tol_newton=self.tol_newton, tol_dcd=self.tol)
        self.classes_ = cls
        self._W = W
        self.coef_ = W[:, :-1].copy()
        self.intercept_ = W[:, -1].copy()
        self.n_iter_ = int(info[:, 0].max()) if len(info) else 0
        self.converged_ = bool((info[:, 3] == 0).all())
        return self

    def decision_function(self, X):
        X = np.asarray(X)
        _, dec = ops.svm_predict_ovr(X.reshape(X.shape[0], -1), self.classes_, self._W,
                                     return_decision=True)
        return dec

    def predict(self, X):
        X = np.asarray(X)
        return ops.svm_predict_ovr(X.reshape(X.shape[0], -1), self.classes_, self._W)


class SVC(BaseEstimator, ClassifierMixin):
    """GPU stand-in for ``sklearn.svm.SVC`` as the reference scripts construct it:
    ``SVC(kernel='rbf', class_weight='balanced')`` (scripts/aligned_decode_svm_ncv.py:313-317,
    the decoder behind the paper's numbers) and ``SVC(kernel='linear')`` inside
    ``BaggingClassifier`` (scripts/aligned_decode_svm.py:262-263; sklearn's own
    ``BaggingClassifier`` clones and drives this class, so its bootstrap index streams are
    sklearn's).  libsvm's C-SVC (SMO, second-order working sets, ``tol`` stopping rule),
    one-vs-one votes with the first maximum winning; solved by ``cpsd_svc_fit_ovo``."""

    def __init__(self, C=1.0, kernel='rbf', gamma='scale', class_weight=None, tol=1e-3,
                 max_iter=-1, decision_function_shape='ovr'):
        self.C = C
        self.kernel = kernel
        self.gamma = gamma
        self.class_weight = class_weight
        self.tol = tol
        self.max_iter = max_iter
        self.decision_function_shape = decision_function_shape

    def fit(self, X, y):
        X = np.asarray(X)
        if self.class_weight not in (None, 'balanced'):
            raise ValueError("class_weight must be None or 'balanced'")
        cap = 10000000 if self.max_iter is None or self.max_iter < 0 else int(self.max_iter)
        m = ops.svc_fit_ovo(X.reshape(X.shape[0], -1), np.asarray(y), C=self.C, kernel=self.kernel,
                            gamma=self.gamma, balanced=self.class_weight == 'balanced', tol=self.tol,
                            max_iter=cap)
        self._model = m
        self.classes_ = m['classes']
        self._gamma = m['gamma']
        coef = m['coef']
        sv = np.nonzero(np.any(coef != 0.0, axis=0))[0]
        # libsvm groups the support vectors by class
        ys = np.asarray(y).astype(np.int32)[sv]
        order = np.argsort(np.searchsorted(self.classes_, ys), kind='stable')
        self.support_ = sv[order].astype(np.int32)
        # sklearn flips the sign of the binary problem so that positive means classes_[1]
        sgn = -1.0 if len(self.classes_) == 2 else 1.0
        self.dual_coef_ = sgn * coef[:, self.support_]
        self.n_support_ = np.array([(ys == c).sum() for c in self.classes_], dtype=np.int32)
        self.intercept_ = -sgn * m['rho']
        self.n_iter_ = m['info'][:, 0].copy()
        self.fit_status_ = int((m['info'][:, 1] == 1).any())
        if (m['info'][:, 1] == 3).any():
            raise RuntimeError('SVC: internal pair-size bound exceeded')
        return self

    def _ovo(self, X):
        X = np.asarray(X)
        return ops.svc_predict_ovo(self._model, X.reshape(X.shape[0], -1), return_decision=True)

    def predict(self, X):
        return self._ovo(X)[0]

    def decision_function(self, X):
        """Pair decisions (``decision_function_shape='ovo'``) or sklearn's monotone
        one-vs-rest transform of them (votes + scaled confidences)."""
        _, dec = self._ovo(X)
        ncls = len(self.classes_)
        if self.decision_function_shape == 'ovo' or ncls == 2:
            return -dec[:, 0] if ncls == 2 else dec
        votes = np.zeros((dec.shape[0], ncls))
        conf = np.zeros((dec.shape[0], ncls))
        p = 0
        for a in range(ncls):
            for b in range(a + 1, ncls):
                conf[:, a] += dec[:, p]
                conf[:, b] -= dec[:, p]
                votes[:, a] += dec[:, p] > 0
                votes[:, b] += dec[:, p] <= 0
                p += 1
        return votes + conf / (3 * (np.abs(conf) + 1))

    def __getstate__(self):
        st = dict(self.__dict__)
        m = st.pop('_model', None)
        if m is not None:            # device tensors -> host copies (joblib / pickle)
            st['_model_host'] = {k: (v.cpu().numpy() if hasattr(v, 'is_cuda') else v) for k, v in m.items()}
        return st

    def __setstate__(self, st):
        mh = st.pop('_model_host', None)
        self.__dict__.update(st)
        if mh is not None:
            from .device import Context
            ctx = Context.get(None)
            self._model = {k: (ctx.upload(v, v.dtype) if isinstance(v, np.ndarray) and k in (
                'St', 'y', 'classes_dev', 'coef_dev', 'rho_dev', 'gamma_dev') else v) for k, v in mh.items()}
