"""One-vs-rest linear SVM estimator backed by the batched CUDA solver (cpsd_svm_fit_ovr):
the objective of ``sklearn.svm.LinearSVC(loss='squared_hinge', dual=True, C,
fit_intercept=True, intercept_scaling=1)`` -- the dual-coordinate-descent decoder the
north star names -- solved to its unique optimum (dual CD epochs + Newton polish).  Use it
as the last step of the decoder pipeline injected into the ``crossPtDecoder_*`` classes."""
import numpy as np
from sklearn.base import BaseEstimator, ClassifierMixin

from . import ops


class LinearSVC(BaseEstimator, ClassifierMixin):
    def __init__(self, C=1.0, tol=1e-4, dcd_epochs=2, max_newton=60, tol_newton=1e-9):
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
