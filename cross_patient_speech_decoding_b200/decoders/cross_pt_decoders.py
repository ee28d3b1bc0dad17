"""Drop-in for the reference's ``decoders/cross_pt_decoders.py``: sklearn-style estimators
that reduce / align every patient, pool the trials and fit an injected decoder; at test
time only the target patient's transform is applied.

Constructor signatures, fitted attributes (``tar_dr``, ``common_dim``, ``algns``,
``joint_dr``, ``aligner``) and the reference's quirks are kept: ``fit`` returns the fitted
*decoder* (cross_pt_decoders.py:58-59) and ``crossPtDecoder_mcca`` replaces its ``aligner``
class by the instance on the first fit (:416-417).  Inject this package's GPU classes
(``decomposition.PCA.PCA``, ``alignment.AlignCCA.AlignCCA``, ``alignment.AlignMCCA.AlignMCCA``,
``svm.LinearSVC``) to run every stage on the B200; for whole CV loops prefer
``cv_align_decode`` / ``engine.CVEngine``, which runs all folds in one batch.
"""
import numpy as np
from sklearn.base import BaseEstimator

from ..decomposition.PCA import PCA


def _flat_trials(x):
    return x.reshape(x.shape[0], -1)


def _pool(X_tar, y, X_cross, cross_pt_data, tar_in_train):
    """Stack target (optionally) and cross-patient trials / labels
    (cross_pt_decoders.py:158-163, 264-269)."""
    ys = [lab for _, lab, _ in cross_pt_data]
    if tar_in_train:
        return np.vstack([X_tar] + X_cross), np.hstack([y] + ys)
    return np.vstack(X_cross), np.hstack(ys)


class crossPtDecoder(BaseEstimator):
    def preprocess_train(self, X, y=None):
        pass

    def preprocess_test(self, X, y=None):
        pass

    def fit(self, X, y, **kwargs):
        X_p, y_p = self.preprocess_train(X, y, **kwargs)
        return self.decoder.fit(X_p, y_p)

    def predict(self, X):
        return self.decoder.predict(self.preprocess_test(X))

    def score(self, X, y, **kwargs):
        return self.decoder.score(self.preprocess_test(X), y, **kwargs)


class crossPtDecoder_sepDimRed(crossPtDecoder):
    """Separate PCA per patient, truncated to the smallest latent size (reference :89-180)."""

    def __init__(self, cross_pt_data, decoder, dim_red=PCA, n_comp=0.8, tar_in_train=True):
        self.cross_pt_data = cross_pt_data
        self.decoder = decoder
        self.dim_red = dim_red
        self.n_comp = n_comp
        self.tar_in_train = tar_in_train

    def _reduce_all(self, X):
        cross = [self.dim_red(n_components=self.n_comp).fit_transform(
            x.reshape(-1, x.shape[-1])) for x, _, _ in self.cross_pt_data]
        self.tar_dr = self.dim_red(n_components=self.n_comp)
        tar = self.tar_dr.fit_transform(X.reshape(-1, X.shape[-1]))
        return tar, cross

    def preprocess_train(self, X, y, **kwargs):
        tar, cross = self._reduce_all(X)
        self.common_dim = min([tar.shape[-1]] + [c.shape[-1] for c in cross])
        d = self.common_dim
        cross = [c[:, :d].reshape(x.shape[0], -1) for c, (x, _, _) in
                 zip(cross, self.cross_pt_data)]
        return _pool(tar[:, :d].reshape(X.shape[0], -1), y, cross, self.cross_pt_data,
                     self.tar_in_train)

    def preprocess_test(self, X):
        X_dr = self.tar_dr.transform(X.reshape(-1, X.shape[-1]))[:, :self.common_dim]
        return X_dr.reshape(X.shape[0], -1)


class crossPtDecoder_sepAlign(crossPtDecoder_sepDimRed):
    """Separate PCA per patient, then one aligner per cross patient maps it into the target's
    latent space (reference :183-285)."""

    def __init__(self, cross_pt_data, decoder, aligner, dim_red=PCA, n_comp=0.8,
                 tar_in_train=True):
        self.cross_pt_data = cross_pt_data
        self.decoder = decoder
        self.dim_red = dim_red
        self.n_comp = n_comp
        self.aligner = aligner
        self.tar_in_train = tar_in_train

    def preprocess_train(self, X, y, y_align=None):
        tar, cross = self._reduce_all(X)
        tar = tar.reshape(X.shape[0], -1, tar.shape[-1])
        cross = [c.reshape(x.shape[0], -1, c.shape[-1]) for c, (x, _, _) in
                 zip(cross, self.cross_pt_data)]
        if y_align is None:
            y_align = y
        self.algns = [self.aligner() for _ in self.cross_pt_data]
        aligned = []
        for algn, c, (_, _, ya) in zip(self.algns, cross, self.cross_pt_data):
            algn.fit(tar, c, y_align, ya)
            aligned.append(_flat_trials(algn.transform(c)))
        return _pool(_flat_trials(tar), y, aligned, self.cross_pt_data, self.tar_in_train)

    def preprocess_test(self, X):
        return self.tar_dr.transform(X.reshape(-1, X.shape[-1])).reshape(X.shape[0], -1)


class _jointBase(crossPtDecoder):
    def _joint(self, model, X, y, y_align):
        if y_align is None:
            y_align = y
        Xs = [X] + [x for x, _, _ in self.cross_pt_data]
        ys = [y_align] + [ya for _, _, ya in self.cross_pt_data]
        out = model.fit_transform(Xs, ys)
        return _pool(_flat_trials(out[0]), y, [_flat_trials(o) for o in out[1:]],
                     self.cross_pt_data, self.tar_in_train)


class crossPtDecoder_jointDimRed(_jointBase):
    """Joint dimensionality reduction of all patients (reference :288-364)."""

    def __init__(self, cross_pt_data, decoder, joint_dr_method, n_comp=0.8, tar_in_train=True):
        self.cross_pt_data = cross_pt_data
        self.decoder = decoder
        self.joint_dr_method = joint_dr_method
        self.n_comp = n_comp
        self.tar_in_train = tar_in_train

    def preprocess_train(self, X, y, y_align=None):
        self.joint_dr = self.joint_dr_method(n_components=self.n_comp)
        return self._joint(self.joint_dr, X, y, y_align)

    def preprocess_test(self, X):
        return _flat_trials(self.joint_dr.transform(X, idx=0))


class crossPtDecoder_mcca(_jointBase):
    """MCCA alignment of all patients into a shared space (reference :367-445)."""

    def __init__(self, cross_pt_data, decoder, aligner, n_comp=10, regs=0.5, pca_var=1,
                 tar_in_train=True):
        self.cross_pt_data = cross_pt_data
        self.decoder = decoder
        self.aligner = aligner
        self.n_comp = n_comp
        self.regs = regs
        self.pca_var = pca_var
        self.tar_in_train = tar_in_train

    def preprocess_train(self, X, y, y_align=None):
        # as in the reference the class is replaced by the instance (a second fit on the same
        # object therefore raises TypeError)
        self.aligner = self.aligner(n_components=self.n_comp, regs=self.regs,
                                    pca_var=self.pca_var)
        return self._joint(self.aligner, X, y, y_align)

    def preprocess_test(self, X):
        return _flat_trials(self.aligner.transform(X, idx=0))
