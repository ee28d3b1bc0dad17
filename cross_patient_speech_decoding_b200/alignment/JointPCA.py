"""Drop-in for the reference's ``alignment/JointPCA.py``: joint-PCA alignment of several
patients to a shared latent space (Pandarinath et al. 2018 read-in matrices).

``fit`` condition-averages every patient over the classes they all share
(alignment_utils.extract_group_conditions), concatenates the averages along the channel
axis, reduces the ``(classes*time, sum_channels)`` matrix with ``dim_red(n_components)``
(JointPCA.py:194-199) and solves one least-squares problem per patient,
``W_p = pinv(X_p) @ latent`` (JointPCA.py:203-206); ``transform`` is ``X @ W_p`` without
centring (JointPCA.py:132,149).  Here the class averages, the PCA (covariance Gram +
eigen-solver) and the least squares (fp64 normal equations + shared-memory Cholesky) run in
the CUDA kernels of libcpsd_b200.so; ``dim_red`` defaults to this package's GPU ``PCA``.
"""
import numpy as np

from .. import ops
from ..decomposition.PCA import PCA
from .alignment_utils import extract_group_conditions


class JointPCA:
    def __init__(self, n_components=40, dim_red=PCA):
        self.n_components = n_components
        self.dim_red = dim_red

    def fit(self, X, y):
        self.transforms = get_joint_PCA_transforms(X, y, n_components=self.n_components,
                                                   dim_red=self.dim_red)

    def transform(self, X, idx=-1):
        if not self._check_fit():
            raise RuntimeError('Must call fit() before transforming data.')
        if idx == -1:
            return self._transform_multiple(X)
        if idx >= len(self.transforms):
            raise IndexError('Input idx is greater than the number of learned '
                             'transforms. For transformation of data from a '
                             'specific session, provide the input idx as the '
                             'index of the session in the input list. If '
                             'transforming multiple sessions, set idx=-1 '
                             '(default).')
        return self._transform_single(X, idx)

    def fit_transform(self, X, y):
        self.fit(X, y)
        return self.transform(X)

    def _transform_multiple(self, X):
        return tuple(self._transform_single(x, i) for i, x in enumerate(X[:len(self.transforms)]))

    def _transform_single(self, X, idx):
        X = np.asarray(X)
        out = ops.project(X.reshape(-1, X.shape[-1]), self.transforms[idx]).astype(np.float64)
        return out.reshape(X.shape[:-1] + (-1,))

    def _check_fit(self):
        return hasattr(self, 'transforms')


def get_joint_PCA_transforms(features, labels, n_components=40, dim_red=PCA):
    """Per-patient channel -> shared-latent read-in matrices ``(C_p, n_components)``."""
    cnd = extract_group_conditions(features, labels)
    views = [np.ascontiguousarray(c.reshape(-1, c.shape[-1])) for c in cnd]
    joint = np.concatenate(views, axis=-1)
    latent = dim_red(n_components=n_components).fit_transform(joint)
    out = []
    for v in views:
        W, status = ops.lstsq_gram(v, latent)
        if status:
            raise np.linalg.LinAlgError('JointPCA: a patient\'s condition averages are rank '
                                        'deficient (X_p^T X_p is not positive definite)')
        out.append(W)
    return tuple(out)
