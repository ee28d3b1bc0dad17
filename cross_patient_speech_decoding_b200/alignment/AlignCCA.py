"""Drop-in for the reference's ``alignment/AlignCCA.py``: pairwise CCA alignment of two
patients' latent dynamics, computed by the CUDA kernels (class averaging, scatter Gram,
Cholesky-whitened one-sided Jacobi SVD) instead of numpy QR/SVD.

Same constructor, methods, fitted attributes (``M_a``, ``M_b``, ``canon_corrs``) and
exceptions as the reference class (AlignCCA.py:11-119).  CCA directions are unique only up
to sign (and rotation inside tied correlations); ``canon_corrs`` and the aligned subspaces
are what match the reference.
"""
import numpy as np

from .. import ops
from ..folds import label2str


class AlignCCA:
    def __init__(self, type='class', return_space='b_to_a'):
        self.type = type
        self.return_space = return_space

    # ------------------------------------------------------------------ fit
    def fit(self, X_a, X_b, y_a, y_b):
        L_a, L_b = reshape_latent_dynamics(X_a, X_b, y_a, y_b, type=self.type)
        M_a, M_b, S, G_ba, G_ab = _cca_from_latents(L_a, L_b)
        self.M_a = M_a
        self.M_b = M_b
        self.canon_corrs = S
        self._G_ba = G_ba            # M_b pinv(M_a)
        self._G_ab = G_ab            # M_a pinv(M_b)

    # ------------------------------------------------------------ transform
    def transform(self, X):
        if not self._check_fit():
            raise RuntimeError('Must call fit() before transforming data.')
        if self.return_space in ['b_to_a', 'a_to_b']:
            return self._transform_single(X)
        return self._transform_shared(X)

    def _transform_single(self, X):
        G = self._G_ba if self.return_space == 'b_to_a' else self._G_ab
        return ops.project(np.asarray(X), G).astype(np.float64)

    def _transform_shared(self, X):
        return (ops.project(np.asarray(X[0]), self.M_a).astype(np.float64),
                ops.project(np.asarray(X[1]), self.M_b).astype(np.float64))

    def _check_fit(self):
        return hasattr(self, 'M_a') and hasattr(self, 'M_b')


def _cca_from_latents(L_a, L_b):
    """CCA_align (reference AlignCCA.py:235-285) in Gram form on the GPU.
    L_a (n, d_a), L_b (n, d_b) sample-major latents."""
    L_a, L_b = np.asarray(L_a, dtype=np.float32), np.asarray(L_b, dtype=np.float32)
    da, db = L_a.shape[1], L_b.shape[1]
    L = np.ascontiguousarray(np.hstack([L_a, L_b]))
    mu = ops.colmean(L)
    S = ops.gram_tn(L, muA=mu, f64=True)             # centred scatter of [L_a | L_b], fp64
    out = ops.cca_solve(S[:da, :da], S[da:, da:], S[:da, da:])
    swp = ops.cca_solve(S[da:, da:], S[:da, :da], S[da:, :da])   # roles swapped: a -> b map
    return (out['Ma'].astype(np.float64), out['Mb'].astype(np.float64),
            out['rho'].astype(np.float64), out['G'].astype(np.float64),
            swp['G'].astype(np.float64))


def CCA_align(L_a, L_b):
    """Same signature as the reference function: inputs are (dims, samples) arrays; returns
    ``(M_a, M_b, S)``.  Unlike the reference it does not centre its inputs in place."""
    M_a, M_b, S, _, _ = _cca_from_latents(np.asarray(L_a).T, np.asarray(L_b).T)
    return M_a, M_b, S


def reshape_latent_dynamics(X_a, X_b, y_a, y_b, type='class'):
    if type == 'class':
        L_a, L_b = extract_latent_dynamics_by_class(X_a, X_b, y_a, y_b)
    elif type == 'trial':
        L_a, L_b = extract_latent_dynamics_by_trial_subselect(X_a, X_b, y_a, y_b)
    else:
        raise ValueError('type must be "class" or "trial".')
    return L_a.reshape(-1, L_a.shape[-1]), L_b.reshape(-1, L_b.shape[-1])


def extract_latent_dynamics_by_class(X_a, X_b, y_a, y_b):
    """Class averages over the classes both datasets contain (AlignCCA.py:156-183)."""
    y_a, y_b = label2str(np.asarray(y_a)), label2str(np.asarray(y_b))
    ca, inv_a = np.unique(y_a, return_inverse=True)
    cb, inv_b = np.unique(y_b, return_inverse=True)
    _, L_a = ops.class_mean(np.asarray(X_a), inv_a)
    _, L_b = ops.class_mean(np.asarray(X_b), inv_b)
    _, ia, ib = np.intersect1d(ca, cb, assume_unique=True, return_indices=True)
    return L_a[ia], L_b[ib]


def extract_latent_dynamics_by_trial_subselect(X_a, X_b, y_a, y_b):
    y_a, y_b = label2str(np.asarray(y_a)), label2str(np.asarray(y_b))
    return shared_trial_subselect(np.asarray(X_a), np.asarray(X_b), y_a, y_b)


def shared_trial_subselect(X_a, X_b, y_a, y_b):
    """Equal trial counts per shared class, trials drawn with ``np.random.permutation`` in
    the reference's call order (AlignCCA.py:205-232) so a seeded run selects the same trials."""
    L_a, L_b = [], []
    for c in np.intersect1d(y_a, y_b):
        cur_a = np.random.permutation(np.where(y_a == c)[0])
        cur_b = np.random.permutation(np.where(y_b == c)[0])
        m = min(cur_a.shape[0], cur_b.shape[0])
        L_a.append(X_a[cur_a[:m]])
        L_b.append(X_b[cur_b[:m]])
    return np.vstack(L_a), np.vstack(L_b)
