"""Drop-in for the reference's ``alignment/AlignMCCA.py``: multi-view CCA alignment.

The reference delegates the arithmetic to ``mvlearn.embed.MCCA`` (AlignMCCA.py:152-153).
Here the whole fit -- condition averages, per-view signal ranks (``n_components_var``,
AlignMCCA.py:156-174), rank-r view bases, the regularised SUMCOR generalised eigenproblem
and the back-projection to channel loadings -- runs in the CUDA kernels; ``self.mcca``
exposes the three members of the mvlearn object that the reference touches
(``loadings_``, ``transform_view``, plus ``means_`` / ``evals_``).
"""
import numpy as np

from .. import ops
from ..engine import CVEngine
from ..folds import label2str


class _FittedMCCA:
    """The slice of ``mvlearn.embed.MCCA`` used by the reference (AlignMCCA.py:78,110,125)."""

    def __init__(self, loadings, means, evals, ranks):
        self.loadings_ = loadings
        self.means_ = means
        self.evals_ = evals
        self.signal_ranks = ranks
        self.n_views_ = len(loadings)

    def transform_view(self, X, view):
        X = np.asarray(X)
        return ops.project(X, self.loadings_[view], self.means_[view]).astype(np.float64)


class AlignMCCA:
    def __init__(self, n_components=10, regs=0.5, pca_var=1):
        self.n_components = n_components
        self.regs = regs
        self.pca_var = pca_var

    def fit(self, X, y):
        self.mcca = get_MCCA_transforms(X, y, n_components=self.n_components, regs=self.regs,
                                        pca_var=self.pca_var)

    def transform(self, X, idx=-1):
        if not self._check_fit():
            raise RuntimeError('Must call fit() before transforming data.')
        if idx == -1:
            return self._transform_multiple(X)
        if idx >= len(self.mcca.loadings_):
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
        out = [self._transform_single(x, i) for i, x in enumerate(X)]
        return (*out,)

    def _transform_single(self, X, idx):
        X = np.asarray(X)
        out = self.mcca.transform_view(X.reshape(-1, X.shape[-1]), idx)
        return out.reshape(X.shape[:-1] + (-1,))

    def _check_fit(self):
        return hasattr(self, 'mcca')


def get_MCCA_transforms(features, labels, n_components=10, regs=0.5, pca_var=1):
    """Fits MCCA on the class averages of every view (AlignMCCA.py:140-154) on the GPU and
    returns the fitted-model view the rest of the reference code expects."""
    views = [(np.asarray(X), np.zeros(len(X), dtype=np.int64), np.asarray(l))
             for X, l in zip(features, labels)]
    eng = CVEngine(views[0], views[1:], method='mcca', n_comp=int(n_components), regs=regs,
                   pca_var=pca_var)
    out = eng.align_mcca()
    P = len(views)
    Cs = [v[0].shape[-1] for v in views]
    loadings = [out['loadings'][0, v, :Cs[v], :].astype(np.float64) for v in range(P)]
    means = [out['mu'][0, v, :Cs[v]].astype(np.float64) for v in range(P)]
    ranks = None if not (0 < pca_var < 1) else [int(r) for r in out['r_eff'][0]]
    # mvlearn's deterministic sign: the largest-|entry| of each common-score column is positive
    strs = [label2str(v[2]) for v in views]
    vocab = eng.vocab
    shared = vocab[out['shared'][0]]
    cols = []
    for v in range(P):
        classes, inv = np.unique(strs[v], return_inverse=True)
        _, cm = ops.class_mean(views[v][0], inv)
        cols.append(cm[np.isin(classes, shared)].reshape(-1, Cs[v]))
    common = ops.project(np.hstack(cols), np.vstack(loadings), np.concatenate(means))
    for q in range(common.shape[1]):
        j = np.argmax(np.abs(common[:, q]))
        if common[j, q] < 0:
            for l in loadings:
                l[:, q] *= -1
    return _FittedMCCA(loadings, means, out['evals_mcca'][0].astype(np.float64), ranks)


def n_components_var(X, var):
    """0-based index of the first cumulative-variance value exceeding ``var`` of the
    uncentred spectrum of X (AlignMCCA.py:156-174; no +1, like the reference)."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    G = ops.gram_tn(X)
    ev, _ = ops.eig_sym(G) if G.shape[0] > 128 else ops.eig_sym(G.astype(np.float64), f64=True)
    return int(ops.select_k(ev[None], var, 1)[0])
