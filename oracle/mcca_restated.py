"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU (numpy/scipy, float64) restatement of ``mvlearn.embed.MCCA`` for the one
call pattern the reference uses (``alignment/AlignMCCA.py:152-153``:
``MCCA(n_components, regs, signal_ranks).fit(views)``; ``transform_view`` at
``AlignMCCA.py:110,125``; ``loadings_`` at ``AlignMCCA.py:78``).

PARITY UNPINNED: mvlearn is a third-party dependency of the reference that is
neither vendored under /root/reference, nor pinned in requirements.txt /
environment.yml, nor installed here.  This file restates the published
algorithm of mvlearn 0.5.0 ``mvlearn/embed/mcca.py`` (``_mcca_gevp``,
``_i_mcca`` with method 'gevp', ``_construct_mcca_gevp_data``, SURVEY.md
Appendix B.5).  It is anchored by known-answer checks against importable
reference code (tests/test_oracle.py): for two views, no regularisation and
full rank the generalised eigenvalues are ``1 + rho`` with ``rho`` the
canonical correlations of the reference's ``CCA_align`` (AlignCCA.py:235).
"""
import numpy as np
import scipy.linalg


def _gevp_blocks(views, regs):
    """LHS/RHS of the SUMCOR generalised eigenproblem.

    Off-diagonal LHS blocks are cross scatters ``X_a^T X_b`` (no 1/n), the
    diagonal blocks of LHS and RHS are both ``(1-r) X_b^T X_b + r I`` (or the
    plain scatter when ``regs`` is None)."""
    dims = [v.shape[1] for v in views]
    offs = np.concatenate([[0], np.cumsum(dims)])
    n = offs[-1]
    lhs = np.zeros((n, n))
    rhs = np.zeros((n, n))
    for a, Xa in enumerate(views):
        sa = slice(offs[a], offs[a + 1])
        for b, Xb in enumerate(views):
            if b <= a:
                continue
            sb = slice(offs[b], offs[b + 1])
            lhs[sa, sb] = Xa.T @ Xb
            lhs[sb, sa] = lhs[sa, sb].T
        scat = Xa.T @ Xa
        if regs is not None:
            scat = (1.0 - regs) * scat + regs * np.eye(dims[a])
        lhs[sa, sa] = scat
        rhs[sa, sa] = scat
    return lhs, rhs, offs


def _det_signs(common, scores, loadings):
    """Sign convention: largest-|entry| of each common-score column > 0."""
    for r in range(common.shape[1]):
        j = np.argmax(np.abs(common[:, r]))
        if common[j, r] < 0:
            common[:, r] *= -1
            for b in range(len(loadings)):
                scores[b][:, r] *= -1
                loadings[b][:, r] *= -1


def mcca_gevp(views, n_components, regs):
    """Top ``n_components`` generalised eigenvectors, split per view."""
    lhs, rhs, offs = _gevp_blocks(views, regs)
    n = lhs.shape[0]
    if n_components > n:
        raise ValueError('n_components=%d exceeds total dimension %d'
                         % (n_components, n))
    evals, vecs = scipy.linalg.eigh(lhs, rhs,
                                    subset_by_index=[n - n_components, n - 1])
    evals, vecs = evals[::-1], vecs[:, ::-1]
    loadings = [vecs[offs[b]:offs[b + 1]].copy() for b in range(len(views))]
    scores = [views[b] @ loadings[b] for b in range(len(views))]
    common = sum(scores)
    norms = np.linalg.norm(common, axis=0)
    common = common / norms
    _det_signs(common, scores, loadings)
    return loadings, evals


class MCCARestated:
    """Drop-in for the three members of ``mvlearn.embed.MCCA`` the reference
    touches: ``fit``, ``transform_view``, ``loadings_`` (plus ``means_`` and
    ``evals_`` for tests)."""

    def __init__(self, n_components=1, regs=None, signal_ranks=None):
        self.n_components = n_components
        self.regs = regs
        self.signal_ranks = signal_ranks

    def fit(self, Xs):
        Xs = [np.asarray(X, dtype=np.float64) for X in Xs]
        self.means_ = [X.mean(axis=0) for X in Xs]
        Xc = [X - m for X, m in zip(Xs, self.means_)]
        if self.signal_ranks is None:
            self.loadings_, self.evals_ = mcca_gevp(Xc, self.n_components,
                                                    self.regs)
            return self
        # informative MCCA: per-view rank-r SVD, GEVP on reduced views
        reduced, backs = [], []
        for X, r in zip(Xc, self.signal_ranks):
            U, D, Vt = np.linalg.svd(X, full_matrices=False)
            U, D, V = U[:, :r], D[:r], Vt[:r].T
            if self.regs is None:      # normalised scores
                reduced.append(U)
                backs.append(V / D)
            else:                       # un-normalised scores
                reduced.append(U * D)
                backs.append(V)
        red_load, self.evals_ = mcca_gevp(reduced, self.n_components,
                                          self.regs)
        self.loadings_ = [B @ W for B, W in zip(backs, red_load)]
        return self

    def transform_view(self, X, view):
        return (np.asarray(X) - self.means_[view]) @ self.loadings_[view]
