"""Nested hyper-parameter search around the batched align -> reduce -> decode path.

The reference's nested CV (scripts/aligned_decode_svm_ncv.py:149-194, 388-405) wraps every outer
fold in ``skopt.BayesSearchCV(model, param_grid, n_iter=25, n_points=5, cv=StratifiedKFold(20,
shuffle=True), refit=False)``: 5 rounds of 5 candidates, every candidate scored by the mean
accuracy over 20 inner folds -- 500 fits per outer fold, the largest fold multiplier in the
repository.  scikit-optimize is not installed here (and its optimiser draws from numpy's global
RNG, so a seeded run could not be reproduced draw for draw anyway); ``BayesSearch`` restates the
loop: a Gaussian-process surrogate (Matern 5/2 + noise, inputs scaled to the unit cube, as skopt
builds it) over the same search-space notation, ``n_initial_points`` uniform random candidates
first, then ``n_points`` proposals per round by expected improvement with the constant-liar
strategy (skopt's ``ask(n_points, strategy='cl_min')``).  Every round is ONE call of
``search_align_decode``: all candidates x inner folds go to the GPU as jobs in flight.

The surrogate and the proposal arithmetic are host control logic (a few hundred flops per
round); the fits and predictions they steer are the CUDA path.
"""
import numpy as np


class Dimension:
    """One axis of the search space in skopt's notation: ``(lo, hi)`` ints -> integer range,
    ``(lo, hi, 'uniform' | 'log-uniform')`` -> real range, any other sequence -> ordered grid."""

    def __init__(self, spec):
        if isinstance(spec, tuple) and len(spec) == 2 and all(isinstance(v, (int, np.integer)) for v in spec):
            self.kind, self.lo, self.hi = 'int', int(spec[0]), int(spec[1])
        elif isinstance(spec, tuple) and len(spec) in (2, 3) and not isinstance(spec[-1], (list, np.ndarray)) \
                and all(isinstance(v, (int, float, np.integer, np.floating)) for v in spec[:2]):
            prior = spec[2] if len(spec) == 3 else 'uniform'
            if prior not in ('uniform', 'log-uniform'):
                raise ValueError("prior must be 'uniform' or 'log-uniform'")
            self.kind, self.lo, self.hi = ('log' if prior == 'log-uniform' else 'real'), float(spec[0]), float(spec[1])
        else:
            self.kind, self.grid = 'grid', list(np.asarray(spec).tolist())
            if not self.grid:
                raise ValueError('empty grid')

    def to_unit(self, v):
        if self.kind == 'int' or self.kind == 'real':
            return (float(v) - self.lo) / max(self.hi - self.lo, 1e-300)
        if self.kind == 'log':
            return (np.log(float(v)) - np.log(self.lo)) / (np.log(self.hi) - np.log(self.lo))
        return self.grid.index(v) / max(len(self.grid) - 1, 1)

    def from_unit(self, u):
        u = float(min(max(u, 0.0), 1.0))
        if self.kind == 'int':
            return int(min(self.hi, max(self.lo, round(self.lo + u * (self.hi - self.lo)))))
        if self.kind == 'real':
            return self.lo + u * (self.hi - self.lo)
        if self.kind == 'log':
            return float(np.exp(np.log(self.lo) + u * (np.log(self.hi) - np.log(self.lo))))
        return self.grid[int(round(u * (len(self.grid) - 1)))]


def _expected_improvement(mu, sd, best, xi=0.01):
    """EI for MINIMISATION of the surrogate (skopt minimises -score)."""
    from scipy.stats import norm
    sd = np.maximum(sd, 1e-12)
    imp = best - mu - xi
    z = imp / sd
    return imp * norm.cdf(z) + sd * norm.pdf(z)


class BayesSearch:
    """ask / tell optimiser over a dict search space (maximises the told scores)."""

    def __init__(self, space, n_initial_points=10, n_candidates=2000, random_state=None):
        self.names = list(space)
        self.dims = [Dimension(space[k]) for k in self.names]
        self.n_initial_points = int(n_initial_points)
        self.n_candidates = int(n_candidates)
        # skopt's default random_state=None also means numpy's global RNG
        if random_state is None:
            self.rng = np.random.mtrand._rand
        elif isinstance(random_state, np.random.RandomState):
            self.rng = random_state
        else:
            self.rng = np.random.RandomState(random_state)
        self.X, self.y = [], []          # unit-cube points, scores

    def _params(self, u):
        return {k: d.from_unit(v) for k, d, v in zip(self.names, self.dims, u)}

    def _snap(self, u):
        """Unit-cube point after the round trip through the parameter values (integers, grids)."""
        return np.array([d.to_unit(d.from_unit(v)) for d, v in zip(self.dims, u)])

    def ask(self, n_points=1):
        out = []
        Xl, yl = [np.asarray(x) for x in self.X], [-s for s in self.y]      # minimise -score
        for _ in range(n_points):
            if len(Xl) < self.n_initial_points or len(set(yl)) < 2:
                u = self._snap(self.rng.uniform(size=len(self.dims)))
            else:
                from sklearn.gaussian_process import GaussianProcessRegressor
                from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel
                kern = ConstantKernel(1.0, (1e-2, 1e3)) * Matern(length_scale=np.ones(len(self.dims)),
                                                                length_scale_bounds=(1e-2, 1e2), nu=2.5) \
                    + WhiteKernel(1e-3, (1e-8, 1e1))
                gp = GaussianProcessRegressor(kern, normalize_y=True, n_restarts_optimizer=2,
                                              random_state=self.rng.randint(2 ** 31 - 1))
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter('ignore')
                    gp.fit(np.vstack(Xl), np.asarray(yl))
                cand = np.vstack([self._snap(c) for c in self.rng.uniform(size=(self.n_candidates, len(self.dims)))])
                mu, sd = gp.predict(cand, return_std=True)
                ei = _expected_improvement(mu, sd, min(yl))
                # never propose a point that was already evaluated / proposed in this round
                seen = np.vstack(Xl)
                far = np.abs(cand[:, None, :] - seen[None, :, :]).max(axis=2).min(axis=1) > 1e-9
                ei = np.where(far, ei, -np.inf)
                u = cand[int(np.argmax(ei))]
            out.append(self._params(u))
            Xl.append(u)
            yl.append(min(yl) if yl else 0.0)          # constant liar: pretend it scored the best
        return out

    def tell(self, params_list, scores):
        for p, s in zip(params_list, scores):
            self.X.append(np.array([d.to_unit(p[k]) for k, d in zip(self.names, self.dims)]))
            self.y.append(float(s))

    @property
    def best_index(self):
        return int(np.argmax(self.y))


# the script's parameter names -> engine keywords (scripts/aligned_decode_svm_ncv.py:149-176)
PARAM_TO_ENGINE = {'n_comp': 'n_comp', 'regs': 'regs', 'pca_var': 'pca_var',
                   'decoder__dimredreshape__n_components': 'decoder_var'}


def engine_keywords(params):
    """Script-style parameter dict -> engine keywords; parameters of decoder stages the batched
    engine does not have (the BaggingClassifier keys of the reference's MCCA grid, which its own
    checked-in pipeline does not contain either) are rejected."""
    kw = {}
    for k, v in params.items():
        if k not in PARAM_TO_ENGINE:
            raise ValueError('parameter %r has no counterpart in the batched engine' % k)
        kw[PARAM_TO_ENGINE[k]] = v
    return kw


def bayes_search_align_decode(target, cross, space, n_folds=20, n_iter=25, n_points=5, method='cca',
                              n_initial_points=10, random_state=None, depth=4, device=None,
                              shard=False, **kw):
    """``BayesSearchCV(model, space, n_iter, n_points, cv=StratifiedKFold(n_folds, shuffle=True),
    refit=False).fit(X, y, y_align=...)`` on the batched engine.  ``target`` holds the TRAIN
    trials of the outer fold.  As in sklearn, every round of candidates gets a fresh shuffled
    split; with ``random_state=None`` splits and proposals draw from numpy's global RNG (as the
    reference's do), an int / RandomState makes the search independent of everything around it
    (what a sharded run needs: every rank sees the same search for the same unit).  Returns ``best_params_`` / ``best_score_`` /
    ``cv_results_``-like lists (script-style parameter names)."""
    from . import search_align_decode
    from .folds import cv_splits
    rs = None if random_state is None else (random_state if isinstance(random_state, np.random.RandomState)
                                            else np.random.RandomState(random_state))
    opt = BayesSearch(space, n_initial_points=n_initial_points, random_state=rs)
    lab = np.asarray(target[1])
    params_all, scores_all = [], []
    left = int(n_iter)
    while left > 0:
        npts = min(n_points, left)
        cands = opt.ask(npts)
        folds = cv_splits(lab, n_folds, random_state=rs)
        ekw = [engine_keywords(c) for c in cands]
        try:
            scores = search_align_decode(target, cross, ekw, folds, method=method, depth=depth,
                                         device=device, shard=shard, **kw)['scores']
        except ValueError:
            # an infeasible candidate (e.g. MCCA n_components above the summed signal ranks, which
            # mvlearn rejects too): score the round one candidate at a time, failures score 0 --
            # sklearn's error_score for a failed fit, mapped to the worst accuracy
            scores = []
            for e in ekw:
                try:
                    scores.append(float(search_align_decode(target, cross, [e], folds, method=method,
                                                            depth=depth, device=device, shard=shard,
                                                            **kw)['scores'][0]))
                except ValueError:
                    scores.append(0.0)
        opt.tell(cands, scores)
        params_all += cands
        scores_all += [float(s) for s in scores]
        left -= npts
    b = int(np.argmax(scores_all))
    return dict(best_params_=params_all[b], best_score_=scores_all[b], best_index_=b,
                params=params_all, mean_test_score=scores_all)
