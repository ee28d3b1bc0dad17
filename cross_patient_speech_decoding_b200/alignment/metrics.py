"""Alignment-quality metrics of the reference (alignment/metrics.py): per-condition Pearson
correlation between a target's and another patient's aligned condition averages.  The
correlations are computed on the GPU (``cpsd_pearson_rows``, fp64); the p-value is scipy's exact
two-sided test for r (a Beta(n/2 - 1, n/2 - 1) law on [-1, 1]) evaluated on the host from r
and n."""
import numpy as np
import torch

from ..device import Context, ptr


def pt_corr(target, to_corr, p_vals=False):
    """target, to_corr: (n_conditions, n_timepoints, n_features) -> r per condition
    (and p-values when ``p_vals``), as alignment/metrics.py:41-68."""
    a = np.ascontiguousarray(target, dtype=np.float64)
    b = np.ascontiguousarray(to_corr, dtype=np.float64)
    if a.shape != b.shape:
        raise ValueError('x and y must have the same length.')
    ncnd = a.shape[0]
    n = int(np.prod(a.shape[1:]))
    if n < 2:
        raise ValueError('x and y must have length at least 2.')
    ctx = Context.get(None)
    r = ctx.empty((ncnd,), torch.float64)
    a_d, b_d = ctx.upload(a.reshape(ncnd, n)), ctx.upload(b.reshape(ncnd, n))
    ctx.call('cpsd_pearson_rows', ptr(a_d), ptr(b_d), ncnd, n, ptr(r))
    r = r.cpu().numpy()
    if not p_vals:
        return r
    from scipy import special
    if n == 2:
        p = np.ones(ncnd)
    else:
        # P(|R| >= |r|) for R ~ 2 Beta(n/2 - 1, n/2 - 1) - 1
        ab = n / 2.0 - 1.0
        p = 2.0 * special.betainc(ab, ab, 0.5 * (1.0 - np.abs(r)))
        p = np.minimum(p, 1.0)
    return r, p


def pt_corr_multi(target, to_corr_list, p_vals=False):
    """alignment/metrics.py:12-38."""
    res = [pt_corr(target, tc, p_vals=p_vals) for tc in to_corr_list]
    if p_vals:
        return [x[0] for x in res], [x[1] for x in res]
    return res
