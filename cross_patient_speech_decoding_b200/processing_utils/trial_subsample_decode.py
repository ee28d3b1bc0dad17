"""Cross-patient trial-count sweep, batched.

The reference's ``scripts/aligned_decode_cross_patient_subsample.py:290-388`` asks how decoding
improves with the amount of pooled cross-patient data: for every trial count ``k`` (5, 30, 55,
... up to the median trial count of the cross patients) and every iteration it draws ``k``
trials of every cross patient (``np.random.choice(N, k, replace=False)``; all trials when the
patient has fewer), splits the target with a shuffled StratifiedKFold and fits / predicts one
aligned decoder per fold; the result is a ``(len(k), n_iter)`` matrix of balanced accuracies.

Here every (k, iteration) is one job of ``cv_align_decode_stream``: the patients stay resident on
the device, a trial subset is a row gather there (``device_subsample.gather_trials``), ``depth``
jobs are in flight.  The numpy global RNG is consumed in the script's order (the cross patients'
draws, the split, one ``SVC.fit`` seed draw per fold), so a seeded run decodes the reference's
trial subsets and folds.  Under torch.distributed the jobs are dealt to the ranks."""
import numpy as np

from .. import cv_align_decode_stream
from ..folds import cv_splits
from . import device_subsample as ds


def trial_counts(cross, trial_step=25, start=5):
    """``np.arange(5, ceil(median(trial counts)) + 1, trial_step)`` (script lines 285-288)."""
    max_trs = int(np.ceil(np.median([c[0].shape[0] for c in cross])))
    return np.arange(start, max_trs + 1, trial_step)


def trial_subsample_jobs(target, cross, k_list, n_iter, n_folds=20, fit_draws=1, device=None,
                         record=None, own=None):
    """Generator of ``(target, cross_views, folds)`` jobs in (k, iteration) order.  ``record``
    receives ``(k, iteration, trial indices per cross patient (None = all), folds)`` for EVERY
    job; ``own`` (set of job numbers) selects the ones that are gathered and yielded."""
    Xt = Xc = None
    lab = np.asarray(target[1])
    j = 0
    for k in k_list:
        for it in range(n_iter):
            mine = own is None or j in own
            if mine and Xt is None:
                Xt = ds.resident(target[0], device)
                Xc = [ds.resident(c[0], device) for c in cross]
            picks = [None if c[0].shape[0] < k else np.random.choice(c[0].shape[0], int(k), replace=False)
                     for c in cross]
            folds = cv_splits(lab, n_folds)
            for _ in range(len(folds) * fit_draws):
                np.random.randint(np.iinfo('i').max)     # SVC.fit's libsvm seed, one per fold
            if record is not None:
                record.append((int(k), it, picks, folds))
            j += 1
            if not mine:
                continue
            cvs = []
            for p, (c, idx) in enumerate(zip(cross, picks)):
                if idx is None:
                    cvs.append((Xc[p], c[1], c[2]))
                else:
                    cvs.append((ds.gather_trials(Xc[p], idx), np.asarray(c[1])[idx], np.asarray(c[2])[idx]))
            yield (Xt, target[1], target[2]), cvs, folds


def trial_subsample_decode(target, cross, k_list=None, n_iter=50, n_folds=20, method='cca', depth=8,
                           fit_draws=1, trial_step=25, device=None, **kw):
    """Returns the script's result fields ``acc_mat`` (len(k) x n_iter balanced accuracies) and
    ``trial_vec`` (pooled cross-patient trials per k) plus ``k_trials_per_pt``, ``y_pred`` and
    ``y_true`` per job."""
    from sklearn.metrics import balanced_accuracy_score
    from .. import sharding
    rank, world, _ = sharding.init_from_env()
    k_list = trial_counts(cross, trial_step) if k_list is None else np.asarray(k_list)
    njob = len(k_list) * n_iter
    mine = sharding.shard_units(njob, 1, rank, world)
    rec = []
    jobs = trial_subsample_jobs(target, cross, k_list, n_iter, n_folds, fit_draws, device, rec,
                                own=set(mine) if world > 1 else None)
    preds = [np.concatenate(r['y_pred']) for r in
             cv_align_decode_stream(jobs, depth=depth, method=method, device=device, **kw)]
    allp = sharding.gather_predictions(mine, preds)
    lab = np.asarray(target[1])
    acc = np.full((len(k_list), n_iter), np.nan)
    tvec = np.full(len(k_list), np.nan)
    y_pred, y_true = [], []
    for j, (k, it, picks, folds) in enumerate(rec):
        yt = np.concatenate([lab[te] for _, te in folds])
        yp = np.asarray(allp[j])
        ki = j // n_iter
        acc[ki, it] = balanced_accuracy_score(yt, yp)
        tvec[ki] = sum(c[0].shape[0] if idx is None else len(idx) for c, idx in zip(cross, picks))
        y_pred.append(yp.tolist())
        y_true.append(yt.tolist())
    return dict(acc_mat=acc, trial_vec=tvec, k_trials_per_pt=np.asarray(k_list), y_pred=y_pred,
                y_true=y_true)
