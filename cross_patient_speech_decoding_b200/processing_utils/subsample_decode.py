"""Electrode-subsampling decode loops (BASELINE config 4), batched.

The reference's subsample scripts (scripts/aligned_decode_grid_subsample.py:280-400 and its
cross_patient / spatialAvg / pitch variants) run, for every channel subsample of the target
patient: a random subsample of every cross patient (``np.random.choice``), a shuffled
StratifiedKFold, and one fit / predict per fold.  Here every subsample is one *job* of
``cv_align_decode_stream``: the patients stay resident on the device, a subsample is a channel
gather there (``device_subsample.gather_channels``), and ``depth`` jobs are in flight at once.
The numpy global RNG is consumed in the script's order (cross-patient choices, the split, one
``SVC.fit`` seed draw per fold), so the folds are the ones the script would have used."""
import numpy as np

from .. import cv_align_decode_stream
from ..folds import cv_splits
from . import device_subsample as ds


def subsample_jobs(target, cross, tar_subsamp_idx_list, cross_subsamp_idx_lists, n_folds=20,
                   fit_draws=1, device=None, record=None, own=None):
    """Generator of ``(target_view, cross_views, folds)`` jobs.  target / cross: ``(X, y,
    y_align)`` triples (host arrays); ``tar_subsamp_idx_list``: channel index arrays for the
    target; ``cross_subsamp_idx_lists[p]``: the candidate index arrays of cross patient p.
    ``record`` (a list) receives ``(chosen cross subsample per patient, folds)`` for EVERY job;
    ``own`` (set of job indices, default all) selects the jobs that are gathered and yielded --
    the numpy global RNG is consumed for all of them, so every rank of a sharded run sees the
    script's stream."""
    Xt = Xc = None
    lab = np.asarray(target[1])
    for j, sub in enumerate(tar_subsamp_idx_list):
        mine = own is None or j in own
        if mine and Xt is None:
            Xt = ds.resident(target[0], device)
            Xc = [ds.resident(c[0], device) for c in cross]
        chosen = [int(np.random.choice(len(cand))) for cand in cross_subsamp_idx_lists]
        folds = cv_splits(lab, n_folds)
        for _ in range(len(folds) * fit_draws):
            np.random.randint(np.iinfo('i').max)         # SVC.fit's libsvm seed, one per fold
        if record is not None:
            record.append((chosen, folds))
        if not mine:
            continue
        tv = (ds.gather_channels(Xt, np.asarray(sub)), target[1], target[2])
        cvs = [(ds.gather_channels(Xc[p], np.asarray(cand[r])), cross[p][1], cross[p][2])
               for p, (cand, r) in enumerate(zip(cross_subsamp_idx_lists, chosen))]
        yield tv, cvs, folds


def subsample_decode(target, cross, tar_subsamp_idx_list, cross_subsamp_idx_lists, n_folds=20,
                     method='cca', depth=8, fit_draws=1, device=None, **kw):
    """Runs every subsample; returns the scripts' result fields ``y_true``, ``y_pred``,
    ``wrong_trs``, ``accs`` (one entry per target subsample, balanced accuracy over its folds,
    aligned_decode_grid_subsample.py:386-400) plus ``chosen`` (cross-patient subsample indices).
    Under torch.distributed (torchrun) the subsamples are dealt to the ranks in contiguous
    blocks, every rank decodes its own on its GPU, and one all_gather of the predicted labels
    gives every rank the full result (sharding.gather_predictions)."""
    from sklearn.metrics import balanced_accuracy_score
    from .. import sharding
    rank, world, _ = sharding.init_from_env()
    lab = np.asarray(target[1])
    rec = []
    njob = len(tar_subsamp_idx_list)
    mine = sharding.shard_units(njob, 1, rank, world)
    jobs = subsample_jobs(target, cross, tar_subsamp_idx_list, cross_subsamp_idx_lists, n_folds,
                          fit_draws, device, rec, own=set(mine) if world > 1 else None)
    preds = [np.concatenate(res['y_pred']) for res in
             cv_align_decode_stream(jobs, depth=depth, method=method, device=device, **kw)]
    for _ in jobs:            # not-owned trailing jobs still consume the RNG
        pass
    allp = sharding.gather_predictions(mine, preds)
    out = dict(y_true=[], y_pred=[], wrong_trs=[], accs=[], chosen=[])
    for j in range(njob):
        chosen, folds = rec[j]
        yt = np.concatenate([lab[te] for _, te in folds])
        yp = np.asarray(allp[j])
        te_all = np.concatenate([te for _, te in folds])
        out['y_true'].append(yt.tolist())
        out['y_pred'].append(yp.tolist())
        out['wrong_trs'].append(te_all[yt != yp].tolist())
        out['accs'].append(balanced_accuracy_score(yt, yp))
        out['chosen'].append(chosen)
    return out
