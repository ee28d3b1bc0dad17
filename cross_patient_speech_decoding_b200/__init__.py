"""B200-native cross-validated align -> reduce -> decode for cross-patient speech decoding.

Drop-in for the hot path of coganlab/cross_patient_speech_decoding: the sklearn-style
classes live in ``alignment``, ``decomposition`` and ``decoders`` (same module and class
names as the reference's ``aligned_decoding`` package); ``cv_align_decode`` is the batched
entry point that runs a whole list of CV folds through the CUDA kernels at once.
"""


def cv_align_decode(target, cross, folds, method='mcca', **kw):
    """Fits + scores every (train_idx, test_idx) fold for ``target`` pooled with ``cross``.

    target / cross entries are ``(X (trials, time, channels), y, y_align)`` host arrays --
    the ``(D, lab, lab_full)`` triples of the reference's ``decoding_data_from_dict``
    (alignment/alignment_utils.py:127-157).  Returns a dict with ``y_pred`` (one array per
    fold), ``k2`` and the bytes moved host<->device.  ``method``: 'mcca'
    (crossPtDecoder_mcca), 'jointpca' (crossPtDecoder_jointDimRed + JointPCA), 'cca'
    (crossPtDecoder_sepAlign + AlignCCA) or 'none' (crossPtDecoder_sepDimRed).  The decoder is
    PCA(decoder_var) followed by ``decoder``: 'linear' (default; one-vs-rest squared-hinge linear
    SVM, the north star's dual-CD decoder), 'svc_rbf' or 'svc_linear' (libsvm-style C-SVC with
    one-vs-one votes -- with ``class_weight='balanced'`` the first is the reference scripts' own
    ``SVC(kernel='rbf', class_weight='balanced')``, scripts/aligned_decode_svm_ncv.py:313-317).
    """
    from .engine import CVEngine
    bag_seeds = kw.pop('bag_seeds', None)      # bagging decoders: per-fold estimator seeds
    eng = CVEngine(target, cross, method=method, **kw)
    out = eng.run(folds, bag_seeds=bag_seeds)
    out['h2d_bytes'] += sum(v.h2d_bytes for v in eng.views)
    return out


def cv_align_decode_stream(jobs, depth=8, method='mcca', device=None, **kw):
    """Pipelined form of ``cv_align_decode`` for many independent jobs (e.g. the 50 CV
    iterations x patient sets of scripts/aligned_decode_svm_ncv.py:332-456): ``jobs`` is an
    iterable of ``(target, cross, folds)`` or ``(target, cross, folds, overrides)`` (a dict of
    engine keywords for that job only); results are yielded in order.  Up to ``depth`` jobs
    are in flight, each on its own CUDA stream with its own upload, so the host->device copy
    and the latency-bound small solvers of one job overlap the kernels of the others."""
    import collections

    import torch

    from .engine import CVEngine, _lane_stream
    from .device import Context
    dev = Context.get(device).device
    pending = collections.deque()
    it = iter(jobs)
    nsub = 0
    done = False

    def submit(job):
        nonlocal nsub
        lane = 8 + (nsub % depth) * 4           # leave room for the engines' extra lanes
        nsub += 1
        with torch.cuda.stream(_lane_stream(dev, lane)):
            jkw = dict(kw)
            if len(job) > 3 and job[3]:
                jkw.update(job[3])                 # per-job engine keywords
            seeds = jkw.pop('bag_seeds', None)
            eng = CVEngine(job[0], job[1], method=jkw.pop('method', method), device=dev, lane=lane,
                           **jkw)
            gen = eng.run_gen(job[2], bag_seeds=seeds)
        return [eng, gen, None, False, None]     # engine, generator, result, done, wait event

    import time
    while True:
        t_iter = time.perf_counter()
        while not done and len(pending) < depth:
            try:
                job = next(it)
            except StopIteration:
                done = True
                break
            pending.append(submit(job))
        if not pending:
            return
        progressed = False
        for ent in pending:
            if ent[3]:
                continue
            # resume a job only when the GPU work it queued before yielding has finished
            if ent[4] is not None and len(pending) > 1 and not ent[4].query():
                continue
            progressed = True
            with torch.cuda.stream(ent[0].stream):
                try:
                    if next(ent[1]) == 'host':
                        ent[4] = None               # pure host step: resumable at once
                    else:
                        ent[4] = torch.cuda.Event()
                        ent[4].record(ent[0].stream)
                except StopIteration as e:
                    out = e.value
                    out['h2d_bytes'] += sum(v.h2d_bytes for v in ent[0].views)
                    ent[2], ent[3] = out, True
        while pending and pending[0][3]:
            progressed = True
            yield pending.popleft()[2]
        if not progressed:
            time.sleep(0)
            cv_align_decode_stream.idle_s += time.perf_counter() - t_iter


def search_align_decode(target, cross, candidates, inner_folds, method='mcca', depth=4, device=None,
                        shard=True, **kw):
    """Hyper-parameter search over the align -> reduce -> decode path, batched: every candidate
    (a dict of engine keywords, e.g. ``{'n_comp': 0.9, 'decoder_var': 0.6}``) is scored on all
    ``inner_folds`` as one job of ``cv_align_decode_stream``; the patients are uploaded once and
    shared by all candidates.  The inner loop of the reference's nested CV
    (scripts/aligned_decode_svm_ncv.py:388-405: a search with ``refit=False`` whose score is the
    mean per-fold accuracy) without one fit / predict call per (candidate, fold).
    Returns ``scores`` (mean per-fold accuracy per candidate), ``best_index``, ``best_params``
    and ``y_pred`` (per candidate, per fold).  Under torch.distributed the candidates are dealt
    to the ranks (one all_gather of the predicted labels at the end) unless ``shard=False`` (a
    search nested inside an outer loop that is itself sharded)."""
    import numpy as np

    from . import sharding
    from .processing_utils import device_subsample as ds
    rank, world, _ = sharding.init_from_env() if shard else (0, 1, 0)
    mine = sharding.shard_units(len(candidates), 1, rank, world)   # candidates dealt to the ranks
    lab = np.asarray(target[1])
    preds = []
    if mine:
        tv = (ds.resident(target[0], device), target[1], target[2])
        cvs = [(ds.resident(c[0], device), c[1], c[2]) for c in cross]
        jobs = ((tv, cvs, inner_folds, dict(candidates[c])) for c in mine)
        for res in cv_align_decode_stream(jobs, depth=depth, method=method, device=device, **kw):
            preds.append(np.concatenate(res['y_pred']) if res['y_pred'] else np.zeros(0, dtype=np.int32))
    allp = sharding.gather_predictions(mine, preds, enabled=shard)
    cuts = np.cumsum([len(te) for _, te in inner_folds])[:-1]
    scores, per_fold = [], []
    for c in range(len(candidates)):
        yps = np.split(np.asarray(allp[c]), cuts)
        accs = [float(np.mean(yp == lab[te])) for yp, (_, te) in zip(yps, inner_folds)]
        scores.append(float(np.mean(accs)))
        per_fold.append(yps)
    scores = np.asarray(scores)
    best = int(np.argmax(scores)) if len(scores) else -1
    return dict(scores=scores, best_index=best, best_params=dict(candidates[best]) if best >= 0 else None,
                y_pred=per_fold)


cv_align_decode_stream.idle_s = 0.0     # host time spent with every in-flight job waiting on the GPU
