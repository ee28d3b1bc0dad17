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


def _same_labels(a, b):
    import numpy as np
    if a is b:
        return True
    if a is None or b is None:
        return False
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool((a == b).all())


def _groupable(job, first):
    """A job can share an engine with ``first`` when it has no per-job overrides, host inputs and
    the same patient shapes and labels (e.g. the CV iterations of one data set)."""
    import torch
    if len(job) > 3 and job[3]:
        return False
    pa, pb = [job[0]] + list(job[1]), [first[0]] + list(first[1])
    if len(pa) != len(pb):
        return False
    for a, b in zip(pa, pb):
        if isinstance(a[0], torch.Tensor) and a[0].is_cuda:
            return False
        if tuple(a[0].shape) != tuple(b[0].shape):
            return False
        if not _same_labels(a[1], b[1]) or not _same_labels(a[2], b[2]):
            return False
    return True


def cv_align_decode_stream(jobs, depth=8, method='mcca', device=None, group=None, **kw):
    """Pipelined form of ``cv_align_decode`` for many independent jobs (e.g. the 50 CV
    iterations x patient sets of scripts/aligned_decode_svm_ncv.py:332-456): ``jobs`` is an
    iterable of ``(target, cross, folds)`` or ``(target, cross, folds, overrides)`` (a dict of
    engine keywords for that job only); results are yielded in order.  Up to ``depth`` jobs
    are in flight, each on its own CUDA stream with its own upload, so the host->device copy
    and the latency-bound small solvers of one job overlap the kernels of the others.

    ``group`` (MCCA with tensor cores; default ``max(1, depth // 2)``, at most 8): consecutive
    jobs with the same patient shapes and labels and no overrides -- each still uploading its
    own inputs -- become REPLICAS of one engine, so their folds share the solver launches of
    one batch (a 20-fold job alone occupies a sixth of the GPU in the tile eigen-solvers; eight
    of them in one batch run at the resident-data rate)."""
    import collections
    import time
    from . import engine as _engine

    import numpy as np
    import torch

    from .engine import CVEngine, _lane_stream
    from .device import Context
    dev = Context.get(device).device
    if group is None:
        group = min(8, max(1, depth // 2)) if (method == 'mcca' and kw.get('use_tensor_cores')) else 1
    if method != 'mcca' or not kw.get('use_tensor_cores') or kw.get('decoder', 'linear').startswith('bag_'):
        group = 1
    slots = depth if group <= 1 else max(1, depth // group)      # engines in flight
    pending = collections.deque()
    it = iter(jobs)
    upload = _lane_stream(dev, 4)                 # one FIFO for every job's host->device copies
    nsub = 0
    done = False
    held = []                   # a job read ahead that did not fit the group being formed

    def next_job():
        if held:
            return held.pop()
        return next(it)

    def submit(members):
        nonlocal nsub
        lane = 8 + (nsub % slots) * 4              # leave room for the engines' extra lanes
        nsub += 1
        first = members[0]
        with torch.cuda.stream(_lane_stream(dev, lane)):
            jkw = dict(kw)
            if len(first) > 3 and first[3]:
                jkw.update(first[3])               # per-job engine keywords (ungrouped jobs only)
            seeds = jkw.pop('bag_seeds', None)
            reps = [(m[0], m[1]) for m in members[1:]]
            if reps:                                # the group is one batch (<= 148 folds, see below)
                jkw['max_batch'] = max(jkw.get('max_batch', 32), sum(len(m[2]) for m in members))
            eng = CVEngine(first[0], first[1], method=jkw.pop('method', method), device=dev, lane=lane,
                           replicas=reps or None, upload_stream=upload, **jkw)
            folds, rep = [], []
            for r, m in enumerate(members):
                folds += list(m[2])
                rep += [r] * len(m[2])
            gen = eng.run_gen(folds, bag_seeds=seeds, rep=rep if reps else None)
        return dict(eng=eng, gen=gen, res=None, done=False, wait=None, counts=[len(m[2]) for m in members])

    def split(ent):
        out, eng = ent['res'], ent['eng']
        n = len(ent['counts'])
        res, o = [], 0
        for r, c in enumerate(ent['counts']):
            res.append({'y_pred': out['y_pred'][o:o + c], 'k2': out['k2'][o:o + c],
                        'h2d_bytes': out['h2d_bytes'] // n + sum(v.h2d_bytes for v in eng.rviews[r]),
                        'd2h_bytes': out['d2h_bytes'] // n})
            o += c
        return res

    while True:
        t_iter = time.perf_counter()
        while not done and len(pending) < slots:
            try:
                members = [next_job()]
            except StopIteration:
                done = True
                break
            while group > 1 and len(members) < group and _groupable(members[0], members[0]):
                try:
                    nxt = next_job()
                except StopIteration:
                    done = True
                    break
                # one batch per group: at most 148 folds (one tile-solver problem per SM)
                if _groupable(nxt, members[0]) and sum(len(m[2]) for m in members) + len(nxt[2]) <= 148:
                    members.append(nxt)
                else:
                    held.append(nxt)
                    break
            pending.append(submit(members))
        if not pending:
            return
        progressed = False
        for pos, ent in enumerate(pending):
            if ent['done']:
                continue
            # grouped engines: only the two oldest compute; the younger ones have their uploads and
            # fold-invariant kernels queued and wait.  (Letting every engine compute at once makes
            # them finish together, then upload together with the SMs idle: measured 1450 folds/s
            # in that lock-step against the staggered pipeline below.)
            if group > 1 and pos >= 2:
                break
            # resume a job only when the GPU work it queued before yielding has finished
            if ent['wait'] is not None and len(pending) > 1 and not ent['wait'].query():
                continue
            progressed = True
            with torch.cuda.stream(ent['eng'].stream):
                try:
                    if next(ent['gen']) == 'host':
                        ent['wait'] = None               # pure host step: resumable at once
                    else:
                        ent['wait'] = torch.cuda.Event()
                        ent['wait'].record(ent['eng'].stream)
                except StopIteration as e:
                    ent['res'], ent['done'] = e.value, True
        while pending and pending[0]['done']:
            progressed = True
            for r in split(pending.popleft()):
                yield r
        if not progressed:
            time.sleep(_engine._IDLE_SLEEP)
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
