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
    eng = CVEngine(target, cross, method=method, **kw)
    out = eng.run(folds)
    out['h2d_bytes'] += sum(v.h2d_bytes for v in eng.views)
    return out


def cv_align_decode_stream(jobs, depth=8, method='mcca', device=None, **kw):
    """Pipelined form of ``cv_align_decode`` for many independent jobs (e.g. the 50 CV
    iterations x patient sets of scripts/aligned_decode_svm_ncv.py:332-456): ``jobs`` is an
    iterable of ``(target, cross, folds)``; results are yielded in order.  Up to ``depth`` jobs
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
            eng = CVEngine(job[0], job[1], method=method, device=dev, lane=lane, **kw)
            gen = eng.run_gen(job[2])
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


cv_align_decode_stream.idle_s = 0.0     # host time spent with every in-flight job waiting on the GPU
