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
    (crossPtDecoder_mcca), 'cca' (crossPtDecoder_sepAlign + AlignCCA) or 'none'
    (crossPtDecoder_sepDimRed); the decoder is PCA(decoder_var) -> one-vs-rest linear SVM.
    """
    from .engine import CVEngine
    eng = CVEngine(target, cross, method=method, **kw)
    out = eng.run(folds)
    out['h2d_bytes'] += sum(v.h2d_bytes for v in eng.views)
    return out
