"""Fold / label bookkeeping on the host (integer work, stays on the CPU by design).

``stratified_kfold`` reproduces ``sklearn.model_selection.StratifiedKFold(n_splits,
shuffle=True)`` -- the splitter of the reference CV loops
(scripts/aligned_decode_svm_ncv.py:336-339, scripts/aligned_decode_svm.py:177) -- draw
for draw on numpy's legacy global ``RandomState``, so that after ``np.random.seed(s)``
the fold indices are bit-identical to the reference's.  ``kfold`` is the reference's
``KFold(shuffle=True)`` fallback (aligned_decode_svm_ncv.py:340-342).
"""
import numpy as np


def _rng(random_state):
    if random_state is None:
        return np.random.mtrand._rand          # what sklearn's check_random_state(None) returns
    if isinstance(random_state, np.random.RandomState):
        return random_state
    return np.random.RandomState(random_state)


def stratified_test_folds(y, n_splits, shuffle=True, random_state=None):
    """Fold id of every sample (sklearn _split.py StratifiedKFold._make_test_folds)."""
    rng = _rng(random_state)
    y = np.asarray(y)
    _, y_idx, y_inv = np.unique(y, return_index=True, return_inverse=True)
    _, class_perm = np.unique(y_idx, return_inverse=True)   # classes by first appearance
    y_enc = class_perm[y_inv]
    n_classes = len(y_idx)
    counts = np.bincount(y_enc)
    if np.all(n_splits > counts):
        raise ValueError('n_splits=%d cannot be greater than the number of members in each '
                         'class.' % n_splits)
    y_order = np.sort(y_enc)
    allocation = np.asarray([np.bincount(y_order[i::n_splits], minlength=n_classes)
                             for i in range(n_splits)])
    test_folds = np.empty(len(y), dtype='i')
    for k in range(n_classes):
        folds_for_class = np.arange(n_splits).repeat(allocation[:, k])
        if shuffle:
            rng.shuffle(folds_for_class)
        test_folds[y_enc == k] = folds_for_class
    return test_folds


def stratified_kfold(y, n_splits, shuffle=True, random_state=None):
    """List of ``(train_idx, test_idx)`` exactly as ``StratifiedKFold.split`` yields."""
    tf = stratified_test_folds(y, n_splits, shuffle, random_state)
    idx = np.arange(len(tf))
    return [(idx[tf != i], idx[tf == i]) for i in range(n_splits)]


def kfold(n_samples, n_splits, shuffle=True, random_state=None):
    """``KFold(n_splits, shuffle=True).split`` replica."""
    indices = np.arange(n_samples)
    if shuffle:
        _rng(random_state).shuffle(indices)
    sizes = np.full(n_splits, n_samples // n_splits, dtype=int)
    sizes[:n_samples % n_splits] += 1
    out, cur = [], 0
    all_idx = np.arange(n_samples)
    for s in sizes:
        mask = np.zeros(n_samples, dtype=bool)
        mask[indices[cur:cur + s]] = True
        out.append((all_idx[~mask], all_idx[mask]))
        cur += s
    return out


def cv_splits(y, n_splits, shuffle=True, random_state=None):
    """The reference's try-stratified-else-plain policy (aligned_decode_svm_ncv.py:336-342)."""
    try:
        return stratified_kfold(y, n_splits, shuffle, random_state)
    except ValueError:
        return kfold(len(y), n_splits, shuffle, random_state)


# ----------------------------------------------------------------------------- labels
_L2S_CACHE = {}


def label2str(labels):
    """Reference alignment_utils.py:64-99: 2-D rows are joined digit strings, 1-D are str.
    (2-D integer labels are joined column-wise with numpy string ops and memoised on their bytes:
    a streamed job re-submits the same label arrays for every CV iteration.)"""
    labels = np.asarray(labels)
    if labels.ndim > 1:
        if labels.dtype.kind in 'iu' and labels.ndim == 2:
            key = (labels.shape, labels.dtype.str, labels.tobytes())
            out = _L2S_CACHE.get(key)
            if out is None:
                cols = labels.astype(str)
                out = cols[:, 0]
                for j in range(1, cols.shape[1]):
                    out = np.char.add(out, cols[:, j])
                if len(_L2S_CACHE) > 256:
                    _L2S_CACHE.clear()
                _L2S_CACHE[key] = out
            return out.copy()
        return np.array([''.join(str(x) for x in row) for row in labels])
    return labels.astype(str)


def class_ids(label_lists):
    """Maps every view's labels to integer ids over the sorted union of label strings.

    ``np.unique`` of strings sorts lexicographically ('10' < '2'), which is the class order
    the reference's ``cnd_avg`` uses (alignment_utils.py:57-60); ids preserve that order.
    Returns ``(ids_per_view, vocabulary)``."""
    strs = [label2str(l) for l in label_lists]
    vocab = np.unique(np.concatenate(strs)) if strs else np.array([], dtype=str)
    ids = [np.searchsorted(vocab, s).astype(np.int32) for s in strs]
    return ids, vocab
