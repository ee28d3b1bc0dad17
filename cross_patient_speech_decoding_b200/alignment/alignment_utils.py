"""Drop-in for the reference's ``alignment/alignment_utils.py`` (same function names and
argument meaning).  Condition averaging runs on the GPU (``cpsd_class_mean``); label
handling, pickling and the data-dictionary unpacking are host bookkeeping.
"""
import pickle
from functools import reduce

import numpy as np

from .. import ops
from ..folds import label2str as _label2str


def label2str(labels):
    """Labels -> 1-D array of strings (reference alignment_utils.py:64-80)."""
    return _label2str(labels)


def label_seq2str(labels):
    """(n_trials, n_labels) integer sequences -> joined strings (alignment_utils.py:83-99)."""
    labels = np.asarray(labels)
    return np.array([''.join(str(v) for v in labels[i, :]) for i in range(labels.shape[0])])


def cnd_avg(data, labels):
    """Per-class mean over the first axis, classes in ``np.unique(labels)`` order
    (alignment_utils.py:42-61).  Returns float64 like the reference."""
    data = np.asarray(data)
    classes, inv = np.unique(np.asarray(labels), return_inverse=True)
    _, means = ops.class_mean(data, inv)
    return means.astype(np.float64)


def extract_group_conditions(Xs, ys):
    """Class averages of every dataset restricted to the classes all share
    (alignment_utils.py:12-39)."""
    ys = [label2str(np.asarray(l)) for l in ys]
    avgs = [cnd_avg(X, l) for X, l in zip(Xs, ys)]
    shared = reduce(np.intersect1d, ys)
    return [a[np.isin(np.unique(l), shared, assume_unique=True)] for a, l in zip(avgs, ys)]


def save_pkl(data, filename):
    with open(filename, 'wb+') as f:
        pickle.dump(data, f, protocol=-1)


def load_pkl(filename):
    with open(filename, 'rb') as f:
        return pickle.load(f)


def phon_to_artic(phon_idx, phon_to_artic_conv):
    return phon_to_artic_conv[phon_idx]


def phon_to_artic_seq(phon_seq):
    """Phoneme ids 1..9 -> articulator ids 1..4 (alignment_utils.py:187-201)."""
    conv = {1: 1, 2: 1, 3: 2, 4: 2, 5: 3, 6: 3, 7: 3, 8: 4, 9: 4}
    phon_seq = np.asarray(phon_seq)
    return np.array([conv[int(p)] for p in phon_seq.flatten()]).reshape(phon_seq.shape)


def get_features_labels(data, p_ind, lab_type, algn_type):
    """One patient's ``(D, lab, lab_full)`` from the data dictionary
    (alignment_utils.py:160-184; schema in SURVEY.md Appendix D)."""
    lab_full = data['y_full_' + algn_type[:-4]]
    if p_ind == -1:
        D = data['X_collapsed']
        lab = data['y_' + lab_type + '_collapsed']
        lab_full = np.tile(lab_full, (3, 1))
    else:
        D = data['X' + str(p_ind)]
        lab = data['y' + str(p_ind)]
    if lab_type == 'artic':
        lab = phon_to_artic_seq(lab)
    return D, lab, lab_full


def decoding_data_from_dict(data_dict, pt, p_ind, lab_type='phon', algn_type='phon_seq'):
    """Target triple + list of pre-training patients' triples (alignment_utils.py:127-157)."""
    tar = get_features_labels(data_dict[pt], p_ind, lab_type, algn_type)
    pre = [get_features_labels(data_dict[p], p_ind, lab_type, algn_type)
           for p in data_dict[pt]['pre_pts']]
    return tar, pre
