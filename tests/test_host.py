"""CPU-side checks: fold-index generation is bit-exact against sklearn, golden fold indices
are reproducible from seeds, and the C-ABI library loads and exports every symbol that
include/cpsd_b200.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re
import sys
import warnings

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, 'golden'))


def test_stratified_kfold_bit_exact():
    from sklearn.model_selection import KFold, StratifiedKFold
    from cross_patient_speech_decoding_b200.folds import cv_splits, kfold, stratified_kfold
    rng = np.random.default_rng(0)
    warnings.simplefilter('ignore')
    for trial in range(30):
        n = int(rng.integers(30, 200))
        y = rng.integers(1, 10, n)
        ns = int(rng.integers(2, 21))
        np.random.seed(trial)
        try:
            ref = list(StratifiedKFold(ns, shuffle=True).split(np.zeros(n), y))
        except ValueError:
            np.random.seed(trial)
            ref = list(KFold(ns, shuffle=True).split(np.zeros(n)))
        np.random.seed(trial)
        got = cv_splits(y, ns)
        assert len(ref) == len(got)
        for (a, b), (c, d) in zip(ref, got):
            assert np.array_equal(a, c) and np.array_equal(b, d)
    # explicit RandomState and int seeds
    y = rng.integers(0, 5, 100)
    ref = list(StratifiedKFold(5, shuffle=True, random_state=3).split(np.zeros(100), y))
    got = stratified_kfold(y, 5, random_state=3)
    assert all(np.array_equal(a[1], b[1]) for a, b in zip(ref, got))
    ref = list(KFold(7, shuffle=True, random_state=4).split(np.zeros(100)))
    got = kfold(100, 7, random_state=4)
    assert all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(ref, got))


def test_label_strings_and_class_order():
    from cross_patient_speech_decoding_b200.folds import class_ids, label2str
    lab = np.array([[1, 2, 3], [10, 1, 1], [2, 9, 9]])
    assert list(label2str(lab)) == ['123', '1011', '299']
    assert list(label2str(np.array([3, 10, 2]))) == ['3', '10', '2']
    ids, vocab = class_ids([np.array([10, 2, 3]), np.array([2, 2, 11])])
    assert list(vocab) == ['10', '11', '2', '3']          # lexicographic, as np.unique of str
    assert list(ids[0]) == [0, 2, 3] and list(ids[1]) == [2, 2, 1]


def test_golden_fold_indices_reproducible():
    import make_golden
    for name, cfg in make_golden.CONFIGS.items():
        path = os.path.join(HERE, 'golden', name + '.npz')
        if not os.path.exists(path):
            pytest.skip('golden file missing: ' + name)
        g = np.load(path)
        pts = [(None, None, None)]
        # only labels are needed: regenerate patient 0 cheaply
        from cross_patient_speech_decoding_b200 import synthetic
        kw = dict(cfg['patients'][cfg.get('target', 0)])
        kw.update(n_time=2, n_chan=2)
        _, y, _ = synthetic.make_patient(**kw)
        from cross_patient_speech_decoding_b200.folds import cv_splits
        np.random.seed(cfg['seed'])
        folds = []
        for _ in range(cfg.get('n_iter', 1)):
            folds += cv_splits(y, cfg['n_splits'])
        for f in range(int(g['n_folds'])):
            assert np.array_equal(folds[f][0], g['train_%d' % f])
            assert np.array_equal(folds[f][1], g['test_%d' % f])


def test_synthetic_labels_do_not_depend_on_shape():
    from cross_patient_speech_decoding_b200 import synthetic
    _, y1, ya1 = synthetic.make_patient(0, n_time=2, n_chan=2)
    _, y2, ya2 = synthetic.make_patient(0, n_time=5, n_chan=3)
    assert np.array_equal(y1, y2) and np.array_equal(ya1, ya2)


def test_library_exports_header_symbols(lib_built):
    hdr = open(os.path.join(ROOT, 'include', 'cpsd_b200.h')).read()
    names = sorted(set(re.findall(r'\b(cpsd_[a-z0-9_]+)\s*\(', hdr)))
    assert len(names) >= 25
    lib = ctypes.CDLL(lib_built)
    for n in names:
        assert hasattr(lib, n), 'missing export ' + n
    from cross_patient_speech_decoding_b200 import _lib
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), set(names) ^ set(_lib.EXPORTED_SYMBOLS)
    handle = _lib.load()      # also checks descriptor sizes against descs.h
    assert handle.cpsd_version() >= 100


def test_block_jacobi_schedule_is_a_tournament(lib_built):
    lib = ctypes.CDLL(lib_built)
    for n_pad in (256, 384, 1152):
        nb = n_pad // 64
        out = np.zeros((nb - 1) * (nb // 2) * 2, dtype=np.int32)
        assert lib.cpsd_bj_schedule(n_pad, out.ctypes.data_as(ctypes.c_void_p)) == 0
        pairs = out.reshape(nb - 1, nb // 2, 2)
        seen = set()
        for r in range(nb - 1):
            assert sorted(pairs[r].ravel().tolist()) == list(range(nb))   # disjoint cover
            for a, b in pairs[r]:
                assert a < b
                seen.add((int(a), int(b)))
        assert len(seen) == nb * (nb - 1) // 2                            # every pair once


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from cross_patient_speech_decoding_b200 import _lib
    from cross_patient_speech_decoding_b200.device import Context
    with pytest.raises(_lib.CpsdError):
        Context(None)


def test_ctypes_signatures_match_header():
    """Every prototype of include/cpsd_b200.h against the ctypes argtypes the Python layer binds
    (argument count and class: pointer / int / long long / float / double)."""
    import ctypes
    import re
    from cross_patient_speech_decoding_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'cpsd_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', ' ', hdr, flags=re.S)
    protos = re.findall(r'\b(?:int|long long|const char\s*\*|void)\s+(cpsd_\w+)\s*\(([^;{]*?)\)\s*;', hdr, flags=re.S)
    assert len(protos) >= 60

    def kind(param):
        p = ' '.join(param.split())
        if '*' in p or 'cudaStream_t' in p:
            return ctypes.c_void_p
        if p.startswith('long long') or p.startswith('const long long'):
            return ctypes.c_longlong
        if p.startswith('double'):
            return ctypes.c_double
        if p.startswith('float'):
            return ctypes.c_float
        assert p.startswith('int') or p.startswith('const int'), p
        return ctypes.c_int

    checked = 0
    for name, params in protos:
        if name not in _lib._SIGS:
            continue
        plist = [] if params.strip() in ('', 'void') else [kind(x) for x in params.split(',')]
        want = list(_lib._SIGS[name])
        assert len(plist) == len(want), (name, len(plist), len(want))
        for i, (a, b) in enumerate(zip(plist, want)):
            assert a is b, (name, i, a, b)
        checked += 1
    assert checked >= 55


def test_bayes_search_space_and_proposals():
    """search.BayesSearch (the BayesSearchCV stand-in of scripts/aligned_decode_svm_ncv.py:388-394):
    skopt's space notation, reproducible proposals under a seed, no repeated candidates, and a
    better optimum than the random start on a smooth objective."""
    from cross_patient_speech_decoding_b200.search import BayesSearch, Dimension, engine_keywords
    assert Dimension((10, 50)).kind == 'int' and Dimension((0.1, 0.95, 'uniform')).kind == 'real'
    assert Dimension((1e-3, 1e5, 'log-uniform')).kind == 'log'
    g = Dimension(np.arange(0.1, 1, 0.1))
    assert g.kind == 'grid' and g.from_unit(0.0) == pytest.approx(0.1) and g.from_unit(1.0) == pytest.approx(0.9)
    d = Dimension((1e-3, 1e5, 'log-uniform'))
    assert d.from_unit(d.to_unit(3.7)) == pytest.approx(3.7)
    space = {'n_comp': (10, 50), 'pca_var': (0.1, 0.95, 'uniform'),
             'decoder__dimredreshape__n_components': (0.1, 0.95, 'uniform')}

    def score(p):     # smooth, maximum 1.0 at (32, 0.7, 0.4)
        return 1.0 - ((p['n_comp'] - 32) / 40.0) ** 2 - (p['pca_var'] - 0.7) ** 2 - \
            (p['decoder__dimredreshape__n_components'] - 0.4) ** 2

    def run(seed):
        opt = BayesSearch(space, n_initial_points=10, random_state=seed)
        hist = []
        for _ in range(5):
            c = opt.ask(5)
            assert all(isinstance(p['n_comp'], int) and 10 <= p['n_comp'] <= 50 for p in c)
            assert all(0.1 <= p['pca_var'] <= 0.95 for p in c)
            opt.tell(c, [score(p) for p in c])
            hist += c
        return opt, hist

    a, ha = run(3)
    b, hb = run(3)
    assert ha == hb                                             # reproducible
    keys = {tuple(sorted(p.items())) for p in ha}
    assert len(keys) == 25                                      # no candidate proposed twice
    assert max(a.y[10:]) > max(a.y[:10])                        # the surrogate rounds improve on the random start
    assert max(a.y) > 0.97
    assert engine_keywords(ha[0]) == {'n_comp': ha[0]['n_comp'], 'pca_var': ha[0]['pca_var'],
                                      'decoder_var': ha[0]['decoder__dimredreshape__n_components']}
    with pytest.raises(ValueError):
        engine_keywords({'decoder__baggingclassifier__n_estimators': 10})


def test_warm_start_plan_of_view_solves():
    """Host logic of the warm-started view solves (engine.plan_view_solves): the first problem of
    every (replica, patient) pair without a basis is cold and founds the pair's basis, everything
    else is warm from its pair's basis; bases persist across batches."""
    from cross_patient_speech_decoding_b200.engine import plan_view_solves
    P = 3
    tab = -np.ones(2 * P, dtype=np.int32)
    # batch 1: replica 0 targets (pair 0) x3, replica 1 target (pair 3), cross problems of pairs 1, 2, 1, 4
    code = np.array([0, 0, 0, 3, 1, 2, 1, 4])
    cold, warm, base_of, tab1 = plan_view_solves(code, tab, 0)
    assert sorted(cold.tolist()) == [0, 3, 4, 5, 7]            # first problem of pairs 0, 3, 1, 2, 4
    assert warm.tolist() == [1, 2, 6]
    assert (tab == -1).all()                                    # the caller's table is not touched
    assert tab1.tolist() == [0, 1, 2, 3, 4, -1]                 # bases numbered in pair order
    assert base_of.tolist() == [-1, 0, 0, -1, -1, -1, 1, -1]
    # the base founded by a cold problem has the index its pair got
    assert [int(tab1[code[i]]) for i in cold] == [0, 1, 2, 3, 4]   # cold problem i founds base n_bases + i
    # batch 2: everything known except pair 5
    code2 = np.array([0, 3, 5, 5, 2])
    cold2, warm2, base2, tab2 = plan_view_solves(code2, tab1, 5)
    assert cold2.tolist() == [2] and warm2.tolist() == [0, 1, 3, 4]
    assert tab2.tolist() == [0, 1, 2, 3, 4, 5] and base2.tolist() == [0, 3, -1, 5, 2]
    # nothing new: no copy needed, nothing cold
    cold3, warm3, base3, tab3 = plan_view_solves(code2, tab2, 6)
    assert len(cold3) == 0 and warm3.tolist() == [0, 1, 2, 3, 4] and tab3 is tab2
