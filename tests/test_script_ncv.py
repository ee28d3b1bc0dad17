"""scripts/aligned_decode_svm_ncv.py (batched) against the golden output of the reference's own
script run unmodified (tests/golden/make_golden_script.py -> script_ncv_cca.npz)."""
import os
import pickle
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))


def _golden():
    return np.load(os.path.join(HERE, 'golden', 'script_ncv_cca.npz'))


def _write_data(tmp_path):
    import make_golden_script as mg
    f = tmp_path / 'pt_decoding_data_S62.pkl'
    with open(f, 'wb') as fh:
        pickle.dump(mg.data_dict(), fh, protocol=-1)
    return str(f)


def test_units_consume_rng_like_reference_script():
    """The test trials of every (iteration, fold) come out in the reference script's order:
    y_true of the golden run is reproduced bit for bit from the seed (no GPU needed)."""
    import make_golden_script as mg
    from cross_patient_speech_decoding_b200.scripts import aligned_decode_svm_ncv as sc
    g = _golden()
    lab = np.asarray(mg.data_dict()['S1']['y1'])
    np.random.seed(int(g['seed']))
    units = sc.make_units(lab, 50, 20, 1.0)
    assert len(units) == 1000
    for j in (0, 1, 17, 49):
        yt = np.concatenate([lab[te] for _, te in units[j * 20:(j + 1) * 20]])
        assert np.array_equal(yt, g['y_true'][j])


def test_trial_subsample_units_match_sklearn_stream():
    """--trial_subsample < 1: one stratified train_test_split per fold, drawn after the
    iteration's split (aligned_decode_svm_ncv.py:336-362)."""
    from sklearn.model_selection import StratifiedKFold, train_test_split
    from cross_patient_speech_decoding_b200.scripts import aligned_decode_svm_ncv as sc
    rng = np.random.default_rng(0)
    lab = rng.integers(1, 5, 80)
    np.random.seed(5)
    units = sc.make_units(lab, 2, 4, 0.5)
    np.random.seed(5)
    k = 0
    for _ in range(2):
        splits = list(StratifiedKFold(n_splits=4, shuffle=True).split(np.zeros((80, 1)), lab))
        for tr, te in splits:
            X = np.arange(80)[tr]
            a, _, b, _ = train_test_split(X, lab[tr], train_size=0.5, stratify=lab[tr], shuffle=True)
            np.random.randint(np.iinfo('i').max)        # SVC.fit's seed draw (sklearn/svm/_base.py)
            assert np.array_equal(units[k][0], a) and np.array_equal(units[k][1], te)
            k += 1


@pytest.mark.gpu
def test_script_matches_reference_script_golden(lib_built, tmp_path):
    from cross_patient_speech_decoding_b200.scripts import aligned_decode_svm_ncv as sc
    g = _golden()
    out_file = str(tmp_path / 'out.pkl')
    res = sc.aligned_decoding(['-pt', 'S1', '-pi', '1', '-po', 'True', '-a', 'True', '-c', 'False',
                               '-f', out_file, '--data_file', _write_data(tmp_path),
                               '--seed', str(int(g['seed']))])
    with open(out_file, 'rb') as fh:
        saved = pickle.load(fh)
    assert sorted(saved['params'].keys()) == list(g['param_keys'])
    assert saved['params']['n_iter'] == 50 and saved['params']['n_folds'] == 20
    yt, yp = np.array(res['y_true']), np.array(res['y_pred'])
    assert np.array_equal(yt, g['y_true'])                      # same trials in the same order
    assert np.mean(yp == g['y_pred']) >= 0.99                   # north-star label bar
    assert np.abs(np.array(res['accs']) - g['accs']).mean() <= 0.005
    assert abs(np.mean(res['accs']) - g['accs'].mean()) <= 0.005
    assert np.abs(np.array([len(w) for w in res['wrong_trs']]) - g['n_wrong']).max() <= 3


@pytest.mark.gpu
def test_script_joint_branch_matches_port(lib_built, tmp_path):
    """-j True: the script's set_params hands n_comp = 0.9 (a variance fraction) to JointPCA
    (aligned_decode_svm_ncv.py:186-190, 372-375, 416); checked fold by fold against the CPU port."""
    import make_golden_script as mg
    from cross_patient_speech_decoding_b200.scripts import aligned_decode_svm_ncv as sc
    from oracle import pipeline_port as port
    d = mg.data_dict()
    res = sc.aligned_decoding(['-pt', 'S1', '-pi', '1', '-po', 'True', '-j', 'True', '-c', 'False',
                               '-f', str(tmp_path / 'out.pkl'), '--data_file', _write_data(tmp_path),
                               '--seed', '3', '--n_iter', '1', '--n_folds', '3', '--decoder', 'linear'])
    tar = (d['S1']['X1'], np.asarray(d['S1']['y1']), d['S1']['y_full_phon'])
    cross = [(d[p]['X1'], np.asarray(d[p]['y1']), d[p]['y_full_phon']) for p in d['S1']['pre_pts']]
    np.random.seed(3)
    units = sc.make_units(tar[1], 1, 3, 1.0)
    ref = np.concatenate([port.run_fold(tar, cross, tr, te, method='jointpca', n_comp=0.9)[0]
                          for tr, te in units])
    got = np.array(res['y_pred'][0])
    assert got.shape == ref.shape and np.mean(got == ref) >= 0.97, float(np.mean(got == ref))


@pytest.mark.gpu
def test_script_nested_bayes_search(lib_built, tmp_path):
    """-cv True (aligned_decode_svm_ncv.py:388-405): per outer fold a Bayesian search over the
    script's space scored on inner folds of the outer-train trials, then one fit / predict with the
    winner.  Checked for consistency: winners lie in the space, and the outer predictions equal a
    plain cv_align_decode call with each unit's winning parameters."""
    import make_golden_script as mg
    from cross_patient_speech_decoding_b200 import cv_align_decode
    from cross_patient_speech_decoding_b200.scripts import aligned_decode_svm_ncv as sc
    from cross_patient_speech_decoding_b200.search import engine_keywords
    res = sc.aligned_decoding(['-pt', 'S1', '-pi', '1', '-po', 'True', '-a', 'True', '-c', 'False',
                               '-cv', 'True', '-f', str(tmp_path / 'out.pkl'), '--data_file',
                               _write_data(tmp_path), '--seed', '5', '--n_iter', '1', '--n_folds', '3',
                               '--decoder', 'linear', '--search_iter', '6', '--search_points', '3'])
    best = res['params']['best_params']
    assert len(best) == 3
    for b in best:
        assert set(b) == {'n_comp', 'decoder__dimredreshape__n_components'}
        assert 0.1 <= b['n_comp'] <= 0.95 and 0.1 <= b['decoder__dimredreshape__n_components'] <= 0.95
    d = mg.data_dict()
    tar = (d['S1']['X1'], np.asarray(d['S1']['y1']), d['S1']['y_full_phon'])
    cross = [(d[p]['X1'], np.asarray(d[p]['y1']), d[p]['y_full_phon']) for p in d['S1']['pre_pts']]
    np.random.seed(5)
    units = sc.make_units(tar[1], 1, 3, 1.0)
    got = np.array(res['y_pred'][0])
    want = np.concatenate([cv_align_decode(tar, cross, [u], method='cca', decoder='linear',
                                           **engine_keywords(b))['y_pred'][0] for u, b in zip(units, best)])
    assert np.array_equal(got, want)
    assert 0.0 <= res['accs'][0] <= 1.0
