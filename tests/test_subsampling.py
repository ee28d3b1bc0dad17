"""Electrode-subsampling front-end: the host index generators must reproduce the reference's
results BIT FOR BIT (golden file made by tests/golden/make_golden_subsampling.py from the
unmodified reference; the same numpy RNG call order under the same seeds); the device gather /
region mean are checked against numpy."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(HERE, 'golden', 'subsampling.npz'))


def test_grid_windows_bit_exact(gold):
    from cross_patient_speech_decoding_b200.processing_utils import grid_subsampling as gs
    for tag, win, step in (('a', (4, 8), (1, 1)), ('b', (3, 5), (2, 3)), ('c', (8, 16), (1, 1))):
        got = gs.grid_susbsample_idxs((8, 16), win, step=step)
        ref = gold['grid_%s' % tag]
        assert len(got) == len(ref)
        for a, b in zip(got, ref):
            assert a.dtype == b.dtype and np.array_equal(a, b)
        lst = gs.sig_channels_in_windows(gold['chanMap'], gold['sigChan'], win, step=step)
        assert len(lst) == int(gold['gridsig_%s_n' % tag])
        for i, a in enumerate(lst):
            assert np.array_equal(a, gold['gridsig_%s_%d' % (tag, i)])


def test_poisson_disk_bit_exact(gold):
    from cross_patient_speech_decoding_b200.processing_utils import poisson_disk_sampling as pds
    for n_elec in (10, 30, 64, 100):
        np.random.seed(100 + n_elec)
        spacing = np.floor(np.sqrt(8 * 16 / n_elec))
        pts = pds.poisson_disk_sampling((8, 16), spacing, n_elec)
        assert np.array_equal(pts, gold['pds_%d' % n_elec])          # exact floats, same order
    for pitch in (1.0, 1.5, 2.0, 3.0, 5.0):
        np.random.seed(int(pitch * 10))
        got = pds.pitch_sig_channels(gold['chanMap'], gold['sigChan'], pitch, 11.3, 22.5, 128)
        assert np.array_equal(got, gold['pitch_%s' % str(pitch).replace('.', 'p')])
    I, D = pds.knn_search(np.array([[0., 0.], [3., 4.], [1., 0.]]), np.array([[0., 0.]]), 2)
    assert list(I[0]) == [0, 2] and np.allclose(D[0], [0., 1.])


def test_spatial_average_regions_bit_exact(gold):
    from cross_patient_speech_decoding_b200.processing_utils import spatial_avg_subsampling as sa
    for cs in (2, 3, 4):
        got = sa.spatial_avg_idxs((8, 16), cs)
        assert np.array_equal(np.stack(got), gold['avg_%d' % cs])
        lst = sa.sig_regions(gold['chanMap'], cs, gold['sigChan'])
        assert len(lst) == int(gold['avgsig_%d_n' % cs])
        for i, a in enumerate(lst):
            assert np.array_equal(a, gold['avgsig_%d_%d' % (cs, i)])


@pytest.mark.skipif(not os.path.isdir('/root/reference/aligned_decoding'),
                    reason='live reference only in the build container')
def test_index_generators_against_live_reference():
    import sys
    import types
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = types.ModuleType('matplotlib')
        mpl.pyplot = types.ModuleType('matplotlib.pyplot')
        sys.modules.setdefault('matplotlib', mpl)
        sys.modules.setdefault('matplotlib.pyplot', mpl.pyplot)
    sys.path.insert(0, '/root/reference/aligned_decoding')
    from processing_utils import grid_subsampling as rg, poisson_disk_sampling as rp
    from cross_patient_speech_decoding_b200.processing_utils import grid_subsampling as gs
    from cross_patient_speech_decoding_b200.processing_utils import poisson_disk_sampling as pds
    rng = np.random.default_rng(0)
    for _ in range(6):
        g = (int(rng.integers(4, 13)), int(rng.integers(6, 25)))
        w = (int(rng.integers(1, g[0] + 1)), int(rng.integers(1, g[1] + 1)))
        st = (int(rng.integers(1, 4)), int(rng.integers(1, 4)))
        a, b = gs.grid_susbsample_idxs(g, w, st), rg.grid_susbsample_idxs(g, w, st)
        assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))
    for seed in range(5):
        dom = (int(rng.integers(6, 13)), int(rng.integers(8, 25)))
        n = int(rng.integers(5, dom[0] * dom[1] // 2))
        spacing = np.floor(np.sqrt(dom[0] * dom[1] / n))
        np.random.seed(seed)
        x = pds.poisson_disk_sampling(dom, spacing, n)
        np.random.seed(seed)
        y = rp.poisson_disk_sampling(dom, spacing, n)
        assert np.array_equal(x, y)


@pytest.mark.gpu
def test_device_gather_and_region_mean(lib_built, gold):
    import torch
    from cross_patient_speech_decoding_b200.processing_utils import device_subsample as ds
    from cross_patient_speech_decoding_b200.processing_utils import spatial_avg_subsampling as sa
    rng = np.random.default_rng(1)
    X = rng.standard_normal((20, 30, 48))
    Xd = ds.resident(X)
    assert Xd.is_cuda and Xd.dtype == torch.float32
    idx = np.array([5, 0, 47, 12, 12, 33])
    sub = ds.gather_channels(Xd, idx).cpu().numpy()
    assert np.array_equal(sub, X.astype(np.float32)[:, :, idx])
    out = sa.spatial_avg_data(gold['avg_data_in'], sa.spatial_avg_idxs((8, 16), 3))
    assert out.shape == gold['avg_data_out'].shape
    assert np.abs(out - gold['avg_data_out']).max() < 1e-12


@pytest.mark.gpu
def test_subsampled_resident_views_match_host_slices(lib_built):
    """A CV job on device-gathered channel subsets == the same job on host-sliced arrays."""
    import sys
    sys.path.insert(0, os.path.join(HERE, 'golden'))
    import make_golden
    from cross_patient_speech_decoding_b200 import cv_align_decode
    from cross_patient_speech_decoding_b200.processing_utils import device_subsample as ds
    from cross_patient_speech_decoding_b200.processing_utils import grid_subsampling as gs
    cfg = make_golden.CONFIGS['mcca_p3_ragged']
    pts, folds = make_golden.build_inputs(cfg)
    kw = dict(method='mcca', n_comp=8, regs=0.5, pca_var=0.8)
    wins = gs.grid_susbsample_idxs((4, 8), (3, 6))      # channel windows on a 4 x 8 layout
    chans = [np.sort((w[:, 0] * 8 + w[:, 1])) for w in wins[:2]]
    res_d = [ds.resident(p[0]) for p in pts]
    for ch in chans:
        host = [(p[0][:, :, ch], p[1], p[2]) for p in pts]
        dev = [(ds.gather_channels(r, ch), p[1], p[2]) for r, p in zip(res_d, pts)]
        a = cv_align_decode(host[0], host[1:], folds, **kw)
        b = cv_align_decode(dev[0], dev[1:], folds, **kw)
        assert a['k2'] == b['k2']
        for x, y in zip(a['y_pred'], b['y_pred']):
            assert np.array_equal(x, y)
        assert b['h2d_bytes'] < a['h2d_bytes']


@pytest.mark.gpu
def test_subsample_decode_matches_per_subsample_port(lib_built):
    """processing_utils.subsample_decode (streamed jobs, device-side channel gathers, the
    scripts' RNG order) against the CPU port run on host-sliced copies with the same choices."""
    from sklearn.model_selection import StratifiedKFold
    from cross_patient_speech_decoding_b200 import synthetic
    from cross_patient_speech_decoding_b200.processing_utils.subsample_decode import subsample_decode
    from oracle import pipeline_port as port
    pts = [synthetic.make_patient(p, n_trials=n, n_time=20, n_chan=32, noise=0.6)
           for p, n in enumerate((72, 80, 64))]
    rng = np.random.default_rng(3)
    tar_list = [np.sort(rng.choice(32, 12, replace=False)) for _ in range(3)]
    cross_lists = [[np.sort(rng.choice(32, 12, replace=False)) for _ in range(4)] for _ in range(2)]
    np.random.seed(9)
    out = subsample_decode(pts[0], pts[1:], tar_list, cross_lists, n_folds=4, method='cca',
                           n_comp=0.9, decoder='svc_rbf', class_weight='balanced', depth=2)
    # replay the reference loop's RNG stream on the host
    np.random.seed(9)
    lab = pts[0][1]
    same = tot = 0
    for j, sub in enumerate(tar_list):
        chosen = [int(np.random.choice(len(c))) for c in cross_lists]
        assert chosen == out['chosen'][j]
        splits = list(StratifiedKFold(n_splits=4, shuffle=True).split(np.zeros((len(lab), 1)), lab))
        tar = (pts[0][0][:, :, sub], pts[0][1], pts[0][2])
        cross = [(pts[p + 1][0][:, :, cross_lists[p][chosen[p]]], pts[p + 1][1], pts[p + 1][2])
                 for p in range(2)]
        yp = []
        for tr, te in splits:
            ref, _ = port.run_fold(tar, cross, tr, te, method='cca', n_comp=0.9, decoder='svc_rbf',
                                   class_weight='balanced')     # sklearn's SVC.fit draws its seed here
            yp.append(ref)
        yp = np.concatenate(yp)
        assert np.array_equal(np.concatenate([lab[te] for _, te in splits]), np.array(out['y_true'][j]))
        same += int((yp == np.array(out['y_pred'][j])).sum())
        tot += len(yp)
    assert same / tot >= 0.99, (same, tot)


@pytest.mark.gpu
def test_trial_count_sweep_matches_port(lib_built):
    """processing_utils.trial_subsample_decode (the batched loops of
    scripts/aligned_decode_cross_patient_subsample.py:290-381: per (k, iteration) a random k-trial
    subset of every cross patient, a shuffled split, one decoder per fold) against the CPU port fed
    with the same np.random stream: identical trial picks and folds, labels and the accuracy matrix."""
    from sklearn.metrics import balanced_accuracy_score
    from sklearn.model_selection import StratifiedKFold
    from cross_patient_speech_decoding_b200 import synthetic
    from cross_patient_speech_decoding_b200.processing_utils import device_subsample as ds
    from cross_patient_speech_decoding_b200.processing_utils.trial_subsample_decode import (
        trial_counts, trial_subsample_decode)
    from oracle import pipeline_port as port
    pts = [synthetic.make_patient(p, n_trials=n, n_time=20, n_chan=24, noise=0.5)
           for p, n in enumerate((72, 90, 50))]
    assert list(trial_counts(pts[1:], 25)) == [5, 30, 55]       # arange(5, ceil(median(90, 50)) + 1, 25)
    # the device row gather itself
    Xd = ds.resident(pts[1][0])
    idx = np.array([7, 0, 89, 33, 12])
    assert np.array_equal(ds.gather_trials(Xd, idx).cpu().numpy(), pts[1][0].astype(np.float32)[idx])
    k_list, n_iter, n_folds = [20, 60], 2, 3
    np.random.seed(21)
    out = trial_subsample_decode(pts[0], pts[1:], k_list=k_list, n_iter=n_iter, n_folds=n_folds,
                                 method='cca', n_comp=0.9, decoder='svc_rbf', class_weight='balanced', depth=3)
    np.random.seed(21)
    lab = pts[0][1]
    same = tot = 0
    for ki, k in enumerate(k_list):
        for it in range(n_iter):
            cross = []
            for X, y, ya in pts[1:]:
                if X.shape[0] < k:
                    cross.append((X, y, ya))
                else:
                    s = np.random.choice(X.shape[0], k, replace=False)
                    cross.append((X[s], y[s], ya[s]))
            splits = list(StratifiedKFold(n_splits=n_folds, shuffle=True).split(np.zeros((len(lab), 1)), lab))
            yp = np.concatenate([port.run_fold(pts[0], cross, tr, te, method='cca', n_comp=0.9,
                                               decoder='svc_rbf', class_weight='balanced')[0]
                                 for tr, te in splits])        # SVC.fit draws libsvm's seed per fold
            yt = np.concatenate([lab[te] for _, te in splits])
            j = ki * n_iter + it
            assert np.array_equal(yt, np.array(out['y_true'][j]))
            got = np.array(out['y_pred'][j])
            same += int((got == yp).sum())
            tot += len(yp)
            assert abs(out['acc_mat'][ki, it] - balanced_accuracy_score(yt, got)) < 1e-12
        assert out['trial_vec'][ki] == sum(min(k, p[0].shape[0]) if p[0].shape[0] >= k else p[0].shape[0]
                                           for p in pts[1:])
    assert same / tot >= 0.99, (same, tot)
