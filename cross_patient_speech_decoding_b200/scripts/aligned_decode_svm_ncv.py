"""Batched counterpart of the reference's ``scripts/aligned_decode_svm_ncv.py``.

Same command line (``-pt -pi -po -t -a -m -j -r -tss -pp -c -f -s``), same data-dictionary
pickle in (``alignment_utils.load_pkl`` / ``decoding_data_from_dict``,
aligned_decode_svm_ncv.py:268-275), same result pickle out
(``{'params', 'y_true', 'y_pred', 'wrong_trs', 'accs'}``, :289-296, :443-456) -- but the
``n_iter x n_folds`` fit/predict units of the two nested loops (:332-442) are generated up
front (the numpy global RNG is consumed in the reference's order: one shuffled
StratifiedKFold per iteration, then one stratified ``train_test_split`` per fold when
``--trial_subsample < 1``) and handed to ``cv_align_decode`` as ONE batch.

``-cv True`` (the nested ``BayesSearchCV`` of :388-405; scikit-optimize is absent) runs
``search.bayes_search_align_decode`` per outer fold: a Gaussian-process / expected-improvement
proposal loop whose candidates x inner folds go to the GPU as one batch per round.
Extra options: ``--data_file``, ``--n_iter``, ``--n_folds``, ``--decoder``, ``--seed``,
``--search_iter``, ``--search_points``.

Multi-GPU: launched under ``torchrun --nproc-per-node N -m
cross_patient_speech_decoding_b200.scripts.aligned_decode_svm_ncv ...`` every rank generates
the same unit list (same RNG stream), runs the CV iterations ``sharding.shard_units`` gives it
on its own GPU with no data-path communication, and one all_gather of the predicted labels
(``sharding.gather_predictions``) gives every rank the full result; rank 0 writes the pickle,
which is identical to the single-GPU one.
"""
import argparse
import os

import numpy as np

from .. import cv_align_decode
from .. import sharding
from ..alignment import alignment_utils as utils
from ..folds import cv_splits


def init_parser():
    ap = argparse.ArgumentParser(description='Aligned decoding SVM (batched, B200)')
    ap.add_argument('-pt', '--patient', type=str, required=True, help='Patient ID')
    ap.add_argument('-pi', '--p_ind', type=int, default=-1, help='Sequence position index')
    ap.add_argument('-po', '--pool_train', type=str, default='False')
    ap.add_argument('-t', '--tar_in_train', type=str, default='True')
    ap.add_argument('-a', '--cca_align', type=str, default='False')
    ap.add_argument('-m', '--MCCA_align', type=str, default='False')
    ap.add_argument('-j', '--joint_dim_red', type=str, default='False')
    ap.add_argument('-r', '--random_data', type=str, default='False')
    ap.add_argument('-n', '--no_S23', type=str, default='False')
    ap.add_argument('-tss', '--trial_subsample', type=float, default=1.0)
    ap.add_argument('-surr', '--surrogate', type=str, default='False')
    ap.add_argument('-pp', '--pooled_patients', type=str, default='all')
    ap.add_argument('-c', '--cluster', type=str, default='True')
    ap.add_argument('-cv', '--cross_validate', type=str, default='False')
    ap.add_argument('-f', '--filename', type=str, default='')
    ap.add_argument('-s', '--suffix', type=str, default='')
    # not in the reference (its values are constants in the script body)
    ap.add_argument('--data_file', type=str, default='')
    ap.add_argument('--n_iter', type=int, default=50)
    ap.add_argument('--n_folds', type=int, default=20)
    ap.add_argument('--decoder', type=str, default='svc_rbf',
                    choices=['svc_rbf', 'svc_linear', 'linear'],
                    help="svc_rbf = the script's SVC(kernel='rbf', class_weight='balanced')")
    ap.add_argument('--seed', type=int, default=None, help='np.random.seed before the loops')
    ap.add_argument('--search_iter', type=int, default=25, help='BayesSearchCV n_iter (-cv True)')
    ap.add_argument('--search_points', type=int, default=5, help='BayesSearchCV n_points (-cv True)')
    return ap


def str2bool(s):
    return s.lower() == 'true'


def make_units(lab_tar, n_iter, n_folds, tr_subsamp_r, fit_draws=1):
    """The (train_idx, test_idx) units of every iteration, drawing from the numpy global RNG in
    the order of aligned_decode_svm_ncv.py:332-442: per iteration one shuffled StratifiedKFold;
    per fold the optional stratified ``train_test_split`` and then the draw sklearn's
    ``SVC.fit`` makes for libsvm's seed (``check_random_state(None).randint(INT_MAX)``,
    sklearn/svm/_base.py) -- ``fit_draws`` of them, so that the folds of iteration j+1 are the
    ones the reference script would have produced."""
    from sklearn.model_selection import train_test_split
    units = []
    for _ in range(n_iter):
        for train_idx, test_idx in cv_splits(lab_tar, n_folds):
            if tr_subsamp_r < 1:
                train_idx, _ = train_test_split(train_idx, train_size=tr_subsamp_r,
                                                stratify=lab_tar[train_idx], shuffle=True)
            for _d in range(fit_draws):
                np.random.randint(np.iinfo('i').max)
            units.append((np.asarray(train_idx), np.asarray(test_idx)))
    return units


def run(inputs, run_units=None):
    """``run_units(target, cross, units, **kw) -> {'y_pred': [...]}`` defaults to
    ``cv_align_decode`` (tests inject a CPU stand-in to exercise the sharding on gloo)."""
    from sklearn.metrics import balanced_accuracy_score
    rank, world, _ = sharding.init_from_env()
    run_units = run_units or cv_align_decode
    cluster = str2bool(inputs['cluster'])
    if cluster:
        data_path, out_path = os.path.expanduser('~') + '/data/', os.path.expanduser('~') + '/workspace/'
    else:
        data_path, out_path = '../data/', '../acc_data/'
    pt, p_ind = inputs['patient'], inputs['p_ind']
    pool_train, tar_in_train = str2bool(inputs['pool_train']), str2bool(inputs['tar_in_train'])
    cca_align, mcca_align = str2bool(inputs['cca_align']), str2bool(inputs['MCCA_align'])
    joint_dim_red = str2bool(inputs['joint_dim_red'])
    do_cv = str2bool(inputs['cross_validate'])
    n_iter, n_folds = inputs['n_iter'], inputs['n_folds']
    if sum([cca_align, mcca_align, joint_dim_red]) > 1:      # the reference's precedence (:218-222)
        cca_align = mcca_align = False
    if do_cv:
        # search spaces of aligned_decode_svm_ncv.py:149-176.  The reference's MCCA grid also lists
        # three 'decoder__baggingclassifier__*' keys, but its checked-in decoder pipeline has no
        # BaggingClassifier step (:299-317), so set_params would reject them: they are left out.
        if mcca_align:
            param_grid = {'n_comp': (10, 50), 'pca_var': (0.1, 0.95, 'uniform'),
                          'decoder__dimredreshape__n_components': (0.1, 0.95, 'uniform')}
        else:
            param_grid = {'n_comp': (0.1, 0.95, 'uniform'),
                          'decoder__dimredreshape__n_components': (0.1, 0.95, 'uniform')}
        if not pool_train or joint_dim_red:
            raise NotImplementedError('-cv True is batched for the pooled CCA / MCCA / unaligned '
                                      'decoders (the search of aligned_decode_svm_ncv.py:388-394)')
    elif mcca_align:
        param_grid = {'n_comp': 30, 'regs': 0.5, 'pca_var': 0.8,
                      'decoder__dimredreshape__n_components': 0.8}
    else:
        param_grid = {'n_comp': 0.9, 'decoder__dimredreshape__n_components': 0.8}
    algn_type, algn_grouping, lab_type, red_method = 'phon_seq', 'class', 'phon', 'PCA'
    if inputs['filename']:
        filename = inputs['filename']
    else:
        prefix = out_path + ('outputs/alignment_accs/%s/' % pt if cluster else 'ncv_accs/%s/' % pt)
        filename = prefix + '%s_%s%s_%s.pkl' % (pt, 'p' if lab_type == 'phon' else 'a',
                                               'All' if p_ind == -1 else p_ind, inputs['suffix'])
    data_file = inputs['data_file'] or data_path + (
        'pt_decoding_data_S62_TME.pkl' if str2bool(inputs['surrogate']) else 'pt_decoding_data_S62.pkl')
    if inputs.get('seed') is not None:
        np.random.seed(inputs['seed'])
    pt_data = utils.load_pkl(data_file)
    (D_tar, lab_tar, lab_tar_full), pre_data = utils.decoding_data_from_dict(
        pt_data, pt, p_ind, lab_type=lab_type, algn_type=algn_type)
    if str2bool(inputs['random_data']):
        cross = [(np.random.rand(*d[0].shape), d[1], d[2]) for d in pre_data]
    elif inputs['pooled_patients'] != 'all':
        pre_pts = pt_data[pt]['pre_pts']
        cross = [pre_data[pre_pts.index(p)] for p in inputs['pooled_patients'].split(',')]
    else:
        cross = pre_data
    out = {'params': {'pt': pt, 'p_ind': p_ind, 'pool_train': pool_train,
                      'tar_in_train': tar_in_train, 'cca_align': cca_align,
                      'joint_dim_red': joint_dim_red, 'n_iter': n_iter, 'n_folds': n_folds,
                      'hyperparams': param_grid, 'algn_type': algn_type,
                      'algn_grouping': algn_grouping, 'lab_type': lab_type,
                      'red_method': red_method}}
    lab_tar = np.asarray(lab_tar)
    units = make_units(lab_tar, n_iter, n_folds, inputs['trial_subsample'])
    mine = sharding.shard_units(len(units), n_folds, rank, world)    # whole iterations per rank
    my_units = [units[u] for u in mine]
    dec_kw = dict(decoder=inputs['decoder'],
                  class_weight='balanced' if inputs['decoder'] == 'svc_rbf' else None)
    if not do_cv:
        dec_kw['decoder_var'] = param_grid['decoder__dimredreshape__n_components']
    # one search seed per unit, drawn up front: a sharded run gives every rank the same searches
    search_seeds = [int(np.random.randint(2 ** 31 - 1)) for _ in units] if do_cv else None
    def make_clf():
        from sklearn.pipeline import make_pipeline
        from ..decomposition.DimRedReshape import DimRedReshape
        from ..decomposition.PCA import PCA
        from ..svm import SVC, LinearSVC
        dec = LinearSVC() if inputs['decoder'] == 'linear' else SVC(
            kernel=inputs['decoder'][4:], class_weight=dec_kw['class_weight'])
        return make_pipeline(DimRedReshape(PCA, n_components=dec_kw['decoder_var']), dec)

    if do_cv:
        # nested search per outer fold (:388-405): 25 candidates in rounds of 5, each scored on
        # n_folds inner folds of the outer-train trials, then one fit / predict with the winner
        from .. import cv_align_decode_stream
        from ..search import bayes_search_align_decode, engine_keywords
        method = 'cca' if cca_align else ('mcca' if mcca_align else 'none')
        base_kw = dict(dec_kw, tar_in_train=tar_in_train, use_tensor_cores=True)
        if mcca_align:
            base_kw['regs'] = 0.5
        best = []
        for u in mine:
            tr, _ = units[u]
            bs = bayes_search_align_decode((D_tar[tr], lab_tar[tr], lab_tar_full[tr]), cross, param_grid,
                                           n_folds=n_folds, n_iter=inputs.get('search_iter', 25),
                                           n_points=inputs.get('search_points', 5), method=method,
                                           random_state=search_seeds[u], shard=False, max_batch=148,
                                           **base_kw)
            best.append(bs['best_params_'])
        jobs = (((D_tar, lab_tar, lab_tar_full), cross, [units[u]], engine_keywords(b))
                for u, b in zip(mine, best))
        y_pred_mine = [r['y_pred'][0] for r in cv_align_decode_stream(jobs, depth=8, method=method,
                                                                     max_batch=8, **base_kw)]
        out['params']['best_params'] = sharding.gather_objects(mine, best)
    elif pool_train and not joint_dim_red:
        if cca_align:
            kw = dict(method='cca', n_comp=param_grid['n_comp'])
        elif mcca_align:
            kw = dict(method='mcca', n_comp=30, regs=0.5, pca_var=0.8)
        else:
            kw = dict(method='none', n_comp=param_grid['n_comp'])
        res = run_units((D_tar, lab_tar, lab_tar_full), cross, my_units,
                        tar_in_train=tar_in_train, use_tensor_cores=True, max_batch=148,
                        **kw, **dec_kw) if my_units else {'y_pred': []}
        y_pred_mine = res['y_pred']
    elif pool_train:
        # joint PCA: the script's set_params hands n_comp = 0.9 (a variance fraction) to
        # JointPCA(n_components=...) (:186-190, :372-375, :416); the batched engine sizes the batch
        # from a fit on all trials and cuts every fold at its own component count
        res = run_units((D_tar, lab_tar, lab_tar_full), cross, my_units, method='jointpca',
                        n_comp=param_grid['n_comp'], tar_in_train=tar_in_train, use_tensor_cores=True,
                        max_batch=148, **dec_kw) if my_units else {'y_pred': []}
        y_pred_mine = res['y_pred']
    else:
        # single-patient branch (:406-428): DimRedReshape(PCA(0.8)) -> decoder on the raw trials
        y_pred_mine = []
        for tr, te in my_units:
            clf = make_clf()
            clf.fit(D_tar[tr], lab_tar[tr])
            y_pred_mine.append(clf.predict(D_tar[te]))
    # the one collective: every rank receives the labels of all units
    allp = sharding.gather_predictions(mine, y_pred_mine)
    y_pred_units = [allp[u] for u in range(len(units))]
    y_true_iter, y_pred_iter, wrong_iter, accs = [], [], [], []
    for j in range(n_iter):
        yt, yp, wrong = [], [], []
        for u in range(j * n_folds, (j + 1) * n_folds):
            te = units[u][1]
            y_test, y_hat = lab_tar[te], np.asarray(y_pred_units[u])
            yt.extend(y_test)
            yp.extend(y_hat)
            wrong.extend(te[np.where(y_test != y_hat)[0]])
        y_true_iter.append(yt)
        y_pred_iter.append(yp)
        wrong_iter.append(wrong)
        accs.append(balanced_accuracy_score(yt, yp))
    out.update(y_true=y_true_iter, y_pred=y_pred_iter, wrong_trs=wrong_iter, accs=accs)
    if rank == 0:
        d = os.path.dirname(filename)
        if d:
            os.makedirs(d, exist_ok=True)
        utils.save_pkl(out, filename)
    return out


def aligned_decoding(argv=None):
    args = init_parser().parse_args(argv)
    return run(dict(vars(args)))


if __name__ == '__main__':
    aligned_decoding()
    print('########## Done ###########')
