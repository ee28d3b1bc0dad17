"""Batched counterpart of the reference's ``scripts/aligned_decode_cross_patient_subsample.py``:
same command line (``-pt -pi -t -a -m -j -r -c -cv -f -s``), same data-dictionary pickle in, same
result pickle out (``{'params', 'acc_mat', 'trial_vec'}``, reference lines 236-245, 383-386); the
``len(k) x n_iter x n_folds`` fit / predict units of its three nested loops (lines 290-381) run as
jobs of ``processing_utils.trial_subsample_decode`` (numpy global RNG consumed in the script's
order).  ``-cv True`` (nested search) and ``-j True`` are not batched here.
Extra options: ``--data_file``, ``--n_iter``, ``--n_folds``, ``--trial_step``, ``--decoder``,
``--seed``.  Runs sharded under torchrun (jobs dealt to the ranks, rank 0 writes the pickle)."""
import argparse
import os

import numpy as np

from .. import sharding
from ..alignment import alignment_utils as utils
from ..processing_utils.trial_subsample_decode import trial_counts, trial_subsample_decode


def init_parser():
    ap = argparse.ArgumentParser(description='Cross-patient decoding with subsampling of pooled '
                                             'data across patients (batched, B200)')
    ap.add_argument('-pt', '--patient', type=str, required=True, help='Patient ID')
    ap.add_argument('-pi', '--p_ind', type=int, default=-1, help='Sequence position index')
    ap.add_argument('-t', '--tar_in_train', type=str, default='True')
    ap.add_argument('-a', '--cca_align', type=str, default='False')
    ap.add_argument('-m', '--MCCA_align', type=str, default='False')
    ap.add_argument('-j', '--joint_dim_red', type=str, default='False')
    ap.add_argument('-r', '--random_data', type=str, default='False')
    ap.add_argument('-c', '--cluster', type=str, default='True')
    ap.add_argument('-cv', '--cross_validate', type=str, default='False')
    ap.add_argument('-f', '--filename', type=str, default='')
    ap.add_argument('-s', '--suffix', type=str, default='')
    ap.add_argument('--data_file', type=str, default='')
    ap.add_argument('--n_iter', type=int, default=50)
    ap.add_argument('--n_folds', type=int, default=20)
    ap.add_argument('--trial_step', type=int, default=25)
    ap.add_argument('--decoder', type=str, default='svc_rbf', choices=['svc_rbf', 'svc_linear', 'linear'])
    ap.add_argument('--seed', type=int, default=None)
    return ap


def str2bool(s):
    return s.lower() == 'true'


def run(inputs):
    rank, world, _ = sharding.init_from_env()
    cluster = str2bool(inputs['cluster'])
    home = os.path.expanduser('~')
    data_path, out_path = (home + '/data/', home + '/workspace/') if cluster else ('../data/', '../acc_data/')
    pt, p_ind = inputs['patient'], inputs['p_ind']
    tar_in_train = str2bool(inputs['tar_in_train'])
    cca_align, mcca_align = str2bool(inputs['cca_align']), str2bool(inputs['MCCA_align'])
    joint_dim_red = str2bool(inputs['joint_dim_red'])
    if str2bool(inputs['cross_validate']):
        raise NotImplementedError('-cv True: use scripts.aligned_decode_svm_ncv for the nested search')
    if sum([cca_align, mcca_align, joint_dim_red]) > 1:
        cca_align = mcca_align = False
    n_iter, n_folds = inputs['n_iter'], inputs['n_folds']
    if mcca_align:
        param_grid = {'n_comp': 30, 'regs': 0.5, 'pca_var': 0.8, 'decoder__dimredreshape__n_components': 0.8}
        kw = dict(method='mcca', n_comp=30, regs=0.5, pca_var=0.8)
    elif joint_dim_red:
        raise NotImplementedError('-j True: JointPCA takes n_comp = 0.9 here; use the drop-in classes')
    else:
        param_grid = {'n_comp': 0.9, 'decoder__dimredreshape__n_components': 0.8}
        kw = dict(method='cca' if cca_align else 'none', n_comp=0.9)
    algn_type, algn_grouping, lab_type, red_method = 'phon_seq', 'class', 'phon', 'PCA'
    if inputs['filename']:
        filename = inputs['filename']
    else:
        prefix = out_path + ('outputs/alignment_accs/%s/' % pt if cluster else 'ncv_accs/%s/' % pt)
        filename = prefix + '%s_%s%s_%s.pkl' % (pt, 'p' if lab_type == 'phon' else 'a',
                                               'All' if p_ind == -1 else p_ind, inputs['suffix'])
    if inputs.get('seed') is not None:
        np.random.seed(inputs['seed'])
    pt_data = utils.load_pkl(inputs['data_file'] or data_path + 'pt_decoding_data_S62.pkl')
    (D_tar, lab_tar, lab_tar_full), pre_data = utils.decoding_data_from_dict(
        pt_data, pt, p_ind, lab_type=lab_type, algn_type=algn_type)
    if str2bool(inputs['random_data']):
        pre_data = [(np.random.rand(*d[0].shape), d[1], d[2]) for d in pre_data]
    out = {'params': {'pt': pt, 'p_ind': p_ind, 'tar_in_train': tar_in_train, 'cca_align': cca_align,
                      'joint_dim_red': joint_dim_red, 'n_iter': n_iter, 'n_folds': n_folds,
                      'hyperparams': param_grid, 'algn_type': algn_type,
                      'algn_grouping': algn_grouping, 'lab_type': lab_type, 'red_method': red_method}}
    res = trial_subsample_decode((D_tar, np.asarray(lab_tar), lab_tar_full), pre_data,
                                 k_list=trial_counts(pre_data, inputs['trial_step']), n_iter=n_iter,
                                 n_folds=n_folds, decoder=inputs['decoder'], decoder_var=0.8,
                                 class_weight='balanced' if inputs['decoder'] == 'svc_rbf' else None,
                                 tar_in_train=tar_in_train, use_tensor_cores=True, **kw)
    out['acc_mat'], out['trial_vec'] = res['acc_mat'], res['trial_vec']
    if rank == 0:
        d = os.path.dirname(filename)
        if d:
            os.makedirs(d, exist_ok=True)
        utils.save_pkl(out, filename)
    return out


def pooled_sampled_decoding(argv=None):
    return run(dict(vars(init_parser().parse_args(argv))))


if __name__ == '__main__':
    pooled_sampled_decoding()
    print('########## Done ###########')
