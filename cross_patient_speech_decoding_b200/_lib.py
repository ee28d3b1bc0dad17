"""ctypes binding of the C-ABI shared library ``libcpsd_b200.so`` (include/cpsd_b200.h).

There is no CPU fallback: if the library is missing or cannot be loaded the product
path raises.  Build it with ``python -m cross_patient_speech_decoding_b200.build`` (nvcc,
sm_100a) or ``__graft_entry__.build()``.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('CPSD_LIB') or os.path.join(_HERE, 'libcpsd_b200.so')

c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_float = ctypes.c_float
c_double = ctypes.c_double
c_void_p = ctypes.c_void_p

# numpy mirrors of csrc/descs.h (all fields naturally aligned, no implicit padding)
GRAM_TN_DESC = np.dtype([
    ('A', 'u8'), ('B', 'u8'), ('segA', 'u8'), ('segB', 'u8'), ('muA', 'u8'), ('muB', 'u8'),
    ('out', 'u8'), ('nseg', 'i4'), ('seg_len', 'i4'), ('p', 'i4'), ('q', 'i4'), ('lda', 'i4'),
    ('ldb', 'i4'), ('ldo', 'i4'), ('sym', 'i4'), ('alpha', 'f4'), ('pad_', 'i4')])
COLSUM_DESC = np.dtype([
    ('A', 'u8'), ('segA', 'u8'), ('out', 'u8'), ('nseg', 'i4'), ('seg_len', 'i4'), ('p', 'i4'),
    ('lda', 'i4'), ('alpha', 'f4'), ('pad_', 'i4')])
PROJ_DESC = np.dtype([
    ('X', 'u8'), ('seg_src', 'u8'), ('seg_dst', 'u8'), ('mu', 'u8'), ('W', 'u8'), ('Y', 'u8'),
    ('nseg', 'i4'), ('seg_len', 'i4'), ('C', 'i4'), ('q', 'i4'), ('ldx', 'i4'), ('ldw', 'i4'),
    ('ldy', 'i4'), ('pad_', 'i4')])
GRAM_NT_DESC = np.dtype([
    ('A', 'u8'), ('B', 'u8'), ('out', 'u8'), ('m', 'i4'), ('n', 'i4'), ('k', 'i4'), ('lda', 'i4'),
    ('ldb', 'i4'), ('ldo', 'i4'), ('sym', 'i4'), ('alpha', 'f4')])
CLASS_MEAN_DESC = np.dtype([
    ('X', 'u8'), ('member_ptr', 'u8'), ('members', 'u8'), ('out', 'u8'), ('nslot', 'i4'),
    ('TC', 'i4'), ('pad0_', 'i4'), ('pad1_', 'i4')])
SVM_DESC = np.dtype([
    ('St', 'u8'), ('y', 'u8'), ('k_dev', 'u8'), ('w', 'u8'), ('info', 'u8'), ('n', 'i4'),
    ('k', 'i4'), ('lds', 'i4'), ('cls', 'i4'), ('C', 'f8'), ('tol_dcd', 'f8'),
    ('tol_newton', 'f8'), ('max_newton', 'i4'), ('dcd_epochs', 'i4')])
CCA_DESC = np.dtype([
    ('Saa', 'u8'), ('Sbb', 'u8'), ('Sab', 'u8'), ('da_dev', 'u8'), ('db_dev', 'u8'),
    ('Ma', 'u8'), ('Mb', 'u8'), ('G', 'u8'), ('rho', 'u8'), ('info', 'u8'),
    ('da', 'i4'), ('db', 'i4'), ('lds', 'i4'), ('ldm', 'i4'), ('ldg', 'i4'), ('mode', 'i4'),
    ('rank_tol', 'f4'), ('pad_', 'i4')])

_EXPECTED_SIZES = {'gram_tn': 96, 'colsum': 48, 'proj': 80, 'gram_nt': 56, 'class_mean': 48,
                   'svm': 88, 'cca': 112}
assert GRAM_TN_DESC.itemsize == 96 and COLSUM_DESC.itemsize == 48
assert PROJ_DESC.itemsize == 80 and GRAM_NT_DESC.itemsize == 56
assert CLASS_MEAN_DESC.itemsize == 48 and SVM_DESC.itemsize == 88
assert CCA_DESC.itemsize == 112

# name -> argtypes (restype is int status unless listed in _RESTYPES)
_P = c_void_p
_SIGS = {
    'cpsd_version': [],
    'cpsd_device_arch': [],
    'cpsd_launch_count': [],
    'cpsd_reset_launch_count': [],
    'cpsd_last_error': [],
    'cpsd_desc_sizes': [_P],
    'cpsd_eig_sym_small': [_P, c_int, c_ll, _P, c_int, c_int, _P, c_int, _P, c_int, c_ll, c_int,
                           c_float, _P, _P],
    'cpsd_eig_sym_small_f64': [_P, c_int, c_ll, _P, c_int, c_int, _P, c_int, _P, c_int, c_ll, c_int,
                               c_float, _P, _P],
    'cpsd_eig_sym_small_f64_warm': [_P, c_int, c_ll, _P, c_int, _P, c_int, _P, _P, c_int, _P, c_int, c_ll,
                                    _P, c_int, c_ll, _P, c_int, c_float, _P, _P],
    'cpsd_dgemm_batched': [c_int, c_int, c_int, c_int, c_double, _P, c_int, c_ll, _P, _P, c_int, c_ll, _P,
                           c_double, _P, c_int, c_ll, _P, _P, c_int, c_ll, _P, c_int, _P],
    'cpsd_cast_f32_f64_idx': [_P, c_ll, _P, _P, c_ll, c_ll, c_int, _P],
    'cpsd_gram_nt_tc_probe': [_P],
    'cpsd_bj_schedule': [c_int, _P],
    'cpsd_eig_sym_block': [_P, c_int, c_ll, c_int, _P, c_int, c_int, _P, _P, _P, _P, _P, _P,
                           c_int, c_int, c_float, _P, _P],
    'cpsd_bj_rlog_elems': [c_int, c_int, c_int],
    'cpsd_bj_eigvecs': [_P, c_int, c_int, _P, _P, _P, c_int, _P, c_int, c_int, _P, c_int, c_ll,
                        c_int, _P],
    'cpsd_select_k': [_P, c_int, _P, c_int, c_float, c_int, c_int, c_int, _P, c_int, c_int, _P],
    'cpsd_select_k_total': [_P, c_int, _P, c_int, _P, c_float, c_int, c_int, c_int, _P, c_int,
                            c_int, _P],
    'cpsd_eig_topk_ws_elems': [c_int, c_int, c_int],
    'cpsd_eig_topk_voff': [c_int, c_int, c_int],
    'cpsd_eig_sym_topk': [_P, c_int, c_ll, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, _P,
                          c_int, _P, _P, _P, c_int, c_float, c_int, _P],
    'cpsd_topk_tc_ws_elems': [c_int, c_int],
    'cpsd_topk_tc_map_bytes': [c_int],
    'cpsd_topk_tc_encode': [_P, c_int, c_ll, c_int, c_int, _P, _P, _P, _P],
    'cpsd_topk_tc_split_k': [_P, c_int, c_ll, c_int, c_int, _P, _P],
    'cpsd_topk_tc_kq': [_P, c_ll, _P, c_ll, c_int, c_int, c_int, _P, _P, _P],
    'cpsd_eig_sym_topk_tc': [_P, c_int, c_ll, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, _P,
                             c_int, _P, _P, _P, c_int, c_float, _P, _P, c_int, c_int, _P],
    'cpsd_sgemm_batched': [c_int, c_int, c_int, c_int, c_float, _P, c_int, c_ll, _P, c_int, c_ll,
                           _P, c_int, c_ll, c_int, _P],
    'cpsd_chol_solve_f64': [_P, c_int, c_ll, c_int, _P, c_int, c_ll, c_int, _P, c_int, c_ll, _P, c_int,
                            _P],
    'cpsd_chol_solve_f64_ws': [_P, c_int, c_ll, c_int, _P, c_int, c_ll, c_int, _P, c_int, c_ll, _P, _P,
                               c_int, _P],
    'cpsd_chol_inv': [_P, c_int, c_ll, c_int, _P, c_int, c_ll, _P, c_int, _P],
    'cpsd_gram_tn': [_P, c_int, c_int, c_int, _P],
    'cpsd_gram_tn_f64': [_P, c_int, c_int, c_int, _P],
    'cpsd_gram_tn_f64_split': [_P, c_int, c_int, c_int, c_int, c_int, _P, _P],
    'cpsd_gram_tn_split_ws_elems': [c_int, c_int, c_int, c_int],
    'cpsd_colsum': [_P, c_int, c_int, _P],
    'cpsd_proj_nn': [_P, c_int, c_int, c_int, c_int, _P],
    'cpsd_gram_nt': [_P, c_int, c_int, c_int, _P],
    'cpsd_class_mean': [_P, c_int, c_int, c_int, _P],
    'cpsd_center_rows': [_P, c_int, c_ll, _P, c_int, _P, c_int, c_int, c_int, c_int, _P],
    'cpsd_copy_rows': [_P, c_int, c_ll, _P, c_int, c_ll, _P, c_int, c_int, c_int, c_int, c_int, _P],
    'cpsd_sum_mats_f64': [_P, _P, c_ll, _P, _P, ctypes.c_double, _P, c_ll, c_int, c_int, _P],
    'cpsd_trial_colsum_f64': [_P, c_int, c_int, c_int, c_int, _P, c_int, _P],
    'cpsd_cov_from_sums': [_P, c_int, c_ll, _P, c_int, _P, c_int, _P, c_int, _P, c_int, c_int, _P],
    'cpsd_gather_channels': [_P, c_int, _P, c_int, _P, c_int, c_ll, _P],
    'cpsd_gather_trials': [_P, c_ll, _P, c_int, _P, _P],
    'cpsd_mask_cols': [_P, c_int, c_ll, c_int, c_int, _P, c_int, c_int, _P],
    'cpsd_predict_fused': [_P, c_int, c_int, c_int, _P, _P, c_int, _P, _P, c_int, _P, _P, c_int, _P, _P,
                           c_int, _P, _P, _P, c_int, _P],
    'cpsd_pearson_rows': [_P, _P, c_int, c_ll, _P, _P],
    'cpsd_region_mean_f64': [_P, c_int, c_int, c_int, _P, _P, c_int, _P, _P],
    'cpsd_cast_f64_f32': [_P, _P, c_ll, _P],
    'cpsd_permute_cols': [_P, c_int, c_ll, _P, c_int, _P, c_int, c_ll, c_int, c_int, c_int, _P],
    'cpsd_mcca_mask': [_P, c_int, c_ll, _P, c_int, _P, _P, c_int, c_int, _P, _P, _P, c_int, _P],
    'cpsd_mcca_mask_idx': [_P, c_int, c_ll, _P, c_int, _P, _P, _P, c_int, c_int, _P, _P, _P, c_int,
                           _P],
    'cpsd_joint_cov': [_P, _P, _P, c_int, _P, c_int, c_int, _P],
    'cpsd_joint_rhs': [_P, _P, _P, c_int, _P, c_int, c_ll, _P, c_int, c_int, c_int, _P, c_int, c_ll,
                       c_int, _P],
    'cpsd_mcca_build': [_P, c_int, c_ll, _P, c_int, c_int, c_float, _P, c_int, c_ll, _P, _P, _P,
                        c_int, _P, c_int, _P],
    'cpsd_mcca_build_split': [_P, c_int, c_ll, _P, c_int, c_ll, _P, _P, c_int, c_int, c_float, _P,
                              c_int, c_ll, _P, _P, _P, c_int, _P, c_int, _P],
    'cpsd_mcca_loadings': [_P, _P, c_int, c_ll, _P, c_int, _P, _P, c_int, c_int, c_int, c_int, _P,
                           c_int, c_int, _P],
    'cpsd_scores_train': [_P, c_int, c_ll, _P, _P, c_int, _P, _P, c_int, c_int, _P, c_int, c_ll,
                          c_int, c_int, _P],
    'cpsd_scores_test': [_P, c_int, c_ll, _P, c_int, c_ll, _P, _P, c_int, _P, _P, c_int, c_int,
                         _P, c_int, c_ll, c_int, c_int, _P],
    'cpsd_svm_fit_ovr': [_P, c_int, c_int, c_int, _P],
    'cpsd_svm_fit_ovr_ex': [_P, c_int, c_int, c_int, c_int, _P],
    'cpsd_svm_predict_ovr': [_P, c_int, c_ll, _P, c_int, c_ll, _P, c_int, _P, c_int, _P, c_int, _P,
                             _P, c_int, _P],
    'cpsd_svc_kernel_matrix': [_P, c_int, c_ll, _P, c_int, _P, c_int, c_int, _P, c_int, _P, c_int, c_int,
                               ctypes.c_double, _P, _P, _P, _P, _P, c_int, c_ll, c_int, _P],
    'cpsd_svc_fit_ovo': [_P, c_int, c_ll, _P, _P, _P, c_int, c_int, ctypes.c_double, c_int,
                         ctypes.c_double, c_int, _P, c_int, _P, _P, c_int, c_int, _P],
    'cpsd_svc_predict_ovo': [_P, c_int, c_ll, _P, c_int, c_ll, _P, c_int, _P, c_int, c_int, _P, c_int,
                             _P, c_int, _P, c_int, c_int, _P, _P, c_int, _P, _P, _P, c_int, c_int, _P],
    'cpsd_bag_gather': [_P, c_int, c_ll, _P, c_int, c_ll, _P, c_int, _P, c_int, _P, _P, _P, c_int, c_int,
                        _P, c_int, _P, c_int, _P, _P, _P, _P, c_int, _P],
    'cpsd_bag_vote': [_P, c_int, _P, c_int, _P, c_int, _P, c_int, _P],
    'cpsd_cca_solve': [_P, c_int, c_int, _P],
    'cpsd_cca_solve_f64_ws_elems': [c_int, c_int],
    'cpsd_cca_solve_f64': [_P, c_int, c_int, _P, _P],
    'cpsd_eig_sym_f64_ws_elems': [c_int, c_int],
    'cpsd_eig_sym_f64': [_P, c_int, c_ll, _P, c_int, c_int, _P, c_int, _P, c_int, c_ll, c_int, _P,
                         c_int, _P, _P],
    'cpsd_pca_basis': [_P, c_int, c_ll, _P, _P, c_int, c_int, _P, c_int, c_int, c_int, _P],
    'cpsd_split_tf32': [_P, _P, _P, c_ll, _P],
    'cpsd_tmap_encode_f32': [_P, _P, c_ll, c_int, c_ll, c_int],
    'cpsd_proj_tc_rep': [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, c_int, _P, _P, _P,
                         c_ll, c_int, _P],
    'cpsd_split_tf32_2d': [_P, c_ll, c_ll, c_int, _P, _P, c_int, _P],
    'cpsd_proj_tc_prep': [_P, c_int, c_ll, _P, _P, c_int, _P, c_int, c_int, _P, _P, _P, c_int, _P],
    'cpsd_proj_tc': [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, c_int, _P, _P, _P, c_ll, c_int,
                     _P],
    'cpsd_gram_nt_tc': [_P, c_int, c_int, c_int, _P, c_ll, _P, _P, _P],
    'cpsd_gram_nt_tc_centered': [_P, c_int, c_int, c_int, _P, c_ll, _P, _P, _P, c_int, _P],
    'cpsd_gram_nt_tc_ws_bytes': [c_int],
}
_RESTYPES = {'cpsd_last_error': ctypes.c_char_p, 'cpsd_launch_count': c_ll,
             'cpsd_bj_rlog_elems': c_ll, 'cpsd_eig_topk_ws_elems': c_ll,
             'cpsd_eig_topk_voff': c_ll, 'cpsd_topk_tc_ws_elems': c_ll,
             'cpsd_gram_tn_split_ws_elems': c_ll, 'cpsd_cca_solve_f64_ws_elems': c_ll,
             'cpsd_eig_sym_f64_ws_elems': c_ll,
             'cpsd_reset_launch_count': None}

EXPORTED_SYMBOLS = sorted(_SIGS)

_lib = None


class CpsdError(RuntimeError):
    pass


def load():
    """Loads the shared library (once).  Raises if it is missing -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CpsdError(
            'libcpsd_b200.so not found at %s; build it with '
            '`python -m cross_patient_speech_decoding_b200.build`' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name, c_int)
    sizes = (c_int * 8)()
    lib.cpsd_desc_sizes(ctypes.cast(sizes, c_void_p))
    got = dict(zip(['gram_tn', 'colsum', 'proj', 'gram_nt', 'class_mean', 'svm', 'cca'], sizes))
    if got != _EXPECTED_SIZES:
        raise CpsdError('descriptor layout mismatch between _lib.py and descs.h: %r' % (got,))
    _lib = lib
    return lib


def check(status, what=''):
    if status != 0:
        msg = load().cpsd_last_error()
        raise CpsdError('%s failed (status %d): %s'
                        % (what or 'cpsd call', status, msg.decode() if msg else ''))
