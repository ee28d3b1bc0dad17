"""Golden output of the reference's own command-line script, UNMODIFIED:
/root/reference/aligned_decoding/scripts/aligned_decode_svm_ncv.py run through runpy on a
small synthetic data dictionary (the script's constants stay: 50 iterations x 20 folds,
SVC(kernel='rbf', class_weight='balanced'), PCA 0.9 / 0.8), pairwise-CCA pooling
(``-po True -a True``).  Environment shims only: a scratch working directory holding
``../data/pt_decoding_data_S62.pkl`` (the script's hard-coded relative path), the reference
package on sys.path, and stub modules for the two absent dependencies the script imports at
module level but does not use on this path (``skopt``; ``mvlearn`` via oracle/reference_path).

Writes tests/golden/script_ncv_cca.npz (+ the input dictionary, script_ncv_data.npz).
Run in the build container: python tests/golden/make_golden_script.py
"""
import os
import pickle
import runpy
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from cross_patient_speech_decoding_b200 import synthetic  # noqa: E402
from oracle import reference_path  # noqa: E402

SEED = 31
PATIENTS = {'S1': dict(p=0, n_trials=72, n_time=24, n_chan=20, noise=2.0),
            'S2': dict(p=1, n_trials=90, n_time=24, n_chan=28, noise=2.0),
            'S3': dict(p=2, n_trials=60, n_time=24, n_chan=16, noise=2.0)}


def data_dict():
    d = {}
    for name, kw in PATIENTS.items():
        X, y, ya = synthetic.make_patient(**kw)
        d[name] = {'X1': X, 'y1': y, 'y_full_phon': ya,
                   'pre_pts': [p for p in PATIENTS if p != name]}
    return d


def main():
    ref_root = '/root/reference/aligned_decoding'
    scratch = os.path.join(ROOT, 'gpurun_out', 'tmp_ref')
    os.makedirs(os.path.join(scratch, 'scripts'), exist_ok=True)
    os.makedirs(os.path.join(scratch, 'data'), exist_ok=True)
    d = data_dict()
    with open(os.path.join(scratch, 'data', 'pt_decoding_data_S62.pkl'), 'wb') as f:
        pickle.dump(d, f, protocol=-1)
    reference_path.load()                                   # mvlearn stand-in (MCCA unused here)
    if 'skopt' not in sys.modules:
        sk = types.ModuleType('skopt')
        sk.BayesSearchCV = None                             # only used with -cv True
        sys.modules['skopt'] = sk
    sys.path.insert(0, ref_root)
    out_file = os.path.join(scratch, 'ref_out.pkl')
    argv = ['aligned_decode_svm_ncv.py', '-pt', 'S1', '-pi', '1', '-po', 'True', '-a', 'True',
            '-c', 'False', '-f', out_file]
    cwd = os.getcwd()
    os.chdir(os.path.join(scratch, 'scripts'))
    old_argv = sys.argv
    sys.argv = argv
    try:
        np.random.seed(SEED)
        runpy.run_path(os.path.join(ref_root, 'scripts', 'aligned_decode_svm_ncv.py'),
                       run_name='__main__')
    finally:
        sys.argv = old_argv
        os.chdir(cwd)
    with open(out_file, 'rb') as f:
        res = pickle.load(f)
    np.savez_compressed(os.path.join(HERE, 'script_ncv_cca.npz'), seed=SEED,
                        accs=np.array(res['accs']), y_true=np.array(res['y_true']),
                        y_pred=np.array(res['y_pred']),
                        n_wrong=np.array([len(w) for w in res['wrong_trs']]),
                        wrong0=np.array(res['wrong_trs'][0]),
                        param_keys=np.array(sorted(res['params'].keys())))
    print('accs mean %.4f (%d iterations), first %s' % (np.mean(res['accs']), len(res['accs']),
                                                        np.round(res['accs'][:4], 4)))


if __name__ == '__main__':
    main()
