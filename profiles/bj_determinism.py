import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cross_patient_speech_decoding_b200 import ops
rng = np.random.default_rng(0)
for n in (259, 384, 640):
    X = rng.standard_normal((n, 3 * n)) * (1.0 / (1.0 + np.arange(3 * n) / 40.0))
    A = (X @ X.T)[None]
    for tc in (False, True):
        r = [ops.eig_sym(A, tensor_cores=tc, return_sweeps=True) for _ in range(3)]
        d_ev = max(np.abs(r[0][0] - x[0]).max() / np.abs(r[0][0]).max() for x in r[1:])
        d_v = max(np.abs(np.abs(r[0][1]) - np.abs(x[1])).max() for x in r[1:])
        print('n=%d tc=%s: evals rel diff between runs %.2e, |V| diff %.2e, sweeps %s' % (
            n, tc, d_ev, d_v, [int(np.ravel(x[2])[0]) for x in r]))
