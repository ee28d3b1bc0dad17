import sys, os, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests/golden')
import numpy as np, torch
import make_golden
from cross_patient_speech_decoding_b200.engine import CVEngine
for name in ('cca_p2_5fold', 'cca_p2_fixed30', 'cca_p2_noisy', 'cca_p3_ragged'):
    cfg = make_golden.CONFIGS[name]
    pts, folds = make_golden.build_inputs(cfg)
    g = np.load('/root/repo/tests/golden/%s.npz' % name)
    nf = int(g['n_folds'])
    yr = np.concatenate([g['y_pred_%d' % f] for f in range(nf)])
    many = folds * 16
    for solver in ('auto', 'topk'):
        eng = CVEngine(pts[0], pts[1:], method='cca', n_comp=cfg['n_comp'], use_tensor_cores=True, pool_solver=solver, max_batch=148)
        res = eng.run(folds)
        yp = np.concatenate(res['y_pred'])
        eng.run(many); torch.cuda.synchronize()
        t0 = time.perf_counter(); eng.run(many); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(name, solver, 'agree %d/%d' % ((yp == yr).sum(), len(yr)), 'k2', res['k2'], list(g['k2']), '%.0f folds/s' % (len(many) / dt),
              {k: eng.stats['topk'].get(k) for k in ('rounds', 'ok', 'why', 'resid')} if 'topk' in eng.stats else None, flush=True)
