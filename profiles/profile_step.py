"""One warm-up step + N measured steps of the headline workload (bench.py's step), for ncu.
Usage: python profiles/profile_step.py [--steps N] [--no-tc] [--folds F]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=1)
ap.add_argument('--same', action='store_true', help='repeat the same folds (warm caches)')
ap.add_argument('--folds', type=int, default=107)
ap.add_argument('--no-tc', action='store_true')
ap.add_argument('--stages', action='store_true')
ap.add_argument('--dcd', type=int, default=0)
ap.add_argument('--decoder', default='linear')
ap.add_argument('--topk-iters', type=int, default=8)
ap.add_argument('--tf32-iters', type=int, default=5)
ap.add_argument('--class-weight', default=None)
a = ap.parse_args()
from cross_patient_speech_decoding_b200.engine import CVEngine  # noqa: E402
pts = bench.make_data()
eng = CVEngine(pts[0], pts[1:], method='mcca', n_comp=30, regs=0.5, pca_var=0.8,
               use_tensor_cores=not a.no_tc, max_batch=a.folds, dcd_epochs=a.dcd,
               decoder=a.decoder, class_weight=a.class_weight, topk_iters=a.topk_iters,
               topk_tf32_iters=a.tf32_iters)
def mk(seed):
    out = []
    while len(out) < a.folds:
        out += bench.step_folds(pts[0][1], seed + len(out))
    return out[:a.folds]


eng.run(mk(1000))
eng.run(mk(1500))        # second use: trial statistics move to the all-trials eigenbasis (once)
for s in range(a.steps):
    eng.profile = a.stages
    fl = mk(2000 if a.same else 2000 + 100 * s)
    res = eng.run(fl, return_details=True)
    if a.stages:
        print('stages_ms', {k: round(v, 3) for k, v in eng.collect_marks().items()})
    if a.decoder != 'linear':
        si = res['details'][0]['svm_info']
        print('svc: SMO iterations per pair mean %.0f max %d; not converged %d' % (
            si[..., 0].mean(), si[..., 0].max(), int((si[..., 1] == 1).sum())))
        continue
    sw = res['details'][0]['bj_sweeps']
    print('acc', sum(int((p == pts[0][1][te]).sum()) for p, (_, te) in zip(res['y_pred'], fl)) / sum(len(te) for _, te in fl))
    print('k2', res['k2'][:4], 'topk', eng.stats.get('topk'), 'bj_sweeps', None if sw is None else sw[:4],
          'svm_newton_max', int(res['details'][0]['svm_info'][..., 0].max()),
          'launches', eng.stats['launches_last_batch'])
    si = res['details'][0]['svm_info']
    print('svm newton its mean %.2f max %d; cg iterations per task mean %.1f max %d; dcd epochs %d' % (
        si[..., 0].mean(), si[..., 0].max(), si[..., 1].mean(), si[..., 1].max(), si[..., 2].max()))
