#!/usr/bin/env python
"""Headline benchmark: CV align+decode folds/sec (MCCA -> PCA -> linear SVM).

Workload (BASELINE.json configs[1]): MCCA alignment of 8 synthetic uECoG patients
(144 trials x 200 time bins x 128 channels each) into a shared latent space, 20-fold CV
cross-patient decode for one target patient.  One "step" = one CV iteration = 20 folds
(20 units), exactly the inner loop of the reference's scripts/aligned_decode_svm_ncv.py:
336-442.  Multi-GPU: every rank runs its own CV iterations (weak scaling, no data-path
collective); per-fold accuracies are all-gathered with NCCL at the end of the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (rank 0).  `value` = folds/s with the patient data resident in HBM;
`e2e` = the same through the public host-buffer call (cv_align_decode), uploads included.
"""
import argparse
import json
import os

# more hardware work queues than the default 8: the streamed e2e path keeps 14+ CUDA streams busy and
# streams that alias onto one queue serialise (measured: 981 folds/s with 2 queues, 1681 with 8, 1774
# with 32).  Must be set before the CUDA context exists.
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PATIENTS = 8
N_FOLDS = 20
METRIC = 'CV align+decode folds/sec (MCCA->PCA->SVM)'
WORKLOAD = ('mcca_8patients_144x200x128_20fold: MCCA(n_comp=30, regs=0.5, pca_var=0.8) -> '
            'PCA(0.8) -> OvR linear SVM, 20 folds per step')
# identical in both arms (the driver compares it)
CONFIG = {'workload': WORKLOAD, 'folds_per_step_per_gpu': N_FOLDS,
          'l2': 'inputs larger than L2: working set per step ~2 GB (20 pooled 1152x6000 matrices + Grams) '
                '>> 126 MB L2, no explicit flush'}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler:
    """SM clocks / throttle reasons while the timed region runs.  Sampled in-process through
    NVML (nvidia_ml_py): spawning nvidia-smi every 200 ms stalls kernel launches for tens of
    milliseconds, which is visible in a sub-second timed region.  Falls back to nvidia-smi."""

    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')
    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

    def __init__(self, index, period=0.1):
        self.index, self.rows, self.stop, self.period = index, [], False, period
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = index
            if vis:
                ids = vis.split(',')
                if index < len(ids) and ids[index].strip().isdigit():
                    phys = int(ids[index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        bits = [getattr(n, 'nvmlClocksThrottleReasonHwSlowdown', 0x8),
                getattr(n, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40),
                getattr(n, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20),
                getattr(n, 'nvmlClocksThrottleReasonSwPowerCap', 0x4)]
        return [str(sm), str(mx)] + ['Active' if (r & b) else 'Not Active' for b in bits]

    def _run(self):
        while not self.stop:
            try:
                if self.nvml is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(['nvidia-smi', '-i', str(self.index),
                                          '--query-gpu=' + self.Q, '--format=csv,noheader,nounits'],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            time.sleep(self.period if self.nvml is not None else 0.5)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        reasons = [n for i, n in enumerate(self.NAMES)
                   if any(len(r) > 2 + i and r[2 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(self.rows), 'source': 'nvml' if self.nvml is not None else 'nvidia-smi'}


def make_data():
    from cross_patient_speech_decoding_b200 import synthetic
    return synthetic.make_patients(N_PATIENTS)


def step_folds(y, step_id):
    from cross_patient_speech_decoding_b200.folds import cv_splits
    import warnings
    np.random.seed(step_id)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return cv_splits(y, N_FOLDS)


# ------------------------------------------------------------------------------ reference
def cpu_sample(pts, n_folds_sample, step_id=0):
    """Times the CPU port of the reference path (oracle/pipeline_port.py) on a bounded sample
    of the same workload.  Returns (folds/s, seconds, accuracy)."""
    from oracle import pipeline_port
    use_all_host_threads()
    folds = step_folds(pts[0][1], step_id)[:n_folds_sample]
    t0 = time.perf_counter()
    ok = tot = 0
    for tr, te in folds:
        yp, _ = pipeline_port.run_fold(pts[0], pts[1:], tr, te, method='mcca', n_comp=30, regs=0.5,
                                       pca_var=0.8)
        ok += int((yp == pts[0][1][te]).sum())
        tot += len(te)
    dt = time.perf_counter() - t0
    return len(folds) / dt, dt, ok / max(tot, 1)


def use_all_host_threads():
    """The CPU arm uses every host core (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass
    return n


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([i.get('num_threads', 1) for i in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank):
    if rank != 0:
        return
    pts = make_data()
    for w in range(args.warmup):
        cpu_sample(pts, 1, 1000 + w)
    t0 = time.perf_counter()
    n = 0
    for s in range(args.steps):
        _, _, _ = cpu_sample(pts, 1, s)
        n += 1
    dt = time.perf_counter() - t0
    val = n / dt
    cores = blas_threads()
    line = {'metric': METRIC, 'value': val, 'unit': 'folds/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / max(n, 1),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'impl': 'reference',
            'config': CONFIG,
            'reference_config': {'step': '1 fold of the 20-fold workload (bounded sample)',
                                 'host': 'numpy/scipy/scikit-learn float64'},
            'cpu_baseline': {'value': val, 'unit': 'folds/s', 'cores': cores, 'kind': 'port',
                             'sample': '%d single-fold steps of the 8-patient 20-fold MCCA workload '
                                       '(oracle/pipeline_port.py; /root/reference is absent on the '
                                       'GPU box)' % n},
            'e2e': {'value': val, 'unit': 'folds/s', 'h2d_bytes_per_step': 0,
                    'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def measure_tf32_peak(dev):
    """TF32 dense matmul throughput of this GPU, measured the way MEASURED_PEAKS.json measured
    bf16 (torch.matmul 8192^3, best of 5, CUDA events): the denominator SURVEY 8(d) asks for the
    Gram kernels.  Library call, outside every timed region."""
    import torch
    try:
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        n = 8192
        a = torch.randn((n, n), device=dev, dtype=torch.float32)
        b = torch.randn((n, n), device=dev, dtype=torch.float32)
        torch.matmul(a, b)
        best = 1e30
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        torch.backends.cuda.matmul.allow_tf32 = old
        del a, b
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    except Exception:
        return None


def load_traffic():
    """profiles/traffic.json: {kernel: {'dram_bytes': read + write of one ncu --set full capture,
    'folds': folds of that launch, 'source': file}} written by profiles/summarize.py."""
    try:
        return json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
    except Exception:
        return {}


def scaled_traffic(traffic, kernel, nfolds, also=None):
    tot = 0.0
    for k in (kernel, also):
        if k is None:
            continue
        t = traffic.get(k)
        if not t:
            return None
        tot += float(t['dram_bytes']) / float(t['folds']) * nfolds
    return tot


def predict_latency(pts):
    """BASELINE configs[4]: per-call latency of the fitted config-2 model from host float64 trials
    to labels (decoders.fused_predict.FusedPredictor: one upload, one kernel, one read-back), batch
    1 and 256, wall clock per call, p50 / p99."""
    try:
        from sklearn.pipeline import make_pipeline
        from cross_patient_speech_decoding_b200.alignment.AlignMCCA import AlignMCCA
        from cross_patient_speech_decoding_b200.decoders.cross_pt_decoders import crossPtDecoder_mcca
        from cross_patient_speech_decoding_b200.decoders.fused_predict import FusedPredictor
        from cross_patient_speech_decoding_b200.decomposition.DimRedReshape import DimRedReshape
        from cross_patient_speech_decoding_b200.decomposition.PCA import PCA
        from cross_patient_speech_decoding_b200.svm import LinearSVC
        Xt, yt, yat = pts[0]
        tr, te = step_folds(yt, 0)[0]
        m = crossPtDecoder_mcca(pts[1:], make_pipeline(DimRedReshape(PCA, n_components=0.8), LinearSVC()),
                                AlignMCCA, n_comp=30, regs=0.5, pca_var=0.8)
        m.fit(Xt[tr], yt[tr], y_align=yat[tr])
        fp = FusedPredictor(m)
        out = {'api': 'decoders.fused_predict.FusedPredictor.predict (host float64 trials -> labels)'}
        for nb in (1, 256):
            Xb = np.ascontiguousarray(np.concatenate([Xt] * 2)[:nb])
            for _ in range(10):
                fp.predict(Xb)
            ts = []
            for _ in range(300 if nb == 1 else 60):
                t0 = time.perf_counter()
                fp.predict(Xb)
                ts.append(1e3 * (time.perf_counter() - t0))
            out['batch%d' % nb] = {'p50_ms': round(float(np.percentile(ts, 50)), 4),
                                   'p99_ms': round(float(np.percentile(ts, 99)), 4),
                                   'trials_per_s': round(nb / (np.median(ts) * 1e-3), 1)}
        return out
    except Exception as e:                                   # the probe never breaks the headline line
        return {'error': str(e)[:120]}


def run_other_config(args, rank, world, local, pts):
    """BASELINE configs[2] / configs[3] through the PRODUCT API, sharded over the ranks the way the
    scripts are (whole CV iterations / subsamples per rank, one all_gather of the labels)."""
    import torch
    import torch.distributed as dist
    from cross_patient_speech_decoding_b200 import cv_align_decode, sharding
    dev = torch.device('cuda', local)

    def timed(fn):
        fn(True)                                            # warm-up with the same structure
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        n = fn(False)
        torch.cuda.synchronize(dev)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return n, float(dt.item())

    lines = []
    y0 = pts[0][1]
    if args.config == 'sweep':
        n_iter = args.steps * world                          # weak scaling: `steps` iterations per GPU
        for method in args.methods.split(','):
            for d in [int(x) for x in args.dims.split(',')]:
                units = []
                for it in range(n_iter):
                    units += step_folds(y0, 5000 + it)
                mine = sharding.shard_units(len(units), N_FOLDS, rank, world)
                kw = dict(method=method, n_comp=d, use_tensor_cores=True, max_batch=args.batch)
                if method == 'mcca':
                    kw.update(regs=0.5, pca_var=0.8)

                def go(warm, kw=kw, units=units, mine=mine):
                    mu = [units[u] for u in (mine[:2 * N_FOLDS] if warm else mine)]
                    res = cv_align_decode(pts[0], pts[1:], mu, **kw)
                    sharding.gather_predictions(mine[:len(mu)], res['y_pred'])
                    return len(units)
                try:
                    n, dt = timed(go)
                    lines.append({'workload': 'config3 sweep: 8 patients, %s, n_comp=%d, 20-fold' % (method, d),
                                  'folds': n, 'value': n / dt})
                except ValueError as e:                      # MCCA n_components above the summed ranks
                    lines.append({'workload': 'config3 sweep: %s n_comp=%d' % (method, d), 'error': str(e)[:80]})
    else:
        from cross_patient_speech_decoding_b200.processing_utils.grid_subsampling import sig_channels_in_windows
        from cross_patient_speech_decoding_b200.processing_utils.subsample_decode import subsample_decode
        chan_map, sig = np.arange(1, 129).reshape(8, 16), np.arange(1, 129)
        subs = [np.sort(np.asarray(s_).ravel()) for s_ in sig_channels_in_windows(chan_map, sig, (4, 8), (2, 4))]
        nsub = args.steps * world                            # subsamples per target, dealt to the ranks

        def go(warm):
            tot = 0
            for tgt in range(1 if warm else N_PATIENTS):
                order = [tgt] + [p for p in range(N_PATIENTS) if p != tgt]
                np.random.seed(300 + tgt)
                pick = [subs[i % len(subs)] for i in range(2 * world if warm else nsub)]
                out = subsample_decode(pts[order[0]], [pts[p] for p in order[1:]], pick, [subs] * 7,
                                       n_folds=N_FOLDS, method='cca', n_comp=0.9, use_tensor_cores=True,
                                       max_batch=args.batch, depth=6)
                tot += N_FOLDS * len(out['accs'])
            return tot
        n, dt = timed(go)
        lines.append({'workload': 'config4 electrode subsampling: %d grid subsamples (32 of 128 channels) x 8 '
                                  'targets x 20 folds, pairwise CCA, host patients uploaded per target' % nsub,
                      'folds': n, 'value': n / dt})
    if rank == 0:
        for ln in lines:
            out = {'metric': METRIC, 'unit': 'folds/s', 'n_gpus': world, 'steps': args.steps,
                   'warmup': 1, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                   'dtype': 'f32', 'data': 'synthetic', 'config': {'workload': ln.pop('workload')},
                   'timing': 'wall clock between device synchronisations, max over ranks, product API '
                             '(host arrays in, labels out)'}
            out.update(ln)
            print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------ ours
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=48)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--no-tc', action='store_true', help='fp32 SIMT pooled Gram instead of tcgen05')
    ap.add_argument('--cpu-folds', type=int, default=2, help='folds in the cpu_baseline sample')
    ap.add_argument('--e2e-steps', type=int, default=None)
    ap.add_argument('--e2e-depth', type=int, default=21, help='steps in flight in the e2e measurement')
    ap.add_argument('--e2e-group', type=int, default=7,
                    help='steps sharing one engine batch in the e2e measurement (replicas: every step still '
                         'uploads its own inputs; 7 x 20 folds = one 140-fold batch)')
    ap.add_argument('--batch', type=int, default=148, help='max folds per engine batch')
    ap.add_argument('--lanes', type=int, default=2, help='execution lanes (batches in flight) of the engine')
    ap.add_argument('--config', default='headline', choices=['headline', 'sweep', 'subsample'],
                    help='headline = BASELINE configs[1] (the driver\'s run); sweep = configs[2] (latent-size '
                         'sweep x methods); subsample = configs[3] (electrode subsampling); the last two run '
                         'through the product API sharded over the ranks')
    ap.add_argument('--dims', default='10,30,60,100', help='--config sweep: latent sizes')
    ap.add_argument('--methods', default='jointpca,cca,mcca', help='--config sweep: alignment methods')
    ap.add_argument('--no-latency', action='store_true', help='skip the config-5 predict latency probe')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return

    # cpu_baseline: rank 0, N = 1 only, before any GPU work is queued (at N > 1 the other ranks
    # would spin in a barrier on the same host cores and contaminate it; the driver's own
    # --impl reference arm covers every N)
    cpu_line = None
    if rank == 0 and world == 1 and args.config == 'headline' and args.cpu_folds > 0:
        pts_cpu = make_data()
        cpu_sample(pts_cpu, 1, 12345)            # warm-up (imports, BLAS threads)
        cpu_val, cpu_dt, cpu_acc = cpu_sample(pts_cpu, args.cpu_folds, 0)
        cpu_line = {'value': cpu_val, 'unit': 'folds/s', 'cores': blas_threads(), 'kind': 'port',
                    'accuracy': cpu_acc,
                    'sample': '%d folds of the same 8-patient 20-fold workload, oracle/pipeline_port.py '
                              '(numpy/scipy/sklearn float64), %.1f s, timed before any GPU work'
                              % (args.cpu_folds, cpu_dt)}
        del pts_cpu
    import torch
    import torch.distributed as dist
    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        dist.barrier()
    from cross_patient_speech_decoding_b200 import cv_align_decode
    from cross_patient_speech_decoding_b200.engine import CVEngine

    pts = make_data()
    y0 = pts[0][1]
    if args.config != 'headline':
        run_other_config(args, rank, world, local, pts)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    kw = dict(method='mcca', n_comp=30, regs=0.5, pca_var=0.8, decoder_var=0.8,
              use_tensor_cores=not args.no_tc, max_batch=args.batch, n_lanes=args.lanes)
    eng = CVEngine(pts[0], pts[1:], device='cuda:%d' % local, **kw)
    dev = eng.ctx.device

    def run_steps(sids):
        """K steps = K CV iterations of 20 folds; the iterations are independent units, so all
        their folds go to the engine in one call and are batched `--batch` at a time."""
        folds = []
        for sid in sids:
            folds += step_folds(y0, sid)
        res = eng.run(folds)
        accs = []
        for i in range(len(sids)):
            fs = folds[i * N_FOLDS:(i + 1) * N_FOLDS]
            ps = res['y_pred'][i * N_FOLDS:(i + 1) * N_FOLDS]
            ok = sum(int((p == y0[te]).sum()) for p, (_, te) in zip(ps, fs))
            accs.append(ok / sum(len(te) for _, te in fs))
        return res, accs

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # size every workspace for a full engine batch before anything is timed
    # (two full batches: both execution lanes allocate their workspaces)
    big = []
    while len(big) < args.lanes * args.batch:
        big += step_folds(y0, 20_000 + len(big))
    eng.run(big[:args.lanes * args.batch])
    if args.warmup:
        run_steps([10_000 + rank * 1000 + w for w in range(args.warmup)])
    sync_all()
    eng.ctx.lib.cpsd_reset_launch_count()
    eng.stats['host_pack_ms'] = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    accs, h2d, d2h = [], 0, 0
    with ClockSampler(local) as clk:
        sync_all()
        e0.record(torch.cuda.current_stream(dev))
        res, accs = run_steps([rank * 100_000 + s for s in range(args.steps)])
        h2d += res['h2d_bytes']
        d2h += res['d2h_bytes']
        # the one collective of the path: gather fixed-size per-iteration records
        from cross_patient_speech_decoding_b200.sharding import gather_records
        rec = np.array([[rank * args.steps + i, int(round(a * 1e6))] for i, a in enumerate(accs)],
                       dtype=np.int32).reshape(-1, 2)
        allrec = gather_records(rec, device=dev)
        acc_all = allrec[:, 1] / 1e6
        e1.record(torch.cuda.current_stream(dev))
        sync_all()
    ms = e0.elapsed_time(e1)
    host_pack_timed = eng.stats.get('host_pack_ms', 0.0)
    launches = int(eng.ctx.lib.cpsd_launch_count())
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    folds_total = world * args.steps * N_FOLDS
    value = folds_total / (ms_max * 1e-3)

    # ---- end to end through the public host-buffer API: every step uploads its inputs (8 patients,
    # float64, pinned host memory), runs its 20 folds and reads its predictions back.  The
    # streaming entry point keeps `depth` independent steps in flight on separate CUDA streams;
    # the blocking one-call-per-step form is reported beside it.
    from cross_patient_speech_decoding_b200 import cv_align_decode_stream
    e2e_steps = args.e2e_steps or max(args.steps, 12)
    host_pts = [(torch.from_numpy(np.ascontiguousarray(X)).pin_memory(), y, ya)
                for X, y, ya in pts]
    kw_e2e = dict(kw, max_batch=N_FOLDS, n_lanes=2)

    def jobs(n, seed0):
        for s in range(n):
            yield host_pts[0], host_pts[1:], step_folds(y0, seed0 + s)

    for _ in cv_align_decode_stream(jobs(args.e2e_depth + 2, 77), depth=args.e2e_depth, device='cuda:%d' % local,
                                    group=args.e2e_group, **kw_e2e):
        pass
    sync_all()
    t0 = time.perf_counter()
    e2e_h2d = e2e_d2h = 0
    e2e_ok = e2e_tot = 0
    for s, r in enumerate(cv_align_decode_stream(jobs(e2e_steps, 500 + rank * 1000),
                                                 depth=args.e2e_depth, device='cuda:%d' % local,
                                                 group=args.e2e_group, **kw_e2e)):
        e2e_h2d += r['h2d_bytes']
        e2e_d2h += r['d2h_bytes']
    sync_all()
    e2e_dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
    e2e_val = world * e2e_steps * N_FOLDS / float(e2e_dt.item())
    # blocking form (one call per step, nothing overlapped)
    nblk = min(e2e_steps, 4)
    cv_align_decode(host_pts[0], host_pts[1:], step_folds(y0, 78), device='cuda:%d' % local, **kw_e2e)
    sync_all()
    t0 = time.perf_counter()
    for s in range(nblk):
        cv_align_decode(host_pts[0], host_pts[1:], step_folds(y0, 700 + rank * 1000 + s),
                        device='cuda:%d' % local, **kw_e2e)
    sync_all()
    e2e_blocking = world * nblk * N_FOLDS / (time.perf_counter() - t0)

    # ---- stage breakdown + roofline of the dominant tensor / HBM kernels (profiling pass)
    # one engine batch of the size the timed region used, CUDA events between the stages
    nprof = min(args.batch, args.steps * N_FOLDS)
    nb = -(-args.steps * N_FOLDS // nprof)
    nprof = -(-args.steps * N_FOLDS // nb)
    pf = []
    while len(pf) < nprof:
        pf += step_folds(y0, 424242 + len(pf))
    eng.profile = True
    eng.run(pf[:nprof])
    stages_batch = eng.collect_marks()
    eng.profile = False
    stages = {k: v * N_FOLDS / nprof for k, v in stages_batch.items()}
    pk, pk_kind = peaks()
    tf32_peak = measure_tf32_peak(dev)
    traffic = load_traffic()
    n_all = 1152
    F = 200 * 30
    stage_total = sum(stages_batch.values()) or 1.0
    # k_gram_tc alone: the engine's profile mode records an event between the hi/lo split and the
    # MMA kernel (cpsd_gram_nt_tc_probe), so 'pool_split' / 'pool_gram' are the two kernels' own times
    gram_ms = stages_batch.get('pool_gram', float('nan'))
    split_ms = stages_batch.get('pool_split', 0.0)
    gram_flops = 2.0 * n_all * n_all * F * nprof              # algorithmic, per launch (nprof folds)
    gram_tf = gram_flops / (gram_ms * 1e-3) / 1e12
    split_bytes = 3.0 * n_all * F * 4 * nprof                 # read the pooled matrix once, write hi and lo
    peak_tf = pk.get('bf16_tflops_sustained', pk.get('bf16_tflops'))
    proj_ms = stages_batch.get('project_pool', float('nan'))
    proj_bytes = nprof * (sum(p[0].size for p in pts) * 4 + n_all * F * 4)   # read X once, write pooled
    roofline = {'kernel': 'k_gram_tc (pooled Gram of the centred matrix, tcgen05 kind::tf32, 3xTF32 on the '
                          'hi/lo operands k_split_tf32_batched writes: see roofline_split)' if not args.no_tc
                else 'k_gram_nt (pooled Gram, fp32 SIMT)',
                'bound': 'tensor', 'achieved': gram_tf, 'peak': peak_tf, 'unit': 'TFLOP/s',
                'frac': gram_tf / peak_tf,
                # dram__bytes_read + write per fold from the committed ncu capture
                # (profiles/traffic.json, written by profiles/summarize.py), times this launch's folds
                'traffic': scaled_traffic(traffic, 'k_gram_tc', nprof),
                'peak_source': '%s bf16_tflops_sustained (MEASURED_PEAKS.json)' % pk_kind,
                # the pipe this kernel runs on: TF32 dense peak measured in this run (torch.matmul,
                # allow_tf32, 8192^3, best of 5); 3xTF32 issues 3 MMAs per algorithmic product and only
                # the 45 upper tiles of the 9 x 9 tile grid are computed
                'tf32_dense_tflops_measured': tf32_peak,
                'frac_of_tf32_peak': (gram_tf / tf32_peak) if tf32_peak else None,
                'executed_tf32_tflops': gram_tf * 3.0 * 45.0 / 81.0,
                'executed_frac_of_tf32_peak': (gram_tf * 3.0 * 45.0 / 81.0 / tf32_peak) if tf32_peak else None,
                'algorithmic_flops_per_launch': gram_flops, 'launch_ms': gram_ms,
                'launch_ms_with_split': gram_ms + split_ms,
                'achieved_with_split': gram_flops / ((gram_ms + split_ms) * 1e-3) / 1e12,
                'share_of_step': gram_ms / stage_total}
    roofline_split = {'kernel': 'k_split_tf32_batched (centre + split the pooled matrix into tf32 hi / lo)',
                      'bound': 'hbm', 'achieved': split_bytes / (split_ms * 1e-3) / 1e9 if split_ms else None,
                      'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                      'frac': split_bytes / (split_ms * 1e-3) / 1e9 / pk['hbm_gbs'] if split_ms else None,
                      'traffic': scaled_traffic(traffic, 'k_split_tf32_batched', nprof),
                      'algorithmic_bytes_per_launch': split_bytes, 'launch_ms': split_ms,
                      'share_of_step': split_ms / stage_total}
    proj_traffic = scaled_traffic(traffic, 'k_proj_tc', nprof)
    # compulsory traffic of the launch: X hi/lo once + every pooled matrix written once
    proj_real = proj_traffic if proj_traffic else (2 * sum(p[0].size for p in pts) * 4 + nprof * n_all * F * 4)
    roofline_hbm = {'kernel': 'k_proj_tc_prep + k_proj_tc (project all trials of all patients into the '
                              'pooled matrices, tcgen05 3xTF32 + TMA)',
                    'traffic': proj_traffic,
                    'bound': 'hbm', 'achieved': proj_real / (proj_ms * 1e-3) / 1e9,
                    'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                    'frac': proj_real / (proj_ms * 1e-3) / 1e9 / pk['hbm_gbs'],
                    'note': 'achieved = the DRAM bytes the kernel really moves (ncu dram read + write of '
                            'the committed capture, scaled to this launch; X hi/lo once + the pooled '
                            'matrices once) / its CUDA-event time.  SURVEY 8(d) counts every fold '
                            're-reading every patient (`survey_bytes_per_launch`, `survey_gbs`): the '
                            'kernel shares each X tile between all folds of the batch instead',
                    'survey_bytes_per_launch': proj_bytes,
                    'survey_gbs': proj_bytes / (proj_ms * 1e-3) / 1e9,
                    'launch_ms': proj_ms, 'share_of_step': proj_ms / stage_total}

    latency = None
    if rank == 0 and world == 1 and not args.no_latency:
        latency = predict_latency(pts)
    # whole-step roofline against SURVEY 8(d)'s per-fold work without reuse: 30 GF of tensor-shaped
    # work and 0.29 GB of streaming traffic per fold
    step_roofline = {'tensor_gflop_per_fold': 30.0, 'hbm_gb_per_fold': 0.29,
                     'tflops': 30.0e9 * value / 1e12, 'frac_of_bf16_sustained': 30.0e9 * value / 1e12 / peak_tf,
                     'frac_of_tf32_measured': (30.0e9 * value / 1e12 / tf32_peak) if tf32_peak else None,
                     'gbs': 0.29 * value, 'frac_of_hbm': 0.29 * value / pk['hbm_gbs'],
                     'note': 'the step is a chain of latency-bound small solvers (tile Jacobi, Cholesky-QR, '
                             'Newton SVM) around two dense kernels; see stages_ms_per_step'}
    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': 'folds/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_max / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': CONFIG,
            'engine_config': {'engine_batch_folds': args.batch,
                              'parallelism': 'folds sharded over %d GPU(s), one NCCL all_gather of '
                                             'accuracies' % world,
                              'precision': 'fp32 storage; fp64 scatter + eigen-solver for the alignment '
                                           'PCA stages; 3xTF32 tcgen05 projection and pooled Gram; top-k '
                                           'subspace iteration (TF32 then 3xTF32 tcgen05) for the decoder '
                                           'PCA; fp64 Newton SVM',
                              'lanes': args.lanes},
            'e2e': {'value': e2e_val, 'unit': 'folds/s',
                    'h2d_bytes_per_step': e2e_h2d // max(e2e_steps, 1),
                    'd2h_bytes_per_step': e2e_d2h // max(e2e_steps, 1), 'steps': e2e_steps,
                    'api': 'cross_patient_speech_decoding_b200.cv_align_decode_stream (host float64 '
                           'arrays in pinned memory -> predictions; every step uploads its own '
                           'inputs; %d steps in flight, %d steps per engine batch)'
                           % (args.e2e_depth, args.e2e_group),
                    'blocking_value': e2e_blocking,
                    'blocking_api': 'cv_align_decode, one blocking call per step'},
            'gpu_launches': launches,
            'h2d_bytes_per_step': h2d // args.steps, 'd2h_bytes_per_step': d2h // args.steps,
            'clocks': clk.summary(),
            'roofline': roofline, 'roofline_hbm': roofline_hbm, 'roofline_split': roofline_split,
            'step_roofline': step_roofline,
            'stages_ms_per_step': {k: round(v, 3) for k, v in stages.items()},
            'stages_batch_folds': nprof,
            'reuse': eng.stats.get('view_solves', 0) and
            'cross-patient view statistics solved once per (view, shared class set) and reused '
            'across folds: %d of %d (fold, view) eigen-problems solved'
            % (eng.stats.get('view_solves', 0), eng.stats.get('view_problems', 0)),
            'accuracy_mean': float(np.mean(acc_all)),
            'host_pack_ms_per_step': round(host_pack_timed / args.steps, 3),
            'cpu_baseline': cpu_line or {'value': None, 'unit': 'folds/s', 'cores': blas_threads(),
                                         'kind': 'port', 'sample': 'measured at N = 1 only (rank 0)'},
            'predict_latency': latency,
            'svm_unconverged': int(eng.stats.get('svm_unconverged', 0)),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
