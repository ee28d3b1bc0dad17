"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: iteration sharding + the single
all_gather of per-unit records."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_iterations_partition():
    from cross_patient_speech_decoding_b200.sharding import shard_iterations
    for n, w in [(50, 8), (7, 2), (3, 4), (16, 1)]:
        got = [shard_iterations(n, r, w) for r in range(w)]
        assert sorted(sum(got, [])) == list(range(n))
        assert max(len(g) for g in got) - min(len(g) for g in got) <= 1


def test_gather_records_single_process():
    from cross_patient_speech_decoding_b200.sharding import gather_records
    rec = np.array([[3, 1, 2], [1, 9, 9]], dtype=np.int32)
    out = gather_records(rec)
    assert out[:, 0].tolist() == [1, 3]


WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, %r)
    from cross_patient_speech_decoding_b200.sharding import gather_records, shard_iterations
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    n_iter, n_folds = 5, 4
    mine = shard_iterations(n_iter, rank, world)
    rows = []
    for it in mine:
        for f in range(n_folds):
            uid = it * n_folds + f
            rows.append([uid, uid * 7 %% 5, 8])         # unit id, n_correct, n_test
    rec = np.array(rows, dtype=np.int32).reshape(-1, 3)
    out = gather_records(rec)
    assert out.shape == (n_iter * n_folds, 3), out.shape
    assert out[:, 0].tolist() == list(range(n_iter * n_folds))
    assert (out[:, 1] == np.arange(n_iter * n_folds) * 7 %% 5).all()
    dist.barrier()
    dist.destroy_process_group()
    print('rank', rank, 'ok', len(mine))
''')


def test_two_rank_gloo_gather(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % ROOT)
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
                        '--nproc-per-node=2', '--master-addr', '127.0.0.1', '--master-port',
                        str(port), str(script)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count('ok') == 2


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _torchrun(script, nproc, timeout=600, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
                           '--nproc-per-node=%d' % nproc, '--master-addr', '127.0.0.1', '--master-port',
                           str(_free_port()), str(script)], capture_output=True, text=True,
                          timeout=timeout, env=e)


def test_shard_units_keeps_iterations_whole():
    from cross_patient_speech_decoding_b200.sharding import shard_units
    for n_iter, n_folds, world in [(50, 20, 8), (5, 4, 2), (3, 7, 4)]:
        got = [shard_units(n_iter * n_folds, n_folds, r, world) for r in range(world)]
        assert sorted(sum(got, [])) == list(range(n_iter * n_folds))
        for g in got:
            assert len(g) % n_folds == 0
            for i in range(0, len(g), n_folds):
                assert g[i] % n_folds == 0 and g[i:i + n_folds] == list(range(g[i], g[i] + n_folds))


SCRIPT_WORKER = textwrap.dedent('''
    import os, pickle, sys
    import numpy as np
    sys.path.insert(0, %(root)r)
    sys.path.insert(0, os.path.join(%(root)r, 'tests', 'golden'))
    from cross_patient_speech_decoding_b200.scripts import aligned_decode_svm_ncv as sc

    def fake_units(target, cross, units, **kw):
        """CPU stand-in for cv_align_decode: nearest class mean of the time-averaged trials."""
        X, y = target[0].mean(axis=1), np.asarray(target[1])
        out = []
        for tr, te in units:
            cls = np.unique(y[tr])
            mu = np.stack([X[tr][y[tr] == c].mean(axis=0) for c in cls])
            d = ((X[te][:, None, :] - mu[None]) ** 2).sum(axis=2)
            out.append(cls[np.argmin(d, axis=1)])
        return {'y_pred': out}

    run_units = fake_units if %(fake)r else None
    out_file = %(out)r + ('.w%%s' %% os.environ.get('WORLD_SIZE', '1'))
    sc.run(dict(vars(sc.init_parser().parse_args(
        ['-pt', 'S1', '-pi', '1', '-po', 'True', '-a', 'True', '-c', 'False', '-f', out_file,
         '--data_file', %(data)r, '--seed', '11', '--n_iter', '5', '--n_folds', '4', '--decoder',
         %(decoder)r]))), run_units=run_units)
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
    print('rank', os.environ.get('RANK', '0'), 'done')
''')


def _script_roundtrip(tmp_path, fake, decoder, env=None):
    import pickle
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
    import make_golden_script as mg
    data = tmp_path / 'data.pkl'
    with open(data, 'wb') as fh:
        pickle.dump(mg.data_dict(), fh, protocol=-1)
    out = str(tmp_path / 'res.pkl')
    script = tmp_path / 'script_worker.py'
    script.write_text(SCRIPT_WORKER % dict(root=ROOT, fake=fake, out=out, data=str(data), decoder=decoder))
    r1 = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=900,
                        env=dict(os.environ, **(env or {})))
    assert r1.returncode == 0, r1.stdout + r1.stderr
    r2 = _torchrun(script, 2, timeout=900, env=env)
    assert r2.returncode == 0, r2.stdout + r2.stderr
    with open(out + '.w1', 'rb') as fh:
        a = pickle.load(fh)
    with open(out + '.w2', 'rb') as fh:
        b = pickle.load(fh)
    assert a['params'] == b['params']
    for k in ('y_true', 'y_pred', 'wrong_trs'):
        assert len(a[k]) == len(b[k]) == 5
        for x, y in zip(a[k], b[k]):
            assert np.array_equal(np.asarray(x), np.asarray(y)), k
    assert np.allclose(a['accs'], b['accs'])
    return a


def test_script_sharded_over_two_ranks_writes_same_pickle(tmp_path):
    """scripts/aligned_decode_svm_ncv under torchrun (world 2, gloo, CPU stand-in for the engine):
    whole CV iterations per rank, one all_gather of the labels, rank 0 writes the pickle --
    identical to the single-process one (aligned_decode_svm_ncv.py:332-456)."""
    _script_roundtrip(tmp_path, True, 'svc_rbf')


@pytest.mark.gpu
def test_script_sharded_over_two_ranks_on_gpu(lib_built, tmp_path):
    """The same with the real engine: two ranks (gloo rendezvous, both on cuda:0 -- the test box
    has one GPU; under nccl every rank owns its own) reproduce the single-process pickle."""
    _script_roundtrip(tmp_path, False, 'linear', env={'CPSD_DIST_BACKEND': 'gloo'})
