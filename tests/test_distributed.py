"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: iteration sharding + the single
all_gather of per-unit records."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

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
