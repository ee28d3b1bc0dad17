"""Builds libcpsd_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libcpsd_b200.so')
SOURCES = ['api.cu', 'jacobi.cu', 'gemm.cu', 'stream.cu', 'svm.cu', 'cca.cu', 'tc_gram.cu', 'subspace.cu', 'tc_proj.cu', 'svc.cu', 'predict.cu', 'solve64.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
EXTRA = os.environ.get('CPSD_NVCC_FLAGS', '').split()
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '--use_fast_math=false' if False else '-Xptxas', '-v']


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library; returns its path."""
    if not force and not _stale():
        return LIB
    objs = []
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    procs = []
    for s in srcs:
        o = os.path.join(CSRC, s.replace('.cu', '.o'))
        cmd = [NVCC] + FLAGS + EXTRA + ['-I', CSRC, '-c', os.path.join(CSRC, s), '-o', o]
        procs.append((s, o, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                             text=True)))
    log = []
    for s, o, p in procs:
        out, _ = p.communicate()
        log.append('== %s ==\n%s' % (s, out))
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError('nvcc failed on %s' % s)
        objs.append(o)
    cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-lcudart', '-lcuda']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError('link failed')
    with open(os.path.join(HERE, 'build.log'), 'w') as f:
        f.write('\n'.join(log))
    if verbose:
        print('\n'.join(log))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
