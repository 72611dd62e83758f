"""Build libdqn_b200.so in-tree with nvcc for sm_100a (no torch involvement: the library is a plain C ABI)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libdqn_b200.so')
SOURCES = ['api.cu', 'replay.cu', 'sumtree.cu', 'nn.cu', 'heads.cu', 'optim.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
    hdrs.append(os.path.join(HERE, '..', 'include', 'dqn_b200.h'))
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest(srcs + hdrs):
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    procs, objs = [], []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + '.o')
        objs.append(o)
        if not force and os.path.exists(o) and os.path.getmtime(o) >= _newest([s] + hdrs):
            continue
        cmd = [nvcc] + NVCC_FLAGS + ['-c', s, '-o', o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = None
    for s, p in procs:
        out, _ = p.communicate()
        with open(os.path.join(objdir, os.path.basename(s) + '.ptxas.log'), 'w') as f:
            f.write(out)
        if verbose:
            print(out)
        if p.returncode != 0:
            sys.stderr.write(out)
            failed = s
    if failed:
        raise RuntimeError('nvcc failed on %s' % failed)
    subprocess.check_call([nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'])
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
