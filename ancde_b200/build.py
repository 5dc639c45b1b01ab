"""Builds libancde_b200.so (hand-written sm_100a kernels + the C ABI of include/ancde_b200.h) in-tree.

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
SOURCES = ['api.cu', 'spline.cu', 'heads.cu', 'ncde_host.cu', 'ncde_fwd.cu', 'ncde_lane.cu', 'ncde_bwd.cu', 'ncde_bwd4.cu', 'ncde_wgrad.cu', 'ncde_wgrad_mma.cu', 'ncde_wgrad_umma.cu', 'ncde_hidden_wgrad_umma.cu', 'ncde3_host.cu', 'ncde3_fwd.cu', 'ncde3_fwd1.cu']
LIB = os.path.join(HERE, 'libancde_b200.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _nvcc():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found; cannot build libancde_b200.so')
    return nvcc


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 'ancde_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=()):
    """Compiles every CUDA source for sm_100a and links the shared library. Returns its path."""
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objs = []
    log = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src[:-3] + '.o')
        cmd = [nvcc] + NVCC_FLAGS + ['-D' + x for x in defines] + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, pr in procs:
        out, _ = pr.communicate()
        log.append(out)
        if pr.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError('nvcc failed on %s' % src)
    cmd = [nvcc, '-shared', '-o', LIB] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError('link failed')
    with open(os.path.join(CSRC, 'build.log'), 'w') as f:
        f.write('\n'.join(log))
    if verbose:
        print('\n'.join(log))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv,
                defines=[a[2:] for a in sys.argv if a.startswith('-D')]))
