"""Build tdrn_b200/libtdrn_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m tdrn_b200.build [--force] [--probe]

``--probe`` (and build_probe()) also builds tdrn_b200/libtdrn_probe.so from csrc/probe/: the hardware probes behind
scripts/umma_*.py and scripts/tma_f32_probe.py (tcgen05.mma issue rates, UMMA descriptor addressing, fp32 TMA boxes).  They
are development aids with their own ``tdrn_debug_*`` exports, linked against the product library but never part of it.

The shared library is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ_DIR = os.path.join(HERE, '_obj')
LIB = os.path.join(HERE, 'libtdrn_b200.so')
PROBE_LIB = os.path.join(HERE, 'libtdrn_probe.so')
PROBE_SRC = os.path.join(CSRC, 'probe')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr'] + os.environ.get('TDRN_NVCC_EXTRA', '').split()


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hdrs.append(os.path.join(HERE, '..', 'include', 'tdrn_b200.h'))
    return max(os.path.getmtime(h) for h in hdrs)


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_m = _deps_mtime()
    jobs = []
    for s in sources():
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ_DIR, s[:-3] + '.o')
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_m):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
        return r.stderr

    with ThreadPoolExecutor(max_workers=max(1, min(8, len(jobs)))) as ex:
        logs = list(ex.map(compile_one, jobs))
    if verbose:
        for l in logs:
            sys.stderr.write(l)
    objs = [os.path.join(OBJ_DIR, s[:-3] + '.o') for s in sources()]
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB] + objs + ['-lcudart_static', '-ldl', '-lrt', '-lpthread']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    return LIB


def build_probe(force=False):
    """csrc/probe/*.cu -> libtdrn_probe.so (links libtdrn_b200.so for the tensor-map encoder and the error plumbing)."""
    build()
    srcs = sorted(os.path.join(PROBE_SRC, f) for f in os.listdir(PROBE_SRC) if f.endswith('.cu'))
    newest = max([os.path.getmtime(f) for f in srcs] + [_deps_mtime()])
    if force or not os.path.exists(PROBE_LIB) or os.path.getmtime(PROBE_LIB) < newest:
        cmd = [NVCC] + FLAGS + ['-shared', '-o', PROBE_LIB] + srcs + ['-L' + HERE, '-ltdrn_b200', '-Xlinker', '-rpath=$ORIGIN',
                                                                     '-lcudart_static', '-ldl', '-lrt', '-lpthread']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('probe build failed:\n%s\n%s' % (r.stdout, r.stderr))
    return PROBE_LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
    if '--probe' in sys.argv:
        print(build_probe(force='--force' in sys.argv))
