"""Compile the reference's OWN CPU NMS (utils/nms/cpu_nms.pyx, the function Detect runs: utils/nms_wrapper.py:23-31
with force_cpu=True) into oracle/_ref/cpu_nms<EXT_SUFFIX> -- TEST INFRASTRUCTURE ONLY.

    python -m oracle.build_ref_nms          (build container only: /root/reference must exist)

The 2017 source does not compile as it lies under Cython 3 / NumPy 2: it spells the index dtype `np.int_t` (buffer
type, cpu_nms.pyx:25,28) and `np.int` (allocation, :29), names NumPy removed.  The recipe therefore cythonizes a
TEMPORARY copy (in a tempfile directory, never in this repository) in which exactly those two tokens are respelled
`np.intp_t` / `np.intp` -- the same 64-bit signed integer `np.int_t` was on Linux x86-64 -- and nothing else: the sort,
the float32 area / IoU arithmetic and the `ovr >= thresh` comparison are the reference's own lines.  The reference's
setup.py (distutils + a CUDA toolchain locator) is not run; gcc is invoked directly on the generated C file.
Output is git-ignored but travels to the GPU box with gpurun (same image, same CPython / NumPy ABI).
"""
import os
import re
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get('TDRN_REFERENCE_ROOT', '/root/reference')
SRC = os.path.join(REFERENCE_ROOT, 'utils', 'nms', 'cpu_nms.pyx')
OUT_DIR = os.path.join(HERE, '_ref')
OUT = os.path.join(OUT_DIR, 'cpu_nms' + (sysconfig.get_config_var('EXT_SUFFIX') or '.so'))

RESPELL = ((r'np\.int_t\b', 'np.intp_t'), (r'dtype=np\.int\)', 'dtype=np.intp)'))


def available():
    return os.path.exists(SRC)


def build(force=False):
    """-> path of the extension module, or None when neither the reference checkout nor a prebuilt file is there."""
    if not available():
        return OUT if os.path.exists(OUT) else None
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(SRC), os.path.getmtime(__file__)):
        return OUT
    import numpy
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix='tdrn_ref_nms_')
    try:
        text = open(SRC).read()
        for pat, rep in RESPELL:
            text, n = re.subn(pat, rep, text)
            if n == 0:
                raise RuntimeError('cpu_nms.pyx does not contain %r any more: review oracle/build_ref_nms.py' % pat)
        pyx = os.path.join(tmp, 'cpu_nms.pyx')
        with open(pyx, 'w') as f:
            f.write(text)
        subprocess.check_call([sys.executable, '-m', 'cython', '-3', pyx], cwd=tmp)
        cmd = ['gcc', '-O2', '-shared', '-fPIC', '-w', '-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION',
               '-I', sysconfig.get_paths()['include'], '-I', numpy.get_include(),
               os.path.join(tmp, 'cpu_nms.c'), '-o', OUT]
        subprocess.check_call(cmd)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return OUT


def load():
    """The compiled module (functions cpu_nms(dets, thresh) and cpu_soft_nms(...)), or None if it was never built."""
    if not os.path.exists(OUT):
        return None
    import importlib.util
    # the module's init symbol is PyInit_cpu_nms: load it under its own name (it is NOT put into sys.modules)
    spec = importlib.util.spec_from_file_location('cpu_nms', OUT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    print(build(force=True))
