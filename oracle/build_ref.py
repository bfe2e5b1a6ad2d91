"""Compile the reference's own native CUDA sources (unmodified, from where they lie under /root/reference)
into oracle/_ref/libtdrn_ref_native.so -- TEST INFRASTRUCTURE ONLY.

    python -m oracle.build_ref          (build container only: /root/reference must exist)

What compiles: utils/deformconv/deform_conv_cuda_kernel.cu (the bilinear im2col kernel K1 and its launcher;
plain CUDA runtime, no THC) and utils/nms/nms_kernel.cu (K6 + host `_nms`).  What does not: the THC host file
utils/deformconv/deform_conv_cuda.c (needs torch 0.4's THC) and utils/nms/cpu_nms.pyx (Cython 0.25 / NumPy 1
source) -- see DESIGN.md section 5.  nvcc is invoked directly on the files (the reference's make.sh / build.py use
torch.utils.ffi and distutils and are not run).  Same flags as the reference's make.sh:11 (none besides -fPIC:
default -fmad=true) plus the sm_100a target.  Output is git-ignored but travels to the GPU box with gpurun.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get('TDRN_REFERENCE_ROOT', '/root/reference')
OUT_DIR = os.path.join(HERE, '_ref')
OUT = os.path.join(OUT_DIR, 'libtdrn_ref_native.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')


def sources():
    return [os.path.join(REFERENCE_ROOT, 'utils', 'deformconv', 'deform_conv_cuda_kernel.cu'),
            os.path.join(REFERENCE_ROOT, 'utils', 'nms', 'nms_kernel.cu')]


def available():
    return all(os.path.exists(s) for s in sources())


def build(force=False):
    """-> path of the .so, or None when the reference checkout is absent (GPU box: the prebuilt file is used)."""
    if not available():
        return OUT if os.path.exists(OUT) else None
    wrap = os.path.join(HERE, 'ref_native', 'wrap.cu')
    deps = sources() + [wrap]
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [NVCC, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-Xcompiler', '-fPIC', '-std=c++11',
           '-I', os.path.join(REFERENCE_ROOT, 'utils', 'deformconv'), '-I', os.path.join(REFERENCE_ROOT, 'utils', 'nms'),
           '-o', OUT] + deps + ['-lcudart_static', '-ldl', '-lrt', '-lpthread']
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    print(build(force=True))
