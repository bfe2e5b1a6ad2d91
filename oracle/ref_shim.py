"""Import the *real* reference Python (model/*.py, layers/*) in place from /root/reference.

TEST INFRASTRUCTURE ONLY, and only usable in the build container: /root/reference does not
exist on the GPU box, so nothing that runs there (pytest -m gpu, smoke(), bench.py) may call
this module.  It is used by tests/test_oracle_vs_reference.py (skipped when the checkout is
absent) and by oracle/make_golden.py to generate tests/golden/*.npz.

Nothing is copied: the reference files are executed where they lie.  Six shims are needed to
run 2018 PyTorch-0.4 code under torch 2.x on CPU:
  1. utils.nms.cpu_nms      -> the reference's own cpu_nms.pyx built by oracle/build_ref_nms.py (two dtype tokens
                               respelled for NumPy 2); oracle.nms_ref.cpu_nms only if that build is impossible
  2. utils.nms.gpu_nms      -> stub (needs a GPU; not on the Detect path)
  3. utils._ext.deform_conv -> stub module (import at model/networks.py:8)
  4. torch.cuda.FloatTensor -> CPU fp32 tensor factory (default arg at detection.py:25)
  5. Tensor.cuda()          -> identity (hard .cuda() at detection.py:60)
  6. model.networks.conv_offset2d -> oracle.deform_conv_ref.conv_offset2d (legacy non-static
     autograd Function + CUDA-only native op)
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('TDRN_REFERENCE_ROOT', '/root/reference')

_loaded = None


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'model'))


def load():
    """Returns a namespace with the reference modules; idempotent."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError('reference checkout not found at %s' % REFERENCE_ROOT)
    import torch
    from . import nms_ref, deform_conv_ref

    sys.dont_write_bytecode = True                      # /root/reference is read-only
    # our own package must not shadow the reference's top-level names
    for name in ('model', 'layers', 'utils', 'data'):
        if name in sys.modules and not getattr(sys.modules[name], '__file__', '').startswith(REFERENCE_ROOT):
            raise RuntimeError('module %r already imported from elsewhere' % name)

    m = types.ModuleType('utils.nms.cpu_nms')
    real = None
    try:                                                # the reference's own Cython NMS, when oracle/build_ref_nms.py built it
        from . import build_ref_nms
        build_ref_nms.build()
        real = build_ref_nms.load()
    except Exception:                                   # no Cython / no compiler: fall back to the restatement
        real = None
    m.cpu_nms = real.cpu_nms if real is not None else (lambda dets, thresh: nms_ref.cpu_nms(dets, thresh))
    m.cpu_soft_nms = None
    sys.modules['utils.nms.cpu_nms'] = m
    m = types.ModuleType('utils.nms.gpu_nms')
    m.gpu_nms = None
    sys.modules['utils.nms.gpu_nms'] = m
    ext = types.ModuleType('utils._ext')
    ext.deform_conv = types.ModuleType('utils._ext.deform_conv')
    sys.modules['utils._ext'] = ext
    sys.modules['utils._ext.deform_conv'] = ext.deform_conv
    if 'cv2' not in sys.modules:
        try:
            import cv2  # noqa: F401
        except Exception:
            sys.modules['cv2'] = types.ModuleType('cv2')  # imported, never used on the hot path

    torch.cuda.FloatTensor = lambda v: torch.tensor(v, dtype=torch.float32)
    torch.Tensor.cuda = lambda self, *a, **k: self

    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import importlib
        ns = types.SimpleNamespace()
        ns.networks = importlib.import_module('model.networks')
        ns.networks.conv_offset2d = deform_conv_ref.conv_offset2d
        ns.drn_vgg = importlib.import_module('model.dualrefinedet_vggbn')
        ns.drn_mobilenet = importlib.import_module('model.dualrefinedet_mobilenet')
        ns.refinedet_vgg = importlib.import_module('model.refinedet_vgg')
        ns.ssd4scale_vgg = importlib.import_module('model.ssd4scale_vgg')
        ns.ssd4scale_mobile = importlib.import_module('model.ssd4scale_mobile')
        ns.layers = importlib.import_module('layers')
        ns.box_utils = importlib.import_module('layers.box_utils')
        ns.Detect = ns.layers.Detect
        ns.PriorBox = ns.layers.PriorBox
        ns.L2Norm = ns.layers.L2Norm
        ns.py_cpu_nms = importlib.import_module('utils.nms.py_cpu_nms').py_cpu_nms
        ns.nms_wrapper = importlib.import_module('utils.nms_wrapper')
        ns.cpu_nms_is_reference_build = real is not None
    finally:
        sys.path.remove(REFERENCE_ROOT)
    _loaded = ns
    return ns
