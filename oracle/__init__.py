"""oracle/ -- CPU restatement of the TDRN DualRefineDet inference hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from here, and only as the checker.
The product (``tdrn_b200``) never imports ``oracle`` and has no CPU fallback.

Contents (each function cites the reference file:line it restates; paths are
relative to the upstream reference checkout):

* ``deform_conv_ref``  -- pure-PyTorch restatement of the deformable conv forward
  (utils/deformconv/deform_conv_cuda_kernel.cu:16-51,157-208 + deform_conv_cuda.c:141-193)
* ``nms_ref``          -- NumPy restatement of utils/nms/cpu_nms.pyx:17-68 (pinned tie rule)
* ``detect_ref``       -- restatement of layers/functions/detection.py:25-70 and
  layers/box_utils.py:16-25,176-195, layers/functions/prior_box.py:33-64
* ``model_ref``        -- functional PyTorch restatement of model/dualrefinedet_vggbn.py,
  model/dualrefinedet_mobilenet.py, model/refinedet_vgg.py, model/ssd4scale_vgg.py
* ``c/oracle.c``       -- plain-C restatement of the sampler, im2col+GEMM, decode and NMS
  (built into ``oracle/_build/liboracle.so`` by ``oracle/build.py``)
* ``ref_shim``         -- imports the *real* reference Python in place from /root/reference
  (only exists in the build container) to validate the restatements and to generate
  ``tests/golden/*.npz`` (see ``oracle/make_golden.py``).

Parity pin status: the reference ships no golden vectors or known-answer tests for this
path.  The Python model / Detect / PriorBox restatements are pinned
against outputs of the reference's own Python files executed in the build container
(fixtures under tests/golden, generator committed).  The two *native* pieces -- the CUDA
deformable conv and the Cython NMS -- cannot be compiled or run anywhere we have access to
(torch.utils.ffi/THC removed; Cython 0.25 source rejects Cython 3/NumPy 2), so for those two
functions the status is "parity unpinned": the restatement follows the source line by line
and is cross-checked against torchvision.ops.deform_conv2d (interior), F.conv2d (zero
offset) and the reference's importable utils/nms/py_cpu_nms.py.
"""
