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
  model/dualrefinedet_mobilenet.py, model/refinedet_vgg.py, model/ssd4scale_vgg.py, model/ssd4scale_mobile.py
* ``preprocess_ref``   -- restatement of data/__init__.py:7-12 (base_transform) incl. OpenCV's 8-bit INTER_LINEAR
  fixed-point resize (pinned against the OpenCV 4.13.0 of this image; see the module header)
* ``eval_ref``         -- restatement of the result scatter of evaluate.py:469-483 (pinned against evaluate_trn.py's
  test_net run from its source)
* ``multi_scale_ref``  -- NumPy restatement of the multi-scale merge + bbox_vote (multi_eval.py:453-494,557-640);
  bbox_vote pinned by ``make_golden_vote`` (runs the reference's own function source), the whole merge by running
  multi_eval.py's test_net from its source (tests/test_oracle_vs_reference.py)
* ``c/oracle.c``       -- plain-C restatement of the sampler, im2col+GEMM, decode and NMS
  (built into ``oracle/_build/liboracle.so`` by ``oracle/build.py``)
* ``ref_shim``         -- imports the *real* reference Python in place from /root/reference
  (only exists in the build container) to validate the restatements and to generate
  ``tests/golden/*.npz`` (see ``oracle/make_golden.py``).

Parity pin status: the reference ships no golden vectors or known-answer tests for this
path.  Pins used instead:
  * Python model / Detect / PriorBox restatements: outputs of the reference's own Python files
    executed in the build container (fixtures under tests/golden, generator committed).
  * Deformable-conv sampler (the parity-critical native piece): the reference's OWN CUDA kernel
    file utils/deformconv/deform_conv_cuda_kernel.cu compiles unmodified with nvcc for sm_100a
    (oracle/build_ref.py -> oracle/_ref/libtdrn_ref_native.so) and is run on the GPU box next to
    the restatements and the product (tests/test_gpu_ref_native.py).  Its THC host wrapper
    (deform_conv_cuda.c) cannot be built (torch 0.4 THC); the wrapper's loop is 20 lines of
    zero + im2col + SGEMM and is restated in oracle/ref_native.deform_conv_forward.
  * NMS: the reference's GPU kernel utils/nms/nms_kernel.cu (`_nms`) is compiled and run the same
    way (IoU +1 convention, bitmask reduction); the Cython CPU variant that Detect actually calls
    (cpu_nms.pyx, suppress at ovr >= thresh instead of >) is built by oracle/build_ref_nms.py from a
    temporary copy with two removed NumPy dtype names respelled (np.int_t / np.int -> np.intp_t /
    np.intp) and nothing else changed; the restatements are compared with it in
    tests/test_ref_cython_nms.py and the reference's Detect runs with it in oracle/ref_shim.py.
  * Pre-processing: the reference's own base_transform source executed with the real cv2 of this image
    (oracle/make_golden_preprocess.py -> tests/golden/base_transform.npz) + live cv2.resize comparisons.
  * Drivers around the path (multi-scale merge, TDRN video loop, result scatter): the reference's own
    `test_net` functions of multi_eval.py / evaluate_trn.py executed from their source with stand-in
    networks (tests/test_oracle_vs_reference.py, build container only).
"""
