// wrap.cu -- extern "C" doors into the reference's OWN native CUDA sources, compiled unmodified from where
// they lie under /root/reference by oracle/build_ref.py into oracle/_ref/libtdrn_ref_native.so.
//
// TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests/test_gpu_ref_native.py on the GPU box
// to pin the oracle's restatements (and the product kernels) against the real reference kernels:
//   * deformable_im2col<float>   utils/deformconv/deform_conv_cuda_kernel.cu:211-238 (kernel :157-208)
//   * _nms                       utils/nms/nms_kernel.cu:91-144 (declared utils/nms/gpu_nms.hpp:1-2)
// Nothing here restates reference code: this file only declares the two reference entry points (C++ linkage in
// the reference) and forwards to them.
#include <cuda_runtime.h>
#include "deform_conv_cuda_kernel.h"   // from /root/reference/utils/deformconv (-I)
#include "gpu_nms.hpp"                 // from /root/reference/utils/nms (-I)

extern "C" int ref_deformable_im2col(const float *data_im, const float *data_offset, int channels, int height,
                                     int width, int ksize_h, int ksize_w, int pad_h, int pad_w, int stride_h,
                                     int stride_w, int dilation_h, int dilation_w, int deformable_group,
                                     float *data_col, void *stream)
{
    deformable_im2col<float>((cudaStream_t)stream, data_im, data_offset, channels, height, width, ksize_h, ksize_w,
                             pad_h, pad_w, stride_h, stride_w, dilation_h, dilation_w, deformable_group, data_col);
    return (int)cudaGetLastError();
}

extern "C" int ref_gpu_nms(int *keep_out_host, int *num_out_host, const float *boxes_host, int boxes_num,
                           int boxes_dim, float nms_overlap_thresh, int device_id)
{
    _nms(keep_out_host, num_out_host, boxes_host, boxes_num, boxes_dim, nms_overlap_thresh, device_id);
    return (int)cudaGetLastError();
}
