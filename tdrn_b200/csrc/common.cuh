// common.cuh -- error plumbing and small device helpers shared by every translation unit of
// libtdrn_b200.so.  No torch headers anywhere in csrc/.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/tdrn_b200.h"

namespace tdrn {

void set_error(const char *fmt, ...);          // api.cu
void count_launch(int n = 1);                  // api.cu

#define TDRN_REQUIRE(cond, ...)                                                     \
    do { if (!(cond)) { ::tdrn::set_error(__VA_ARGS__); return TDRN_EINVAL; } } while (0)

#define TDRN_CUDA(expr)                                                             \
    do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) {                        \
        ::tdrn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
        return TDRN_ECUDA; } } while (0)

// After a <<<>>> launch: surface launch-configuration errors (the reference only printf'd them,
// utils/deformconv/deform_conv_cuda_kernel.cu:233-237).
#define TDRN_LAUNCH_CHECK()                                                         \
    do { ::tdrn::count_launch(); TDRN_CUDA(cudaGetLastError()); } while (0)

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---- programmatic dependent launch (TDRN_PDL=1) ---------------------------------------------------------------------
// A kernel launched through launch_pdl() may be scheduled while its predecessor in the stream still runs; pdl_sync() at the end
// of its prologue (barrier init, TMEM allocation, descriptor prefetch -- nothing that touches data another kernel produces) lets
// ITS successor do the same and then waits until the predecessor's results are visible.  What overlaps is the launch latency
// and the prologue of every kernel: the batch-1 video paths are chains of ~50 kernels that each run for a few microseconds.
// Only kernels that call pdl_sync() may be launched with launch_pdl(); every other combination behaves like a plain launch.
__device__ __forceinline__ void pdl_sync()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

bool pdl_enabled();                            // api.cu

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline cudaStream_t as_stream(tdrn_stream_t s) { return (cudaStream_t)s; }

}  // namespace tdrn
