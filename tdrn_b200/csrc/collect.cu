// collect.cu -- device-side result scatter of the evaluation drivers: what evaluate.py:469-483 and
// evaluate_coco.py:140-159 do per image and class with masked_select + four in-place multiplies + .cpu():
//     dets = detections[i, j];  keep rows with score > 0;  boxes * (w, h, w, h);  (boxes, score) -> host
// i.e. B x (C-1) tiny device ops and D2H copies per batch.  Here: two launches and ONE copy for the whole batch.
//   collect_count_kernel   : rows with score > 0 per (image, class) segment (class 0 = background is skipped)
//   collect_compact_kernel : each segment sums the counts before it (B*C is a few thousand), then writes its rows in
//                            rank order:  (image, class, x1*w, y1*h, x2*w, y2*h, score)
#include "common.cuh"

namespace tdrn {

__global__ void __launch_bounds__(128) collect_count_kernel(const float *__restrict__ det, int C, int top_k, int *__restrict__ counts)
{
    const int seg = blockIdx.x, cl = seg % C;
    __shared__ int s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    int n = 0;
    if (cl != 0)
        for (int r = threadIdx.x; r < top_k; r += blockDim.x) n += det[((long long)seg * top_k + r) * 5] > 0.f;
    n = __reduce_add_sync(0xffffffffu, n);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(&s_n, n);
    __syncthreads();
    if (threadIdx.x == 0) counts[seg] = s_n;
}

__global__ void __launch_bounds__(128) collect_compact_kernel(const float *__restrict__ det, const float *__restrict__ wh, int B, int C,
                                                              int top_k, const int *__restrict__ counts, float *__restrict__ out,
                                                              int max_rows, int *__restrict__ total)
{
    const int seg = blockIdx.x, img = seg / C, cl = seg - img * C;
    __shared__ int s_base, s_run;
    int part = 0;
    for (int s = threadIdx.x; s < seg; s += blockDim.x) part += counts[s];
    part = __reduce_add_sync(0xffffffffu, part);
    if (threadIdx.x == 0) { s_base = 0; s_run = 0; }
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && part) atomicAdd(&s_base, part);
    __syncthreads();
    if (seg == B * C - 1 && threadIdx.x == 0) *total = s_base + counts[seg];
    if (cl == 0 || counts[seg] == 0) return;
    const float w = wh[2 * img], h = wh[2 * img + 1];
    const float *src = det + (long long)seg * top_k * 5;
    // rank order = row order; rows are visited 128 at a time, a warp-level ballot keeps them stable
    for (int r0 = 0; r0 < top_k; r0 += blockDim.x) {
        const int r = r0 + threadIdx.x;
        const bool take = r < top_k && src[r * 5] > 0.f;
        const unsigned ballot = __ballot_sync(0xffffffffu, take);
        __shared__ int s_warp[4];
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) s_warp[warp] = __popc(ballot);
        __syncthreads();
        int before = s_run;
        for (int k = 0; k < warp; ++k) before += s_warp[k];
        if (take) {
            const long long row = (long long)s_base + before + __popc(ballot & ((1u << lane) - 1u));
            if (row < max_rows) {
                float *o = out + row * 7;
                o[0] = (float)img; o[1] = (float)cl;
                o[2] = __fmul_rn(src[r * 5 + 1], w); o[3] = __fmul_rn(src[r * 5 + 2], h);     // evaluate.py:474-477
                o[4] = __fmul_rn(src[r * 5 + 3], w); o[5] = __fmul_rn(src[r * 5 + 4], h);
                o[6] = src[r * 5];
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_run += s_warp[0] + s_warp[1] + s_warp[2] + s_warp[3];
        __syncthreads();
    }
}

}  // namespace tdrn

using namespace tdrn;

extern "C" size_t tdrn_collect_workspace_bytes(int B, int C) { return (size_t)(B > 0 && C > 0 ? B * C : 0) * sizeof(int); }

extern "C" int tdrn_collect_detections(const float *det, const float *wh, int B, int C, int top_k, float *out_rows, int max_rows,
                                       int *count, void *workspace, size_t workspace_bytes, tdrn_stream_t stream)
{
    TDRN_REQUIRE(det && wh && out_rows && count && workspace && B > 0 && C > 0 && top_k > 0 && max_rows >= 0,
                 "tdrn_collect_detections: bad argument");
    if (workspace_bytes < tdrn_collect_workspace_bytes(B, C)) {
        set_error("tdrn_collect_detections: workspace too small (%zu < %zu)", workspace_bytes, tdrn_collect_workspace_bytes(B, C));
        return TDRN_EWORKSPACE;
    }
    int *counts = (int *)workspace;
    cudaStream_t st = as_stream(stream);
    collect_count_kernel<<<B * C, 128, 0, st>>>(det, C, top_k, counts);
    TDRN_LAUNCH_CHECK();
    collect_compact_kernel<<<B * C, 128, 0, st>>>(det, wh, B, C, top_k, counts, out_rows, max_rows, count);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}
