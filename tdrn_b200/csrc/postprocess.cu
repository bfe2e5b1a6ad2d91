// postprocess.cu -- two-stage box decode, per-class candidate selection, sort, greedy NMS, top-k.
//
// Replaces the Python loops of Detect.forward (layers/functions/detection.py:25-70: B x (C-1)
// device->host round trips + single-core Cython NMS, utils/nms/cpu_nms.pyx:17-68) with two kernels:
//
//   decode_transpose_kernel : boxes[B,P,4] = decode(loc, center_size(decode(arm_loc, priors)))
//                             (layers/box_utils.py:176-195, :16-25) and scoresT[B,C,P] (class-major
//                             copy of conf so every (image, class) segment is one coalesced row).
//   nms_segment_kernel      : one CTA per (image, class): threshold + compaction, bitonic sort on
//                             (score, ~index) 64-bit keys in shared memory, chunked greedy NMS with
//                             a 256x256 suppression bitmask per chunk resolved by one warp with
//                             shuffles, early exit once top_k boxes are kept (exactly equivalent,
//                             SURVEY.md 8a "Exactness note for A8").
//
// Bit-exactness: every fp32 operation of the reference's IoU (cpu_nms.pyx:24,57-65) and of decode
// is issued with explicit round-to-nearest intrinsics in the reference's order so that nvcc cannot
// contract them into FMAs; `ovr >= thresh` is the reference's float-vs-double compare.
#include "common.cuh"
#include <math.h>

namespace tdrn {

constexpr int NMS_THREADS = 256;           // == chunk size of the greedy scan
constexpr int NMS_WORDS = NMS_THREADS / 32;

__device__ __forceinline__ float4 decode_box(float4 l, float4 p)
{
    // box_utils.py:190-195: cxcy = p_xy + (l_xy*0.1)*p_wh ; wh = p_wh*exp(l_wh*0.2);
    // x1y1 = cxcy - wh/2 ; x2y2 = wh + x1y1
    const float cx = __fadd_rn(p.x, __fmul_rn(__fmul_rn(l.x, 0.1f), p.z));
    const float cy = __fadd_rn(p.y, __fmul_rn(__fmul_rn(l.y, 0.1f), p.w));
    const float w = __fmul_rn(p.z, expf(__fmul_rn(l.z, 0.2f)));
    const float h = __fmul_rn(p.w, expf(__fmul_rn(l.w, 0.2f)));
    const float x1 = __fsub_rn(cx, __fmul_rn(w, 0.5f));
    const float y1 = __fsub_rn(cy, __fmul_rn(h, 0.5f));
    return make_float4(x1, y1, __fadd_rn(w, x1), __fadd_rn(h, y1));
}

__device__ __forceinline__ float4 center_size_box(float4 b)
{
    // box_utils.py:24-25
    return make_float4(__fmul_rn(__fadd_rn(b.z, b.x), 0.5f), __fmul_rn(__fadd_rn(b.w, b.y), 0.5f),
                       __fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

constexpr int DEC_TP = 128;   // priors per CTA

// grid (ceil(P/DEC_TP), B).  scoresT may be NULL (plain tdrn_decode).
__global__ void __launch_bounds__(256) decode_transpose_kernel(const float4 *__restrict__ loc, const float4 *__restrict__ priors,
                                                               const float4 *__restrict__ arm_loc, const float *__restrict__ conf,
                                                               float4 *__restrict__ boxes, float *__restrict__ scoresT, int P, int C)
{
    extern __shared__ float s_conf[];          // [DEC_TP * C]
    const int b = blockIdx.y, p0 = blockIdx.x * DEC_TP;
    const int np = min(DEC_TP, P - p0);
    if (threadIdx.x < np) {
        const int p = p0 + threadIdx.x;
        float4 prior = priors[p];
        if (arm_loc) prior = center_size_box(decode_box(arm_loc[(long long)b * P + p], prior));
        boxes[(long long)b * P + p] = decode_box(loc[(long long)b * P + p], prior);
    }
    if (scoresT) {
        const float *src = conf + ((long long)b * P + p0) * C;
        for (int i = threadIdx.x; i < np * C; i += blockDim.x) s_conf[i] = src[i];
        __syncthreads();
        for (int i = threadIdx.x; i < np * (C - 1); i += blockDim.x) {
            const int cl = 1 + i / np, pl = i - (cl - 1) * np;
            scoresT[((long long)b * C + cl) * P + p0 + pl] = s_conf[pl * C + cl];
        }
    }
}

struct NmsP {
    // detect mode
    const float4 *boxes;      // [B,P,4] normalised
    const float *scoresT;     // [B,C,P]
    float *out;               // [B,C,top_k,5]
    int P, C, top_k;
    float conf_thresh;
    float4 scale;
    // standalone mode
    const float *dets;        // [n,5]
    int n;
    int *keep; int *num_keep;
    float *kept_ws;           // [max_keep*5] global scratch for the kept list
    // common
    int max_keep;
    float thr_up;             // smallest float >= (double) nms threshold
    int n_pad_max;            // keys[] capacity (power of two)
};

__device__ __forceinline__ bool iou_ge(float4 a, float aarea, float4 b, float barea, float thr)
{
    // cpu_nms.pyx:57-65
    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    const float w = fmaxf(0.0f, __fadd_rn(__fsub_rn(xx2, xx1), 1.0f));
    const float h = fmaxf(0.0f, __fadd_rn(__fsub_rn(yy2, yy1), 1.0f));
    const float inter = __fmul_rn(w, h);
    const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, barea), inter));
    return ovr >= thr;
}

__device__ __forceinline__ float box_area(float4 b)
{
    // cpu_nms.pyx:24
    return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.0f), __fadd_rn(__fsub_rn(b.w, b.y), 1.0f));
}

// One CTA per segment.  DETECT: grid (C, B); class 0 only zero-fills.  Standalone: grid (1).
template <bool DETECT>
__global__ void __launch_bounds__(NMS_THREADS) nms_segment_kernel(const NmsP p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *keys = (unsigned long long *)smem_raw;                       // [n_pad_max]
    float *kept_s = (float *)(keys + p.n_pad_max);                                    // DETECT: [max_keep*5]
    __shared__ float4 cbox[NMS_THREADS];
    __shared__ float carea[NMS_THREADS];
    __shared__ int cidx[NMS_THREADS];
    __shared__ unsigned char calive[NMS_THREADS];
    __shared__ unsigned cmask[NMS_THREADS][NMS_WORDS];
    __shared__ int s_count, s_kept;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int cl = 0, b = 0;
    float *out_seg = nullptr;
    const float *sc = nullptr;
    int n_in;
    if (DETECT) {
        cl = blockIdx.x; b = blockIdx.y;
        out_seg = p.out + ((long long)b * p.C + cl) * p.top_k * 5;
        if (cl == 0) {                                             // background row stays zero (detection.py:37,52)
            for (int i = tid; i < p.top_k * 5; i += NMS_THREADS) out_seg[i] = 0.f;
            return;
        }
        sc = p.scoresT + ((long long)b * p.C + cl) * p.P;
        n_in = p.P;
    } else {
        n_in = p.n;
    }
    float *kept = DETECT ? kept_s : p.kept_ws;

    // ---- 1. threshold + compaction into 64-bit keys: (score bits << 32) | ~index -------------------
    if (tid == 0) { s_count = 0; s_kept = 0; }
    __syncthreads();
    for (int i0 = 0; i0 < n_in; i0 += NMS_THREADS) {
        const int i = i0 + tid;
        bool take = false; float s = 0.f;
        if (i < n_in) {
            s = DETECT ? sc[i] : p.dets[5 * (long long)i + 4];
            take = DETECT ? (s > p.conf_thresh) : true;          // strict fp32 compare, detection.py:53
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, take);
        int base = 0;
        if (lane == 0 && ballot) base = atomicAdd(&s_count, __popc(ballot));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (take) {
            // order-preserving map of fp32 to u32 (handles negative scores in standalone mode)
            unsigned u = __float_as_uint(s);
            u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
            keys[base + __popc(ballot & ((1u << lane) - 1))] = ((unsigned long long)u << 32) | (unsigned)(0xffffffffu - (unsigned)i);
        }
    }
    __syncthreads();
    const int count = s_count;
    int n_pad = 32;
    while (n_pad < count) n_pad <<= 1;
    for (int i = count + tid; i < n_pad; i += NMS_THREADS) keys[i] = 0ull;
    __syncthreads();

    // ---- 2. bitonic sort, descending ----------------------------------------------------------------
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (n_pad >> 1); t += NMS_THREADS) {
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int hi = lo | j;
                const unsigned long long a = keys[lo], c = keys[hi];
                const bool desc = (lo & k) == 0;
                if (desc ? (a < c) : (a > c)) { keys[lo] = c; keys[hi] = a; }
            }
            __syncthreads();
        }
    }

    // ---- 3. chunked greedy NMS ----------------------------------------------------------------------
    int kept_n = 0;
    const int max_keep = p.max_keep > 0 ? p.max_keep : count;
    for (int c0 = 0; c0 < count && kept_n < max_keep; c0 += NMS_THREADS) {
        const int cnt = min(NMS_THREADS, count - c0);
        float4 bx = make_float4(0, 0, 0, 0); float area = 0.f; bool alive = false; int src = 0;
        if (tid < cnt) {
            src = (int)(0xffffffffu - (unsigned)(keys[c0 + tid] & 0xffffffffull));
            if (DETECT) {
                const float4 nb = p.boxes[(long long)b * p.P + src];
                bx = make_float4(__fmul_rn(nb.x, p.scale.x), __fmul_rn(nb.y, p.scale.y),
                                 __fmul_rn(nb.z, p.scale.z), __fmul_rn(nb.w, p.scale.w));   // detection.py:59
            } else {
                const float *d = p.dets + 5 * (long long)src;
                bx = make_float4(d[0], d[1], d[2], d[3]);
            }
            area = box_area(bx);
            alive = true;
            for (int q = 0; q < kept_n; ++q) {                      // phase 1: vs boxes kept in earlier chunks
                const float4 kb = make_float4(kept[q * 5 + 0], kept[q * 5 + 1], kept[q * 5 + 2], kept[q * 5 + 3]);
                if (iou_ge(kb, kept[q * 5 + 4], bx, area, p.thr_up)) { alive = false; break; }
            }
        }
        cbox[tid] = bx; carea[tid] = area; cidx[tid] = src; calive[tid] = alive ? 1 : 0;
        __syncthreads();
        if (alive) {                                                // phase 2a: my suppression row (j > tid)
#pragma unroll
            for (int w = 0; w < NMS_WORDS; ++w) {
                unsigned m = 0;
                if (w >= (tid >> 5)) {
                    const int jb = w * 32;
                    for (int jj = 0; jj < 32; ++jj) {
                        const int j = jb + jj;
                        if (j > tid && j < cnt && iou_ge(bx, area, cbox[j], carea[j], p.thr_up)) m |= 1u << jj;
                    }
                }
                cmask[tid][w] = m;
            }
        }
        __syncthreads();
        if (warp == 0) {                                            // phase 2b: serial resolve, one warp
            unsigned removed = 0;                                   // lane w (< NMS_WORDS) owns word w
            for (int i = 0; i < cnt; ++i) {
                const unsigned r = __shfl_sync(0xffffffffu, removed, i >> 5);
                if (calive[i] && !((r >> (i & 31)) & 1u)) {
                    if (lane == 0) {
                        const float4 kb = cbox[i];
                        kept[kept_n * 5 + 0] = kb.x; kept[kept_n * 5 + 1] = kb.y; kept[kept_n * 5 + 2] = kb.z;
                        kept[kept_n * 5 + 3] = kb.w; kept[kept_n * 5 + 4] = carea[i];
                        if (DETECT) {
                            const float4 nb = p.boxes[(long long)b * p.P + cidx[i]];
                            float *o = out_seg + kept_n * 5;         // detection.py:61-63
                            o[0] = sc[cidx[i]]; o[1] = nb.x; o[2] = nb.y; o[3] = nb.z; o[4] = nb.w;
                        } else {
                            p.keep[kept_n] = cidx[i];
                        }
                    }
                    ++kept_n;
                    if (lane < NMS_WORDS) removed |= cmask[i][lane];
                    if (kept_n >= max_keep) break;
                }
            }
            if (lane == 0) s_kept = kept_n;
        }
        __threadfence_block();
        __syncthreads();
        kept_n = s_kept;
    }

    if (DETECT) {
        for (int i = kept_n * 5 + tid; i < p.top_k * 5; i += NMS_THREADS) out_seg[i] = 0.f;
    } else if (tid == 0) {
        *p.num_keep = kept_n;
    }
}

static float thresh_up(double t)
{
    float f = (float)t;
    if ((double)f < t) f = nextafterf(f, INFINITY);
    return f;
}

static int pow2_at_least(int n) { int v = 32; while (v < n) v <<= 1; return v; }

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace tdrn

using namespace tdrn;

extern "C" int tdrn_decode(const float *loc, const float *priors, const float *arm_loc, int B, int P, float *boxes, tdrn_stream_t stream)
{
    TDRN_REQUIRE(loc && priors && boxes && B > 0 && P > 0, "tdrn_decode: bad argument");
    dim3 grid(ceil_div(P, DEC_TP), B);
    decode_transpose_kernel<<<grid, 256, 0, as_stream(stream)>>>((const float4 *)loc, (const float4 *)priors, (const float4 *)arm_loc,
                                                                 nullptr, (float4 *)boxes, nullptr, P, 0);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

extern "C" size_t tdrn_detect_workspace_bytes(int B, int P, int C, int top_k)
{
    (void)top_k;
    return align_up((size_t)B * P * 4 * sizeof(float), 256) + align_up((size_t)B * C * P * sizeof(float), 256);
}

extern "C" int tdrn_detect(const float *loc, const float *conf, const float *priors, const float *arm_loc,
                           const float *scale_host, int B, int P, int C, int top_k, float conf_thresh,
                           double nms_thresh, float *out, void *workspace, size_t workspace_bytes, tdrn_stream_t stream)
{
    TDRN_REQUIRE(loc && conf && priors && scale_host && out, "tdrn_detect: null argument");
    TDRN_REQUIRE(B > 0 && P > 0 && C > 1 && top_k > 0, "tdrn_detect: bad shape B=%d P=%d C=%d top_k=%d", B, P, C, top_k);
    TDRN_REQUIRE(nms_thresh > 0, "nms_threshold must be non negative.");          // detection.py:20-21
    if (!workspace || workspace_bytes < tdrn_detect_workspace_bytes(B, P, C, top_k)) {
        set_error("tdrn_detect: workspace too small (%zu < %zu)", workspace_bytes, tdrn_detect_workspace_bytes(B, P, C, top_k));
        return TDRN_EWORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    float *boxes = (float *)workspace;
    float *scoresT = (float *)((char *)workspace + align_up((size_t)B * P * 4 * sizeof(float), 256));

    const size_t dec_smem = (size_t)DEC_TP * C * sizeof(float);
    TDRN_REQUIRE(dec_smem <= 200 * 1024, "tdrn_detect: too many classes (%d)", C);
    if (dec_smem > 48 * 1024)
        TDRN_CUDA(cudaFuncSetAttribute(decode_transpose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec_smem));
    dim3 dgrid(ceil_div(P, DEC_TP), B);
    decode_transpose_kernel<<<dgrid, 256, dec_smem, st>>>((const float4 *)loc, (const float4 *)priors, (const float4 *)arm_loc,
                                                           conf, (float4 *)boxes, scoresT, P, C);
    TDRN_LAUNCH_CHECK();

    NmsP p{};
    p.boxes = (const float4 *)boxes; p.scoresT = scoresT; p.out = out; p.P = P; p.C = C; p.top_k = top_k;
    p.conf_thresh = conf_thresh;
    p.scale = make_float4(scale_host[0], scale_host[1], scale_host[2], scale_host[3]);
    p.max_keep = top_k; p.thr_up = thresh_up(nms_thresh); p.n_pad_max = pow2_at_least(P);
    const size_t smem = (size_t)p.n_pad_max * 8 + (size_t)top_k * 5 * sizeof(float);
    TDRN_REQUIRE(smem <= 200 * 1024, "tdrn_detect: P=%d / top_k=%d exceed the shared-memory sort capacity", P, top_k);
    TDRN_CUDA(cudaFuncSetAttribute(nms_segment_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nms_segment_kernel<true><<<dim3(C, B), NMS_THREADS, smem, st>>>(p);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

extern "C" size_t tdrn_nms_workspace_bytes(int n) { return align_up((size_t)(n > 0 ? n : 1) * 5 * sizeof(float), 256); }

extern "C" int tdrn_nms(const float *dets, int n, double thresh, int max_keep, int *keep, int *num_keep,
                        void *workspace, size_t workspace_bytes, tdrn_stream_t stream)
{
    TDRN_REQUIRE(keep && num_keep && n >= 0, "tdrn_nms: bad argument");
    cudaStream_t st = as_stream(stream);
    if (n == 0) {                                                      // nms_wrapper.py:26-27
        TDRN_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int), st));
        return TDRN_OK;
    }
    TDRN_REQUIRE(dets != nullptr, "tdrn_nms: null dets");
    if (!workspace || workspace_bytes < tdrn_nms_workspace_bytes(n)) {
        set_error("tdrn_nms: workspace too small (%zu < %zu)", workspace_bytes, tdrn_nms_workspace_bytes(n));
        return TDRN_EWORKSPACE;
    }
    NmsP p{};
    p.dets = dets; p.n = n; p.keep = keep; p.num_keep = num_keep; p.kept_ws = (float *)workspace;
    p.max_keep = max_keep; p.thr_up = thresh_up(thresh); p.n_pad_max = pow2_at_least(n);
    const size_t smem = (size_t)p.n_pad_max * 8;
    if (smem > 200 * 1024) {
        set_error("tdrn_nms: n=%d exceeds the single-CTA shared-memory sort capacity (25600 boxes)", n);
        return TDRN_EUNSUPPORTED;
    }
    TDRN_CUDA(cudaFuncSetAttribute(nms_segment_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nms_segment_kernel<false><<<1, NMS_THREADS, smem, st>>>(p);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

// Host-pointer, synchronous variant with the argument order of the reference's `_nms`
// (utils/nms/gpu_nms.hpp:1-2), except that the input need not be pre-sorted and the threshold rule
// is the CPU one Detect uses.
extern "C" int tdrn_nms_host(int *keep_out, int *num_out, const float *dets_host, int boxes_num, int boxes_dim,
                             double thresh, int device_id)
{
    TDRN_REQUIRE(keep_out && num_out && boxes_num >= 0, "tdrn_nms_host: bad argument");
    TDRN_REQUIRE(boxes_dim == 5, "tdrn_nms_host: boxes_dim must be 5");
    if (boxes_num == 0) { *num_out = 0; return TDRN_OK; }
    TDRN_REQUIRE(dets_host != nullptr, "tdrn_nms_host: null dets");
    if (device_id >= 0) TDRN_CUDA(cudaSetDevice(device_id));
    float *d_dets = nullptr; int *d_keep = nullptr; void *d_ws = nullptr;
    const size_t ws = tdrn_nms_workspace_bytes(boxes_num);
    int rc = TDRN_OK;
    cudaError_t e;
    if ((e = cudaMalloc(&d_dets, sizeof(float) * 5 * boxes_num)) != cudaSuccess ||
        (e = cudaMalloc(&d_keep, sizeof(int) * (boxes_num + 1))) != cudaSuccess ||
        (e = cudaMalloc(&d_ws, ws)) != cudaSuccess) {
        set_error("tdrn_nms_host: cudaMalloc failed: %s", cudaGetErrorString(e));
        rc = TDRN_ECUDA;
    }
    if (rc == TDRN_OK && (e = cudaMemcpy(d_dets, dets_host, sizeof(float) * 5 * boxes_num, cudaMemcpyHostToDevice)) != cudaSuccess) {
        set_error("tdrn_nms_host: H2D failed: %s", cudaGetErrorString(e)); rc = TDRN_ECUDA;
    }
    if (rc == TDRN_OK) rc = tdrn_nms(d_dets, boxes_num, thresh, 0, d_keep + 1, d_keep, d_ws, ws, nullptr);
    if (rc == TDRN_OK) {
        if ((e = cudaMemcpy(num_out, d_keep, sizeof(int), cudaMemcpyDeviceToHost)) != cudaSuccess ||
            (e = cudaMemcpy(keep_out, d_keep + 1, sizeof(int) * (*num_out), cudaMemcpyDeviceToHost)) != cudaSuccess) {
            set_error("tdrn_nms_host: D2H failed: %s", cudaGetErrorString(e)); rc = TDRN_ECUDA;
        }
    }
    cudaFree(d_dets); cudaFree(d_keep); cudaFree(d_ws);
    return rc;
}
