// conv_stem_tc.cu -- the Cin = 3 stem convolution (VGG conv1_1 + folded BN + ReLU, model/networks.py:146 with
// in_channels = 3, first entry of vgg_base) on the tensor cores.
//
// The layer is 0.35 GFLOP/frame but writes the largest activation of the network (B x 320 x 320 x 64 bf16 =
// 419 MB at b32), so it is HBM-write bound; the CUDA-core version (conv_first.cu) was FMA-issue bound at ~0.9 TB/s.
// Here the 27-tap dot product runs as a K = 32 (27 + zero pad) tcgen05 GEMM:
//   A [128 pixels x 32] bf16 : im2col rows built by 4 producer warps from a fp32 NCHW patch that a 4-deep TMA
//                              ring stages in shared memory (3-D boxes of the reference's own input layout; the
//                              conv zero padding is the TMA out-of-bounds fill; no layout/cast pass),
//                              written in the 128B-swizzled K-major UMMA layout (logical chunks 0..3 of each row);
//   B [64 couts x 32]   bf16 : converted once per CTA from the packed fp32 weights [27][64];
//   D [128 x 64] fp32 in TMEM, double buffered; 4 epilogue warps add bias, ReLU, cast to bf16 into a swizzled
//   staging tile and ONE thread issues a TMA store (cp.async.bulk.tensor ... global <- shared) of the 16 KB box,
//   so the NHWC output is written as full 128-byte lines.
// Persistent grid (2 CTAs per SM), tile = bw x bh output pixels of one image (64x2, 32x4 or 16x8).
//
// SPLIT (fp32-accurate path, out_dtype TDRN_BF16_SPLIT): x = hi + lo and w = hi + lo to 16 mantissa bits each.  An A row is
// [x_hi (k 0..31) | x_lo (k 0..31)] = K 64, B1 rows are [w_hi | w_hi], B2 rows [w_lo | -]:  A x B1 (four K = 16 steps) gives
// x_hi*w_hi + x_lo*w_hi, the first half of A x B2 (two steps) adds x_hi*w_lo -- six accumulation steps into one accumulator
// (nothing like the hundreds of a long-K layer, so no accumulator spreading is needed here).  The epilogue writes the
// (hi | lo) operand of conv1_2, [B,H,W,128] bf16, as two TMA stores (channels 0..63 and 64..127).
//
// S = 2, COUT = 32: the MobileNet stem (dualrefinedet_mobilenet.py:20, conv_bn(3, 32, 2)): the patch box covers
// (bw*2 + 8) x (bh*2 + 1) input pixels, the im2col rows step two input pixels per output pixel, N = 32; the 64-byte output
// rows are staged un-swizzled (their tensor map is SWIZZLE_NONE).  The CUDA-core stem took 0.29 ms of the 3.48 ms b64 step.
#include "tc_common.cuh"
#include <stdlib.h>

namespace tdrn {
namespace tc {

struct StemP {
    const float *x;        // [B,3,H,W] fp32 (reference layout)
    const float *w;        // [27][64] fp32, k = (i*3+j)*3 + c
    const float *bias;     // [64] or NULL
    int B, H, W;
    int bw, bh, pw, tiles_w, tiles_h, total;   // pw = patch pitch (bw + halo, padded so that pw*4 B is TMA-legal)
    int relu;
    int f16;               // operands (image, weights) and output as IEEE half instead of bf16 (TDRN_F16; not with SPLIT)
};

constexpr int ST_THREADS = 320;          // warps 0-3 producers, 4 MMA, 5-8 epilogue, 9 patch TMA
constexpr int ST_PSTAGES = 4;            // input-patch ring depth (hides HBM latency of the fp32 image reads)
// patch stage: >= 3 ch x ((bh-1)*S + 3) rows x (bw*S + 8) cols x 4 B for the three tile shapes (64x2, 32x4, 16x8), 128B aligned
constexpr int st_patch_bytes(int S) { return S == 1 ? 4096 : 8192; }
constexpr int st_smem(int S) { return 2 * 16384 + 2 * 16384 + 2 * 8192 + ST_PSTAGES * st_patch_bytes(S) + 256 + 1024; }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// low parts of two fp32 values: bf16(v - bf16(v)), packed like pack_bf16x2
__device__ __forceinline__ uint32_t split_lo2(float a, float b)
{
    const float ha = __bfloat162float(__float2bfloat16_rn(a)), hb = __bfloat162float(__float2bfloat16_rn(b));
    return pack_bf16x2(__fsub_rn(a, ha), __fsub_rn(b, hb));
}

template <bool SPLIT, int S, int ST_COUT>
__global__ void __launch_bounds__(ST_THREADS, 2) conv_stem_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmO, const StemP p)
{
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t a_full[2], a_empty[2], t_full[2], t_empty[2];
    __shared__ __align__(8) uint64_t p_full[ST_PSTAGES], p_empty[ST_PSTAGES];
    __shared__ uint32_t tmem_base_s;

    uint8_t *base = (uint8_t *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = base;                                   // 2 x [128 rows][128 B]
    uint8_t *sO = base + 2 * 16384;                       // 2 x [128 rows][128 B] output staging (SPLIT: the hi and the lo box of one tile)
    uint8_t *sB = base + 4 * 16384;                       // [64 rows][128 B] (SPLIT: B1 = [w_hi | w_hi], then B2 = [w_lo | -])
    constexpr int ST_PATCH_BYTES = st_patch_bytes(S);
    static_assert(!SPLIT || (S == 1 && ST_COUT == 64), "split precision: the VGG stem only");
    uint8_t *sP = base + 4 * 16384 + 2 * 8192;            // ST_PSTAGES x patch [3][(bh-1)*S+3][bw*S+8] fp32
    float *sBias = (float *)(sP + ST_PSTAGES * ST_PATCH_BYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int f16 = SPLIT ? 0 : p.f16;
    const int PH = (p.bh - 1) * S + 3, PW = p.pw;
    const int tiles_per_img = p.tiles_w * p.tiles_h;

    if (tid == 0) {
        tma_prefetch_desc(&tmO);
        tma_prefetch_desc(&tmX);
#pragma unroll
        for (int s = 0; s < ST_PSTAGES; ++s) { mbar_init(&p_full[s], 1); mbar_init(&p_empty[s], 1); }
#pragma unroll
        for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 4); }
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(&tmem_base_s, 128);
    // B operand: w[k][n] fp32 -> bf16 rows n, logical chunks 0..3 (k 0..31, zero beyond 27)
    for (int e = tid; e < ST_COUT * 4; e += ST_THREADS) {
        const int n = e >> 2, chunk = e & 3;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int k = chunk * 8 + j; v[j] = k < 27 ? __ldg(p.w + k * ST_COUT + n) : 0.f; }
        uint4 q;
        q.x = pack16x2(v[0], v[1], f16); q.y = pack16x2(v[2], v[3], f16); q.z = pack16x2(v[4], v[5], f16); q.w = pack16x2(v[6], v[7], f16);
        *(uint4 *)(sB + sw128_offset(n, chunk)) = q;
        if (SPLIT) {
            *(uint4 *)(sB + sw128_offset(n, chunk + 4)) = q;
            uint4 l;
            l.x = split_lo2(v[0], v[1]); l.y = split_lo2(v[2], v[3]); l.z = split_lo2(v[4], v[5]); l.w = split_lo2(v[6], v[7]);
            *(uint4 *)(sB + 8192 + sw128_offset(n, chunk)) = l;
            *(uint4 *)(sB + 8192 + sw128_offset(n, chunk + 4)) = make_uint4(0, 0, 0, 0);
        }
    }
    if (tid < ST_COUT) sBias[tid] = p.bias ? p.bias[tid] : 0.f;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    pdl_sync();                               // programmatic dependent launch (common.cuh): the image may still be in flight

    if (warp < 4) {
        // ===================== producers: fp32 NCHW patch (TMA ring) -> bf16 im2col rows =====================
        const int wl = tid % p.bw, hl = tid / p.bw;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.total; tile += gridDim.x, ++it) {
            const int s = it & 1;
            const uint32_t ps = it % ST_PSTAGES, pph = (it / ST_PSTAGES) & 1u;
            mbar_wait(&p_full[ps], pph);
            mbar_wait(&a_empty[s], ((it >> 1) & 1u) ^ 1u);
            const float *P = (const float *)(sP + ps * ST_PATCH_BYTES);
            uint32_t kw[16], kl[16];
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) {
                float v[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int k = 2 * k2 + h;                           // compile-time after unrolling
                    if (k < 27) { const int t = k / 3, c = k - t * 3, i = t / 3, j = t - i * 3; v[h] = P[(c * PH + hl * S + i) * PW + wl * S + j + 3]; }
                    else v[h] = 0.f;
                }
                kw[k2] = pack16x2(v[0], v[1], f16);
                if (SPLIT) kl[k2] = split_lo2(v[0], v[1]);
            }
            uint8_t *a = sA + s * 16384;
#pragma unroll
            for (int chunk = 0; chunk < 4; ++chunk) {
                *(uint4 *)(a + sw128_offset(tid, chunk)) = make_uint4(kw[4 * chunk], kw[4 * chunk + 1], kw[4 * chunk + 2], kw[4 * chunk + 3]);
                if (SPLIT)
                    *(uint4 *)(a + sw128_offset(tid, chunk + 4)) = make_uint4(kl[4 * chunk], kl[4 * chunk + 1], kl[4 * chunk + 2], kl[4 * chunk + 3]);
            }
            fence_proxy_async_smem();
            named_bar(1, 128);                                          // A tile complete, patch[ps] fully consumed
            if (tid == 0) { mbar_arrive(&a_full[s]); mbar_arrive(&p_empty[ps]); }
        }
    } else if (warp == 4) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_16(128, ST_COUT, f16);
            const uint64_t bdesc = umma_desc_sw128(smem_u32(sB));
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.total; tile += gridDim.x, ++it) {
                const uint32_t s = it & 1u, ph = (it >> 1) & 1u;
                mbar_wait(&t_empty[s], ph ^ 1u);
                mbar_wait(&a_full[s], ph);
                tc_fence_after();
                const uint64_t adesc = umma_desc_sw128(smem_u32(sA + s * 16384));
                umma_bf16(tmem_base + s * ST_COUT, adesc, bdesc, idesc, 0u);
                umma_bf16(tmem_base + s * ST_COUT, adesc + 2, bdesc + 2, idesc, 1u);
                if (SPLIT) {
                    const uint64_t bdesc2 = umma_desc_sw128(smem_u32(sB + 8192));
                    umma_bf16(tmem_base + s * ST_COUT, adesc + 4, bdesc + 4, idesc, 1u);      // x_lo * w_hi
                    umma_bf16(tmem_base + s * ST_COUT, adesc + 6, bdesc + 6, idesc, 1u);
                    umma_bf16(tmem_base + s * ST_COUT, adesc, bdesc2, idesc, 1u);             // x_hi * w_lo
                    umma_bf16(tmem_base + s * ST_COUT, adesc + 2, bdesc2 + 2, idesc, 1u);
                }
                umma_commit(&a_empty[s]);
                umma_commit(&t_full[s]);
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // ===================== input-patch TMA producer =====================
        if (elect_one()) {
            const uint32_t patch_bytes = (uint32_t)(3 * PH * PW * 4);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.total; tile += gridDim.x, ++it) {
                const uint32_t ps = it % ST_PSTAGES, pph = (it / ST_PSTAGES) & 1u;
                const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
                mbar_wait(&p_empty[ps], pph ^ 1u);
                mbar_expect_tx(&p_full[ps], patch_bytes);
                tma_load_4d(sP + ps * ST_PATCH_BYTES, &tmX, &p_full[ps], (rem % p.tiles_w) * p.bw * S - 4, (rem / p.tiles_w) * p.bh * S - 1, 0, b);
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue: TMEM -> bias/ReLU/bf16 -> swizzled smem -> TMA store =====================
        const int et = tid - 160;                 // warps 5-8
        const int quad = warp & 3, r = quad * 32 + lane;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.total; tile += gridDim.x, ++it) {
            const uint32_t s = it & 1u, ph = (it >> 1) & 1u;
            if (et == 0) { if (SPLIT) bulk_wait_read<0>(); else bulk_wait_read<1>(); }   // the store that last read these boxes has drained
            named_bar(2, 128);
            mbar_wait(&t_full[s], ph);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + s * ST_COUT;
            uint8_t *o = SPLIT ? sO : sO + s * 16384;
#pragma unroll
            for (int c0 = 0; c0 < ST_COUT; c0 += 16) {
                float v[16];
                tmem_ld16(trow + (uint32_t)c0, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) { v[j] += sBias[c0 + j]; if (p.relu) v[j] = fmaxf(v[j], 0.f); }
                // COUT = 64: 128-byte rows, 128B swizzle; COUT = 32: 64-byte rows, no swizzle (tensor map SWIZZLE_NONE)
                const uint32_t o0 = ST_COUT == 64 ? sw128_offset(r, c0 >> 3) : (uint32_t)(r * 64 + (c0 >> 3) * 16);
                const uint32_t o1 = ST_COUT == 64 ? sw128_offset(r, (c0 >> 3) + 1) : o0 + 16u;
                *(uint4 *)(o + o0) =
                    make_uint4(pack16x2(v[0], v[1], f16), pack16x2(v[2], v[3], f16), pack16x2(v[4], v[5], f16), pack16x2(v[6], v[7], f16));
                *(uint4 *)(o + o1) =
                    make_uint4(pack16x2(v[8], v[9], f16), pack16x2(v[10], v[11], f16), pack16x2(v[12], v[13], f16), pack16x2(v[14], v[15], f16));
                if (SPLIT) {
                    *(uint4 *)(o + 16384 + sw128_offset(r, c0 >> 3)) =
                        make_uint4(split_lo2(v[0], v[1]), split_lo2(v[2], v[3]), split_lo2(v[4], v[5]), split_lo2(v[6], v[7]));
                    *(uint4 *)(o + 16384 + sw128_offset(r, (c0 >> 3) + 1)) =
                        make_uint4(split_lo2(v[8], v[9]), split_lo2(v[10], v[11]), split_lo2(v[12], v[13]), split_lo2(v[14], v[15]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[s]);
            fence_proxy_async_smem();
            named_bar(2, 128);
            if (et == 0) {
                const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
                tma_store_4d(&tmO, o, 0, (rem % p.tiles_w) * p.bw, (rem / p.tiles_w) * p.bh, b);
                if (SPLIT) tma_store_4d(&tmO, o + 16384, ST_COUT, (rem % p.tiles_w) * p.bw, (rem / p.tiles_w) * p.bh, b);
                bulk_commit();
            }
        }
        if (et == 0) bulk_wait_read<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}

template <bool SPLIT, int S, int COUT>
static int launch_stem(const CUtensorMap &tmX, const CUtensorMap &tmO, const StemP &p, int grid, cudaStream_t st)
{
    TDRN_CUDA(cudaFuncSetAttribute(conv_stem_tc_kernel<SPLIT, S, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, st_smem(S)));
    TDRN_CUDA(launch_pdl(conv_stem_tc_kernel<SPLIT, S, COUT>, dim3(grid), dim3(ST_THREADS), st_smem(S), st, tmX, tmO, p));
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

// stride 1 / Cout 64 (VGG conv1_1; split = the fp32-accurate form) or stride 2 / Cout 32 (MobileNet stem).
// -> TDRN_EUNSUPPORTED when the shape does not tile (caller falls back to the CUDA-core stem)
int launch_conv_stem_tc(const float *x, const float *w, const float *bias, void *out, int B, int H, int W, int relu,
                        bool split, int stride, int cout, cudaStream_t st, bool f16)
{
    if (!((stride == 1 && cout == 64) || (stride == 2 && cout == 32 && !split)) || (f16 && split)) return TDRN_EUNSUPPORTED;
    const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
    if (stride == 2 && ((H | W) & 1)) return TDRN_EUNSUPPORTED;
    StemP p{};
    if (Wo % 64 == 0 && Ho % 2 == 0) { p.bw = 64; p.bh = 2; }
    else if (Wo % 32 == 0 && Ho % 4 == 0) { p.bw = 32; p.bh = 4; }
    else if (Wo % 16 == 0 && Ho % 8 == 0) { p.bw = 16; p.bh = 8; }
    else return TDRN_EUNSUPPORTED;
    p.x = x; p.w = w; p.bias = bias; p.B = B; p.H = H; p.W = W; p.relu = relu; p.f16 = f16 ? 1 : 0;
    p.tiles_w = Wo / p.bw; p.tiles_h = Ho / p.bh; p.total = p.tiles_w * p.tiles_h * B;
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return TDRN_ECUDA; }
    CUtensorMap tmX;
    {   // fp32 NCHW image: dims (W, H, 3, B); box (bw*S + 8, (bh-1)*S + 3, 3, 1) starting at (x0*S - 4, y0*S - 1): out-of-bounds = conv padding
        if ((W * 4) % 16 != 0 || ((uintptr_t)x & 15)) return TDRN_EUNSUPPORTED;
        const cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)B};
        const cuuint64_t gstr[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
        p.pw = p.bw * stride + 8;   // columns x0*S-4 .. : TMA needs the innermost start coordinate 16-byte aligned (4 floats)
        const int ph = (p.bh - 1) * stride + 3;
        const cuuint32_t bdim[4] = {(cuuint32_t)p.pw, (cuuint32_t)ph, 3, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        if (3 * ph * p.pw * 4 > st_patch_bytes(stride)) return TDRN_EUNSUPPORTED;
        CUresult r = enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(x), gdim, gstr, bdim, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (stem input) failed (CUresult %d)", (int)r); return TDRN_ECUDA; }
    }
    CUtensorMap tmO;
    const uint64_t oc = split ? 2 * cout : cout;                     // channels per output pixel in memory
    if (cout == 64) {
        const uint64_t dims[4] = {oc, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)B};
        const uint64_t str[3] = {oc * 2, (uint64_t)Wo * oc * 2, (uint64_t)Ho * Wo * oc * 2};
        const uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, 1};
        int rc = make_tmap_bf16(&tmO, out, 4, dims, str, box, nullptr);
        if (rc) return rc;
    } else {   // 32 channels = 64-byte rows: staged un-swizzled
        if ((uintptr_t)out & 15) return TDRN_EUNSUPPORTED;
        const cuuint64_t gdim[4] = {(cuuint64_t)oc, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)B};
        const cuuint64_t gstr[3] = {(cuuint64_t)oc * 2, (cuuint64_t)Wo * oc * 2, (cuuint64_t)Ho * Wo * oc * 2};
        const cuuint32_t bdim[4] = {(cuuint32_t)cout, (cuuint32_t)p.bw, (cuuint32_t)p.bh, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out, gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (stem output) failed (CUresult %d)", (int)r); return TDRN_ECUDA; }
    }
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        TDRN_CUDA(cudaGetDevice(&dev));
        TDRN_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int grid = p.total < 2 * num_sms ? p.total : 2 * num_sms;
    if (stride == 2) return launch_stem<false, 2, 32>(tmX, tmO, p, grid, st);
    if (split) return launch_stem<true, 1, 64>(tmX, tmO, p, grid, st);
    return launch_stem<false, 1, 64>(tmX, tmO, p, grid, st);
}

}  // namespace tc
}  // namespace tdrn
