// conv_first.cu -- the Cin = 3 stem convolution (VGG conv1_1, model/networks.py:146 with in_channels=3;
// MobileNet backbone[0], model/dualrefinedet_mobilenet.py:20) as a register-tiled direct convolution.
//
// Reads the reference's NCHW fp32 image directly (no separate layout/cast pass), writes NHWC in the
// activation dtype, folded-BN bias and ReLU fused.  K = 27 is too short for the tensor pipe to pay
// off, so this is a CUDA-core kernel: a CTA owns a 64-pixel strip of 8 output rows, keeps the
// [27][Cout] weights in shared memory, stages the 3 x (rows + halo) x (strip + halo) input patch once, and each
// thread accumulates 2 pixels (32 apart) x 16 output channels (864 FMAs per 135 shared-memory reads,
// weights read as broadcast LDS.128).
#include "common.cuh"

namespace tdrn {

constexpr int CF_TW = 64;      // output pixels per strip
constexpr int CF_ROWS = 8;     // output rows per CTA

template <typename TOut>
__global__ void __launch_bounds__(128) conv_first_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                         const float *__restrict__ bias, TOut *__restrict__ out,
                                                         int H, int W, int Cout, int Ho, int Wo, int stride, int relu)
{
    extern __shared__ __align__(16) float sm[];
    float *sw = sm;                               // [27][Cout]
    const int PW = CF_TW * stride + 2;            // patch width incl. halo
    const int RH = (CF_ROWS - 1) * stride + 3;    // patch rows: all 8 output rows of the CTA are staged at once (one barrier)
    float *sp = sm + 27 * Cout;                   // [3 ch][RH rows][PW]
    const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;       // warp = 16-channel group, lane = pixel pair
    const int b = blockIdx.z, x0 = blockIdx.x * CF_TW, y0 = blockIdx.y * CF_ROWS;
    for (int i = tid; i < 27 * Cout; i += blockDim.x) sw[i] = w[i];

    const float *xb = x + (long long)b * 3 * H * W;
    const int iy0 = y0 * stride - 1, ix0 = x0 * stride - 1;
    for (int cr = 0; cr < 3 * RH; ++cr) {         // cr = c * RH + row: whole rows, consecutive threads -> consecutive columns
        const int c = cr / RH, r = cr - c * RH;
        const int iy = iy0 + r;
        const bool row_ok = iy >= 0 && iy < H;
        const float *src = xb + ((long long)c * H + (row_ok ? iy : 0)) * W;
        for (int px = tid; px < PW; px += blockDim.x) {
            const int ix = ix0 + px;
            sp[cr * PW + px] = (row_ok && ix >= 0 && ix < W) ? __ldg(src + ix) : 0.f;
        }
    }
    float bv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) bv[j] = bias ? bias[g * 16 + j] : 0.f;
    __syncthreads();

    const int p0 = lane * stride, p1 = (lane + 32) * stride;         // pixels lane and lane + 32 of the strip (bank-friendly)
    for (int yy = 0; yy < CF_ROWS; ++yy) {
        const int y = y0 + yy;
        if (y >= Ho) break;
        float acc0[16], acc1[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { acc0[j] = bv[j]; acc1[j] = bv[j]; }
        const float *spr = sp + yy * stride * PW;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float a0 = spr[(c * RH + r) * PW + p0 + s], a1 = spr[(c * RH + r) * PW + p1 + s];
                    const float4 *wp = (const float4 *)(sw + ((r * 3 + s) * 3 + c) * Cout + g * 16);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 wv = wp[q];
                        acc0[4 * q + 0] = fmaf(a0, wv.x, acc0[4 * q + 0]); acc1[4 * q + 0] = fmaf(a1, wv.x, acc1[4 * q + 0]);
                        acc0[4 * q + 1] = fmaf(a0, wv.y, acc0[4 * q + 1]); acc1[4 * q + 1] = fmaf(a1, wv.y, acc1[4 * q + 1]);
                        acc0[4 * q + 2] = fmaf(a0, wv.z, acc0[4 * q + 2]); acc1[4 * q + 2] = fmaf(a1, wv.z, acc1[4 * q + 2]);
                        acc0[4 * q + 3] = fmaf(a0, wv.w, acc0[4 * q + 3]); acc1[4 * q + 3] = fmaf(a1, wv.w, acc1[4 * q + 3]);
                    }
                }
        if (relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) { acc0[j] = fmaxf(acc0[j], 0.f); acc1[j] = fmaxf(acc1[j], 0.f); }
        }
#pragma unroll
        for (int pp = 0; pp < 2; ++pp) {
            const int xo = x0 + lane + 32 * pp;
            if (xo >= Wo) continue;
            const float *a = pp ? acc1 : acc0;
            TOut *op = out + (((long long)b * Ho + y) * Wo + xo) * Cout + g * 16;
            if (sizeof(TOut) == 2) {
                uint4 q[2];
                __nv_bfloat162 *qb = (__nv_bfloat162 *)q;
#pragma unroll
                for (int j = 0; j < 8; ++j) qb[j] = __floats2bfloat162_rn(a[2 * j], a[2 * j + 1]);
                ((uint4 *)op)[0] = q[0]; ((uint4 *)op)[1] = q[1];
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) ((float4 *)op)[j] = make_float4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
            }
        }
    }
}

int launch_conv_first(const float *x, const float *w, const float *bias, void *out, int B, int H, int W, int Cout,
                      int Ho, int Wo, int stride, int relu, int out_dtype, cudaStream_t st)
{
    const int groups = Cout / 16;
    dim3 grid(ceil_div(Wo, CF_TW), ceil_div(Ho, CF_ROWS), B), block(32 * groups);
    const size_t smem = (size_t)(27 * Cout + 3 * ((CF_ROWS - 1) * stride + 3) * (CF_TW * stride + 2)) * sizeof(float);
    if (out_dtype == TDRN_BF16)
        conv_first_kernel<__nv_bfloat16><<<grid, block, smem, st>>>(x, w, bias, (__nv_bfloat16 *)out, H, W, Cout, Ho, Wo, stride, relu);
    else
        conv_first_kernel<float><<<grid, block, smem, st>>>(x, w, bias, (float *)out, H, W, Cout, Ho, Wo, stride, relu);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

}  // namespace tdrn
