// simt_ops.cu -- fp32-accurate CUDA-core kernels of the hot path (sm_100a).
//
//   * conv_simt_kernel : strided implicit-GEMM convolution / ConvTranspose2d(k2,s2) / deformable
//     convolution with fp32 accumulation.  It is the fp32 ("1e-4") path for every layer and the
//     fallback-free implementation of the odd shapes the tcgen05 kernel does not take (Cin=3,
//     Cin%64!=0).  Generic element strides let the same kernel consume the reference's NCHW
//     tensors (tdrn_deform_conv_forward) and the library's NHWC activations.
//   * depthwise 3x3, maxpool 2x2 (ceil), L2Norm, row softmax, NCHW<->NHWC.
//
// Reference semantics restated here (paths relative to the upstream checkout):
//   deformable sampler     utils/deformconv/deform_conv_cuda_kernel.cu:16-51, :157-208
//   L2Norm                 layers/modules/l2norm.py:17-21
//   vgg()/conv_dw()        model/networks.py:136-163, :736-745
#include "common.cuh"
#include <math.h>
#include <stdlib.h>

namespace tdrn {

struct ConvP {
    const void *in; const float *w; const float *bias; const void *res; const float *off; void *out;
    int B, Cin, H, W, Cout, kh, kw, stride, pad, dil, Ho, Wo;
    int M, N, K;
    long long in_sb, in_sy, in_sx, in_sc;
    long long w_sn, w_sc, w_st;
    long long out_sb, out_sy, out_sx, out_sc;
    long long off_sb, off_sy, off_sx, off_sc;
    int dg, cpg, relu, deconv;
};

// Bilinear sample with the reference's border rules.  (y0,x0) = top-left tap origin
// (h_in,w_in at .cu:177-178), (ti,tj) tap, (oh,ow) learned offsets.  Arithmetic is kept in the
// reference's order with explicit round-to-nearest ops so the sampled column value is bit-identical
// to the scalar C oracle (no FMA contraction).
template <typename TIn>
__device__ __forceinline__ float deform_sample(const TIn *__restrict__ plane, long long sy, long long sx,
                                               int H, int W, int y0, int x0, int di, int dj, float oh, float ow)
{
    const float h_im = __fadd_rn((float)(y0 + di), oh);          // .cu:195
    const float w_im = __fadd_rn((float)(x0 + dj), ow);          // .cu:196
    if (!(h_im >= 0.f && w_im >= 0.f && h_im < (float)H && w_im < (float)W)) return 0.f;   // .cu:197
    float h = __fadd_rn((float)di, oh);                          // map_h .cu:198
    float w = __fadd_rn((float)dj, ow);                          // map_w .cu:199
    const int cur_h = H - y0, cur_w = W - x0;                    // .cu:200-201
    int h_low = (int)floorf(h), w_low = (int)floorf(w);          // .cu:21-22
    int h_high, w_high;
    if (h_low >= cur_h - 1) { h_high = h_low = cur_h - 1; h = (float)h_low; } else { h_high = h_low + 1; }
    if (w_low >= cur_w - 1) { w_high = w_low = cur_w - 1; w = (float)w_low; } else { w_high = w_low + 1; }
    const float lh = __fsub_rn(h, (float)h_low), lw = __fsub_rn(w, (float)w_low);
    const float hh = __fsub_rn(1.f, lh), hw = __fsub_rn(1.f, lw);
    // absolute coordinates; the clamp only matters for the measure-zero fp32 rounding case where the
    // reference itself would read one element outside the plane with weight 0.
    const int ya = min(max(y0 + h_low, 0), H - 1), yb = min(max(y0 + h_high, 0), H - 1);
    const int xa = min(max(x0 + w_low, 0), W - 1), xb = min(max(x0 + w_high, 0), W - 1);
    const float v1 = to_f32(plane[ya * sy + xa * sx]);
    const float v2 = to_f32(plane[ya * sy + xb * sx]);
    const float v3 = to_f32(plane[yb * sy + xa * sx]);
    const float v4 = to_f32(plane[yb * sy + xb * sx]);
    const float w1 = __fmul_rn(hh, hw), w2 = __fmul_rn(hh, lw), w3 = __fmul_rn(lh, hw), w4 = __fmul_rn(lh, lw);
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, v1), __fmul_rn(w2, v2)), __fmul_rn(w3, v3)), __fmul_rn(w4, v4));
}

template <typename TIn, typename TOut, int BM, int BN, bool DEFORM>
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvP p)
{
    constexpr int BK = 16, TM = BM / 16, TN = BN / 16;
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    __shared__ int row_b[BM], row_y[BM], row_x[BM];

    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const TIn *__restrict__ in = (const TIn *)p.in;

    for (int r = tid; r < BM; r += 256) {
        const int m = m0 + r;
        if (m < p.M) {
            const int hw = p.Ho * p.Wo;
            const int b = m / hw, rem = m - b * hw;
            row_b[r] = b; row_y[r] = rem / p.Wo; row_x[r] = rem % p.Wo;
        } else {
            row_b[r] = -1; row_y[r] = 0; row_x[r] = 0;
        }
    }
    __syncthreads();

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int kl = tid % BK, r0 = tid / BK;
    for (int k0 = 0; k0 < p.K; k0 += BK) {
        // ---- A tile: As[kl][r] = im2col(row r, k0+kl) ----
        {
            const int k = k0 + kl;
            const bool kvalid = k < p.K;
            const int tap = kvalid ? k / p.Cin : 0;
            const int c = kvalid ? k - tap * p.Cin : 0;
            const int ti = tap / p.kw, tj = tap - ti * p.kw;
#pragma unroll
            for (int i = 0; i < BM / 16; ++i) {
                const int r = r0 + i * 16;
                const int b = row_b[r];
                float v = 0.f;
                if (kvalid && b >= 0) {
                    const int y0 = row_y[r] * p.stride - p.pad, x0 = row_x[r] * p.stride - p.pad;
                    if (!DEFORM) {
                        const int yy = y0 + ti * p.dil, xx = x0 + tj * p.dil;
                        if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W)
                            v = to_f32(in[b * p.in_sb + yy * p.in_sy + xx * p.in_sx + c * p.in_sc]);
                    } else {
                        const int g = c / p.cpg;
                        const float *op = p.off + b * p.off_sb + row_y[r] * p.off_sy + row_x[r] * p.off_sx
                                          + (long long)(g * 2 * p.kh * p.kw + 2 * tap) * p.off_sc;
                        const float oh = op[0], ow = op[p.off_sc];
                        v = deform_sample<TIn>(in + b * p.in_sb + c * p.in_sc, p.in_sy, p.in_sx, p.H, p.W,
                                               y0, x0, ti * p.dil, tj * p.dil, oh, ow);
                    }
                }
                As[kl][r] = v;
            }
        }
        // ---- B tile: Bs[k][n] = weight(k0+k, n0+n) ----
        for (int idx = tid; idx < BK * BN; idx += 256) {
            const int kk = idx / BN, nl = idx - kk * BN;
            const int k = k0 + kk, n = n0 + nl;
            float v = 0.f;
            if (k < p.K && n < p.N) {
                const int tap = k / p.Cin, c = k - tap * p.Cin;
                if (!p.deconv) v = p.w[n * p.w_sn + c * p.w_sc + tap * p.w_st];
                else { const int ij = n / p.Cout, co = n - ij * p.Cout; v = p.w[co * p.w_sn + c * p.w_sc + ij * p.w_st]; }
            }
            Bs[kk][nl] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], bv[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    TOut *__restrict__ out = (TOut *)p.out;
    const TOut *__restrict__ res = (const TOut *)p.res;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int r = ty * TM + i;
        const int b = row_b[r];
        if (b < 0) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n >= p.N) continue;
            int co = n, oy = row_y[r], ox = row_x[r];
            if (p.deconv) { const int ij = n / p.Cout; co = n - ij * p.Cout; oy = 2 * oy + (ij >> 1); ox = 2 * ox + (ij & 1); }
            const long long o = b * p.out_sb + oy * p.out_sy + ox * p.out_sx + co * p.out_sc;
            float v = acc[i][j];
            if (p.bias) v += p.bias[co];
            if (res) v += to_f32(res[o]);
            if (p.relu) v = fmaxf(v, 0.f);
            out[o] = from_f32<TOut>(v);
        }
    }
}

template <typename TIn, typename TOut, bool DEFORM>
static int launch_conv_t(const ConvP &p, cudaStream_t st)
{
    if (p.N > 16) {
        dim3 grid(ceil_div(p.M, 128), ceil_div(p.N, 64));
        conv_simt_kernel<TIn, TOut, 128, 64, DEFORM><<<grid, 256, 0, st>>>(p);
    } else {
        dim3 grid(ceil_div(p.M, 64), 1);
        conv_simt_kernel<TIn, TOut, 64, 16, DEFORM><<<grid, 256, 0, st>>>(p);
    }
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

static int launch_conv(const ConvP &p, int in_dtype, int out_dtype, cudaStream_t st)
{
    const bool d = p.dg > 0;
    if (in_dtype == TDRN_F32 && out_dtype == TDRN_F32)
        return d ? launch_conv_t<float, float, true>(p, st) : launch_conv_t<float, float, false>(p, st);
    if (in_dtype == TDRN_BF16 && out_dtype == TDRN_BF16)
        return d ? launch_conv_t<__nv_bfloat16, __nv_bfloat16, true>(p, st) : launch_conv_t<__nv_bfloat16, __nv_bfloat16, false>(p, st);
    if (in_dtype == TDRN_BF16 && out_dtype == TDRN_F32)
        return d ? launch_conv_t<__nv_bfloat16, float, true>(p, st) : launch_conv_t<__nv_bfloat16, float, false>(p, st);
    if (in_dtype == TDRN_F32 && out_dtype == TDRN_BF16 && !d)
        return launch_conv_t<float, __nv_bfloat16, false>(p, st);
    set_error("conv: unsupported dtype combination in=%d out=%d deform=%d", in_dtype, out_dtype, (int)d);
    return TDRN_EUNSUPPORTED;
}

static int conv_out_dim(int in, int k, int stride, int pad, int dil) { return (in + 2 * pad - (dil * (k - 1) + 1)) / stride + 1; }

namespace tc { int launch_conv_stem_tc(const float *x, const float *w, const float *bias, void *out, int B, int H, int W, int relu,
                                       bool split, int stride, int cout, cudaStream_t st, bool f16 = false); }   // conv_stem_tc.cu
int launch_conv_first(const float *x, const float *w, const float *bias, void *out, int B, int H, int W, int Cout,
                      int Ho, int Wo, int stride, int relu, int out_dtype, cudaStream_t st);   // conv_first.cu

// ------------------------------------------------------------------------------------------------
// depthwise 3x3, pad 1 (conv_dw first half, model/networks.py:738-740), NHWC, weight [9][C]
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void dwconv3x3_kernel(const T *__restrict__ in, const float *__restrict__ w, const float *__restrict__ bias,
                                 T *__restrict__ out, int B, int H, int W, int C, int Ho, int Wo, int stride, int relu)
{
    const long long total = (long long)B * Ho * Wo * C;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        long long pix = idx / C;
        const int x = (int)(pix % Wo); pix /= Wo;
        const int y = (int)(pix % Ho);
        const int b = (int)(pix / Ho);
        float acc = bias ? bias[c] : 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int yy = y * stride - 1 + i;
            if (yy < 0 || yy >= H) continue;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int xx = x * stride - 1 + j;
                if (xx < 0 || xx >= W) continue;
                acc = fmaf(to_f32(in[(((long long)b * H + yy) * W + xx) * C + c]), w[(i * 3 + j) * C + c], acc);
            }
        }
        if (relu) acc = fmaxf(acc, 0.f);
        out[idx] = from_f32<T>(acc);
    }
}

// bf16 fast path: a thread owns 8 channels (one 16-byte vector) of XT consecutive output pixels of a row, so every
// input vector it loads feeds up to three outputs; consecutive threads take consecutive channel groups (coalesced
// 16-byte accesses); packed fp32x2 FMAs.  HBM-bound by design (the scalar kernel above was ~30x off).
template <bool F16>
__device__ __forceinline__ unsigned long long dw_unpack(uint32_t a)      // packed bf16 (or IEEE half) pair -> (lo, hi) fp32
{
    unsigned long long r;
    if (F16) {
        asm("{\n\t.reg .b16 l, h;\n\t.reg .f32 fl, fh;\n\tmov.b32 {l, h}, %1;\n\tcvt.f32.f16 fl, l;\n\tcvt.f32.f16 fh, h;\n\tmov.b64 %0, {fl, fh};\n\t}"
            : "=l"(r) : "r"(a));
    } else {
        asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a << 16), "r"(a & 0xffff0000u));
    }
    return r;
}
__device__ __forceinline__ unsigned long long dw_fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long dw_pair(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}

template <int STRIDE, bool F16IN, bool F16OUT>
__global__ void __launch_bounds__(128) dwconv3x3_bf16_kernel(const uint4 *__restrict__ in, const float *__restrict__ w,
                                                             const float *__restrict__ bias, uint4 *__restrict__ out,
                                                             int B, int H, int W, int C, int Ho, int Wo, int relu)
{
    // grid (ceil(XS * CG / 128), Ho, B): one 32-bit division per thread (the flat 64-bit index of the first version cost three
    // 64-bit divisions = ~200 of the kernel's ~1 000 instructions per thread, and the kernel is issue-bound: ncu r02z2)
    constexpr int XT = 4, NCOL = (XT - 1) * STRIDE + 3;
    const int CG = C >> 3, XS = (Wo + XT - 1) / XT;
    const unsigned e = blockIdx.x * 128u + threadIdx.x;
    if (e >= (unsigned)(XS * CG)) return;
    const int xs = (int)(e / (unsigned)CG), cg = (int)(e - (unsigned)xs * (unsigned)CG);
    const int y = blockIdx.y, b = blockIdx.z;
    const int x0 = xs * XT;
    unsigned long long acc[XT][4];
    {
        const float4 b0 = bias ? __ldg((const float4 *)(bias + cg * 8)) : make_float4(0, 0, 0, 0);
        const float4 b1 = bias ? __ldg((const float4 *)(bias + cg * 8 + 4)) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int p = 0; p < XT; ++p) { acc[p][0] = dw_pair(b0.x, b0.y); acc[p][1] = dw_pair(b0.z, b0.w); acc[p][2] = dw_pair(b1.x, b1.y); acc[p][3] = dw_pair(b1.z, b1.w); }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int yy = y * STRIDE - 1 + i;
        if (yy < 0 || yy >= H) continue;
        unsigned long long wt[3][4];                              // this kernel row's 3 taps x 8 channels
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float4 w0 = __ldg((const float4 *)(w + (i * 3 + j) * C + cg * 8)), w1 = __ldg((const float4 *)(w + (i * 3 + j) * C + cg * 8 + 4));
            wt[j][0] = dw_pair(w0.x, w0.y); wt[j][1] = dw_pair(w0.z, w0.w); wt[j][2] = dw_pair(w1.x, w1.y); wt[j][3] = dw_pair(w1.z, w1.w);
        }
        const uint4 *row = in + (((long long)b * H + yy) * W) * CG + cg;
        uint4 v[NCOL];                                            // all columns of this row in flight together (requesting all
                                                                  // three rows up front was measured slower: fewer resident warps)
#pragma unroll
        for (int c = 0; c < NCOL; ++c) {
            const int xx = x0 * STRIDE - 1 + c;
            v[c] = (xx >= 0 && xx < W) ? __ldg(row + (long long)xx * CG) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int c = 0; c < NCOL; ++c) {
            const unsigned long long u[4] = {dw_unpack<F16IN>(v[c].x), dw_unpack<F16IN>(v[c].y), dw_unpack<F16IN>(v[c].z), dw_unpack<F16IN>(v[c].w)};
#pragma unroll
            for (int p = 0; p < XT; ++p) {
                const int j = c - p * STRIDE;                      // tap column of input column c for output p
                if (j < 0 || j > 2) continue;                      // compile-time after unrolling
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] = dw_fma2(wt[j][q], u[q], acc[p][q]);
            }
        }
    }
#pragma unroll
    for (int p = 0; p < XT; ++p) {
        const int x = x0 + p;
        if (x >= Wo) break;
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float lo, hi;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[p][q]));
            if (F16OUT) {
                asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(o[q]) : "f"(hi), "f"(lo));
                if (relu) asm("max.f16x2 %0, %0, %1;" : "+r"(o[q]) : "r"(0u));
            } else {
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o[q]) : "f"(hi), "f"(lo));
                if (relu) asm("max.bf16x2 %0, %0, %1;" : "+r"(o[q]) : "r"(0u));     // max commutes with the (monotonic) rounding
            }
        }
        out[(((long long)b * Ho + y) * Wo + x) * CG + cg] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// IEEE-half in and out (the MobileNet trunk's format, TDRN_F16): the nine taps are packed-half FMAs on the loaded vectors
// themselves -- no unpacking, no 64-bit register pairs, half the FMA issue slots of the fp32x2 form above, which was issue-bound
// (~1 000 instructions per thread for 32 outputs, 2.6 TB/s).  Weights and bias (fp32 in memory, as for the other kernels) are
// rounded to half per thread; the accumulator is half: every FMA rounds to 11 significand bits.  CPU emulation of exactly this
// arithmetic inside the whole detector: conf 1.36e-2 against 1.33e-2 with fp32 accumulation.  The result
// is clamped to the largest finite half.
__device__ __forceinline__ uint32_t h2_fma(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t h2_pack(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

template <int STRIDE>
__global__ void __launch_bounds__(128) dwconv3x3_half_kernel(const uint4 *__restrict__ in, const float *__restrict__ w,
                                                             const float *__restrict__ bias, uint4 *__restrict__ out,
                                                             int B, int H, int W, int C, int Ho, int Wo, int relu)
{
    constexpr int XT = 4, NCOL = (XT - 1) * STRIDE + 3;
    const int CG = C >> 3, XS = (Wo + XT - 1) / XT;
    const unsigned e = blockIdx.x * 128u + threadIdx.x;
    pdl_sync();
    if (e >= (unsigned)(XS * CG)) return;
    const int xs = (int)(e / (unsigned)CG), cg = (int)(e - (unsigned)xs * (unsigned)CG);
    const int y = blockIdx.y, b = blockIdx.z;
    const int x0 = xs * XT;
    uint32_t acc[XT][4];
    {
        const float4 b0 = bias ? __ldg((const float4 *)(bias + cg * 8)) : make_float4(0, 0, 0, 0);
        const float4 b1 = bias ? __ldg((const float4 *)(bias + cg * 8 + 4)) : make_float4(0, 0, 0, 0);
        const uint32_t bh[4] = {h2_pack(b0.x, b0.y), h2_pack(b0.z, b0.w), h2_pack(b1.x, b1.y), h2_pack(b1.z, b1.w)};
#pragma unroll
        for (int p = 0; p < XT; ++p)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[p][q] = bh[q];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int yy = y * STRIDE - 1 + i;
        if (yy < 0 || yy >= H) continue;
        uint32_t wt[3][4];                                        // this kernel row's 3 taps x 8 channels, packed half
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float4 w0 = __ldg((const float4 *)(w + (i * 3 + j) * C + cg * 8)), w1 = __ldg((const float4 *)(w + (i * 3 + j) * C + cg * 8 + 4));
            wt[j][0] = h2_pack(w0.x, w0.y); wt[j][1] = h2_pack(w0.z, w0.w); wt[j][2] = h2_pack(w1.x, w1.y); wt[j][3] = h2_pack(w1.z, w1.w);
        }
        const uint4 *row = in + (((long long)b * H + yy) * W) * CG + cg;
        uint4 v[NCOL];
#pragma unroll
        for (int c = 0; c < NCOL; ++c) {
            const int xx = x0 * STRIDE - 1 + c;
            v[c] = (xx >= 0 && xx < W) ? __ldg(row + (long long)xx * CG) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int c = 0; c < NCOL; ++c) {
            const uint32_t u[4] = {v[c].x, v[c].y, v[c].z, v[c].w};
#pragma unroll
            for (int p = 0; p < XT; ++p) {
                const int j = c - p * STRIDE;                      // tap column of input column c for output p
                if (j < 0 || j > 2) continue;                      // compile-time after unrolling
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] = h2_fma(wt[j][q], u[q], acc[p][q]);
            }
        }
    }
#pragma unroll
    for (int p = 0; p < XT; ++p) {
        const int x = x0 + p;
        if (x >= Wo) break;
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            o[q] = acc[p][q];
            if (relu) asm("max.f16x2 %0, %0, %1;" : "+r"(o[q]) : "r"(0u));
            else asm("max.f16x2 %0, %0, %1;" : "+r"(o[q]) : "r"(0xfbfffbffu));       // -Inf -> -65504
            asm("min.f16x2 %0, %0, %1;" : "+r"(o[q]) : "r"(0x7bff7bffu));               // +Inf -> 65504
        }
        out[(((long long)b * Ho + y) * Wo + x) * CG + cg] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// Stride-1 form for large batches: a thread walks YT consecutive output rows of its 4-pixel x 8-channel strip and keeps the
// three input rows it needs in registers, so every input vector is loaded ONCE per thread (not once per output row), the
// weights are converted once per thread, and the next input row is requested a whole output row ahead.  Four row buffers
// rotate by name through a 4x unrolled loop (no register moves).
template <int YT>
__global__ void __launch_bounds__(128, 3) dwconv3x3_half_roll_kernel(const uint4 *__restrict__ in, const float *__restrict__ w,
                                                                     const float *__restrict__ bias, uint4 *__restrict__ out,
                                                                     int B, int H, int W, int C, int relu)
{
    constexpr int XT = 4, NCOL = XT + 2;
    const int CG = C >> 3, XS = (W + XT - 1) / XT;                  // stride 1, pad 1: Ho = H, Wo = W
    const unsigned e = blockIdx.x * 128u + threadIdx.x;
    pdl_sync();
    if (e >= (unsigned)(XS * CG)) return;
    const int xs = (int)(e / (unsigned)CG), cg = (int)(e - (unsigned)xs * (unsigned)CG);
    const int y_begin = blockIdx.y * YT, y_end = min(H, y_begin + YT), b = blockIdx.z;
    const int x0 = xs * XT;
    uint32_t wt[9][4], bh[4];
    {
        const float4 b0 = bias ? __ldg((const float4 *)(bias + cg * 8)) : make_float4(0, 0, 0, 0);
        const float4 b1 = bias ? __ldg((const float4 *)(bias + cg * 8 + 4)) : make_float4(0, 0, 0, 0);
        bh[0] = h2_pack(b0.x, b0.y); bh[1] = h2_pack(b0.z, b0.w); bh[2] = h2_pack(b1.x, b1.y); bh[3] = h2_pack(b1.z, b1.w);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const float4 w0 = __ldg((const float4 *)(w + t * C + cg * 8)), w1 = __ldg((const float4 *)(w + t * C + cg * 8 + 4));
            wt[t][0] = h2_pack(w0.x, w0.y); wt[t][1] = h2_pack(w0.z, w0.w); wt[t][2] = h2_pack(w1.x, w1.y); wt[t][3] = h2_pack(w1.z, w1.w);
        }
    }
    bool cv[NCOL];                                                  // column inside the map?
#pragma unroll
    for (int c = 0; c < NCOL; ++c) cv[c] = (unsigned)(x0 - 1 + c) < (unsigned)W;
    const uint4 *col0 = in + ((long long)b * H * W + (x0 - 1)) * CG + cg;      // (image b, row 0, column x0 - 1)
    const long long rstride = (long long)W * CG;
    struct Row { uint4 v[NCOL]; };
    auto load_row = [&](Row &r, int yy) {
        const bool rv = (unsigned)yy < (unsigned)H;
        const uint4 *row = col0 + (long long)yy * rstride;
#pragma unroll
        for (int c = 0; c < NCOL; ++c) r.v[c] = (rv && cv[c]) ? __ldg(row + c * CG) : make_uint4(0, 0, 0, 0);
    };
    auto emit = [&](const Row &ra, const Row &rb, const Row &rc, int y) {      // output row y from input rows y-1, y, y+1
        uint32_t acc[XT][4];
#pragma unroll
        for (int p = 0; p < XT; ++p)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[p][q] = bh[q];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const Row &r = i == 0 ? ra : (i == 1 ? rb : rc);
#pragma unroll
            for (int c = 0; c < NCOL; ++c) {
                const uint32_t u[4] = {r.v[c].x, r.v[c].y, r.v[c].z, r.v[c].w};
#pragma unroll
                for (int p = 0; p < XT; ++p) {
                    const int j = c - p;
                    if (j < 0 || j > 2) continue;
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[p][q] = h2_fma(wt[i * 3 + j][q], u[q], acc[p][q]);
                }
            }
        }
        uint4 *orow = out + (((long long)b * H + y) * W + x0) * CG + cg;
#pragma unroll
        for (int p = 0; p < XT; ++p) {
            if (x0 + p >= W) break;
            uint32_t o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                o[q] = acc[p][q];
                if (relu) asm("max.f16x2 %0, %0, %1;" : "+r"(o[q]) : "r"(0u));
                else asm("max.f16x2 %0, %0, %1;" : "+r"(o[q]) : "r"(0xfbfffbffu));
                asm("min.f16x2 %0, %0, %1;" : "+r"(o[q]) : "r"(0x7bff7bffu));
            }
            orow[p * CG] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    };
    Row r0, r1, r2, r3;
    load_row(r0, y_begin - 1); load_row(r1, y_begin); load_row(r2, y_begin + 1);
    for (int y = y_begin; y < y_end; y += 4) {
        load_row(r3, y + 2); emit(r0, r1, r2, y);
        if (y + 1 >= y_end) break;
        load_row(r0, y + 3); emit(r1, r2, r3, y + 1);
        if (y + 2 >= y_end) break;
        load_row(r1, y + 4); emit(r2, r3, r0, y + 2);
        if (y + 3 >= y_end) break;
        load_row(r2, y + 5); emit(r3, r0, r1, y + 3);
    }
}

// ------------------------------------------------------------------------------------------------
// MaxPool2d(2, 2, ceil_mode) NHWC
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void maxpool2x2_kernel(const T *__restrict__ in, T *__restrict__ out, int B, int H, int W, int C, int Ho, int Wo)
{
    const long long total = (long long)B * Ho * Wo * C;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        long long pix = idx / C;
        const int x = (int)(pix % Wo); pix /= Wo;
        const int y = (int)(pix % Ho);
        const int b = (int)(pix / Ho);
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int yy = 2 * y + i, xx = 2 * x + j;
                if (yy < H && xx < W) m = fmaxf(m, to_f32(in[(((long long)b * H + yy) * W + xx) * C + c]));
            }
        out[idx] = from_f32<T>(m);
    }
}

// ------------------------------------------------------------------------------------------------
// L2Norm (layers/modules/l2norm.py:17-21): one warp per pixel
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void l2norm_kernel(const T *__restrict__ in, const float *__restrict__ weight, T *__restrict__ out, long long pixels, int C)
{
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = warp; p < pixels; p += nwarps) {
        const T *row = in + p * C;
        float s = 0.f;
        for (int c = lane; c < C; c += 32) { const float v = to_f32(row[c]); s = fmaf(v, v, s); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float norm = sqrtf(s) + 1e-10f;
        for (int c = lane; c < C; c += 32) out[p * C + c] = from_f32<T>(weight[c] * (to_f32(row[c]) / norm));
    }
}

// ------------------------------------------------------------------------------------------------
// L2Norm + the MaxPool2d(2,2) that follows the same tensor (conv4_3 -> {L2Norm_4_3, pool4}, conv5_3 ->
// {L2Norm_5_3, pool5}; model/dualrefinedet_vggbn.py:130-148): x is read ONCE, both consumers' inputs are written.
// bf16 NHWC, one warp per 2x2 window, 16-byte vector accesses (lane l owns channels v*256 + 8l .. +7).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void l2_unpack(uint32_t a, int f16, float &lo, float &hi)     // packed bf16 / IEEE half pair -> fp32
{
    if (f16) {
        asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(lo), "=f"(hi) : "r"(a));
    } else {
        lo = __uint_as_float(a << 16); hi = __uint_as_float(a & 0xffff0000u);
    }
}

template <int CV>
__global__ void __launch_bounds__(256, 4) l2norm_pool_kernel(const uint4 *__restrict__ in, const float *__restrict__ weight,
                                                             uint4 *__restrict__ out_norm, uint4 *__restrict__ out_pool,
                                                             int B, int H, int W, int f16in = 0, int f16out = 0)
{
    const int lane = threadIdx.x & 31;
    const int Ho = H >> 1, Wo = W >> 1;
    const long long win = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (win >= (long long)B * Ho * Wo) return;
    const int b = (int)(win / (Ho * Wo)), rem = (int)(win - (long long)b * Ho * Wo);
    const int py = rem / Wo, px = rem - py * Wo;
    constexpr int PV = CV * 32;                                  // uint4 per pixel
    const long long p00 = ((long long)b * H + 2 * py) * W + 2 * px;
    // pass 1: per-pixel sum of squares + running 2x2 max (x is re-read from L1 in pass 2: keeps the register
    // footprint small enough for 32 resident warps per SM, which is what a streaming kernel needs)
    float ss[4];
    uint4 mx[CV];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 *src = in + (p00 + (q >> 1) * W + (q & 1)) * PV + lane;
        float s = 0.f;
#pragma unroll
        for (int v = 0; v < CV; ++v) {
            const uint4 xv = __ldg(src + v * 32);
            const uint32_t w4[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float lo, hi;
                l2_unpack(w4[j], f16in, lo, hi);
                s = fmaf(lo, lo, s); s = fmaf(hi, hi, s);
            }
            if (q == 0) mx[v] = xv;
            else if (out_pool) {
                const __nv_bfloat162 *x2 = (const __nv_bfloat162 *)&xv;
                __nv_bfloat162 *m2 = (__nv_bfloat162 *)&mx[v];
#pragma unroll
                for (int j = 0; j < 4; ++j) m2[j] = __hmax2(m2[j], x2[j]);
            }
        }
        ss[q] = s;
    }
    if (out_pool) {                                              // NULL: plain L2Norm (any grouping of 4 pixels per warp)
        const long long po = ((long long)b * Ho + py) * Wo + px;
#pragma unroll
        for (int v = 0; v < CV; ++v) out_pool[po * PV + v * 32 + lane] = mx[v];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int q = 0; q < 4; ++q) ss[q] += __shfl_xor_sync(0xffffffffu, ss[q], o);
    // pass 2: normalise.  x / norm is evaluated as x * (1 / norm): the result is rounded to bf16 (2^-9), the
    // reciprocal's extra 2^-24 is invisible; the fp32 path keeps the exact division (l2norm_kernel).
#pragma unroll
    for (int v = 0; v < CV; ++v) {
        const float4 wa = __ldg((const float4 *)(weight + v * 256 + lane * 8)), wb = __ldg((const float4 *)(weight + v * 256 + lane * 8 + 4));
        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float inv = 1.f / (sqrtf(ss[q]) + 1e-10f);     // l2norm.py:18
            const long long e = (p00 + (q >> 1) * W + (q & 1)) * PV + v * 32 + lane;
            const uint4 xv = in[e];
            const uint32_t w4[4] = {xv.x, xv.y, xv.z, xv.w};
            uint32_t o4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float lo, hi;
                l2_unpack(w4[j], f16in, lo, hi);
                const float rl = wv[2 * j] * (lo * inv), rh = wv[2 * j + 1] * (hi * inv);                               // :19-20
                if (f16out) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(o4[j]) : "f"(rh), "f"(rl));
                else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o4[j]) : "f"(rh), "f"(rl));
            }
            out_norm[e] = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Row softmax (nn.Softmax(dim=1)): one warp per row
// ------------------------------------------------------------------------------------------------
__global__ void softmax_rows_kernel(const float *__restrict__ in, float *__restrict__ out, long long rows, int C)
{
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp; r < rows; r += nwarps) {
        const float *row = in + r * C;
        float m = -INFINITY;
        for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += expf(row[c] - m);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        for (int c = lane; c < C; c += 32) out[r * C + c] = expf(row[c] - m) / s;
    }
}

// ------------------------------------------------------------------------------------------------
// layout transforms (tiled through shared memory so both sides are coalesced)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T *__restrict__ in, float *__restrict__ out, int HW, int C)
{
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int p = p0 + i, c = c0 + threadIdx.x;
        if (p < HW && c < C) tile[i][threadIdx.x] = to_f32(in[((long long)b * HW + p) * C + c]);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, p = p0 + threadIdx.x;
        if (p < HW && c < C) out[((long long)b * C + c) * HW + p] = tile[threadIdx.x][i];
    }
}

template <typename T>
__global__ void nchw_to_nhwc_kernel(const float *__restrict__ in, T *__restrict__ out, int C, int HW)
{
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, p = p0 + threadIdx.x;
        if (p < HW && c < C) tile[i][threadIdx.x] = in[((long long)b * C + c) * HW + p];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int p = p0 + i, c = c0 + threadIdx.x;
        if (p < HW && c < C) out[((long long)b * HW + p) * C + c] = from_f32<T>(tile[threadIdx.x][i]);
    }
}

static int grid_1d(long long total, int block) { long long g = (total + block - 1) / block; return (int)(g > 148LL * 32 ? 148 * 32 : (g < 1 ? 1 : g)); }

}  // namespace tdrn

using namespace tdrn;

extern "C" int tdrn_conv2d(const tdrn_conv_desc *d, const void *in, const float *weight, const float *bias,
                           const void *residual, const float *offsets, void *out, tdrn_stream_t stream)
{
    TDRN_REQUIRE(d && in && weight && out, "tdrn_conv2d: null argument");
    TDRN_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "tdrn_conv2d: bad shape");
    TDRN_REQUIRE((d->dg > 0) == (offsets != nullptr), "tdrn_conv2d: offsets must be given iff dg > 0");
    TDRN_REQUIRE(d->dg == 0 || d->Cin % d->dg == 0, "tdrn_conv2d: Cin %% dg != 0");
    TDRN_REQUIRE(!(d->deconv2x2 && d->dg), "tdrn_conv2d: deconv and deform are exclusive");
    if (d->pool2x2) { set_error("tdrn_conv2d: fused max-pool is only implemented by tdrn_conv2d_tc"); return TDRN_EUNSUPPORTED; }
    ConvP p{};
    p.in = in; p.w = weight; p.bias = bias; p.res = residual; p.off = offsets; p.out = out;
    p.B = d->B; p.Cin = d->Cin; p.H = d->H; p.W = d->W; p.Cout = d->Cout;
    p.relu = d->relu; p.deconv = d->deconv2x2; p.dg = d->dg; p.cpg = d->dg > 0 ? d->Cin / d->dg : d->Cin;
    int out_w;
    if (d->deconv2x2) {
        p.kh = p.kw = 1; p.stride = 1; p.pad = 0; p.dil = 1; p.Ho = d->H; p.Wo = d->W;
        p.N = 4 * d->Cout; p.K = d->Cin;
        p.w_sc = 4LL * d->Cout; p.w_st = d->Cout; p.w_sn = 1;      // packed [Cin][ij][Cout]
        out_w = 2 * d->W;
    } else {
        p.kh = d->kh; p.kw = d->kw; p.stride = d->stride; p.pad = d->pad; p.dil = d->dil;
        p.Ho = conv_out_dim(d->H, d->kh, d->stride, d->pad, d->dil);
        p.Wo = conv_out_dim(d->W, d->kw, d->stride, d->pad, d->dil);
        TDRN_REQUIRE(p.Ho > 0 && p.Wo > 0, "convolution input is too small (output would be %dx%d)", p.Ho, p.Wo);
        p.N = d->Cout; p.K = d->kh * d->kw * d->Cin;
        p.w_st = (long long)d->Cin * d->Cout; p.w_sc = d->Cout; p.w_sn = 1;   // packed [tap][cin][cout]
        out_w = p.Wo;
    }
    p.M = d->B * p.Ho * p.Wo;
    p.in_sb = d->in_sb > 0 ? d->in_sb : (long long)d->H * d->W * d->Cin; p.in_sy = (long long)d->W * d->Cin; p.in_sx = d->Cin; p.in_sc = 1;
    p.out_sb = d->out_sb; p.out_sx = d->out_sp; p.out_sy = (long long)out_w * d->out_sp; p.out_sc = 1;
    if (d->dg > 0) {
        const long long oc = (long long)d->dg * 2 * d->kh * d->kw;
        p.off_sb = (long long)p.Ho * p.Wo * oc; p.off_sy = (long long)p.Wo * oc; p.off_sx = oc; p.off_sc = 1;
    }
    return launch_conv(p, d->in_dtype, d->out_dtype, as_stream(stream));
}

extern "C" int tdrn_conv_first(const float *x, const float *weight, const float *bias, void *out, int B, int H,
                               int W, int Cout, int stride, int relu, int out_dtype, tdrn_stream_t stream)
{
    TDRN_REQUIRE(x && weight && out && B > 0 && H > 0 && W > 0 && Cout > 0, "tdrn_conv_first: bad argument");
    TDRN_REQUIRE(stride == 1 || stride == 2, "tdrn_conv_first: stride must be 1 or 2");
    if (out_dtype == TDRN_BF16_SPLIT) {      // fp32-accurate path: (hi | lo) operand of conv1_2, tensor cores only
        if (Cout != 64 || stride != 1) { set_error("tdrn_conv_first: TDRN_BF16_SPLIT output needs Cout = 64, stride 1"); return TDRN_EUNSUPPORTED; }
        return tc::launch_conv_stem_tc(x, weight, bias, out, B, H, W, relu, true, 1, 64, as_stream(stream));
    }
    if (((Cout == 64 && stride == 1) || (Cout == 32 && stride == 2)) && (out_dtype == TDRN_BF16 || out_dtype == TDRN_F16) && !getenv("TDRN_STEM_SIMT")) {
        // tcgen05 K = 32 stem (conv_stem_tc.cu): VGG conv1_1, MobileNet conv_bn(3, 32, 2); TDRN_F16: half operands and output
        const int rc = tc::launch_conv_stem_tc(x, weight, bias, out, B, H, W, relu, false, stride, Cout, as_stream(stream), out_dtype == TDRN_F16);
        if (rc != TDRN_EUNSUPPORTED) return rc;
    }
    if (out_dtype == TDRN_F16) { set_error("tdrn_conv_first: half output needs a map the tensor-core stem tiles"); return TDRN_EUNSUPPORTED; }
    if (Cout % 16 == 0 && Cout <= 64)            // register-tiled direct kernel (conv_first.cu)
        return launch_conv_first(x, weight, bias, out, B, H, W, Cout, conv_out_dim(H, 3, stride, 1, 1),
                                 conv_out_dim(W, 3, stride, 1, 1), stride, relu, out_dtype, as_stream(stream));
    ConvP p{};
    p.in = x; p.w = weight; p.bias = bias; p.out = out;
    p.B = B; p.Cin = 3; p.H = H; p.W = W; p.Cout = Cout; p.kh = p.kw = 3; p.stride = stride; p.pad = 1; p.dil = 1;
    p.Ho = conv_out_dim(H, 3, stride, 1, 1); p.Wo = conv_out_dim(W, 3, stride, 1, 1);
    p.M = B * p.Ho * p.Wo; p.N = Cout; p.K = 27; p.relu = relu; p.cpg = 3;
    p.in_sb = 3LL * H * W; p.in_sc = (long long)H * W; p.in_sy = W; p.in_sx = 1;        // NCHW image
    p.w_st = 3LL * Cout; p.w_sc = Cout; p.w_sn = 1;
    p.out_sb = (long long)p.Ho * p.Wo * Cout; p.out_sy = (long long)p.Wo * Cout; p.out_sx = Cout; p.out_sc = 1;
    return launch_conv(p, TDRN_F32, out_dtype, as_stream(stream));
}

extern "C" int tdrn_deform_conv_forward(const float *input, const float *weight, const float *offset, float *output,
                                        int B, int Cin, int H, int W, int Cout, int kW, int kH, int dW, int dH,
                                        int padW, int padH, int dilationH, int dilationW, int deformable_group,
                                        tdrn_stream_t stream)
{
    // shape_check, utils/deformconv/deform_conv_cuda.c:7-96
    TDRN_REQUIRE(input && weight && offset && output, "deform_conv_forward: null tensor");
    TDRN_REQUIRE(kW > 0 && kH > 0, "kernel size should be greater than zero, but got kH: %d kW: %d", kH, kW);
    TDRN_REQUIRE(dW > 0 && dH > 0, "stride should be greater than zero, but got dH: %d dW: %d", dH, dW);
    TDRN_REQUIRE(dilationW > 0 && dilationH > 0, "dilation should be greater than 0, but got dilationH: %d dilationW: %d", dilationH, dilationW);
    TDRN_REQUIRE(B > 0 && Cin > 0 && H > 0 && W > 0 && Cout > 0, "deform_conv_forward: bad shape");
    TDRN_REQUIRE(deformable_group > 0 && Cin % deformable_group == 0, "input channels must divide deformable group size");
    TDRN_REQUIRE(dW == dH && padW == padH && dilationW == dilationH, "deform_conv_forward: only square stride/pad/dilation are implemented");
    ConvP p{};
    p.in = input; p.w = weight; p.off = offset; p.out = output;
    p.B = B; p.Cin = Cin; p.H = H; p.W = W; p.Cout = Cout; p.kh = kH; p.kw = kW; p.stride = dH; p.pad = padH; p.dil = dilationH;
    p.Ho = conv_out_dim(H, kH, dH, padH, dilationH); p.Wo = conv_out_dim(W, kW, dW, padW, dilationW);
    TDRN_REQUIRE(p.Ho > 0 && p.Wo > 0, "Given input size: (%d x %d x %d). Calculated output size: (%d x %d x %d). Output size is too small",
                 Cin, H, W, Cout, p.Ho, p.Wo);
    p.M = B * p.Ho * p.Wo; p.N = Cout; p.K = kH * kW * Cin; p.dg = deformable_group; p.cpg = Cin / deformable_group;
    p.in_sb = (long long)Cin * H * W; p.in_sc = (long long)H * W; p.in_sy = W; p.in_sx = 1;
    p.w_sn = (long long)Cin * kH * kW; p.w_sc = (long long)kH * kW; p.w_st = 1;             // reference [Cout,Cin,kH,kW]
    const long long hw = (long long)p.Ho * p.Wo;
    p.out_sb = Cout * hw; p.out_sc = hw; p.out_sy = p.Wo; p.out_sx = 1;
    p.off_sb = (long long)deformable_group * 2 * kH * kW * hw; p.off_sc = hw; p.off_sy = p.Wo; p.off_sx = 1;
    return launch_conv(p, TDRN_F32, TDRN_F32, as_stream(stream));
}

template <int STRIDE>
static void launch_dw16(const void *in, const float *weight, const float *bias, void *out, int B, int H, int W, int C, int Ho, int Wo,
                        int relu, bool f16in, bool f16out, cudaStream_t st)
{
    const dim3 grid((unsigned)((((Wo + 3) / 4) * (C / 8) + 127) / 128), (unsigned)Ho, (unsigned)B);
#define TDRN_DW(FI, FO) dwconv3x3_bf16_kernel<STRIDE, FI, FO><<<grid, 128, 0, st>>>((const uint4 *)in, weight, bias, (uint4 *)out, B, H, W, C, Ho, Wo, relu)
    static const bool f32acc = getenv("TDRN_DW_F32ACC") != nullptr;    // half in/out with fp32 accumulation (the form the packed-half kernel replaced)
    if (f16in && f16out && !f32acc) {
        // TDRN_DW_ROLL_MIN (read per call so that tests can switch it): CTAs from which the row-walking form is used; 0 = never
        const char *rm = getenv("TDRN_DW_ROLL_MIN");
        const long long roll_min = rm ? atoll(rm) : 1024;
        constexpr int YT = 8;
        const long long roll_blocks = (long long)grid.x * ((Ho + YT - 1) / YT) * B;
        if (STRIDE == 1 && roll_min > 0 && roll_blocks >= roll_min) {   // enough CTAs to fill the GPU several times over: row-walking form
            const dim3 g2(grid.x, (unsigned)((Ho + YT - 1) / YT), (unsigned)B);
            launch_pdl(dwconv3x3_half_roll_kernel<YT>, g2, dim3(128), 0, st, (const uint4 *)in, weight, bias, (uint4 *)out, B, H, W, C, relu);
            return;
        }
        launch_pdl(dwconv3x3_half_kernel<STRIDE>, grid, dim3(128), 0, st, (const uint4 *)in, weight, bias, (uint4 *)out, B, H, W, C, Ho, Wo, relu);
        return;
    }
    if (f16in) { if (f16out) TDRN_DW(true, true); else TDRN_DW(true, false); }
    else { if (f16out) TDRN_DW(false, true); else TDRN_DW(false, false); }
#undef TDRN_DW
}

extern "C" int tdrn_dwconv3x3_io(const void *in, const float *weight, const float *bias, void *out, int B, int H, int W,
                                 int C, int stride, int relu, int in_dtype, int out_dtype, tdrn_stream_t stream)
{
    TDRN_REQUIRE(in && weight && out && B > 0 && H > 0 && W > 0 && C > 0, "tdrn_dwconv3x3: bad argument");
    const int Ho = conv_out_dim(H, 3, stride, 1, 1), Wo = conv_out_dim(W, 3, stride, 1, 1);
    const long long total = (long long)B * Ho * Wo * C;
    const bool in16 = in_dtype == TDRN_BF16 || in_dtype == TDRN_F16, out16 = out_dtype == TDRN_BF16 || out_dtype == TDRN_F16;
    if (in16 && out16 && C % 8 == 0 && (stride == 1 || stride == 2)) {
        TDRN_REQUIRE(Ho <= 65535 && B <= 65535, "tdrn_dwconv3x3: map too tall / batch too large for the grid");
        if (stride == 1) launch_dw16<1>(in, weight, bias, out, B, H, W, C, Ho, Wo, relu, in_dtype == TDRN_F16, out_dtype == TDRN_F16, as_stream(stream));
        else launch_dw16<2>(in, weight, bias, out, B, H, W, C, Ho, Wo, relu, in_dtype == TDRN_F16, out_dtype == TDRN_F16, as_stream(stream));
        TDRN_LAUNCH_CHECK();
        return TDRN_OK;
    }
    if (in_dtype != out_dtype || in_dtype == TDRN_F16) {
        set_error("tdrn_dwconv3x3: half / mixed formats need C %% 8 == 0 and stride 1 or 2 (got C=%d stride=%d in=%d out=%d)", C, stride, in_dtype, out_dtype);
        return TDRN_EUNSUPPORTED;
    }
    if (in_dtype == TDRN_F32)
        dwconv3x3_kernel<float><<<grid_1d(total, 256), 256, 0, as_stream(stream)>>>((const float *)in, weight, bias, (float *)out, B, H, W, C, Ho, Wo, stride, relu);
    else
        dwconv3x3_kernel<__nv_bfloat16><<<grid_1d(total, 256), 256, 0, as_stream(stream)>>>((const __nv_bfloat16 *)in, weight, bias, (__nv_bfloat16 *)out, B, H, W, C, Ho, Wo, stride, relu);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

extern "C" int tdrn_dwconv3x3(const void *in, const float *weight, const float *bias, void *out, int B, int H, int W,
                              int C, int stride, int relu, int dtype, tdrn_stream_t stream)
{
    return tdrn_dwconv3x3_io(in, weight, bias, out, B, H, W, C, stride, relu, dtype, dtype, stream);
}

extern "C" int tdrn_maxpool2x2(const void *in, void *out, int B, int H, int W, int C, int ceil_mode, int dtype, tdrn_stream_t stream)
{
    TDRN_REQUIRE(in && out && B > 0 && H > 0 && W > 0 && C > 0, "tdrn_maxpool2x2: bad argument");
    const int Ho = ceil_mode ? (H + 1) / 2 : H / 2, Wo = ceil_mode ? (W + 1) / 2 : W / 2;
    const long long total = (long long)B * Ho * Wo * C;
    if (dtype == TDRN_F32)
        maxpool2x2_kernel<float><<<grid_1d(total, 256), 256, 0, as_stream(stream)>>>((const float *)in, (float *)out, B, H, W, C, Ho, Wo);
    else
        maxpool2x2_kernel<__nv_bfloat16><<<grid_1d(total, 256), 256, 0, as_stream(stream)>>>((const __nv_bfloat16 *)in, (__nv_bfloat16 *)out, B, H, W, C, Ho, Wo);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

extern "C" int tdrn_l2norm_io(const void *in, const float *weight, void *out, long long pixels, int C, int in_dtype, int out_dtype,
                              tdrn_stream_t stream)
{
    TDRN_REQUIRE(in && weight && out && pixels > 0 && C > 0, "tdrn_l2norm: bad argument");
    const bool in16 = in_dtype == TDRN_BF16 || in_dtype == TDRN_F16, out16 = out_dtype == TDRN_BF16 || out_dtype == TDRN_F16;
    if (in16 && out16 && (C == 256 || C == 512 || C == 1024) && pixels % 4 == 0 && pixels / 2 < (1LL << 30)) {
        // vectorised path: the fused L2Norm+pool kernel without its pool output, 4 consecutive pixels per warp
        const int Wv = (int)(pixels / 2);
        const int g = (int)((pixels / 4 * 32 + 255) / 256);
        const int fi = in_dtype == TDRN_F16, fo = out_dtype == TDRN_F16;
        cudaStream_t st = as_stream(stream);
        if (C == 256) l2norm_pool_kernel<1><<<g, 256, 0, st>>>((const uint4 *)in, weight, (uint4 *)out, nullptr, 1, 2, Wv, fi, fo);
        else if (C == 512) l2norm_pool_kernel<2><<<g, 256, 0, st>>>((const uint4 *)in, weight, (uint4 *)out, nullptr, 1, 2, Wv, fi, fo);
        else l2norm_pool_kernel<4><<<g, 256, 0, st>>>((const uint4 *)in, weight, (uint4 *)out, nullptr, 1, 2, Wv, fi, fo);
        TDRN_LAUNCH_CHECK();
        return TDRN_OK;
    }
    if (in_dtype != out_dtype || in_dtype == TDRN_F16) {
        set_error("tdrn_l2norm: half / mixed formats need C in {256, 512, 1024} and pixels %% 4 == 0 (got C=%d pixels=%lld in=%d out=%d)",
                  C, pixels, in_dtype, out_dtype);
        return TDRN_EUNSUPPORTED;
    }
    const int grid = grid_1d(pixels * 32, 256);
    if (in_dtype == TDRN_F32)
        l2norm_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float *)in, weight, (float *)out, pixels, C);
    else
        l2norm_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>((const __nv_bfloat16 *)in, weight, (__nv_bfloat16 *)out, pixels, C);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

extern "C" int tdrn_l2norm(const void *in, const float *weight, void *out, long long pixels, int C, int dtype, tdrn_stream_t stream)
{
    return tdrn_l2norm_io(in, weight, out, pixels, C, dtype, dtype, stream);
}

extern "C" int tdrn_l2norm_pool2x2(const void *in, const float *weight, void *out_norm, void *out_pool, int B, int H, int W,
                                   int C, int dtype, tdrn_stream_t stream)
{
    TDRN_REQUIRE(in && weight && out_norm && out_pool && B > 0 && H > 0 && W > 0 && C > 0, "tdrn_l2norm_pool2x2: bad argument");
    if (dtype != TDRN_BF16 || (C != 256 && C != 512 && C != 1024) || (H & 1) || (W & 1)) {
        set_error("tdrn_l2norm_pool2x2: needs bf16, C in {256,512,1024}, even H and W (got dtype=%d C=%d %dx%d)", dtype, C, H, W);
        return TDRN_EUNSUPPORTED;
    }
    const long long windows = (long long)B * (H / 2) * (W / 2);
    const int grid = (int)((windows * 32 + 255) / 256);
    cudaStream_t st = as_stream(stream);
    if (C == 256) l2norm_pool_kernel<1><<<grid, 256, 0, st>>>((const uint4 *)in, weight, (uint4 *)out_norm, (uint4 *)out_pool, B, H, W);
    else if (C == 512) l2norm_pool_kernel<2><<<grid, 256, 0, st>>>((const uint4 *)in, weight, (uint4 *)out_norm, (uint4 *)out_pool, B, H, W);
    else l2norm_pool_kernel<4><<<grid, 256, 0, st>>>((const uint4 *)in, weight, (uint4 *)out_norm, (uint4 *)out_pool, B, H, W);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

extern "C" int tdrn_softmax(const float *in, float *out, long long rows, int C, tdrn_stream_t stream)
{
    TDRN_REQUIRE(in && out && rows > 0 && C > 0, "tdrn_softmax: bad argument");
    softmax_rows_kernel<<<grid_1d(rows * 32, 256), 256, 0, as_stream(stream)>>>(in, out, rows, C);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

extern "C" int tdrn_nhwc_to_nchw_f32(const void *in, float *out, int B, int H, int W, int C, int dtype, tdrn_stream_t stream)
{
    TDRN_REQUIRE(in && out && B > 0 && H > 0 && W > 0 && C > 0, "tdrn_nhwc_to_nchw_f32: bad argument");
    dim3 grid(ceil_div(H * W, 32), ceil_div(C, 32), B), block(32, 8);
    if (dtype == TDRN_F32) nhwc_to_nchw_kernel<float><<<grid, block, 0, as_stream(stream)>>>((const float *)in, out, H * W, C);
    else nhwc_to_nchw_kernel<__nv_bfloat16><<<grid, block, 0, as_stream(stream)>>>((const __nv_bfloat16 *)in, out, H * W, C);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

extern "C" int tdrn_nchw_f32_to_nhwc(const float *in, void *out, int B, int C, int H, int W, int dtype, tdrn_stream_t stream)
{
    TDRN_REQUIRE(in && out && B > 0 && H > 0 && W > 0 && C > 0, "tdrn_nchw_f32_to_nhwc: bad argument");
    dim3 grid(ceil_div(H * W, 32), ceil_div(C, 32), B), block(32, 8);
    if (dtype == TDRN_F32) nchw_to_nhwc_kernel<float><<<grid, block, 0, as_stream(stream)>>>(in, (float *)out, C, H * W);
    else nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, block, 0, as_stream(stream)>>>(in, (__nv_bfloat16 *)out, C, H * W);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

// ---------------------------------------------------------------------------------------------------------
// fp32 -> (hi | lo) bf16 split: the A operand of the fp32-accurate tensor-core convs (conv_tc.cu, split3).
// One thread = 8 channels of one pixel: two float4 loads, one 16-byte store of the high parts at [p][c0..], one of
// the low parts at [p][C + c0..].
// ---------------------------------------------------------------------------------------------------------
namespace tdrn {
__global__ void __launch_bounds__(256) split_bf16_kernel(const float4 *__restrict__ in, uint4 *__restrict__ out, long long pixels, int C8)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pixels * C8) return;
    const long long p = i / C8;
    const int c8 = (int)(i - p * C8);
    const float4 a = in[2 * i], b = in[2 * i + 1];
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint4 hi, lo;
    __nv_bfloat162 *h2 = (__nv_bfloat162 *)&hi, *l2 = (__nv_bfloat162 *)&lo;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * j]), h1 = __float2bfloat16_rn(v[2 * j + 1]);
        h2[j] = __halves2bfloat162(h0, h1);
        l2[j] = __floats2bfloat162_rn(__fsub_rn(v[2 * j], __bfloat162float(h0)), __fsub_rn(v[2 * j + 1], __bfloat162float(h1)));
    }
    out[p * (2 * C8) + c8] = hi;
    out[p * (2 * C8) + C8 + c8] = lo;
}
}  // namespace tdrn

extern "C" int tdrn_split_bf16(const float *in, void *out, long long pixels, int C, tdrn_stream_t stream)
{
    TDRN_REQUIRE(in && out && pixels > 0 && C > 0 && C % 8 == 0, "tdrn_split_bf16: bad argument (C %% 8 == 0 required, got %d)", C);
    const long long n = pixels * (C / 8);
    tdrn::split_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, tdrn::as_stream(stream)>>>((const float4 *)in, (uint4 *)out, pixels, C / 8);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

// ---------------------------------------------------------------------------------------------------------
// offset / offset2 1x1 convs of all pyramid levels in ONE launch (dualrefinedet_vggbn.py:160-164, ssd4scale_vgg.py:72-76):
// per pixel a 12-vector (the ARM regression of its three anchors, read straight out of the flattened [B,P,4] tensor) times
// a [c1 | c2] x 12 matrix + bias -> NHWC fp32 offset maps for the deformable heads and, for the first head, the NCHW copy
// the reference returns.  Replaces 8 conv launches + 4 layout copies per step (each a few microseconds of work on its own
// branch stream).  One thread per (image, pixel, output channel).
// ---------------------------------------------------------------------------------------------------------
namespace tdrn {
struct OffsetLevels {
    int n, c1, c2, B, P;
    int H[TDRN_MAX_OFFSET_LEVELS], W[TDRN_MAX_OFFSET_LEVELS], prior_off[TDRN_MAX_OFFSET_LEVELS];
    long long first[TDRN_MAX_OFFSET_LEVELS + 1];           // running (pixel, channel) work-item offsets per image
    const float *w1[TDRN_MAX_OFFSET_LEVELS], *b1[TDRN_MAX_OFFSET_LEVELS], *w2[TDRN_MAX_OFFSET_LEVELS], *b2[TDRN_MAX_OFFSET_LEVELS];
    float *o1[TDRN_MAX_OFFSET_LEVELS], *o2[TDRN_MAX_OFFSET_LEVELS], *o1_nchw[TDRN_MAX_OFFSET_LEVELS];
};

constexpr int OFF_PX = 64;      // pixels per block

// grid (blocks of all levels, B); a block owns OFF_PX pixels of one level of one image: the level's [c1 + c2] x 12 weights,
// biases and the block's 12-vectors are staged in shared memory once, then one (pixel, channel) output per thread and step.
__global__ void __launch_bounds__(256) offset_convs_kernel(const float *__restrict__ arm_loc, const OffsetLevels L)
{
    extern __shared__ float off_smem[];            // w [cc][12] | bias [cc] | arm [OFF_PX][12]
    const int b = blockIdx.y;
    int k = 0;
    while (k + 1 < L.n && (long long)blockIdx.x >= L.first[k + 1]) ++k;      // first[] = running block counts here
    const int cc = L.c1 + L.c2;
    const int HW = L.H[k] * L.W[k];
    const int pix0 = (int)(blockIdx.x - L.first[k]) * OFF_PX;
    const int npx = min(OFF_PX, HW - pix0);
    float *w = off_smem, *bias = w + cc * 12, *arm = bias + cc;
    for (int i = threadIdx.x; i < cc * 12; i += blockDim.x) w[i] = i < L.c1 * 12 ? L.w1[k][i] : L.w2[k][i - L.c1 * 12];
    for (int i = threadIdx.x; i < cc; i += blockDim.x)
        bias[i] = i < L.c1 ? (L.b1[k] ? L.b1[k][i] : 0.f) : (L.b2[k] ? L.b2[k][i - L.c1] : 0.f);
    const float *a = arm_loc + ((long long)b * L.P + L.prior_off[k] + (long long)pix0 * 3) * 4;     // npx * 12 consecutive floats
    for (int i = threadIdx.x; i < npx * 12; i += blockDim.x) arm[i] = a[i];
    __syncthreads();
    for (int i = threadIdx.x; i < npx * cc; i += blockDim.x) {
        const int px = i / cc, ch = i - px * cc;
        const float *av = arm + px * 12, *wv = w + ch * 12;
        float acc = bias[ch];
#pragma unroll
        for (int c = 0; c < 12; ++c) acc = fmaf(av[c], wv[c], acc);
        const int pix = pix0 + px;
        if (ch >= L.c1) {
            L.o2[k][((long long)b * HW + pix) * L.c2 + (ch - L.c1)] = acc;
        } else {
            L.o1[k][((long long)b * HW + pix) * L.c1 + ch] = acc;
            if (L.o1_nchw[k]) L.o1_nchw[k][((long long)b * L.c1 + ch) * HW + pix] = acc;
        }
    }
}
}  // namespace tdrn

extern "C" int tdrn_offset_convs(const float *arm_loc, int B, int P, int n_levels, const tdrn_offset_level *lv, int c1, int c2,
                                 tdrn_stream_t stream)
{
    TDRN_REQUIRE(arm_loc && lv && B > 0 && P > 0 && n_levels > 0 && n_levels <= TDRN_MAX_OFFSET_LEVELS && c1 > 0 && c2 >= 0,
                 "tdrn_offset_convs: bad argument");
    tdrn::OffsetLevels L{};
    L.n = n_levels; L.c1 = c1; L.c2 = c2; L.B = B; L.P = P;
    long long run = 0;
    for (int k = 0; k < n_levels; ++k) {
        TDRN_REQUIRE(lv[k].H > 0 && lv[k].W > 0 && lv[k].w1 && lv[k].out1 && (c2 == 0 || (lv[k].w2 && lv[k].out2)) &&
                     lv[k].prior_off >= 0 && lv[k].prior_off + 3ll * lv[k].H * lv[k].W <= P, "tdrn_offset_convs: bad level %d", k);
        L.H[k] = lv[k].H; L.W[k] = lv[k].W; L.prior_off[k] = lv[k].prior_off;
        L.w1[k] = lv[k].w1; L.b1[k] = lv[k].b1; L.w2[k] = lv[k].w2; L.b2[k] = lv[k].b2;
        L.o1[k] = lv[k].out1; L.o2[k] = lv[k].out2; L.o1_nchw[k] = lv[k].out1_nchw;
        L.first[k] = run;
        run += ((long long)lv[k].H * lv[k].W + tdrn::OFF_PX - 1) / tdrn::OFF_PX;      // blocks of this level
    }
    L.first[n_levels] = run;
    const size_t smem = (size_t)((c1 + c2) * 13 + tdrn::OFF_PX * 12) * sizeof(float);
    TDRN_REQUIRE(smem <= 48 * 1024, "tdrn_offset_convs: too many offset channels (%d + %d)", c1, c2);
    tdrn::offset_convs_kernel<<<dim3((unsigned)run, (unsigned)B), 256, smem, tdrn::as_stream(stream)>>>(arm_loc, L);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}
