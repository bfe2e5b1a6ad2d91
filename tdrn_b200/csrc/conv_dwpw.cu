// conv_dwpw.cu -- one conv_dw block of the MobileNet trunks (model/networks.py:736-745: depthwise 3x3 + BN + ReLU, then
// pointwise 1x1 + BN + ReLU; dualrefinedet_mobilenet.py:23-35, ssd4scale_mobile.py) as ONE kernel.
//
// The two-kernel path (tdrn_dwconv3x3, then tdrn_conv2d_tc) writes the depthwise output to HBM and reads it back:
// 2 x B*Ho*Wo*Cin bf16 per block, as much as the block's own input + output.  Here the depthwise result never leaves the SM:
//   warp 18: TMA of the input halo patch of a tile and channel block -- box (64 ch, bw + 2, bh + 2, bn) of the NHWC input, the
//       conv zero padding is the out-of-bounds fill; 3-stage ring; the block's 9 x 64 depthwise weights and 64 biases are
//       copied next to the patch by the producers themselves (a first version read the input with global loads from the producer warps: 8 warps cannot keep enough
//       bytes in flight, it ran at 11 000 cycles per channel block instead of the ~1 100 the arithmetic needs);
//   warps 10-17 (256 threads): depthwise 3x3 on the CUDA cores out of that patch (16-byte shared loads of 8 channels, fp32
//       FMAs in the same order as dwconv3x3_bf16_kernel -> the same values), bias + ReLU, rounded to bf16 and written as the
//       128B-swizzled K-major A tile [128 output pixels][64 channels] of a 2-stage ring;
//   warp 0: TMA of the pointwise weight boxes [<= 256 couts][64 channels] (2-stage ring);
//   warp 1: tcgen05.mma, all couts of a pass (<= 512 = the whole TMEM) accumulate at once, so every A tile is produced
//       once per pass (Cout = 1024: two passes); passes of <= 256 couts alternate between two accumulator sets;
//   warps 2-9: epilogue, TMEM -> bias / ReLU -> bf16 -> swizzled staging box -> TMA store (as conv_tc.cu).
// Same operands, same accumulation order over K as the two-kernel path: the output is bit-identical to it (tested).
//
// MEASURED (B200, b64, scripts/dwpw_timing.py, profiles/r02y_dwpw_timing.txt): correct but SLOWER than the two kernels on every
// MobileNet-320 layer (512 -> 512 @40x40: 0.249 ms against 0.081 + 0.074 ms), so the model keeps the two-kernel path and this one
// is opt-in (TDRN_DWPW=1).  With the producers' arithmetic switched off the kernel takes 0.117 ms, with the epilogue off as well
// 0.079 ms (= the pointwise GEMM's own L2 -> SM operand traffic: every 128-pixel tile re-reads the 512 KB weight matrix); the
// depthwise arithmetic adds 0.115 ms: per 64-channel block a producer warp issues ~830 instructions (144 FFMA, ~150 unpack
// shifts / masks, 36 LDS, swizzled addresses), two warps per scheduler, FFMA and the unpack at one instruction per two cycles
// -> >= 2 300 cycles per block against 512 for the block's MMAs.  The stand-alone depthwise kernel spends the same
// instructions but across all four schedulers with 16+ resident warps each.
#include "tc_common.cuh"
#include <stdlib.h>

namespace tdrn {
namespace tc {

void pick_box(int B, int H, int W, int max_w, int max_h, int &bw, int &bh, int &bn);      // conv_tc.cu

struct DwPwP {
    const uint4 *in;              // [B,H,W,Cin] bf16
    const float *dw_w, *dw_b;     // [9][Cin] fp32 (BN folded), [Cin] or NULL
    const float *pw_b;            // [Cout] or NULL
    int B, H, W, Cin, Cout, stride, Ho, Wo;
    int bw, bh, bn, tiles_w, tiles_h, m_tiles;
    int cblocks;                  // ceil(Cin / 64)
    int n_pass, np;               // passes over the couts; couts per pass (multiple of 16, <= 512)
    uint32_t b_box_bytes;         // bytes one weight box deposits (min(np, 256) rows x 128 B; rows beyond the tensor are zero-filled)
    int relu_dw, relu_pw;
    int dbg;                      // timing experiments (TDRN_DWPW_DEBUG; results are wrong): 1 producers skip the arithmetic, 2 no weight
                                  // loads / MMAs wait for nothing of B, 4 epilogue skips its work
};

constexpr int DP_THREADS = 608;
constexpr int DP_PROD0 = 320;        // first producer thread
constexpr int DP_A_STAGES = 2, DP_B_STAGES = 2, DP_P_STAGES = 3;
constexpr int DP_A_BYTES = 128 * 128, DP_B_BYTES = 256 * 128, DP_O_BYTES = 128 * 128;
constexpr int DP_PATCH_PX = 216;                             // pixels of the largest halo patch (40 x 3 tile: 42 x 5 = 210)
constexpr int DP_W_OFF = DP_PATCH_PX * 128;                  // stage layout: patch | 9 x 64 fp32 weights | 64 fp32 biases
constexpr int DP_BIAS_OFF = DP_W_OFF + 9 * 256;
constexpr int DP_P_BYTES = (DP_BIAS_OFF + 256 + 1023) & ~1023;    // 30 KB
constexpr int DP_SMEM = 1024 + DP_A_STAGES * DP_A_BYTES + DP_B_STAGES * DP_B_BYTES + DP_P_STAGES * DP_P_BYTES + 2 * DP_O_BYTES;

// explicit shared-space accesses: the ring pointers are derived from an aligned-up address, and through that cast the
// compiler no longer knows they are shared (it emitted generic LD.E.128 for the patch reads)
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}


__device__ __forceinline__ unsigned long long dp_unpack(uint32_t a)      // two bf16 -> two fp32 (packed pair)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a << 16), "r"(a & 0xffff0000u));
    return r;
}
__device__ __forceinline__ unsigned long long dp_fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long dp_pair(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}

__global__ void __launch_bounds__(DP_THREADS, 1) conv_dwpw_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                  const __grid_constant__ CUtensorMap tmB,
                                                                  const __grid_constant__ CUtensorMap tmO, const DwPwP p)
{
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t a_full[DP_A_STAGES], a_empty[DP_A_STAGES];
    __shared__ __align__(8) uint64_t b_full[DP_B_STAGES], b_empty[DP_B_STAGES];
    __shared__ __align__(8) uint64_t p_full[DP_P_STAGES], p_empty[DP_P_STAGES];
    __shared__ __align__(8) uint64_t t_full[2], t_empty[2];
    __shared__ uint32_t tmem_base_s;

    uint8_t *base = (uint8_t *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = base;
    uint8_t *sB = sA + DP_A_STAGES * DP_A_BYTES;
    uint8_t *sP = sB + DP_B_STAGES * DP_B_BYTES;
    uint8_t *sO = sP + DP_P_STAGES * DP_P_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_sub = (p.np + 255) >> 8;                       // MMA sub-tiles (<= 256 couts each) of a pass
    const uint32_t set_cols = (uint32_t)n_sub * 256u;          // TMEM columns of one accumulator set
    const uint32_t nbuf = set_cols <= 256u ? 2u : 1u;
    const int total = p.m_tiles * p.n_pass;                    // unit = pass * m_tiles + mt: concurrent CTAs share a weight slice

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmO);
#pragma unroll
        for (int s = 0; s < DP_A_STAGES; ++s) { mbar_init(&a_full[s], 8); mbar_init(&a_empty[s], 1); }
#pragma unroll
        for (int s = 0; s < DP_B_STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
#pragma unroll
        for (int s = 0; s < DP_P_STAGES; ++s) { mbar_init(&p_full[s], 1); mbar_init(&p_empty[s], 8); }
#pragma unroll
        for (int b = 0; b < 2; ++b) { mbar_init(&t_full[b], 1); mbar_init(&t_empty[b], 8); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===================== pointwise-weight TMA producer =====================
        if (elect_one()) {
            uint32_t jt = 0;
            for (int unit = blockIdx.x; unit < total; unit += gridDim.x) {
                const int n0 = (unit / p.m_tiles) * p.np;
                for (int kb = 0; kb < p.cblocks; ++kb)
                    for (int h = 0; h < n_sub; ++h, ++jt) {
                        const uint32_t s = jt % DP_B_STAGES, ph = (jt / DP_B_STAGES) & 1u;
                        mbar_wait(&b_empty[s], ph ^ 1u);
                        mbar_expect_tx(&b_full[s], p.b_box_bytes);
                        tma_load_2d(sB + (size_t)s * DP_B_BYTES, &tmB, &b_full[s], kb * 64, n0 + h * 256);
                    }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (elect_one()) {
            uint32_t it = 0, jt = 0, ucount = 0;
            for (int unit = blockIdx.x; unit < total; unit += gridDim.x, ++ucount) {
                const uint32_t buf = ucount % nbuf;
                mbar_wait(&t_empty[buf], ((ucount / nbuf) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * set_cols;
                for (int kb = 0; kb < p.cblocks; ++kb, ++it) {
                    const uint32_t sa = it % DP_A_STAGES, pha = (it / DP_A_STAGES) & 1u;
                    mbar_wait(&a_full[sa], pha);
                    tc_fence_after();
                    const uint64_t adesc = umma_desc_sw128(smem_u32(sA + (size_t)sa * DP_A_BYTES));
                    for (int h = 0; h < n_sub; ++h, ++jt) {
                        const uint32_t sb = jt % DP_B_STAGES, phb = (jt / DP_B_STAGES) & 1u;
                        mbar_wait(&b_full[sb], phb);
                        tc_fence_after();
                        const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + (size_t)sb * DP_B_BYTES));
                        const uint32_t idesc = umma_idesc_bf16(128, min(256, p.np - h * 256));
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(d_tmem + (uint32_t)h * 256u, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                        umma_commit(&b_empty[sb]);
                    }
                    umma_commit(&a_empty[sa]);
                }
                umma_commit(&t_full[buf]);
            }
        }
        __syncwarp();
    } else if (warp < 10) {
        // ===================== epilogue (warps 2..9): two warps per TMEM lane quadrant =====================
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int r = quad * 32 + lane;
        const bool leader = threadIdx.x == 64;
        const int tiles_per_img = p.tiles_w * p.tiles_h;
        uint32_t ucount = 0, git = 0;
        for (int unit = blockIdx.x; unit < total; unit += gridDim.x, ++ucount) {
            const int pass = unit / p.m_tiles, mt = unit - pass * p.m_tiles;
            const int tn = mt / tiles_per_img, rem = mt - tn * tiles_per_img;
            const int x0 = (rem % p.tiles_w) * p.bw, y0 = (rem / p.tiles_w) * p.bh, b0 = tn * p.bn;
            const int n0 = pass * p.np;
            const uint32_t buf = ucount % nbuf;
            mbar_wait(&t_full[buf], (ucount / nbuf) & 1u);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * set_cols;
            const int groups = (p.np + 63) >> 6;
            for (int g = 0; g < groups; ++g, ++git) {
                uint8_t *o = sO + (git & 1u) * DP_O_BYTES;
                if (leader) bulk_wait_read<1>();                 // the store that last read this box has drained
                named_bar(1, 256);
                const int c0 = g * 64 + half * 32;
                if (p.dbg & 4) {
                    if (g == groups - 1) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(&t_empty[buf]); }
                } else
                if (c0 < p.np) {                                 // warp-uniform; columns beyond Cout are clipped by the store
                    float v[32];
                    tmem_ld32(trow + (uint32_t)c0, v);
                    if (g == groups - 1) {                       // all tcgen05.ld of this unit are complete
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&t_empty[buf]);
                    }
                    const int n = n0 + c0;
                    if (p.pw_b) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (n + j < p.Cout) v[j] += __ldg(p.pw_b + n + j);
                    }
                    if (p.relu_pw) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 w;
                        __nv_bfloat162 *wb = (__nv_bfloat162 *)&w;
#pragma unroll
                        for (int j = 0; j < 4; ++j) wb[j] = __floats2bfloat162_rn(v[q * 8 + 2 * j], v[q * 8 + 2 * j + 1]);
                        *(uint4 *)(o + sw128_offset(r, half * 4 + q)) = w;
                    }
                } else if (g == groups - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&t_empty[buf]);
                }
                fence_proxy_async_smem();
                named_bar(1, 256);
                if (leader && !(p.dbg & 4)) {
                    tma_store_4d(&tmO, o, n0 + g * 64, x0, y0, b0);
                    bulk_commit();
                }
            }
        }
        if (leader) bulk_wait_read<0>();
    } else if (warp == 18) {
        // ===================== input halo patch + depthwise weights: TMA producer =====================
        if (elect_one()) {
            const int tiles_per_img = p.tiles_w * p.tiles_h;
            const uint32_t patch_bytes = (uint32_t)((p.bw + 2) * (p.bh + 2) * p.bn) * 128u;
            uint32_t it = 0;
            for (int unit = blockIdx.x; unit < total; unit += gridDim.x) {
                const int mt = unit % p.m_tiles;
                const int tn = mt / tiles_per_img, rem = mt - tn * tiles_per_img;
                const int x0 = (rem % p.tiles_w) * p.bw, y0 = (rem / p.tiles_w) * p.bh, b0 = tn * p.bn;
                for (int kb = 0; kb < p.cblocks; ++kb, ++it) {
                    const uint32_t ps = it % DP_P_STAGES, php = (it / DP_P_STAGES) & 1u;
                    mbar_wait(&p_empty[ps], php ^ 1u);
                    uint8_t *st = sP + (size_t)ps * DP_P_BYTES;
                    mbar_expect_tx(&p_full[ps], patch_bytes);
                    tma_load_4d(st, &tmX, &p_full[ps], kb * 64, x0 - 1, y0 - 1, b0);
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== depthwise producers (warps 10..17) =====================
        // thread = (channel group of 8, pixel slot): rows slot, slot + 32, slot + 64, slot + 96 of the A tile
        const int pt = threadIdx.x - DP_PROD0;
        const int cg = pt & 7, slot = pt >> 3;
        const int PW = p.bw + 2, PHW = (p.bw + 2) * (p.bh + 2);
        int pix[4];                                              // patch pixel of tap (0, 0) of each of the thread's rows; < 0: no pixel
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = slot + 32 * q;
            const int wl = r % p.bw, hl = (r / p.bw) % p.bh, nl = r / (p.bw * p.bh);
            pix[q] = nl < p.bn ? nl * PHW + hl * PW + wl : -1;
        }
        uint32_t it = 0;
        for (int unit = blockIdx.x; unit < total; unit += gridDim.x) {
            for (int kb = 0; kb < p.cblocks; ++kb, ++it) {
                const uint32_t sa = it % DP_A_STAGES, pha = (it / DP_A_STAGES) & 1u;
                const uint32_t ps = it % DP_P_STAGES, php = (it / DP_P_STAGES) & 1u;
                const bool chan_ok = kb * 64 + cg * 8 < p.Cin;
                const uint32_t st = smem_u32(sP) + ps * (uint32_t)DP_P_BYTES;
                const uint32_t a = smem_u32(sA) + sa * (uint32_t)DP_A_BYTES;
                {   // this block's 9 x 64 depthwise weights + 64 biases -> the stage (160 float4, one per thread; zeros beyond Cin)
                    float4 wv = make_float4(0, 0, 0, 0);
                    if (pt < 160) {
                        const int row = pt >> 4, ch = kb * 64 + (pt & 15) * 4;
                        if (ch < p.Cin) {
                            if (row < 9) wv = __ldg((const float4 *)(p.dw_w + (size_t)row * p.Cin + ch));
                            else if (p.dw_b) wv = __ldg((const float4 *)(p.dw_b + ch));
                        }
                    }
                    mbar_wait(&p_full[ps], php);                 // the halo patch has landed (the weight load above overlaps the wait)
                    if (pt < 160)
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st + DP_W_OFF + pt * 16), "f"(wv.x), "f"(wv.y), "f"(wv.z), "f"(wv.w) : "memory");
                    named_bar(2, 256);                           // all producers: weights visible; nobody is still in the previous block
                }
                float4 b0v = make_float4(0, 0, 0, 0), b1v = b0v;
                if (chan_ok && p.dw_b) { b0v = lds128f(st + DP_BIAS_OFF + cg * 32); b1v = lds128f(st + DP_BIAS_OFF + cg * 32 + 16); }
#pragma unroll 1
                for (int q2 = 0; q2 < 4; q2 += 2) {              // two pixels at a time (register budget)
                    // scalar FMAs: the packed fma.rn.f32x2 form needs its operands in aligned register pairs, and assembling those
                    // from the unpacked bf16 halves cost more moves than the packing saved (same IEEE results either way)
                    float acc[2][8];
#pragma unroll
                    for (int qq = 0; qq < 2; ++qq) {
                        acc[qq][0] = b0v.x; acc[qq][1] = b0v.y; acc[qq][2] = b0v.z; acc[qq][3] = b0v.w;
                        acc[qq][4] = b1v.x; acc[qq][5] = b1v.y; acc[qq][6] = b1v.z; acc[qq][7] = b1v.w;
                    }
                    const int pq[2] = {q2 == 0 ? pix[0] : pix[2], q2 == 0 ? pix[1] : pix[3]};
                    if (chan_ok && !(p.dbg & 1)) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
#pragma unroll
                            for (int j = 0; j < 3; ++j) {
                                const float4 w0 = lds128f(st + DP_W_OFF + (i * 3 + j) * 256 + cg * 32);
                                const float4 w1 = lds128f(st + DP_W_OFF + (i * 3 + j) * 256 + cg * 32 + 16);
                                const float wt[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                                for (int qq = 0; qq < 2; ++qq) {
                                    if (pq[qq] < 0) continue;
                                    const uint4 v = lds128(st + sw128_offset(pq[qq] + i * PW + j, cg));
                                    const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        acc[qq][2 * e] = fmaf(wt[2 * e], __uint_as_float(vw[e] << 16), acc[qq][2 * e]);
                                        acc[qq][2 * e + 1] = fmaf(wt[2 * e + 1], __uint_as_float(vw[e] & 0xffff0000u), acc[qq][2 * e + 1]);
                                    }
                                }
                            }
                        }
                    }
                    if (q2 == 0) mbar_wait(&a_empty[sa], pha ^ 1u);   // the MMAs that read this stage last have retired
#pragma unroll
                    for (int qq = 0; qq < 2; ++qq) {
                        uint32_t o[4] = {0u, 0u, 0u, 0u};
                        if (chan_ok && pq[qq] >= 0) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float lo = acc[qq][2 * e], hi = acc[qq][2 * e + 1];
                                if (p.relu_dw) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
                                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o[e]) : "f"(hi), "f"(lo));
                            }
                        }
                        sts128(a + sw128_offset(slot + 32 * (q2 + qq), cg), make_uint4(o[0], o[1], o[2], o[3]));
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&a_full[sa]); mbar_arrive(&p_empty[ps]); }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace tc
}  // namespace tdrn

using namespace tdrn;
using namespace tdrn::tc;

extern "C" int tdrn_conv_dwpw(const tdrn_dwpw_desc *d, const void *in, const float *dw_weight, const float *dw_bias,
                              const void *pw_weight, const float *pw_bias, void *out, tdrn_stream_t stream)
{
    TDRN_REQUIRE(d && in && dw_weight && pw_weight && out, "tdrn_conv_dwpw: null argument");
    TDRN_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "tdrn_conv_dwpw: bad shape");
    if (d->Cin % 8 != 0 || d->Cout % 8 != 0 || d->stride != 1 || ((uintptr_t)in & 15) || ((uintptr_t)out & 15) ||
        ((uintptr_t)dw_weight & 15) || ((uintptr_t)dw_bias & 15)) {
        set_error("tdrn_conv_dwpw: needs Cin %% 8 == 0, Cout %% 8 == 0, stride 1, 16-byte aligned tensors (got Cin=%d Cout=%d stride=%d)",
                  d->Cin, d->Cout, d->stride);
        return TDRN_EUNSUPPORTED;
    }
    DwPwP p{};
    p.in = (const uint4 *)in; p.dw_w = dw_weight; p.dw_b = dw_bias; p.pw_b = pw_bias;
    p.B = d->B; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout; p.stride = d->stride;
    p.Ho = (d->H + 2 - 3) / d->stride + 1; p.Wo = (d->W + 2 - 3) / d->stride + 1;
    p.relu_dw = d->relu_dw; p.relu_pw = d->relu_pw;
    { const char *e = getenv("TDRN_DWPW_DEBUG"); p.dbg = e ? atoi(e) : 0; }
    // tile = box of <= 128 output pixels whose halo patch (bw + 2) x (bh + 2) x bn fits a ring stage
    pick_box(p.B, p.Ho, p.Wo, 40, 16, p.bw, p.bh, p.bn);
    if ((p.bw + 2) * (p.bh + 2) * p.bn > DP_PATCH_PX) pick_box(p.B, p.Ho, p.Wo, 16, 16, p.bw, p.bh, p.bn);
    while ((p.bw + 2) * (p.bh + 2) * p.bn > DP_PATCH_PX && p.bn > 1) --p.bn;
    p.tiles_w = (p.Wo + p.bw - 1) / p.bw; p.tiles_h = (p.Ho + p.bh - 1) / p.bh;
    p.m_tiles = p.tiles_w * p.tiles_h * ((p.B + p.bn - 1) / p.bn);
    p.cblocks = (d->Cin + 63) / 64;
    const int n_pad16 = (d->Cout + 15) & ~15;
    p.n_pass = (n_pad16 + 511) / 512;
    p.np = n_pad16;
    if (p.n_pass > 1) p.np = ((n_pad16 + p.n_pass - 1) / p.n_pass + 63) & ~63;     // whole 64-column store groups per pass
    p.b_box_bytes = (uint32_t)(p.np < 256 ? p.np : 256) * 128u;
    TDRN_REQUIRE((p.bw + 2) * (p.bh + 2) * p.bn <= DP_PATCH_PX, "tdrn_conv_dwpw: halo patch of a %dx%dx%d tile does not fit", p.bw, p.bh, p.bn);
    CUtensorMap tmX, tmB, tmO;
    {   // input halo patch: box (64 ch, bw + 2, bh + 2, bn) at (kb*64, x0 - 1, y0 - 1, b0); out of bounds = zero = the conv padding
        const uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
        const uint64_t str[3] = {(uint64_t)d->Cin * 2, (uint64_t)d->W * d->Cin * 2, (uint64_t)d->H * d->W * d->Cin * 2};
        const uint32_t box[4] = {64, (uint32_t)(p.bw + 2), (uint32_t)(p.bh + 2), (uint32_t)p.bn};
        int rc = make_tmap_bf16(&tmX, in, 4, dims, str, box, nullptr);
        if (rc) return rc;
    }
    {
        const uint64_t K = (uint64_t)p.cblocks * 64;                  // packed [Cout_pad16][Cin_pad64] bf16
        const uint64_t dims[2] = {K, (uint64_t)n_pad16};
        const uint64_t str[1] = {K * 2};
        const uint32_t box[2] = {64, (uint32_t)(p.np < 256 ? p.np : 256)};
        int rc = make_tmap_bf16(&tmB, pw_weight, 2, dims, str, box, nullptr);
        if (rc) return rc;
    }
    {
        const uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)p.B};
        const uint64_t str[3] = {(uint64_t)d->Cout * 2, (uint64_t)p.Wo * d->Cout * 2, (uint64_t)p.Ho * p.Wo * d->Cout * 2};
        const uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
        int rc = make_tmap_bf16(&tmO, out, 4, dims, str, box, nullptr);
        if (rc) return rc;
    }
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        TDRN_CUDA(cudaGetDevice(&dev));
        TDRN_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int total = p.m_tiles * p.n_pass;
    TDRN_CUDA(cudaFuncSetAttribute(conv_dwpw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DP_SMEM));
    conv_dwpw_kernel<<<total < num_sms ? total : num_sms, DP_THREADS, DP_SMEM, as_stream(stream)>>>(tmX, tmB, tmO, p);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}
