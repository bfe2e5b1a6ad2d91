// halo_common.cuh -- pieces shared by the halo-tile convolution kernels (conv_halo_tc.cu, conv_stem_pair.cu): tile
// geometry, the kernel parameter block, the shifted / strided UMMA descriptor and the per-tile epilogue.
#pragma once
#include "tc_common.cuh"

namespace tdrn {
namespace tc {

constexpr int HL_BW = 8, HL_BH = 16;                 // output tile
constexpr int HL_PW = HL_BW + 2, HL_PH = HL_BH + 2;  // halo tile (pixels)
constexpr int HL_A_BYTES = HL_PW * HL_PH * 128;      // 23040
constexpr int HL_A_STRIDE = (HL_A_BYTES + 1023) & ~1023;
constexpr int HL_THREADS = 320;                     // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int HL_MAX_STAGES = 4;

struct HaloP {
    int H, W, B;
    int tiles_w, tiles_h, total;
    int Cin, cblocks, Cout, n_pad16;
    int stages;
    uint32_t w_bytes;            // resident weight bytes = 9 * cblocks * n_pad16 * 128
    const float *bias; void *out;
    long long out_sb, out_sp; int out_w;
    int relu, out_f32, pool;
    int debug_skip_epilogue;     // timing experiments only (TDRN_HALO_DEBUG=1): results are NOT written
};

// SWIZZLE_128B K-major descriptor with an arbitrary (16-byte aligned) start and stride between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc_sw128_sbo(uint32_t smem_addr, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// Epilogue of one 128-pixel tile: accumulator columns [0, ncols) at `trow` are output channels n0 .. n0+ncols-1.
// `half` selects which of the two warps of this TMEM lane quadrant handles a 32-column chunk.
__device__ __forceinline__ void halo_epilogue_tile(const HaloP &p, uint32_t trow, const float *s_bias, int n0, int ncols, int half,
                                           int b, int x, int y, int wl, int hl)
{
    const bool valid = x < p.W && y < p.H;
    for (int c0 = half * 32; c0 < (p.debug_skip_epilogue == 1 ? 0 : ncols); c0 += 64) {
        float v[32];
        tmem_ld32(trow + (uint32_t)c0, v);
        if (n0 + c0 >= p.Cout) continue;                     // warp-uniform
        const int nv = min(32, p.Cout - n0 - c0);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 bq = *(const float4 *)(s_bias + n0 + c0 + j);
            v[j] += bq.x; v[j + 1] += bq.y; v[j + 2] += bq.z; v[j + 3] += bq.w;
        }
        if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        bool store = valid;
        int oy = y, ox = x;
        if (p.pool) { store = valid && !(wl & 1) && !(hl & 1); oy = y >> 1; ox = x >> 1; }
        const long long o = (long long)b * p.out_sb + ((long long)oy * p.out_w + ox) * p.out_sp + n0 + c0;
        if (p.out_f32) {
            if (p.pool) {
                // MaxPool2d(2,2): the 2x2 partners of pixel (hl, wl) are lanes ^1 and ^8 of this warp
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float m = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
                    v[j] = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, HL_BW));
                }
            }
            if (!store) continue;
            float *op = (float *)p.out + o;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < nv) op[j] = v[j];
        } else {
            uint32_t q[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                q[j] = *(const uint32_t *)&h2;
            }
            if (p.pool) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    uint32_t o1 = __shfl_xor_sync(0xffffffffu, q[j], 1);
                    __nv_bfloat162 m = __hmax2(*(const __nv_bfloat162 *)&q[j], *(const __nv_bfloat162 *)&o1);
                    uint32_t mw = *(const uint32_t *)&m;
                    uint32_t o2 = __shfl_xor_sync(0xffffffffu, mw, HL_BW);
                    m = __hmax2(m, *(const __nv_bfloat162 *)&o2);
                    q[j] = *(const uint32_t *)&m;
                }
            }
            if (!store) continue;
            __nv_bfloat16 *op = (__nv_bfloat16 *)p.out + o;
            if (nv == 32 && ((o & 7) == 0)) {
#pragma unroll
                for (int j = 0; j < 4; ++j) ((uint4 *)op)[j] = make_uint4(q[4 * j], q[4 * j + 1], q[4 * j + 2], q[4 * j + 3]);
            } else {
                const __nv_bfloat16 *qb = (const __nv_bfloat16 *)q;
#pragma unroll
                for (int j = 0; j < 32; ++j) if (j < nv) op[j] = qb[j];
            }
        }
    }
}

}  // namespace tc
}  // namespace tdrn
