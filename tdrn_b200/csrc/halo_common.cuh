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
    int tma_out;                 // epilogue stages bf16 tiles in shared memory and stores them with TMA (plain NHWC bf16 output)
    long long *dbg;              // TDRN_HALO_TIMING=1: per CTA {cycles total, waiting for t_empty, waiting for a_full, issuing, tiles}
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

// bias / ReLU / (pool) / store of one 32-column chunk held in v[]
__device__ __forceinline__ void halo_epilogue_chunk(const HaloP &p, float (&v)[32], const float *s_bias, int n0, int c0,
                                                    int b, int x, int y, int wl, int hl, bool valid);

// Epilogue of one 128-pixel tile: accumulator columns [0, ncols) at `trow` are output channels n0 .. n0+ncols-1.
// `half` selects which of the two warps of this TMEM lane quadrant handles a 32-column chunk.  A warp's chunks (one at
// ncols = 64, two at 128) are LOADED TOGETHER before the single tcgen05.wait::ld: next to running MMAs a tcgen05.ld takes
// several hundred cycles (ncu r02n, conv2_1: LDTM + the first use of its result = 26 % of all stall samples, the epilogue
// took longer than the tile's MMAs), and back-to-back load -> wait -> process pairs paid that latency twice per tile.
__device__ __forceinline__ void halo_epilogue_tile(const HaloP &p, uint32_t trow, const float *s_bias, int n0, int ncols, int half,
                                           int b, int x, int y, int wl, int hl)
{
    const bool valid = x < p.W && y < p.H;
    if (p.debug_skip_epilogue == 1) return;
    const int ca = half * 32, cb2 = half * 32 + 64;
    if (cb2 < ncols) {
        uint32_t ra[32], rb[32];
        const long long c0_ = p.dbg ? clock64() : 0;
        tmem_ld32_issue(trow + (uint32_t)ca, ra);
        tmem_ld32_issue(trow + (uint32_t)cb2, rb);
        tmem_ld_wait32(ra);
        tmem_ld_pin32(rb);
        if (p.dbg && (threadIdx.x & 31) == 0) atomicAdd((unsigned long long *)&p.dbg[148 * 5 + (threadIdx.x >> 5)], (unsigned long long)(clock64() - c0_));
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(ra[i]);
        if (n0 + ca < p.Cout) halo_epilogue_chunk(p, v, s_bias, n0, ca, b, x, y, wl, hl, valid);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rb[i]);
        if (n0 + cb2 < p.Cout) halo_epilogue_chunk(p, v, s_bias, n0, cb2, b, x, y, wl, hl, valid);
    } else if (ca < ncols) {
        float v[32];
        tmem_ld32(trow + (uint32_t)ca, v);
        if (n0 + ca < p.Cout) halo_epilogue_chunk(p, v, s_bias, n0, ca, b, x, y, wl, hl, valid);
    }
}

__device__ __forceinline__ void halo_epilogue_chunk(const HaloP &p, float (&v)[32], const float *s_bias, int n0, int c0,
                                                    int b, int x, int y, int wl, int hl, bool valid)
{
    {
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
            if (!store) return;
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
            if (!store) return;
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

// ---------------------------------------------------------------------------------------------------------
// TMA-store epilogue of one 128-pixel tile (plain NHWC bf16 output, optional fused 2x2 max-pool).
// Measured r02 (TDRN_HALO_TIMING, conv2_1): the register-store epilogue took ~3 900 cycles per tile against 2 300 for the
// tile's MMAs, and not because of tcgen05.ld (40 cycles): a lane stores 16-byte pieces of ITS pixel, so every store
// instruction touches 32 different 128-byte lines and the L1 pipeline spends ~32 cycles on it (64 such instructions per tile).
// Here the eight epilogue warps convert their 32 rows x 32 columns of a 64-channel group into a 128B-swizzled
// [rows][64 ch] staging box (16-byte shared stores, conflict-free) and one thread hands the box to TMA: whole lines leave the SM.
//   stage: two staging boxes `stage_stride` bytes apart (128 rows x 128 B = 16 KB; 4 KB is enough for pooled tiles), used
//   alternately; `git` counts groups over the whole kernel.
//   leader: the one thread that issues / waits for the bulk stores.  Barrier id 1 is shared by the 256 epilogue threads.
// Pooled tiles: the 8 x 16 pixel tile becomes 4 x 8 = 32 rows; the caller's tensor map has box (64, 4, 8, 1).
// `after_last_ld` is called by every warp right after its last tcgen05.ld of the tile (hands the accumulator back).
// ---------------------------------------------------------------------------------------------------------
template <typename F>
__device__ __forceinline__ void halo_epilogue_tile_tma(const HaloP &p, const CUtensorMap *tmO, uint8_t *stage, uint32_t stage_stride,
                                                       uint32_t &git, bool leader,
                                                       uint32_t trow, const float *s_bias, int n0, int ncols, int quad, int half, int lane,
                                                       int b, int x0, int y0, F after_last_ld)
{
    const int r = quad * 32 + lane;
    const int wl = r & (HL_BW - 1), hl = r >> 3;
    const int groups = (ncols + 63) >> 6;
    for (int g = 0; g < groups; ++g, ++git) {
        uint8_t *o = stage + (git & 1u) * stage_stride;
        if (leader) bulk_wait_read<1>();                         // the store that last read this box has drained
        named_bar(1, 256);
        const int c0 = g * 64 + half * 32;
        if (c0 < ncols) {                                        // warp-uniform
            float v[32];
            tmem_ld32(trow + (uint32_t)c0, v);
            if (g == groups - 1) after_last_ld();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 bq = *(const float4 *)(s_bias + n0 + c0 + j);
                v[j] += bq.x; v[j + 1] += bq.y; v[j + 2] += bq.z; v[j + 3] += bq.w;
            }
            if (p.relu) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            uint32_t q[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                q[j] = *(const uint32_t *)&h2;
            }
            int row = r;
            bool wr = true;
            if (p.pool) {                                        // MaxPool2d(2,2): partners are lanes ^1 and ^8 (bf16 max commutes with rounding)
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    uint32_t o1 = __shfl_xor_sync(0xffffffffu, q[j], 1);
                    __nv_bfloat162 m = __hmax2(*(const __nv_bfloat162 *)&q[j], *(const __nv_bfloat162 *)&o1);
                    uint32_t mw = *(const uint32_t *)&m;
                    uint32_t o2 = __shfl_xor_sync(0xffffffffu, mw, HL_BW);
                    m = __hmax2(m, *(const __nv_bfloat162 *)&o2);
                    q[j] = *(const uint32_t *)&m;
                }
                wr = !(wl & 1) && !(hl & 1);
                row = (hl >> 1) * (HL_BW / 2) + (wl >> 1);
            }
            if (wr) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    *(uint4 *)(o + sw128_offset(row, half * 4 + k)) = make_uint4(q[4 * k], q[4 * k + 1], q[4 * k + 2], q[4 * k + 3]);
            }
        } else if (g == groups - 1) {
            after_last_ld();
        }
        fence_proxy_async_smem();
        named_bar(1, 256);
        if (leader) {
            tma_store_4d(tmO, o, n0 + g * 64, p.pool ? x0 >> 1 : x0, p.pool ? y0 >> 1 : y0, b);
            bulk_commit();
        }
    }
}

}  // namespace tc
}  // namespace tdrn
