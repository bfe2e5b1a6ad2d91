// conv_halo_tc.cu -- 3x3 / stride 1 / pad 1 implicit-GEMM convolution for the high-resolution, narrow layers
// (VGG conv1_2 64->64 @320^2, conv2_1 64->128 @160^2; model/networks.py:136-163) on tcgen05, with
//   * ONE shared-memory halo tile per (output tile, 64-channel block): a 4-D TMA box of 10 x 18 input pixels
//     (8 x 16 outputs + the 3x3 halo, zero padding = TMA out-of-bounds fill).  The nine taps are nine UMMA
//     A-descriptors into that same tile: start address shifted by (r*10 + s) 128-byte rows, stride between
//     8-row core-matrix groups (SBO) = 10 rows = 1280 B.  The 128B swizzle is a function of the absolute
//     shared-memory address on this hardware (scripts/umma_probe.py), so shifted / 1280-strided views read
//     exactly what TMA wrote.  Compared with one box per tap (conv_tc.cu) the A operand crosses L2->SM once
//     instead of nine times;
//   * the whole weight tensor RESIDENT in shared memory (9 taps x Cin/64 blocks x Cout rows x 128 B, loaded
//     once per persistent CTA), so the steady-state L2->SM traffic is 23 KB per 128-pixel tile.
// conv_tc.cu's generic kernel moved 9 x (16 KB + Cout x 128 B) per tile and was L2-bandwidth bound on these
// layers (385 TFLOP/s on conv1_2); here the tensor pipe / epilogue are the limit.
// Warp roles: warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2-9 epilogue (two per TMEM lane
// quadrant; bias, ReLU, optional fused 2x2 max-pool by warp shuffles, bf16/fp32 store); TMEM accumulator double-buffered.
#include "halo_common.cuh"
#include <stdlib.h>
#include <stdio.h>

// Compile-time experiments for the main loops (all OFF in the shipped build: the code below is then removed by the
// preprocessor; build with e.g. TDRN_NVCC_EXTRA="-DTDRN_HALO_PREFETCH=6" python -m tdrn_b200.build --force):
//   TDRN_HALO_PREFETCH=k   the A producer asks L2 for the halo box of the tile k grid-strides ahead (no correctness effect):
//                          tests whether the issuer waits for HBM latency with only 2-4 stages in flight;
//   TDRN_HALO_NO_A_LOADS   the A producer signals `a_full` without loading anything (TIMING ONLY, results are garbage):
//                          what the loop costs when the operand is always there;
//   TDRN_HALO_ALIGNED_A    every tap reads the nearest 1024-byte-aligned view with SBO 1024 (TIMING ONLY): do the shifted /
//                          1280-byte-strided views of the halo tile read slower than aligned ones?
// DESIGN.md section 4 ("What one narrow MMA really costs") says why these three.

namespace tdrn {
namespace tc {

template <int BN>
__global__ void __launch_bounds__(HL_THREADS, 1) conv_halo_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmB,
                                                                  const __grid_constant__ CUtensorMap tmO, const HaloP p)
{
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t a_full[HL_MAX_STAGES], a_empty[HL_MAX_STAGES];
    __shared__ __align__(8) uint64_t t_full[2], t_empty[2], w_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float s_bias[BN];

    uint8_t *base = (uint8_t *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    uint8_t *sW = base;                               // [9*cblocks][n_pad16 rows][128 B]
    uint8_t *sA = base + p.w_bytes;                   // stages x HL_A_STRIDE (w_bytes is a multiple of 2048)
    uint8_t *sO = sA + (size_t)p.stages * HL_A_STRIDE;   // tma_out: two 16 KB staging boxes of the TMA-store epilogue
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = p.tiles_w * p.tiles_h;
    const uint32_t kb_bytes = (uint32_t)p.n_pad16 * 128u;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
#pragma unroll
        for (int b = 0; b < 2; ++b) { mbar_init(&t_full[b], 1); mbar_init(&t_empty[b], 8); }
        mbar_init(&w_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, 2 * BN);
    for (int i = threadIdx.x; i < BN; i += HL_THREADS) s_bias[i] = (p.bias && i < p.Cout) ? p.bias[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            mbar_expect_tx(&w_bar, p.w_bytes);
            for (int kb = 0; kb < 9 * p.cblocks; ++kb) tma_load_2d(sW + (size_t)kb * kb_bytes, &tmB, &w_bar, kb * 64, 0);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.total; tile += gridDim.x) {
                const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
                const int y0 = (rem / p.tiles_w) * HL_BH - 1, x0 = (rem % p.tiles_w) * HL_BW - 1;
                for (int cb = 0; cb < p.cblocks; ++cb, ++it) {
                    const uint32_t s = it % (uint32_t)p.stages, ph = (it / (uint32_t)p.stages) & 1u;
                    mbar_wait(&a_empty[s], ph ^ 1u);
#ifdef TDRN_HALO_NO_A_LOADS
                    mbar_arrive(&a_full[s]);
#else
                    mbar_expect_tx(&a_full[s], HL_A_BYTES);
                    tma_load_4d(sA + (size_t)s * HL_A_STRIDE, &tmA, &a_full[s], cb * 64, x0, y0, b);
#endif
#if defined(TDRN_HALO_PREFETCH) && TDRN_HALO_PREFETCH > 0
                    {
                        const int tp = tile + TDRN_HALO_PREFETCH * (int)gridDim.x;
                        if (tp < p.total) {
                            const int bp = tp / tiles_per_img, rp = tp - bp * tiles_per_img;
                            tma_prefetch_l2_4d(&tmA, cb * 64, (rp % p.tiles_w) * HL_BW - 1, (rp / p.tiles_w) * HL_BH - 1, bp);
                        }
                    }
#endif
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_bf16(128, p.n_pad16);
            mbar_wait(&w_bar, 0);
            tc_fence_after();
            const uint32_t sW_u = smem_u32(sW), sA_u = smem_u32(sA);
            uint32_t it = 0, tcount = 0;
            long long c_te = 0, c_af = 0, c_start = p.dbg ? clock64() : 0, c_mark = c_start;    // TDRN_HALO_TIMING buckets
            for (int tile = blockIdx.x; tile < p.total; tile += gridDim.x, ++tcount) {
                const uint32_t buf = tcount & 1u;
                mbar_wait(&t_empty[buf], ((tcount >> 1) & 1u) ^ 1u);
                if (p.dbg) { const long long c = clock64(); c_te += c - c_mark; c_mark = c; }
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * BN;
                for (int cb = 0; cb < p.cblocks; ++cb, ++it) {
                    const uint32_t s = it % (uint32_t)p.stages, ph = (it / (uint32_t)p.stages) & 1u;
                    if (p.dbg) c_mark = clock64();
                    mbar_wait(&a_full[s], ph);
                    if (p.dbg) { const long long c = clock64(); c_af += c - c_mark; c_mark = c; }
                    tc_fence_after();
                    const uint32_t a0 = sA_u + s * (uint32_t)HL_A_STRIDE;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const int tr = tap / 3, ts = tap - tr * 3;
#ifdef TDRN_HALO_ALIGNED_A
                        const uint64_t adesc = umma_desc_sw128_sbo(a0 + ((uint32_t)(tr * HL_PW + ts) & ~7u) * 128u, 1024u);
#else
                        const uint64_t adesc = umma_desc_sw128_sbo(a0 + (uint32_t)(tr * HL_PW + ts) * 128u, HL_PW * 128u);
#endif
                        const uint64_t bdesc = umma_desc_sw128(sW_u + (uint32_t)(tap * p.cblocks + cb) * kb_bytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (cb | tap | k) != 0);
                    }
                    umma_commit(&a_empty[s]);
                }
                umma_commit(&t_full[buf]);
                if (p.dbg) c_mark = clock64();
            }
            if (p.dbg) {
                const long long tot = clock64() - c_start;
                p.dbg[blockIdx.x * 5 + 0] = tot; p.dbg[blockIdx.x * 5 + 1] = c_te; p.dbg[blockIdx.x * 5 + 2] = c_af;
                p.dbg[blockIdx.x * 5 + 3] = tot - c_te - c_af; p.dbg[blockIdx.x * 5 + 4] = tcount;
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (warps 2..9) =====================
        // Two warps per TMEM lane quadrant (hardware rule: a warp reads lanes 32*(warp%4)..+31); they split the
        // accumulator's 32-column chunks between them.  bf16 outputs are packed BEFORE the 2x2 max-pool, so the
        // pool is 2 shuffles + 2 packed max per channel PAIR (rounding is monotonic: max commutes with it).
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int r = quad * 32 + lane;
        const int wl = r & (HL_BW - 1), hl = r >> 3;
        uint32_t tcount = 0, git = 0;
        const bool leader = threadIdx.x == 64;
        if (p.tma_out) {
            if (leader) tma_prefetch_desc(&tmO);
            for (int tile = blockIdx.x; tile < p.total; tile += gridDim.x, ++tcount) {
                const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
                const uint32_t buf = tcount & 1u;
                mbar_wait(&t_full[buf], (tcount >> 1) & 1u);
                tc_fence_after();
                const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * BN;
                halo_epilogue_tile_tma(p, &tmO, sO, 128u * 128u, git, leader, trow, s_bias, 0, p.n_pad16, quad, half, lane, b,
                                       (rem % p.tiles_w) * HL_BW, (rem / p.tiles_w) * HL_BH, [&]() {
                                           tc_fence_before();
                                           __syncwarp();
                                           if (lane == 0) mbar_arrive(&t_empty[buf]);
                                       });
            }
            if (leader) bulk_wait_read<0>();
        } else
        for (int tile = blockIdx.x; tile < p.total; tile += gridDim.x, ++tcount) {
            const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
            const int x = (rem % p.tiles_w) * HL_BW + wl, y = (rem / p.tiles_w) * HL_BH + hl;
            const uint32_t buf = tcount & 1u;
            const long long e0 = p.dbg ? clock64() : 0;
            mbar_wait(&t_full[buf], (tcount >> 1) & 1u);
            const long long e1 = p.dbg ? clock64() : 0;
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * BN;
            halo_epilogue_tile(p, trow, s_bias, 0, p.n_pad16, half, b, x, y, wl, hl);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[buf]);
            if (p.dbg && lane == 0) {
                atomicAdd((unsigned long long *)&p.dbg[148 * 5 + 16 + warp], (unsigned long long)(e1 - e0));             // waiting for the tile
                atomicAdd((unsigned long long *)&p.dbg[148 * 5 + 32 + warp], (unsigned long long)(clock64() - e1));       // working on it
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 2 * BN); }
}

// ---------------------------------------------------------------------------------------------------------
// Streamed-weight variant for layers whose weights do not fit in shared memory (conv2_2 128->128 @160^2,
// conv3_1 128->256 @80^2).  A work unit is a 16 x 16 pixel super-tile (two 8 x 16 M tiles side by side, ONE 18 x 18
// halo box, tile 1's descriptors start 8 rows further) times one 128-wide N tile: each weight k-block (16 KB)
// fetched from L2 feeds 2 x 4 MMAs, so the L2->SM traffic per MMA is 2.5x lower than conv_tc.cu's and the main
// loop is no longer L2 bound (measured r01f: ~98 cycles per M=128, N=128, K=16 instruction against a hardware cost of 64;
// DESIGN.md section 4 lists what is suspected).
// TMEM: 2 tiles x 128 columns, double buffered = all 512 columns.
// Warps: 0 A-halo TMA, 1 MMA, 2-9 epilogue, 10 weight TMA.
// ---------------------------------------------------------------------------------------------------------
constexpr int HS_PW = 2 * HL_BW + 2;                      // 18
constexpr int HS_A_BYTES = HS_PW * HL_PH * 128;           // 41472
constexpr int HS_A_STRIDE = (HS_A_BYTES + 1023) & ~1023;  // 41984
constexpr int HS_A_STAGES = 2;
constexpr int HS_BN = 128;
constexpr int HS_B_BYTES = HS_BN * 128;                   // 16 KB per (tap, channel block)
constexpr int HS_B_STAGES = 6;                           // (7 before the TMA-store epilogue took 32 KB for its staging boxes)
constexpr int HS_THREADS = 352;
constexpr int HS_SMEM = 1024 + HS_A_STAGES * HS_A_STRIDE + HS_B_STAGES * HS_B_BYTES + 2 * 128 * 128;

__global__ void __launch_bounds__(HS_THREADS, 1) conv_halo_stream_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                         const __grid_constant__ CUtensorMap tmB,
                                                                         const __grid_constant__ CUtensorMap tmO, const HaloP p)
{
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t a_full[HS_A_STAGES], a_empty[HS_A_STAGES];
    __shared__ __align__(8) uint64_t b_full[HS_B_STAGES], b_empty[HS_B_STAGES];
    __shared__ __align__(8) uint64_t t_full[2], t_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float s_bias[512];

    uint8_t *base = (uint8_t *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = base;
    uint8_t *sB = base + HS_A_STAGES * HS_A_STRIDE;
    uint8_t *sO = sB + HS_B_STAGES * HS_B_BYTES;                  // two 16 KB staging boxes of the TMA-store epilogue
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pairs_w = p.tiles_w >> 1;                           // super-tiles per row
    const int units_per_img = pairs_w * p.tiles_h;
    const int n_tiles = p.n_pad16 / HS_BN;
    const int m_units = units_per_img * p.B;
    const int total = m_units * n_tiles;                          // unit = (n tile, super-tile); n-major so that
                                                                  // concurrently running CTAs share the weight slice

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
#pragma unroll
        for (int s = 0; s < HS_A_STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
#pragma unroll
        for (int s = 0; s < HS_B_STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
#pragma unroll
        for (int b = 0; b < 2; ++b) { mbar_init(&t_full[b], 1); mbar_init(&t_empty[b], 8); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, 512);
    for (int i = threadIdx.x; i < 512; i += HS_THREADS) s_bias[i] = (p.bias && i < p.Cout) ? p.bias[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===================== A (halo) TMA producer =====================
        if (elect_one()) {
            uint32_t it = 0;
            for (int unit = blockIdx.x; unit < total; unit += gridDim.x) {
                const int mu = unit % m_units;
                const int b = mu / units_per_img, rem = mu - b * units_per_img;
                const int y0 = (rem / pairs_w) * HL_BH - 1, x0 = (rem % pairs_w) * (2 * HL_BW) - 1;
                for (int cb = 0; cb < p.cblocks; ++cb, ++it) {
                    const uint32_t s = it % HS_A_STAGES, ph = (it / HS_A_STAGES) & 1u;
                    mbar_wait(&a_empty[s], ph ^ 1u);
#ifdef TDRN_HALO_NO_A_LOADS
                    mbar_arrive(&a_full[s]);
#else
                    mbar_expect_tx(&a_full[s], HS_A_BYTES);
                    tma_load_4d(sA + (size_t)s * HS_A_STRIDE, &tmA, &a_full[s], cb * 64, x0, y0, b);
#endif
#if defined(TDRN_HALO_PREFETCH) && TDRN_HALO_PREFETCH > 0
                    if (cb == 0) {                                       // every channel block of the unit TDRN_HALO_PREFETCH strides ahead
                        const int up = unit + TDRN_HALO_PREFETCH * (int)gridDim.x;
                        if (up < total) {
                            const int mp = up % m_units, bp = mp / units_per_img, rp = mp - bp * units_per_img;
                            for (int c = 0; c < p.cblocks; ++c)
                                tma_prefetch_l2_4d(&tmA, c * 64, (rp % pairs_w) * (2 * HL_BW) - 1, (rp / pairs_w) * HL_BH - 1, bp);
                        }
                    }
#endif
                }
            }
        }
        __syncwarp();
    } else if (warp == 10) {
        // ===================== B (weight k-block) TMA producer =====================
        if (elect_one()) {
            uint32_t jt = 0;
            for (int unit = blockIdx.x; unit < total; unit += gridDim.x) {
                const int n0 = (unit / m_units) * HS_BN;
                for (int cb = 0; cb < p.cblocks; ++cb)
                    for (int tap = 0; tap < 9; ++tap, ++jt) {
                        const uint32_t s = jt % HS_B_STAGES, ph = (jt / HS_B_STAGES) & 1u;
                        mbar_wait(&b_empty[s], ph ^ 1u);
                        if (p.debug_skip_epilogue == 2 && (jt & 1u)) { mbar_arrive(&b_full[s]); continue; }   // timing experiment: half the weight traffic
                        mbar_expect_tx(&b_full[s], HS_B_BYTES);
                        tma_load_2d(sB + (size_t)s * HS_B_BYTES, &tmB, &b_full[s], (tap * p.cblocks + cb) * 64, n0);
                    }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_bf16(128, HS_BN);
            const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
            uint32_t it = 0, jt = 0, tcount = 0;
            for (int unit = blockIdx.x; unit < total; unit += gridDim.x, ++tcount) {
                const uint32_t buf = tcount & 1u;
                mbar_wait(&t_empty[buf], ((tcount >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * (2 * HS_BN);
                for (int cb = 0; cb < p.cblocks; ++cb, ++it) {
                    const uint32_t sa = it % HS_A_STAGES, pha = (it / HS_A_STAGES) & 1u;
                    mbar_wait(&a_full[sa], pha);
                    tc_fence_after();
                    const uint32_t a0 = sA_u + sa * (uint32_t)HS_A_STRIDE;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap, ++jt) {
                        const int tr = tap / 3, ts = tap - tr * 3;
                        const uint32_t sb = jt % HS_B_STAGES, phb = (jt / HS_B_STAGES) & 1u;
                        mbar_wait(&b_full[sb], phb);
                        tc_fence_after();
                        const uint64_t bdesc = umma_desc_sw128(sB_u + sb * (uint32_t)HS_B_BYTES);
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            const uint64_t adesc = umma_desc_sw128_sbo(a0 + (uint32_t)(tr * HS_PW + ts + mt * HL_BW) * 128u, HS_PW * 128u);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16(d_tmem + mt * HS_BN, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (cb | tap | k) != 0);
                        }
                        umma_commit(&b_empty[sb]);
                    }
                    umma_commit(&a_empty[sa]);
                }
                umma_commit(&t_full[buf]);
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (warps 2..9) =====================
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int r = quad * 32 + lane;
        const int wl = r & (HL_BW - 1), hl = r >> 3;
        uint32_t tcount = 0, git = 0;
        const bool leader = threadIdx.x == 64;
        if (p.tma_out) {
            if (leader) tma_prefetch_desc(&tmO);
            for (int unit = blockIdx.x; unit < total; unit += gridDim.x, ++tcount) {
                const int mu = unit % m_units, n0 = (unit / m_units) * HS_BN;
                const int b = mu / units_per_img, rem = mu - b * units_per_img;
                const int y0 = (rem / pairs_w) * HL_BH, x0 = (rem % pairs_w) * (2 * HL_BW);
                const uint32_t buf = tcount & 1u;
                mbar_wait(&t_full[buf], (tcount >> 1) & 1u);
                tc_fence_after();
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * (2 * HS_BN) + mt * HS_BN;
                    halo_epilogue_tile_tma(p, &tmO, sO, 128u * 128u, git, leader, trow, s_bias, n0, HS_BN, quad, half, lane, b, x0 + mt * HL_BW, y0, [&]() {
                        if (mt == 1) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&t_empty[buf]);
                        }
                    });
                }
            }
            if (leader) bulk_wait_read<0>();
        } else
        for (int unit = blockIdx.x; unit < total; unit += gridDim.x, ++tcount) {
            const int mu = unit % m_units, n0 = (unit / m_units) * HS_BN;
            const int b = mu / units_per_img, rem = mu - b * units_per_img;
            const int y = (rem / pairs_w) * HL_BH + hl, xb = (rem % pairs_w) * (2 * HL_BW) + wl;
            const uint32_t buf = tcount & 1u;
            mbar_wait(&t_full[buf], (tcount >> 1) & 1u);
            tc_fence_after();
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * (2 * HS_BN) + mt * HS_BN;
                halo_epilogue_tile(p, trow, s_bias, n0, HS_BN, half, b, xb + mt * HL_BW, y, wl, hl);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[buf]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

static int g_halo_sms = 0;

template <int BN>
static int launch_halo(const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmO, const HaloP &p, size_t smem, cudaStream_t st)
{
    TDRN_CUDA(cudaFuncSetAttribute(conv_halo_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_halo_kernel<BN><<<p.total < g_halo_sms ? p.total : g_halo_sms, HL_THREADS, smem, st>>>(tmA, tmB, tmO, p);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

// Returns TDRN_EUNSUPPORTED when the layer does not fit this kernel (the caller then uses the generic conv_tc path).
int conv_halo_try(const tdrn_conv_desc *d, const void *in, const void *weight, const float *bias, const void *residual,
                  void *out, cudaStream_t st)
{
    if (d->deconv2x2 || d->kh != 3 || d->kw != 3 || d->stride != 1 || d->pad != 1 || d->dil != 1 || residual ||
        d->Cin % 64 != 0 || d->in_sb != 0)
        return TDRN_EUNSUPPORTED;
    // Exact tiling is needed by the fused pool and by the streamed variant; the resident-weight kernel also takes
    // ragged maps (partial tiles: TMA zero-fills beyond the map, the epilogue masks), which is what makes the
    // Cout = 12 ARM heads on 40x40 / 20x20 maps cheap: their cost is A traffic, cut ~6x by the halo tile.
    const bool exact = d->W % HL_BW == 0 && d->H % HL_BH == 0;
    if (!exact && (d->pool2x2 || d->Cout > 32)) return TDRN_EUNSUPPORTED;
    HaloP p{};
    p.H = d->H; p.W = d->W; p.B = d->B; p.Cin = d->Cin; p.cblocks = d->Cin / 64; p.Cout = d->Cout;
    p.n_pad16 = (d->Cout + 15) & ~15;
    p.w_bytes = 9u * (uint32_t)p.cblocks * (uint32_t)p.n_pad16 * 128u;
    // TMA-store epilogue: plain contiguous NHWC bf16 output (pooled or not) on exactly tiled maps; costs 32 KB of staging boxes
    static const bool no_tma_out = getenv("TDRN_NO_TMA_STORE") != nullptr;
    const int Ho = d->pool2x2 ? d->H / 2 : d->H, Wo = d->pool2x2 ? d->W / 2 : d->W;
    const bool tma_out = !no_tma_out && exact && d->out_dtype == TDRN_BF16 && d->Cout % 8 == 0 && d->out_sp == d->Cout &&
                         d->out_sb == (long long)Ho * Wo * d->Cout && ((uintptr_t)out & 15) == 0;
    const size_t stage_bytes = tma_out ? 2 * 128 * 128 : 0;
    const size_t budget = 225 * 1024;
    int stages = HL_MAX_STAGES;
    while (stages >= 2 && 1024 + (size_t)p.w_bytes + (size_t)stages * HL_A_STRIDE + stage_bytes > budget) --stages;
    const bool resident = stages >= 2 && d->Cout <= 128;
    // weights too large to stay resident: streamed variant (two M tiles per weight k-block), 128-wide N tiles
    const bool streamed = !resident && exact && d->Cout % HS_BN == 0 && d->Cout <= 512 && d->W % (2 * HL_BW) == 0 && !getenv("TDRN_NO_HALO_STREAM");
    if (!resident && !streamed) return TDRN_EUNSUPPORTED;
    p.stages = stages;
    p.tiles_w = (d->W + HL_BW - 1) / HL_BW; p.tiles_h = (d->H + HL_BH - 1) / HL_BH; p.total = p.tiles_w * p.tiles_h * d->B;
    p.bias = bias; p.out = out; p.out_sb = d->out_sb; p.out_sp = d->out_sp;
    p.relu = d->relu; p.out_f32 = d->out_dtype == TDRN_F32; p.pool = d->pool2x2;
    p.out_w = p.pool ? d->W / 2 : d->W;
    { static const int dbg = getenv("TDRN_HALO_DEBUG") ? atoi(getenv("TDRN_HALO_DEBUG")) : 0; p.debug_skip_epilogue = dbg; }
    if (!g_halo_sms) {
        int dev = 0;
        TDRN_CUDA(cudaGetDevice(&dev));
        TDRN_CUDA(cudaDeviceGetAttribute(&g_halo_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    CUtensorMap tmA, tmB;
    {
        const uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
        const uint64_t str[3] = {(uint64_t)d->Cin * 2, (uint64_t)d->W * d->Cin * 2, (uint64_t)d->H * d->W * d->Cin * 2};
        const uint32_t box[4] = {64, (uint32_t)(resident ? HL_PW : HS_PW), HL_PH, 1};
        int rc = make_tmap_bf16(&tmA, in, 4, dims, str, box, nullptr);
        if (rc) return rc;
    }
    {
        const uint64_t K = 9ull * d->Cin;
        const uint64_t dims[2] = {K, (uint64_t)p.n_pad16};
        const uint64_t str[1] = {K * 2};
        const uint32_t box[2] = {64, (uint32_t)(resident ? p.n_pad16 : HS_BN)};
        int rc = make_tmap_bf16(&tmB, weight, 2, dims, str, box, nullptr);
        if (rc) return rc;
    }
    p.tma_out = tma_out;
    CUtensorMap tmO = tmA;
    if (tma_out) {
        const uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)d->B};
        const uint64_t str[3] = {(uint64_t)d->Cout * 2, (uint64_t)Wo * d->Cout * 2, (uint64_t)Ho * Wo * d->Cout * 2};
        const uint32_t box[4] = {64, (uint32_t)(d->pool2x2 ? HL_BW / 2 : HL_BW), (uint32_t)(d->pool2x2 ? HL_BH / 2 : HL_BH), 1};
        int rc = make_tmap_bf16(&tmO, out, 4, dims, str, box, nullptr);
        if (rc) return rc;
    }
    if (!resident) {
        const int total = (p.tiles_w / 2) * p.tiles_h * d->B * (p.n_pad16 / HS_BN);
        TDRN_CUDA(cudaFuncSetAttribute(conv_halo_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HS_SMEM));
        conv_halo_stream_kernel<<<total < g_halo_sms ? total : g_halo_sms, HS_THREADS, HS_SMEM, st>>>(tmA, tmB, tmO, p);
        TDRN_LAUNCH_CHECK();
        return TDRN_OK;
    }
    const size_t smem = 1024 + (size_t)p.w_bytes + (size_t)stages * HL_A_STRIDE + stage_bytes;
    static const bool timing = getenv("TDRN_HALO_TIMING") != nullptr;        // development aid: where does the MMA issuer's time go?
    if (timing) {
        static long long *dbg = nullptr;
        if (!dbg) TDRN_CUDA(cudaMalloc(&dbg, (148 * 5 + 48) * sizeof(long long)));
        TDRN_CUDA(cudaMemsetAsync(dbg, 0, (148 * 5 + 48) * sizeof(long long), st));
        p.dbg = dbg;
        int rc = p.n_pad16 > 64 ? launch_halo<128>(tmA, tmB, tmO, p, smem, st) : launch_halo<64>(tmA, tmB, tmO, p, smem, st);
        long long h[148 * 5 + 48];
        TDRN_CUDA(cudaMemcpyAsync(h, dbg, sizeof(h), cudaMemcpyDeviceToHost, st));
        TDRN_CUDA(cudaStreamSynchronize(st));
        double t[5] = {0, 0, 0, 0, 0};
        for (int i = 0; i < 148; ++i) for (int j = 0; j < 5; ++j) t[j] += (double)h[i * 5 + j] / 148.0;
        fprintf(stderr, "halo timing Cin %d Cout %d @%dx%d: MMA-issuer cycles per tile: total %.0f = waiting for the epilogue (t_empty) %.0f + "
                        "waiting for the halo box (a_full) %.0f + issuing %.0f  (%.0f tiles per CTA, %d MMAs per tile)\n", p.Cin, p.Cout, p.H, p.W,
                t[0] / t[4], t[1] / t[4], t[2] / t[4], t[3] / t[4], t[4], 36 * p.cblocks);
        const double tiles_all = t[4] * 148.0;
        fprintf(stderr, "   epilogue warp 2, cycles per tile: waiting for the tile %.0f, working on it %.0f, of which tcgen05.ld x2 + wait %.0f\n",
                (double)h[148 * 5 + 16 + 2] / tiles_all, (double)h[148 * 5 + 32 + 2] / tiles_all, (double)h[148 * 5 + 2] / tiles_all);
        return rc;
    }
    return p.n_pad16 > 64 ? launch_halo<128>(tmA, tmB, tmO, p, smem, st) : launch_halo<64>(tmA, tmB, tmO, p, smem, st);
}

}  // namespace tc
}  // namespace tdrn
