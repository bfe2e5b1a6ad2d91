// conv_stem_pair.cu -- VGG conv1_1 (3 -> 64) and conv1_2 (64 -> 64, + optional MaxPool2d(2,2)) in ONE kernel
// (model/networks.py:136-163, first three entries of the VGG cfg; folded BN + ReLU after each conv).
//
// conv1_1 writes the largest activation of the network (B x 320 x 320 x 64 bf16 = 419 MB at b32) only for conv1_2 to
// read it back: the stem kernel (conv_stem_tc.cu) is bound by exactly that HBM write.  Here the conv1_1 outputs a conv1_2
// tile needs -- the 10 x 18 halo of an 8 x 16 output tile -- are computed on the spot and never leave the SM:
//   patch   : fp32 NCHW image box (20 x 20 x 3, the reference's own input layout) through a 4-deep TMA ring; the conv1_1
//             zero padding is the TMA out-of-bounds fill;
//   A'      : 4 warps build the 27-tap im2col rows of the 180 halo pixels (K = 32, 128B-swizzled K-major, two M = 128 halves);
//   pre-MMA : D' [256 x 64] = A' x W1^T on tcgen05 (4 instructions), accumulators in TMEM;
//   mid     : 4 more warps read D', add the folded-BN bias, ReLU, ZERO the halo pixels that lie outside the image
//             (conv1_2's zero padding applies to conv1_1's OUTPUT map), cast to bf16 and store them as the 10 x 18 x 64
//             halo tile in exactly the layout a TMA box of the NHWC tensor would have (128-byte rows, absolute-address
//             128B swizzle), so
//   main    : the nine taps of conv1_2 are the nine shifted / 1280-byte-strided UMMA descriptors of conv_halo_tc.cu,
//             weights resident in shared memory, accumulator double-buffered in TMEM;
//   epilogue: 8 warps, bias + ReLU + fused 2x2 max-pool, bf16 NHWC stores (halo_epilogue_tile).
// The pre-MMA of tile i+1 is issued before the main MMAs of tile i, so the builder / mid warps work one tile ahead of the
// tensor pipe.  Arithmetic is identical to the two-kernel path (same bf16 rounding of conv1_1's output, same K order), so
// the results are bit-identical to conv_stem_tc + conv_halo_kernel; tests assert exactly that.
//
// Warps: 0 patch TMA + weight TMA, 1 TMEM allocator + MMA issuer, 2-9 epilogue, 10-13 im2col builders, 14-17 mid stage.
//
// Measured (B200, b32, 320x320): 0.410 ms against 0.118 (conv_stem_tc) + 0.308 (conv_halo_kernel) = 0.426 ms for the two
// kernels.  The limiter is shared-memory bandwidth: an M=128, N=64, K=16 MMA reads 6 KB of operands (48 cycles at the 128 B/clk the
// tensor core gets from shared memory, r01f probe; "73 cycles" in earlier notes was a probe artefact), and the LSU traffic of the builder / mid / epilogue warps competes for the rest (ncu: 45 % LSU-shared
// wavefront utilisation on top of the tensor core's reads at 0.431 ms; loading the conv1_1 bias as float4 instead of scalars
// and a patch pitch of 20 floats -- no 2-way bank conflicts in the im2col reads -- brought 0.431 -> 0.410).  It also removes
// 838 MB of HBM traffic per step (40 % of the step's total); whole step 2.79-2.82 vs 2.85-2.86 ms on a power-capped
// part (same box, alternating runs, before the shared-memory tweaks).  TDRN_NO_STEM_PAIR=1 selects the two-kernel path.
#include "halo_common.cuh"
#include <stdlib.h>
#include <stdio.h>

namespace tdrn {
namespace tc {

constexpr int SP_THREADS = 704;                            // 22 warps: TMA, MMA, 8 epilogue, 4 im2col builders, 8 mid-stage
constexpr int SP_C = 64;                                   // channels of conv1_1's output = conv1_2's input and output
constexpr int SP_ROWS = HL_PW * HL_PH;                     // 180 halo pixels
constexpr int SP_PXW = 20, SP_PXH = HL_PH + 2;             // patch: x0-4 .. x0+15 (16-byte aligned start; 12 columns are needed, a pitch
                                                           // of 20 floats keeps rows r, r+1, r+2 of a warp's LDS on different banks), y0-2 .. y0+17
constexpr int SP_PATCH_BYTES = 3 * SP_PXH * SP_PXW * 4;    // 4800
constexpr int SP_PATCH_STRIDE = 5120;
constexpr int SP_PSTAGES = 4;
constexpr int SP_W2_BYTES = 9 * SP_C * 128;                // 73728: conv1_2 weights, [tap][64 rows][128 B]
constexpr int SP_HSTAGES = 2;                              // halo A tiles
constexpr int SP_AP_BYTES = 2 * 16384;                     // one A' tile: two M = 128 halves of [128 rows][128 B]
constexpr int SP_OUT_BYTES = 2 * 4096;                     // two staging boxes of the TMA-store epilogue (pooled tile: 32 rows x 128 B)
constexpr int SP_SMEM = 1024 + SP_W2_BYTES + SP_HSTAGES * HL_A_STRIDE + 2 * SP_AP_BYTES + 8192 + SP_PSTAGES * SP_PATCH_STRIDE + SP_OUT_BYTES;
constexpr int SP_TMEM_PRE = 128;                           // TMEM: main accumulators [0,128), D' slot s half h at 128 + s*128 + h*64

struct StemPairP {
    HaloP h;                     // conv1_2 / output description (Cin = Cout = 64, cblocks = 1)
    const float *w1;             // [27][64] fp32, k = (i*3+j)*3 + c
    const float *b1;             // [64] or NULL
    int relu1;
};

__device__ __forceinline__ uint32_t sp_pack(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// max(x, 0) on a packed bf16 pair
__device__ __forceinline__ uint32_t sp_relu2(uint32_t a)
{
    uint32_t r;
    asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(0u));
    return r;
}

__global__ void __launch_bounds__(SP_THREADS, 1) conv_stem_pair_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                       const __grid_constant__ CUtensorMap tmW2,
                                                                       const __grid_constant__ CUtensorMap tmO, const StemPairP q)
{
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t p_full[SP_PSTAGES], p_empty[SP_PSTAGES];
    __shared__ __align__(8) uint64_t ap_full[2], ap_empty[2], pre_full[2], pre_empty[2];
    __shared__ __align__(8) uint64_t a_full[SP_HSTAGES], a_empty[SP_HSTAGES], t_full[2], t_empty[2], w_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float s_bias[SP_C];           // conv1_2 bias (halo_epilogue_tile)
    __shared__ __align__(16) float s_bias1[SP_C];

    const HaloP &p = q.h;
    uint8_t *base = (uint8_t *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    uint8_t *sW2 = base;
    uint8_t *sH = sW2 + SP_W2_BYTES;                       // SP_HSTAGES x HL_A_STRIDE (1024-aligned: 73728 = 72 * 1024)
    uint8_t *sAp = sH + SP_HSTAGES * HL_A_STRIDE;          // 2 slots x 2 halves x 16 KB
    uint8_t *sB1 = sAp + 2 * SP_AP_BYTES;                  // [64 rows][128 B]
    uint8_t *sP = sB1 + 8192;                              // patch ring
    uint8_t *sO = sP + SP_PSTAGES * SP_PATCH_STRIDE;       // two 4 KB staging boxes of the TMA-store epilogue (pooled tiles)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_per_img = p.tiles_w * p.tiles_h;
    const int n_my = blockIdx.x < p.total ? (p.total - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;   // tiles of this CTA

    if (tid == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW2);
#pragma unroll
        for (int s = 0; s < SP_PSTAGES; ++s) { mbar_init(&p_full[s], 1); mbar_init(&p_empty[s], 1); }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            mbar_init(&ap_full[s], 1); mbar_init(&ap_empty[s], 1); mbar_init(&pre_full[s], 1); mbar_init(&pre_empty[s], 1);
            mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 8);
        }
        mbar_init(&w_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, 512);
    // conv1_1 weights: w1[k][n] fp32 -> bf16 rows n, logical chunks 0..3 (k 0..31, zero beyond 27)
    for (int e = tid; e < SP_C * 4; e += SP_THREADS) {
        const int n = e >> 2, chunk = e & 3;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int k = chunk * 8 + j; v[j] = k < 27 ? __ldg(q.w1 + k * SP_C + n) : 0.f; }
        *(uint4 *)(sB1 + sw128_offset(n, chunk)) =
            make_uint4(sp_pack(v[0], v[1]), sp_pack(v[2], v[3]), sp_pack(v[4], v[5]), sp_pack(v[6], v[7]));
    }
    if (tid < SP_C) { s_bias[tid] = p.bias ? p.bias[tid] : 0.f; s_bias1[tid] = q.b1 ? q.b1[tid] : 0.f; }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===================== TMA producer: conv1_2 weights once, then one image patch per tile =====================
        if (elect_one()) {
            mbar_expect_tx(&w_bar, SP_W2_BYTES);
            for (int tap = 0; tap < 9; ++tap) tma_load_2d(sW2 + tap * (SP_C * 128), &tmW2, &w_bar, tap * 64, 0);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.total; tile += gridDim.x, ++it) {
                const uint32_t ps = it % SP_PSTAGES, pph = (it / SP_PSTAGES) & 1u;
                const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
                mbar_wait(&p_empty[ps], pph ^ 1u);
                mbar_expect_tx(&p_full[ps], SP_PATCH_BYTES);
                tma_load_4d(sP + ps * SP_PATCH_STRIDE, &tmX, &p_full[ps], (rem % p.tiles_w) * HL_BW - 4,
                            (rem / p.tiles_w) * HL_BH - 2, 0, b);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread): pre(0); { pre(it+1); main(it) } =====================
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_bf16(128, SP_C);
            const uint64_t b1desc = umma_desc_sw128(smem_u32(sB1));
            const uint32_t sW_u = smem_u32(sW2), sH_u = smem_u32(sH), sAp_u = smem_u32(sAp);
            long long c_ap = 0, c_pe = 0, c_te = 0, c_af = 0;            // TDRN_HALO_TIMING: what the issuer waits for
            const long long c_start = p.dbg ? clock64() : 0;
            auto pre = [&](uint32_t it) {
                const uint32_t s = it & 1u, ph = (it >> 1) & 1u;
                long long c0 = p.dbg ? clock64() : 0;
                mbar_wait(&ap_full[s], ph);
                if (p.dbg) { const long long c = clock64(); c_ap += c - c0; c0 = c; }
                mbar_wait(&pre_empty[s], ph ^ 1u);
                if (p.dbg) c_pe += clock64() - c0;
                tc_fence_after();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint64_t adesc = umma_desc_sw128(sAp_u + (s * 2 + h) * 16384u);
                    const uint32_t d = tmem_base + SP_TMEM_PRE + s * 128u + h * 64u;
                    umma_bf16(d, adesc, b1desc, idesc, 0u);
                    umma_bf16(d, adesc + 2, b1desc + 2, idesc, 1u);
                }
                umma_commit(&ap_empty[s]);
                umma_commit(&pre_full[s]);
            };
            mbar_wait(&w_bar, 0);
            tc_fence_after();
            if (n_my > 0) pre(0);
            for (uint32_t it = 0; it < (uint32_t)n_my; ++it) {
                if (it + 1 < (uint32_t)n_my) pre(it + 1);
                const uint32_t s = it & 1u, ph = (it >> 1) & 1u;
                long long c0 = p.dbg ? clock64() : 0;
                mbar_wait(&t_empty[s], ph ^ 1u);
                if (p.dbg) { const long long c = clock64(); c_te += c - c0; c0 = c; }
                mbar_wait(&a_full[s], ph);
                if (p.dbg) c_af += clock64() - c0;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + s * SP_C;
                const uint32_t a0 = sH_u + s * (uint32_t)HL_A_STRIDE;
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int tr = tap / 3, ts = tap - tr * 3;
                    const uint64_t adesc = umma_desc_sw128_sbo(a0 + (uint32_t)(tr * HL_PW + ts) * 128u, HL_PW * 128u);
                    const uint64_t bdesc = umma_desc_sw128(sW_u + (uint32_t)tap * (SP_C * 128u));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (tap | k) != 0);
                }
                umma_commit(&a_empty[s]);
                umma_commit(&t_full[s]);
            }
            if (p.dbg) {
                long long *o = p.dbg + blockIdx.x * 6;
                o[0] = clock64() - c_start; o[1] = c_ap; o[2] = c_pe; o[3] = c_te; o[4] = c_af; o[5] = n_my;
            }
        }
        __syncwarp();
    } else if (warp < 10) {
        // ===================== epilogue (warps 2..9): conv1_2 bias + ReLU + 2x2 max-pool + store =====================
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int r = quad * 32 + lane;
        const int wl = r & (HL_BW - 1), hl = r >> 3;
        uint32_t it = 0, git = 0;
        const bool leader = tid == 64;
        if (p.tma_out) {
            if (leader) tma_prefetch_desc(&tmO);
            for (int tile = blockIdx.x; tile < p.total; tile += gridDim.x, ++it) {
                const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
                const uint32_t s = it & 1u;
                mbar_wait(&t_full[s], (it >> 1) & 1u);
                tc_fence_after();
                const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + s * SP_C;
                halo_epilogue_tile_tma(p, &tmO, sO, 4096u, git, leader, trow, s_bias, 0, SP_C, quad, half, lane, b,
                                       (rem % p.tiles_w) * HL_BW, (rem / p.tiles_w) * HL_BH, [&]() {
                                           tc_fence_before();
                                           __syncwarp();
                                           if (lane == 0) mbar_arrive(&t_empty[s]);
                                       });
            }
            if (leader) bulk_wait_read<0>();
        } else
        for (int tile = blockIdx.x; tile < p.total; tile += gridDim.x, ++it) {
            const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
            const int x = (rem % p.tiles_w) * HL_BW + wl, y = (rem / p.tiles_w) * HL_BH + hl;
            const uint32_t s = it & 1u;
            mbar_wait(&t_full[s], (it >> 1) & 1u);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + s * SP_C;
            halo_epilogue_tile(p, trow, s_bias, 0, SP_C, half, b, x, y, wl, hl);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[s]);
        }
    } else if (warp < 14) {
        // ===================== im2col builders (warps 10..13, 128 threads) =====================
        const int bt = tid - 320;                          // 0..127
        const uint32_t sP_u = smem_u32(sP), sAp_u = smem_u32(sAp);
        auto build = [&](uint32_t it, int tile) {
            const uint32_t s = it & 1u, ps = it % SP_PSTAGES;
            mbar_wait(&p_full[ps], (it / SP_PSTAGES) & 1u);
            mbar_wait(&ap_empty[s], ((it >> 1) & 1u) ^ 1u);
            const uint32_t P_u = sP_u + ps * (uint32_t)SP_PATCH_STRIDE;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int row = h * 128 + bt;              // halo pixel index r * 10 + c
                uint32_t kw[16];
                if (row < SP_ROWS) {
                    const int hr = row / HL_PW, hc = row - hr * HL_PW;
                    const uint32_t p0 = P_u + (uint32_t)((hr * SP_PXW + hc + 2) * 4);
#pragma unroll
                    for (int k2 = 0; k2 < 16; ++k2) {
                        float v[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int k = 2 * k2 + e;                       // compile-time after unrolling
                            if (k < 27) { const int t = k / 3, c = k - t * 3, i = t / 3, j = t - i * 3; v[e] = lds_f32(p0 + (uint32_t)(((c * SP_PXH + i) * SP_PXW + j) * 4)); }
                            else v[e] = 0.f;
                        }
                        kw[k2] = sp_pack(v[0], v[1]);
                    }
                } else {
#pragma unroll
                    for (int k2 = 0; k2 < 16; ++k2) kw[k2] = 0u;
                }
                const uint32_t a = sAp_u + (uint32_t)((s * 2 + h) * 16384);
#pragma unroll
                for (int chunk = 0; chunk < 4; ++chunk)
                    sts128(a + sw128_offset(bt, chunk), kw[4 * chunk], kw[4 * chunk + 1], kw[4 * chunk + 2], kw[4 * chunk + 3]);
            }
            fence_proxy_async_smem();
            named_bar(2, 128);                             // A' complete, patch[ps] fully consumed
            if (bt == 0) { mbar_arrive(&ap_full[s]); mbar_arrive(&p_empty[ps]); }
            (void)tile;
        };
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.total; tile += gridDim.x, ++it) build(it, tile);
    } else {
        // ===================== mid stage (warps 14..21, 256 threads): D' -> bias/ReLU/mask -> bf16 halo tile =====================
        // r02: eight warps instead of four -- two per TMEM lane quadrant, each takes one 32-column half of its rows.  The halo
        // tile this stage produces is what the main MMAs waited for (1 230 of 4 184 cycles per tile, TDRN_HALO_TIMING).
        // r02 (2nd pass): ncu source view -- 42 % of these warps' stall samples sat on the FADDs that consume the bias LDS.128
        // (short scoreboard: next to the tensor core's operand streaming a shared-memory load takes hundreds of cycles) and the
        // stage was never waiting for its input: it was the bottleneck of the kernel.  The bias of a warp's 32 columns now lives
        // in registers, ReLU is one packed bf16x2 max after the cast (max commutes with the monotonic rounding), the halo tile
        // is written with st.shared (the generic ST.E the compiler emitted for these pointers is gone).
        const int bt = tid - 448;                          // 0..255
        const int quad = warp & 3;                         // TMEM lane quadrant this warp may read
        const int chalf = (warp - 14) >> 2;                // which 32-column half
        const int c0 = chalf * 32;
        const uint32_t sH_u = smem_u32(sH);
        float bias_r[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) bias_r[j] = s_bias1[c0 + j];
        auto mid = [&](uint32_t it, int tile) {
            const uint32_t s = it & 1u, ph = (it >> 1) & 1u;
            const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
            const int x0 = (rem % p.tiles_w) * HL_BW - 1, y0 = (rem / p.tiles_w) * HL_BH - 1;
            (void)b;
            mbar_wait(&pre_full[s], ph);
            mbar_wait(&a_empty[s], ph ^ 1u);
            tc_fence_after();
            const uint32_t ht = sH_u + s * (uint32_t)HL_A_STRIDE;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h * 128 + quad * 32 >= SP_ROWS) continue;              // warp-uniform: rows 192.. do not exist
                const int row = h * 128 + quad * 32 + lane;
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + SP_TMEM_PRE + s * 128u + h * 64u;
                const int hr = row / HL_PW, hc = row - hr * HL_PW;
                const bool inside = row < SP_ROWS && (unsigned)(x0 + hc) < (unsigned)p.W && (unsigned)(y0 + hr) < (unsigned)p.H;
                float v[32];
                tmem_ld32(taddr + (uint32_t)c0, v);
                if (row < SP_ROWS) {
#pragma unroll
                    for (int cq = 0; cq < 4; ++cq) {
                        uint32_t w[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int c = cq * 8 + 2 * j;
                            w[j] = sp_pack(v[c] + bias_r[c], v[c + 1] + bias_r[c + 1]);
                            if (q.relu1) w[j] = sp_relu2(w[j]);
                        }
                        const uint32_t dst = ht + sw128_offset(row, (c0 >> 3) + cq);
                        if (inside) sts128(dst, w[0], w[1], w[2], w[3]);
                        else sts128(dst, 0u, 0u, 0u, 0u);                   // conv1_2's zero padding
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();
            named_bar(3, 256);                             // halo tile complete, D' slot s fully read
            if (bt == 0) { mbar_arrive(&a_full[s]); mbar_arrive(&pre_empty[s]); }
        };
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.total; tile += gridDim.x, ++it) mid(it, tile);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace tc
}  // namespace tdrn

using namespace tdrn;
using namespace tdrn::tc;

extern "C" int tdrn_conv_stem_pair(const float *x, const float *w1, const float *b1, const void *w2, const float *b2, void *out,
                                   int B, int H, int W, int relu1, int relu2, int pool, tdrn_stream_t stream)
{
    TDRN_REQUIRE(x && w1 && w2 && out, "tdrn_conv_stem_pair: null argument");
    TDRN_REQUIRE(B > 0 && H > 0 && W > 0, "tdrn_conv_stem_pair: bad shape");
    if (W % HL_BW != 0 || H % HL_BH != 0 || (W * 4) % 16 != 0 || ((uintptr_t)x & 15)) {
        set_error("tdrn_conv_stem_pair: needs W %% 8 == 0, H %% 16 == 0 and a 16-byte aligned image (got %dx%d)", H, W);
        return TDRN_EUNSUPPORTED;
    }
    StemPairP q{};
    HaloP &p = q.h;
    p.H = H; p.W = W; p.B = B; p.Cin = SP_C; p.cblocks = 1; p.Cout = SP_C; p.n_pad16 = SP_C;
    p.w_bytes = SP_W2_BYTES; p.stages = SP_HSTAGES;
    p.tiles_w = W / HL_BW; p.tiles_h = H / HL_BH; p.total = p.tiles_w * p.tiles_h * B;
    p.bias = b2; p.out = out; p.relu = relu2; p.out_f32 = 0; p.pool = pool;
    p.out_w = pool ? W / 2 : W;
    p.out_sp = SP_C; p.out_sb = (long long)(pool ? (H / 2) * (W / 2) : H * W) * SP_C;
    q.w1 = w1; q.b1 = b1; q.relu1 = relu1;
    CUtensorMap tmX, tmW2;
    {   // fp32 NCHW image: dims (W, H, 3, B); box (20, 20, 3, 1) starting at (x0-4, y0-2): out-of-bounds = conv1_1 padding
        EncodeTiledFn enc = get_encode_tiled();
        if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return TDRN_ECUDA; }
        const cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)B};
        const cuuint64_t gstr[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
        const cuuint32_t bdim[4] = {SP_PXW, SP_PXH, 3, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(x), gdim, gstr, bdim, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (stem pair input) failed (CUresult %d)", (int)r); return TDRN_ECUDA; }
    }
    {
        const uint64_t K = 9ull * SP_C;
        const uint64_t dims[2] = {K, (uint64_t)SP_C};
        const uint64_t str[1] = {K * 2};
        const uint32_t box[2] = {64, (uint32_t)SP_C};
        int rc = make_tmap_bf16(&tmW2, w2, 2, dims, str, box, nullptr);
        if (rc) return rc;
    }
    CUtensorMap tmO = tmW2;
    {   // TMA-store epilogue for the pooled output (its staging boxes are 32 rows; the un-pooled form keeps register stores)
        static const bool no_tma_out = getenv("TDRN_NO_TMA_STORE") != nullptr;
        p.tma_out = !no_tma_out && pool && ((uintptr_t)out & 15) == 0;
        if (p.tma_out) {
            const uint64_t dims[4] = {(uint64_t)SP_C, (uint64_t)(W / 2), (uint64_t)(H / 2), (uint64_t)B};
            const uint64_t str[3] = {(uint64_t)SP_C * 2, (uint64_t)(W / 2) * SP_C * 2, (uint64_t)(H / 2) * (W / 2) * SP_C * 2};
            const uint32_t box[4] = {64, HL_BW / 2, HL_BH / 2, 1};
            int rc = make_tmap_bf16(&tmO, out, 4, dims, str, box, nullptr);
            if (rc) return rc;
        }
    }
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        TDRN_CUDA(cudaGetDevice(&dev));
        TDRN_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    TDRN_CUDA(cudaFuncSetAttribute(conv_stem_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SP_SMEM));
    static const bool timing = getenv("TDRN_HALO_TIMING") != nullptr;        // development aid (see conv_halo_tc.cu)
    if (timing) {
        static long long *dbg = nullptr;
        if (!dbg) TDRN_CUDA(cudaMalloc(&dbg, 148 * 6 * sizeof(long long)));
        TDRN_CUDA(cudaMemsetAsync(dbg, 0, 148 * 6 * sizeof(long long), as_stream(stream)));
        p.dbg = dbg;
        conv_stem_pair_kernel<<<p.total < num_sms ? p.total : num_sms, SP_THREADS, SP_SMEM, as_stream(stream)>>>(tmX, tmW2, tmO, q);
        TDRN_LAUNCH_CHECK();
        long long h[148 * 6];
        TDRN_CUDA(cudaMemcpyAsync(h, dbg, sizeof(h), cudaMemcpyDeviceToHost, as_stream(stream)));
        TDRN_CUDA(cudaStreamSynchronize(as_stream(stream)));
        double t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 148; ++i) for (int j = 0; j < 6; ++j) t[j] += (double)h[i * 6 + j] / 148.0;
        fprintf(stderr, "stem pair timing: MMA-issuer cycles per tile: total %.0f = waiting for the im2col rows (ap_full) %.0f + for the mid stage to "
                        "drain D' (pre_empty) %.0f + for the epilogue (t_empty) %.0f + for the halo tile (a_full) %.0f + issuing %.0f  (%.0f tiles per CTA)\n",
                t[0] / t[5], t[1] / t[5], t[2] / t[5], t[3] / t[5], t[4] / t[5], (t[0] - t[1] - t[2] - t[3] - t[4]) / t[5], t[5]);
        return TDRN_OK;
    }
    conv_stem_pair_kernel<<<p.total < num_sms ? p.total : num_sms, SP_THREADS, SP_SMEM, as_stream(stream)>>>(tmX, tmW2, tmO, q);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}
