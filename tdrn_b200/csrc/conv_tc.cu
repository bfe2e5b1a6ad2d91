// conv_tc.cu -- implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 / TMEM / TMA).
//
// Replaces the cuDNN calls behind every dense nn.Conv2d / nn.ConvTranspose2d(k2,s2) of the detectors
// (model/networks.py:136-163 vgg(), model/dualrefinedet_vggbn.py:30-50,97-114)
// for bf16 NHWC activations.  BatchNorm is folded into weight/bias on the host; bias, residual add
// (`up(x) + t`, dualrefinedet_vggbn.py:177), ReLU, the following MaxPool2d(2,2) and the output cast
// are the epilogue.
//
// GEMM view:  D[M = 128 output pixels, N <= BN couts] = sum over (tap, 64-channel block)
//             A[128 px, 64 ch] * B[N, 64 ch]^T,  fp32 accumulation in TMEM.
//   A : no im2col buffer anywhere.  One 4-D TMA box (64 ch, bw, bh, bn) of the NHWC input, shifted by
//       the tap offset (element stride 2 for the stride-2 conv); out-of-bounds rows/cols (the conv zero
//       padding) are zero-filled by TMA.  The box lands in shared memory as 128-byte rows with the
//       128B swizzle = the canonical K-major UMMA layout, pixel index (n, y, x) = tile row.
//   B : packed weights [Cout_pad][kh*kw*Cin] bf16 (k = tap*Cin + c), 2-D TMA box (64, BN).
//
// Persistent, warp-specialised: grid = min(#tiles, #SMs), each CTA walks tiles blockIdx.x, +grid, ...
//   warp 0   TMA producer            (ring of STAGES smem stages, full/empty mbarriers)
//   warp 1   TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2-9 epilogue (two per TMEM lane quadrant): tcgen05.ld -> bias / residual / ReLU / 2x2 max-pool -> global
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the
// main loop of tile i+1; the smem ring is 192 KB deep so TMA latency is hidden even when a layer has
// fewer tiles than SMs (the 10x10 / 5x5 pyramid levels).
#include "tc_common.cuh"
#include <stdlib.h>
#include <stdio.h>

namespace tdrn {
namespace tc {

EncodeTiledFn get_encode_tiled()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int make_tmap_bf16(CUtensorMap *m, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                   const uint32_t *box, const uint32_t *elem_strides)
{
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return TDRN_ECUDA; }
    cuuint64_t gdim[5]; cuuint64_t gstr[4]; cuuint32_t bdim[5]; cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = elem_strides ? elem_strides[i] : 1; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstr, bdim, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d) rank=%d dims=(%llu,%llu,%llu,%llu) box=(%u,%u,%u,%u)", (int)r, rank,
                  (unsigned long long)dims[0], (unsigned long long)dims[1], rank > 2 ? (unsigned long long)dims[2] : 0ull,
                  rank > 3 ? (unsigned long long)dims[3] : 0ull, box[0], box[1], rank > 2 ? box[2] : 0u, rank > 3 ? box[3] : 0u);
        return TDRN_ECUDA;
    }
    return TDRN_OK;
}

struct TcConvP {
    int H, W, B;                 // output (= tile) space, before pooling
    int bw, bh, bn;              // A box: bw*bh*bn <= 128 pixels
    int tiles_w, tiles_h, m_tiles, n_tiles;
    int Cin, Cout, n_total;      // n_total = Cout, or 4*Cout for the k2s2 deconv
    int kw, taps, pad, dil, stride;
    uint32_t a_bytes;            // bytes one A box deposits (bw*bh*bn*128)
    // ragged-tail tiles: when H = qh*bh + rr (rr > 0) with full-width single-image boxes, the rr leftover rows of g2
    // consecutive images are batched into one tile (second tensor map, box (64, bw, rr, g2)); nA = #regular tiles
    int nA, qh, rr, g2; uint32_t a_bytes2;
    int l2_prefetch;             // producer prefetches its first weight slice into L2 (few-tile layers)
    int mt2;                     // work unit = two M tiles sharing every weight box (kernel template MT = 2)
    int tma_out;                 // epilogue stages bf16 tiles in shared memory and writes them with TMA stores (plain NHWC output)
    int b_resident;              // few k-blocks (wide 1x1 layers): the weight boxes of an N tile live in their own shared-memory
                                 // region and stay there across consecutive M tiles (CTAs walk contiguous tile ranges); the rest
    int a_stages;                // of the 192 KB is a ring of a_stages activation boxes (16 KB each)
    uint32_t b_bytes;            // bytes one B box deposits (min(BN, n_pad16)*128)
    int splitk;                  // > 1: K is split over a thread-block cluster of this many CTAs (conv_splitk_kernel)
    int split_nacc;              // split mode: accumulators the hi * W_hi products are spread over (1 or 3)
    int split_g;                 // > 0: output written as (hi | lo) bf16 pairs in groups of split_g channels (tdrn_conv_desc.split_out)
    int split_cb;                // > 0: fp32-accurate mode (tdrn_conv_desc.split3) with split_cb channel blocks per tap: a ring stage holds
                                 // FOUR boxes of one (tap, channel block) -- x_hi, x_lo, W_hi, W_lo; the activation tensor is [hi | lo],
                                 // the weight K axis is taps x [W_hi | W_lo] -- and feeds three products: hi*W_hi, hi*W_lo, lo*W_hi
    const float *bias; const void *res; void *out;
    long long out_sb, out_sp; int out_w;
    int relu, deconv, out_f32, pool;
    int f16_in, f16_out;         // TDRN_F16: operands (activations AND packed weights) / the 16-bit output are IEEE half instead of bf16
    int mt_major;                // tile index = mu * n_tiles + nt (the N tiles of one M tile run at the same time on neighbouring CTAs: the
                                 // activation boxes of the second one hit L2) instead of nt * m_units + mu
};

// split mode (fp32-accurate): ring stage = x_hi | x_lo | W_hi | W_lo boxes of one (tap, channel block); fp32 / (hi|lo) register-store epilogue
template <int BN> struct TcSplitCfg {
    static constexpr int A_BYTES = 128 * 128;
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = (192 * 1024) / STAGE_BYTES;          // 2 / 3 / 4 stages for BN = 256 / 128 / 64
    static constexpr int OUT_STAGE_BYTES = 2 * 128 * 128;              // (hi box | lo box) of the TMA-store epilogue, [128 px][64 ch] bf16 each
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_STAGE_BYTES + 1024;
};

template <int BN, int MT = 1> struct TcCfg {
    static constexpr int A_BYTES = 128 * 128;
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = MT * A_BYTES + B_BYTES;         // MT activation boxes (M tiles) share one weight box
    static constexpr int STAGES = (192 * 1024) / STAGE_BYTES;          // 4 / 6 / 8 stages for BN = 256 / 128 / 64 (MT = 2, BN = 256: 3)
    static constexpr int TMEM_COLS = 2 * BN;                           // MT = 1: double-buffered accumulator; MT = 2: two accumulators
    static constexpr int OUT_STAGE_BYTES = 128 * 128;                  // one [128 px][64 ch] bf16 box of the TMA-store epilogue
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * OUT_STAGE_BYTES + 1024;   // + slack for manual 1024B alignment
};

constexpr int TC_THREADS = 320;           // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue

// CL = 2: the kernel runs as thread-block clusters of two CTAs that own two different M tiles of the SAME N tile and
// walk the k-blocks in lock step; each CTA fetches half of every weight box and multicasts it into both CTAs'
// shared memory, so the weight (B) traffic L2->SM per CTA is halved (the 40x40 / 20x20 layers are L2->SM bound).
// RES: resident weight boxes (TcConvP::b_resident) -- a template parameter so that the ring arithmetic of the common
// streamed case stays compile-time (the single MMA-issuing thread has ~500 cycles per k-block for everything it does).
// MT = 2: a work unit is TWO consecutive M tiles of one N tile: every weight box is multiplied with two activation boxes
// (64 instead of 94 operand bytes per MMA cycle from L2, as in conv_halo_stream_kernel); the two accumulators fill TMEM, so
// the epilogue of a unit is not overlapped with the next unit's main loop -- worth it for the long-K 3x3 layers on the
// 40x40 / 20x20 maps (conv4_x, conv5_x, TCB), which are bound by operand delivery.  TMA-store epilogue only.
// SPLIT (fp32-accurate mode, tdrn_conv_desc.split3): the tensor core adds every K = 16 step into the fp32 accumulator with
// truncation, so an accumulator that takes all 3 * K/16 steps of a long-K layer drifts by ~1e-5 per layer (measured: 1.3e-4
// over the 17 stacked layers, above the 1e-4 bar).  The big products (hi * W_hi) are therefore spread round-robin over
// NACC = 3 accumulators (each takes a third of the steps; 3 + 1 accumulators of BN = 128 columns fill TMEM), the two small products
// (hi * W_lo, lo * W_hi: 2^-8 of the magnitude, their truncation does not matter) share one more, and the epilogue adds the
// accumulators in fp32 registers.  All 512 TMEM columns belong to one tile: no accumulator double-buffering in this mode.
// Operands: each (tap, channel block) brings x_hi, x_lo, W_hi, W_lo ONCE (64 KB at BN = 128) for its three products -- 85 bytes per
// MMA cycle from L2 instead of the 128 of one (A, B) pair per product, which is what bound the first version (r02i: 560 TFLOP/s
// of products on the 40x40 layers against 1 400 for the bf16 kernel).
template <int BN, int CL, bool RES, int MT, int SPLIT = 0>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmA2,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const __grid_constant__ CUtensorMap tmO,
                                                                const __grid_constant__ CUtensorMap tmO2, const TcConvP p)
{
    static_assert(MT == 1 || (CL == 1 && !RES), "two-M-tile units: single CTA, streamed weights");
    static_assert(!SPLIT || ((SPLIT + 1) * BN <= 512 && CL == 1 && !RES && MT == 1), "split mode: SPLIT + 1 accumulators of BN columns, single CTA, streamed weights");
    using Cfg = TcCfg<BN, MT>;
    constexpr int NACC = SPLIT ? SPLIT : 1;                           // split mode: accumulators of the hi * W_hi products (the host picks
                                                                      // ONE value for all tile shapes: results must not depend on BN, i.e. on the batch size)
    constexpr uint32_t ACC_COLS = SPLIT ? (uint32_t)((SPLIT + 1) * BN) : (uint32_t)BN;     // TMEM columns of one tile's accumulator set
    // accumulator sets a unit alternates between (the epilogue of tile i overlaps the main loop of tile i+1): two where TMEM
    // holds them -- in split mode that is BN = 64 with 3 + 1 accumulators, BN <= 128 with 1 + 1
    constexpr uint32_t NBUF = MT == 2 ? 1u : (SPLIT ? (2u * ACC_COLS <= 512u ? 2u : 1u) : 2u);
    constexpr int TMEM_COLS = SPLIT ? (NBUF * ACC_COLS > 256 ? 512 : (NBUF * ACC_COLS > 128 ? 256 : 128)) : Cfg::TMEM_COLS;
    extern __shared__ uint8_t smem_dyn[];
    constexpr int MAX_STAGES = Cfg::STAGES > 8 ? Cfg::STAGES : 8;
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_s;

    uint8_t *tiles = (uint8_t *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_pad16 = (p.n_total + 15) & ~15;
    const int cblocks = p.Cin >> 6;
    const int num_kb = p.taps * cblocks;
    // work unit = CL M tiles x one N tile; CTA `cr` of the cluster takes M tile mu*CL + cr (may be a dummy beyond
    // m_tiles: its A boxes are entirely out of bounds -> zeros, its epilogue stores nothing)
    const int cr = CL == 2 ? (int)cluster_ctarank() : 0;
    const int m_units = (p.m_tiles + CL * MT - 1) / (CL * MT);
    const int total_tiles = m_units * p.n_tiles;
    const int unit0 = blockIdx.x / CL, unit_step = gridDim.x / CL;
    // tile walk: interleaved (tile = unit0, +unit_step, ...) or, with resident weights, one contiguous range per CTA so
    // that its tiles share the N tile (tile = nt * m_units + mt)
    const int tile_begin = RES ? (int)((long long)total_tiles * unit0 / unit_step) : unit0;
    const int tile_end = RES ? (int)((long long)total_tiles * (unit0 + 1) / unit_step) : total_tiles;
    const int tile_step = RES ? 1 : unit_step;
    const bool mt_major = !RES && p.mt_major;
    auto tile_nt = [&](int t) { return mt_major ? t % p.n_tiles : t / m_units; };     // N tile of work item t
    auto tile_mu = [&](int t) { return mt_major ? t / p.n_tiles : t % m_units; };     // its M unit
    // shared-memory layout: ring of (A box | B box) stages, or -- resident weights -- num_kb B boxes followed by a ring of A boxes
    const uint32_t n_stages = RES ? (uint32_t)p.a_stages : (SPLIT ? (uint32_t)TcSplitCfg<BN>::STAGES : (uint32_t)Cfg::STAGES);
    constexpr uint32_t a_stride = RES ? (uint32_t)Cfg::A_BYTES : (SPLIT ? (uint32_t)TcSplitCfg<BN>::STAGE_BYTES : (uint32_t)Cfg::STAGE_BYTES);
    uint8_t *a_base = RES ? tiles + num_kb * Cfg::B_BYTES : tiles;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        if (p.rr) tma_prefetch_desc(&tmA2);
        tma_prefetch_desc(&tmB);
#pragma unroll
        for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CL); }
#pragma unroll
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 8); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();          // the peer's barriers are initialised before anything is multicast at them
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    pdl_sync();                               // programmatic dependent launch: everything above overlaps the previous kernel's tail

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            // The weights of the small-map layers are cold in L2 every step (1 GB of activations went through it since
            // their last use) and each CTA streams them through a ring that only holds a few k-blocks: prefetch this
            // CTA's first weight slice into L2 now, all boxes at once.
            if (p.l2_prefetch && unit0 < total_tiles) {
                const int n0p = tile_nt(unit0) * BN + (CL == 2 ? cr * (int)(p.b_bytes >> 8) : 0);
                for (int kb = 0; kb < num_kb; ++kb) tma_prefetch_l2_2d(&tmB, kb * 64, n0p);
            }
            uint32_t it = 0, s = 0, ph = 0;                    // running k-block counter across tiles; its ring stage and phase
            int nt_in_smem = -1;                               // b_resident: N tile whose weight boxes sit in the stages
            for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
                const int nt = tile_nt(tile);
                int w0[MT], h0[MT], b0[MT];
                const CUtensorMap *mapA[MT];
                uint32_t a_bytes = 0;
                bool have[MT];
#pragma unroll
                for (int sub = 0; sub < MT; ++sub) {
                    const int mt = tile_mu(tile) * (CL * MT) + (MT == 2 ? sub : cr);
                    have[sub] = MT == 1 || mt < p.m_tiles;          // an odd M tile count leaves the last unit half empty
                    const bool tail = p.rr && mt >= p.nA;           // ragged-tail tile (leftover rows of g2 images)
                    const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
                    w0[sub] = tw * p.bw * p.stride - p.pad;
                    h0[sub] = (tail ? p.qh * p.bh : th * p.bh) * p.stride - p.pad;
                    b0[sub] = tail ? (mt - p.nA) * p.g2 : tn * p.bn;
                    mapA[sub] = tail ? &tmA2 : &tmA;
                    if (have[sub]) a_bytes += tail ? p.a_bytes2 : p.a_bytes;
                }
                const int n0 = nt * BN;
                const bool load_b = !(RES && nt == nt_in_smem);
                nt_in_smem = nt;
                if (RES && load_b) {
                    // new N tile: its weight boxes overwrite the resident region -> every MMA issued so far must have retired
                    for (uint32_t k = it > n_stages ? it - n_stages : 0u; k < it; ++k)
                        mbar_wait(&empty_bar[k % n_stages], (k / n_stages) & 1u);
                }
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    mbar_wait(&empty_bar[s], ph ^ 1u);        // CL = 2: both CTAs' MMAs have released stage s
                    uint8_t *sa = a_base + s * a_stride;
                    uint8_t *sb = RES ? tiles + kb * Cfg::B_BYTES : sa + MT * Cfg::A_BYTES;
                    const int tap = kb / cblocks, cb = kb - tap * cblocks;
                    const int tr = tap / p.kw, ts = tap - tr * p.kw;
                    if (SPLIT) {
                        // x_hi | x_lo | W_hi | W_lo of this (tap, channel block); weight column = (tap * 2 + half) * Cin + cb * 64
                        mbar_expect_tx(&full_bar[s], 2u * a_bytes + 2u * p.b_bytes);
                        tma_load_4d(sa, mapA[0], &full_bar[s], cb * 64, w0[0] + ts * p.dil, h0[0] + tr * p.dil, b0[0]);
                        tma_load_4d(sa + Cfg::A_BYTES, mapA[0], &full_bar[s], (cblocks + cb) * 64, w0[0] + ts * p.dil, h0[0] + tr * p.dil, b0[0]);
                        tma_load_2d(sa + 2 * Cfg::A_BYTES, &tmB, &full_bar[s], (2 * tap * cblocks + cb) * 64, n0);
                        tma_load_2d(sa + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &tmB, &full_bar[s], ((2 * tap + 1) * cblocks + cb) * 64, n0);
                        if (++s == n_stages) { s = 0; ph ^= 1u; }
                        continue;
                    }
                    mbar_expect_tx(&full_bar[s], a_bytes + (load_b ? p.b_bytes : 0u));
#pragma unroll
                    for (int sub = 0; sub < MT; ++sub)
                        if (have[sub])
                            tma_load_4d(sa + sub * Cfg::A_BYTES, mapA[sub], &full_bar[s], cb * 64, w0[sub] + ts * p.dil,
                                        h0[sub] + tr * p.dil, b0[sub]);
                    if (!load_b) {
                        // this stage still holds k-block kb of the same N tile
                    } else if (CL == 2) {
                        const uint32_t half_rows = p.b_bytes >> 8;           // (b_bytes / 128) / 2 rows of the weight box
                        tma_load_2d_mc(sb + cr * half_rows * 128u, &tmB, &full_bar[s], kb * 64,
                                       n0 + cr * (int)half_rows, (uint16_t)3);
                    } else {
                        tma_load_2d(sb, &tmB, &full_bar[s], kb * 64, n0);
                    }
                    if (++s == n_stages) { s = 0; ph ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (elect_one()) {
            uint32_t s = 0, ph = 0, tcount = 0;
            for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++tcount) {
                const int nt = tile_nt(tile);
                const int n_eff = min(BN, n_pad16 - nt * BN);           // UMMA N (multiple of 16)
                const uint32_t idesc = umma_idesc_16(128, n_eff, p.f16_in);
                const uint32_t buf = tcount % NBUF;
                mbar_wait(&tmem_empty_bar[buf], ((tcount / NBUF) & 1u) ^ 1u);   // epilogue drained this buffer
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * ACC_COLS;
                const bool have1 = MT == 2 && tile_mu(tile) * 2 + 1 < p.m_tiles;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(a_base + s * a_stride);
                    const uint64_t adesc = umma_desc_sw128(sa);
                    const uint64_t bdesc = umma_desc_sw128(RES ? smem_u32(tiles + kb * Cfg::B_BYTES) : sa + MT * Cfg::A_BYTES);
                    if (SPLIT) {
                        // hi * W_hi -> accumulator kb % NACC (each takes a third of the truncating steps); hi * W_lo and lo * W_hi -> the last one
                        const uint64_t alo = umma_desc_sw128(sa + Cfg::A_BYTES);
                        const uint64_t bhi = umma_desc_sw128(sa + 2 * Cfg::A_BYTES), blo = umma_desc_sw128(sa + 2 * Cfg::A_BYTES + Cfg::B_BYTES);
                        const uint32_t a = (uint32_t)kb % (uint32_t)NACC;
                        const uint32_t d_big = d_tmem + a * BN, d_small = d_tmem + (uint32_t)NACC * BN;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(d_big, adesc + (uint64_t)(k * 2), bhi + (uint64_t)(k * 2), idesc, kb >= NACC || k != 0);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(d_small, adesc + (uint64_t)(k * 2), blo + (uint64_t)(k * 2), idesc, (kb | k) != 0);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(d_small, alo + (uint64_t)(k * 2), bhi + (uint64_t)(k * 2), idesc, true);
                        umma_commit(&empty_bar[s]);
                        if (++s == n_stages) { s = 0; ph ^= 1u; }
                        continue;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k)      // 4 x (K = 16 bf16 = 32 bytes) inside the 128-byte swizzle atom
                        umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    if (have1) {                     // second M tile of the unit: same weight box, accumulator in the upper columns
                        const uint64_t adesc1 = umma_desc_sw128(sa + Cfg::A_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(d_tmem + BN, adesc1 + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    }
                    if (CL == 2) umma_commit_mc(&empty_bar[s], (uint16_t)3);   // both producers multicast into this stage
                    else umma_commit(&empty_bar[s]);     // frees the smem stage when these MMAs retire
                    if (++s == n_stages) { s = 0; ph ^= 1u; }
                }
                umma_commit(&tmem_full_bar[buf]);   // accumulator of this tile complete
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (warps 2..9) =====================
        // Two warps per TMEM lane quadrant (a warp may only read lanes 32*(warp%4)..+31); they take alternate
        // 16-column chunks.  The residual (`up(x) + t`) of the NEXT chunk is fetched before the current chunk is
        // processed, so its global-memory latency overlaps the TMEM load / arithmetic / stores.
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int r = quad * 32 + lane;
        if (SPLIT != 0 && p.tma_out) {
            // ---- split mode, (hi | lo) bf16 output [B,Ho,Wo,2*Cout] (optionally 2x2 max-pooled): sum of the accumulators -> bias /
            // ReLU / pool in fp32 -> hi = bf16(v), lo = bf16(v - hi) -> two 128B-swizzled staging boxes -> two TMA stores (channels
            // n .. n+63 and Cout + n .. Cout + n+63).  Same arithmetic, in the same order, as the register-store path below.
            // Unpooled tiles use one (hi | lo) box pair and wait for its previous store to be read out right before refilling it
            // (the loads and the arithmetic of the group hide that); pooled tiles (32 rows) alternate between two pairs.
            uint8_t *stage_base = tiles + TcSplitCfg<BN>::STAGES * TcSplitCfg<BN>::STAGE_BYTES;
            const bool leader = threadIdx.x == 64;
            const int used = num_kb < NACC ? num_kb : NACC;
            const int wl = r % p.bw, hl = (r / p.bw) % p.bh;
            const uint32_t box_bytes = p.pool ? 32u * 128u : 128u * 128u;
            uint32_t tcount = 0, git = 0;
            for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++tcount) {
                const int mt = tile_mu(tile), nt = tile_nt(tile);
                const int n0 = nt * BN;
                const int n_eff = min(BN, n_pad16 - n0);
                const uint32_t buf = tcount % NBUF;
                mbar_wait(&tmem_full_bar[buf], (tcount / NBUF) & 1u);
                tc_fence_after();
                const int groups = (n_eff + 63) >> 6;
                const bool tail = p.rr && mt >= p.nA;
                const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
                const int x0 = tw * p.bw, y0 = tail ? p.qh * p.bh : th * p.bh, b0 = tail ? (mt - p.nA) * p.g2 : tn * p.bn;
                const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * ACC_COLS;
                for (int g = 0; g < groups; ++g, ++git) {
                    uint8_t *o = stage_base + (p.pool ? (git & 1u) * 2u * box_bytes : 0u);
                    const int c0 = g * 64 + half * 32;
                    const bool have = c0 < n_eff;                    // warp-uniform
                    uint32_t qh[16], ql[16];
                    if (have) {
                        float v[32];
                        {
                            uint32_t ra[32], rb[32];
                            tmem_ld32_issue(trow + (uint32_t)c0, ra);
                            tmem_ld32_issue(trow + (uint32_t)((NACC == 1 ? 1 : 1) * BN + c0), rb);   // accumulator 1 (NACC = 1: the small-product one)
                            tmem_ld_wait32(ra);
                            tmem_ld_pin32(rb);
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(ra[j]);
                            if (NACC == 1 || used > 1) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(rb[j]);
                            }
                        }
                        if (NACC == 3) {
                            uint32_t ra[32], rb[32];
                            tmem_ld32_issue(trow + (uint32_t)(2 * BN + c0), ra);
                            tmem_ld32_issue(trow + (uint32_t)(3 * BN + c0), rb);
                            tmem_ld_wait32(ra);
                            tmem_ld_pin32(rb);
                            if (used > 2) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(ra[j]);
                            }
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(rb[j]);
                        }
                        if (g == groups - 1) {                          // all tcgen05.ld of this tile are complete
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
                        }
                        if (p.bias) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 bq = __ldg((const float4 *)(p.bias + n0 + c0 + j));
                                v[j] += bq.x; v[j + 1] += bq.y; v[j + 2] += bq.z; v[j + 3] += bq.w;
                            }
                        }
                        if (p.pool) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                float m = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
                                v[j] = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, p.bw));
                            }
                        }
                        if (p.relu) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * j]), h1 = __float2bfloat16_rn(v[2 * j + 1]);
                            const __nv_bfloat16 l0 = __float2bfloat16_rn(__fsub_rn(v[2 * j], __bfloat162float(h0)));
                            const __nv_bfloat16 l1 = __float2bfloat16_rn(__fsub_rn(v[2 * j + 1], __bfloat162float(h1)));
                            qh[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                            ql[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                        }
                    } else if (g == groups - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
                    }
                    if (leader) { if (p.pool) bulk_wait_read<1>(); else bulk_wait_read<0>(); }   // the stores that last read these boxes have drained
                    named_bar(1, 256);
                    if (have) {
                        int row = r;
                        bool wr = true;
                        if (p.pool) { wr = !(wl & 1) && !(hl & 1); row = (hl >> 1) * (p.bw >> 1) + (wl >> 1); }
                        if (wr) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                *(uint4 *)(o + sw128_offset(row, half * 4 + k)) = make_uint4(qh[4 * k], qh[4 * k + 1], qh[4 * k + 2], qh[4 * k + 3]);
                                *(uint4 *)(o + box_bytes + sw128_offset(row, half * 4 + k)) = make_uint4(ql[4 * k], ql[4 * k + 1], ql[4 * k + 2], ql[4 * k + 3]);
                            }
                        }
                    }
                    fence_proxy_async_smem();
                    named_bar(1, 256);
                    if (leader) {
                        const CUtensorMap *m = tail ? &tmO2 : &tmO;
                        const int xo = p.pool ? x0 >> 1 : x0, yo = p.pool ? y0 >> 1 : y0;
                        tma_store_4d(m, o, n0 + g * 64, xo, yo, b0);
                        tma_store_4d(m, o + box_bytes, p.Cout + n0 + g * 64, xo, yo, b0);
                        bulk_commit();
                    }
                }
            }
            if (leader) bulk_wait_read<0>();
        } else if (p.tma_out) {
            // ---- plain NHWC bf16 output: TMEM -> bias/ReLU -> bf16 -> 128B-swizzled [128 px][64 ch] staging box -> TMA store.
            // Whole 128-byte lines leave the SM (a per-thread store writes 32 bytes of 32 different lines), rows beyond the
            // map / channels beyond Cout are clipped by TMA.  Two staging boxes: the store of column group g drains while
            // group g+1 is converted.  Each warp converts 32 rows x 32 columns of the group.
            uint8_t *stage_base = tiles + Cfg::STAGES * Cfg::STAGE_BYTES;
            const bool leader = threadIdx.x == 64;
            uint32_t tcount = 0, git = 0;
            for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++tcount) {
                const int nt = tile_nt(tile);
                const int n0 = nt * BN;
                const int n_eff = min(BN, n_pad16 - n0);
                const uint32_t buf = tcount % NBUF;
                mbar_wait(&tmem_full_bar[buf], (tcount / NBUF) & 1u);
                tc_fence_after();
                const int groups = (n_eff + 63) >> 6;
                const int nsub = MT == 2 && tile_mu(tile) * 2 + 1 < p.m_tiles ? 2 : 1;
                for (int sub = 0; sub < nsub; ++sub) {
                    const int mt = tile_mu(tile) * (CL * MT) + (MT == 2 ? sub : cr);
                    const bool tail = p.rr && mt >= p.nA;
                    const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
                    const int x0 = tw * p.bw, y0 = tail ? p.qh * p.bh : th * p.bh, b0 = tail ? (mt - p.nA) * p.g2 : tn * p.bn;
                    const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (MT == 2 ? sub * BN : buf * BN);
                    // r03: on layers with a bias the accumulator load and the bias loads of a group are issued BEFORE the barrier that
                    // waits for the staging box -- on short-K layers (MobileNet's pointwise convs) the epilogue's chain of latencies
                    // (barrier -> tcgen05.ld -> bias from global -> convert -> store -> barrier), four times per tile, was longer than
                    // the tile's MMAs (ncu source view r03m: the bias FADDs wait on the long scoreboard; 512 -> 512 @40x40 b64:
                    // 0.074 -> 0.064 ms).  Measured and NOT kept: issuing group g + 1's loads while group g is converted (pointwise
                    // convs another 3 % faster, but the bias-free per-tap projection GEMM 0.096 -> 0.132 ms), and the early issue on
                    // that bias-free GEMM (+4 %): it keeps the original order.
                    const bool early = p.bias != nullptr;
                    for (int g = 0; g < groups; ++g, ++git) {
                        uint8_t *o = stage_base + (git & 1u) * Cfg::OUT_STAGE_BYTES;
                        const int c0 = g * 64 + half * 32;
                        const bool have = c0 < n_eff;                    // warp-uniform; columns >= n_eff are >= Cout: clipped
                        const int n = n0 + c0;
                        if (early) {
                            uint32_t ra[32];
                            float bv[32];
                            if (have) {
                                tmem_ld32_issue(trow + (uint32_t)c0, ra);
                                if (n + 32 <= p.n_total) {
#pragma unroll
                                    for (int j = 0; j < 32; j += 4) {
                                        const float4 b4 = __ldg((const float4 *)(p.bias + n + j));      // n is a multiple of 32
                                        bv[j] = b4.x; bv[j + 1] = b4.y; bv[j + 2] = b4.z; bv[j + 3] = b4.w;
                                    }
                                } else {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) bv[j] = n + j < p.n_total ? __ldg(p.bias + n + j) : 0.f;
                                }
                            }
                            if (leader) bulk_wait_read<1>();             // the store that last read this box has drained
                            named_bar(1, 256);
                            if (have) {
                                tmem_ld_wait32(ra);
                                float v[32];
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(ra[j]) + bv[j];
                                if (p.relu) {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                                }
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    uint32_t w[4];
#pragma unroll
                                    for (int j = 0; j < 4; ++j) w[j] = pack16x2(v[q * 8 + 2 * j], v[q * 8 + 2 * j + 1], p.f16_out);
                                    *(uint4 *)(o + sw128_offset(r, half * 4 + q)) = make_uint4(w[0], w[1], w[2], w[3]);
                                }
                            }
                        } else {
                            if (leader) bulk_wait_read<1>();             // the store that last read this box has drained
                            named_bar(1, 256);
                            if (have) {
                                float v[32];
                                tmem_ld32(trow + (uint32_t)c0, v);
                                if (p.relu) {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                                }
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    uint32_t w[4];
#pragma unroll
                                    for (int j = 0; j < 4; ++j) w[j] = pack16x2(v[q * 8 + 2 * j], v[q * 8 + 2 * j + 1], p.f16_out);
                                    *(uint4 *)(o + sw128_offset(r, half * 4 + q)) = make_uint4(w[0], w[1], w[2], w[3]);
                                }
                            }
                        }
                        if (g == groups - 1 && sub == nsub - 1) {        // all tcgen05.ld of this unit are complete
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
                        }
                        fence_proxy_async_smem();
                        named_bar(1, 256);
                        if (leader) {
                            tma_store_4d(tail ? &tmO2 : &tmO, o, n0 + g * 64, x0, y0, b0);
                            bulk_commit();
                        }
                    }
                }
            }
            if (leader) bulk_wait_read<0>();
        } else {
        const int wl = r % p.bw, hl = (r / p.bw) % p.bh, nl = r / (p.bw * p.bh);
        const int hl2 = p.rr ? (r / p.bw) % p.rr : 0, nl2 = p.rr ? r / (p.bw * p.rr) : 0;      // ragged-tail tiles
        const bool res_bf16 = p.res != nullptr && !p.out_f32;
        uint32_t tcount = 0;
        for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++tcount) {
            const int mt = tile_mu(tile) * CL + cr, nt = tile_nt(tile);
            const bool tail = p.rr && mt >= p.nA;
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
            const int x = tw * p.bw + wl;
            const int y = tail ? p.qh * p.bh + hl2 : th * p.bh + hl;
            const int b = tail ? (mt - p.nA) * p.g2 + nl2 : tn * p.bn + nl;
            const int n0 = nt * BN;
            const int n_eff = min(BN, n_pad16 - n0);
            const bool valid = (tail ? nl2 < p.g2 : nl < p.bn) && x < p.W && y < p.H && b < p.B;
            // element offset of output channel n of this thread's pixel (deconv: pixel-shuffled position)
            auto out_off = [&](int n) -> long long {
                int co = n, oy = y, ox = x;
                if (p.deconv) { const int ij = n / p.Cout; co = n - ij * p.Cout; oy = 2 * y + (ij >> 1); ox = 2 * x + (ij & 1); }
                else if (p.pool) { oy = y >> 1; ox = x >> 1; }
                if (p.split_g) co = n + (n / p.split_g) * p.split_g;       // high halves; the low halves sit split_g further
                return (long long)b * p.out_sb + ((long long)oy * p.out_w + ox) * p.out_sp + co;
            };
            uint4 rn[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
            auto prefetch_res = [&](int c0) {
                const int n = n0 + c0;
                if (!res_bf16 || !valid || c0 >= n_eff || n + 16 > p.n_total) return;
                const long long o = out_off(n);
                if (o & 7) return;
                const uint4 *rp = (const uint4 *)((const __nv_bfloat16 *)p.res + o);
                rn[0] = rp[0]; rn[1] = rp[1];
            };
            prefetch_res(half * 16);
            const uint32_t buf = tcount % NBUF;
            mbar_wait(&tmem_full_bar[buf], (tcount / NBUF) & 1u);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * ACC_COLS;
            for (int c0 = half * 16; c0 < n_eff; c0 += 32) {
                const uint4 rc[2] = {rn[0], rn[1]};
                prefetch_res(c0 + 32);
                float v[16];
                tmem_ld16(trow + (uint32_t)c0, v);
                if (SPLIT) {                                         // + the other hi * W_hi accumulators in use + the small-product one
                    const int used = num_kb < NACC ? num_kb : NACC;
                    for (int a = 1; a <= NACC; ++a) {
                        if (a >= used && a != NACC) continue;
                        float u[16];
                        tmem_ld16(trow + (uint32_t)(a * BN + c0), u);
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += u[j];
                    }
                }
                const int n = n0 + c0;
                if (n >= p.n_total) continue;                       // warp-uniform
                const int co = p.deconv ? n % p.Cout : n;
                const int nv = min(16, p.n_total - n);
                if (p.bias) {
                    if (nv == 16) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 bq = __ldg((const float4 *)(p.bias + co + j));      // co is a multiple of 16 here
                            v[j] += bq.x; v[j + 1] += bq.y; v[j + 2] += bq.z; v[j + 3] += bq.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) if (j < nv) v[j] += __ldg(p.bias + co + j);
                    }
                }
                bool store = valid;
                if (p.pool) {
                    // MaxPool2d(2,2) fused: the 2x2 partners of pixel (hl, wl) live in lanes ^1 and ^bw of this warp
                    // (bw is a power of two <= 16, bh even, tile rows are (h, w)-ordered). relu/max commute.
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float m = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
                        v[j] = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, p.bw));
                    }
                    store = valid && !(wl & 1) && !(hl & 1);
                }
                if (!store) continue;
                const long long o = out_off(n);
                if (p.out_f32) {
                    float *op = (float *)p.out + o;
                    if (p.res) { const float *rp = (const float *)p.res + o;
#pragma unroll
                        for (int j = 0; j < 16; ++j) if (j < nv) v[j] += rp[j]; }
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (j < nv) op[j] = p.relu ? fmaxf(v[j], 0.f) : v[j];
                } else {
                    __nv_bfloat16 *op = (__nv_bfloat16 *)p.out + o;
                    const bool vec = nv == 16 && ((o & 7) == 0);
                    if (p.res) {
                        if (vec) {                                   // prefetched one chunk ago
                            const __nv_bfloat16 *rb = (const __nv_bfloat16 *)rc;
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] += __bfloat162float(rb[j]);
                        } else {
                            const __nv_bfloat16 *rp = (const __nv_bfloat16 *)p.res + o;
#pragma unroll
                            for (int j = 0; j < 16; ++j) if (j < nv) v[j] += __bfloat162float(rp[j]);
                        }
                    }
                    if (p.relu) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
                    }
                    if (vec && p.split_g) {                           // (hi | lo) pair: x = hi + lo to 16 mantissa bits
                        uint4 q[2], ql[2];
                        __nv_bfloat16 *qb = (__nv_bfloat16 *)q, *lb = (__nv_bfloat16 *)ql;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            qb[j] = __float2bfloat16_rn(v[j]);
                            lb[j] = __float2bfloat16_rn(__fsub_rn(v[j], __bfloat162float(qb[j])));
                        }
                        ((uint4 *)op)[0] = q[0]; ((uint4 *)op)[1] = q[1];
                        ((uint4 *)(op + p.split_g))[0] = ql[0]; ((uint4 *)(op + p.split_g))[1] = ql[1];
                    } else if (vec) {
                        uint32_t q[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) q[j] = pack16x2(v[2 * j], v[2 * j + 1], p.f16_out);
                        ((uint4 *)op)[0] = make_uint4(q[0], q[1], q[2], q[3]); ((uint4 *)op)[1] = make_uint4(q[4], q[5], q[6], q[7]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) if (j < nv) ((uint16_t *)op)[j] = pack16(v[j], p.f16_out);
                    }
                }
            }
            // all tcgen05.ld of this tile are complete (tmem_ld16 waits): hand the buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
        }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();          // no CTA leaves while its peer can still signal its barriers
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------------------
// Split-K variant for the small pyramid levels (10x10 / 5x5 maps at 320, 16x16 / 8x8 at 512: M = B * H * W is a few
// thousand rows, so a layer has 7..100 output tiles for 148 SMs while its K loop is 36..144 k-blocks long; measured r01:
// arm_loc.2 1024 -> 12 @10x10 runs 144 k-blocks on 25 SMs, 0.39 us per k-block = the latency of eight TMA boxes in flight,
// 12 TFLOP/s).  The k-blocks of ONE output tile are split over a thread-block cluster of S CTAs (S SMs pull operands for
// the tile); every CTA accumulates its K slice in its own TMEM, ranks 1..S-1 park the fp32 partial tile in their shared
// memory, and rank 0 adds them through distributed shared memory (ld.shared::cluster) in rank order -- deterministic --
// before the usual epilogue.  S depends on the layer's shape only (never on the batch size): results stay bit-identical
// across batch sizes.  BN = 64 (these layers are small: more N tiles = more SMs at work).
// ---------------------------------------------------------------------------------------------------------
constexpr int SK_BN = 64;
constexpr int SK_STAGES = 8;
constexpr int SK_STAGE_BYTES = 128 * 128 + SK_BN * 128;              // one A box + one B box
constexpr int SK_SMEM_BYTES = SK_STAGES * SK_STAGE_BYTES + 1024;

__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

__global__ void __launch_bounds__(TC_THREADS, 1) conv_splitk_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                    const __grid_constant__ CUtensorMap tmA2,
                                                                    const __grid_constant__ CUtensorMap tmB, const TcConvP p)
{
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[SK_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[SK_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_s;

    uint8_t *tiles = (uint8_t *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.splitk;
    const int rank = (int)cluster_ctarank();
    const int unit = blockIdx.x / S;                       // output tile: unit = nt * m_tiles + mt
    const int nt = unit / p.m_tiles, mt = unit - nt * p.m_tiles;
    const int n_pad16 = (p.n_total + 15) & ~15;
    const int cblocks = p.Cin >> 6;
    const int num_kb = p.taps * cblocks;
    const int kb0 = (int)((long long)num_kb * rank / S), kb1 = (int)((long long)num_kb * (rank + 1) / S);
    const int n0 = nt * SK_BN;
    const int n_eff = min(SK_BN, n_pad16 - n0);

    const bool tail = p.rr && mt >= p.nA;                  // ragged-tail tile (leftover rows of g2 images)
    const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        if (p.rr) tma_prefetch_desc(&tmA2);
        tma_prefetch_desc(&tmB);
#pragma unroll
        for (int s = 0; s < SK_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, SK_BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===================== TMA producer: k-blocks kb0 .. kb1-1 of this tile =====================
        if (elect_one()) {
            const int w0 = tw * p.bw * p.stride - p.pad;
            const int h0 = (tail ? p.qh * p.bh : th * p.bh) * p.stride - p.pad;
            const int b0 = tail ? (mt - p.nA) * p.g2 : tn * p.bn;
            const CUtensorMap *mapA = tail ? &tmA2 : &tmA;
            const uint32_t a_bytes = tail ? p.a_bytes2 : p.a_bytes;
            uint32_t s = 0, ph = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&empty_bar[s], ph ^ 1u);
                uint8_t *sa = tiles + s * SK_STAGE_BYTES;
                const int tap = kb / cblocks, cb = kb - tap * cblocks;
                const int tr = tap / p.kw, ts = tap - tr * p.kw;
                mbar_expect_tx(&full_bar[s], a_bytes + p.b_bytes);
                tma_load_4d(sa, mapA, &full_bar[s], cb * 64, w0 + ts * p.dil, h0 + tr * p.dil, b0);
                tma_load_2d(sa + 128 * 128, &tmB, &full_bar[s], kb * 64, n0);
                if (++s == SK_STAGES) { s = 0; ph ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_bf16(128, n_eff);
            uint32_t s = 0, ph = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(tiles + s * SK_STAGE_BYTES);
                const uint64_t adesc = umma_desc_sw128(sa);
                const uint64_t bdesc = umma_desc_sw128(sa + 128 * 128);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb > kb0) || k != 0);
                umma_commit(&empty_bar[s]);
                if (++s == SK_STAGES) { s = 0; ph ^= 1u; }
            }
            umma_commit(&tmem_full_bar);
        }
        __syncwarp();
    }

    // ===================== partial tiles: TMEM -> registers (-> shared memory on ranks 1..S-1) =====================
    // Two warps per TMEM lane quadrant; each takes two of the four 16-column chunks.  Partial layout in shared memory
    // (the TMA ring is idle by then): [chunk 0..3][row 0..127][16 floats] so that a warp's read of one chunk is 2 KB contiguous.
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    float v[2][16];
    float *partial = (float *)tiles;
    if (warp >= 2) {
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c0 = half * 16 + i * 32;
            if (c0 < n_eff) tmem_ld16(trow + (uint32_t)c0, v[i]);
        }
        if (rank != 0) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int c0 = half * 16 + i * 32;
                if (c0 < n_eff) {
                    float4 *dst = (float4 *)(partial + ((c0 >> 4) * 128 + r) * 16);
#pragma unroll
                    for (int q = 0; q < 4; ++q) dst[q] = make_float4(v[i][4 * q], v[i][4 * q + 1], v[i][4 * q + 2], v[i][4 * q + 3]);
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();                                   // every rank's partial tile is visible cluster-wide
    if (rank == 0 && warp >= 2) {
        const uint32_t local = smem_u32(partial);
        for (int rr = 1; rr < S; ++rr) {                  // fixed order: deterministic sum
            const uint32_t remote = mapa_shared(local, (uint32_t)rr);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int c0 = half * 16 + i * 32;
                if (c0 < n_eff) {
                    const uint32_t a = remote + (uint32_t)(((c0 >> 4) * 128 + r) * 64);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 u = ld_dsmem_f4(a + q * 16);
                        v[i][4 * q] += u.x; v[i][4 * q + 1] += u.y; v[i][4 * q + 2] += u.z; v[i][4 * q + 3] += u.w;
                    }
                }
            }
        }
    }
    cluster_sync_all();                                   // nobody leaves (and frees its shared memory) while rank 0 reads

    if (rank == 0 && warp >= 2) {
        // ===================== epilogue (rank 0): bias / residual / ReLU / pixel shuffle / store =====================
        const int wl = r % p.bw, hl = (r / p.bw) % p.bh, nl = r / (p.bw * p.bh);
        const int hl2 = p.rr ? (r / p.bw) % p.rr : 0, nl2 = p.rr ? r / (p.bw * p.rr) : 0;
        const int x = tw * p.bw + wl;
        const int y = tail ? p.qh * p.bh + hl2 : th * p.bh + hl;
        const int b = tail ? (mt - p.nA) * p.g2 + nl2 : tn * p.bn + nl;
        const bool valid = (tail ? nl2 < p.g2 : nl < p.bn) && x < p.W && y < p.H && b < p.B;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c0 = half * 16 + i * 32;
            const int n = n0 + c0;
            if (c0 >= n_eff || n >= p.n_total || !valid) continue;
            const int nv = min(16, p.n_total - n);
            int co = n, oy = y, ox = x;
            if (p.deconv) { const int ij = n / p.Cout; co = n - ij * p.Cout; oy = 2 * y + (ij >> 1); ox = 2 * x + (ij & 1); }
            const long long o = (long long)b * p.out_sb + ((long long)oy * p.out_w + ox) * p.out_sp + co;
            if (p.bias) {
#pragma unroll
                for (int j = 0; j < 16; ++j) if (j < nv) v[i][j] += __ldg(p.bias + co + j);
            }
            if (p.out_f32) {
                float *op = (float *)p.out + o;
                if (p.res) { const float *rp = (const float *)p.res + o;
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (j < nv) v[i][j] += rp[j]; }
#pragma unroll
                for (int j = 0; j < 16; ++j) if (j < nv) op[j] = p.relu ? fmaxf(v[i][j], 0.f) : v[i][j];
            } else {
                __nv_bfloat16 *op = (__nv_bfloat16 *)p.out + o;
                if (p.res) { const __nv_bfloat16 *rp = (const __nv_bfloat16 *)p.res + o;
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (j < nv) v[i][j] += __bfloat162float(rp[j]); }
                if (p.relu) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[i][j] = fmaxf(v[i][j], 0.f);
                }
                if (nv == 16 && ((o & 7) == 0)) {
                    uint4 q[2];
                    __nv_bfloat162 *qb = (__nv_bfloat162 *)q;
#pragma unroll
                    for (int j = 0; j < 8; ++j) qb[j] = __floats2bfloat162_rn(v[i][2 * j], v[i][2 * j + 1]);
                    ((uint4 *)op)[0] = q[0]; ((uint4 *)op)[1] = q[1];
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (j < nv) op[j] = __float2bfloat16_rn(v[i][j]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, SK_BN); }
}

// Pick the A box (bw, bh, bn), bw*bh*bn <= 128, maximising useful rows; bn > 1 only when one image's
// whole map fits (the 10x10 and 5x5 pyramid levels).  Ties -> wider rows.
void pick_box(int B, int H, int W, int max_w, int max_h, int &bw, int &bh, int &bn)
{
    double best = -1;
    bw = bh = bn = 1;
    for (int w = 1; w <= W && w <= 128 && w <= max_w; ++w)
        for (int h = 1; h <= H && w * h <= 128 && h <= max_h; ++h) {
            const int nmax = (w == W && h == H) ? 128 / (w * h) : 1;
            for (int n = 1; n <= nmax && n <= B; ++n) {
                const long long tiles = (long long)((W + w - 1) / w) * ((H + h - 1) / h) * ((B + n - 1) / n);
                const double eff = (double)B * H * W / (double)(tiles * 128) + 1e-6 * w;
                if (eff > best) { best = eff; bw = w; bh = h; bn = n; }
            }
        }
}

static int g_num_sms = 0;

template <int BN>
static int launch_tc(const CUtensorMap &tmA, const CUtensorMap &tmA2, const CUtensorMap &tmB, const CUtensorMap &tmO,
                     const CUtensorMap &tmO2, const TcConvP &p, bool use_cluster, cudaStream_t st)
{
    using Cfg = TcCfg<BN>;
    if (!g_num_sms) {
        int dev = 0;
        TDRN_CUDA(cudaGetDevice(&dev));
        TDRN_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    if (use_cluster) {
        static int max_clusters[3] = {0, 0, 0};
        const int slot = BN == 256 ? 2 : (BN == 128 ? 1 : 0);
        TDRN_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, 2, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        if (!max_clusters[slot]) {
            cfg.gridDim = dim3(g_num_sms);
            int n = 0;
            TDRN_CUDA(cudaOccupancyMaxActiveClusters(&n, conv_tc_kernel<BN, 2, false, 1>, &cfg));
            max_clusters[slot] = n > 0 ? n : 1;
        }
        const int units = ((p.m_tiles + 1) / 2) * p.n_tiles;
        const int clusters = units < max_clusters[slot] ? units : max_clusters[slot];
        cfg.gridDim = dim3(2 * clusters);
        TDRN_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, 2, false, 1>, tmA, tmA2, tmB, tmO, tmO2, p));
        count_launch();
        return TDRN_OK;
    }
    if constexpr (BN == 256) {
        if (p.mt2) {
            using Cfg2 = TcCfg<256, 2>;
            const int units = ((p.m_tiles + 1) / 2) * p.n_tiles;
            TDRN_CUDA(cudaFuncSetAttribute(conv_tc_kernel<256, 1, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2::SMEM_BYTES));
            TDRN_CUDA(launch_pdl(conv_tc_kernel<256, 1, false, 2>, dim3(units < g_num_sms ? units : g_num_sms), dim3(TC_THREADS), Cfg2::SMEM_BYTES, st, tmA, tmA2, tmB, tmO, tmO2, p));
            TDRN_LAUNCH_CHECK();
            return TDRN_OK;
        }
    }
    const int total = p.m_tiles * p.n_tiles;
    const int grid = total < g_num_sms ? total : g_num_sms;
    if (p.split_cb) {
        constexpr int smem = TcSplitCfg<BN>::SMEM_BYTES;
        if constexpr (BN <= 128) {
            if (p.split_nacc == 3) {
                TDRN_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, 1, false, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                TDRN_CUDA(launch_pdl(conv_tc_kernel<BN, 1, false, 1, 3>, dim3(grid), dim3(TC_THREADS), smem, st, tmA, tmA2, tmB, tmO, tmO2, p));
                TDRN_LAUNCH_CHECK();
                return TDRN_OK;
            }
        }
        TDRN_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, 1, false, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        TDRN_CUDA(launch_pdl(conv_tc_kernel<BN, 1, false, 1, 1>, dim3(grid), dim3(TC_THREADS), smem, st, tmA, tmA2, tmB, tmO, tmO2, p));
        TDRN_LAUNCH_CHECK();
        return TDRN_OK;
    }
    if (p.b_resident) {
        TDRN_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, 1, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        TDRN_CUDA(launch_pdl(conv_tc_kernel<BN, 1, true, 1>, dim3(grid), dim3(TC_THREADS), Cfg::SMEM_BYTES, st, tmA, tmA2, tmB, tmO, tmO2, p));
    } else {
        TDRN_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, 1, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        TDRN_CUDA(launch_pdl(conv_tc_kernel<BN, 1, false, 1>, dim3(grid), dim3(TC_THREADS), Cfg::SMEM_BYTES, st, tmA, tmA2, tmB, tmO, tmO2, p));
    }
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

}  // namespace tc
}  // namespace tdrn

namespace tdrn { namespace tc {
int conv_halo_try(const tdrn_conv_desc *d, const void *in, const void *weight, const float *bias, const void *residual,
                  void *out, cudaStream_t st);      // conv_halo_tc.cu
} }

using namespace tdrn;
using namespace tdrn::tc;

extern "C" int tdrn_conv2d_tc(const tdrn_conv_desc *d, const void *in, const void *weight, const float *bias,
                              const void *residual, void *out, tdrn_stream_t stream)
{
    TDRN_REQUIRE(d && in && weight && out, "tdrn_conv2d_tc: null argument");
    // Cin that is a multiple of 8 but not of 64 (MobileNet's 32-channel stem output): the channel box still asks for 64
    // channels, TMA zero-fills the ones beyond Cin, and the packed weights carry zero rows for them (K padded per tap).
    const bool f16_in = d->in_dtype == TDRN_F16, f16_out = d->out_dtype == TDRN_F16;
    if (f16_in || f16_out) {     // half operands / output: the plain conv kernel only (what the MobileNet trunk's 1x1 convs need)
        if (d->split3 || d->pool2x2 || d->deconv2x2 || residual || d->dg) {
            set_error("tdrn_conv2d_tc: TDRN_F16 is not available with split3 / fused pool / deconv / residual");
            return TDRN_EUNSUPPORTED;
        }
    }
    if ((d->in_dtype != TDRN_BF16 && !f16_in) || d->Cin % 8 != 0 || d->dg != 0 || d->in_sb != 0 ||
        (!d->deconv2x2 && d->stride != 1 && d->stride != 2)) {
        set_error("tdrn_conv2d_tc: needs bf16 input, Cin %% 8 == 0, stride 1 or 2, no offsets (got dtype=%d Cin=%d stride=%d dg=%d)",
                  d->in_dtype, d->Cin, d->stride, d->dg);
        return TDRN_EUNSUPPORTED;
    }
    TDRN_REQUIRE(!d->split3 || d->Cin % 64 == 0, "tdrn_conv2d_tc: split3 needs Cin %% 64 == 0 (got %d)", d->Cin);
    TDRN_REQUIRE(!d->split_out || (d->split3 && d->out_dtype == TDRN_BF16 && d->split_out % 16 == 0 && d->Cout % d->split_out == 0 &&
                                   d->out_sp == 2ll * d->Cout && !d->deconv2x2 && !residual && ((uintptr_t)out & 15) == 0),
                 "tdrn_conv2d_tc: split_out needs split3, bf16 out, g %% 16 == 0, Cout %% g == 0, out_sp == 2*Cout, no deconv / residual");
    {   // narrow high-resolution 3x3 layers: halo tile + resident weights (conv_halo_tc.cu)
        static const bool no_halo = getenv("TDRN_NO_HALO") != nullptr;
        if (!no_halo && !d->split3 && !f16_in && !f16_out) {
            const int rc = conv_halo_try(d, in, weight, bias, residual, out, as_stream(stream));
            if (rc != TDRN_EUNSUPPORTED) return rc;
        }
    }
    TcConvP p{};
    const int kh = d->deconv2x2 ? 1 : d->kh, kw = d->deconv2x2 ? 1 : d->kw;
    const int pad = d->deconv2x2 ? 0 : d->pad, dil = d->deconv2x2 ? 1 : d->dil;
    const int stride = d->deconv2x2 ? 1 : d->stride;
    p.H = (d->H + 2 * pad - (dil * (kh - 1) + 1)) / stride + 1;
    p.W = (d->W + 2 * pad - (dil * (kw - 1) + 1)) / stride + 1;
    TDRN_REQUIRE(p.H > 0 && p.W > 0, "convolution input is too small (output would be %dx%d)", p.H, p.W);
    // p.Cin is padded to 64: the k-block count and the weight K use it, the activation tensor map uses the real Cin
    p.B = d->B; p.Cin = (d->Cin + 63) & ~63; p.Cout = d->Cout; p.n_total = d->deconv2x2 ? 4 * d->Cout : d->Cout;
    const int cin_mem = d->split3 ? 2 * d->Cin : d->Cin;        // channels of the activation tensor in memory ([hi | lo] when split)
    if (d->split3) { p.split_cb = d->Cin >> 6; p.split_g = d->split_out; }
    p.kw = kw; p.taps = kh * kw; p.pad = pad; p.dil = dil; p.stride = stride;
    p.bias = bias; p.res = residual; p.out = out;
    p.out_sb = d->out_sb; p.out_sp = d->out_sp;
    p.relu = d->relu; p.deconv = d->deconv2x2; p.out_f32 = d->out_dtype == TDRN_F32; p.pool = d->pool2x2;
    p.f16_in = f16_in; p.f16_out = f16_out;
    TDRN_REQUIRE(!d->deconv2x2 || d->Cout % 16 == 0, "tdrn_conv2d_tc: deconv needs Cout %% 16 == 0");
    if (p.pool) {
        if (d->deconv2x2 || residual || p.W % 16 != 0 || p.H % 8 != 0) {
            set_error("tdrn_conv2d_tc: fused 2x2 max-pool needs W %% 16 == 0, H %% 8 == 0, no residual/deconv (got %dx%d)", p.H, p.W);
            return TDRN_EUNSUPPORTED;
        }
        p.bw = 16; p.bh = 8; p.bn = 1;
        p.out_w = p.W / 2;
    } else {
        // element-strided boxes are limited to 256 traversed elements per dimension
        pick_box(p.B, p.H, p.W, 256 / stride, 256 / stride, p.bw, p.bh, p.bn);
        p.out_w = d->deconv2x2 ? 2 * p.W : p.W;
    }
    p.tiles_w = (p.W + p.bw - 1) / p.bw; p.tiles_h = (p.H + p.bh - 1) / p.bh;
    const int tiles_n = (p.B + p.bn - 1) / p.bn;
    p.a_bytes = (uint32_t)(p.bw * p.bh * p.bn) * 128u;
    p.m_tiles = p.tiles_w * p.tiles_h * tiles_n;
    // ragged tail: full-width single-image boxes that do not divide H (40x40 -> 13 boxes of 3 rows + 1 leftover row):
    // batch the leftover rows of several images into one tile instead of spending a 128-row tile on rr*bw pixels
    static const bool no_tail = getenv("TDRN_NO_TAIL_TILES") != nullptr;
    if (!no_tail && !p.pool && !d->deconv2x2 && stride == 1 && p.bn == 1 && p.bw == p.W && p.H % p.bh != 0 && p.H > p.bh) {
        p.qh = p.H / p.bh; p.rr = p.H - p.qh * p.bh;
        p.g2 = 128 / (p.bw * p.rr);
        if (p.g2 > p.B) p.g2 = p.B;
        if (p.g2 >= 2) {
            p.tiles_h = p.qh;                                        // regular tiles: mt = tn * qh + th
            p.nA = p.qh * p.B;
            p.m_tiles = p.nA + (p.B + p.g2 - 1) / p.g2;
            p.a_bytes2 = (uint32_t)(p.bw * p.rr * p.g2) * 128u;
        } else {
            p.rr = 0;
        }
    }

    const int n_pad16 = (p.n_total + 15) & ~15;
    // split mode: 3 + 1 accumulators (BN <= 128) or, with TDRN_X3_NACC=1, 1 + 1 (any BN).  Measured on B200 (b32 step / worst
    // max-norm error over the test-suite, bar 1e-4): one accumulator for everything 23.0 ms / 1.3e-4; 1 + 1: 13.3 ms / 1.2e-4
    // (one 704 x 704 case above the bar); 3 + 1: 17.2 ms / every case below the bar -- the default.  BN <= 128 doubles the operand
    // traffic of the wide layers, which is what the extra 4 ms are.
    static const int x3_nacc = getenv("TDRN_X3_NACC") ? atoi(getenv("TDRN_X3_NACC")) : 3;
    if (d->split3) p.split_nacc = x3_nacc == 1 ? 1 : 3;
    int BN = n_pad16 > 128 && !(d->split3 && p.split_nacc == 3) ? 256 : (n_pad16 > 64 ? 128 : 64);
    // Small maps (the 10x10 / 5x5 pyramid levels, M <= 3200 rows at b32) yield a handful of 128-row tiles: with the
    // widest N tile only 7..25 SMs would stream the whole weight tensor.  Narrower N tiles put 2-4x more SMs to
    // work (the A re-reads this costs are tiny at these sizes).
    if (!g_num_sms) {
        int dev = 0;
        TDRN_CUDA(cudaGetDevice(&dev));
        TDRN_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    while (BN > 64 && p.m_tiles * ((n_pad16 + BN - 1) / BN) * 2 <= g_num_sms) BN >>= 1;
    p.n_tiles = (n_pad16 + BN - 1) / BN;

    {   // small maps with a long K loop: split K over a cluster (conv_splitk_kernel).  S is a function of the LAYER's shape only.
        // OPT-IN (TDRN_SPLITK=1, read per call so that tests can switch it).  Measured r02n (b32, two steps in flight): the summed
        // conv time drops (2.662 -> 2.616 ms) but the STEP gets slower (2.728 -> 2.780 ms), and batch-1 MobileNet latency too (0.531
        // -> 0.567 ms): a cluster launch costs ~3 us more than a plain one and its 8 CTAs need 8 free SMs of one GPC at once, which
        // stalls behind the persistent trunk kernels of the other step in flight.
        const char *sk_env = getenv("TDRN_SPLITK");
        const bool no_splitk = !(sk_env && sk_env[0] == '1');
        const int num_kb = p.taps * (p.Cin >> 6);
        const int n_tiles64 = (n_pad16 + SK_BN - 1) / SK_BN;
        // Measured r02l (b32): splitting pays where the cluster grid still fits one wave of 148 SMs -- the Cout <= 64 heads on the
        // 10x10 / 16x16 maps (arm_loc.2: 0.063 -> 0.043 ms) and everything on the 5x5 / 8x8 maps (512 -> 256: 0.038 -> 0.023) -- and
        // loses where it does not (256 -> 256 @10x10 as 400 one-tile CTAs: 0.024 -> 0.056).  The rule may only look at the layer.
        int S = 1;
        if (!no_splitk && !d->split3 && !p.pool && num_kb >= 16 && !f16_in && !f16_out) {
            if (p.H * p.W <= 64) { S = 8; while (S > 1 && (num_kb / S < 8 || n_tiles64 * S > 16)) S >>= 1; }
            else if (p.H * p.W <= 256 && n_tiles64 == 1) { S = 8; while (S > 1 && num_kb / S < 8) S >>= 1; }
        }
        p.splitk = S;
        if (S > 1) { BN = SK_BN; p.n_tiles = n_tiles64; }
    }
    CUtensorMap tmA, tmA2, tmB;
    bool use_cluster = false;
    {
        const uint64_t dims[4] = {(uint64_t)cin_mem, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
        const uint64_t str[3] = {(uint64_t)cin_mem * 2, (uint64_t)d->W * cin_mem * 2, (uint64_t)d->H * d->W * cin_mem * 2};
        // with element stride s TMA loads ceil(box / s) elements: box = n * s loads exactly n
        const uint32_t box[4] = {64, (uint32_t)(p.bw * stride), (uint32_t)(p.bh * stride), (uint32_t)p.bn};
        const uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
        int rc = make_tmap_bf16(&tmA, in, 4, dims, str, box, es);
        if (rc) return rc;
        tmA2 = tmA;
        if (p.rr) {
            const uint32_t box2[4] = {64, (uint32_t)p.bw, (uint32_t)p.rr, (uint32_t)p.g2};
            rc = make_tmap_bf16(&tmA2, in, 4, dims, str, box2, es);
            if (rc) return rc;
        }
    }
    {
        const uint64_t K = (uint64_t)p.taps * p.Cin * (d->split3 ? 2 : 1);       // split mode: taps x [W_hi | W_lo]
        const uint64_t dims[2] = {K, (uint64_t)n_pad16};
        const uint64_t str[1] = {K * 2};
        const uint32_t b_rows = (uint32_t)(n_pad16 < BN ? n_pad16 : BN);
        // cluster pairs (two M tiles share every weight box, each CTA fetches and multicasts half of it): only when
        // there are M-tile pairs, and the half box stays aligned to the 8-row swizzle atom
        // Measured on B200 (profiles/r01b_*): no gain -- these layers are bound by tile fill and wave quantisation, not by
        // L2->SM weight traffic -- so the cluster path is opt-in (TDRN_CLUSTER=1) and kept as a tested option.
        static const bool want_cluster = getenv("TDRN_CLUSTER") != nullptr;
        use_cluster = want_cluster && !d->split3 && p.splitk == 1 && p.m_tiles >= 2 && (b_rows % 16u) == 0;
        const uint32_t box[2] = {64, use_cluster ? b_rows / 2 : b_rows};
        p.b_bytes = b_rows * 128u;
        int rc = make_tmap_bf16(&tmB, weight, 2, dims, str, box, nullptr);
        if (rc) return rc;
    }
    {   // opt-in experiment (TDRN_L2_PREFETCH=1): measured slower on B200 (3.13 vs 3.06 ms per step), so it stays off
        static const bool pf = getenv("TDRN_L2_PREFETCH") != nullptr;
        p.l2_prefetch = pf && p.m_tiles * p.n_tiles <= 2 * g_num_sms;
    }
    {   // wide 1x1 layers with few k-blocks (the per-tap projection of the deformable heads: Cin 256 -> 2720; MobileNet's
        // pointwise convs) are bound by operand traffic and by the depth of the TMA ring, not by the MMAs: keep the weight
        // boxes of an N tile resident and give the rest of the ring to activation boxes.  Measured (b32, 256 -> 2720 @40x40):
        // 0.131 ms without, 0.124 with resident weights, 0.096 with the TMA-store epilogue on top = cuBLAS (0.093) and 77 % of
        // a pure 278 MB memset (0.074 ms: HBM write bandwidth is the bound); BN = 128 with a two-tile ring is slower (0.157)
        static const bool no_res = getenv("TDRN_NO_RESIDENT_B") != nullptr;
        const int num_kb = p.taps * (p.Cin >> 6);
        const int ring_boxes = (192 * 1024 - num_kb * BN * 128) / (128 * 128);
        p.b_resident = !no_res && !d->split3 && p.splitk == 1 && !use_cluster && ring_boxes >= 4 && ring_boxes >= num_kb && p.m_tiles * p.n_tiles >= 4 * g_num_sms;
        p.a_stages = ring_boxes < 8 ? ring_boxes : 8;
    }
    CUtensorMap tmO = tmA, tmO2 = tmA;
    {   // TMA-store epilogue: plain contiguous NHWC bf16 output (no pool / pixel shuffle / residual), 16-byte aligned rows
        static const bool no_tma_out = getenv("TDRN_NO_TMA_STORE") != nullptr;
        p.tma_out = !no_tma_out && !d->split3 && p.splitk == 1 && !d->deconv2x2 && !p.pool && !p.out_f32 && !residual && d->Cout % 8 == 0 &&
                    d->out_sp == d->Cout && d->out_sb == (long long)p.H * p.W * d->Cout && ((uintptr_t)out & 15) == 0;
        if (!no_tma_out && d->split3 && d->split_out == d->Cout && d->Cout % 64 == 0) {
            // split mode writing the whole (hi | lo) operand of the next layer, [B,Ho,Wo,2*Cout] contiguous (optionally pooled)
            const int Ho = p.pool ? p.H / 2 : p.H, Wo = p.pool ? p.W / 2 : p.W;
            if (d->out_sb == (long long)Ho * Wo * 2 * d->Cout) {
                p.tma_out = 1;
                const uint64_t dims[4] = {(uint64_t)2 * d->Cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)p.B};
                const uint64_t str[3] = {(uint64_t)d->Cout * 4, (uint64_t)Wo * d->Cout * 4, (uint64_t)Ho * Wo * d->Cout * 4};
                const uint32_t box[4] = {64, (uint32_t)(p.pool ? p.bw / 2 : p.bw), (uint32_t)(p.pool ? p.bh / 2 : p.bh), (uint32_t)p.bn};
                int rc = make_tmap_bf16(&tmO, out, 4, dims, str, box, nullptr);
                if (rc) return rc;
                tmO2 = tmO;
                if (p.rr) {
                    const uint32_t box2[4] = {64, (uint32_t)p.bw, (uint32_t)p.rr, (uint32_t)p.g2};
                    rc = make_tmap_bf16(&tmO2, out, 4, dims, str, box2, nullptr);
                    if (rc) return rc;
                }
            }
        } else if (p.tma_out) {
            const uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.B};
            const uint64_t str[3] = {(uint64_t)d->Cout * 2, (uint64_t)p.W * d->Cout * 2, (uint64_t)p.H * p.W * d->Cout * 2};
            const uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
            int rc = make_tmap_bf16(&tmO, out, 4, dims, str, box, nullptr);
            if (rc) return rc;
            tmO2 = tmO;
            if (p.rr) {
                const uint32_t box2[4] = {64, (uint32_t)p.bw, (uint32_t)p.rr, (uint32_t)p.g2};
                rc = make_tmap_bf16(&tmO2, out, 4, dims, str, box2, nullptr);
                if (rc) return rc;
            }
        }
    }
    {   // long-K layers with enough M tiles: pair two M tiles per weight box (see the kernel's MT parameter)
        static const bool no_mt2 = getenv("TDRN_NO_MT2") != nullptr;
        static const int mt2_min = getenv("TDRN_MT2_MIN") ? atoi(getenv("TDRN_MT2_MIN")) : 70;   // paired units per 100 SMs from which pairing pays
        const int num_kb = p.taps * (p.Cin >> 6);
        const int units2 = ((p.m_tiles + 1) / 2) * p.n_tiles;
        // r03: pairing must not cost a round of tiles: 512 -> 256 @40x40 at b32 is 427 tiles = 3 rounds on 148 SMs unpaired but 214
        // units = 2 rounds of TWO tiles paired (TDRN_MT2_ROUNDS=0 switches the test off)
        static const bool rounds_rule = !(getenv("TDRN_MT2_ROUNDS") && getenv("TDRN_MT2_ROUNDS")[0] == '0');
        const int tiles1 = p.m_tiles * p.n_tiles;
        const bool rounds_ok = !rounds_rule || 2 * ((units2 + g_num_sms - 1) / g_num_sms) <= (tiles1 + g_num_sms - 1) / g_num_sms;
        p.mt2 = !no_mt2 && !d->split3 && BN == 256 && p.tma_out && !p.b_resident && !use_cluster && num_kb >= 64 && units2 * 100 >= g_num_sms * mt2_min &&
                rounds_ok;   // measured: K = 2304 (36 k-blocks) loses 5 %, K = 4608 gains 5-10 %
    }
    {   // r03: N tiles of the same M tile next to each other in the tile walk.  Measured on the MobileNet pointwise convs at b64
        // (512 -> 512 @40x40: a 105 MB input, two N tiles): with the N-tile-major walk the second N tile re-read the whole input from
        // DRAM half a kernel later (ncu: 208.6 MB read for a 105 MB input).  TDRN_MT_MAJOR=0 / 1 forces the walk (read per call).
        const char *e = getenv("TDRN_MT_MAJOR");
        p.mt_major = e ? (e[0] == '1') : (p.n_tiles >= 2 && !p.b_resident && p.taps == 1);
        if (use_cluster || p.splitk > 1) p.mt_major = 0;
    }
    cudaStream_t st = as_stream(stream);
    {   // development aid (TDRN_TC_VERBOSE=1): which tiling / variant a layer gets
        static const bool verbose = getenv("TDRN_TC_VERBOSE") != nullptr;
        if (verbose)
            fprintf(stderr, "conv_tc %dx%d k%d @%dx%d b%d: box %dx%dx%d, m_tiles %d (tail rr %d g2 %d), BN %d x %d, kb %d, mt2 %d cluster %d resident %d tma_out %d splitk %d mt_major %d\n",
                    d->Cin, d->Cout, kh, p.H, p.W, p.B, p.bw, p.bh, p.bn, p.m_tiles, p.rr, p.g2, BN, p.n_tiles, p.taps * (p.Cin >> 6), p.mt2,
                    (int)use_cluster, p.b_resident, p.tma_out, p.splitk, p.mt_major);
    }
    if (p.splitk > 1) {
        TDRN_CUDA(cudaFuncSetAttribute(conv_splitk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM_BYTES));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(p.m_tiles * p.n_tiles * p.splitk));
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = SK_SMEM_BYTES;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)p.splitk; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        TDRN_CUDA(cudaLaunchKernelEx(&cfg, conv_splitk_kernel, tmA, tmA2, tmB, p));
        count_launch();
        return TDRN_OK;
    }
    if (BN == 256) return launch_tc<256>(tmA, tmA2, tmB, tmO, tmO2, p, use_cluster, st);
    if (BN == 128) return launch_tc<128>(tmA, tmA2, tmB, tmO, tmO2, p, use_cluster, st);
    return launch_tc<64>(tmA, tmA2, tmB, tmO, tmO2, p, use_cluster, st);
}
