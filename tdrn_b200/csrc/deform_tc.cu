// deform_tc.cu -- fused deformable detection head on tcgen05 (bf16 in, fp32 accumulate).
//
// Replaces, for the ODM heads of DualRefineDet (model/dualrefinedet_vggbn.py:180-197) and the
// temporal heads of TDRN (model/ssd4scale_vgg.py:106-109), the reference's per-sample loop
//   zero(output); deformable_im2col -> columns[K, HW] fp32 in HBM; SGEMM     (deform_conv_cuda.c:157-193)
// executed once for `loc` and once again -- same sampling -- for `conf`.  Here:
//   * loc and conf weights are concatenated along N (12 + 3*C rows), so every sampled value feeds
//     all outputs of its pixel;
//   * the bilinear-sampled im2col tile (sampler rules of deform_conv_cuda_kernel.cu:16-51,195-203)
//     is produced by 16 warps straight into shared memory in the 128B-swizzled K-major UMMA layout
//     (each corner read is a 16-byte channels-last vector, the 8 lanes of a quarter-warp cover one
//     128-byte line) and is
//     consumed in place by tcgen05.mma: the column buffer never exists in HBM;
//   * the optional 5x5 "multihead" (l(ob,f) + l2(ob,f2), :182-183) simply continues the K loop into
//     the same TMEM accumulator;
//   * the epilogue stages the 128 x N fp32 tile through shared memory, applies the class softmax
//     (nn.Softmax(dim=1) on view(-1, C), :196) and writes loc [B,P,4] / conf [B,P,C] rows coalesced
//     at their prior offsets (the reference's permute(0,2,3,1).contiguous().view + cat).
//
// Measured dead ends (B200, level 0 = 0.34 ms): blending in packed bf16 (HFMA2.BF16) is only 13 % faster and adds
// 50 % error; a register double-buffered software pipeline (corner loads of k-block i+1 issued before k-block i is
// blended) spills at the 113-register cap of 576 threads and runs 60 % slower.  The kernel sits at 65 % issue
// utilisation with ~215 instructions per warp per k-block, 104 of them the fp32 blend.
//
// Warp roles (576 threads): warps 0-15 A producers (then epilogue), warp 16 weight TMA, warp 17 TMEM
// allocator + MMA issuer.
#include "tc_common.cuh"
#include <stdlib.h>

namespace tdrn {
namespace tc {

struct DeformP {
    const __nv_bfloat16 *feat;       // [B,H,W,Cin]
    const float *off[2];             // [B,H,W,dg*2*taps]
    int B, H, W, Cin, dg, cpg;
    int k[2], pad[2], taps[2];       // head 1 / head 2 (taps[1] == 0: absent)
    int M;                           // B*H*W
    int n_total, n_pad16, C;         // 12 + 3*C
    int P, prior_off, softmax;
    float *loc_out, *conf_out;
    uint32_t b_bytes;
    int geo_per_row;                 // (taps0 + taps1) * dg geometry entries per tile row
    int stages;                      // smem ring depth (3, or 2 when the geometry cache is large: dg = 8)
};

template <int NMAX> struct DfCfg {
    static constexpr int A_BYTES = 128 * 128;
    static constexpr int B_BYTES = NMAX * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = 3;
    static constexpr int TMEM_COLS = NMAX;
    static constexpr int SMEM_STAGES_BYTES = STAGES * STAGE_BYTES + 1024;   // + geometry cache (run-time size)
};

// Packed-pair fp32 arithmetic (Blackwell FFMA2 / FMUL2): one instruction blends both channels of a bf16x2 word.
__device__ __forceinline__ uint64_t pair_of(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
// bf16x2 word -> (lo, hi) fp32 pair: a bf16 is the upper half of the fp32 with the same value
__device__ __forceinline__ uint64_t unpack_bf16x2(uint32_t a)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a << 16), "r"(a & 0xffff0000u));
    return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b)
{
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint32_t blend2s(uint32_t a, uint32_t b, uint32_t c, uint32_t d, float w1, float w2, float w3, float w4)
{
    float lo = w1 * __uint_as_float(a << 16), hi = w1 * __uint_as_float(a & 0xffff0000u);
    lo = fmaf(w2, __uint_as_float(b << 16), lo); hi = fmaf(w2, __uint_as_float(b & 0xffff0000u), hi);
    lo = fmaf(w3, __uint_as_float(c << 16), lo); hi = fmaf(w3, __uint_as_float(c & 0xffff0000u), hi);
    lo = fmaf(w4, __uint_as_float(d << 16), lo); hi = fmaf(w4, __uint_as_float(d & 0xffff0000u), hi);
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// Bilinear blend of two packed bf16 channels from the 4 corners (.cu:49), fp32 arithmetic, one bf16x2 result.
__device__ __forceinline__ uint32_t blend2(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint64_t w1, uint64_t w2, uint64_t w3, uint64_t w4)
{
    uint64_t acc = mul2(w1, unpack_bf16x2(a));
    acc = fma2(w2, unpack_bf16x2(b), acc);
    acc = fma2(w3, unpack_bf16x2(c), acc);
    acc = fma2(w4, unpack_bf16x2(d), acc);
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc));
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

constexpr int DF_PRODUCER_WARPS = 16;
constexpr int DF_THREADS = (DF_PRODUCER_WARPS + 2) * 32;

template <int NMAX, bool F2>
__global__ void __launch_bounds__(DF_THREADS, 1) deform_head_kernel(const __grid_constant__ CUtensorMap tmB0,
                                                                 const __grid_constant__ CUtensorMap tmB1, const DeformP p)
{
    using Cfg = DfCfg<NMAX>;
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[Cfg::STAGES];
    __shared__ __align__(8) uint64_t empty_bar[Cfg::STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_s;

    uint8_t *tiles = (uint8_t *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128;
    const int cblocks = p.Cin >> 6;
    const int kb_head0 = p.taps[0] * cblocks;
    const int num_kb = kb_head0 + p.taps[1] * cblocks;
    const int HW = p.H * p.W;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmB0);
        if (p.taps[1]) tma_prefetch_desc(&tmB1);
#pragma unroll
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full_bar[s], DF_PRODUCER_WARPS + 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == DF_PRODUCER_WARPS + 1) tmem_alloc(&tmem_base_s, Cfg::TMEM_COLS);

    // ---- per-tile sampling geometry cache: one 16-byte entry per (row, head, tap, deformable group) ----------
    //   x = byte offset of the (low,low) corner pixel in feat | dx << 30 | dy << 31
    //   y = in * (1 - lh), z = in * lh, w = lw      (in = 0 when the sample lies outside the map, .cu:197)
    uint4 *geo = (uint4 *)(tiles + p.stages * Cfg::STAGE_BYTES);
    for (int e = threadIdx.x; e < 128 * p.geo_per_row; e += DF_THREADS) {
        const int r = e / p.geo_per_row, gi = e - r * p.geo_per_row;
        const int head = gi >= p.taps[0] * p.dg ? 1 : 0;
        const int gl = head ? gi - p.taps[0] * p.dg : gi;
        const int tap = gl / p.dg, g = gl - tap * p.dg;
        const int m = m0 + r;
        const bool rvalid = m < p.M;
        const int mm = rvalid ? m : 0;
        const int rb = mm / HW, rem = mm - rb * HW;
        const int ry = rem / p.W, rx = rem - ry * p.W;
        const int kk = p.k[head];
        const int ti = tap / kk, tj = tap - ti * kk;
        const int oc = p.dg * 2 * p.taps[head];
        const float *op = p.off[head] + ((long long)rb * HW + rem) * oc + (g * 2 * p.taps[head] + 2 * tap);
        const float oh = rvalid ? __ldg(op) : 0.f, ow = rvalid ? __ldg(op + 1) : 0.f;
        const int y0 = ry - p.pad[head], x0 = rx - p.pad[head];                          // stride 1
        const float h_im = (float)(y0 + ti) + oh, w_im = (float)(x0 + tj) + ow;          // .cu:195-196 (dilation 1)
        const bool inside = rvalid && h_im >= 0.f && w_im >= 0.f && h_im < (float)p.H && w_im < (float)p.W;   // .cu:197
        float h = (float)ti + oh, w = (float)tj + ow;                                     // map_h / map_w .cu:198-199
        const int cur_h = p.H - y0, cur_w = p.W - x0;
        int h_low = (int)floorf(h), w_low = (int)floorf(w), h_high, w_high;               // .cu:21-37
        if (h_low >= cur_h - 1) { h_high = h_low = cur_h - 1; h = (float)h_low; } else { h_high = h_low + 1; }
        if (w_low >= cur_w - 1) { w_high = w_low = cur_w - 1; w = (float)w_low; } else { w_high = w_low + 1; }
        const float lh = h - (float)h_low, lw = w - (float)w_low;
        const int ya = min(max(y0 + h_low, 0), p.H - 1), yb = min(max(y0 + h_high, 0), p.H - 1);
        const int xa = min(max(x0 + w_low, 0), p.W - 1), xb = min(max(x0 + w_high, 0), p.W - 1);
        const unsigned base = (unsigned)(rb * HW + ya * p.W + xa) * (unsigned)(p.Cin * 2);
        uint4 ent;
        ent.x = inside ? (base | ((unsigned)(xb - xa) << 30) | ((unsigned)(yb - ya) << 31)) : 0u;
        ent.y = __float_as_uint(inside ? 1.f - lh : 0.f);
        ent.z = __float_as_uint(inside ? lh : 0.f);
        ent.w = __float_as_uint(lw);
        geo[e] = ent;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp < DF_PRODUCER_WARPS) {
        // ===================== A producers: bilinear-sampled im2col straight into smem =====================
        // k-blocks run channel-block-major (cb, tap): all taps of one 64-channel slab are sampled back to
        // back, so the slab's footprint (tile rows +- kernel radius +- offsets, ~50 KB) stays L1-resident.
        // 16 warps x 32 lanes; per k-block a lane produces the 16-byte channel chunk `chunk` of two tile rows.  The
        // 8 lanes of a quarter-warp cover one row's whole 128-byte line, so every corner read is ONE L1 wavefront
        // per quarter-warp (a full line) and the swizzled 16-byte stores are bank-conflict free.
        const int chunk = lane & 7, sub = lane >> 3;
        const int blocks_per_group = p.cpg >> 6;
        const int r0 = warp * 8 + sub, r1 = r0 + 4;
        const uint32_t geo_s = smem_u32(geo);
        const uint32_t geo_row0 = geo_s + (uint32_t)r0 * (uint32_t)p.geo_per_row * 16u;
        const uint32_t geo_row1 = geo_s + (uint32_t)r1 * (uint32_t)p.geo_per_row * 16u;
        const uint32_t tiles_s = smem_u32(tiles);
        const uint32_t st_off0 = sw128_offset(r0, chunk), st_off1 = sw128_offset(r1, chunk);
        const char *fbytes = (const char *)p.feat;
        const uint32_t cin2 = (uint32_t)p.Cin * 2u, row2 = (uint32_t)p.W * cin2;
        uint32_t s = 0, ph = 1;                                  // ring stage and the parity empty_bar[s] must have passed
        // Flat walk over the k-blocks (head, channel block, tap).  The corner loads of row-item i of the NEXT k-block are
        // issued right after item i of the current k-block has been blended (same registers, no double buffer), so
        // their L1/L2 latency overlaps the other item's blend, the fence/arrive and the ring wait instead of being
        // paid in full at the top of every k-block.
        const int taps0 = p.taps[0], taps1 = p.taps[1];
        const uint32_t gstep = (uint32_t)p.dg * 16u;
        const uint32_t ghead1 = (uint32_t)(taps0 * p.dg) * 16u;
        int head = 0, cb = 0, tap = 0;
        uint32_t goff = 0u, lane_off = (uint32_t)(chunk << 4);
        bool valid = true;
        auto advance = [&]() {                                    // -> (goff, lane_off) of the next k-block, or valid = false
            const int taps = head ? taps1 : taps0;
            if (++tap < taps) { goff += gstep; return; }
            tap = 0;
            if (++cb == cblocks) { cb = 0; ++head; if (head == 2 || taps1 == 0) { valid = false; return; } }
            goff = (head ? ghead1 : 0u) + (uint32_t)(cb / blocks_per_group) * 16u;
            lane_off = (uint32_t)(cb << 7) + (uint32_t)(chunk << 4);
        };
        uint4 cv[2][4];                                           // 2 rows x 4 corners: 8 independent 16-byte loads in flight
        float f[2][4];
        auto load_item = [&](int i) {
            uint32_t e0, e1, e2, e3;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(e0), "=r"(e1), "=r"(e2), "=r"(e3) : "r"((i ? geo_row1 : geo_row0) + goff));
            const uint32_t oa = (e0 & 0x3fffffffu) + lane_off;
            const uint32_t dxb = (uint32_t)((int32_t)(e0 << 1) >> 31) & cin2;
            const uint32_t dyb = (uint32_t)((int32_t)e0 >> 31) & row2;
            const uint32_t oc = oa + dyb;
            cv[i][0] = __ldg((const uint4 *)(fbytes + oa));
            cv[i][1] = __ldg((const uint4 *)(fbytes + (oa + dxb)));
            cv[i][2] = __ldg((const uint4 *)(fbytes + oc));
            cv[i][3] = __ldg((const uint4 *)(fbytes + (oc + dxb)));
            const float hh = __uint_as_float(e1), lh = __uint_as_float(e2), lw = __uint_as_float(e3), hw = 1.f - lw;
            f[i][0] = hh * hw; f[i][1] = hh * lw; f[i][2] = lh * hw; f[i][3] = lh * lw;      // .cu:47 (x inside flag)
        };
        auto blend_item = [&](int i, uint32_t sa) {
            uint32_t o0, o1, o2, o3;                             // .cu:49, two channels per 32-bit lane
            if (F2) {
                const uint64_t w1 = pair_of(f[i][0], f[i][0]), w2 = pair_of(f[i][1], f[i][1]);
                const uint64_t w3 = pair_of(f[i][2], f[i][2]), w4 = pair_of(f[i][3], f[i][3]);
                o0 = blend2(cv[i][0].x, cv[i][1].x, cv[i][2].x, cv[i][3].x, w1, w2, w3, w4);
                o1 = blend2(cv[i][0].y, cv[i][1].y, cv[i][2].y, cv[i][3].y, w1, w2, w3, w4);
                o2 = blend2(cv[i][0].z, cv[i][1].z, cv[i][2].z, cv[i][3].z, w1, w2, w3, w4);
                o3 = blend2(cv[i][0].w, cv[i][1].w, cv[i][2].w, cv[i][3].w, w1, w2, w3, w4);
            } else {
                o0 = blend2s(cv[i][0].x, cv[i][1].x, cv[i][2].x, cv[i][3].x, f[i][0], f[i][1], f[i][2], f[i][3]);
                o1 = blend2s(cv[i][0].y, cv[i][1].y, cv[i][2].y, cv[i][3].y, f[i][0], f[i][1], f[i][2], f[i][3]);
                o2 = blend2s(cv[i][0].z, cv[i][1].z, cv[i][2].z, cv[i][3].z, f[i][0], f[i][1], f[i][2], f[i][3]);
                o3 = blend2s(cv[i][0].w, cv[i][1].w, cv[i][2].w, cv[i][3].w, f[i][0], f[i][1], f[i][2], f[i][3]);
            }
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sa + (i ? st_off1 : st_off0)), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
        };
        load_item(0);
        load_item(1);
        while (valid) {
            mbar_wait(&empty_bar[s], ph);
            const uint32_t sa = tiles_s + s * (uint32_t)Cfg::STAGE_BYTES;
            advance();                                            // (goff, lane_off) now describe the NEXT k-block
            blend_item(0, sa);
            if (valid) load_item(0);
            blend_item(1, sa);
            if (valid) load_item(1);
            fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the tensor-core (async) proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[s]);
            if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; }
        }
    } else if (warp == DF_PRODUCER_WARPS) {
        // ===================== weight (B operand) TMA producer =====================
        if (elect_one()) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % p.stages;
                const uint32_t ph = (uint32_t)(kb / p.stages) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                uint8_t *sb = tiles + s * Cfg::STAGE_BYTES + Cfg::A_BYTES;
                mbar_expect_tx(&full_bar[s], p.b_bytes);
                if (kb < kb_head0) tma_load_2d(sb, &tmB0, &full_bar[s], kb * 64, 0);
                else tma_load_2d(sb, &tmB1, &full_bar[s], (kb - kb_head0) * 64, 0);
            }
        }
        __syncwarp();
    } else {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_bf16(128, p.n_pad16);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % p.stages;
                const uint32_t ph = (uint32_t)(kb / p.stages) & 1u;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(tiles + s * Cfg::STAGE_BYTES);
                const uint64_t adesc = umma_desc_sw128(sa);
                const uint64_t bdesc = umma_desc_sw128(sa + Cfg::A_BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&tmem_full_bar);
        }
        __syncwarp();
    }

    // ===================== epilogue =====================
    // All MMAs (hence all smem operand reads) are complete once tmem_full fires; the stage buffers are
    // reused as a [128][n_pad16+1] fp32 staging tile.
    // (The mbarrier chain producer-arrive -> MMA wait -> tcgen05.commit -> tmem_full already orders the producers' A-tile
    // stores before these staging stores; the CTA barrier makes that order visible to compute-sanitizer's racecheck, which
    // does not follow mbarriers -- it reported the pair as a write-after-write hazard -- and costs one barrier per tile.)
    __syncthreads();
    float *stg = (float *)tiles;
    const int lds = p.n_pad16 + 1;
    if (warp < 4) {
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        const int r = warp * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < p.n_pad16; c0 += 16) {
            float v[16];
            tmem_ld16(trow + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) stg[r * lds + c0 + j] = v[j];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == DF_PRODUCER_WARPS + 1) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }

    const int C = p.C;
    if (p.softmax) {
        for (int t = threadIdx.x; t < 128 * 3; t += DF_THREADS) {
            const int r = t / 3, a = t - r * 3;
            float *row = stg + r * lds + 12 + a * C;
            float mx = -INFINITY;
            for (int c = 0; c < C; ++c) mx = fmaxf(mx, row[c]);
            float sum = 0.f;
            for (int c = 0; c < C; ++c) { const float e = expf(row[c] - mx); row[c] = e; sum += e; }
            const float inv = 1.f / sum;
            for (int c = 0; c < C; ++c) row[c] *= inv;
        }
        __syncthreads();
    }
    const int rows = min(128, p.M - m0);
    for (int e = threadIdx.x; e < rows * 12; e += DF_THREADS) {
        const int r = e / 12, j = e - r * 12;
        const int m = m0 + r, b = m / HW, pix = m - b * HW;
        p.loc_out[((long long)b * p.P + p.prior_off + pix * 3) * 4 + j] = stg[r * lds + j];
    }
    const int nc = 3 * C;
    for (int e = threadIdx.x; e < rows * nc; e += DF_THREADS) {
        const int r = e / nc, j = e - r * nc;
        const int m = m0 + r, b = m / HW, pix = m - b * HW;
        p.conf_out[((long long)b * p.P + p.prior_off + pix * 3) * C + j] = stg[r * lds + 12 + j];
    }
}

template <int NMAX, bool F2>
static int launch_deform(const CUtensorMap &t0, const CUtensorMap &t1, const DeformP &p, cudaStream_t st)
{
    using Cfg = DfCfg<NMAX>;
    const size_t geo_bytes = (size_t)128 * p.geo_per_row * 16;
    DeformP q = p;
    q.stages = Cfg::STAGES;
    size_t smem = (size_t)q.stages * Cfg::STAGE_BYTES + 1024 + geo_bytes;
    if (smem > 227 * 1024) { q.stages = 2; smem = (size_t)q.stages * Cfg::STAGE_BYTES + 1024 + geo_bytes; }
    if (smem > 227 * 1024) {
        set_error("tdrn_deform_head: geometry cache does not fit (%zu bytes of shared memory needed)", smem);
        return TDRN_EUNSUPPORTED;
    }
    TDRN_CUDA(cudaFuncSetAttribute(deform_head_kernel<NMAX, F2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    deform_head_kernel<NMAX, F2><<<(p.M + 127) / 128, DF_THREADS, smem, st>>>(t0, t1, q);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

}  // namespace tc
}  // namespace tdrn

using namespace tdrn;
using namespace tdrn::tc;

extern "C" int tdrn_deform_head(const tdrn_deform_head_desc *d, const void *feat, const float *offsets,
                                const void *weight, const float *offsets2, const void *weight2,
                                float *loc_out, float *conf_out, tdrn_stream_t stream)
{
    TDRN_REQUIRE(d && feat && offsets && weight && loc_out && conf_out, "tdrn_deform_head: null argument");
    TDRN_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->num_classes > 0 && d->dg > 0, "tdrn_deform_head: bad shape");
    TDRN_REQUIRE(d->kh > 0 && 2 * d->pad == d->kh - 1, "tdrn_deform_head: head 1 must be 'same' (2*pad == k-1)");
    TDRN_REQUIRE(d->kh2 == 0 || (2 * d->pad2 == d->kh2 - 1 && offsets2 && weight2), "tdrn_deform_head: bad second head");
    TDRN_REQUIRE(d->Cin % d->dg == 0, "input channels must divide deformable group size");
    const int cpg = d->Cin / d->dg;
    DeformP p{};
    p.n_total = 12 + 3 * d->num_classes;
    p.n_pad16 = (p.n_total + 15) & ~15;
    if (d->Cin % 64 != 0 || cpg % 64 != 0 || p.n_pad16 > 256) {
        set_error("tdrn_deform_head: needs Cin %% 64 == 0, (Cin/dg) %% 64 == 0 and 12+3*C <= 256 (got Cin=%d dg=%d C=%d)",
                  d->Cin, d->dg, d->num_classes);
        return TDRN_EUNSUPPORTED;
    }
    p.feat = (const __nv_bfloat16 *)feat; p.off[0] = offsets; p.off[1] = offsets2;
    p.B = d->B; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.dg = d->dg; p.cpg = cpg;
    p.k[0] = d->kh; p.pad[0] = d->pad; p.taps[0] = d->kh * d->kh;
    p.k[1] = d->kh2 ? d->kh2 : 1; p.pad[1] = d->pad2; p.taps[1] = d->kh2 * d->kh2;
    p.M = d->B * d->H * d->W; p.C = d->num_classes; p.P = d->P; p.prior_off = d->prior_off; p.softmax = d->softmax;
    p.loc_out = loc_out; p.conf_out = conf_out;
    p.geo_per_row = (p.taps[0] + p.taps[1]) * d->dg;
    const int nmax = p.n_pad16 > 128 ? 256 : 128;
    const int b_rows = p.n_pad16;
    p.b_bytes = (uint32_t)b_rows * 128u;

    CUtensorMap t0, t1;
    for (int h = 0; h < 2; ++h) {
        if (h == 1 && !d->kh2) { t1 = t0; break; }
        const uint64_t K = (uint64_t)p.taps[h] * d->Cin;
        const uint64_t dims[2] = {K, (uint64_t)p.n_pad16};
        const uint64_t str[1] = {K * 2};
        const uint32_t box[2] = {64, (uint32_t)b_rows};
        int rc = make_tmap_bf16(h ? &t1 : &t0, h ? weight2 : weight, 2, dims, str, box, nullptr);
        if (rc) return rc;
    }
    cudaStream_t st = as_stream(stream);
    static const bool f2 = getenv("TDRN_DEFORM_SCALAR_BLEND") == nullptr;   // default: packed fp32x2 (FFMA2) blend
    if (f2) return nmax == 256 ? launch_deform<256, true>(t0, t1, p, st) : launch_deform<128, true>(t0, t1, p, st);
    return nmax == 256 ? launch_deform<256, false>(t0, t1, p, st) : launch_deform<128, false>(t0, t1, p, st);
}
