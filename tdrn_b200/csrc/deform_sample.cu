// deform_sample.cu -- second half of the "project, then sample" deformable detection head.
//
// Bilinear sampling is linear in the feature map and acts on every channel alike, so for one deformable group
//   out[p,o] = sum_tap sum_c W[o,c,tap] * S(X[c]; p,tap)  =  sum_tap S( (W_tap X)[o]; p,tap )
// (S = the reference sampler, deform_conv_cuda_kernel.cu:16-51,195-203; the sum over c is the SGEMM of
// deform_conv_cuda.c:181-190).  When the head is narrow (12 + 3C = 75 outputs against Cin = 256 at VOC-21) it is
// 3.4x cheaper to project FIRST -- one dense 1x1 tcgen05 GEMM  Y[p, tap, o] = sum_c W[o,c,tap] X[p,c]
// (tdrn_conv2d_tc, Cout = taps * n_pad) -- and to sample the 80-channel projections than to sample 256 channels
// into an im2col tile (deform_tc.cu, which stays the path for wide heads and for dg > 1).
//
// This kernel is the sampler: per output pixel and tap it reads the four corner vectors Y[corner, tap, 0..n_pad)
// (n_pad bf16 = n_pad/8 lanes x 16 bytes, contiguous), blends them in fp32 with the reference's border rules and
// accumulates over the taps of both heads (3x3 and, for "multihead", 5x5: dualrefinedet_vggbn.py:182-183) in
// registers; the epilogue applies the class softmax (:196) and writes loc [B,P,4] / conf [B,P,C] rows at their
// prior offsets (the reference's permute(0,2,3,1).contiguous().view + cat, :186-189).
//
// A CTA owns a TH x TW pixel patch so that the corner reads of neighbouring pixels hit L1; the warps sweep the taps
// together (the bytes of one tap are only touched while that tap is processed).  HBM-/L2-bound:
// algorithmic bytes per pixel = taps * n_pad * 2 (each projection is read once) + offsets + outputs.
#include "common.cuh"
#include <stdlib.h>

namespace tdrn {

struct SampleP {
    const uint4 *y;                  // [B,H,W,taps_total,n_pad] bf16 in 16-byte units
    const float *off[2];             // [B,H,W,2*taps] fp32 (dg = 1)
    int B, H, W;
    int k[2], pad[2], taps[2];
    int taps_total, n_pad, lpp, ppw; // lanes per pixel (n_pad / 8), pixels per warp (32 / lpp)
    int taps_pad;                    // taps_total rounded up to even (padding entry: weight 0)
    int TW, TH, tiles_w, tiles_h, slots;
    int n_total, C, P, prior_off, softmax;
    int f16;                         // the projections are IEEE half instead of bf16 (tdrn_deform_head_desc.split & 2)
    int merge_g;                     // > 0: the projections are (hi | lo) pairs, columns [g, 2g) are added to [0, g) before the softmax
    float *loc_out, *conf_out;
};

__device__ __forceinline__ uint64_t s_pair(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
template <bool F16>
__device__ __forceinline__ uint64_t s_unpack(uint32_t a)     // bf16x2 (or IEEE half x2) -> (lo, hi) fp32
{
    uint64_t r;
    if (F16) {
        asm("{\n\t.reg .b16 l, h;\n\t.reg .f32 fl, fh;\n\tmov.b32 {l, h}, %1;\n\tcvt.f32.f16 fl, l;\n\tcvt.f32.f16 fh, h;\n\tmov.b64 %0, {fl, fh};\n\t}"
            : "=l"(r) : "r"(a));
    } else {
        asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a << 16), "r"(a & 0xffff0000u));
    }
    return r;
}
__device__ __forceinline__ uint64_t s_fma2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
template <bool F16>
__device__ __forceinline__ void s_acc(uint64_t (&acc)[4], const uint4 v, float w)
{
    const uint64_t ww = s_pair(w, w);
    acc[0] = s_fma2(ww, s_unpack<F16>(v.x), acc[0]);
    acc[1] = s_fma2(ww, s_unpack<F16>(v.y), acc[1]);
    acc[2] = s_fma2(ww, s_unpack<F16>(v.z), acc[2]);
    acc[3] = s_fma2(ww, s_unpack<F16>(v.w), acc[3]);
}

constexpr int DS_MAX_WARPS = 16;

// One CTA's patch of one pyramid level.  `block` = index of the patch inside the level; blockDim may be larger than the level
// needs (grouped launch: the widest level decides): the extra warps own no pixel slot and only take part in the barriers.
template <bool F16>
__device__ __forceinline__ void deform_sample_body(const SampleP &p, int block)
{
    extern __shared__ uint4 ds_smem[];           // geometry [slots][taps_total], later the fp32 staging tile
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int HW = p.H * p.W;
    int t = block;
    const int tb = t / (p.tiles_w * p.tiles_h);
    t -= tb * p.tiles_w * p.tiles_h;
    const int ty0 = (t / p.tiles_w) * p.TH, tx0 = (t % p.tiles_w) * p.TW;
    const uint32_t px_stride = (uint32_t)(p.taps_total * p.lpp);           // 16-byte units per pixel of Y
    const uint32_t row_stride = (uint32_t)p.W * px_stride;

    // ---- sampling geometry, once per (pixel, tap): 16-byte entries ---------------------------------------------
    //   x = index (16-byte units) of the (low,low) corner's tap vector in Y | dx << 30 | dy << 31
    //   y = in * (1 - lh), z = in * lh, w = lw        (in = 0 when the sample lies outside the map, .cu:197)
    for (int e = threadIdx.x; e < p.slots * p.taps_pad; e += blockDim.x) {
        const int slot = e / p.taps_pad, gt = e - slot * p.taps_pad;
        const int head = gt >= p.taps[0] ? 1 : 0;
        const int tap = head ? gt - p.taps[0] : gt;
        const int sy = slot / p.TW, sx = slot - sy * p.TW;
        const int ry = ty0 + sy, rx = tx0 + sx;
        const bool rvalid = gt < p.taps_total && sy < p.TH && ry < p.H && rx < p.W;
        uint4 ent = make_uint4(0u, 0u, 0u, 0u);
        if (rvalid) {
            const int kk = p.k[head];
            const int ti = tap / kk, tj = tap - ti * kk;
            const float *op = p.off[head] + ((long long)tb * HW + ry * p.W + rx) * (2 * p.taps[head]) + 2 * tap;
            const float oh = __ldg(op), ow = __ldg(op + 1);
            const int y0 = ry - p.pad[head], x0 = rx - p.pad[head];                          // stride 1
            const float h_im = (float)(y0 + ti) + oh, w_im = (float)(x0 + tj) + ow;          // .cu:195-196 (dilation 1)
            const bool inside = h_im >= 0.f && w_im >= 0.f && h_im < (float)p.H && w_im < (float)p.W;   // .cu:197
            float h = (float)ti + oh, w = (float)tj + ow;                                     // map_h / map_w .cu:198-199
            const int cur_h = p.H - y0, cur_w = p.W - x0;
            int h_low = (int)floorf(h), w_low = (int)floorf(w), h_high, w_high;               // .cu:21-37
            if (h_low >= cur_h - 1) { h_high = h_low = cur_h - 1; h = (float)h_low; } else { h_high = h_low + 1; }
            if (w_low >= cur_w - 1) { w_high = w_low = cur_w - 1; w = (float)w_low; } else { w_high = w_low + 1; }
            const float lh = h - (float)h_low, lw = w - (float)w_low;
            const int ya = min(max(y0 + h_low, 0), p.H - 1), yb = min(max(y0 + h_high, 0), p.H - 1);
            const int xa = min(max(x0 + w_low, 0), p.W - 1), xb = min(max(x0 + w_high, 0), p.W - 1);
            if (inside) {
                const unsigned base = (unsigned)(tb * HW + ya * p.W + xa) * px_stride + (unsigned)(gt * p.lpp);
                ent.x = base | ((unsigned)(xb - xa) << 30) | ((unsigned)(yb - ya) << 31);
                ent.y = __float_as_uint(1.f - lh);
                ent.z = __float_as_uint(lh);
                ent.w = __float_as_uint(lw);
            }
        }
        ds_smem[e] = ent;
    }
    __syncthreads();

    // ---- sampling: lane = (pixel slot of this warp, 16-byte channel chunk) --------------------------------------
    const int grp = lane / p.lpp, sub = lane - grp * p.lpp;
    const bool active = grp < p.ppw;
    const int slot = warp * p.ppw + (active ? grp : 0);
    uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};
    if (active && slot < p.slots) {
        const uint4 *geo = ds_smem + slot * p.taps_pad;
        const uint4 *__restrict__ yb = p.y;                  // uniform base + 32-bit index: one IMAD.WIDE per address
        const uint32_t usub = (uint32_t)sub;
        // Measured on B200 (level 0, b32): this burst form (8 corner loads, then their 64 FMAs) runs in 0.132 ms; a rolling
        // pipeline that re-issues tap t+2's loads right after tap t is blended 0.146 ms, four taps in flight at one CTA
        // per SM (111 registers) 0.164 ms -- the kernel is not latency-bound but balanced between L1 wavefronts (58 %),
        // issue slots (55 %) and the FMA pipe (45 %).
#pragma unroll 2
        for (int gt = 0; gt < p.taps_pad; ++gt) {
            const uint4 e = geo[gt];
            const uint32_t oa = (e.x & 0x3fffffffu) + usub;
            const uint32_t dxs = (uint32_t)((int32_t)(e.x << 1) >> 31) & px_stride;
            const uint32_t oc = oa + ((uint32_t)((int32_t)e.x >> 31) & row_stride);
            // a sample outside the map (deform_conv_cuda_kernel.cu:195) and the padding taps are null entries (all weights 0,
            // hh = 1 - lh > 0 for every real one): their loads are predicated off, so that they contribute exactly 0 as in the
            // reference even if the projection they would have pointed at (Y[0]) holds an Inf / NaN
            const bool real = e.y != 0u;
            const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
            const uint4 v1 = real ? __ldg(yb + oa) : zero4;
            const uint4 v2 = real ? __ldg(yb + (oa + dxs)) : zero4;
            const uint4 v3 = real ? __ldg(yb + oc) : zero4;
            const uint4 v4 = real ? __ldg(yb + (oc + dxs)) : zero4;
            const float hh = __uint_as_float(e.y), lh = __uint_as_float(e.z), lw = __uint_as_float(e.w), hw = 1.f - lw;
            s_acc<F16>(acc, v1, hh * hw);                                                     // .cu:47-49
            s_acc<F16>(acc, v2, hh * lw);
            s_acc<F16>(acc, v3, lh * hw);
            s_acc<F16>(acc, v4, lh * lw);
        }
    }
    __syncthreads();                              // every warp is done with the geometry: reuse it as the staging tile

    // ---- epilogue: fp32 staging tile [slots][n_pad + 1], softmax, flattened stores --------------------------------
    float *stg = (float *)ds_smem;
    const int lds = p.n_pad + 1;
    if (active && slot < p.slots) {
        float *row = stg + slot * lds + sub * 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float lo, hi;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[j]));
            row[2 * j] = lo; row[2 * j + 1] = hi;
        }
    }
    __syncthreads();
    const int C = p.C;
    if (p.merge_g) {                              // fp32-accurate heads: sampled high halves + sampled low halves
        for (int e = threadIdx.x; e < p.slots * p.merge_g; e += blockDim.x) {
            const int r = e / p.merge_g, c = e - r * p.merge_g;
            stg[r * lds + c] += stg[r * lds + p.merge_g + c];
        }
        __syncthreads();
    }
    if (p.softmax) {
        for (int q = threadIdx.x; q < p.slots * 3; q += blockDim.x) {
            const int r = q / 3, a = q - r * 3;
            float *row = stg + r * lds + 12 + a * C;
            float mx = -INFINITY;
            for (int c = 0; c < C; ++c) mx = fmaxf(mx, row[c]);
            float sum = 0.f;
            for (int c = 0; c < C; ++c) { const float ev = expf(row[c] - mx); row[c] = ev; sum += ev; }
            const float inv = 1.f / sum;
            for (int c = 0; c < C; ++c) row[c] *= inv;
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < p.slots * 12; e += blockDim.x) {
        const int r = e / 12, j = e - r * 12;
        const int sy = r / p.TW, sx = r - sy * p.TW;
        const int ry = ty0 + sy, rx = tx0 + sx;
        if (sy < p.TH && ry < p.H && rx < p.W)
            p.loc_out[((long long)tb * p.P + p.prior_off + (ry * p.W + rx) * 3) * 4 + j] = stg[r * lds + j];
    }
    const int nc = 3 * C;
    for (int e = threadIdx.x; e < p.slots * nc; e += blockDim.x) {
        const int r = e / nc, j = e - r * nc;
        const int sy = r / p.TW, sx = r - sy * p.TW;
        const int ry = ty0 + sy, rx = tx0 + sx;
        if (sy < p.TH && ry < p.H && rx < p.W)
            p.conf_out[((long long)tb * p.P + p.prior_off + (ry * p.W + rx) * 3) * C + j] = stg[r * lds + 12 + j];
    }
}

template <bool F16>
__global__ void __launch_bounds__(DS_MAX_WARPS * 32, 2) deform_sample_kernel(const SampleP p)
{
    deform_sample_body<F16>(p, (int)blockIdx.x);
}

// All pyramid levels of a detector in ONE launch (the 20x20 / 10x10 / 5x5 levels are a quarter of the pixels and, launched
// on their own, 40 % of the summed sampler time: launch + tail latency of kernels that live ~20 us).
struct SampleGroup {
    int n;
    int first[TDRN_MAX_OFFSET_LEVELS + 1];       // running CTA counts
    SampleP lv[TDRN_MAX_OFFSET_LEVELS];
};

template <bool F16>
__global__ void __launch_bounds__(DS_MAX_WARPS * 32, 2) deform_sample_group_kernel(const __grid_constant__ SampleGroup g)
{
    int k = 0;
    while (k + 1 < g.n && (int)blockIdx.x >= g.first[k + 1]) ++k;
    deform_sample_body<F16>(g.lv[k], (int)blockIdx.x - g.first[k]);
}

}  // namespace tdrn

using namespace tdrn;

// Validates one level and fills its kernel parameters; nw = warps the level needs, smem = its dynamic shared memory.
static int sample_setup(const tdrn_deform_head_desc *d, const void *proj, int n_pad, const float *offsets, const float *offsets2,
                        float *loc_out, float *conf_out, SampleP &p, int &nw_out, size_t &smem_out)
{
    TDRN_REQUIRE(d && proj && offsets && loc_out && conf_out, "tdrn_deform_head_sample: null argument");
    TDRN_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->num_classes > 0, "tdrn_deform_head_sample: bad shape");
    TDRN_REQUIRE(d->kh > 0 && 2 * d->pad == d->kh - 1, "tdrn_deform_head_sample: head 1 must be 'same' (2*pad == k-1)");
    TDRN_REQUIRE(d->kh2 == 0 || (2 * d->pad2 == d->kh2 - 1 && offsets2), "tdrn_deform_head_sample: bad second head");
    p = SampleP{};
    p.n_total = 12 + 3 * d->num_classes;
    const int split = d->split & 1;
    TDRN_REQUIRE(d->split >= 0 && d->split <= 2, "tdrn_deform_head_sample: split is 0, 1 (hi | lo bf16 pairs) or 2 (IEEE-half projections)");
    p.f16 = d->split == 2;
    if (split && (n_pad % 16 != 0 || n_pad / 2 < p.n_total)) {
        set_error("tdrn_deform_head_sample: split projections need n_pad = 2*g with 12+3*C <= g (got C=%d n_pad=%d)", d->num_classes, n_pad);
        return TDRN_EUNSUPPORTED;
    }
    p.merge_g = split ? n_pad / 2 : 0;
    if (d->dg != 1 || n_pad % 8 != 0 || n_pad < p.n_total || n_pad > 256) {
        set_error("tdrn_deform_head_sample: needs one deformable group and 12+3*C <= n_pad <= 256, n_pad %% 8 == 0 "
                  "(got dg=%d C=%d n_pad=%d)", d->dg, d->num_classes, n_pad);
        return TDRN_EUNSUPPORTED;
    }
    p.y = (const uint4 *)proj; p.off[0] = offsets; p.off[1] = offsets2;
    p.B = d->B; p.H = d->H; p.W = d->W;
    p.k[0] = d->kh; p.pad[0] = d->pad; p.taps[0] = d->kh * d->kh;
    p.k[1] = d->kh2 ? d->kh2 : 1; p.pad[1] = d->pad2; p.taps[1] = d->kh2 * d->kh2;
    p.taps_total = p.taps[0] + p.taps[1];
    p.n_pad = n_pad; p.lpp = n_pad / 8; p.ppw = 32 / p.lpp;
    p.C = d->num_classes; p.P = d->P; p.prior_off = d->prior_off; p.softmax = d->softmax;
    p.loc_out = loc_out; p.conf_out = conf_out;
    const long long units = (long long)d->B * d->H * d->W * p.taps_total * p.lpp;
    TDRN_REQUIRE(units < (1ll << 30), "tdrn_deform_head_sample: projection tensor too large for one call (%lld 16-byte units); "
                                      "split the batch", units);
    // pixel patch of a CTA: as many of the 16 warps' pixel slots as possible, few wasted slots, small halo
    const int max_slots = DS_MAX_WARPS * p.ppw;
    double best = -1.0;
    for (int tw = 1; tw <= d->W && tw <= max_slots; ++tw)
        for (int th = 1; th <= d->H && tw * th <= max_slots; ++th) {
            const int nw = (tw * th + p.ppw - 1) / p.ppw;
            const int tiles = ((d->W + tw - 1) / tw) * ((d->H + th - 1) / th);
            const double fill = (double)(d->W * d->H) / ((double)tiles * nw * p.ppw);
            const double halo = (double)(tw * th) / ((double)(tw + 2) * (th + 2));
            const double score = fill * halo;
            if (score > best) { best = score; p.TW = tw; p.TH = th; }
        }
    p.tiles_w = (d->W + p.TW - 1) / p.TW; p.tiles_h = (d->H + p.TH - 1) / p.TH;
    const int nw = (p.TW * p.TH + p.ppw - 1) / p.ppw;
    p.slots = nw * p.ppw;                                  // slots beyond TW*TH are idle (rvalid false)
    p.taps_pad = (p.taps_total + 1) & ~1;                  // the tap loop is unrolled by two
    const size_t geo_bytes = (size_t)p.slots * p.taps_pad * 16;
    const size_t stg_bytes = (size_t)p.slots * (n_pad + 1) * 4;
    smem_out = geo_bytes > stg_bytes ? geo_bytes : stg_bytes;
    nw_out = nw;
    return TDRN_OK;
}

extern "C" int tdrn_deform_head_sample(const tdrn_deform_head_desc *d, const void *proj, int n_pad,
                                       const float *offsets, const float *offsets2, float *loc_out, float *conf_out,
                                       tdrn_stream_t stream)
{
    SampleP p;
    int nw = 0;
    size_t smem = 0;
    const int rc = sample_setup(d, proj, n_pad, offsets, offsets2, loc_out, conf_out, p, nw, smem);
    if (rc != TDRN_OK) return rc;
    if (p.f16) {
        TDRN_CUDA(cudaFuncSetAttribute(deform_sample_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        deform_sample_kernel<true><<<d->B * p.tiles_w * p.tiles_h, nw * 32, smem, as_stream(stream)>>>(p);
    } else {
        TDRN_CUDA(cudaFuncSetAttribute(deform_sample_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        deform_sample_kernel<false><<<d->B * p.tiles_w * p.tiles_h, nw * 32, smem, as_stream(stream)>>>(p);
    }
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

extern "C" int tdrn_deform_head_sample_group(int n_levels, const tdrn_deform_head_desc *descs, const void *const *projs, int n_pad,
                                             const float *const *offsets, const float *const *offsets2, float *loc_out, float *conf_out,
                                             tdrn_stream_t stream)
{
    TDRN_REQUIRE(n_levels > 0 && n_levels <= TDRN_MAX_OFFSET_LEVELS && descs && projs && offsets, "tdrn_deform_head_sample_group: bad argument");
    SampleGroup g{};
    g.n = n_levels;
    int nw_max = 0, run = 0;
    size_t smem_max = 0;
    for (int k = 0; k < n_levels; ++k) {
        int nw = 0;
        size_t smem = 0;
        const int rc = sample_setup(&descs[k], projs[k], n_pad, offsets[k], offsets2 ? offsets2[k] : nullptr, loc_out, conf_out, g.lv[k], nw, smem);
        if (rc != TDRN_OK) return rc;
        g.first[k] = run;
        run += descs[k].B * g.lv[k].tiles_w * g.lv[k].tiles_h;
        nw_max = nw > nw_max ? nw : nw_max;
        smem_max = smem > smem_max ? smem : smem_max;
    }
    g.first[n_levels] = run;
    for (int k = 1; k < n_levels; ++k)
        TDRN_REQUIRE(g.lv[k].f16 == g.lv[0].f16, "tdrn_deform_head_sample_group: all levels must use the same projection format");
    if (g.lv[0].f16) {
        TDRN_CUDA(cudaFuncSetAttribute(deform_sample_group_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        deform_sample_group_kernel<true><<<run, nw_max * 32, smem_max, as_stream(stream)>>>(g);
    } else {
        TDRN_CUDA(cudaFuncSetAttribute(deform_sample_group_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        deform_sample_group_kernel<false><<<run, nw_max * 32, smem_max, as_stream(stream)>>>(g);
    }
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}
