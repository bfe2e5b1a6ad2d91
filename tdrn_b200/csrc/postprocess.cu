// postprocess.cu -- two-stage box decode, per-class candidate selection, sort, greedy NMS, top-k.
//
// Replaces the Python loops of Detect.forward (layers/functions/detection.py:25-70: B x (C-1)
// device->host round trips + single-core Cython NMS, utils/nms/cpu_nms.pyx:17-68) with two kernels:
//
//   detect_front_kernel : ONE pass over [B,P,(4 + 4 + C)]: boxes[B,P,4] = decode(loc, center_size(decode(arm_loc,
//                         priors))) (layers/box_utils.py:176-195, :16-25) and, for every (prior, class >= 1) with
//                         score > conf_thresh (strict, detection.py:53), a 64-bit key (score bits << 32 | ~prior) appended
//                         to the (image, class) segment's candidate list (warp-aggregated atomics).  The lists are
//                         what everything downstream reads: no class-major copy of the scores is made.
//   nms_segment_kernel  : one CTA per (image, class) segment, two paths:
//     small (<= NMS_WCAP candidates -- every segment of a trained detector, ~1 % of the priors pass the threshold):
//       sort the keys, gather the boxes once, all-pairs suppression bit matrix with the whole CTA, one warp walks the
//       candidates in score order over the bit rows (nms_small_path);
//     large (random-init scores: all P priors pass): greedy NMS only ever consumes the highest-scoring candidates until
//       top_k boxes are kept (SURVEY.md 8a "Exactness note for A8"), so instead of sorting all candidates the kernel
//       works in descending score BATCHES:
//       1. radix-select (12/12/8/8.. bit digits, shared-memory histograms) the cut-off such that
//          the next <= 1024 candidates in (score desc, index asc) order are selected;
//       2. bitonic-sort that batch on the 64-bit keys in shared memory;
//       3. greedy NMS in chunks of 32: 8 lanes per candidate scan the kept list, a 32x32
//          suppression matrix resolves the chunk with warp shuffles;
//       4. stop when top_k boxes are kept or the candidates are exhausted, else next batch.
//     The visiting order is exactly the reference's (descending score, pinned tie rule), so the
//     result is identical to a full sort + full scan.
//
// Bit-exactness: every fp32 operation of the reference's IoU (cpu_nms.pyx:24,57-65) and of decode
// is issued with explicit round-to-nearest intrinsics in the reference's order so that nvcc cannot
// contract them into FMAs; `ovr >= thresh` is the reference's float-vs-double compare (the division
// is skipped only when the outcome is certain by a 1e-6 relative margin, 16x the rounding error).
#include "common.cuh"
#include <math.h>

namespace tdrn {

constexpr int NMS_THREADS = 256;
constexpr int NMS_CAP = 1024;              // candidates sorted per batch
constexpr int NMS_KSEL = 512;              // every batch holds at least min(KSEL, remaining) candidates
constexpr int NMS_BINS = 4096;             // histogram bins (12-bit digit)
constexpr int NMS_CH = 32;                 // candidates per greedy chunk
constexpr int NMS_WCAP = 256;              // segments up to this size take the all-pairs path (nms_small_path)

__device__ __forceinline__ float4 decode_box(float4 l, float4 p)
{
    // box_utils.py:190-195: cxcy = p_xy + (l_xy*0.1)*p_wh ; wh = p_wh*exp(l_wh*0.2);
    // x1y1 = cxcy - wh/2 ; x2y2 = wh + x1y1
    const float cx = __fadd_rn(p.x, __fmul_rn(__fmul_rn(l.x, 0.1f), p.z));
    const float cy = __fadd_rn(p.y, __fmul_rn(__fmul_rn(l.y, 0.1f), p.w));
    const float w = __fmul_rn(p.z, expf(__fmul_rn(l.z, 0.2f)));
    const float h = __fmul_rn(p.w, expf(__fmul_rn(l.w, 0.2f)));
    const float x1 = __fsub_rn(cx, __fmul_rn(w, 0.5f));
    const float y1 = __fsub_rn(cy, __fmul_rn(h, 0.5f));
    return make_float4(x1, y1, __fadd_rn(w, x1), __fadd_rn(h, y1));
}

__device__ __forceinline__ float4 center_size_box(float4 b)
{
    // box_utils.py:24-25
    return make_float4(__fmul_rn(__fadd_rn(b.z, b.x), 0.5f), __fmul_rn(__fadd_rn(b.w, b.y), 0.5f),
                       __fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

constexpr int DEC_TP = 128;   // priors per CTA

// order-preserving fp32 -> u32 (handles negative scores in standalone mode)
__device__ __forceinline__ unsigned score_key(float s)
{
    const unsigned u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_score(unsigned long long c)   // inverse of score_key for the key's upper half
{
    const unsigned k = (unsigned)(c >> 32);
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// grid (ceil(P/DEC_TP), B).  conf == NULL: plain tdrn_decode (boxes only).
// cand [B*C][P] u64 candidate keys per (image, class) segment, cnt [B*C] their number (zeroed by the caller).
__global__ void __launch_bounds__(256) detect_front_kernel(const float4 *__restrict__ loc, const float4 *__restrict__ priors,
                                                           const float4 *__restrict__ arm_loc, const float *__restrict__ conf,
                                                           float4 *__restrict__ boxes, unsigned long long *__restrict__ cand,
                                                           unsigned *__restrict__ cnt, int P, int C, float conf_thresh)
{
    extern __shared__ __align__(16) float s_raw[];  // [4 + DEC_TP * C] scores, then 3 * C counters
    const int b = blockIdx.y, p0 = blockIdx.x * DEC_TP;
    const int np = min(DEC_TP, P - p0);
    float *s_conf = s_raw;
    if (conf) {                                // issue the score tile's loads first: the decode below overlaps them
        // The tile starts (b * P + p0) * C floats into conf: rarely on a 16-byte boundary (C = 21).  The shared copy is placed
        // at the SAME offset from a 16-byte boundary as the global tile, so that all but the first / last few elements move as
        // aligned float4 (a scalar copy loop was 18 % of this kernel's stall samples and most of its instructions).
        const float *src = conf + ((long long)b * P + p0) * C;
        const int n = np * C;
        const int k = (int)((((uintptr_t)src) >> 2) & 3);          // floats past a 16-byte boundary
        const int head = min((4 - k) & 3, n);                      // scalar elements in front of the first aligned float4
        s_conf = s_raw + k;
        const int nv = (n - head) >> 2;
        const float4 *src4 = (const float4 *)(src + head);
        float4 *dst4 = (float4 *)(s_conf + head);
        for (int i = threadIdx.x; i < nv; i += blockDim.x) dst4[i] = __ldg(src4 + i);
        if (threadIdx.x < head) s_conf[threadIdx.x] = src[threadIdx.x];
        const int tail0 = head + 4 * nv;
        if (threadIdx.x < n - tail0) s_conf[tail0 + threadIdx.x] = src[tail0 + threadIdx.x];
    }
    if (threadIdx.x < np) {
        const int p = p0 + threadIdx.x;
        float4 prior = priors[p];
        if (arm_loc) prior = center_size_box(decode_box(arm_loc[(long long)b * P + p], prior));
        boxes[(long long)b * P + p] = decode_box(loc[(long long)b * P + p], prior);
    }
    if (!conf) return;
    // Candidate compaction with ONE global atomic per (CTA, class): pass 1 counts the CTA's candidates of every class in
    // shared memory, the first C-1 threads then reserve the CTA's range in every segment list at once (one round trip to
    // L2 for the whole CTA), pass 2 repeats the ballots and writes the keys at range base + rank inside the CTA.
    unsigned *s_cnt = (unsigned *)(s_raw + 4 + DEC_TP * C);  // [C] candidates of class cl in this tile
    unsigned *s_pos = s_cnt + C;                           // [C] running rank inside the CTA (pass 2)
    unsigned *s_base = s_pos + C;                          // [C] start of the CTA's range in the segment list
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_cnt[i] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    // item i = (class cl, prior pl) with the prior fastest: a warp looks at 32 consecutive priors of one class
    // (DEC_TP is a multiple of 32 and the tile is padded up to it, so a warp never straddles two classes)
    for (int i = threadIdx.x; i < DEC_TP * (C - 1); i += blockDim.x) {
        const int cl = 1 + i / DEC_TP, pl = i - (cl - 1) * DEC_TP;
        const bool take = pl < np && s_conf[pl * C + cl] > conf_thresh;    // strict fp32 compare, detection.py:53
        const unsigned ballot = __ballot_sync(0xffffffffu, take);
        if (lane == 0 && ballot) atomicAdd(&s_cnt[cl], (unsigned)__popc(ballot));
    }
    __syncthreads();
    for (int cl = 1 + threadIdx.x; cl < C; cl += blockDim.x)
        s_base[cl] = s_cnt[cl] ? atomicAdd(&cnt[b * C + cl], s_cnt[cl]) : 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < DEC_TP * (C - 1); i += blockDim.x) {
        const int cl = 1 + i / DEC_TP, pl = i - (cl - 1) * DEC_TP;
        const float sc = pl < np ? s_conf[pl * C + cl] : 0.f;
        const bool take = pl < np && sc > conf_thresh;
        const unsigned ballot = __ballot_sync(0xffffffffu, take);
        if (ballot == 0u) continue;                                        // warp-uniform
        unsigned off = 0;
        if (lane == 0) off = atomicAdd(&s_pos[cl], (unsigned)__popc(ballot));
        off = __shfl_sync(0xffffffffu, off, 0);
        if (take)
            cand[(long long)(b * C + cl) * P + s_base[cl] + off + __popc(ballot & ((1u << lane) - 1u))] =
                ((unsigned long long)score_key(sc) << 32) | (unsigned)(0xffffffffu - (unsigned)(p0 + pl));
    }
}

struct NmsP {
    // detect mode
    const float4 *boxes;      // [B,P,4] normalised
    const unsigned long long *cand;   // [B*C][P] candidate keys (detect_front_kernel)
    const unsigned *cnt;      // [B*C]
    float *out;               // [B,C,top_k,5]
    int P, C, top_k;
    float conf_thresh;
    float4 scale;
    // standalone mode
    const float *dets;        // [n,5]
    int n;
    int *keep; int *num_keep;
    float *kept_ws;           // global scratch for the kept list: float4 box[kept_cap] then float area[kept_cap]
    // common
    int max_keep;             // <= 0: unlimited
    int kept_cap;             // stride of the kept SoA arrays
    float thr_up;             // smallest float >= (double) nms threshold
};

__device__ __forceinline__ bool iou_ge(float4 a, float aarea, float4 b, float barea, float thr)
{
    // cpu_nms.pyx:57-65
    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    const float w = fmaxf(0.0f, __fadd_rn(__fsub_rn(xx2, xx1), 1.0f));
    const float h = fmaxf(0.0f, __fadd_rn(__fsub_rn(yy2, yy1), 1.0f));
    const float inter = __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(aarea, barea), inter);
    if (uni > 0.f) {
        // ovr = rn(inter / uni) carries <= 2^-24 relative error; decide without the division when the
        // real quotient is further than 1e-6 (relative) from the threshold.
        if (inter == 0.f) return 0.f >= thr;
        const float t = __fmul_rn(thr, uni);
        if (inter > __fmul_rn(t, 1.000001f)) return true;
        if (inter < __fmul_rn(t, 0.999999f)) return false;
    }
    return __fdiv_rn(inter, uni) >= thr;
}

__device__ __forceinline__ float box_area(float4 b)
{
    // cpu_nms.pyx:24
    return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.0f), __fadd_rn(__fsub_rn(b.w, b.y), 1.0f));
}

// Small segment (n <= NMS_WCAP candidates; every segment of a trained detector, where ~1 % of the priors pass the
// threshold): no selection rounds and no serial scan of a kept list.  The CTA sorts the n keys, gathers the n boxes once,
// computes the upper triangle of the n x n suppression matrix with all 256 threads (independent IoU tests: the latency
// of one does not wait for the previous one), and warp 0 then walks the candidates in score order over the bit rows.
// Same visiting order, same fp32 IoU and the same early exit at top_k as the batch path below.
// keys [>= NMS_WCAP], scratch [>= 13 KB, 16-byte aligned], klist [>= NMS_WCAP] are shared-memory regions of the caller.
__device__ __forceinline__ void nms_small_path(const NmsP &p, int b, int seg, int n, float *out_seg, unsigned long long *keys,
                                               unsigned char *scratch, unsigned *klist, unsigned *s_kept)
{
    constexpr int MW = NMS_WCAP / 32;                               // words per matrix row
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float4 *sbox = (float4 *)scratch;                               // [NMS_WCAP] boxes in NMS (pixel) units
    float *sarea = (float *)(scratch + NMS_WCAP * 16);              // [NMS_WCAP]
    unsigned *M = (unsigned *)(scratch + NMS_WCAP * 20);            // [NMS_WCAP][MW]: bit j of row i = candidate j (> i) overlaps i
    const unsigned long long *cand = p.cand + (long long)seg * p.P;
    const float4 *boxes = p.boxes + (long long)b * p.P;
    const int nw = (n + 31) >> 5;
    // rank sort (n <= NMS_WCAP = NMS_THREADS): thread t owns candidate t; its position in (score desc, prior asc) order is
    // the number of keys larger than its own (keys are distinct: they carry the prior index).  One pass of broadcast
    // shared-memory reads and two barriers, against 28-36 barrier stages of a bitonic network.
    unsigned long long *unsorted = keys + NMS_WCAP;                 // keys[] has NMS_CAP = 4 * NMS_WCAP slots
    const unsigned long long mine = tid < n ? cand[tid] : 0ull;
    if (tid < n) unsorted[tid] = mine;
    __syncthreads();
    if (tid < n) {
        int rank = 0;
#pragma unroll 4
        for (int j = 0; j < n; ++j) rank += unsorted[j] > mine ? 1 : 0;
        keys[rank] = mine;
    }
    __syncthreads();
    for (int i = tid; i < n; i += NMS_THREADS) {
        const float4 nb = boxes[(int)(0xffffffffu - (unsigned)(keys[i] & 0xffffffffull))];
        const float4 bx = make_float4(__fmul_rn(nb.x, p.scale.x), __fmul_rn(nb.y, p.scale.y),
                                      __fmul_rn(nb.z, p.scale.z), __fmul_rn(nb.w, p.scale.w));   // detection.py:59
        sbox[i] = bx; sarea[i] = box_area(bx);
    }
    __syncthreads();
    // one warp per (candidate i, word w >= i / 32 of its row): lane jj tests the pair (i, 32 w + jj), the ballot IS the word
    {
        int item = warp;                                            // items enumerate (i, w) with w >= i >> 5, row-major
        for (int i = 0; i < n; ++i) {
            const int w0 = i >> 5;
            const int cnt_i = nw - w0;
            if (item >= cnt_i) { item -= cnt_i; continue; }
            const float4 bi = sbox[i];
            const float ai = sarea[i];
            for (; item < cnt_i; item += NMS_THREADS / 32) {
                const int w = w0 + item, j = (w << 5) + lane;
                const bool hit = j > i && j < n && iou_ge(bi, ai, sbox[j], sarea[j], p.thr_up);
                const unsigned bits = __ballot_sync(0xffffffffu, hit);
                if (lane == 0) M[i * MW + w] = bits;
            }
            item -= cnt_i;
        }
    }
    __syncthreads();
    if (warp == 0) {
        const unsigned max_keep = (unsigned)p.max_keep;
        unsigned removed = 0u, kept_n = 0u;                         // lane l holds word l of the suppressed set
        for (int w = 0; w < nw && kept_n < max_keep; ++w) {
            const unsigned r = __shfl_sync(0xffffffffu, removed, w);
            const unsigned valid = (w == nw - 1 && (n & 31)) ? ((1u << (n & 31)) - 1u) : 0xffffffffu;
            const unsigned diag = M[((w << 5) + lane) * MW + w];    // lane l: which later candidates of this word overlap candidate l
            unsigned alive = ~r & valid, keepmask = 0u;
            while (alive && kept_n < max_keep) {                    // uniform: the lowest pending candidate of the word is kept
                const int i = __ffs(alive) - 1;
                keepmask |= 1u << i;
                ++kept_n;
                alive &= ~(__shfl_sync(0xffffffffu, diag, i) | (1u << i));
            }
            if ((keepmask >> lane) & 1u) klist[kept_n - __popc(keepmask >> lane)] = (unsigned)((w << 5) + lane);
            if (lane > w && lane < nw) {                            // the kept candidates' rows knock out later words
                for (unsigned km = keepmask; km; km &= km - 1u) removed |= M[((w << 5) + __ffs(km) - 1) * MW + lane];
            }
        }
        if (lane == 0) *s_kept = kept_n;
    }
    __syncthreads();
    const unsigned kept_n = *s_kept;
    for (unsigned r = tid; r < kept_n; r += NMS_THREADS) {          // detection.py:61-63
        const unsigned long long key = keys[klist[r]];
        const float4 nb = boxes[(int)(0xffffffffu - (unsigned)(key & 0xffffffffull))];
        float *o = out_seg + r * 5;
        o[0] = key_score(key); o[1] = nb.x; o[2] = nb.y; o[3] = nb.z; o[4] = nb.w;
    }
    for (int i = kept_n * 5 + tid; i < p.top_k * 5; i += NMS_THREADS) out_seg[i] = 0.f;
}

// One CTA per segment.  DETECT: grid (C, B); class 0 only zero-fills.  Standalone: grid (1).
template <bool DETECT>
__global__ void __launch_bounds__(NMS_THREADS, 5) nms_segment_kernel(const NmsP p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *kept_s = (float *)smem_raw;                                      // DETECT: [5][kept_cap]
    __shared__ unsigned long long keys[NMS_CAP];
    __shared__ __align__(16) unsigned hist[NMS_BINS];
    __shared__ unsigned part[NMS_THREADS];
    __shared__ float4 cbox2[2][NMS_CH];      // chunk boxes / areas / source indices, double-buffered: warp 1 fetches the next
    __shared__ float carea2[2][NMS_CH];      // chunk while warp 0 resolves the current one
    __shared__ int cidx2[2][NMS_CH];
    __shared__ unsigned csup[NMS_CH];        // suppressed by an earlier-kept box
    __shared__ unsigned crow[NMS_CH];        // bit j: candidate j (> i) overlaps candidate i
    __shared__ unsigned long long s_lo;
    __shared__ unsigned s_nsel, s_total, s_kept;
    __shared__ int s_done;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int cl = 0, b = 0;
    float *out_seg = nullptr;
    const unsigned long long *cand = nullptr;
    int n_in;
    if (DETECT) {
        cl = blockIdx.x; b = blockIdx.y;
        out_seg = p.out + ((long long)b * p.C + cl) * p.top_k * 5;
        if (cl == 0) {                                             // background row stays zero (detection.py:37,52)
            for (int i = tid; i < p.top_k * 5; i += NMS_THREADS) out_seg[i] = 0.f;
            return;
        }
        n_in = (int)p.cnt[b * p.C + cl];
        if (n_in <= NMS_WCAP) {                                    // uniform across the CTA
            nms_small_path(p, b, b * p.C + cl, n_in, out_seg, keys, (unsigned char *)hist, part, &s_kept);
            return;
        }
        cand = p.cand + (long long)(b * p.C + cl) * p.P;
    } else {
        n_in = p.n;
    }
    float *kept = DETECT ? kept_s : p.kept_ws;
    const int kc = p.kept_cap;
    float4 *kbox = (float4 *)kept;          // kept boxes as float4 (one 16-byte load per IoU test) + areas behind them
    float *kar = kept + 4 * kc;

    // composite key of element i, or 0 if it is not a candidate / not below the current bound
    auto key_of = [&](int i, unsigned long long hi_incl) -> unsigned long long {
        unsigned long long c;
        if (DETECT) {
            c = cand[i];                                            // already thresholded and keyed by detect_front_kernel
        } else {
            const float s = p.dets[5 * (long long)i + 4];
            c = ((unsigned long long)score_key(s) << 32) | (unsigned)(0xffffffffu - (unsigned)i);
        }
        return c <= hi_incl ? c : 0ull;
    };

    unsigned long long hi_incl = ~0ull;
    unsigned kept_n = 0, processed = 0, total = 0;
    const unsigned max_keep = p.max_keep > 0 ? (unsigned)p.max_keep : 0xffffffffu;
    bool first = true;

    while (true) {
        // ---- 1. radix select the lower bound `lo` of the next batch: {c : lo <= c <= hi_incl}, size <= CAP ----------
        unsigned long long prefix_val = 0ull, prefix_mask = 0ull;
        unsigned need = NMS_KSEL, above_total = 0;
        const int shifts[7] = {52, 40, 32, 24, 16, 8, 0};
        const int bits[7] = {12, 12, 8, 8, 8, 8, 8};
        for (int lv = 0; lv < 7; ++lv) {
            for (int i = tid; i < NMS_BINS; i += NMS_THREADS) hist[i] = 0;
            __syncthreads();
            const int sh = shifts[lv];
            const unsigned dmask = (1u << bits[lv]) - 1u;
            for (int i = tid; i < n_in; i += NMS_THREADS) {
                const unsigned long long c = key_of(i, hi_incl);
                if (c != 0ull && (c & prefix_mask) == prefix_val) atomicAdd(&hist[(unsigned)(c >> sh) & dmask], 1u);
            }
            __syncthreads();
            // per-thread partial sums over 16 bins, grouped from the top bin downwards
            {
                unsigned s = 0;
                const int base = NMS_BINS - 1 - tid * (NMS_BINS / NMS_THREADS);
#pragma unroll
                for (int k = 0; k < NMS_BINS / NMS_THREADS; ++k) s += hist[base - k];
                part[tid] = s;
            }
            __syncthreads();
            if (tid == 0) {
                unsigned cum = 0, tot = 0;
                for (int t = 0; t < NMS_THREADS; ++t) tot += part[t];
                if (lv == 0 && first) s_total = tot;
                int done = 0;
                unsigned long long lo = 0ull;
                if (lv == 0 && tot <= (unsigned)NMS_CAP) {          // everything that is left fits in one batch
                    done = 1; lo = 1ull; s_nsel = tot;
                } else {
                    int t = 0;
                    while (cum + part[t] < need) { cum += part[t]; ++t; }   // group holding the need-th largest
                    int d = NMS_BINS - 1 - t * (NMS_BINS / NMS_THREADS);
                    while (cum + hist[d] < need) { cum += hist[d]; --d; }   // digit holding it
                    const unsigned inbin = hist[d];
                    const unsigned long long pv = prefix_val | ((unsigned long long)d << sh);
                    if (above_total + cum + inbin <= (unsigned)NMS_CAP) {   // take the whole bucket: batch complete
                        done = 1; lo = pv; s_nsel = above_total + cum + inbin;
                        if (lo == 0ull) lo = 1ull;
                    } else {                                         // refine inside bucket d at the next digit
                        lo = pv;
                        s_nsel = cum;                                // (re-used to pass `cum` to all threads)
                    }
                }
                s_lo = lo; s_done = done;
            }
            __syncthreads();
            if (s_done) break;
            prefix_val = s_lo;
            prefix_mask |= (unsigned long long)dmask << sh;
            need -= s_nsel; above_total += s_nsel;
            __syncthreads();
        }
        if (first) { total = s_total; first = false; }
        const unsigned long long lo = s_lo;
        const unsigned n_sel = s_nsel;
        if (n_sel == 0) break;
        __syncthreads();

        // ---- 2. gather the batch and bitonic-sort it (descending) ------------------------------------------------
        if (tid == 0) s_nsel = 0;
        __syncthreads();
        for (int i0 = 0; i0 < n_in; i0 += NMS_THREADS) {
            const int i = i0 + tid;
            const unsigned long long c = i < n_in ? key_of(i, hi_incl) : 0ull;
            const bool take = c >= lo && c != 0ull;
            const unsigned ballot = __ballot_sync(0xffffffffu, take);
            unsigned base = 0;
            if (lane == 0 && ballot) base = atomicAdd(&s_nsel, (unsigned)__popc(ballot));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (take) keys[base + __popc(ballot & ((1u << lane) - 1u))] = c;
        }
        __syncthreads();
        int n_pad = 32;
        while (n_pad < (int)n_sel) n_pad <<= 1;
        for (int i = n_sel + tid; i < n_pad; i += NMS_THREADS) keys[i] = 0ull;
        __syncthreads();
        for (int k = 2; k <= n_pad; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < (n_pad >> 1); t += NMS_THREADS) {
                    const int lo_i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int hi_i = lo_i | j;
                    const unsigned long long a = keys[lo_i], c = keys[hi_i];
                    const bool desc = (lo_i & k) == 0;
                    if (desc ? (a < c) : (a > c)) { keys[lo_i] = c; keys[hi_i] = a; }
                }
                __syncthreads();
            }
        }

        // ---- 3. greedy NMS over the sorted batch, 32 candidates per chunk -----------------------------------------
        auto load_chunk = [&](unsigned c0, int buf, int t) {               // t = 0..31: candidate of the chunk
            if (c0 + (unsigned)t >= n_sel) return;
            const int src = (int)(0xffffffffu - (unsigned)(keys[c0 + t] & 0xffffffffull));
            float4 bx;
            if (DETECT) {
                const float4 nb = p.boxes[(long long)b * p.P + src];
                bx = make_float4(__fmul_rn(nb.x, p.scale.x), __fmul_rn(nb.y, p.scale.y),
                                 __fmul_rn(nb.z, p.scale.z), __fmul_rn(nb.w, p.scale.w));   // detection.py:59
            } else {
                const float *d = p.dets + 5 * (long long)src;
                bx = make_float4(d[0], d[1], d[2], d[3]);
            }
            cbox2[buf][t] = bx; carea2[buf][t] = box_area(bx); cidx2[buf][t] = src;
        };
        if (tid < NMS_CH) load_chunk(0, 0, tid);
        __syncthreads();
        int cur = 0;
        for (unsigned c0 = 0; c0 < n_sel && kept_n < max_keep; c0 += NMS_CH, cur ^= 1) {
            const int cnt = min(NMS_CH, (int)(n_sel - c0));
            const float4 *cbox = cbox2[cur];
            const float *carea = carea2[cur];
            const int *cidx = cidx2[cur];
            {
                // thread (ci, l): candidate ci = tid / 8, sub-lane l = tid % 8
                const int ci = tid >> 3, l = tid & 7;
                bool sup = false;
                unsigned row = 0;
                if (ci < cnt) {
                    const float4 bx = cbox[ci];
                    const float ar = carea[ci];
                    // (a group vote that stops all 8 sub-lanes at the first hit was measured slower: 0.276 vs 0.233 ms; an
                    // area-ratio pre-test -- ovr <= min(area)/max(area), so cross-level pairs cannot reach the threshold --
                    // gave nothing either: the lanes of a warp diverge and the warp still pays for the full test)
                    for (unsigned q = l; q < kept_n && !sup; q += 8)        // vs boxes kept so far
                        sup = iou_ge(kbox[q], kar[q], bx, ar, p.thr_up);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {                         // my 4 columns of the 32x32 chunk matrix
                        const int j = l * 4 + jj;
                        if (j > ci && j < cnt && iou_ge(bx, ar, cbox[j], carea[j], p.thr_up)) row |= 1u << j;
                    }
                }
                // combine the 8 sub-lanes of each candidate (they sit in the same warp, aligned groups of 8)
                unsigned s = sup ? 1u : 0u;
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) { s |= __shfl_xor_sync(0xffffffffu, s, o); row |= __shfl_xor_sync(0xffffffffu, row, o); }
                if (l == 0 && ci < NMS_CH) { csup[ci] = s; crow[ci] = row; }
            }
            __syncthreads();
            if (warp == 1) load_chunk(c0 + NMS_CH, cur ^ 1, lane);          // next chunk's boxes, hidden behind the resolution
            if (warp == 0) {
                const bool alive = lane < cnt && !csup[lane];
                const unsigned alive_mask = __ballot_sync(0xffffffffu, alive);
                const unsigned myrow = crow[lane];
                // greedy resolution of the chunk in candidate order, visiting only the candidates that are still pending
                // (typically a handful of the 32): the lowest pending one is kept and knocks out the later ones it overlaps
                unsigned pending = alive_mask, keepmask = 0, k = kept_n;
                while (pending && k < max_keep) {                            // uniform across the warp
                    const int i = __ffs(pending) - 1;
                    const unsigned ri = __shfl_sync(0xffffffffu, myrow, i);
                    keepmask |= 1u << i;
                    ++k;
                    pending &= ~((1u << i) | ri);
                }
                if ((keepmask >> lane) & 1u) {
                    const unsigned pos = kept_n + __popc(keepmask & ((1u << lane) - 1u));
                    const float4 kb = cbox[lane];
                    kbox[pos] = kb; kar[pos] = carea[lane];
                    if (DETECT) {
                        const float4 nb = p.boxes[(long long)b * p.P + cidx[lane]];
                        float *o = out_seg + pos * 5;                       // detection.py:61-63
                        o[0] = key_score(keys[c0 + lane]); o[1] = nb.x; o[2] = nb.y; o[3] = nb.z; o[4] = nb.w;
                    } else {
                        p.keep[pos] = cidx[lane];
                    }
                }
                if (lane == 0) s_kept = k;
            }
            __syncthreads();
            kept_n = s_kept;
        }
        processed += n_sel;
        if (kept_n >= max_keep || processed >= total) break;
        hi_incl = lo - 1ull;
        __syncthreads();
    }

    if (DETECT) {
        for (int i = kept_n * 5 + tid; i < p.top_k * 5; i += NMS_THREADS) out_seg[i] = 0.f;
    } else if (tid == 0) {
        *p.num_keep = (int)kept_n;
    }
}

static float thresh_up(double t)
{
    float f = (float)t;
    if ((double)f < t) f = nextafterf(f, INFINITY);
    return f;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace tdrn

using namespace tdrn;

extern "C" int tdrn_decode(const float *loc, const float *priors, const float *arm_loc, int B, int P, float *boxes, tdrn_stream_t stream)
{
    TDRN_REQUIRE(loc && priors && boxes && B > 0 && P > 0, "tdrn_decode: bad argument");
    dim3 grid(ceil_div(P, DEC_TP), B);
    detect_front_kernel<<<grid, 256, 0, as_stream(stream)>>>((const float4 *)loc, (const float4 *)priors, (const float4 *)arm_loc,
                                                             nullptr, (float4 *)boxes, nullptr, nullptr, P, 0, 0.f);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

// workspace: decoded boxes [B,P,4] f32 | candidate counts [B*C] u32 | candidate keys [B*C][P] u64
extern "C" size_t tdrn_detect_workspace_bytes(int B, int P, int C, int top_k)
{
    (void)top_k;
    return align_up((size_t)B * P * 4 * sizeof(float), 256) + align_up((size_t)B * C * sizeof(unsigned), 256) +
           align_up((size_t)B * C * P * sizeof(unsigned long long), 256);
}

extern "C" int tdrn_detect(const float *loc, const float *conf, const float *priors, const float *arm_loc,
                           const float *scale_host, int B, int P, int C, int top_k, float conf_thresh,
                           double nms_thresh, float *out, void *workspace, size_t workspace_bytes, tdrn_stream_t stream)
{
    TDRN_REQUIRE(loc && conf && priors && scale_host && out, "tdrn_detect: null argument");
    TDRN_REQUIRE(B > 0 && P > 0 && C > 1 && top_k > 0, "tdrn_detect: bad shape B=%d P=%d C=%d top_k=%d", B, P, C, top_k);
    TDRN_REQUIRE(nms_thresh > 0, "nms_threshold must be non negative.");          // detection.py:20-21
    if (!workspace || workspace_bytes < tdrn_detect_workspace_bytes(B, P, C, top_k)) {
        set_error("tdrn_detect: workspace too small (%zu < %zu)", workspace_bytes, tdrn_detect_workspace_bytes(B, P, C, top_k));
        return TDRN_EWORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    char *w = (char *)workspace;
    float *boxes = (float *)w;
    w += align_up((size_t)B * P * 4 * sizeof(float), 256);
    unsigned *cnt = (unsigned *)w;
    w += align_up((size_t)B * C * sizeof(unsigned), 256);
    unsigned long long *cand = (unsigned long long *)w;

    const size_t dec_smem = (size_t)(4 + DEC_TP * C) * sizeof(float) + 3 * (size_t)C * sizeof(unsigned);
    TDRN_REQUIRE(dec_smem <= 200 * 1024, "tdrn_detect: too many classes (%d)", C);
    if (dec_smem > 48 * 1024)
        TDRN_CUDA(cudaFuncSetAttribute(detect_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dec_smem));
    TDRN_CUDA(cudaMemsetAsync(cnt, 0, (size_t)B * C * sizeof(unsigned), st));
    dim3 dgrid(ceil_div(P, DEC_TP), B);
    detect_front_kernel<<<dgrid, 256, dec_smem, st>>>((const float4 *)loc, (const float4 *)priors, (const float4 *)arm_loc,
                                                       conf, (float4 *)boxes, cand, cnt, P, C, conf_thresh);
    TDRN_LAUNCH_CHECK();

    NmsP p{};
    p.boxes = (const float4 *)boxes; p.cand = cand; p.cnt = cnt; p.out = out; p.P = P; p.C = C; p.top_k = top_k;
    p.conf_thresh = conf_thresh;
    p.scale = make_float4(scale_host[0], scale_host[1], scale_host[2], scale_host[3]);
    p.max_keep = top_k; p.kept_cap = top_k; p.thr_up = thresh_up(nms_thresh);
    // one CTA per (image, class) segment: all-pairs path for small segments, selection batches for large ones
    const size_t smem = (size_t)top_k * 5 * sizeof(float);
    TDRN_REQUIRE(smem <= 150 * 1024, "tdrn_detect: top_k=%d exceeds the shared-memory kept-list capacity", top_k);
    if (smem > 16 * 1024)
        TDRN_CUDA(cudaFuncSetAttribute(nms_segment_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {   // five 32 KB CTAs per SM (one wave for 672 segments at b32 / VOC-21) need the large shared-memory carve-out
        static bool carve = false;
        if (!carve) { cudaFuncSetAttribute(nms_segment_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); carve = true; }
    }
    nms_segment_kernel<true><<<dim3(C, B), NMS_THREADS, smem, st>>>(p);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

extern "C" size_t tdrn_nms_workspace_bytes(int n) { return align_up((size_t)(n > 0 ? n : 1) * 5 * sizeof(float), 256); }

// rule 0: suppress when ovr >= thresh, thresh compared as a double (cpu_nms.pyx:65, what Detect uses);
// rule 1: suppress when ovr > (float)thresh (the reference's GPU kernel, nms_kernel.cu:71, what nms() runs with force_cpu=False).
// Both become `ovr >= t` for a float t: the smallest float >= thresh, or the float after (float)thresh.
static float rule_threshold(double thresh, int rule)
{
    return rule == 1 ? nextafterf((float)thresh, INFINITY) : thresh_up(thresh);
}

extern "C" int tdrn_nms_rule(const float *dets, int n, double thresh, int rule, int max_keep, int *keep, int *num_keep,
                             void *workspace, size_t workspace_bytes, tdrn_stream_t stream)
{
    TDRN_REQUIRE(keep && num_keep && n >= 0 && (rule == 0 || rule == 1), "tdrn_nms: bad argument");
    cudaStream_t st = as_stream(stream);
    if (n == 0) {                                                      // nms_wrapper.py:26-27
        TDRN_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int), st));
        return TDRN_OK;
    }
    TDRN_REQUIRE(dets != nullptr, "tdrn_nms: null dets");
    if (!workspace || workspace_bytes < tdrn_nms_workspace_bytes(n)) {
        set_error("tdrn_nms: workspace too small (%zu < %zu)", workspace_bytes, tdrn_nms_workspace_bytes(n));
        return TDRN_EWORKSPACE;
    }
    NmsP p{};
    p.dets = dets; p.n = n; p.keep = keep; p.num_keep = num_keep; p.kept_ws = (float *)workspace;
    p.max_keep = max_keep; p.kept_cap = n; p.thr_up = rule_threshold(thresh, rule);
    nms_segment_kernel<false><<<1, NMS_THREADS, 0, st>>>(p);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

extern "C" int tdrn_nms(const float *dets, int n, double thresh, int max_keep, int *keep, int *num_keep,
                        void *workspace, size_t workspace_bytes, tdrn_stream_t stream)
{
    return tdrn_nms_rule(dets, n, thresh, 0, max_keep, keep, num_keep, workspace, workspace_bytes, stream);
}

// Host-pointer, synchronous variant with the argument order of the reference's `_nms`
// (utils/nms/gpu_nms.hpp:1-2), except that the input need not be pre-sorted and the threshold rule
// is the CPU one Detect uses.
extern "C" int tdrn_nms_host_rule(int *keep_out, int *num_out, const float *dets_host, int boxes_num, int boxes_dim,
                                  double thresh, int device_id, int rule)
{
    TDRN_REQUIRE(keep_out && num_out && boxes_num >= 0, "tdrn_nms_host: bad argument");
    TDRN_REQUIRE(boxes_dim == 5, "tdrn_nms_host: boxes_dim must be 5");
    if (boxes_num == 0) { *num_out = 0; return TDRN_OK; }
    TDRN_REQUIRE(dets_host != nullptr, "tdrn_nms_host: null dets");
    if (device_id >= 0) TDRN_CUDA(cudaSetDevice(device_id));
    float *d_dets = nullptr; int *d_keep = nullptr; void *d_ws = nullptr;
    const size_t ws = tdrn_nms_workspace_bytes(boxes_num);
    int rc = TDRN_OK;
    cudaError_t e;
    if ((e = cudaMalloc(&d_dets, sizeof(float) * 5 * boxes_num)) != cudaSuccess ||
        (e = cudaMalloc(&d_keep, sizeof(int) * (boxes_num + 1))) != cudaSuccess ||
        (e = cudaMalloc(&d_ws, ws)) != cudaSuccess) {
        set_error("tdrn_nms_host: cudaMalloc failed: %s", cudaGetErrorString(e));
        rc = TDRN_ECUDA;
    }
    if (rc == TDRN_OK && (e = cudaMemcpy(d_dets, dets_host, sizeof(float) * 5 * boxes_num, cudaMemcpyHostToDevice)) != cudaSuccess) {
        set_error("tdrn_nms_host: H2D failed: %s", cudaGetErrorString(e)); rc = TDRN_ECUDA;
    }
    if (rc == TDRN_OK) rc = tdrn_nms_rule(d_dets, boxes_num, thresh, rule, 0, d_keep + 1, d_keep, d_ws, ws, nullptr);
    if (rc == TDRN_OK) {
        if ((e = cudaMemcpy(num_out, d_keep, sizeof(int), cudaMemcpyDeviceToHost)) != cudaSuccess ||
            (e = cudaMemcpy(keep_out, d_keep + 1, sizeof(int) * (*num_out), cudaMemcpyDeviceToHost)) != cudaSuccess) {
            set_error("tdrn_nms_host: D2H failed: %s", cudaGetErrorString(e)); rc = TDRN_ECUDA;
        }
    }
    cudaFree(d_dets); cudaFree(d_keep); cudaFree(d_ws);
    return rc;
}

extern "C" int tdrn_nms_host(int *keep_out, int *num_out, const float *dets_host, int boxes_num, int boxes_dim,
                             double thresh, int device_id)
{
    return tdrn_nms_host_rule(keep_out, num_out, dets_host, boxes_num, boxes_dim, thresh, device_id, 0);
}
