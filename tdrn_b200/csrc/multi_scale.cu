// multi_scale.cu -- (next row) the merge step of multi-scale / flip testing on the device:
//   per class, gather the detections of the K (scale, flip) passes of ONE image (multi_eval.py:557-640: rows with
//   score > 0, un-flip `x1' = 1 - x2, x2' = 1 - x1`, scale to pixels, per-scale size rule), then bbox_vote
//   (multi_eval.py:453-494): visit by descending score, merge every remaining box whose overlap with the head is >= 0.45
//   into one score-weighted box carrying the head's score.
// The reference does this on the host in NumPy float32 after B*C device->host copies; here one CTA per class works on the
// K Detect outputs where they are.  Arithmetic order is NumPy's (the oracle restatement is NumPy itself): IoU with the +1
// convention evaluated operation by operation, coordinate sums row after row, the score sum with NumPy's pairwise scheme
// (oracle/multi_scale_ref.py), float32 division -- results are bit-identical to the restatement.
// Latency-bound by nature (a sequential merge over at most K*top_k boxes per class); it replaces host work, not a hot kernel.
#include "common.cuh"
#include <math.h>

namespace tdrn {

constexpr int MS_THREADS = 256;

struct MsP {
    const float *dets;           // [K, C, top_k, 5] (score, x1, y1, x2, y2) normalised
    const int *flip;             // [K]
    const int *rule;             // [K]  0: longer side > thr   1: shorter side < thr
    const float *rule_thr;       // [K]
    int K, C, top_k, cap;        // cap = K * top_k candidates per class at most
    float w, h, thr;             // image size, vote threshold (float32(0.45))
    float4 *box;                 // workspace [C, cap]
    float *score;                // workspace [C, cap]
    unsigned long long *keys;    // workspace [C, cap_pow2]
    int *member;                 // workspace [C, cap]
    unsigned char *alive;        // workspace [C, cap]
    int cap_pow2;
    float *out;                  // [C, max_out, 5] (x1, y1, x2, y2, score)
    int *out_count;              // [C]
    int max_out;
};

// NumPy's float32 pairwise sum (numpy/_core/src/umath/loops_utils.h.src, *_pairwise_sum), see oracle/multi_scale_ref.py
__device__ float ms_pairwise(const float *a, const int *idx, int n)
{
    if (n < 8) {
        float r = 0.f;
        for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[idx[i]]);
        return r;
    }
    if (n <= 128) {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = a[idx[j]];
        int i = 8;
        for (; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[idx[i + j]]);
        }
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __fadd_rn(res, a[idx[i]]);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(ms_pairwise(a, idx, n2), ms_pairwise(a, idx + n2, n - n2));
}

__global__ void __launch_bounds__(MS_THREADS) multiscale_vote_kernel(const MsP p)
{
    __shared__ unsigned warp_cnt[MS_THREADS / 32];
    __shared__ int s_n, s_m, s_head, s_out;
    const int cls = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (cls == 0) { if (tid == 0) p.out_count[0] = 0; return; }          // background: nothing (multi_eval.py:557 starts at 1)
    float4 *box = p.box + (size_t)cls * p.cap;
    float *score = p.score + (size_t)cls * p.cap;
    unsigned long long *keys = p.keys + (size_t)cls * p.cap_pow2;
    int *member = p.member + (size_t)cls * p.cap;
    unsigned char *alive = p.alive + (size_t)cls * p.cap;
    float *out = p.out + (size_t)cls * p.max_out * 5;

    // block-wide ordered compaction of a predicate: returns this thread's slot (or -1) and adds the count to *total
    auto compact = [&](bool take, int *total) -> int {
        const unsigned ballot = __ballot_sync(0xffffffffu, take);
        if (lane == 0) warp_cnt[warp] = __popc(ballot);
        __syncthreads();
        int base = *total;
        for (int wv = 0; wv < warp; ++wv) base += warp_cnt[wv];
        const int slot = take ? base + __popc(ballot & ((1u << lane) - 1u)) : -1;
        __syncthreads();
        if (tid == 0) { int s = 0; for (int wv = 0; wv < MS_THREADS / 32; ++wv) s += warp_cnt[wv]; *total += s; }
        __syncthreads();
        return slot;
    };

    // ---- 1. gather, pass after pass, rank after rank (the reference's concatenation order) --------------------------
    if (tid == 0) s_n = 0;
    __syncthreads();
    for (int k = 0; k < p.K; ++k) {
        const float *d = p.dets + ((size_t)k * p.C + cls) * p.top_k * 5;
        const int fl = p.flip[k], rule = p.rule[k];
        const float rthr = p.rule_thr[k];
        for (int t0 = 0; t0 < p.top_k; t0 += MS_THREADS) {
            const int t = t0 + tid;
            bool take = false;
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            float sc = 0.f;
            if (t < p.top_k) {
                sc = d[t * 5];
                if (sc > 0.f) {                                             // :563
                    float x1 = d[t * 5 + 1], y1 = d[t * 5 + 2], x2 = d[t * 5 + 3], y2 = d[t * 5 + 4];
                    if (fl) { const float nx0 = __fsub_rn(1.f, x1), nx2 = __fsub_rn(1.f, x2); x1 = nx2; x2 = nx0; }   // :566-571
                    x1 = __fmul_rn(x1, p.w); x2 = __fmul_rn(x2, p.w); y1 = __fmul_rn(y1, p.h); y2 = __fmul_rn(y2, p.h);
                    const float sw = __fadd_rn(__fsub_rn(x2, x1), 1.f), sh = __fadd_rn(__fsub_rn(y2, y1), 1.f);
                    take = rule == 0 ? fmaxf(sw, sh) > rthr : fminf(sw, sh) < rthr;                                 // :574-625
                    b = make_float4(x1, y1, x2, y2);
                }
            }
            const int slot = compact(take, &s_n);
            if (slot >= 0) { box[slot] = b; score[slot] = sc; }
        }
    }
    __syncthreads();
    const int n = s_n;
    if (n == 0) { if (tid == 0) p.out_count[cls] = 0; return; }

    // ---- 2. order: descending score, ties -> lower index (bitonic sort of (score bits, ~index) keys) ----------------
    int npad = 1;
    while (npad < n) npad <<= 1;
    for (int i = tid; i < npad; i += MS_THREADS) {
        unsigned long long key = 0ull;
        if (i < n) key = ((unsigned long long)__float_as_uint(score[i]) << 32) | (unsigned)(0xffffffffu - (unsigned)i);   // score > 0
        keys[i] = key;
    }
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (npad >> 1); t += MS_THREADS) {
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;
                const unsigned long long a = keys[lo], c = keys[hi];
                const bool desc = (lo & k) == 0;
                if (desc ? (a < c) : (a > c)) { keys[lo] = c; keys[hi] = a; }
            }
            __syncthreads();
        }
    for (int i = tid; i < n; i += MS_THREADS) alive[i] = 1;
    if (tid == 0) { s_head = 0; s_out = 0; }
    __syncthreads();
    auto src_of = [&](int i) -> int { return (int)(0xffffffffu - (unsigned)(keys[i] & 0xffffffffull)); };
    auto area_of = [&](float4 b) -> float { return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f)); };

    // ---- 3. bbox_vote ------------------------------------------------------------------------------------------------
    if (n == 1) {                                                            // :454-455: returned unchanged
        if (tid == 0) {
            const float4 b = box[0];
            out[0] = b.x; out[1] = b.y; out[2] = b.z; out[3] = b.w; out[4] = score[0];
            p.out_count[cls] = 1;
        }
        return;
    }
    while (true) {
        const int head = s_head;
        if (head >= n) break;
        const float4 hb = box[src_of(head)];
        const float ha = area_of(hb);
        if (tid == 0) s_m = 0;
        __syncthreads();
        for (int i0 = head; i0 < n; i0 += MS_THREADS) {
            const int i = i0 + tid;
            bool mem = false;
            if (i < n && alive[i]) {
                const float4 b = box[src_of(i)];
                const float xx1 = fmaxf(hb.x, b.x), yy1 = fmaxf(hb.y, b.y), xx2 = fminf(hb.z, b.z), yy2 = fminf(hb.w, b.w);
                const float iw = fmaxf(0.f, __fadd_rn(__fsub_rn(xx2, xx1), 1.f)), ih = fmaxf(0.f, __fadd_rn(__fsub_rn(yy2, yy1), 1.f));
                const float inter = __fmul_rn(iw, ih);
                const float o = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ha, area_of(b)), inter));
                mem = o >= p.thr;                                            // false for NaN, as in NumPy
            }
            const int slot = compact(mem, &s_m);
            if (slot >= 0) member[slot] = i;
        }
        __syncthreads();
        int m = s_m;
        if (m == 0) { if (tid == 0) { member[0] = head; s_m = 1; } m = 1; __syncthreads(); }    // degenerate head: emit it alone
        for (int q = tid; q < m; q += MS_THREADS) alive[member[q]] = 0;
        __syncthreads();                                                     // member[] fully read, alive[] updated
        if (tid == 0) {
            const int o_idx = s_out;
            if (o_idx < p.max_out) {
                float *row = out + (size_t)o_idx * 5;
                if (m == 1) {
                    const int s0 = src_of(member[0]);
                    const float4 b = box[s0];
                    row[0] = b.x; row[1] = b.y; row[2] = b.z; row[3] = b.w; row[4] = score[s0];
                } else {
                    // coordinates: (x * score) summed row after row from the first member; scores: NumPy pairwise
                    float sx1 = 0.f, sy1 = 0.f, sx2 = 0.f, sy2 = 0.f, mx = 0.f;
                    for (int q = 0; q < m; ++q) {
                        const int s0 = src_of(member[q]);
                        const float4 b = box[s0];
                        const float sc = score[s0];
                        const float a = __fmul_rn(b.x, sc), bb = __fmul_rn(b.y, sc), c = __fmul_rn(b.z, sc), dd = __fmul_rn(b.w, sc);
                        if (q == 0) { sx1 = a; sy1 = bb; sx2 = c; sy2 = dd; mx = sc; }
                        else { sx1 = __fadd_rn(sx1, a); sy1 = __fadd_rn(sy1, bb); sx2 = __fadd_rn(sx2, c); sy2 = __fadd_rn(sy2, dd); mx = fmaxf(mx, sc); }
                        member[q] = s0;                                      // from here on: index into score[]
                    }
                    const float ssum = ms_pairwise(score, member, m);
                    row[0] = __fdiv_rn(sx1, ssum); row[1] = __fdiv_rn(sy1, ssum); row[2] = __fdiv_rn(sx2, ssum); row[3] = __fdiv_rn(sy2, ssum);
                    row[4] = mx;
                }
            }
            s_out = o_idx + 1;
            int nh = head;
            while (nh < n && !alive[nh]) ++nh;                               // alive[] of this round was cleared by the loop above
            s_head = nh;
        }
        __syncthreads();
    }
    if (tid == 0) p.out_count[cls] = s_out;
}

}  // namespace tdrn

using namespace tdrn;

static size_t ms_align(size_t v) { return (v + 255) & ~(size_t)255; }
static int ms_pow2(int n) { int p = 1; while (p < n) p <<= 1; return p; }

extern "C" size_t tdrn_multiscale_vote_workspace_bytes(int K, int C, int top_k)
{
    const size_t cap = (size_t)K * top_k;
    return ms_align((size_t)C * cap * 16) + ms_align((size_t)C * cap * 4) + ms_align((size_t)C * ms_pow2((int)cap) * 8) +
           ms_align((size_t)C * cap * 4) + ms_align((size_t)C * cap);
}

extern "C" int tdrn_multiscale_vote(const float *dets, const int *flip, const int *rule, const float *rule_thr, int K, int C,
                                    int top_k, float w, float h, float vote_thresh, float *out, int *out_count, int max_out,
                                    void *workspace, size_t workspace_bytes, tdrn_stream_t stream)
{
    TDRN_REQUIRE(dets && flip && rule && rule_thr && out && out_count && workspace, "tdrn_multiscale_vote: null argument");
    TDRN_REQUIRE(K > 0 && C > 0 && top_k > 0 && max_out > 0, "tdrn_multiscale_vote: bad shape");
    TDRN_REQUIRE(workspace_bytes >= tdrn_multiscale_vote_workspace_bytes(K, C, top_k), "tdrn_multiscale_vote: workspace too small");
    MsP p{};
    p.dets = dets; p.flip = flip; p.rule = rule; p.rule_thr = rule_thr; p.K = K; p.C = C; p.top_k = top_k; p.cap = K * top_k;
    p.w = w; p.h = h; p.thr = vote_thresh; p.out = out; p.out_count = out_count; p.max_out = max_out;
    p.cap_pow2 = ms_pow2(p.cap);
    unsigned char *ws = (unsigned char *)workspace;
    const size_t cap = (size_t)p.cap;
    p.box = (float4 *)ws; ws += ms_align((size_t)C * cap * 16);
    p.score = (float *)ws; ws += ms_align((size_t)C * cap * 4);
    p.keys = (unsigned long long *)ws; ws += ms_align((size_t)C * p.cap_pow2 * 8);
    p.member = (int *)ws; ws += ms_align((size_t)C * cap * 4);
    p.alive = ws;
    multiscale_vote_kernel<<<C, MS_THREADS, 0, as_stream(stream)>>>(p);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}
