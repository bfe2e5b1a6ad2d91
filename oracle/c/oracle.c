/*
 * oracle.c -- plain-C scalar restatement of the native pieces of the TDRN hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Built by oracle/build.py into
 * oracle/_build/liboracle.so; loaded by tests, smoke() and bench.py's cpu_baseline leg.
 *
 * Follows (paths relative to the reference checkout):
 *   oracle_bilinear            utils/deformconv/deform_conv_cuda_kernel.cu:16-51
 *   oracle_deform_im2col       utils/deformconv/deform_conv_cuda_kernel.cu:157-208
 *   oracle_deform_conv_forward utils/deformconv/deform_conv_cuda.c:141-193
 *   oracle_decode              layers/box_utils.py:176-195, :16-25; layers/functions/detection.py:43-48
 *   oracle_cpu_nms             utils/nms/cpu_nms.pyx:17-68
 *   oracle_detect              layers/functions/detection.py:37-63
 *   oracle_prior_box           layers/functions/prior_box.py:33-64
 *
 * Build flags must keep IEEE semantics: -O2 -ffp-contract=off, no -ffast-math.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- sampler: .cu:16-51.  data points at (h_in, w_in) of one channel plane. ---------------- */
static float oracle_bilinear(const float *data, int data_width, int height, int width, float h, float w)
{
    int h_low = (int)floorf(h);
    int w_low = (int)floorf(w);
    int h_high, w_high;
    if (h_low >= height - 1) { h_high = h_low = height - 1; h = (float)h_low; } else { h_high = h_low + 1; }
    if (w_low >= width - 1)  { w_high = w_low = width - 1;  w = (float)w_low; } else { w_high = w_low + 1; }
    float lh = h - h_low, lw = w - w_low;
    float hh = 1 - lh, hw = 1 - lw;
    float v1 = data[h_low * data_width + w_low];
    float v2 = data[h_low * data_width + w_high];
    float v3 = data[h_high * data_width + w_low];
    float v4 = data[h_high * data_width + w_high];
    float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
    return (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
}

/* ---- im2col for one sample: .cu:157-208.  im [C,H,W], off [dg*2*kh*kw,Ho,Wo] -> col [C*kh*kw, Ho*Wo] */
void oracle_deform_im2col(const float *im, const float *off, int C, int H, int W, int kh, int kw,
                          int pad, int stride, int dil, int dg, int Ho, int Wo, float *col)
{
    int cpg = C / dg;
    for (int c_im = 0; c_im < C; ++c_im)
        for (int h_col = 0; h_col < Ho; ++h_col)
            for (int w_col = 0; w_col < Wo; ++w_col) {
                int g = c_im / cpg;
                int h_in = h_col * stride - pad;
                int w_in = w_col * stride - pad;
                float *col_ptr = col + ((size_t)(c_im * kh * kw) * Ho + h_col) * Wo + w_col;
                const float *im_ptr = im + ((size_t)c_im * H + h_in) * W + w_in;   /* may point before row 0 */
                const float *off_ptr = off + (size_t)g * 2 * kh * kw * Ho * Wo;
                for (int i = 0; i < kh; ++i)
                    for (int j = 0; j < kw; ++j) {
                        float offset_h = off_ptr[((size_t)(2 * (i * kw + j)) * Ho + h_col) * Wo + w_col];
                        float offset_w = off_ptr[((size_t)(2 * (i * kw + j) + 1) * Ho + h_col) * Wo + w_col];
                        float val = 0.f;
                        float h_im = h_in + i * dil + offset_h;
                        float w_im = w_in + j * dil + offset_w;
                        if (h_im >= 0 && w_im >= 0 && h_im < H && w_im < W) {
                            float map_h = i * dil + offset_h;
                            float map_w = j * dil + offset_w;
                            val = oracle_bilinear(im_ptr, W, H - h_in, W - w_in, map_h, map_w);
                        }
                        *col_ptr = val;
                        col_ptr += (size_t)Ho * Wo;
                    }
            }
}

/* ---- forward: deform_conv_cuda.c:157-193.  out[b] = weight[Cout, C*kh*kw] x col (beta = 0 after zero). */
int oracle_deform_conv_forward(const float *in, const float *off, const float *weight, float *out,
                               int B, int C, int H, int W, int Cout, int kh, int kw,
                               int stride, int pad, int dil, int dg)
{
    int Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) / stride + 1;
    int Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) / stride + 1;
    if (Ho <= 0 || Wo <= 0 || C % dg) return -1;
    size_t K = (size_t)C * kh * kw, HW = (size_t)Ho * Wo;
    float *col = (float *)malloc(K * HW * sizeof(float));
    if (!col) return -2;
    for (int b = 0; b < B; ++b) {
        oracle_deform_im2col(in + (size_t)b * C * H * W, off + (size_t)b * dg * 2 * kh * kw * HW,
                             C, H, W, kh, kw, pad, stride, dil, dg, Ho, Wo, col);
        float *o = out + (size_t)b * Cout * HW;
        memset(o, 0, (size_t)Cout * HW * sizeof(float));
        for (int m = 0; m < Cout; ++m)
            for (size_t k = 0; k < K; ++k) {
                float wv = weight[m * K + k];
                const float *cr = col + k * HW;
                float *orow = o + m * HW;
                for (size_t n = 0; n < HW; ++n) orow[n] += wv * cr[n];
            }
    }
    free(col);
    return 0;
}

/* ---- decode: box_utils.py:190-195 then center_size :24-25, fp32, same operation order ------- */
static void decode_one(const float *loc, const float *pr, float *box)
{
    const float v0 = 0.1f, v1 = 0.2f;
    float cx = pr[0] + loc[0] * v0 * pr[2];
    float cy = pr[1] + loc[1] * v0 * pr[3];
    float w = pr[2] * expf(loc[2] * v1);
    float h = pr[3] * expf(loc[3] * v1);
    float x1 = cx - w / 2, y1 = cy - h / 2;
    box[0] = x1; box[1] = y1; box[2] = w + x1; box[3] = h + y1;
}

/* loc [B,P,4], priors [P,4], arm_loc [B,P,4] or NULL -> boxes [B,P,4] (detection.py:43-48) */
void oracle_decode(const float *loc, const float *priors, const float *arm_loc, int B, int P, float *boxes)
{
    for (int b = 0; b < B; ++b)
        for (int p = 0; p < P; ++p) {
            size_t o = ((size_t)b * P + p) * 4;
            float def[4];
            if (arm_loc) {
                float t[4];
                decode_one(arm_loc + o, priors + p * 4, t);
                def[0] = (t[2] + t[0]) / 2; def[1] = (t[3] + t[1]) / 2;
                def[2] = t[2] - t[0];       def[3] = t[3] - t[1];
            } else {
                memcpy(def, priors + p * 4, sizeof def);
            }
            decode_one(loc + o, def, boxes + o);
        }
}

/* ---- NMS: cpu_nms.pyx:17-68 with the pinned tie rule (stable descending) -------------------- */
typedef struct { float s; int i; } sidx_t;
static int cmp_desc_stable(const void *a, const void *b)
{
    const sidx_t *x = (const sidx_t *)a, *y = (const sidx_t *)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return (x->i > y->i) - (x->i < y->i);
}

/* dets [N,5] (x1,y1,x2,y2,score).  keep_out needs room for min(N,max_keep) ints.  Returns #kept. */
int oracle_cpu_nms(const float *dets, int n, double thresh, int max_keep, int *keep_out)
{
    if (n <= 0) return 0;
    if (max_keep <= 0 || max_keep > n) max_keep = n;
    float *areas = (float *)malloc(sizeof(float) * n);
    sidx_t *ord = (sidx_t *)malloc(sizeof(sidx_t) * n);
    unsigned char *supp = (unsigned char *)calloc(n, 1);
    for (int i = 0; i < n; ++i) {
        const float *d = dets + 5 * (size_t)i;
        areas[i] = (d[2] - d[0] + 1) * (d[3] - d[1] + 1);          /* pyx:24 */
        ord[i].s = d[4]; ord[i].i = i;
    }
    qsort(ord, n, sizeof(sidx_t), cmp_desc_stable);                  /* pyx:25 */
    int nk = 0;
    for (int _i = 0; _i < n; ++_i) {
        int i = ord[_i].i;
        if (supp[i]) continue;
        keep_out[nk++] = i;
        if (nk >= max_keep) break;
        const float *di = dets + 5 * (size_t)i;
        float ix1 = di[0], iy1 = di[1], ix2 = di[2], iy2 = di[3], iarea = areas[i];
        for (int _j = _i + 1; _j < n; ++_j) {
            int j = ord[_j].i;
            if (supp[j]) continue;
            const float *dj = dets + 5 * (size_t)j;
            float xx1 = ix1 >= dj[0] ? ix1 : dj[0];
            float yy1 = iy1 >= dj[1] ? iy1 : dj[1];
            float xx2 = ix2 <= dj[2] ? ix2 : dj[2];
            float yy2 = iy2 <= dj[3] ? iy2 : dj[3];
            float w = xx2 - xx1 + 1; if (!(w >= 0.0f)) w = 0.0f;       /* max(0.0, .) pyx:61 */
            float h = yy2 - yy1 + 1; if (!(h >= 0.0f)) h = 0.0f;
            float inter = w * h;
            float ovr = inter / (iarea + areas[j] - inter);
            if ((double)ovr >= thresh) supp[j] = 1;                   /* pyx:65: float vs double */
        }
    }
    free(areas); free(ord); free(supp);
    return nk;
}

/* ---- Detect given decoded boxes: detection.py:37-63 ------------------------------------------
 * boxes [B,P,4] normalised, conf [B*P,C], scale[4] -> out [B,C,top_k,5] (zero-initialised here). */
int oracle_detect(const float *boxes, const float *conf, const float *scale, int B, int P, int C,
                  int top_k, float conf_thresh, double nms_thresh, float *out)
{
    memset(out, 0, sizeof(float) * (size_t)B * C * top_k * 5);
    float *dets = (float *)malloc(sizeof(float) * 5 * (size_t)P);
    int *src = (int *)malloc(sizeof(int) * (size_t)P);
    int *keep = (int *)malloc(sizeof(int) * (size_t)P);
    if (!dets || !src || !keep) return -2;
    for (int b = 0; b < B; ++b)
        for (int cl = 1; cl < C; ++cl) {
            int n = 0;
            for (int p = 0; p < P; ++p) {
                float s = conf[((size_t)b * P + p) * C + cl];
                if (s > conf_thresh) {                                /* strict, :53 */
                    const float *bx = boxes + ((size_t)b * P + p) * 4;
                    dets[5 * n + 0] = bx[0] * scale[0]; dets[5 * n + 1] = bx[1] * scale[1];
                    dets[5 * n + 2] = bx[2] * scale[2]; dets[5 * n + 3] = bx[3] * scale[3];
                    dets[5 * n + 4] = s;
                    src[n++] = p;
                }
            }
            if (n == 0) continue;
            int nk = oracle_cpu_nms(dets, n, nms_thresh, top_k, keep);
            for (int r = 0; r < nk && r < top_k; ++r) {
                int p = src[keep[r]];
                float *o = out + (((size_t)b * C + cl) * top_k + r) * 5;
                const float *bx = boxes + ((size_t)b * P + p) * 4;
                o[0] = conf[((size_t)b * P + p) * C + cl];
                o[1] = bx[0]; o[2] = bx[1]; o[3] = bx[2]; o[4] = bx[3];
            }
        }
    free(dets); free(src); free(keep);
    return 0;
}

/* ---- PriorBox: prior_box.py:33-64 (float64 arithmetic, one fp32 conversion, clamp) -----------
 * Boxes per cell: [s,s], (if n_max) [sqrt(s*s'),..], per ar: [s*sqrt(ar), s/sqrt(ar)], flip.  Returns P. */
int oracle_prior_box(int image_size, int n_levels, const int *feature_maps, const int *steps,
                     const double *min_sizes, const double *max_sizes /* or NULL */,
                     const int *n_ar, const double *ars /* [n_levels][4] */, int flip, int clip, float *out)
{
    int n = 0;
    for (int k = 0; k < n_levels; ++k) {
        int f = feature_maps[k];
        for (int i = 0; i < f; ++i)
            for (int j = 0; j < f; ++j) {
                double f_k = (double)image_size / steps[k];
                double cx = (j + 0.5) / f_k, cy = (i + 0.5) / f_k;
                double s_k = min_sizes[k] / image_size;
                double b[16][2]; int nb = 0;
                b[nb][0] = s_k; b[nb][1] = s_k; nb++;
                if (max_sizes) {
                    double sp = sqrt(s_k * (max_sizes[k] / image_size));
                    b[nb][0] = sp; b[nb][1] = sp; nb++;
                }
                for (int a = 0; a < n_ar[k]; ++a) {
                    double r = sqrt(ars[k * 4 + a]);
                    b[nb][0] = s_k * r; b[nb][1] = s_k / r; nb++;
                    if (flip) { b[nb][0] = s_k / r; b[nb][1] = s_k * r; nb++; }
                }
                for (int q = 0; q < nb; ++q) {
                    float v[4] = { (float)cx, (float)cy, (float)b[q][0], (float)b[q][1] };
                    for (int t = 0; t < 4; ++t) {
                        if (clip) { if (v[t] > 1.f) v[t] = 1.f; if (v[t] < 0.f) v[t] = 0.f; }
                        if (out) out[(size_t)n * 4 + t] = v[t];
                    }
                    n++;
                }
            }
    }
    return n;
}
