// api.cu -- library-wide plumbing of libtdrn_b200.so (error string, launch counter) and the one
// host-only entry point, PriorBox.
#include "common.cuh"
#include <atomic>
#include <math.h>
#include <string.h>
#include <stdlib.h>

namespace tdrn {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool pdl_enabled()
{
    static const bool on = [] { const char *e = getenv("TDRN_PDL"); return e && e[0] == '1'; }();
    return on;
}

}  // namespace tdrn

using namespace tdrn;

extern "C" const char *tdrn_last_error(void) { return g_err; }
extern "C" int tdrn_version(void) { return 100; }
extern "C" long long tdrn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// PriorBox.forward, layers/functions/prior_box.py:33-64.  The reference accumulates Python floats
// (IEEE double) and converts the whole list to fp32 once (torch.Tensor(mean), :61), then clamps
// (:62-63).  Same operations, same order, in C double -> bit-identical fp32 output.
extern "C" int tdrn_prior_box(int image_size, int n_levels, const int *feature_maps, const int *steps,
                              const double *min_sizes, const double *max_sizes, const int *n_ar, const double *ars,
                              int flip, int clip, float *out, int *num_priors)
{
    TDRN_REQUIRE(image_size > 0 && n_levels > 0 && feature_maps && steps && min_sizes && n_ar && ars && num_priors,
                 "tdrn_prior_box: bad argument");
    long long n = 0;
    for (int k = 0; k < n_levels; ++k) {
        TDRN_REQUIRE(feature_maps[k] > 0 && steps[k] > 0 && n_ar[k] >= 0 && n_ar[k] <= 4, "tdrn_prior_box: bad level %d", k);
        const int f = feature_maps[k];
        const double f_k = (double)image_size / (double)steps[k];           // :39
        const double s_k = min_sizes[k] / (double)image_size;               // :46
        double bw[10], bh[10];
        int nb = 0;
        bw[nb] = s_k; bh[nb] = s_k; ++nb;                                   // :47
        if (max_sizes) {                                                    // :51-53
            const double sp = sqrt(s_k * (max_sizes[k] / (double)image_size));
            bw[nb] = sp; bh[nb] = sp; ++nb;
        }
        for (int a = 0; a < n_ar[k]; ++a) {                                 // :56-59
            const double r = sqrt(ars[k * 4 + a]);
            bw[nb] = s_k * r; bh[nb] = s_k / r; ++nb;
            if (flip) { bw[nb] = s_k / r; bh[nb] = s_k * r; ++nb; }
        }
        for (int i = 0; i < f; ++i)
            for (int j = 0; j < f; ++j) {
                const double cx = ((double)j + 0.5) / f_k, cy = ((double)i + 0.5) / f_k;   // :41-42
                for (int q = 0; q < nb; ++q, ++n) {
                    if (!out) continue;
                    float v[4] = {(float)cx, (float)cy, (float)bw[q], (float)bh[q]};
                    for (int t = 0; t < 4; ++t) {
                        if (clip) v[t] = v[t] > 1.f ? 1.f : (v[t] < 0.f ? 0.f : v[t]);
                        out[n * 4 + t] = v[t];
                    }
                }
            }
    }
    *num_priors = (int)n;
    return TDRN_OK;
}
