// post.cuh -- device code shared by the post-processing translation units (post.cu, bloom.cu):
// the lens flare evaluated per pixel and the flare parameters both composite kernels consume.
#pragma once
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// lens flare, render.py:3925-4028 (float64 like the numpy original, f32 accumulation of `flare`)
// ---------------------------------------------------------------------------------------------
struct FlareParams {
    int enabled;
    double light_x, light_y, scx, scy, scale, intensity, streak_alpha, streak_len;
};

static __device__ __forceinline__ double np_mod(double a, double b) {   // numpy.mod for b > 0
    double m = fmod(a, b);
    if (m != 0.0 && m < 0.0) m += b;
    return m;
}

// Every term is zero for most pixels; squared-distance / slope pre-tests skip the sqrt / atan2 /
// exp of terms that cannot contribute (a skipped term adds exactly 0, as in numpy).  The streak
// mask is a discontinuity and is decided in float64 exactly like the reference.
static __device__ __noinline__ void flare_pixel(const FlareParams& F, int x, int y, float fl[3]) {
    const double PI = 3.14159265358979323846;
    const double GUARD = 1.0 + 1e-9;
    fl[0] = fl[1] = fl[2] = 0.0f;
    // Ghosts, rings and the hexagon are continuous functions of the pixel position: centres are
    // formed in float64, the per-pixel distance algebra runs in float32 (error ~1e-6 of a term
    // <= 1.5, far inside the 2/255 gate); the `flare` accumulation keeps numpy's f64-add/f32-store.
    const double gc[3] = {1.0, 0.9, 0.7};
    const float inten = (float)F.intensity, fscale = (float)F.scale;
    for (int g = 0; g < 8; ++g) {
        double t = (g + 1) * 0.15;
        double gx = F.light_x + (F.scx - F.light_x) * t, gy = F.light_y + (F.scy - F.light_y) * t;
        float size = (float)(25 + g * 30) * fscale;
        float dx = (float)(x - gx), dy = (float)(y - gy);
        float d2 = dx * dx + dy * dy;
        if (d2 >= size * size) continue;
        float u = 1.0f - sqrtf(d2) / size;
        float alpha = u * u * (float)(1 - g * 0.08) * inten;
        for (int c = 0; c < 3; ++c) fl[c] = (float)((double)fl[c] + (double)alpha * gc[c]);
    }
    const double rc[3][3] = {{0.3, 0.4, 1.0}, {0.5, 0.5, 0.9}, {0.7, 0.5, 0.8}};
    for (int k = 0; k < 3; ++k) {
        double t = 0.35 + k * 0.15;
        double rx = F.light_x + (F.scx - F.light_x) * t, ry = F.light_y + (F.scy - F.light_y) * t;
        float rr = (float)(60 + k * 40) * fscale, rw = (float)(6 + k * 3) * fscale;
        float dx = (float)(x - rx), dy = (float)(y - ry);
        float d2 = dx * dx + dy * dy;
        float lo = rr - rw, hi = rr + rw;
        if (d2 >= hi * hi || (lo > 0.0f && d2 <= lo * lo)) continue;
        float u = fminf(fmaxf(1.0f - fabsf(sqrtf(d2) - rr) / rw, 0.0f), 1.0f);
        double ra = (double)(u * u * 0.5f * inten * (float)(1 - k * 0.25));
        for (int c = 0; c < 3; ++c) fl[c] = (float)((double)fl[c] + ra * rc[k][c]);
    }
    {
        const double hc[3] = {0.6, 0.7, 1.0};
        double hx = F.light_x + (F.scx - F.light_x) * 0.5, hy = F.light_y + (F.scy - F.light_y) * 0.5;
        float hr = 100.0f * fscale, hw = 15.0f * fscale;
        float dx = (float)(x - hx), dy = (float)(y - hy);
        float d2 = dx * dx + dy * dy;
        float lo = hr - hw, hi = hr + hw;
        if (!(d2 >= hi * hi || (lo > 0.0f && d2 <= lo * lo))) {
            float angle = atan2f(dy, dx);
            float m = fmodf(angle, 1.0471976f);
            if (m < 0.0f) m += 1.0471976f;
            float hf = fminf(fmaxf(1.0f - fabsf(m - 0.5235988f) / 0.2f, 0.0f), 1.0f);
            float u = fminf(fmaxf(1.0f - fabsf(sqrtf(d2) - hr) / hw, 0.0f), 1.0f);
            double ra = (double)(u * u * hf * 0.3f * inten);
            for (int c = 0; c < 3; ++c) fl[c] = (float)((double)fl[c] + ra * hc[c]);
        }
    }
    {
        const double sc[3] = {1.0, 0.95, 0.9};
        const double main_angles[4] = {0.0, PI / 2, PI, 3 * PI / 2};
        double dx = x - F.light_x, dy = y - F.light_y;
        // within 0.05 rad of one of the four axes through the light?  tan(0.05) = 0.050041708
        const double T = 0.0500418 * GUARD;
        if (fabs(dy) <= T * fabs(dx) || fabs(dx) <= T * fabs(dy)) {
            double dist = sqrt(dx * dx + dy * dy);
            double angle = atan2(dy, dx);
            double falloff = exp(-dist / F.streak_len);
            for (int a = 0; a < 4; ++a) {
                double diff = fabs(np_mod(angle - main_angles[a] + PI, 2 * PI) - PI);
                for (int c = 0; c < 3; ++c) {
                    double add = diff < 0.05 ? falloff * F.streak_alpha * sc[c] : 0.0;
                    fl[c] = (float)((double)fl[c] + add);
                }
            }
        }
    }
}


}  // namespace
