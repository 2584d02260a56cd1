// nfk_knots.cuh -- arithmetic of the knot-table kernel (see nfk_knots.cu), host/device so that the
// CPU test harness (tests/cpu_harness) runs the very same code.
#pragma once
#include <math.h>
#include "nfk_math.cuh"

namespace nfk {

constexpr int kMaxKnots = 256;

struct KnotArgs {
    const float *wx, *wy, *wd;
    int K;
    float xlo, xw, ylo, yw;
};

NFK_HD void knots_softmax(const float* w, int n, double* p) {
    double m = w[0];
    for (int i = 1; i < n; ++i) m = fmax(m, (double)w[i]);
    double z = 0.0;
    for (int i = 0; i < n; ++i) { p[i] = exp((double)w[i] - m); z += p[i]; }
    for (int i = 0; i < n; ++i) p[i] /= z;
}

NFK_HD void knots_fwd_body(const KnotArgs& a, float* table) {
    const int K = a.K, n = K - 1;
    double p[kMaxKnots], q[kMaxKnots];
    knots_softmax(a.wx, n, p);
    knots_softmax(a.wy, n, q);
    double lx = 0.0, ly = 0.0;
    for (int j = 0; j < n; ++j) {
        table[0 * K + j] = (float)(a.xlo + a.xw * lx);
        table[1 * K + j] = (float)(a.ylo + a.yw * ly);
        lx += p[j];
        ly += q[j];
    }
    table[0 * K + n] = a.xlo + a.xw;
    table[1 * K + n] = a.ylo + a.yw;
    double rx = 0.0, ry = 0.0;
    table[3 * K + n] = 0.f;
    table[4 * K + n] = 0.f;
    for (int j = n - 1; j >= 0; --j) {
        rx += p[j];
        ry += q[j];
        table[3 * K + j] = (float)(a.xw * rx);
        table[4 * K + j] = (float)(a.yw * ry);
    }
    if (a.wd) {
        const double ln2 = 0.6931471805599453;
        for (int j = 0; j < K; ++j) {
            const double z = ln2 * (double)a.wd[j];
            table[2 * K + j] = (float)(z > 20.0 ? (double)a.wd[j] : log1p(exp(z)) / ln2);
        }
    } else {
        double prev = 0.0;
        for (int i = 0; i < n; ++i) {
            const double s = ((double)a.yw * q[i]) / ((double)a.xw * p[i]);
            table[2 * K + i] = (float)(i == 0 ? s : 0.5 * (s + prev));
            prev = s;
        }
        table[2 * K + n] = (float)prev;
    }
}

NFK_HD void knots_bwd_body(const KnotArgs& a, const float* g, float* gwx, float* gwy, float* gwd) {
    const int K = a.K, n = K - 1;
    double p[kMaxKnots], q[kMaxKnots], gp[kMaxKnots], gq[kMaxKnots];
    knots_softmax(a.wx, n, p);
    knots_softmax(a.wy, n, q);
    // prefix sums from the left knots (j > i, j <= n-1) and the right complements (j <= i)
    double sx = 0.0, sy = 0.0;
    for (int i = n - 1; i >= 0; --i) {          // sum_{j=i+1}^{n-1} g_kx[j]
        gp[i] = a.xw * sx;
        gq[i] = a.yw * sy;
        sx += g[0 * K + i];
        sy += g[1 * K + i];
    }
    sx = sy = 0.0;
    for (int i = 0; i < n; ++i) {               // sum_{j=0}^{i} g_cx[j]
        sx += g[3 * K + i];
        sy += g[4 * K + i];
        gp[i] += a.xw * sx;
        gq[i] += a.yw * sy;
    }
    if (a.wd) {
        const double ln2 = 0.6931471805599453;
        for (int j = 0; j < K; ++j) {
            const double z = ln2 * (double)a.wd[j];
            gwd[j] = (float)((double)g[2 * K + j] * (z > 20.0 ? 1.0 : 1.0 / (1.0 + exp(-z))));
        }
    } else {
        for (int i = 0; i < n; ++i) {
            const double s = ((double)a.yw * q[i]) / ((double)a.xw * p[i]);
            double gs = 0.0;
            if (i == 0) gs += g[2 * K + 0];
            if (i >= 1) gs += 0.5 * g[2 * K + i];
            if (i + 1 <= n - 1) gs += 0.5 * g[2 * K + i + 1];
            if (i == n - 1) gs += g[2 * K + n];
            gp[i] -= gs * s / p[i];
            gq[i] += gs * s / q[i];
        }
    }
    double dx = 0.0, dy = 0.0;
    for (int i = 0; i < n; ++i) { dx += p[i] * gp[i]; dy += q[i] * gq[i]; }
    for (int i = 0; i < n; ++i) {
        gwx[i] = (float)(p[i] * (gp[i] - dx));
        gwy[i] = (float)(q[i] * (gq[i] - dy));
    }
}

}  // namespace nfk
