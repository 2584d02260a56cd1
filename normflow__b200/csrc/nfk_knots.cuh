// nfk_knots.cuh -- arithmetic of the knot-table kernel (see nfk_knots.cu), host/device so that the
// CPU test harness (tests/cpu_harness) runs the very same code.
//
// The work is split between `nl` co-operating lanes (one warp on the device, a single "lane" on
// the host): exponentials and the final element-wise formulas are strided over the lanes, the
// running sums (a few hundred double additions at most) are done by lane 0.  `Sync` separates
// the phases (__syncwarp on the device, nothing on the host).  Scratch arrays live in `KnotScratch`
// (shared memory on the device).
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

struct KnotScratch {
    double p[kMaxKnots], q[kMaxKnots], gp[kMaxKnots], gq[kMaxKnots];
    double dx, dy;
};

// unnormalised softmax terms exp(w - max w), strided over the lanes (every lane finds the max itself)
NFK_HD void knots_exp_terms(const float* w, int n, double* e, int lane, int nl) {
    float m = w[0];
    for (int i = 1; i < n; ++i) m = fmaxf(m, w[i]);
    for (int i = lane; i < n; i += nl) e[i] = exp((double)w[i] - (double)m);
}

NFK_HD void knots_normalise(double* e, int n) {
    double z = 0.0;
    for (int i = 0; i < n; ++i) z += e[i];
    const double inv = 1.0 / z;
    for (int i = 0; i < n; ++i) e[i] *= inv;
}

NFK_HD double knots_softplus_ln2(double w) {
    const double ln2 = 0.6931471805599453;
    const double z = ln2 * w;
    return z > 20.0 ? w : log1p(exp(z)) / ln2;
}

NFK_HD double knots_bin_slope(const KnotArgs& a, const KnotScratch& s, int i) {
    return ((double)a.yw * s.q[i]) / ((double)a.xw * s.p[i]);
}

template <class Sync>
NFK_HD void knots_fwd_body(const KnotArgs& a, float* table, KnotScratch& s, int lane, int nl, Sync sync) {
    const int K = a.K, n = K - 1;
    knots_exp_terms(a.wx, n, s.p, lane, nl);
    knots_exp_terms(a.wy, n, s.q, lane, nl);
    sync();
    if (lane == 0) {
        knots_normalise(s.p, n);
        knots_normalise(s.q, n);
        double lx = 0.0, ly = 0.0;
        for (int j = 0; j < n; ++j) {
            table[0 * K + j] = (float)(a.xlo + a.xw * lx);
            table[1 * K + j] = (float)(a.ylo + a.yw * ly);
            lx += s.p[j];
            ly += s.q[j];
        }
        table[0 * K + n] = a.xlo + a.xw;
        table[1 * K + n] = a.ylo + a.yw;
        double rx = 0.0, ry = 0.0;
        table[3 * K + n] = 0.f;
        table[4 * K + n] = 0.f;
        for (int j = n - 1; j >= 0; --j) {
            rx += s.p[j];
            ry += s.q[j];
            table[3 * K + j] = (float)(a.xw * rx);
            table[4 * K + j] = (float)(a.yw * ry);
        }
    }
    sync();
    if (a.wd) {
        for (int j = lane; j < K; j += nl) table[2 * K + j] = (float)knots_softplus_ln2((double)a.wd[j]);
    } else {
        for (int j = lane; j < K; j += nl) {
            const double hi = knots_bin_slope(a, s, j < n ? j : n - 1);
            const double lo = knots_bin_slope(a, s, j > 0 ? j - 1 : 0);
            table[2 * K + j] = (float)(0.5 * (hi + lo));      // end knots: hi == lo == the end bin's slope
        }
    }
}

template <class Sync>
NFK_HD void knots_bwd_body(const KnotArgs& a, const float* g, float* gwx, float* gwy, float* gwd,
                           KnotScratch& s, int lane, int nl, Sync sync) {
    const int K = a.K, n = K - 1;
    knots_exp_terms(a.wx, n, s.p, lane, nl);
    knots_exp_terms(a.wy, n, s.q, lane, nl);
    sync();
    if (lane == 0) {
        knots_normalise(s.p, n);
        knots_normalise(s.q, n);
        // knots_x[j] (j <= n-1) holds p_i for i < j; the right complements [j] hold p_i for i >= j
        double sx = 0.0, sy = 0.0;
        for (int i = n - 1; i >= 0; --i) {
            s.gp[i] = a.xw * sx;
            s.gq[i] = a.yw * sy;
            sx += g[0 * K + i];
            sy += g[1 * K + i];
        }
        sx = sy = 0.0;
        for (int i = 0; i < n; ++i) {
            sx += g[3 * K + i];
            sy += g[4 * K + i];
            s.gp[i] += a.xw * sx;
            s.gq[i] += a.yw * sy;
        }
    }
    sync();
    if (a.wd) {
        const double ln2 = 0.6931471805599453;
        for (int j = lane; j < K; j += nl) {
            const double z = ln2 * (double)a.wd[j];
            gwd[j] = (float)((double)g[2 * K + j] * (z > 20.0 ? 1.0 : 1.0 / (1.0 + exp(-z))));
        }
    } else {
        for (int i = lane; i < n; i += nl) {
            const double sl = knots_bin_slope(a, s, i);
            double gs = 0.0;
            if (i == 0) gs += g[2 * K + 0];
            if (i >= 1) gs += 0.5 * g[2 * K + i];
            if (i + 1 <= n - 1) gs += 0.5 * g[2 * K + i + 1];
            if (i == n - 1) gs += g[2 * K + n];
            s.gp[i] -= gs * sl / s.p[i];
            s.gq[i] += gs * sl / s.q[i];
        }
    }
    sync();
    if (lane == 0) {
        double dx = 0.0, dy = 0.0;
        for (int i = 0; i < n; ++i) { dx += s.p[i] * s.gp[i]; dy += s.q[i] * s.gq[i]; }
        s.dx = dx;
        s.dy = dy;
    }
    sync();
    for (int i = lane; i < n; i += nl) {
        gwx[i] = (float)(s.p[i] * (s.gp[i] - s.dx));
        gwy[i] = (float)(s.q[i] * (s.gq[i] - s.dy));
    }
}

struct KnotNoSync {
    NFK_HD void operator()() const {}
};

}  // namespace nfk
