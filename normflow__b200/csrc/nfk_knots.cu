// nfk_knots.cu -- SplineNet.make_spline (reference src/nn/scalar/modules.py:369-391) for one shared
// 1-D spline: raw parameters -> knot table, in ONE launch instead of ~25 tiny softmax / cumsum /
// cat launches (the parameter-only part of DistConvertor_, MeanFieldNet_ and IPSD; it dominated
// the launch-bound configurations).
//
//   table[0] = knots_x    = xlo + xw * sum_{i<j} p_i          p = softmax(weights_x)   (K-1 terms)
//   table[1] = knots_y    = ylo + yw * sum_{i<j} q_i          q = softmax(weights_y)
//   table[2] = knots_d    = softplus_{beta=ln2}(weights_d)    or, smooth (weights_d == NULL), the mean
//                           of the adjacent bin slopes s_i = (yw q_i)/(xw p_i), end slopes at the ends
//   table[3] = xlo + xw - knots_x, accumulated from the right end (sum_{i>=j} p_i), exact 0 at the end
//   table[4] = ylo + yw - knots_y, likewise
//
// K is a few tens at most: one thread walks the arrays in double precision.
#include "nfk_common.cuh"

#define NFK_STREAM(s) reinterpret_cast<cudaStream_t>(s)

namespace nfk {

constexpr int kMaxKnots = 256;

struct KnotArgs {
    const float *wx, *wy, *wd;
    int K;
    float xlo, xw, ylo, yw;
};

__device__ void knots_softmax(const float* w, int n, double* p) {
    double m = w[0];
    for (int i = 1; i < n; ++i) m = fmax(m, (double)w[i]);
    double z = 0.0;
    for (int i = 0; i < n; ++i) { p[i] = exp((double)w[i] - m); z += p[i]; }
    for (int i = 0; i < n; ++i) p[i] /= z;
}

__global__ void knots_fwd_kernel(KnotArgs a, float* __restrict__ table) {
    if (threadIdx.x != 0) return;
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

__global__ void knots_bwd_kernel(KnotArgs a, const float* __restrict__ g, float* __restrict__ gwx,
                                 float* __restrict__ gwy, float* __restrict__ gwd) {
    if (threadIdx.x != 0) return;
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

using namespace nfk;

static bool knots_ok(const float* wx, const float* wy, int K, float xw, float yw) {
    return wx && wy && K >= 2 && K <= kMaxKnots && xw > 0.f && yw > 0.f;
}

extern "C" int nfk_knots_fwd(const float* wx, const float* wy, const float* wd, int K, float xlo, float xw,
                             float ylo, float yw, float* table, void* stream) {
    if (!knots_ok(wx, wy, K, xw, yw) || !table) return NFK_EINVAL;
    KnotArgs a{wx, wy, wd, K, xlo, xw, ylo, yw};
    knots_fwd_kernel<<<1, 32, 0, NFK_STREAM(stream)>>>(a, table);
    return check_launch();
}

extern "C" int nfk_knots_bwd(const float* wx, const float* wy, const float* wd, int K, float xlo, float xw,
                             float ylo, float yw, const float* g_table, float* gwx, float* gwy, float* gwd,
                             void* stream) {
    if (!knots_ok(wx, wy, K, xw, yw) || !g_table || !gwx || !gwy || (wd && !gwd)) return NFK_EINVAL;
    KnotArgs a{wx, wy, wd, K, xlo, xw, ylo, yw};
    knots_bwd_kernel<<<1, 32, 0, NFK_STREAM(stream)>>>(a, g_table, gwx, gwy, gwd);
    return check_launch();
}
