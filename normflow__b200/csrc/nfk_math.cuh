// nfk_math.cuh -- per-site arithmetic of the normflow hot path, fp32.
//
// Everything here is `__host__ __device__`: the CUDA kernels call it per thread,
// and tests/cpu_harness compiles the very same header with g++ so the arithmetic
// is checked against the oracle on a machine without a GPU (test infrastructure
// only -- the product never runs this on the host).
//
// Reference semantics (paths under /root/reference/src):
//   softplus(beta=ln2), softmax/cumsum knots ... nn/scalar/couplings_.py:173-176, 230-245
//   bin search (searchsorted right=False, clamp) lib/spline/spline.py:154-172
//   Pade[2,2] segment value / derivative ........ lib/spline/spline.py:185-220
//   inverse ..................................... lib/spline/spline.py:222-287 (stable root here)
//   linear / anti extrapolation ................. lib/spline/spline.py:458-532
//   Expit_ / Logit_ ............................. nn/scalar/modules_.py:93-114
//   affine ...................................... nn/scalar/couplings_.py:123-139
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define NFK_HD __host__ __device__ __forceinline__
#else
#define NFK_HD inline
#endif

namespace nfk {

constexpr float kLn2 = 0.69314718055994531f;
constexpr float kInvLn2 = 1.44269504088896341f;
constexpr float kLogSqrt2Pi = 0.91893853320467274f;

constexpr int kExtrapNone = 0;
constexpr int kExtrapLinear = 1;
constexpr int kExtrapAnti = 2;
constexpr int kExtrapPeriodic = 3;   // 1-D shared-knot splines only (spline.py:502-508, 518-524): even mirror image

// ---------------------------------------------------------------------------
// softplus with beta = ln 2:  log2(1 + 2^z); torch switches to the identity when
// beta*z > 20 (couplings_.py:173-176).
NFK_HD float softplus_ln2(float z) {
    if (z * kLn2 > 20.f) return z;
    return log1pf(exp2f(z)) * kInvLn2;
}
NFK_HD float softplus_ln2_grad(float z) {
    if (z * kLn2 > 20.f) return 1.f;
    const float t = exp2f(z);
    return t / (1.f + t);
}

// ---------------------------------------------------------------------------
// One spline segment [X0, X0+w] -> [Y0, Y0+h] with end derivatives D0, D1.
struct RqSeg {
    float X0, w, Y0, h, D0, D1;
};

// Adjoint of a segment (same parametrisation) plus the adjoint of the argument.
struct RqSegGrad {
    float gX0, gw, gY0, gh, gD0, gD1, gx;
};

// Pade[2,2] value and log-derivative (spline.py:204-212).
//   m = h/w, theta = (x - X0)/w, omega = 1 - theta
//   y    = Y0 + h theta (m theta + D0 omega) / (m + (D0 + D1 - 2m) theta omega)
//   dydx = m^2 (D1 theta^2 + 2 m theta omega + D0 omega^2) / (...)^2
// `m` = h / w is handed in so that callers who already hold 1 / w do not divide twice.
NFK_HD void rq_eval_theta_m(const RqSeg& s, float m, float th, float om, float& y, float& logg) {
    const float sig = s.D0 + s.D1 - 2.f * m;
    const float tom = th * om;
    const float rden = 1.f / (m + sig * tom);
    const float N = m * th * th + s.D0 * tom;
    y = s.Y0 + s.h * N * rden;
    const float Q = s.D1 * th * th + 2.f * m * tom + s.D0 * om * om;
    const float mr = m * rden;
    logg = logf(mr * mr * Q);
}
NFK_HD void rq_eval_theta(const RqSeg& s, float th, float om, float& y, float& logg) {
    rq_eval_theta_m(s, s.h / s.w, th, om, y, logg);
}

NFK_HD void rq_forward(const RqSeg& s, float x, float& y, float& logg) {
    const float rw = 1.f / s.w;
    const float th = (x - s.X0) * rw;
    rq_eval_theta_m(s, s.h * rw, th, 1.f - th, y, logg);
}

// theta solving  a2 theta^2 + a1 theta + a0 = 0  for eta = (y - Y0)/h
// (spline.py:243-271), written as 2 a0 / (-a1 + sqrt(disc)) which has no
// cancellation and no a2 == 0 special case.
NFK_HD float rq_theta_from_eta(float m, float D0, float D1, float eta) {
    const float sig = D0 + D1 - 2.f * m;
    const float a2 = -sig * eta + (D0 - m);
    const float a1 = -a2 - m;
    const float a0 = m * eta;
    float disc = a1 * a1 - 4.f * a0 * a2;
    disc = disc > 0.f ? disc : 0.f;
    return 2.f * a0 / (-a1 + sqrtf(disc));
}

// Inverse of rq_forward: x with y = g0(x), and loginv = -log g1(x).
NFK_HD void rq_inverse(const RqSeg& s, float y, float& x, float& loginv) {
    const float m = s.h / s.w;
    const float eta = (y - s.Y0) / s.h;
    const float th = rq_theta_from_eta(m, s.D0, s.D1, eta);
    const float om = 1.f - th;
    const float sig = s.D0 + s.D1 - 2.f * m;
    const float mr = m / (m + sig * th * om);
    const float Q = s.D1 * th * th + 2.f * m * th * om + s.D0 * om * om;
    x = s.X0 + s.w * th;
    loginv = -logf(mr * mr * Q);
}

// Vector-Jacobian product of (y, logg) = rq_forward(seg, x):
// given gy = dL/dy and gl = dL/dlogg, the adjoints of x and of the six segment
// parameters.  Reverse sweep through
//   a = th^2, b = th - a, c = 1 - 2 th + a
//   N = m a + D0 b ; den = m + sig b ; y = Y0 + h N / den
//   Q = D1 a + 2 m b + D0 c ; logg = 2 log m + log Q - 2 log den
// `th`, `om` = theta and 1 - theta, handed in so callers can form them without cancellation.
NFK_HD RqSegGrad rq_vjp_theta(const RqSeg& s, float th, float om, float gy, float gl) {
    const float rw = 1.f / s.w;
    const float m = s.h * rw;
    const float a = th * th;
    const float b = th * om;
    const float c = om * om;
    const float sig = s.D0 + s.D1 - 2.f * m;
    const float den = m + sig * b;
    const float N = m * a + s.D0 * b;
    const float Q = s.D1 * a + 2.f * m * b + s.D0 * c;
    const float rden = 1.f / den;

    float gh = gy * N * rden;
    const float gN = gy * s.h * rden;
    const float gden = -gy * s.h * N * rden * rden - 2.f * gl * rden;
    const float gQ = gl / Q;
    float gm = 2.f * gl / m + gN * a + gden + 2.f * gQ * b;
    const float gsig = gden * b;
    float ga = gN * m + gQ * s.D1;
    float gb = gN * s.D0 + gden * sig + 2.f * gQ * m;
    const float gc = gQ * s.D0;
    RqSegGrad r;
    r.gD0 = gN * b + gQ * c + gsig;
    r.gD1 = gQ * a + gsig;
    gm -= 2.f * gsig;
    ga += gc - gb;
    const float gth = -2.f * gc + gb + 2.f * th * ga;
    gh += gm * rw;
    r.gw = -gm * m * rw - gth * th * rw;
    r.gx = gth * rw;
    r.gX0 = -gth * rw;
    r.gY0 = gy;
    r.gh = gh;
    return r;
}
NFK_HD RqSegGrad rq_forward_vjp(const RqSeg& s, float x, float gy, float gl) {
    const float th = (x - s.X0) / s.w;          // (rq_vjp_theta forms its own 1 / w; the compiler merges the two)
    return rq_vjp_theta(s, th, 1.f - th, gy, gl);
}

// ---------------------------------------------------------------------------
// Coupling spline whose knots come from the conditioner's raw channels
// (couplings_.py:211-262).  `Ld` is a callable: Ld(c) -> raw channel c at this site.
//   channels [0, K-1)        -> bin widths  (softmax)
//   channels [K-1, 2K-2)     -> bin heights (softmax)
//   channels [2K-2, 3K-2)    -> knot derivatives (softplus)
// Ld also offers pick(base, n, j) = channel base + j for a run-time j in [0, n): a plain
// load for memory-backed channels, a select chain for channels held in registers.
struct RqsCfg {
    float xlim0, xw, ylim0, yw;
    int left, right;   // kExtrapNone | kExtrapLinear
};

// What was selected for this site: needed again by the backward pass.
template <int K>
struct RqsSite {
    float ex[K - 1], ey[K - 1];   // exp(raw - max): unnormalised softmax terms
    float sx, sy;                 // their sums
    int j;                        // segment index, or -1 / K-1 for the linear tails
    float cumx, cumy;             // unnormalised prefix sums below the segment
};

// Builds the softmax terms and finds the segment by x (by_y = false) or by y.
template <int K, class Ld>
NFK_HD void rqs_select(const Ld& ld, const RqsCfg& cfg, float v, bool by_y, RqsSite<K>& st) {
    float mx = -INFINITY, my = -INFINITY;
#pragma unroll
    for (int c = 0; c < K - 1; ++c) {
        st.ex[c] = ld(c);
        st.ey[c] = ld(K - 1 + c);
        mx = fmaxf(mx, st.ex[c]);
        my = fmaxf(my, st.ey[c]);
    }
    st.sx = 0.f;
    st.sy = 0.f;
#pragma unroll
    for (int c = 0; c < K - 1; ++c) {
        st.ex[c] = expf(st.ex[c] - mx);
        st.ey[c] = expf(st.ey[c] - my);
        st.sx += st.ex[c];
        st.sy += st.ey[c];
    }
    const float lo = by_y ? cfg.ylim0 : cfg.xlim0;
    const float wd = by_y ? cfg.yw : cfg.xw;
    st.cumx = 0.f;
    st.cumy = 0.f;
    if (cfg.left == kExtrapLinear && v <= lo) {
        st.j = -1;
        return;
    }
    if (cfg.right == kExtrapLinear && v > lo + wd) {
        st.j = K - 1;
        return;
    }
    // number of interior knots strictly below v (searchsorted right=False, then
    // clamp(.,1,n_seg)-1): knots are increasing, so the test is monotone in c.
    const float scale = wd / (by_y ? st.sy : st.sx);
    int j = 0;
    float run = 0.f;              // prefix sum up to and including term c: knot c+1
#pragma unroll
    for (int c = 0; c < K - 2; ++c) {
        run += by_y ? st.ey[c] : st.ex[c];
        if (lo + run * scale < v) j = c + 1;
    }
    st.j = j;
    float cx = 0.f, cy = 0.f;
#pragma unroll
    for (int c = 0; c < K - 2; ++c) {
        if (c < j) {
            cx += st.ex[c];
            cy += st.ey[c];
        }
    }
    st.cumx = cx;
    st.cumy = cy;
}

template <int K>
NFK_HD float pick(const float (&a)[K], int j) {
    float r = a[0];
#pragma unroll
    for (int c = 1; c < K; ++c) r = (c == j) ? a[c] : r;
    return r;
}

template <int K, class Ld>
NFK_HD RqSeg rqs_segment(const Ld& ld, const RqsCfg& cfg, const RqsSite<K>& st) {
    RqSeg s;
    const float rx = cfg.xw / st.sx, ry = cfg.yw / st.sy;
    s.X0 = cfg.xlim0 + st.cumx * rx;
    s.w = pick<K - 1>(st.ex, st.j) * rx;
    s.Y0 = cfg.ylim0 + st.cumy * ry;
    s.h = pick<K - 1>(st.ey, st.j) * ry;
    s.D0 = softplus_ln2(ld.pick(2 * K - 2, K, st.j));
    s.D1 = softplus_ln2(ld.pick(2 * K - 2, K, st.j + 1));
    return s;
}

// forward: y = spline(x), logg = log dy/dx
template <int K, class Ld>
NFK_HD void rqs_site_forward(const Ld& ld, const RqsCfg& cfg, float x, float& y, float& logg) {
    RqsSite<K> st;
    rqs_select<K>(ld, cfg, x, false, st);
    if (st.j < 0) {            // linear tail on the left (spline.py:466-470)
        const float D = softplus_ln2(ld(2 * K - 2));
        y = cfg.ylim0 + D * (x - cfg.xlim0);
        logg = logf(D);
        return;
    }
    if (st.j == K - 1) {       // linear tail on the right (spline.py:476-478)
        const float D = softplus_ln2(ld(3 * K - 3));
        y = (cfg.ylim0 + cfg.yw) + D * (x - (cfg.xlim0 + cfg.xw));
        logg = logf(D);
        return;
    }
    rq_forward(rqs_segment<K>(ld, cfg, st), x, y, logg);
}

// inverse: x = spline^{-1}(y), loginv = log dx/dy
template <int K, class Ld>
NFK_HD void rqs_site_inverse(const Ld& ld, const RqsCfg& cfg, float y, float& x, float& loginv) {
    RqsSite<K> st;
    rqs_select<K>(ld, cfg, y, true, st);
    if (st.j < 0) {
        const float D = softplus_ln2(ld(2 * K - 2));
        x = cfg.xlim0 + (y - cfg.ylim0) / D;
        loginv = -logf(D);
        return;
    }
    if (st.j == K - 1) {
        const float D = softplus_ln2(ld(3 * K - 3));
        x = (cfg.xlim0 + cfg.xw) + (y - (cfg.ylim0 + cfg.yw)) / D;
        loginv = -logf(D);
        return;
    }
    rq_inverse(rqs_segment<K>(ld, cfg, st), y, x, loginv);
}

// VJP of rqs_site_forward w.r.t. x and all 3K-2 raw channels.
// `St` is a callable St(c, value) storing the gradient of raw channel c.
template <int K, class Ld, class St>
NFK_HD float rqs_site_backward(const Ld& ld, const RqsCfg& cfg, float x, float gy, float gl,
                               const St& store) {
    RqsSite<K> st;
    rqs_select<K>(ld, cfg, x, false, st);
    if (st.j < 0 || st.j == K - 1) {
        const int ch = st.j < 0 ? 2 * K - 2 : 3 * K - 3;
        const float raw = ld(ch);
        const float D = softplus_ln2(raw);
        const float dx = st.j < 0 ? x - cfg.xlim0 : x - (cfg.xlim0 + cfg.xw);
        const float gD = gy * dx + gl / D;
#pragma unroll
        for (int c = 0; c < 3 * K - 2; ++c) store(c, c == ch ? gD * softplus_ln2_grad(raw) : 0.f);
        return gy * D;
    }
    const float raw0 = ld.pick(2 * K - 2, K, st.j), raw1 = ld.pick(2 * K - 2, K, st.j + 1);
    RqSeg s;
    const float isx = 1.f / st.sx, isy = 1.f / st.sy;              // one reciprocal per softmax normaliser
    const float rx = cfg.xw * isx, ry = cfg.yw * isy;
    const float exj = pick<K - 1>(st.ex, st.j), eyj = pick<K - 1>(st.ey, st.j);
    s.X0 = cfg.xlim0 + st.cumx * rx;
    s.w = exj * rx;
    s.Y0 = cfg.ylim0 + st.cumy * ry;
    s.h = eyj * ry;
    s.D0 = softplus_ln2(raw0);
    s.D1 = softplus_ln2(raw1);
    const RqSegGrad g = rq_forward_vjp(s, x, gy, gl);
    // chain through X0 = xlim0 + xw sum_{i<j} p_i, w = xw p_j, p = softmax(raw):
    //   d/draw_k = xw p_k ( gX0 ([k<j] - C_j) + gw ([k==j] - p_j) )
    const float Cx = st.cumx * isx, px = exj * isx;
    const float Cy = st.cumy * isy, py = eyj * isy;
#pragma unroll
    for (int k = 0; k < K - 1; ++k) {
        const float below = k < st.j ? 1.f : 0.f, here = k == st.j ? 1.f : 0.f;
        store(k, st.ex[k] * rx * (g.gX0 * (below - Cx) + g.gw * (here - px)));
        store(K - 1 + k, st.ey[k] * ry * (g.gY0 * (below - Cy) + g.gh * (here - py)));
    }
    const float gr0 = g.gD0 * softplus_ln2_grad(raw0), gr1 = g.gD1 * softplus_ln2_grad(raw1);
#pragma unroll
    for (int k = 0; k < K; ++k) store(2 * K - 2 + k, k == st.j ? gr0 : (k == st.j + 1 ? gr1 : 0.f));
    return g.gx;
}

// ---------------------------------------------------------------------------
// Shared 1-D spline with explicit knots (SplineNet_ / DistConvertor_).
struct Spline1dCfg {
    int K;
    int left, right;   // kExtrapNone | kExtrapLinear | kExtrapAnti | kExtrapPeriodic
    int logistic;      // wrap as expit -> spline -> logit
};

// segment index for v among knots k[0..K-1]: #interior knots below v.
NFK_HD int knots_segment(const float* k, int K, float v) {
    int lo = 1, hi = K - 1;   // count of c in [1, K-2] with k[c] < v, by bisection
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (k[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo - 1;
}

NFK_HD RqSeg knots_seg(const float* kx, const float* ky, const float* kd, int j) {
    RqSeg s;
    s.X0 = kx[j];
    s.w = kx[j + 1] - kx[j];
    s.Y0 = ky[j];
    s.h = ky[j + 1] - ky[j];
    s.D0 = kd[j];
    s.D1 = kd[j + 1];
    return s;
}

// Plain (non-logistic) forward: value and log-derivative, with extrapolation.
// `refl` reports whether the point was mirrored (anti-periodic side).
NFK_HD void spline1d_forward(const float* kx, const float* ky, const float* kd,
                             const Spline1dCfg& cfg, float x, float& y, float& logg) {
    const int K = cfg.K;
    const float Xa = kx[0], Xb = kx[K - 1], Ya = ky[0], Yb = ky[K - 1];
    int refl = 0;
    if (x <= Xa) {
        if (cfg.left == kExtrapLinear) { y = Ya + kd[0] * (x - Xa); logg = logf(kd[0]); return; }
        if (cfg.left == kExtrapAnti && x < Xa) { x = 2.f * Xa - x; refl = 1; }
        if (cfg.left == kExtrapPeriodic && x < Xa) { x = 2.f * Xa - x; refl = 3; }
    } else if (x > Xb) {
        if (cfg.right == kExtrapLinear) { y = Yb + kd[K - 1] * (x - Xb); logg = logf(kd[K - 1]); return; }
        if (cfg.right == kExtrapAnti) { x = 2.f * Xb - x; refl = 2; }
        if (cfg.right == kExtrapPeriodic) { x = 2.f * Xb - x; refl = 3; }
    }
    const int j = knots_segment(kx, K, x);
    rq_forward(knots_seg(kx, ky, kd, j), x, y, logg);
    if (refl == 1) y = 2.f * Ya - y;
    if (refl == 2) y = 2.f * Yb - y;
    // periodic: y(2 X_end - x) = y(x), the slope changes sign -- its logarithm does not exist (the
    // reference's torch.log of the negative derivative is NaN as well)
    if (refl == 3) logg = logf(-1.f);
}

NFK_HD void spline1d_inverse(const float* kx, const float* ky, const float* kd,
                             const Spline1dCfg& cfg, float y, float& x, float& loginv) {
    const int K = cfg.K;
    const float Xa = kx[0], Xb = kx[K - 1], Ya = ky[0], Yb = ky[K - 1];
    int refl = 0;
    if (y <= Ya) {
        if (cfg.left == kExtrapLinear) { x = Xa + (y - Ya) / kd[0]; loginv = -logf(kd[0]); return; }
        if (cfg.left == kExtrapAnti && y < Ya) { y = 2.f * Ya - y; refl = 1; }
    } else if (y > Yb) {
        if (cfg.right == kExtrapLinear) { x = Xb + (y - Yb) / kd[K - 1]; loginv = -logf(kd[K - 1]); return; }
        if (cfg.right == kExtrapAnti) { y = 2.f * Yb - y; refl = 2; }
    }
    const int j = knots_segment(ky, K, y);
    rq_inverse(knots_seg(kx, ky, kd, j), y, x, loginv);
    if (refl == 1) x = 2.f * Xa - x;
    if (refl == 2) x = 2.f * Xb - x;
}

// VJP of spline1d_forward.  `Acc(i, value)` accumulates into gkx (i in [0,K)),
// gky (i in [K,2K)) and gkd (i in [2K,3K)).  Returns gx.
template <class Acc>
NFK_HD float spline1d_backward(const float* kx, const float* ky, const float* kd,
                               const Spline1dCfg& cfg, float x, float gy, float gl, const Acc& acc) {
    const int K = cfg.K;
    const float Xa = kx[0], Xb = kx[K - 1];
    int refl = 0;
    if (x <= Xa) {
        if (cfg.left == kExtrapLinear) {
            const float D = kd[0];
            acc(0, -gy * D); acc(K, gy); acc(2 * K, gy * (x - Xa) + gl / D);
            return gy * D;
        }
        if (cfg.left == kExtrapAnti && x < Xa) { x = 2.f * Xa - x; refl = 1; }
        if (cfg.left == kExtrapPeriodic && x < Xa) { x = 2.f * Xa - x; refl = 3; }
    } else if (x > Xb) {
        if (cfg.right == kExtrapLinear) {
            const float D = kd[K - 1];
            acc(K - 1, -gy * D); acc(2 * K - 1, gy); acc(3 * K - 1, gy * (x - Xb) + gl / D);
            return gy * D;
        }
        if (cfg.right == kExtrapAnti) { x = 2.f * Xb - x; refl = 2; }
        if (cfg.right == kExtrapPeriodic) { x = 2.f * Xb - x; refl = 4; }
    }
    if (refl >= 3) {                      // periodic: y = f(2 X_end - x); no log-derivative term
        const int jp = knots_segment(kx, K, x);
        const RqSegGrad gp = rq_forward_vjp(knots_seg(kx, ky, kd, jp), x, gy, 0.f);
        acc(jp, gp.gX0 - gp.gw);
        acc(jp + 1, gp.gw);
        acc(K + jp, gp.gY0 - gp.gh);
        acc(K + jp + 1, gp.gh);
        acc(2 * K + jp, gp.gD0);
        acc(2 * K + jp + 1, gp.gD1);
        acc(refl == 3 ? 0 : K - 1, 2.f * gp.gx);
        return -gp.gx;
    }
    // mirrored: y = 2 Y_end - f(2 X_end - x)  ->  the inner value receives -gy
    const float gyi = refl ? -gy : gy;
    const int j = knots_segment(kx, K, x);
    const RqSegGrad g = rq_forward_vjp(knots_seg(kx, ky, kd, j), x, gyi, gl);
    acc(j, g.gX0 - g.gw);
    acc(j + 1, g.gw);
    acc(K + j, g.gY0 - g.gh);
    acc(K + j + 1, g.gh);
    acc(2 * K + j, g.gD0);
    acc(2 * K + j + 1, g.gD1);
    if (refl) {
        const int e = refl == 1 ? 0 : K - 1;
        acc(e, 2.f * g.gx);        // d(2 X_e - x)/dX_e
        acc(K + e, 2.f * gy);      // d(2 Y_e - f)/dY_e
        return -g.gx;
    }
    return g.gx;
}

// ---------------------------------------------------------------------------
// Logistic chain  y = logit(f(expit(x)))  in complement form.
// A point of (0,1) is carried as the pair (s, c) with c = 1 - s, both accurate to
// fp32 relative precision, so the tails do not lose digits.
struct Unit {
    float s, c;
};
NFK_HD Unit expit_pair(float x) {
    const float e = expf(-fabsf(x));
    const float big = 1.f / (1.f + e), small = e * big;
    Unit u;
    u.s = x >= 0.f ? big : small;
    u.c = x >= 0.f ? small : big;
    return u;
}
// log(s (1-s)) for s = expit(x):  -|x| - 2 log(1 + e^{-|x|})
NFK_HD float log_sc_from_x(float x) {
    const float a = fabsf(x);
    return -a - 2.f * log1pf(expf(-a));
}

// Knots of the chain are handed over twice: kx, ky measured from the lower end and
// cx = x_hi - kx, cy = y_hi - ky measured from the upper end (x_hi = y_hi = 1), both
// built from cumulative softmax sums on their own side, so differences such as
// (X1 - s) for s -> 1 carry no cancellation.  Without this a 1-ulp error of a knot near
// 1 is amplified by 1/(bin width) and by 1/(1 - s').
struct UnitKnots {
    const float *kx, *ky, *kd, *cx, *cy;
    int K;
};

// one located segment, offsets formed on the accurate side
struct UnitSeg {
    RqSeg g;         // X0, Y0 from the lower end; w, h from the accurate side
    float cY1;       // y_hi - Y1
    float cX1;       // x_hi - X1
    float th, om;    // theta, 1 - theta  (by x for the forward map)
    int j;
    bool upx, upy;   // which representation was used (for the adjoints)
};

// number of interior knots strictly below the point (s, c = 1 - s)
NFK_HD int unit_segment(const float* k, const float* ck, int K, Unit u) {
    int lo = 1, hi = K - 1;
    const bool up = u.s > 0.5f;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const bool below = up ? (ck[mid] > u.c) : (k[mid] < u.s);
        if (below) lo = mid + 1; else hi = mid;
    }
    return lo - 1;
}

NFK_HD void unit_widths(const UnitKnots& k, int j, UnitSeg& sg) {
    sg.j = j;
    sg.upx = k.kx[j] >= 0.5f;
    sg.upy = k.ky[j] >= 0.5f;
    sg.g.X0 = k.kx[j];
    sg.g.Y0 = k.ky[j];
    sg.g.w = sg.upx ? k.cx[j] - k.cx[j + 1] : k.kx[j + 1] - k.kx[j];
    sg.g.h = sg.upy ? k.cy[j] - k.cy[j + 1] : k.ky[j + 1] - k.ky[j];
    sg.g.D0 = k.kd[j];
    sg.g.D1 = k.kd[j + 1];
    sg.cX1 = k.cx[j + 1];
    sg.cY1 = k.cy[j + 1];
}

// f on (s, c): returns (s', c') and log f'(s).
NFK_HD UnitSeg spline_unit_forward(const UnitKnots& k, Unit u, Unit& v, float& logg) {
    UnitSeg sg;
    unit_widths(k, unit_segment(k.kx, k.cx, k.K, u), sg);
    const RqSeg& g = sg.g;
    const bool up = u.s > 0.5f;
    const float rw = 1.f / g.w;                                   // one reciprocal serves theta, omega and the slope
    sg.th = (up ? k.cx[sg.j] - u.c : u.s - g.X0) * rw;           // (s - X0)/w
    sg.om = (up ? u.c - sg.cX1 : k.kx[sg.j + 1] - u.s) * rw;     // (X1 - s)/w
    const float th = sg.th, om = sg.om;
    const float m = g.h * rw;
    const float sig = g.D0 + g.D1 - 2.f * m;
    const float rden = 1.f / (m + sig * th * om);
    v.s = g.Y0 + g.h * th * (m * th + g.D0 * om) * rden;
    v.c = sg.cY1 + g.h * om * (m * om + g.D1 * th) * rden;       // Y1 - g0: an exact identity
    const float Q = g.D1 * th * th + 2.f * m * th * om + g.D0 * om * om;
    const float mr = m * rden;
    logg = logf(mr * mr * Q);
    return sg;
}

// inverse of the above on (s', c').
NFK_HD void spline_unit_inverse(const UnitKnots& k, Unit v, Unit& u, float& loginv) {
    UnitSeg sg;
    unit_widths(k, unit_segment(k.ky, k.cy, k.K, v), sg);
    const RqSeg& g = sg.g;
    const bool up = v.s > 0.5f;
    const float m = g.h / g.w;
    const float rh = 1.f / g.h;
    const float eta = (up ? k.cy[sg.j] - v.c : v.s - g.Y0) * rh;
    const float ome = (up ? v.c - sg.cY1 : k.ky[sg.j + 1] - v.s) * rh;      // 1 - eta
    float th, om;
    if (eta <= 0.5f) {
        th = rq_theta_from_eta(m, g.D0, g.D1, eta);
        om = 1.f - th;
    } else {                                                 // mirrored segment: D0 <-> D1
        om = rq_theta_from_eta(m, g.D1, g.D0, ome);
        th = 1.f - om;
    }
    const float sig = g.D0 + g.D1 - 2.f * m;
    const float den = m + sig * th * om;
    const float Q = g.D1 * th * th + 2.f * m * th * om + g.D0 * om * om;
    u.s = g.X0 + g.w * th;
    u.c = sg.cX1 + g.w * om;
    const float mr = m / den;
    loginv = -logf(mr * mr * Q);
}

// DistConvertor_ forward (inverse = false) or ModuleList_.backward (inverse = true):
//   y = logit(F(expit(x))),  logj = log(s c) + log F'(s) - log(s' c')
// anti != 0: odd extension about 0 (xlim0 = ylim0 = 0.5, extrap left 'anti').
NFK_HD void distconv_eval(const UnitKnots& k, bool anti, bool inverse, float x, float& y, float& logj) {
    const float xa = anti ? fabsf(x) : x;
    // expit_pair and log_sc_from_x share e^{-|x|}
    const float a = fabsf(xa), e = expf(-a);
    const float big = 1.f / (1.f + e), small = e * big;
    Unit u;
    u.s = xa >= 0.f ? big : small;
    u.c = xa >= 0.f ? small : big;
    Unit v;
    float lg;
    if (!inverse) spline_unit_forward(k, u, v, lg);
    else spline_unit_inverse(k, u, v, lg);
    const float ls = logf(v.s), lc = logf(v.c);
    y = anti ? copysignf(ls - lc, x) : ls - lc;
    logj = (-a - 2.f * log1pf(e)) + lg - (ls + lc);
}

// VJP of distconv_eval (forward direction).  `acc(i, v)` accumulates into a [5K]
// buffer laid out gkx | gky | gkd | gcx | gcy; the knot arrays of the two ends are
// independent inputs here (the host chains both back to the weights).
template <class Acc>
NFK_HD float distconv_backward(const UnitKnots& k, bool anti, float x, float gy, float gl, const Acc& acc) {
    const int K = k.K;
    const float xa = anti ? fabsf(x) : x;
    const float sgn = (anti && x < 0.f) ? -1.f : 1.f;
    const Unit u = expit_pair(xa);
    Unit v;
    float lg;
    const UnitSeg sg = spline_unit_forward(k, u, v, lg);
    // y = sgn (log s' - log c'),  logj = log(s c) + lg - log s' - log c',  c' = 1 - s'
    const float gya = sgn * gy;
    const float gsp = gya * (1.f / v.s + 1.f / v.c) - gl * (1.f / v.s - 1.f / v.c);
    const RqSegGrad g = rq_vjp_theta(sg.g, sg.th, sg.om, gsp, gl);
    const int j = sg.j;
    if (sg.upx) {            // X0 = 1 - cx[j], w = cx[j] - cx[j+1]
        acc(3 * K + j, g.gw - g.gX0);
        acc(3 * K + j + 1, -g.gw);
    } else {                 // X0 = kx[j], w = kx[j+1] - kx[j]
        acc(j, g.gX0 - g.gw);
        acc(j + 1, g.gw);
    }
    if (sg.upy) {
        acc(4 * K + j, g.gh - g.gY0);
        acc(4 * K + j + 1, -g.gh);
    } else {
        acc(K + j, g.gY0 - g.gh);
        acc(K + j + 1, g.gh);
    }
    acc(2 * K + j, g.gD0);
    acc(2 * K + j + 1, g.gD1);
    const float gs = g.gx + gl * (1.f / u.s - 1.f / u.c);
    return sgn * gs * u.s * u.c;
}

// ---------------------------------------------------------------------------
// Expit_ / Logit_ alone (modules_.py:93-114)
NFK_HD void expit_eval(float x, float& y, float& logj) {
    y = expit_pair(x).s;
    logj = log_sc_from_x(x);                 // -x + 2 log y
}
NFK_HD void logit_eval(float x, float& y, float& logj) {
    const float lx = logf(x), lc = log1pf(-x);
    y = lx - lc;
    logj = -(lx + lc);
}

// ---------------------------------------------------------------------------
// Philox4x32-10 counter-based generator (Salmon et al. 2011) and Box-Muller.
struct Philox {
    uint32_t c[4];
};
NFK_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    const uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
}
NFK_HD Philox philox4x32_10(uint64_t counter_lo, uint64_t counter_hi, uint64_t key) {
    uint32_t c0 = (uint32_t)counter_lo, c1 = (uint32_t)(counter_lo >> 32);
    uint32_t c2 = (uint32_t)counter_hi, c3 = (uint32_t)(counter_hi >> 32);
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0, l0, h1, l1;
        mulhilo(0xD2511F53u, c0, h0, l0);
        mulhilo(0xCD9E8D57u, c2, h1, l1);
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox p;
    p.c[0] = c0; p.c[1] = c1; p.c[2] = c2; p.c[3] = c3;
    return p;
}
// two standard normals from two 32-bit words
NFK_HD void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
    const float u1 = (float)a * 2.3283064365386963e-10f + 1.1641532182693481e-10f;  // (a + 0.5) 2^-32
    const float u2 = (float)b * 2.3283064365386963e-10f + 1.1641532182693481e-10f;
#if defined(__CUDA_ARCH__)
    // MUFU-based log / sin / cos: the draw is a sampler, not a parity path (the density that
    // goes with it is computed from the very z it returns), and the generator is HBM-bound only
    // with the fast forms.  The angle is kept in (-pi, pi] where __sincosf is accurate to 2^-21.
    const float r = sqrtf(-2.f * __logf(u1));
    const float ang = 6.283185307179586f * (u2 - 0.5f);
    float sn, cs;
    __sincosf(ang, &sn, &cs);
#else
    const float r = sqrtf(-2.f * logf(u1));
    const float ang = 6.283185307179586f * (u2 - 0.5f);
    const float sn = sinf(ang), cs = cosf(ang);
#endif
    z0 = r * cs;
    z1 = r * sn;
}

// ---------------------------------------------------------------------------
// Lattice geometry helpers (row-major sites, periodic).
struct Lat {
    int ndim;
    int shape[4];
    int stride[4];
};
NFK_HD Lat make_lat(int ndim, const int* shape) {
    Lat l;
    l.ndim = ndim;
    int st = 1;
    for (int d = 3; d >= 0; --d) {
        if (d < ndim) {
            l.shape[d] = shape[d];
            l.stride[d] = st;
            st *= shape[d];
        } else {
            l.shape[d] = 1;
            l.stride[d] = 0;
        }
    }
    return l;
}
NFK_HD void site_coords(const Lat& l, int s, int* c) {
    for (int d = 0; d < 4; ++d) c[d] = 0;
    for (int d = l.ndim - 1; d >= 0; --d) {
        c[d] = s % l.shape[d];
        s /= l.shape[d];
    }
}
// site reached from coords c by moving `delta` along axis d (periodic)
NFK_HD int shifted_site(const Lat& l, int s, const int* c, int d, int delta) {
    int n = c[d] + delta;
    const int L = l.shape[d];
    n %= L;
    if (n < 0) n += L;
    return s + (n - c[d]) * l.stride[d];
}
// (1 - parity + sum(ind) [- ind[exclude]]) mod 2, python semantics (mask.py:55-58)
NFK_HD uint8_t evenodd_bit(const Lat& l, int s, int parity, int exclude_mu) {
    int c[4];
    site_coords(l, s, c);
    int sum = 1 - parity;
    for (int d = 0; d < l.ndim; ++d)
        if (d != exclude_mu) sum += c[d];
    return (uint8_t)(((sum % 2) + 2) % 2);
}
NFK_HD uint8_t alongaxis_bit(const Lat& l, int s, int parity, int mu) {
    int c[4];
    site_coords(l, s, c);
    const int v = 1 - parity + c[mu];
    return (uint8_t)(((v % 2) + 2) % 2);
}

// activations of ConvAct (modules.py:43-54) and their derivatives expressed
// through the post-activation value where that is possible (tanh, relu, ...).
NFK_HD float act_apply(int kind, float v) {
    switch (kind) {
        case 1: return tanhf(v);
        case 2: return v > 0.f ? v : 0.f;
        case 3: return v > 0.f ? v : 0.01f * v;
        case 4: return v > 20.f ? v : log1pf(expf(v));
        case 5: return fabsf(v);
        default: return v;
    }
}

}  // namespace nfk
