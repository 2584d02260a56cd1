// nfk_ops.cuh -- per-(sample, site) operations of the hot path.
//
// Each functor does the loads, arithmetic and stores for ONE site (or a short run
// of consecutive sites) of ONE sample and returns that site's contribution to the
// per-sample sum (log|det J|, log-prob, action).  The CUDA kernels in
// nfk_kernels.cu wrap them in a one-CTA-per-(sample, chunk) loop with a
// warp-shuffle / shared-memory reduction; tests/cpu_harness wraps them in a plain
// host loop so indexing + arithmetic are checked against the oracle without a GPU.
#pragma once

#include "nfk_math.cuh"

namespace nfk {

#if defined(__CUDA_ARCH__)
#define NFK_LDG(p) __ldg(p)
#else
#define NFK_LDG(p) (*(p))
#endif

// is site s updated by a coupling step of parity p?  (couplings_.py:56-64, mask.py:36-37)
NFK_HD bool site_active(const uint8_t* mask, int64_t s, int active_val) {
    return NFK_LDG(mask + s) == (uint8_t)active_val;
}

// ------------------------------------------------------------------ mask.split
struct MaskSelectOp {
    const float* x;
    const uint8_t* mask;
    int keep;
    float* y;
    int64_t V;
    NFK_HD float operator()(int64_t b, int64_t s) const {
        const int64_t i = b * V + s;
        y[i] = site_active(mask, s, keep) ? NFK_LDG(x + i) : 0.f;
        return 0.f;
    }
};

// ------------------------------------------------------------------ prior
// Prior.log_prob density at one site (prior.py:30-36)
struct PriorLogProbOp {
    const float* x;
    const float* loc;     // [V] or null
    const float* scale;   // [V] or null
    int64_t V;
    NFK_HD float operator()(int64_t b, int64_t s) const {
        const float mu = loc ? NFK_LDG(loc + s) : 0.f;
        const float sg = scale ? NFK_LDG(scale + s) : 1.f;
        const float z = (NFK_LDG(x + b * V + s) - mu) / sg;
        return -0.5f * z * z - (scale ? logf(sg) : 0.f) - kLogSqrt2Pi;
    }
};

// four consecutive sites (s % 4 == 0) with one 128-bit load when there is no loc / scale
struct PriorLogProbOp4 {
    const float* x;
    const float* loc;
    const float* scale;
    int64_t V;
    NFK_HD float operator()(int64_t b, int64_t s, int n) const {
#if defined(__CUDA_ARCH__)
        if (n == 4 && !loc && !scale && (((b * V + s) & 3) == 0)) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(x + b * V + s));
            return -0.5f * (v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w) - 4.f * kLogSqrt2Pi;
        }
#endif
        const PriorLogProbOp one{x, loc, scale, V};
        float acc = 0.f;
        for (int i = 0; i < n; ++i) acc += one(b, s + i);
        return acc;
    }
};

// NormalPrior.sample_: four consecutive sites per Philox call.
struct PriorSampleOp {
    float* x;
    const float* loc;
    const float* scale;
    int64_t V;
    uint64_t seed, offset;
    const uint64_t* state;    // device {seed, offset} (overrides the two values above) or null
    // sites [s, s+n), n <= 4, s % 4 == 0
    NFK_HD float operator()(int64_t b, int64_t s, int n) const {
        const uint64_t quads = (uint64_t)((V + 3) / 4);
        const uint64_t sd = state ? NFK_LDG(state) : seed, of = state ? NFK_LDG(state + 1) : offset;
        const Philox r = philox4x32_10((uint64_t)b * quads + (uint64_t)(s >> 2), of, sd);
        float z[4];
        box_muller(r.c[0], r.c[1], z[0], z[1]);
        box_muller(r.c[2], r.c[3], z[2], z[3]);
        float acc = 0.f;
#if defined(__CUDA_ARCH__)
        if (n == 4 && !loc && !scale && (((b * V + s) & 3) == 0)) {       // the common case: one 128-bit store
            *reinterpret_cast<float4*>(x + b * V + s) = make_float4(z[0], z[1], z[2], z[3]);
            return -0.5f * (z[0] * z[0] + z[1] * z[1] + z[2] * z[2] + z[3] * z[3]) - 4.f * kLogSqrt2Pi;
        }
#endif
        for (int i = 0; i < n; ++i) {
            const float mu = loc ? NFK_LDG(loc + s + i) : 0.f;
            const float sg = scale ? NFK_LDG(scale + s + i) : 1.f;
            x[b * V + s + i] = mu + sg * z[i];
            acc += -0.5f * z[i] * z[i] - (scale ? logf(sg) : 0.f) - kLogSqrt2Pi;
        }
        return acc;
    }
};

// ------------------------------------------------------------------ affine
// MODE 0: forward, 1: inverse  (couplings_.py:123-139)
template <int MODE>
struct AffineOp {
    const float* x;
    const float* out;     // [B][2][V]
    const uint8_t* mask;
    int active_val;
    int frozen_copy;
    float* y;
    int64_t V;
    NFK_HD float operator()(int64_t b, int64_t s) const {
        const int64_t i = b * V + s;
        const float xv = NFK_LDG(x + i);
        if (!site_active(mask, s, active_val)) {
            y[i] = frozen_copy ? xv : 0.f;
            return 0.f;
        }
        const float t = NFK_LDG(out + (2 * b) * V + s);
        const float sc = fabsf(NFK_LDG(out + (2 * b + 1) * V + s));
        if (MODE == 0) {
            y[i] = t + xv * expf(-sc);
            return -sc;
        }
        y[i] = (xv - t) * expf(sc);
        return sc;
    }
};

// four consecutive sites (s % 4 == 0, V % 4 == 0): 128-bit loads / stores of x, t, s, y
template <int MODE>
struct AffineOp4 {
    const float* x;
    const float* out;     // [B][2][V]
    const uint8_t* mask;
    int active_val;
    int frozen_copy;
    float* y;
    int64_t V;
    NFK_HD float operator()(int64_t b, int64_t s, int n) const {
#if defined(__CUDA_ARCH__)
        if (n == 4) {
            const int64_t i = b * V + s;
            const float4 xv = __ldg(reinterpret_cast<const float4*>(x + i));
            const float4 tv = __ldg(reinterpret_cast<const float4*>(out + (2 * b) * V + s));
            const float4 sv = __ldg(reinterpret_cast<const float4*>(out + (2 * b + 1) * V + s));
            const uint32_t m = __ldg(reinterpret_cast<const uint32_t*>(mask + s));
            const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, ta[4] = {tv.x, tv.y, tv.z, tv.w};
            const float sa[4] = {sv.x, sv.y, sv.z, sv.w};
            float ya[4], acc = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool active = (int)((m >> (8 * k)) & 0xFF) == active_val;
                const float sc = fabsf(sa[k]);
                const float tr = MODE == 0 ? ta[k] + xa[k] * expf(-sc) : (xa[k] - ta[k]) * expf(sc);
                ya[k] = active ? tr : (frozen_copy ? xa[k] : 0.f);
                acc += active ? (MODE == 0 ? -sc : sc) : 0.f;
            }
            *reinterpret_cast<float4*>(y + i) = make_float4(ya[0], ya[1], ya[2], ya[3]);
            return acc;
        }
#endif
        const AffineOp<MODE> one{x, out, mask, active_val, frozen_copy, y, V};
        float acc = 0.f;
        for (int k = 0; k < n; ++k) acc += one(b, s + k);
        return acc;
    }
};

struct AffineBwdOp {
    const float* x;
    const float* out;
    const uint8_t* mask;
    int active_val;
    int frozen_copy;
    const float* gy;
    const float* glog;    // [B] or null
    float* gx;
    float* gout;
    int64_t V;
    NFK_HD float operator()(int64_t b, int64_t s) const {
        const int64_t i = b * V + s;
        const float g = NFK_LDG(gy + i);
        if (!site_active(mask, s, active_val)) {
            gx[i] = frozen_copy ? g : 0.f;
            gout[(2 * b) * V + s] = 0.f;
            gout[(2 * b + 1) * V + s] = 0.f;
            return 0.f;
        }
        const float sraw = NFK_LDG(out + (2 * b + 1) * V + s);
        const float e = expf(-fabsf(sraw));
        const float gl = glog ? NFK_LDG(glog + b) : 0.f;
        // y = t + x e^{-|s|}; log = -|s|
        gx[i] = g * e;
        gout[(2 * b) * V + s] = g;
        const float sgn = sraw > 0.f ? 1.f : (sraw < 0.f ? -1.f : 0.f);
        gout[(2 * b + 1) * V + s] = sgn * (-g * NFK_LDG(x + i) * e - gl);
        return 0.f;
    }
};

// ------------------------------------------------------------------ shift
struct ShiftOp {
    const float* x;
    const float* out;     // [B][1][V]
    const uint8_t* mask;
    int active_val;
    int frozen_copy;
    float sign;
    float* y;
    int64_t V;
    NFK_HD float operator()(int64_t b, int64_t s) const {
        const int64_t i = b * V + s;
        const float xv = NFK_LDG(x + i);
        if (site_active(mask, s, active_val)) y[i] = xv + sign * NFK_LDG(out + i);
        else y[i] = frozen_copy ? xv : 0.f;
        return 0.f;
    }
};

// ------------------------------------------------------------------ RQ spline coupling
struct ChanLoad {
    const float* base;    // &out[b][0][s]
    int64_t stride;       // V
    NFK_HD float operator()(int c) const { return NFK_LDG(base + c * stride); }
    NFK_HD float pick(int c0, int, int j) const { return NFK_LDG(base + (c0 + j) * stride); }
};
struct ChanStore {
    float* base;
    int64_t stride;
    NFK_HD void operator()(int c, float v) const { base[c * stride] = v; }
};

// MODE 0: forward, 1: inverse  (couplings_.py:178-200)
template <int K, int MODE>
struct RqsOp {
    const float* x;
    const float* out;     // [B][3K-2][V]
    const uint8_t* mask;
    int active_val;
    int frozen_copy;
    RqsCfg cfg;
    float* y;
    int64_t V;
    NFK_HD float operator()(int64_t b, int64_t s) const {
        const int64_t i = b * V + s;
        const float xv = NFK_LDG(x + i);
        if (!site_active(mask, s, active_val)) {
            y[i] = frozen_copy ? xv : 0.f;
            return 0.f;
        }
        const ChanLoad ld{out + (int64_t)(3 * K - 2) * b * V + s, V};
        float yv, lg;
        if (MODE == 0) rqs_site_forward<K>(ld, cfg, xv, yv, lg);
        else rqs_site_inverse<K>(ld, cfg, xv, yv, lg);
        y[i] = yv;
        return lg;
    }
};

template <int K>
struct RqsBwdOp {
    const float* x;
    const float* out;
    const uint8_t* mask;
    int active_val;
    int frozen_copy;
    RqsCfg cfg;
    const float* gy;
    const float* glog;
    float* gx;
    float* gout;
    int64_t V;
    NFK_HD float operator()(int64_t b, int64_t s) const {
        const int64_t i = b * V + s;
        const float g = NFK_LDG(gy + i);
        const int64_t o = (int64_t)(3 * K - 2) * b * V + s;
        const ChanStore st{gout + o, V};
        if (!site_active(mask, s, active_val)) {
            gx[i] = frozen_copy ? g : 0.f;
#pragma unroll
            for (int c = 0; c < 3 * K - 2; ++c) st(c, 0.f);
            return 0.f;
        }
        const ChanLoad ld{out + o, V};
        const float gl = glog ? NFK_LDG(glog + b) : 0.f;
        gx[i] = rqs_site_backward<K>(ld, cfg, NFK_LDG(x + i), g, gl, st);
        return 0.f;
    }
};

// ------------------------------------------------------------------ Expit_ / Logit_
struct LogisticOp {
    const float* x;
    int which;            // 0 expit, 1 logit
    float* y;
    int64_t V;
    NFK_HD float operator()(int64_t b, int64_t s) const {
        const int64_t i = b * V + s;
        float yv, lj;
        if (which == 0) expit_eval(NFK_LDG(x + i), yv, lj);
        else logit_eval(NFK_LDG(x + i), yv, lj);
        y[i] = yv;
        return lj;
    }
};
struct LogisticBwdOp {
    const float* x;
    int which;
    const float* gy;
    const float* glog;
    float* gx;
    int64_t V;
    NFK_HD float operator()(int64_t b, int64_t s) const {
        const int64_t i = b * V + s;
        const float xv = NFK_LDG(x + i), g = NFK_LDG(gy + i);
        const float gl = glog ? NFK_LDG(glog + b) : 0.f;
        if (which == 0) {          // y = s, logj = log(s c): dy/dx = s c, dlogj/dx = c - s
            const Unit u = expit_pair(xv);
            gx[i] = g * u.s * u.c + gl * (u.c - u.s);
        } else {                   // y = log x - log(1-x), logj = -log x - log(1-x)
            const float c = 1.f - xv;
            gx[i] = g * (1.f / xv + 1.f / c) - gl * (1.f / xv - 1.f / c);
        }
        return 0.f;
    }
};

// ------------------------------------------------------------------ shared 1-D spline
// knots: [5][K] = kx | ky | kd | cx | cy  (cx, cy only read by the logistic chain)
struct Spline1dOp {
    const float* x;
    const float* knots;          // shared memory on the device
    Spline1dCfg cfg;
    int inverse;
    float* y;
    int64_t V;
    NFK_HD float operator()(int64_t b, int64_t s) const {
        const int64_t i = b * V + s;
        const float xv = NFK_LDG(x + i);
        const int K = cfg.K;
        const float *kx = knots, *ky = knots + K, *kd = knots + 2 * K;
        float yv, lj;
        if (cfg.logistic) {
            const UnitKnots uk{kx, ky, kd, knots + 3 * K, knots + 4 * K, K};
            distconv_eval(uk, cfg.left == kExtrapAnti, inverse != 0, xv, yv, lj);
        } else if (inverse) {
            spline1d_inverse(kx, ky, kd, cfg, xv, yv, lj);
        } else {
            spline1d_forward(kx, ky, kd, cfg, xv, yv, lj);
        }
        y[i] = yv;
        return lj;
    }
};

// accumulate into a [5K] buffer: atomics on the device, plain adds on the host
struct KnotAcc {
    float* buf;
    NFK_HD void operator()(int i, float v) const {
#if defined(__CUDA_ARCH__)
        atomicAdd(buf + i, v);
#else
        buf[i] += v;
#endif
    }
};

struct Spline1dBwdOp {
    const float* x;
    const float* knots;
    Spline1dCfg cfg;
    const float* gy;
    const float* glog;
    float* gx;
    float* gknots;        // [5K] accumulation buffer (shared memory on the device)
    int64_t V;
    NFK_HD float operator()(int64_t b, int64_t s) const {
        const int64_t i = b * V + s;
        const float xv = NFK_LDG(x + i), g = NFK_LDG(gy + i);
        const float gl = glog ? NFK_LDG(glog + b) : 0.f;
        const KnotAcc acc{gknots};
        const int K = cfg.K;
        const float *kx = knots, *ky = knots + K, *kd = knots + 2 * K;
        if (cfg.logistic) {
            const UnitKnots uk{kx, ky, kd, knots + 3 * K, knots + 4 * K, K};
            gx[i] = distconv_backward(uk, cfg.left == kExtrapAnti, xv, g, gl, acc);
        } else {
            gx[i] = spline1d_backward(kx, ky, kd, cfg, xv, g, gl, acc);
        }
        return 0.f;
    }
};

// ------------------------------------------------------------------ phi^4 action
// S density at one site with backward neighbours only (scalar_action.py:40-46)
struct Phi4Op {
    const float* phi;
    Lat lat;
    float w0, w2, w4;
    int64_t V;
    NFK_HD float operator()(int64_t b, int64_t s) const {
        const float* p = phi + b * V;
        const float v = NFK_LDG(p + s);
        int c[4];
        site_coords(lat, (int)s, c);
        float nb = 0.f;
        for (int d = 0; d < lat.ndim; ++d) nb += NFK_LDG(p + shifted_site(lat, (int)s, c, d, -1));
        const float v2 = v * v;
        return v2 * (w2 + w4 * v2) - w0 * v * nb;
    }
};
struct Phi4BwdOp {
    const float* phi;
    Lat lat;
    float w0, w2, w4;
    const float* gS;
    float* gphi;
    int64_t V;
    NFK_HD float operator()(int64_t b, int64_t s) const {
        const float* p = phi + b * V;
        const float v = NFK_LDG(p + s);
        int c[4];
        site_coords(lat, (int)s, c);
        float nb = 0.f;
        for (int d = 0; d < lat.ndim; ++d)
            nb += NFK_LDG(p + shifted_site(lat, (int)s, c, d, -1)) + NFK_LDG(p + shifted_site(lat, (int)s, c, d, +1));
        gphi[b * V + s] = NFK_LDG(gS + b) * (v * (2.f * w2 + 4.f * w4 * v * v) - w0 * nb);
        return 0.f;
    }
};

// ------------------------------------------------------------------ circular convolution
// derivative of an activation expressed through its OUTPUT h (post-activation)
NFK_HD float act_grad_from_post(int kind, float h) {
    switch (kind) {
        case 1: return 1.f - h * h;                 // tanh
        case 2: return h > 0.f ? 1.f : 0.f;         // relu
        case 3: return h > 0.f ? 1.f : 0.01f;       // leaky_relu (sign preserved)
        case 4: return 1.f - expf(-h);              // softplus: sigmoid(v) = 1 - e^{-h}
        default: return 1.f;
    }
}

// tap t of a ksize^ndim kernel -> periodic neighbour of site s (coords c)
NFK_HD int tap_neighbor(const Lat& lat, int s, const int* c, int t, int ksize) {
    int n = s;
    const int half = ksize / 2;
    for (int d = lat.ndim - 1; d >= 0; --d) {
        const int td = t % ksize;
        t /= ksize;
        int nc = c[d] + td - half;
        const int L = lat.shape[d];
        nc %= L;
        if (nc < 0) nc += L;
        n += (nc - c[d]) * lat.stride[d];
    }
    return n;
}

// acc[co] += sum_{ci,t} w[(ci*T + t)*CO + co] * in[b][ci][nbr(s,t)]   for one site.
// The taps are walked as four nested loops (unused dimensions have one tap), each level keeping
// its periodic coordinate incrementally: one modulo per level and site instead of ndim per tap,
// which is what made the 27- and 81-tap (3-D, 4-D) convolutions index-bound.  Tap order = the
// weight layout: last dimension fastest.
template <int CO>
NFK_HD void conv_site(const float* in_b, const float* wt, const uint8_t* in_mask, int in_keep,
                      const Lat& lat, int s, int Ci, int T, int ksize, int64_t V, float* acc) {
    int c[4];
    site_coords(lat, s, c);
    const int half = ksize / 2;
    int k[4], start[4];
    for (int d = 0; d < 4; ++d) {
        k[d] = d < lat.ndim ? ksize : 1;
        int v = d < lat.ndim ? (c[d] - half) % lat.shape[d] : 0;
        start[d] = v < 0 ? v + lat.shape[d] : v;
    }
    int t = 0;
    int x0 = start[0];
    for (int t0 = 0; t0 < k[0]; ++t0, x0 = x0 + 1 == lat.shape[0] ? 0 : x0 + 1) {
        const int n0 = x0 * lat.stride[0];
        int x1 = start[1];
        for (int t1 = 0; t1 < k[1]; ++t1, x1 = x1 + 1 == lat.shape[1] ? 0 : x1 + 1) {
            const int n1 = n0 + x1 * lat.stride[1];
            int x2 = start[2];
            for (int t2 = 0; t2 < k[2]; ++t2, x2 = x2 + 1 == lat.shape[2] ? 0 : x2 + 1) {
                const int n2 = n1 + x2 * lat.stride[2];
                int x3 = start[3];
                for (int t3 = 0; t3 < k[3]; ++t3, ++t, x3 = x3 + 1 == lat.shape[3] ? 0 : x3 + 1) {
                    const int n = n2 + x3 * lat.stride[3];
                    if (in_mask && NFK_LDG(in_mask + n) != (uint8_t)in_keep) continue;
                    for (int ci = 0; ci < Ci; ++ci) {
                        const float v = NFK_LDG(in_b + ci * V + n);
                        const float* wr = wt + (ci * T + t) * CO;
#pragma unroll
                        for (int co = 0; co < CO; ++co) acc[co] = fmaf(v, wr[co], acc[co]);
                    }
                }
            }
        }
    }
}

}  // namespace nfk
