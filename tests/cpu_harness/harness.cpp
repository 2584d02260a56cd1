// tests/cpu_harness/harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the very same per-site functors the CUDA kernels use
// (normflow__b200/csrc/nfk_ops.cuh, nfk_math.cuh) with g++ and drives them with a
// plain host loop, so the fp32 arithmetic and the indexing can be checked against
// the oracle on a machine without a GPU.  It mirrors the C ABI with a `cpu_`
// prefix.  The product never loads this library.

#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/normflow_b200.h"
#include "../../normflow__b200/csrc/nfk_ops.cuh"
#include "../../normflow__b200/csrc/nfk_fused.cuh"
#include "../../normflow__b200/csrc/nfk_knots.cuh"
#include "../../normflow__b200/csrc/nfk_psd.cuh"

using namespace nfk;

template <class Op>
static void run_sites(const Op& op, int64_t B, int64_t V, const float* log_in, float* log_out) {
    for (int64_t b = 0; b < B; ++b) {
        float acc = 0.f;
        for (int64_t s = 0; s < V; ++s) acc += op(b, s);
        if (log_out) log_out[b] = (log_in ? log_in[b] : 0.f) + acc;
    }
}

static RqsCfg to_cfg(const nfk_rqs_params& p) {
    RqsCfg c;
    c.xlim0 = p.xlim0; c.xw = p.xlim1 - p.xlim0;
    c.ylim0 = p.ylim0; c.yw = p.ylim1 - p.ylim0;
    c.left = p.extrap_left; c.right = p.extrap_right;
    return c;
}

#define FOR_EACH_K(X) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(14) X(16) X(20) X(24) X(32)

extern "C" {

int cpu_psd_weights_fwd(const float* ipsd, int64_t Kc, int Lh, int inverse, float* w, float* logj) {
    double acc = 0.0;
    for (int64_t k = 0; k < Kc; ++k) {
        w[k] = psd_weight(ipsd[k], inverse);
        acc += psd_logj_term(k, Lh, ipsd[k]);
    }
    logj[0] = (float)((inverse ? 0.5 : -0.5) * acc);
    return 0;
}
int cpu_psd_weights_bwd(const float* ipsd, const float* w, const float* gw, float glogj, int64_t Kc, int Lh,
                        int inverse, float* g_ipsd) {
    for (int64_t k = 0; k < Kc; ++k) g_ipsd[k] = psd_weight_grad(k, Lh, ipsd[k], w[k], gw[k], glogj, inverse);
    return 0;
}

int cpu_knots_fwd(const float* wx, const float* wy, const float* wd, int K, float xlo, float xw, float ylo,
                  float yw, float* table) {
    KnotScratch scratch;
    knots_fwd_body(KnotArgs{wx, wy, wd, K, xlo, xw, ylo, yw}, table, scratch, 0, 1, KnotNoSync{});
    return 0;
}
int cpu_knots_bwd(const float* wx, const float* wy, const float* wd, int K, float xlo, float xw, float ylo,
                  float yw, const float* g, float* gwx, float* gwy, float* gwd) {
    KnotScratch scratch;
    knots_bwd_body(KnotArgs{wx, wy, wd, K, xlo, xw, ylo, yw}, g, gwx, gwy, gwd, scratch, 0, 1, KnotNoSync{});
    return 0;
}

int cpu_mask_evenodd(uint8_t* mask, nfk_lattice lat, int parity, int exclude_mu) {
    const Lat l = make_lat(lat.ndim, lat.shape);
    int V = 1;
    for (int d = 0; d < lat.ndim; ++d) V *= lat.shape[d];
    for (int s = 0; s < V; ++s) mask[s] = evenodd_bit(l, s, parity, exclude_mu < 0 ? -1 : exclude_mu);
    return 0;
}
int cpu_mask_alongaxis(uint8_t* mask, nfk_lattice lat, int parity, int mu) {
    const Lat l = make_lat(lat.ndim, lat.shape);
    int V = 1;
    for (int d = 0; d < lat.ndim; ++d) V *= lat.shape[d];
    for (int s = 0; s < V; ++s) mask[s] = alongaxis_bit(l, s, parity, mu);
    return 0;
}
int cpu_mask_select(const float* x, const uint8_t* mask, int keep, float* y, int64_t B, int64_t V) {
    run_sites(MaskSelectOp{x, mask, keep, y, V}, B, V, nullptr, nullptr);
    return 0;
}
int cpu_prior_normal_sample(float* x, float* logr, int64_t B, int64_t V, const float* loc,
                            const float* scale, uint64_t seed, uint64_t offset) {
    const PriorSampleOp op{x, loc, scale, V, seed, offset};
    for (int64_t b = 0; b < B; ++b) {
        float acc = 0.f;
        for (int64_t s = 0; s < V; s += 4) acc += op(b, s, V - s < 4 ? (int)(V - s) : 4);
        if (logr) logr[b] = acc;
    }
    return 0;
}
int cpu_prior_normal_logprob(const float* x, float* logr, int64_t B, int64_t V, const float* loc,
                             const float* scale) {
    run_sites(PriorLogProbOp{x, loc, scale, V}, B, V, nullptr, logr);
    return 0;
}
int cpu_affine_fwd(const float* x, const float* out, const uint8_t* mask, int parity, int frozen_mode,
                   const float* log_in, float* y, float* log_out, int64_t B, int64_t V) {
    run_sites(AffineOp<0>{x, out, mask, parity == 0 ? 1 : 0, frozen_mode, y, V}, B, V, log_in, log_out);
    return 0;
}
int cpu_affine_inv(const float* x, const float* out, const uint8_t* mask, int parity, int frozen_mode,
                   const float* log_in, float* y, float* log_out, int64_t B, int64_t V) {
    run_sites(AffineOp<1>{x, out, mask, parity == 0 ? 1 : 0, frozen_mode, y, V}, B, V, log_in, log_out);
    return 0;
}
int cpu_affine_bwd(const float* x, const float* out, const uint8_t* mask, int parity, int frozen_mode,
                   const float* gy, const float* glog, float* gx, float* gout, int64_t B, int64_t V) {
    run_sites(AffineBwdOp{x, out, mask, parity == 0 ? 1 : 0, frozen_mode, gy, glog, gx, gout, V}, B, V,
              nullptr, nullptr);
    return 0;
}
int cpu_shift_apply(const float* x, const float* out, const uint8_t* mask, int parity, int frozen_mode,
                    float sign, float* y, int64_t B, int64_t V) {
    run_sites(ShiftOp{x, out, mask, parity == 0 ? 1 : 0, frozen_mode, sign, y, V}, B, V, nullptr, nullptr);
    return 0;
}
int cpu_rqs_fwd(const float* x, const float* out, const uint8_t* mask, int parity, int frozen_mode,
                nfk_rqs_params prm, const float* log_in, float* y, float* log_out, int64_t B, int64_t V) {
    const RqsCfg cfg = to_cfg(prm);
    const int av = parity == 0 ? 1 : 0;
    switch (prm.n_knots) {
#define X(KK) case KK: run_sites(RqsOp<KK, 0>{x, out, mask, av, frozen_mode, cfg, y, V}, B, V, log_in, log_out); return 0;
        FOR_EACH_K(X)
#undef X
    }
    return NFK_EUNSUPPORTED;
}
int cpu_rqs_inv(const float* x, const float* out, const uint8_t* mask, int parity, int frozen_mode,
                nfk_rqs_params prm, const float* log_in, float* y, float* log_out, int64_t B, int64_t V) {
    const RqsCfg cfg = to_cfg(prm);
    const int av = parity == 0 ? 1 : 0;
    switch (prm.n_knots) {
#define X(KK) case KK: run_sites(RqsOp<KK, 1>{x, out, mask, av, frozen_mode, cfg, y, V}, B, V, log_in, log_out); return 0;
        FOR_EACH_K(X)
#undef X
    }
    return NFK_EUNSUPPORTED;
}
int cpu_rqs_bwd(const float* x, const float* out, const uint8_t* mask, int parity, int frozen_mode,
                nfk_rqs_params prm, const float* gy, const float* glog, float* gx, float* gout,
                int64_t B, int64_t V) {
    const RqsCfg cfg = to_cfg(prm);
    const int av = parity == 0 ? 1 : 0;
    switch (prm.n_knots) {
#define X(KK) case KK: run_sites(RqsBwdOp<KK>{x, out, mask, av, frozen_mode, cfg, gy, glog, gx, gout, V}, B, V, nullptr, nullptr); return 0;
        FOR_EACH_K(X)
#undef X
    }
    return NFK_EUNSUPPORTED;
}
int cpu_logistic_fwd(const float* x, int which, const float* log_in, float* y, float* log_out,
                     int64_t B, int64_t V) {
    run_sites(LogisticOp{x, which, y, V}, B, V, log_in, log_out);
    return 0;
}
int cpu_logistic_bwd(const float* x, int which, const float* gy, const float* glog, float* gx,
                     int64_t B, int64_t V) {
    run_sites(LogisticBwdOp{x, which, gy, glog, gx, V}, B, V, nullptr, nullptr);
    return 0;
}
int cpu_spline1d_fwd(const float* x, const float* knots, int K, int extrap_left, int extrap_right,
                     int logistic, int inverse, const float* log_in, float* y, float* log_out,
                     int64_t B, int64_t V) {
    const Spline1dCfg cfg{K, extrap_left, extrap_right, logistic};
    run_sites(Spline1dOp{x, knots, cfg, inverse, y, V}, B, V, log_in, log_out);
    return 0;
}
int cpu_spline1d_bwd(const float* x, const float* knots, int K, int extrap_left, int extrap_right,
                     int logistic, const float* gy, const float* glog, float* gx, float* gk,
                     int64_t B, int64_t V) {
    const Spline1dCfg cfg{K, extrap_left, extrap_right, logistic};
    run_sites(Spline1dBwdOp{x, knots, cfg, gy, glog, gx, gk, V}, B, V, nullptr, nullptr);
    return 0;
}
int cpu_phi4_action_fwd(const float* phi, nfk_lattice lat, float w0, float w2, float w4, float* S, int64_t B) {
    const Lat l = make_lat(lat.ndim, lat.shape);
    int64_t V = 1;
    for (int d = 0; d < lat.ndim; ++d) V *= lat.shape[d];
    run_sites(Phi4Op{phi, l, w0, w2, w4, V}, B, V, nullptr, S);
    return 0;
}
int cpu_phi4_action_bwd(const float* phi, nfk_lattice lat, float w0, float w2, float w4, const float* gS,
                        float* gphi, int64_t B) {
    const Lat l = make_lat(lat.ndim, lat.shape);
    int64_t V = 1;
    for (int d = 0; d < lat.ndim; ++d) V *= lat.shape[d];
    run_sites(Phi4BwdOp{phi, l, w0, w2, w4, gS, gphi, V}, B, V, nullptr, nullptr);
    return 0;
}
// same staging as conv_fwd_kernel<8>: weights re-laid as [Ci*T][8] per channel block
int cpu_conv_circ_fwd(const float* in, const float* w, int w_transposed, const float* bias, const uint8_t* in_mask, int in_keep,
                      int act, const float* dact_from, int dact_kind, float* out, nfk_lattice lat, int ksize,
                      int Ci, int Co, int64_t B) {
    const Lat l = make_lat(lat.ndim, lat.shape);
    int V = 1, T = 1;
    for (int d = 0; d < lat.ndim; ++d) { V *= lat.shape[d]; T *= ksize; }
    constexpr int CO = 8;
    std::vector<float> wt((size_t)Ci * T * CO);
    for (int co0 = 0; co0 < Co; co0 += CO) {
        for (int i = 0; i < Ci * T * CO; ++i) {
            const int co = i % CO, r = i / CO;
            float v = 0.f;
            if (co0 + co < Co) {
                if (!w_transposed) {
                    v = w[(int64_t)(co0 + co) * Ci * T + r];
                } else {
                    const int ci = r / T, t = r % T;
                    v = w[((int64_t)ci * Co + co0 + co) * T + (T - 1 - t)];
                }
            }
            wt[i] = v;
        }
        for (int64_t b = 0; b < B; ++b)
            for (int s = 0; s < V; ++s) {
                float acc[CO];
                for (int co = 0; co < CO; ++co) acc[co] = (bias && co0 + co < Co) ? bias[co0 + co] : 0.f;
                conv_site<CO>(in + b * Ci * (int64_t)V, wt.data(), in_mask, in_keep, l, s, Ci, T, ksize, V, acc);
                for (int co = 0; co < CO && co0 + co < Co; ++co) {
                    const int64_t o = (b * Co + co0 + co) * (int64_t)V + s;
                    float v = act_apply(act, acc[co]);
                    if (dact_from) v *= act_grad_from_post(dact_kind, dact_from[o]);
                    out[o] = v;
                }
            }
    }
    return 0;
}

}  // extern "C"

// host walk through the phases of fused2d_kernel (one "CTA" per sample, items in order)
template <int KIND, int K>
static void fused_host(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                       const float* w3, const float* b3, FusedGeom g, FusedXform xf, const float* log_in,
                       float* y, float* log_out, int64_t B) {
    constexpr int P = KIND == 0 ? 2 : 3 * K - 2;
    constexpr int PP = (P + 3) / 4 * 4;
    std::vector<float> buf(fused_smem_floats<PP>(g.L1, g.R), 0.f);
    const FusedSmem m = fused_carve<PP>(buf.data(), g.L1, g.R);
    for (int e = 0; e < fused_weight_elems<PP>(); ++e) fused_load_weight<P, PP>(m, w1, w2, w3, b1, b2, b3, e);
    const int64_t V = (int64_t)g.L0 * g.L1;
    for (int64_t b = 0; b < B; ++b) {
        float lacc = 0.f;
        for (int r0 = 0; r0 < g.L0; r0 += g.R) {
            const int rows = g.L0 - r0 < g.R ? g.L0 - r0 : g.R;
            for (int e = 0; e < (rows + 6) * g.WS; ++e) fused_load_x(g, m, x + b * V, r0, rows, e);
            for (int it = 0; it < (rows + 4) * g.ncg; ++it)
                fused_hidden_item<1>(g, m.xf, g.R + 6, m.w1s, m.b1s, m.h1s, g.R + 4, rows + 4, it);
            for (int it = 0; it < (rows + 2) * g.ncg; ++it)
                fused_hidden_item<kFH>(g, m.h1s, g.R + 4, m.w2s, m.b2s, m.h2s, g.R + 2, rows + 2, it);
            for (int it = 0; it < rows * g.ncg; ++it) lacc += fused_out_item<KIND, K>(g, m, xf, r0, rows, y + b * V, it);
        }
        if (log_out) log_out[b] = (log_in ? log_in[b] : 0.f) + lacc;
    }
}

extern "C" {

int cpu_fused2d_step(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                     const float* w3, const float* b3, int H, int kind, nfk_rqs_params prm, int mask_parity,
                     int parity, int inverse, const float* log_in, float* y, float* log_out, int L0, int L1,
                     int64_t B, int R) {
    if (H != kFH || L1 % 4 != 0) return NFK_EUNSUPPORTED;
    FusedGeom g;
    g.L0 = L0; g.L1 = L1; g.WS = L1 + 4; g.ncg = L1 / 4; g.R = R > L0 ? L0 : R;
    g.mask_parity = mask_parity; g.active_val = parity == 0 ? 1 : 0;
    FusedXform xf;
    xf.inverse = inverse;
    xf.cfg = to_cfg(prm);
    if (kind == 0) { fused_host<0, 2>(x, w1, b1, w2, b2, w3, b3, g, xf, log_in, y, log_out, B); return 0; }
    switch (prm.n_knots) {
        case 4: fused_host<1, 4>(x, w1, b1, w2, b2, w3, b3, g, xf, log_in, y, log_out, B); return 0;
        case 10: fused_host<1, 10>(x, w1, b1, w2, b2, w3, b3, g, xf, log_in, y, log_out, B); return 0;
    }
    return NFK_EUNSUPPORTED;
}

}  // extern "C"
