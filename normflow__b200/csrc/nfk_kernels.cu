// nfk_kernels.cu -- C-ABI entry points for the pointwise / per-sample kernels:
// masks, prior, affine / shift / RQ-spline couplings, shared 1-D spline chain
// (DistConvertor_), phi^4 action, Metropolis scan, row gather.
// Target: sm_100a (B200).  See include/normflow_b200.h for the contract.

#include "nfk_common.cuh"

namespace nfk {
std::atomic<unsigned long long> g_launches{0};
}
using namespace nfk;

#define NFK_STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" const char* nfk_strerror(int code) {
    switch (code) {
        case NFK_OK: return "ok";
        case NFK_EINVAL: return "invalid argument";
        case NFK_EUNSUPPORTED: return "unsupported configuration";
        case NFK_ECUDA: return "CUDA launch error";
        default: return "unknown error";
    }
}
extern "C" int nfk_version(void) { return 100; }
extern "C" uint64_t nfk_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// ============================================================== masks
__global__ void mask_kernel(uint8_t* mask, Lat lat, int V, int parity, int mu, int along) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= V) return;
    mask[s] = along ? alongaxis_bit(lat, s, parity, mu) : evenodd_bit(lat, s, parity, mu);
}
extern "C" int nfk_mask_evenodd(uint8_t* mask, nfk_lattice lat, int parity, int exclude_mu, void* stream) {
    if (!mask || !lat_ok(lat) || exclude_mu >= lat.ndim) return NFK_EINVAL;
    const int V = (int)lat_volume(lat);
    mask_kernel<<<(V + 255) / 256, 256, 0, NFK_STREAM(stream)>>>(mask, to_lat(lat), V, parity,
                                                                exclude_mu < 0 ? -1 : exclude_mu, 0);
    return check_launch();
}
extern "C" int nfk_mask_alongaxis(uint8_t* mask, nfk_lattice lat, int parity, int mu, void* stream) {
    if (!mask || !lat_ok(lat) || mu < 0 || mu >= lat.ndim) return NFK_EINVAL;
    const int V = (int)lat_volume(lat);
    mask_kernel<<<(V + 255) / 256, 256, 0, NFK_STREAM(stream)>>>(mask, to_lat(lat), V, parity, mu, 1);
    return check_launch();
}
extern "C" int nfk_mask_select(const float* x, const uint8_t* mask, int keep, float* y,
                               int64_t B, int64_t V, void* stream) {
    if (!x || !mask || !y) return NFK_EINVAL;
    return launch_sites(MaskSelectOp{x, mask, keep, y, V}, B, V, nullptr, nullptr, NFK_STREAM(stream));
}

// ============================================================== prior
// Standard normal (no loc / scale), 1024 <= V <= 65536, V % 4 == 0: one CTA per sample, a thread draws quads of
// sites.  The SAME numbers as PriorSampleOp up to the last bits of sqrt (same Philox counters, same uniforms, same
// Box-Muller), with the generic kernel's per-call overheads removed: it executed 58 thread instructions per value
// (ncu: issue slots 80 % busy, profiles/r02_prior_ncu_full.txt) where the draw itself needs about 25.
__global__ void __launch_bounds__(256) prior_normal_fast_kernel(float* __restrict__ x, float* __restrict__ logr, int quads,
                                                                uint64_t seed, uint64_t offset,
                                                                const uint64_t* __restrict__ state) {
    const uint64_t sd = state ? NFK_LDG(state) : seed, of = state ? NFK_LDG(state + 1) : offset;
    const int64_t b = blockIdx.x;
    const uint64_t c_base = (uint64_t)b * (uint64_t)quads;
    float4* xb = reinterpret_cast<float4*>(x) + b * (int64_t)quads;
    float acc = 0.f;
    // the ten round keys depend on the seed only: out of the loop (philox4x32_10 re-derives them per call)
    uint32_t rk0[10], rk1[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        rk0[i] = (uint32_t)sd + (uint32_t)i * 0x9E3779B9u;
        rk1[i] = (uint32_t)(sd >> 32) + (uint32_t)i * 0xBB67AE85u;
    }
#pragma unroll 2
    for (int q = threadIdx.x; q < quads; q += 256) {
        Philox r;
        {
            const uint64_t ctr = c_base + (uint64_t)q;
            uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = (uint32_t)of, c3 = (uint32_t)(of >> 32);
#pragma unroll
            for (int i = 0; i < 10; ++i) {
                uint32_t h0, l0, h1, l1;
                mulhilo(0xD2511F53u, c0, h0, l0);
                mulhilo(0xCD9E8D57u, c2, h1, l1);
                const uint32_t n0 = h1 ^ c1 ^ rk0[i], n2 = h0 ^ c3 ^ rk1[i];
                c0 = n0; c1 = l1; c2 = n2; c3 = l0;
            }
            r.c[0] = c0; r.c[1] = c1; r.c[2] = c2; r.c[3] = c3;
        }
        float z[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float u1 = (float)r.c[2 * h] * 2.3283064365386963e-10f + 1.1641532182693481e-10f;
            const float u2 = (float)r.c[2 * h + 1] * 2.3283064365386963e-10f + 1.1641532182693481e-10f;
            float rad;
            const float t = -1.3862943611198906f * __log2f(u1);             // -2 ln u1 >= 0 (u1 may round to 1)
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(t));
            float sn, cs;
            __sincosf(6.283185307179586f * (u2 - 0.5f), &sn, &cs);
            z[2 * h] = rad * cs;
            z[2 * h + 1] = rad * sn;
        }
        xb[q] = make_float4(z[0], z[1], z[2], z[3]);
        acc = fmaf(z[0], z[0], fmaf(z[1], z[1], fmaf(z[2], z[2], fmaf(z[3], z[3], acc))));
    }
    if (logr == nullptr) return;
    acc = block_sum(acc);
    if (threadIdx.x == 0) logr[b] = -0.5f * acc - 4.f * (float)quads * kLogSqrt2Pi;
}
static bool prior_fast_ok(const float* x, int64_t B, int64_t V, const float* loc, const float* scale) {
    return !loc && !scale && V >= 1024 && V <= 65536 && (V & 3) == 0 && B < (1LL << 31) &&
           (reinterpret_cast<uintptr_t>(x) & 15) == 0;
}

extern "C" int nfk_prior_normal_sample(float* x, float* logr, int64_t B, int64_t V,
                                       const float* loc, const float* scale,
                                       uint64_t seed, uint64_t offset, void* stream) {
    if (!x) return NFK_EINVAL;
    if (B > 0 && prior_fast_ok(x, B, V, loc, scale)) {
        prior_normal_fast_kernel<<<(unsigned)B, 256, 0, NFK_STREAM(stream)>>>(x, logr, (int)(V >> 2), seed, offset, nullptr);
        return check_launch();
    }
    return launch_sites_vec<PriorSampleOp, 4>(PriorSampleOp{x, loc, scale, V, seed, offset, nullptr}, B, V,
                                              nullptr, logr, NFK_STREAM(stream));
}

// The same draw with the Philox key and stream offset read from DEVICE memory, and the offset
// advanced by a one-thread kernel behind it: a captured CUDA graph then draws fresh numbers at
// every replay (host-side (seed, offset) arguments would be frozen into the graph).
static __global__ void advance_offset_kernel(uint64_t* state) { state[1] += 1; }
extern "C" int nfk_prior_normal_sample_dev(float* x, float* logr, int64_t B, int64_t V,
                                           const float* loc, const float* scale,
                                           uint64_t* state, void* stream) {
    if (!x || !state) return NFK_EINVAL;
    int rc;
    if (B > 0 && prior_fast_ok(x, B, V, loc, scale)) {
        prior_normal_fast_kernel<<<(unsigned)B, 256, 0, NFK_STREAM(stream)>>>(x, logr, (int)(V >> 2), 0, 0, state);
        rc = check_launch();
    } else {
        rc = launch_sites_vec<PriorSampleOp, 4>(PriorSampleOp{x, loc, scale, V, 0, 0, state}, B, V,
                                                nullptr, logr, NFK_STREAM(stream));
    }
    if (rc != NFK_OK) return rc;
    advance_offset_kernel<<<1, 1, 0, NFK_STREAM(stream)>>>(state);
    return check_launch();
}
// Standard normal, mid-sized samples: ONE WARP per sample (8 loads of 16 bytes in flight per lane, a shuffle
// reduction, no block barrier) -- a CTA per sample spends as long in its two barriers as in its 16 KB of loads.
__global__ void __launch_bounds__(256) prior_logprob_warp_kernel(const float* __restrict__ x, float* __restrict__ logr,
                                                                int64_t B, int64_t V) {
    const int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= B) return;
    const int lane = threadIdx.x & 31;
    const float4* p = reinterpret_cast<const float4*>(x + b * V);
    const int nq = (int)(V >> 2);
    float acc = 0.f;
#pragma unroll 8
    for (int q = lane; q < nq; q += 32) {
        const float4 v = __ldg(p + q);
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    acc = warp_sum(acc);
    if (lane == 0) logr[b] = -0.5f * acc - (float)V * kLogSqrt2Pi;
}

extern "C" int nfk_prior_normal_logprob(const float* x, float* logr, int64_t B, int64_t V,
                                        const float* loc, const float* scale, void* stream) {
    if (!x || !logr) return NFK_EINVAL;
    if (!loc && !scale && B >= 1024 && V % 4 == 0 && V >= 512 && V <= 32768 && ((uintptr_t)x % 16) == 0) {
        prior_logprob_warp_kernel<<<(unsigned)((B + 7) / 8), 256, 0, NFK_STREAM(stream)>>>(x, logr, B, V);
        return check_launch();
    }
    return launch_sites_vec<PriorLogProbOp4, 4>(PriorLogProbOp4{x, loc, scale, V}, B, V, nullptr, logr, NFK_STREAM(stream));
}

// ============================================================== affine / shift
static bool affine_vec_ok(const float* x, const float* out, const uint8_t* mask, const float* y, int64_t V) {
    return V % 4 == 0 && V > kSmallV && ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0 &&
           ((uintptr_t)y % 16) == 0 && ((uintptr_t)mask % 4) == 0;
}
extern "C" int nfk_affine_fwd(const float* x, const float* out, const uint8_t* mask, int parity,
                              int frozen_mode, const float* log_in, float* y, float* log_out,
                              int64_t B, int64_t V, void* stream) {
    if (!x || !out || !mask || !y) return NFK_EINVAL;
    if (affine_vec_ok(x, out, mask, y, V))
        return launch_sites_vec<AffineOp4<0>, 4>(AffineOp4<0>{x, out, mask, parity == 0 ? 1 : 0, frozen_mode, y, V},
                                                 B, V, log_in, log_out, NFK_STREAM(stream));
    return launch_sites(AffineOp<0>{x, out, mask, parity == 0 ? 1 : 0, frozen_mode, y, V}, B, V, log_in,
                        log_out, NFK_STREAM(stream));
}
extern "C" int nfk_affine_inv(const float* x, const float* out, const uint8_t* mask, int parity,
                              int frozen_mode, const float* log_in, float* y, float* log_out,
                              int64_t B, int64_t V, void* stream) {
    if (!x || !out || !mask || !y) return NFK_EINVAL;
    if (affine_vec_ok(x, out, mask, y, V))
        return launch_sites_vec<AffineOp4<1>, 4>(AffineOp4<1>{x, out, mask, parity == 0 ? 1 : 0, frozen_mode, y, V},
                                                 B, V, log_in, log_out, NFK_STREAM(stream));
    return launch_sites(AffineOp<1>{x, out, mask, parity == 0 ? 1 : 0, frozen_mode, y, V}, B, V, log_in,
                        log_out, NFK_STREAM(stream));
}
extern "C" int nfk_affine_bwd(const float* x, const float* out, const uint8_t* mask, int parity,
                              int frozen_mode, const float* gy, const float* glog,
                              float* gx, float* gout, int64_t B, int64_t V, void* stream) {
    if (!x || !out || !mask || !gy || !gx || !gout) return NFK_EINVAL;
    return launch_sites(AffineBwdOp{x, out, mask, parity == 0 ? 1 : 0, frozen_mode, gy, glog, gx, gout, V},
                        B, V, nullptr, nullptr, NFK_STREAM(stream));
}
extern "C" int nfk_shift_apply(const float* x, const float* out, const uint8_t* mask, int parity,
                               int frozen_mode, float sign, float* y, int64_t B, int64_t V, void* stream) {
    if (!x || !out || !mask || !y) return NFK_EINVAL;
    return launch_sites(ShiftOp{x, out, mask, parity == 0 ? 1 : 0, frozen_mode, sign, y, V}, B, V, nullptr,
                        nullptr, NFK_STREAM(stream));
}

// ============================================================== RQ-spline coupling
static bool rqs_cfg(const nfk_rqs_params& p, RqsCfg& c) {
    if (p.n_knots < 2 || !(p.xlim1 > p.xlim0) || !(p.ylim1 > p.ylim0)) return false;
    if (p.extrap_left != NFK_EXTRAP_NONE && p.extrap_left != NFK_EXTRAP_LINEAR) return false;
    if (p.extrap_right != NFK_EXTRAP_NONE && p.extrap_right != NFK_EXTRAP_LINEAR) return false;
    c.xlim0 = p.xlim0; c.xw = p.xlim1 - p.xlim0;
    c.ylim0 = p.ylim0; c.yw = p.ylim1 - p.ylim0;
    c.left = p.extrap_left; c.right = p.extrap_right;
    return true;
}

// the knot count is a template parameter (per-site knot arrays live in registers)
#define NFK_FOR_EACH_K(X) \
    X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(14) X(16) X(20) X(24) X(32)

// Two adjacent sites per thread: on a checkerboard one of them runs the spline and the other is a copy, so no
// lane of a warp idles through the long branch (with one thread per site every other lane does).
template <class Op>
struct PairSites {
    Op op;
    NFK_HD float operator()(int64_t b, int64_t p) const { return op(b, 2 * p) + op(b, 2 * p + 1); }
};
template <class Op>
static int launch_site_pairs(const Op& op, int64_t B, int64_t V, const float* log_in, float* log_out, cudaStream_t st) {
    if (V % 2 == 0 && V >= 128) return launch_sites(PairSites<Op>{op}, B, V / 2, log_in, log_out, st);
    return launch_sites(op, B, V, log_in, log_out, st);
}

template <int MODE>
static int rqs_apply(const float* x, const float* out, const uint8_t* mask, int parity, int frozen_mode,
                     nfk_rqs_params prm, const float* log_in, float* y, float* log_out,
                     int64_t B, int64_t V, cudaStream_t st) {
    RqsCfg cfg;
    if (!x || !out || !mask || !y) return NFK_EINVAL;
    if (!rqs_cfg(prm, cfg)) return NFK_EINVAL;
    const int av = parity == 0 ? 1 : 0;
    switch (prm.n_knots) {
#define X(KK) case KK: return launch_site_pairs(RqsOp<KK, MODE>{x, out, mask, av, frozen_mode, cfg, y, V}, B, V, log_in, log_out, st);
        NFK_FOR_EACH_K(X)
#undef X
        default: return NFK_EUNSUPPORTED;
    }
}
extern "C" int nfk_rqs_fwd(const float* x, const float* out, const uint8_t* mask, int parity,
                           int frozen_mode, nfk_rqs_params prm, const float* log_in,
                           float* y, float* log_out, int64_t B, int64_t V, void* stream) {
    return rqs_apply<0>(x, out, mask, parity, frozen_mode, prm, log_in, y, log_out, B, V, NFK_STREAM(stream));
}
extern "C" int nfk_rqs_inv(const float* x, const float* out, const uint8_t* mask, int parity,
                           int frozen_mode, nfk_rqs_params prm, const float* log_in,
                           float* y, float* log_out, int64_t B, int64_t V, void* stream) {
    return rqs_apply<1>(x, out, mask, parity, frozen_mode, prm, log_in, y, log_out, B, V, NFK_STREAM(stream));
}
// Spline VJP with TWO adjacent sites per thread.  On a checkerboard exactly one site of such a pair is
// active: one thread per site leaves every other lane of a warp idle through the (long) spline adjoint, while
// one thread per pair keeps all lanes busy and writes each gradient channel as one 8-byte store (value, 0).
struct PairStore {
    float2* base;        // &gout[b][0][s0] viewed as pairs
    int64_t stride2;     // V / 2
    int slot;            // which site of the pair is active
    __device__ __forceinline__ void operator()(int c, float v) const {
        base[c * stride2] = slot == 0 ? make_float2(v, 0.f) : make_float2(0.f, v);
    }
};

template <int K>
__global__ void __launch_bounds__(256) rqs_bwd_pair_kernel(RqsBwdOp<K> op, int64_t n_pairs, int64_t half_v) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_pairs) return;
    const int64_t b = idx / half_v, s0 = 2 * (idx - b * half_v);
    const bool a0 = site_active(op.mask, s0, op.active_val), a1 = site_active(op.mask, s0 + 1, op.active_val);
    if (a0 == a1) {                          // not a checkerboard pair: the two sites one after the other
        op(b, s0);
        op(b, s0 + 1);
        return;
    }
    const int slot = a0 ? 0 : 1;
    const int64_t i0 = b * op.V + s0;
    const float2 g = *reinterpret_cast<const float2*>(op.gy + i0);
    const int64_t o = (int64_t)(3 * K - 2) * b * op.V + s0;
    const ChanLoad ld{op.out + o + slot, op.V};
    const PairStore st{reinterpret_cast<float2*>(op.gout + o), half_v, slot};
    const float gl = op.glog ? __ldg(op.glog + b) : 0.f;
    const float ga = rqs_site_backward<K>(ld, op.cfg, __ldg(op.x + i0 + slot), slot == 0 ? g.x : g.y, gl, st);
    const float gf = op.frozen_copy ? (slot == 0 ? g.y : g.x) : 0.f;
    *reinterpret_cast<float2*>(op.gx + i0) = slot == 0 ? make_float2(ga, gf) : make_float2(gf, ga);
}

template <int K>
static int launch_rqs_bwd(const RqsBwdOp<K>& op, int64_t B, int64_t V, cudaStream_t st) {
    const bool pairs = V % 2 == 0 && V >= 64 &&
                       (((uintptr_t)op.gy | (uintptr_t)op.gx | (uintptr_t)op.gout) % 8 == 0);
    if (!pairs) return launch_sites(op, B, V, nullptr, nullptr, st);
    const int64_t n_pairs = B * (V / 2);
    if (n_pairs <= 0) return NFK_OK;
    rqs_bwd_pair_kernel<K><<<(unsigned)((n_pairs + 255) / 256), 256, 0, st>>>(op, n_pairs, V / 2);
    return check_launch();
}

extern "C" int nfk_rqs_bwd(const float* x, const float* out, const uint8_t* mask, int parity,
                           int frozen_mode, nfk_rqs_params prm, const float* gy, const float* glog,
                           float* gx, float* gout, int64_t B, int64_t V, void* stream) {
    RqsCfg cfg;
    if (!x || !out || !mask || !gy || !gx || !gout) return NFK_EINVAL;
    if (!rqs_cfg(prm, cfg)) return NFK_EINVAL;
    const int av = parity == 0 ? 1 : 0;
    switch (prm.n_knots) {
#define X(KK) case KK: return launch_rqs_bwd(RqsBwdOp<KK>{x, out, mask, av, frozen_mode, cfg, gy, glog, gx, gout, V}, B, V, NFK_STREAM(stream));
        NFK_FOR_EACH_K(X)
#undef X
        default: return NFK_EUNSUPPORTED;
    }
}

// ============================================================== Expit_ / Logit_
extern "C" int nfk_logistic_fwd(const float* x, int which, const float* log_in, float* y, float* log_out,
                                int64_t B, int64_t V, void* stream) {
    if (!x || !y || (which != 0 && which != 1)) return NFK_EINVAL;
    return launch_sites(LogisticOp{x, which, y, V}, B, V, log_in, log_out, NFK_STREAM(stream));
}
extern "C" int nfk_logistic_bwd(const float* x, int which, const float* gy, const float* glog, float* gx,
                                int64_t B, int64_t V, void* stream) {
    if (!x || !gy || !gx || (which != 0 && which != 1)) return NFK_EINVAL;
    return launch_sites(LogisticBwdOp{x, which, gy, glog, gx, V}, B, V, nullptr, nullptr, NFK_STREAM(stream));
}

// ============================================================== shared 1-D spline chain
// The knots (3K floats) are staged in shared memory by every CTA; elements are
// walked grid-stride over the flattened [B*V] array when V is tiny, else with the
// per-sample plan so the log-Jacobian reduces without atomics.
struct SplineArgs {
    const float *x, *knots;      // knots: [5][K] = kx | ky | kd | cx | cy
    Spline1dCfg cfg;
    int inverse;
    const float *gy, *glog;
    float *y, *gx, *gk;     // gk: global [5K] accumulation (backward)
    const float* log_in;
    float* log_out;
    int64_t B, V;
    int chunks;
    int64_t chunk_len;
};

template <bool BWD, bool SMALL>
__global__ void __launch_bounds__(256) spline1d_kernel(SplineArgs a) {
    __shared__ float knots[5 * NFK_MAX_KNOTS];
    __shared__ float gacc[5 * NFK_MAX_KNOTS];
    const int K = a.cfg.K;
    const int nk = a.cfg.logistic ? 5 * K : 3 * K;
    for (int i = threadIdx.x; i < 5 * K; i += blockDim.x) knots[i] = i < nk ? a.knots[i] : 0.f;
    if (BWD)
        for (int i = threadIdx.x; i < 5 * K; i += blockDim.x) gacc[i] = 0.f;
    __syncthreads();

    if (SMALL) {            // one thread per sample
        const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (b < a.B) {
            float acc = 0.f;
            for (int64_t s = 0; s < a.V; ++s) {
                if (BWD) Spline1dBwdOp{a.x, knots, a.cfg, a.gy, a.glog, a.gx, gacc, a.V}(b, s);
                else acc += Spline1dOp{a.x, knots, a.cfg, a.inverse, a.y, a.V}(b, s);
            }
            if (!BWD && a.log_out) a.log_out[b] = (a.log_in ? a.log_in[b] : 0.f) + acc;
        }
    } else {
        const int64_t b = blockIdx.x / a.chunks;
        const int chunk = (int)(blockIdx.x % a.chunks);
        const int64_t s0 = chunk * a.chunk_len;
        const int64_t s1 = s0 + a.chunk_len < a.V ? s0 + a.chunk_len : a.V;
        float acc = 0.f;
        for (int64_t s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
            if (BWD) Spline1dBwdOp{a.x, knots, a.cfg, a.gy, a.glog, a.gx, gacc, a.V}(b, s);
            else acc += Spline1dOp{a.x, knots, a.cfg, a.inverse, a.y, a.V}(b, s);
        }
        if (!BWD && a.log_out) {
            acc = block_sum(acc);
            if (threadIdx.x == 0) {
                if (a.chunks == 1) a.log_out[b] = (a.log_in ? a.log_in[b] : 0.f) + acc;
                else atomicAdd(a.log_out + b, acc);
            }
        }
    }
    if (BWD) {
        __syncthreads();
        for (int i = threadIdx.x; i < 5 * K; i += blockDim.x)
            if (gacc[i] != 0.f) atomicAdd(a.gk + i, gacc[i]);
    }
}

static bool spline_cfg(int K, int left, int right, int logistic, Spline1dCfg& c) {
    if (K < 2 || K > NFK_MAX_KNOTS) return false;
    if (left < 0 || left > NFK_EXTRAP_PERIODIC || right < 0 || right > NFK_EXTRAP_PERIODIC) return false;
    if (logistic && (right != NFK_EXTRAP_NONE || left == NFK_EXTRAP_LINEAR || left == NFK_EXTRAP_PERIODIC)) return false;
    c.K = K; c.left = left; c.right = right; c.logistic = logistic;
    return true;
}

template <bool BWD>
static int spline1d_launch(SplineArgs a, cudaStream_t st) {
    if (a.B <= 0 || a.V <= 0) return NFK_OK;
    const Plan p = make_plan(a.V, 1);
    a.chunks = p.chunks;
    a.chunk_len = p.chunk_len;
    if (p.small) {
        spline1d_kernel<BWD, true><<<(unsigned)((a.B + 127) / 128), 128, 0, st>>>(a);
        return check_launch();
    }
    if (!BWD && p.chunks > 1 && a.log_out) {
        init_log_kernel<<<(unsigned)((a.B + 255) / 256), 256, 0, st>>>(a.log_in, a.log_out, a.B);
        if (int e = check_launch()) return e;
    }
    spline1d_kernel<BWD, false><<<(unsigned)(a.B * p.chunks), p.threads, 0, st>>>(a);
    return check_launch();
}

extern "C" int nfk_spline1d_fwd(const float* x, const float* knots, int K, int extrap_left, int extrap_right,
                                int logistic, int inverse, const float* log_in, float* y, float* log_out,
                                int64_t B, int64_t V, void* stream) {
    SplineArgs a{};
    if (!x || !knots || !y) return NFK_EINVAL;
    if (!spline_cfg(K, extrap_left, extrap_right, logistic, a.cfg)) return NFK_EINVAL;
    if (inverse && (extrap_left == NFK_EXTRAP_PERIODIC || extrap_right == NFK_EXTRAP_PERIODIC)) return NFK_EINVAL;
    a.x = x; a.knots = knots; a.inverse = inverse;
    a.y = y; a.log_in = log_in; a.log_out = log_out; a.B = B; a.V = V;
    return spline1d_launch<false>(a, NFK_STREAM(stream));
}
extern "C" int nfk_spline1d_bwd(const float* x, const float* knots, int K, int extrap_left, int extrap_right,
                                int logistic, const float* gy, const float* glog,
                                float* gx, float* gknots, int64_t B, int64_t V, void* stream) {
    SplineArgs a{};
    if (!x || !knots || !gy || !gx || !gknots) return NFK_EINVAL;
    if (!spline_cfg(K, extrap_left, extrap_right, logistic, a.cfg)) return NFK_EINVAL;
    a.x = x; a.knots = knots; a.gy = gy; a.glog = glog;
    a.gx = gx; a.gk = gknots; a.B = B; a.V = V;
    return spline1d_launch<true>(a, NFK_STREAM(stream));
}

// ============================================================== phi^4 action
// 2-D lattices with L1 % 4 == 0: one CTA per sample, a thread owns 4 consecutive columns of a row
// (128-bit loads of its row and the row above, one scalar for the left neighbour; the re-reads are
// L1/L2 hits), so the kernel moves 4 bytes per site from HBM and does no index division per site.
template <bool BWD>
__global__ void __launch_bounds__(256) phi4_2d_kernel(const float* phi, int L0, int L1, float w0, float w2, float w4,
                                                      const float* gS, float* out) {
    const int64_t b = blockIdx.x;
    const float* p = phi + b * (int64_t)L0 * L1;
    const int nq = L1 >> 2;                                   // column quads per row
    const float gs = BWD ? __ldg(gS + b) : 0.f;
    float acc = 0.f;
#pragma unroll 4
    for (int it = threadIdx.x; it < L0 * nq; it += blockDim.x) {
        const int r = it / nq, c0 = (it - r * nq) * 4;
        const int ru = r == 0 ? L0 - 1 : r - 1;
        const float4 v = __ldg(reinterpret_cast<const float4*>(p + r * L1 + c0));
        const float4 u = __ldg(reinterpret_cast<const float4*>(p + ru * L1 + c0));
        const float lf = __ldg(p + r * L1 + (c0 == 0 ? L1 - 1 : c0 - 1));
        if (!BWD) {
            // sum_x w2 v^2 + w4 v^4 - w0 v (v(x - e0) + v(x - e1))   (scalar_action.py:40-46)
            const float q0 = v.x * v.x, q1 = v.y * v.y, q2 = v.z * v.z, q3 = v.w * v.w;
            acc += q0 * (w2 + w4 * q0) + q1 * (w2 + w4 * q1) + q2 * (w2 + w4 * q2) + q3 * (w2 + w4 * q3);
            acc -= w0 * (v.x * (u.x + lf) + v.y * (u.y + v.x) + v.z * (u.z + v.y) + v.w * (u.w + v.z));
        } else {
            const int rd = r == L0 - 1 ? 0 : r + 1;
            const float4 d = __ldg(reinterpret_cast<const float4*>(p + rd * L1 + c0));
            const float rt = __ldg(p + r * L1 + (c0 + 4 == L1 ? 0 : c0 + 4));
            float4 g;
            g.x = gs * (v.x * (2.f * w2 + 4.f * w4 * v.x * v.x) - w0 * (u.x + d.x + lf + v.y));
            g.y = gs * (v.y * (2.f * w2 + 4.f * w4 * v.y * v.y) - w0 * (u.y + d.y + v.x + v.z));
            g.z = gs * (v.z * (2.f * w2 + 4.f * w4 * v.z * v.z) - w0 * (u.z + d.z + v.y + v.w));
            g.w = gs * (v.w * (2.f * w2 + 4.f * w4 * v.w * v.w) - w0 * (u.w + d.w + v.z + rt));
            *reinterpret_cast<float4*>(out + b * (int64_t)L0 * L1 + r * L1 + c0) = g;
        }
    }
    if (!BWD) {
        acc = block_sum(acc);
        if (threadIdx.x == 0) out[b] = acc;
    }
}
// forward action, one warp per sample (same arithmetic as phi4_2d_kernel<false>; no block barrier)
// SH >= 0: the number of column quads per row is 2^SH (row / column from shifts instead of a division)
template <int SH>
__global__ void __launch_bounds__(256) phi4_2d_warp_kernel(const float* phi, int L0, int L1, float w0, float w2, float w4,
                                                           float* out, int64_t B) {
    const int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= B) return;
    const int lane = threadIdx.x & 31;
    const float* p = phi + b * (int64_t)L0 * L1;
    const int nq = SH >= 0 ? (1 << SH) : (L1 >> 2);
    float acc = 0.f;
#pragma unroll 4
    for (int it = lane; it < L0 * nq; it += 32) {
        const int r = SH >= 0 ? (it >> SH) : it / nq, c0 = (it - r * nq) * 4;
        const int ru = r == 0 ? L0 - 1 : r - 1;
        const float4 v = __ldg(reinterpret_cast<const float4*>(p + r * L1 + c0));
        const float4 u = __ldg(reinterpret_cast<const float4*>(p + ru * L1 + c0));
        const float lf = __ldg(p + r * L1 + (c0 == 0 ? L1 - 1 : c0 - 1));
        const float q0 = v.x * v.x, q1 = v.y * v.y, q2 = v.z * v.z, q3 = v.w * v.w;
        acc += q0 * (w2 + w4 * q0) + q1 * (w2 + w4 * q1) + q2 * (w2 + w4 * q2) + q3 * (w2 + w4 * q3);
        acc -= w0 * (v.x * (u.x + lf) + v.y * (u.y + v.x) + v.z * (u.z + v.y) + v.w * (u.w + v.z));
    }
    acc = warp_sum(acc);
    if (lane == 0) out[b] = acc;
}

// forward action, one warp per sample, every site loaded ONCE: a lane owns one column quad over a contiguous
// range of rows and walks down it (the row above is the previous iteration's registers, the left neighbour one
// shuffle).  Needs 32 % (L1 / 4) == 0 and L0 divisible by the 32 / (L1 / 4) row ranges of a warp.
__global__ void __launch_bounds__(256) phi4_2d_colwalk_kernel(const float* phi, int L0, int L1, float w0, float w2,
                                                              float w4, float* out, int64_t B) {
    const int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= B) return;                                      // (whole warps leave together)
    const int lane = threadIdx.x & 31;
    const int nq = L1 >> 2, g = lane / nq, q = lane - g * nq;
    const int rows_per = L0 / (32 / nq), r_begin = g * rows_per;
    const int src_left = g * nq + (q == 0 ? nq - 1 : q - 1);
    const float* col = phi + b * (int64_t)L0 * L1 + 4 * q;
    float4 up = __ldg(reinterpret_cast<const float4*>(col + (int64_t)((r_begin == 0 ? L0 : r_begin) - 1) * L1));
    float acc = 0.f;
#pragma unroll 8
    for (int i = 0; i < rows_per; ++i) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(col + (int64_t)(r_begin + i) * L1));
        const float lf = __shfl_sync(0xffffffffu, v.w, src_left);
        const float q0 = v.x * v.x, q1 = v.y * v.y, q2 = v.z * v.z, q3 = v.w * v.w;
        acc += q0 * (w2 + w4 * q0) + q1 * (w2 + w4 * q1) + q2 * (w2 + w4 * q2) + q3 * (w2 + w4 * q3);
        acc -= w0 * (v.x * (up.x + lf) + v.y * (up.y + v.x) + v.z * (up.z + v.y) + v.w * (up.w + v.z));
        up = v;
    }
    acc = warp_sum(acc);
    if (lane == 0) out[b] = acc;
}

static bool phi4_2d_ok(const nfk_lattice& lat, const float* phi, const float* out) {
    return lat.ndim == 2 && lat.shape[1] % 4 == 0 && lat.shape[1] >= 4 && lat.shape[0] >= 2 &&
           ((uintptr_t)phi % 16) == 0 && ((uintptr_t)out % 16) == 0;
}

extern "C" int nfk_phi4_action_fwd(const float* phi, nfk_lattice lat, float w0, float w2, float w4,
                                   float* S, int64_t B, void* stream) {
    if (!phi || !S || !lat_ok(lat)) return NFK_EINVAL;
    const int64_t V = lat_volume(lat);
    if (B >= 1024 && phi4_2d_ok(lat, phi, phi) && V >= 512 && V <= 32768) {
        const unsigned blocks = (unsigned)((B + 7) / 8);
        const int nq = lat.shape[1] / 4;
        if (nq <= 32 && 32 % nq == 0 && lat.shape[0] % (32 / nq) == 0) {
            phi4_2d_colwalk_kernel<<<blocks, 256, 0, NFK_STREAM(stream)>>>(phi, lat.shape[0], lat.shape[1], w0, w2, w4, S, B);
            return check_launch();
        }
        if (lat.shape[1] == 64)
            phi4_2d_warp_kernel<4><<<blocks, 256, 0, NFK_STREAM(stream)>>>(phi, lat.shape[0], 64, w0, w2, w4, S, B);
        else if (lat.shape[1] == 32)
            phi4_2d_warp_kernel<3><<<blocks, 256, 0, NFK_STREAM(stream)>>>(phi, lat.shape[0], 32, w0, w2, w4, S, B);
        else
            phi4_2d_warp_kernel<-1><<<blocks, 256, 0, NFK_STREAM(stream)>>>(phi, lat.shape[0], lat.shape[1], w0, w2, w4, S, B);
        return check_launch();
    }
    if (B > 0 && phi4_2d_ok(lat, phi, phi)) {
        const int items = lat.shape[0] * (lat.shape[1] / 4);
        const int th = items >= 256 ? 256 : (items + 31) / 32 * 32;
        phi4_2d_kernel<false><<<(unsigned)B, th, 0, NFK_STREAM(stream)>>>(phi, lat.shape[0], lat.shape[1], w0, w2, w4,
                                                                           nullptr, S);
        return check_launch();
    }
    return launch_sites(Phi4Op{phi, to_lat(lat), w0, w2, w4, V}, B, V, nullptr, S, NFK_STREAM(stream));
}
extern "C" int nfk_phi4_action_bwd(const float* phi, nfk_lattice lat, float w0, float w2, float w4,
                                   const float* gS, float* gphi, int64_t B, void* stream) {
    if (!phi || !gS || !gphi || !lat_ok(lat)) return NFK_EINVAL;
    const int64_t V = lat_volume(lat);
    if (B > 0 && phi4_2d_ok(lat, phi, gphi)) {
        const int items = lat.shape[0] * (lat.shape[1] / 4);
        const int th = items >= 256 ? 256 : (items + 31) / 32 * 32;
        phi4_2d_kernel<true><<<(unsigned)B, th, 0, NFK_STREAM(stream)>>>(phi, lat.shape[0], lat.shape[1], w0, w2, w4,
                                                                          gS, gphi);
        return check_launch();
    }
    return launch_sites(Phi4BwdOp{phi, to_lat(lat), w0, w2, w4, gS, gphi, V}, B, V, nullptr, nullptr,
                        NFK_STREAM(stream));
}

// ============================================================== Metropolis scan
// One warp.  The chain is sequential, but its inputs are not: the lanes stream
// l = logq - logp and log u into shared memory 32 at a time (double precision,
// mcmc.py:64 hands float64 to numpy), lane 0 walks the 32 decisions, and the
// lanes then publish accept flags / indices coalesced.
__global__ void metropolis_kernel(const float* logq, const float* logp, const double* log_u,
                                  double* ref_inout, uint8_t* accept, int64_t* idx, int64_t* n_accept,
                                  int64_t B) {
    __shared__ double sl[32], su[32];
    __shared__ uint8_t sa[32];
    __shared__ int64_t si[32];
    const int lane = threadIdx.x;
    double ref = ref_inout[0];
    bool has_ref = ref_inout[1] != 0.0;
    int64_t last = -1, count = 0;
    for (int64_t base = 0; base < B; base += 32) {
        const int64_t i = base + lane;
        if (i < B) {
            sl[lane] = (double)logq[i] - (double)logp[i];
            su[lane] = log_u[i];
        }
        __syncwarp();
        if (lane == 0) {
            const int n = B - base < 32 ? (int)(B - base) : 32;
            for (int k = 0; k < n; ++k) {
                const double l = sl[k];
                // mcmc.py:308-309: the proposal that initialises the chain is its own reference, and
                // log u < 0 accepts it -- made unconditional so that a NaN log q - log p (a diverged
                // model) cannot leave the chain without a state for the gather below
                const bool init = !has_ref;
                if (init) { ref = l; has_ref = true; }
                const bool acc = init || su[k] < ref - l;       // mcmc.py:313
                if (acc) { ref = l; last = base + k; ++count; }
                sa[k] = acc;
                si[k] = last;
            }
        }
        __syncwarp();
        if (i < B) {
            accept[i] = sa[lane];
            idx[i] = si[lane];
        }
        __syncwarp();
    }
    if (lane == 0) {
        ref_inout[0] = ref;
        ref_inout[1] = has_ref ? 1.0 : 0.0;
        if (n_accept) *n_accept = count;
    }
}
extern "C" int nfk_metropolis_scan(const float* logq, const float* logp, const double* log_u,
                                   double* ref_inout, uint8_t* accept, int64_t* idx, int64_t* n_accept,
                                   int64_t B, void* stream) {
    if (!logq || !logp || !log_u || !ref_inout || !accept || !idx) return NFK_EINVAL;
    if (B <= 0) return NFK_OK;
    metropolis_kernel<<<1, 32, 0, NFK_STREAM(stream)>>>(logq, logp, log_u, ref_inout, accept, idx, n_accept, B);
    return check_launch();
}

// ---- acceptance rate of R shuffled chains (MCMCSampler.estimate_accept_rate, mcmc.py:117-124)
// The reference builds R chains from R random permutations of the same log q - log p values and
// walks each one in a host loop; here chain r is one CTA of one warp: the lanes gather
// l[perm[r][i]] and log u[r][i] 32 at a time, lane 0 walks the decisions (as metropolis_kernel).
// Every chain starts without a reference, so its first proposal is accepted (mcmc.py:308-313).
__global__ void metropolis_rates_kernel(const double* l, const int64_t* perm, const double* log_u,
                                        double* rates, int64_t N) {
    __shared__ double sl[32], su[32];
    const int lane = threadIdx.x;
    const int64_t* pr = perm ? perm + (int64_t)blockIdx.x * N : nullptr;
    const double* ur = log_u + (int64_t)blockIdx.x * N;
    double ref = 0.0;
    int64_t count = 0;
    for (int64_t base = 0; base < N; base += 32) {
        const int64_t i = base + lane;
        if (i < N) {
            sl[lane] = l[pr ? pr[i] : i];
            su[lane] = ur[i];
        }
        __syncwarp();
        if (lane == 0) {
            const int n = N - base < 32 ? (int)(N - base) : 32;
            for (int k = 0; k < n; ++k) {
                const double v = sl[k];
                const bool init = base == 0 && k == 0;
                if (init) ref = v;
                if (init || su[k] < ref - v) { ref = v; ++count; }
            }
        }
        __syncwarp();
    }
    if (lane == 0) rates[blockIdx.x] = (double)count / (double)N;
}
extern "C" int nfk_metropolis_rates(const double* logqp, const int64_t* perm, const double* log_u,
                                    double* rates, int64_t N, int64_t R, void* stream) {
    if (!logqp || !log_u || !rates) return NFK_EINVAL;
    if (N <= 0 || R <= 0) return NFK_OK;
    metropolis_rates_kernel<<<(unsigned)R, 32, 0, NFK_STREAM(stream)>>>(logqp, perm, log_u, rates, N);
    return check_launch();
}

// ============================================================== row gather
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* src, const int64_t* idx, const float* prev,
                                                          float* dst, int64_t row, int vec_ok) {
    const int64_t i = blockIdx.x;
    const int64_t j = idx[i];
    const float* from = (j >= 0 || prev == nullptr) ? src + (j >= 0 ? j : 0) * row : prev;   // no state yet: row 0, as mcmc.py:67-68
    float* to = dst + i * row;
    if (vec_ok) {
        const float4* f4 = reinterpret_cast<const float4*>(from);
        float4* t4 = reinterpret_cast<float4*>(to);
        for (int64_t k = threadIdx.x; k < row / 4; k += blockDim.x) t4[k] = f4[k];
    } else {
        for (int64_t k = threadIdx.x; k < row; k += blockDim.x) to[k] = from[k];
    }
}
__global__ void gather_scalar_kernel(const float* src, const int64_t* idx, const float* prev, float* dst, int64_t B) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const int64_t j = idx[i];
    dst[i] = j >= 0 ? src[j] : (prev ? prev[0] : src[0]);
}
extern "C" int nfk_gather_rows(const float* src, const int64_t* idx, const float* prev, float* dst,
                               int64_t B, int64_t row_elems, void* stream) {
    if (!src || !idx || !dst || row_elems < 1) return NFK_EINVAL;
    if (B <= 0) return NFK_OK;
    if (row_elems == 1) {
        gather_scalar_kernel<<<(unsigned)((B + 255) / 256), 256, 0, NFK_STREAM(stream)>>>(src, idx, prev, dst, B);
        return check_launch();
    }
    const bool aligned = ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0) &&
                         (prev == nullptr || (uintptr_t)prev % 16 == 0) && row_elems % 4 == 0;
    const int threads = row_elems >= 1024 ? 256 : (row_elems >= 256 ? 64 : 32);
    gather_rows_kernel<<<(unsigned)B, threads, 0, NFK_STREAM(stream)>>>(src, idx, prev, dst, row_elems, aligned ? 1 : 0);
    return check_launch();
}
