// nfk_psd.cu -- spectral part of PSDBlock_ / FFTNet_ (reference src/nn/scalar/psd_.py:17-40,
// fftflow_.py:121-131,167-180).  The FFTs themselves stay cuFFT (called by the host
// package); these kernels are what sits between rfftn and irfftn:
//
//   psd_weights : ipsd[Kc] -> w = ipsd^-1/2 (or ^+1/2 for the inverse map) and the
//                 log-Jacobian  sum_k m_k log w_k,  m_k = 2 - [col==0] - [col==Lh-1]
//                 (half-spectrum multiplicities, fftflow_.py:172-178).
//   psd_scale   : Y[b,k] = X[b,k] * w[k]; optionally the zero mode is REPLACED by
//                 zero_scale * zero_mode[b] (PSDBlock_: mean field handled by its own
//                 net and added back = writing V*y_mf into k=0; subtracting the mean
//                 before the transform = dropping X[b,0]).
//   psd_scale_bwd: adjoint of the above incl. the batch reduction for d/dw.
//
// Layout: X, Y are the complex64 half-spectra [B, Kc] (float2, Kc = prod(L[:-1]) * Lh,
// Lh = L[-1]/2+1), w real [Kc].  All three are HBM-bound streams: 16 B per spectrum
// element forward (8 in place), 24 B backward.
#include "nfk_common.cuh"
#include "nfk_psd.cuh"

#define NFK_STREAM(s) reinterpret_cast<cudaStream_t>(s)

namespace nfk {

constexpr int kPsdThreads = 256;
constexpr int kPsdUnroll = 4;

__device__ __forceinline__ double block_sum_f64(double v) {
    __shared__ double part[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) part[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? part[threadIdx.x] : 0.0;
    if (wid == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    return v;
}

// one CTA; Kc is at most a few 10^4 and this runs once per flow evaluation
__global__ void psd_weights_kernel(const float* __restrict__ ipsd, int64_t Kc, int Lh, int inverse,
                                   float* __restrict__ w, float* __restrict__ logj) {
    double acc = 0.0;
    for (int64_t k = threadIdx.x; k < Kc; k += blockDim.x) {
        const float s = ipsd[k];
        w[k] = psd_weight(s, inverse);
        acc += psd_logj_term(k, Lh, s);
    }
    acc = block_sum_f64(acc);
    if (threadIdx.x == 0) logj[0] = (float)((inverse ? 0.5 : -0.5) * acc);
}

__global__ void psd_weights_bwd_kernel(const float* __restrict__ ipsd, const float* __restrict__ w,
                                       const float* __restrict__ gw, const float* __restrict__ glogj,
                                       int64_t Kc, int Lh, int inverse, float* __restrict__ g_ipsd) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Kc) return;
    g_ipsd[k] = psd_weight_grad(k, Lh, ipsd[k], w[k], gw ? gw[k] : 0.f, glogj ? glogj[0] : 0.f, inverse);
}

// grid = (ceil(Kc / threads), sample chunks); a thread owns one k and walks its chunk of
// samples, so w[k] is read once and every load/store is a coalesced 8-byte stream.
__global__ void __launch_bounds__(kPsdThreads)
psd_scale_kernel(const float2* __restrict__ X, const float* __restrict__ w,
                 const float* __restrict__ zero_mode, float zero_scale,
                 float2* __restrict__ Y, int64_t B, int64_t Kc, int64_t per_chunk) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Kc) return;
    const int64_t b0 = (int64_t)blockIdx.y * per_chunk;
    const int64_t b1 = b0 + per_chunk < B ? b0 + per_chunk : B;
    if (k == 0 && zero_mode) {
        for (int64_t b = b0; b < b1; ++b) Y[b * Kc] = make_float2(zero_scale * zero_mode[b], 0.f);
        return;
    }
    const float wk = w[k];
    int64_t b = b0;
    for (; b + kPsdUnroll <= b1; b += kPsdUnroll) {
        float2 v[kPsdUnroll];
#pragma unroll
        for (int u = 0; u < kPsdUnroll; ++u) v[u] = X[(b + u) * Kc + k];
#pragma unroll
        for (int u = 0; u < kPsdUnroll; ++u) Y[(b + u) * Kc + k] = make_float2(v[u].x * wk, v[u].y * wk);
    }
    for (; b < b1; ++b) {
        const float2 v = X[b * Kc + k];
        Y[b * Kc + k] = make_float2(v.x * wk, v.y * wk);
    }
}

// gX = gY * w (0 at a replaced zero mode), g_zero[b] = zero_scale * Re gY[b,0],
// gw_part[chunk,k] = sum_{b in chunk} Re(conj(X) gY)
__global__ void __launch_bounds__(kPsdThreads)
psd_scale_bwd_kernel(const float2* __restrict__ X, const float2* __restrict__ gY,
                     const float* __restrict__ w, int replace_zero, float zero_scale,
                     float2* __restrict__ gX, float* __restrict__ gw_part, float* __restrict__ g_zero,
                     int64_t B, int64_t Kc, int64_t per_chunk) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Kc) return;
    const int64_t b0 = (int64_t)blockIdx.y * per_chunk;
    const int64_t b1 = b0 + per_chunk < B ? b0 + per_chunk : B;
    if (k == 0 && replace_zero) {
        for (int64_t b = b0; b < b1; ++b) {
            if (g_zero) g_zero[b] = zero_scale * gY[b * Kc].x;
            if (gX) gX[b * Kc] = make_float2(0.f, 0.f);
        }
        if (gw_part) gw_part[(int64_t)blockIdx.y * Kc] = 0.f;
        return;
    }
    const float wk = w[k];
    float acc = 0.f;
    int64_t b = b0;
    for (; b + kPsdUnroll <= b1; b += kPsdUnroll) {
        float2 g[kPsdUnroll], x[kPsdUnroll];
#pragma unroll
        for (int u = 0; u < kPsdUnroll; ++u) g[u] = gY[(b + u) * Kc + k];
        if (gw_part) {
#pragma unroll
            for (int u = 0; u < kPsdUnroll; ++u) x[u] = X[(b + u) * Kc + k];
#pragma unroll
            for (int u = 0; u < kPsdUnroll; ++u) acc += x[u].x * g[u].x + x[u].y * g[u].y;
        }
        if (gX) {
#pragma unroll
            for (int u = 0; u < kPsdUnroll; ++u) gX[(b + u) * Kc + k] = make_float2(g[u].x * wk, g[u].y * wk);
        }
    }
    for (; b < b1; ++b) {
        const float2 g = gY[b * Kc + k];
        if (gw_part) {
            const float2 x = X[b * Kc + k];
            acc += x.x * g.x + x.y * g.y;
        }
        if (gX) gX[b * Kc + k] = make_float2(g.x * wk, g.y * wk);
    }
    if (gw_part) gw_part[(int64_t)blockIdx.y * Kc + k] = acc;
}

__global__ void psd_reduce_parts_kernel(const float* __restrict__ part, int chunks, int64_t Kc,
                                        float* __restrict__ gw) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Kc) return;
    float acc = 0.f;
    for (int c = 0; c < chunks; ++c) acc += part[(int64_t)c * Kc + k];
    gw[k] = acc;
}

// ---------------------------------------------------------------- per-sample mean / shift
// MeanFieldNet_ used on a whole field (meanfield_.py:26-32): mean over the lattice, then
// x + delta[b].  T threads co-operate on one sample (T = 32 for small lattices).
template <int T>
__global__ void __launch_bounds__(256)
sample_mean_kernel(const float* __restrict__ x, int64_t B, int64_t V, float inv_v, float* __restrict__ mean) {
    constexpr int kPer = 256 / T;
    const int sub = threadIdx.x / T, lane = threadIdx.x % T;
    const int64_t b = (int64_t)blockIdx.x * kPer + sub;
    float acc = 0.f;
    if (b < B) {
        const float* row = x + b * V;
        for (int64_t i = lane; i < V; i += T) acc += row[i];
    }
    if (T == 32) {
        acc = warp_sum(acc);
    } else {
        acc = block_sum(acc);
    }
    if (lane == 0 && b < B) mean[b] = acc * inv_v;
}

__global__ void sample_shift_kernel(const float* __restrict__ x, const float* __restrict__ delta,
                                    float* __restrict__ y, int64_t n, int64_t V) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        y[i] = x[i] + delta[i / V];
}
__global__ void sample_shift4_kernel(const float4* __restrict__ x, const float* __restrict__ delta,
                                     float4* __restrict__ y, int64_t n4, int64_t V4) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float d = delta[i / V4];
        float4 v = x[i];
        v.x += d; v.y += d; v.z += d; v.w += d;
        y[i] = v;
    }
}

}  // namespace nfk

using namespace nfk;

// Sample chunks of the two batch-walking kernels: enough CTAs for two waves of 148 SMs at
// 8 resident CTAs each, at least kPsdUnroll samples per chunk, at most 65535 chunks.
extern "C" int nfk_psd_chunks(int64_t B, int64_t Kc) {
    if (B <= 0 || Kc <= 0) return 0;
    const int64_t kblocks = (Kc + kPsdThreads - 1) / kPsdThreads;
    int64_t want = (2 * 148 * 8 + kblocks - 1) / kblocks;
    const int64_t max_chunks = (B + kPsdUnroll - 1) / kPsdUnroll;
    if (want > max_chunks) want = max_chunks;
    if (want > 65535) want = 65535;
    if (want < 1) want = 1;
    const int64_t per = (B + want - 1) / want;
    return (int)((B + per - 1) / per);
}

static inline int64_t psd_per_chunk(int64_t B, int chunks) { return (B + chunks - 1) / chunks; }

extern "C" int nfk_psd_weights_fwd(const float* ipsd, int64_t Kc, int last_half, int inverse,
                                   float* w, float* logj, void* stream) {
    if (!ipsd || !w || !logj || Kc < 1 || last_half < 1 || Kc % last_half != 0) return NFK_EINVAL;
    psd_weights_kernel<<<1, 1024, 0, NFK_STREAM(stream)>>>(ipsd, Kc, last_half, inverse, w, logj);
    return check_launch();
}

extern "C" int nfk_psd_weights_bwd(const float* ipsd, const float* w, const float* gw, const float* glogj,
                                   int64_t Kc, int last_half, int inverse, float* g_ipsd, void* stream) {
    if (!ipsd || !w || !g_ipsd || Kc < 1 || last_half < 1 || Kc % last_half != 0) return NFK_EINVAL;
    psd_weights_bwd_kernel<<<(unsigned)((Kc + 255) / 256), 256, 0, NFK_STREAM(stream)>>>(
        ipsd, w, gw, glogj, Kc, last_half, inverse, g_ipsd);
    return check_launch();
}

extern "C" int nfk_psd_scale(const float* X, const float* w, const float* zero_mode, float zero_scale,
                             float* Y, int64_t B, int64_t Kc, void* stream) {
    if (!X || !w || !Y || Kc < 1) return NFK_EINVAL;
    if (((uintptr_t)X | (uintptr_t)Y) % 8 != 0) return NFK_EINVAL;
    if (B <= 0) return NFK_OK;
    const int chunks = nfk_psd_chunks(B, Kc);
    dim3 grid((unsigned)((Kc + kPsdThreads - 1) / kPsdThreads), (unsigned)chunks);
    psd_scale_kernel<<<grid, kPsdThreads, 0, NFK_STREAM(stream)>>>(
        reinterpret_cast<const float2*>(X), w, zero_mode, zero_scale, reinterpret_cast<float2*>(Y), B, Kc,
        psd_per_chunk(B, chunks));
    return check_launch();
}

extern "C" int nfk_psd_scale_bwd(const float* X, const float* gY, const float* w, int replace_zero,
                                 float zero_scale, float* gX, float* gw, float* gw_part, float* g_zero,
                                 int64_t B, int64_t Kc, void* stream) {
    if (!gY || !w || Kc < 1) return NFK_EINVAL;
    if (gw && (!X || !gw_part)) return NFK_EINVAL;
    if (replace_zero && !g_zero) return NFK_EINVAL;
    if (((uintptr_t)X | (uintptr_t)gY | (uintptr_t)gX) % 8 != 0) return NFK_EINVAL;
    if (B <= 0) {
        if (gw) {
            if (cudaMemsetAsync(gw, 0, sizeof(float) * Kc, NFK_STREAM(stream)) != cudaSuccess) return NFK_ECUDA;
        }
        return NFK_OK;
    }
    const int chunks = nfk_psd_chunks(B, Kc);
    dim3 grid((unsigned)((Kc + kPsdThreads - 1) / kPsdThreads), (unsigned)chunks);
    psd_scale_bwd_kernel<<<grid, kPsdThreads, 0, NFK_STREAM(stream)>>>(
        reinterpret_cast<const float2*>(X), reinterpret_cast<const float2*>(gY), w, replace_zero, zero_scale,
        reinterpret_cast<float2*>(gX), gw ? gw_part : nullptr, g_zero, B, Kc, psd_per_chunk(B, chunks));
    int rc = check_launch();
    if (rc != NFK_OK || !gw) return rc;
    psd_reduce_parts_kernel<<<(unsigned)((Kc + 255) / 256), 256, 0, NFK_STREAM(stream)>>>(gw_part, chunks, Kc, gw);
    return check_launch();
}

extern "C" int nfk_sample_mean(const float* x, int64_t B, int64_t V, float scale, float* mean, void* stream) {
    if (!x || !mean || V < 1) return NFK_EINVAL;
    if (B <= 0) return NFK_OK;
    if (V >= 2048) {
        sample_mean_kernel<256><<<(unsigned)B, 256, 0, NFK_STREAM(stream)>>>(x, B, V, scale, mean);
    } else {
        sample_mean_kernel<32><<<(unsigned)((B + 7) / 8), 256, 0, NFK_STREAM(stream)>>>(x, B, V, scale, mean);
    }
    return check_launch();
}

extern "C" int nfk_sample_shift(const float* x, const float* delta, float* y, int64_t B, int64_t V,
                                void* stream) {
    if (!x || !delta || !y || V < 1) return NFK_EINVAL;
    if (B <= 0) return NFK_OK;
    const int64_t n = B * V;
    const bool vec = V % 4 == 0 && ((uintptr_t)x | (uintptr_t)y) % 16 == 0;
    const int64_t items = vec ? n / 4 : n;
    int64_t blocks = (items + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (vec) {
        sample_shift4_kernel<<<(unsigned)blocks, 256, 0, NFK_STREAM(stream)>>>(
            reinterpret_cast<const float4*>(x), delta, reinterpret_cast<float4*>(y), items, V / 4);
    } else {
        sample_shift_kernel<<<(unsigned)blocks, 256, 0, NFK_STREAM(stream)>>>(x, delta, y, n, V);
    }
    return check_launch();
}
