// nfk_common.cuh -- launch plumbing shared by the .cu files.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/normflow_b200.h"
#include "nfk_ops.cuh"

namespace nfk {

extern std::atomic<unsigned long long> g_launches;   // counted on the host at every launch

inline int check_launch() {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError() == cudaSuccess ? NFK_OK : NFK_ECUDA;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel: raise it once per
// (kernel, device), whichever thread gets there first.  `Kernel` is the kernel itself (a non-type
// template parameter), so every instantiation has its own flags.
template <auto Kernel>
inline int ensure_dynamic_smem(int bytes) {
    static std::atomic<unsigned long long> done[2] = {{0ull}, {0ull}};     // devices 0..127
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return NFK_ECUDA;
    const bool tracked = dev >= 0 && dev < 128;
    if (tracked && (done[dev >> 6].load(std::memory_order_acquire) >> (dev & 63) & 1ull)) return NFK_OK;
    if (cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) {
        cudaGetLastError();
        return NFK_ECUDA;
    }
    if (tracked) done[dev >> 6].fetch_or(1ull << (dev & 63), std::memory_order_release);
    return NFK_OK;
}

inline Lat to_lat(const nfk_lattice& l) { return make_lat(l.ndim, l.shape); }
inline bool lat_ok(const nfk_lattice& l) {
    if (l.ndim < 1 || l.ndim > NFK_MAX_DIM) return false;
    int64_t v = 1;
    for (int d = 0; d < l.ndim; ++d) {
        if (l.shape[d] < 1) return false;
        v *= l.shape[d];
    }
    return v < (int64_t(1) << 31);
}
inline int64_t lat_volume(const nfk_lattice& l) {
    int64_t v = 1;
    for (int d = 0; d < l.ndim; ++d) v *= l.shape[d];
    return v;
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// sum over the CTA (blockDim.x multiple of 32, <= 1024); result valid in thread 0
__device__ __forceinline__ float block_sum(float v) {
    __shared__ float part[32];
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) part[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? part[threadIdx.x] : 0.f;
    if (wid == 0) v = warp_sum(v);
    return v;
}

// ---------------------------------------------------------------- per-sample launch plan
// One CTA per (sample, chunk of sites); a thread walks its chunk with stride
// blockDim.  V <= kSmallV: one thread per sample instead (launch-bound regime).
constexpr int kSmallV = 32;
struct Plan {
    int threads;
    int chunks;          // CTAs per sample
    int64_t chunk_len;   // sites per CTA (multiple of threads*vec)
    bool small;
};
inline Plan make_plan(int64_t V, int vec) {
    Plan p;
    p.small = V <= kSmallV;
    if (p.small) {
        p.threads = 128;
        p.chunks = 1;
        p.chunk_len = V;
        return p;
    }
    const int64_t lanes = (V + vec - 1) / vec;
    int th = 256;
    if (lanes < 256) th = (int)((lanes + 31) / 32 * 32);
    p.threads = th;
    const int64_t per_cta = (int64_t)th * vec * 16;   // <= 16 trips per thread
    p.chunks = (int)((V + per_cta - 1) / per_cta);
    p.chunk_len = per_cta;
    return p;
}

// writes log_out[b] = log_in[b] (or 0): initialises the atomic accumulation when a
// sample is spread over several CTAs
static __global__ void init_log_kernel(const float* log_in, float* log_out, int64_t B) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) log_out[b] = log_in ? log_in[b] : 0.f;
}

template <class Op>
__global__ void __launch_bounds__(256) site_kernel(Op op, int64_t V, int chunks, int64_t chunk_len,
                                                   const float* log_in, float* log_out) {
    const int64_t b = blockIdx.x / chunks;
    const int chunk = (int)(blockIdx.x % chunks);
    const int64_t s0 = chunk * chunk_len;
    const int64_t s1 = s0 + chunk_len < V ? s0 + chunk_len : V;
    float acc = 0.f;
    for (int64_t s = s0 + threadIdx.x; s < s1; s += blockDim.x) acc += op(b, s);
    if (log_out == nullptr) return;            // uniform: no reduction wanted
    acc = block_sum(acc);
    if (threadIdx.x == 0) {
        if (chunks == 1) log_out[b] = (log_in ? log_in[b] : 0.f) + acc;
        else atomicAdd(log_out + b, acc);
    }
}

// vector variant: op(b, s, n) handles sites [s, s+n), n <= VEC
template <class Op, int VEC>
__global__ void __launch_bounds__(256) site_kernel_vec(Op op, int64_t V, int chunks, int64_t chunk_len,
                                                       const float* log_in, float* log_out) {
    const int64_t b = blockIdx.x / chunks;
    const int chunk = (int)(blockIdx.x % chunks);
    const int64_t s0 = chunk * chunk_len;
    const int64_t s1 = s0 + chunk_len < V ? s0 + chunk_len : V;
    float acc = 0.f;
#pragma unroll 4
    for (int64_t s = s0 + (int64_t)threadIdx.x * VEC; s < s1; s += (int64_t)blockDim.x * VEC) {
        const int64_t left = s1 - s;
        acc += op(b, s, left < VEC ? (int)left : VEC);
    }
    if (log_out == nullptr) return;
    acc = block_sum(acc);
    if (threadIdx.x == 0) {
        if (chunks == 1) log_out[b] = (log_in ? log_in[b] : 0.f) + acc;
        else atomicAdd(log_out + b, acc);
    }
}

// tiny lattices: one thread per sample
template <class Op>
__global__ void __launch_bounds__(128) sample_kernel(Op op, int64_t B, int64_t V,
                                                     const float* log_in, float* log_out) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float acc = 0.f;
    for (int64_t s = 0; s < V; ++s) acc += op(b, s);
    if (log_out) log_out[b] = (log_in ? log_in[b] : 0.f) + acc;
}
template <class Op, int VEC>
__global__ void __launch_bounds__(128) sample_kernel_vec(Op op, int64_t B, int64_t V,
                                                         const float* log_in, float* log_out) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float acc = 0.f;
    for (int64_t s = 0; s < V; s += VEC) acc += op(b, s, V - s < VEC ? (int)(V - s) : VEC);
    if (log_out) log_out[b] = (log_in ? log_in[b] : 0.f) + acc;
}

// Runs `op` over every (sample, site) and reduces its return values per sample
// into log_out (may be null: pure map).
template <class Op>
int launch_sites(const Op& op, int64_t B, int64_t V, const float* log_in, float* log_out,
                 cudaStream_t st) {
    if (B <= 0 || V <= 0) return NFK_OK;
    const Plan p = make_plan(V, 1);
    if (p.small) {
        sample_kernel<Op><<<(unsigned)((B + 127) / 128), 128, 0, st>>>(op, B, V, log_in, log_out);
        return check_launch();
    }
    if (p.chunks > 1 && log_out) {
        init_log_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(log_in, log_out, B);
        if (int e = check_launch()) return e;
    }
    site_kernel<Op><<<(unsigned)(B * p.chunks), p.threads, 0, st>>>(op, V, p.chunks, p.chunk_len,
                                                                    log_in, log_out);
    return check_launch();
}

template <class Op, int VEC>
int launch_sites_vec(const Op& op, int64_t B, int64_t V, const float* log_in, float* log_out,
                     cudaStream_t st) {
    if (B <= 0 || V <= 0) return NFK_OK;
    const Plan p = make_plan(V, VEC);
    if (p.small) {
        sample_kernel_vec<Op, VEC><<<(unsigned)((B + 127) / 128), 128, 0, st>>>(op, B, V, log_in, log_out);
        return check_launch();
    }
    if (p.chunks > 1 && log_out) {
        init_log_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(log_in, log_out, B);
        if (int e = check_launch()) return e;
    }
    site_kernel_vec<Op, VEC><<<(unsigned)(B * p.chunks), p.threads, 0, st>>>(op, V, p.chunks, p.chunk_len,
                                                                             log_in, log_out);
    return check_launch();
}

}  // namespace nfk
