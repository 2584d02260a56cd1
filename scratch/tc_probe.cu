// tc_probe.cu -- throw-away probe of tcgen05 on B200 for the fused conditioner design:
//   1. is a K-major SWIZZLE_NONE A operand with an arbitrary 16-byte-aligned (site-shifted)
//      start address read the way the implicit-GEMM convolution needs?
//   2. accuracy of  tf32(A_t x B_t) + bf16([A_bf | A_r] x [B_r ; B_bf])  against fp64
//   3. cycles per MMA for M=128, N in {16,32,64}, kind tf32 (K=8) and f16 (K=16)
//   4. tcgen05.ld throughput
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tc_probe.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // version 1 (Blackwell)
    return d;                        // base_offset 0, lbo_mode 0, SWIZZLE_NONE
}
// kind: 2 = tf32, 1 = bf16
__host__ __device__ inline uint32_t make_idesc(int fmt, int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                    // D = f32
    d |= (uint32_t)fmt << 7;         // A format
    d |= (uint32_t)fmt << 10;        // B format
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;                        // K-major A and B, dense, no negate
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    for (int it = 0; it < (1 << 16); ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;                    // timed out: never hang the box
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

constexpr int WS = 67;             // odd padded row stride (sites)
constexpr int NSITE = 128 + 2 * WS + 8;   // rows of A available: tile + one row of halo each side
constexpr int NB = 32;             // output channels (N)
constexpr int S0 = WS + 1;         // first site of the output tile

// shared layout (bytes)
//   At  : 2 planes [NSITE][4 f32]          (tf32 operand, channel halves 0-3 / 4-7)
//   Abf : [NSITE][8 bf16]  bf16(a)
//   Ar  : [NSITE][8 bf16]  bf16(a - a_t)
//   Bt  : 9 taps x 2 planes [NB][4 f32]
//   Bc  : 9 taps x 2 planes [NB][8 bf16]   plane 0 = bf16(b - b_t), plane 1 = bf16(b)
struct Smem {
    float At[2][NSITE][4];
    __nv_bfloat16 Abf[NSITE][8];
    __nv_bfloat16 Ar[NSITE][8];
    float Bt[9][2][NB][4];
    __nv_bfloat16 Bc[9][2][NB][8];
    uint64_t bar[4];
    uint32_t tmem_base;
};

__device__ __forceinline__ float tf32_rn(float v) {
    uint32_t u = __float_as_uint(v);
    u = (u + 0x1000u) & 0xFFFFE000u;
    return __uint_as_float(u);
}

// mode 0: tf32 only with ROUNDED operands stored; 1: tf32 with RAW fp32 operands stored (what does the
// hardware do with the low 13 bits?); 2: tf32 rounded + bf16 corrections
__global__ void __launch_bounds__(128) probe_correct(const float* a /*[NSITE][8]*/, const float* b /*[9][NB][8]*/,
                                                     float* d /*[128][NB]*/, int mode, int* err) {
    extern __shared__ __align__(1024) uint8_t raw[];
    Smem& s = *reinterpret_cast<Smem*>(raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < NSITE * 8; i += 128) {
        const int site = i / 8, c = i % 8;
        const float v = a[i], vt = tf32_rn(v);
        s.At[c / 4][site][c % 4] = (mode == 1) ? v : vt;
        s.Abf[site][c] = __float2bfloat16_rn(v);
        s.Ar[site][c] = __float2bfloat16_rn(v - vt);
    }
    for (int i = tid; i < 9 * NB * 8; i += 128) {
        const int t = i / (NB * 8), n = (i / 8) % NB, c = i % 8;
        const float v = b[i], vt = tf32_rn(v);
        s.Bt[t][c / 4][n][c % 4] = (mode == 1) ? v : vt;
        s.Bc[t][0][n][c] = __float2bfloat16_rn(v - vt);
        s.Bc[t][1][n][c] = __float2bfloat16_rn(v);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s.tmem_base)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) mbar_init(smem_u32(&s.bar[0]), 1);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    if (tid == 0) {
        const uint32_t id_t = make_idesc(2, 128, NB), id_b = make_idesc(1, 128, NB);
        uint32_t acc = 0;
        for (int t = 0; t < 9; ++t) {
            const int shift = (t / 3 - 1) * WS + (t % 3 - 1);
            const uint64_t da = make_desc(smem_u32(&s.At[0][S0 + shift][0]), sizeof(s.At[0]), 128);
            const uint64_t db = make_desc(smem_u32(&s.Bt[t][0][0][0]), sizeof(s.Bt[0][0]), 128);
            mma_tf32(tmem, da, db, id_t, acc);
            acc = 1;
            if (mode == 2) {
                const uint64_t ca = make_desc(smem_u32(&s.Abf[S0 + shift][0]),
                                              (uint32_t)((const uint8_t*)&s.Ar[0][0] - (const uint8_t*)&s.Abf[0][0]), 128);
                const uint64_t cb = make_desc(smem_u32(&s.Bc[t][0][0][0]), sizeof(s.Bc[0][0]), 128);
                mma_bf16(tmem, ca, cb, id_b, 1);
            }
        }
        mma_commit(smem_u32(&s.bar[0]));
    }
    __syncwarp();
    if (!mbar_wait(smem_u32(&s.bar[0]), 0)) { if (tid == 0) *err = 1; }
    tc_fence_after();
    uint32_t r[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), r);
    tmem_ld_wait();
    for (int j = 0; j < NB; ++j) d[tid * NB + j] = __uint_as_float(r[j]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(64));
}

// ---------------------------------------------------------------- timing
// issues `reps` x 9 MMAs of the given kind / N and times issue -> completion
//   VAR 0: thread 0, taps chained on one accumulator      VAR 1: thread 0, 4 accumulators interleaved
//   VAR 2: warp 0 + elect.sync, chained                   VAR 3: warp 0 + elect.sync, 4 interleaved
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred;
}
template <int KIND, int VAR>
__global__ void __launch_bounds__(128) probe_time(int N, int reps, long long* out, int* err) {
    extern __shared__ __align__(1024) uint8_t raw[];
    uint32_t* z = reinterpret_cast<uint32_t*>(raw);
    for (int i = threadIdx.x; i < (96 * 1024) / 4; i += 128) z[i] = 0;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) mbar_init(smem_u32(&bar), 1);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    const bool issuer = (VAR < 2) ? (tid == 0) : (warp == 0);
    if (issuer) {
        const uint32_t idesc = make_idesc(KIND, 128, N);
        const uint32_t abase = smem_u32(raw), bbase = abase + 64 * 1024;
        const uint64_t da0 = make_desc(abase, 24 * 1024, 128), db0 = make_desc(bbase, 1024, 128);
        const bool lead = (VAR < 2) ? true : (elect_one() != 0);
        long long t0 = clock64();
        if (VAR == 0 || VAR == 2) {
            for (int r = 0; r < reps; ++r) {
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const uint64_t da = da0 + (uint64_t)((t / 3) * WS + (t % 3));
                    const uint64_t db = db0 + (uint64_t)(t * 128);
                    if (lead) {
                        if (KIND == 2) mma_tf32(tmem, da, db, idesc, t > 0);
                        else mma_bf16(tmem, da, db, idesc, t > 0);
                    }
                }
            }
        } else {
            for (int r = 0; r < reps / 4; ++r) {
#pragma unroll
                for (int t = 0; t < 9; ++t) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint64_t da = da0 + (uint64_t)((t / 3) * WS + (t % 3) + q * 128);
                        const uint64_t db = db0 + (uint64_t)(t * 128);
                        if (lead) {
                            if (KIND == 2) mma_tf32(tmem + q * 128, da, db, idesc, t > 0);
                            else mma_bf16(tmem + q * 128, da, db, idesc, t > 0);
                        }
                    }
                }
            }
        }
        long long t1 = clock64();
        if (lead) mma_commit(smem_u32(&bar));
        if (VAR >= 2) __syncwarp();
        bool ok = mbar_wait(smem_u32(&bar), 0);
        long long t2 = clock64();
        if (!ok) *err = 2;
        if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

// tcgen05.ld throughput: `nwarps` warps each read `reps` x 32 columns of their subpartition
__global__ void __launch_bounds__(512) probe_ld(int reps, int wide, long long* out, float* sink) {
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (wide) {
            uint32_t v[32];
            tmem_ld32(tmem + (uint32_t)((r * 32) & 511), v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc += __uint_as_float(v[j]);
        } else {
            uint32_t v[16];
            tmem_ld16(tmem + (uint32_t)((r * 16) & 511), v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) acc += __uint_as_float(v[j]);
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (tid == 0) out[0] = t1 - t0;
    sink[tid] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512));
}

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    int *derr; long long* dout; float* sink;
    CK(cudaMalloc(&derr, 4)); CK(cudaMemset(derr, 0, 4));
    CK(cudaMalloc(&dout, 64)); CK(cudaMalloc(&sink, 4096));
    // ---- correctness
    std::vector<float> a(NSITE * 8), b(9 * NB * 8);
    srand(1);
    for (auto& v : a) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : b) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.3f;
    float *da, *db, *dd;
    CK(cudaMalloc(&da, a.size() * 4)); CK(cudaMalloc(&db, b.size() * 4)); CK(cudaMalloc(&dd, 128 * NB * 4));
    CK(cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(probe_correct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem) + 1024));
    std::vector<double> ref(128 * NB), mag(128 * NB);
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < NB; ++n) {
            double acc = 0, ab = 0;
            for (int t = 0; t < 9; ++t) {
                const int site = S0 + m + (t / 3 - 1) * WS + (t % 3 - 1);
                for (int c = 0; c < 8; ++c) {
                    const double p = (double)a[site * 8 + c] * (double)b[(t * NB + n) * 8 + c];
                    acc += p; ab += fabs(p);
                }
            }
            ref[m * NB + n] = acc; mag[m * NB + n] = ab;
        }
    for (int mode = 0; mode < 3; ++mode) {
        CK(cudaMemset(dd, 0, 128 * NB * 4));
        probe_correct<<<1, 128, sizeof(Smem) + 1024>>>(da, db, dd, mode, derr);
        CK(cudaDeviceSynchronize());
        std::vector<float> d(128 * NB);
        CK(cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost));
        double worst = 0, worst_rel_mag = 0;
        for (size_t i = 0; i < d.size(); ++i) {
            worst = fmax(worst, fabs(d[i] - ref[i]));
            worst_rel_mag = fmax(worst_rel_mag, fabs(d[i] - ref[i]) / mag[i]);
        }
        printf("correct mode %d: max abs err %.3e   max err / sum|terms| %.3e   (d[0]=%f ref[0]=%f)\n", mode, worst,
               worst_rel_mag, d[0], ref[0]);
    }
    int herr = 0;
    CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
    printf("err flag after correctness: %d\n", herr);

    // ---- timing
    auto run_time = [&](int kind, int var, int N) {
        for (int rep = 0; rep < 2; ++rep) {
#define LAUNCH(K, V) { CK(cudaFuncSetAttribute(probe_time<K, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); \
                       probe_time<K, V><<<1, 128, 96 * 1024>>>(N, 20, dout, derr); }
            if (kind == 2) { if (var == 0) LAUNCH(2, 0) else if (var == 1) LAUNCH(2, 1) else if (var == 2) LAUNCH(2, 2) else LAUNCH(2, 3) }
            else { if (var == 0) LAUNCH(1, 0) else if (var == 1) LAUNCH(1, 1) else if (var == 2) LAUNCH(1, 2) else LAUNCH(1, 3) }
            CK(cudaDeviceSynchronize());
        }
        long long h[2];
        CK(cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost));
        int e = 0;
        CK(cudaMemcpy(&e, derr, 4, cudaMemcpyDeviceToHost));
        CK(cudaMemset(derr, 0, 4));
        printf("time kind=%s var=%d N=%3d: 180 MMAs issue %lld cyc, done %lld cyc -> %.1f cyc/MMA  err=%d\n",
               kind == 2 ? "tf32" : "bf16", var, N, h[0], h[1], h[1] / 180.0, e);
    };
    const int Ns[] = {16, 32, 64, 128, 256};
    for (int var = 0; var < 4; ++var)
        for (int kind = 2; kind >= 1; --kind)
            for (int N : Ns) {
                if (N == 256 && (var & 1)) continue;     // 4 x 256 columns do not fit the spread layout
                run_time(kind, var, N);
            }
    CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
    printf("err flag after timing: %d\n", herr);

    // ---- tcgen05.ld
    for (int wide = 0; wide < 2; ++wide)
        for (int threads : {128, 256, 512}) {
            for (int rep = 0; rep < 2; ++rep) {
                probe_ld<<<1, threads>>>(256, wide, dout, sink);
                CK(cudaDeviceSynchronize());
            }
            long long h;
            CK(cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost));
            const double bytes = 256.0 * (wide ? 32 : 16) * 4 * threads;
            printf("tcgen05.ld x%d, %d threads: %lld cyc for %.0f bytes -> %.1f B/cyc\n", wide ? 32 : 16, threads, h, bytes,
                   bytes / h);
        }
    printf("probe done\n");
    return 0;
}
