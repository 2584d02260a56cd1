// mn_probe.cu -- throw-away probe for the tensor-core WEIGHT GRADIENT design: MN-major SWIZZLE_NONE fp16 operands read
// straight from site-major 16-byte records (8 consecutive sites x 8 channels = one core matrix):
//   A: M blocks = channel groups (block stride = plane stride), K = sites;   B: N blocks = taps (block stride = ONE
//   record = 16 bytes: overlapping core matrices), K = sites.  Which of LBO / SBO is the MN-block stride, and where do
//   the rows of an M = 64 accumulator live in TMEM?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mn_probe mn_probe.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
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
    d |= (uint64_t)1 << 46;
    return d;
}
__host__ __device__ inline uint32_t make_idesc(int M, int N, int amn, int bmn) {
    uint32_t d = 0;
    d |= 1u << 4;                    // D = f32, A = B = f16
    d |= (uint32_t)amn << 15;
    d |= (uint32_t)bmn << 16;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
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
    for (int it = 0; it < (1 << 18); ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

constexpr int NPOS = 64;          // record positions per plane
constexpr int AG = 8;             // A planes (M blocks), plane stride = NPOS records
constexpr int NCOL = 64;          // accumulator columns read back

// A records: a[AG][NPOS][8]; B records: b[NPOS][8].  variant 0: SBO = MN-block stride, LBO = K-block stride (CuTe's
// canonical form); 1: the two swapped.  M in {64, 128}, N = 8 nb.
__global__ void __launch_bounds__(128) probe(const __half* a, const __half* b, float* d /*[128][NCOL]*/, int variant, int M,
                                             int nb, int kstart, int* err) {
    __shared__ __align__(128) __half As[AG][NPOS][8];
    __shared__ __align__(128) __half Bs[NPOS][8];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < AG * NPOS * 8; i += 128) (&As[0][0][0])[i] = a[i];
    for (int i = tid; i < NPOS * 8; i += 128) (&Bs[0][0])[i] = b[i];
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    // clear the accumulator columns through a dummy read?  (not needed: accumulate = 0 overwrites rows the MMA writes;
    // rows it does not write keep stale data -- the host looks at matches only)
    if (tid == 0) {
        const uint32_t a_mn = NPOS * 16, k_blk = 128, b_mn = 16;
        const uint64_t da = variant == 0 ? make_desc(smem_u32(&As[0][kstart][0]), k_blk, a_mn) : make_desc(smem_u32(&As[0][kstart][0]), a_mn, k_blk);
        const uint64_t db = variant == 0 ? make_desc(smem_u32(&Bs[kstart][0]), k_blk, b_mn) : make_desc(smem_u32(&Bs[kstart][0]), b_mn, k_blk);
        mma_f16(tmem, da, db, make_idesc(M, 8 * nb, 1, 1), 0);
        mma_commit(smem_u32(&bar));
    }
    __syncthreads();
    if (!mbar_wait(smem_u32(&bar), 0)) { if (tid == 0) *err = 1; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < NCOL; c0 += 8) {
        uint32_t r[8];
        tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
        for (int j = 0; j < 8; ++j) d[tid * NCOL + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(64));
}


// cycles per MMA: `reps` dependent (same accumulator) or independent (ncolsets accumulators round robin) MMAs
__global__ void __launch_bounds__(128) probe_time(int M, int N, int amn, int bmn, int b_sbo, int b_off, int nsets, int reps, long long* out) {
    __shared__ __align__(128) __half As[16][NPOS][8];
    __shared__ __align__(128) __half Bs[8][NPOS][8];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 16 * NPOS * 8; i += 128) (&As[0][0][0])[i] = __float2half(0.001f * (i % 97));
    for (int i = tid; i < 8 * NPOS * 8; i += 128) (&Bs[0][0][0])[i] = __float2half(0.001f * (i % 89));
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        const uint32_t plane = NPOS * 16;
        // MN-major: LBO = K-block stride (128 B), SBO = MN-block stride; K-major: LBO = K-group distance, SBO = 128
        const uint64_t da = amn ? make_desc(smem_u32(&As[0][0][0]), 128, plane) : make_desc(smem_u32(&As[0][0][0]), plane, 128);
        const uint64_t db = bmn ? make_desc(smem_u32(&Bs[0][0][0]) + b_off, 128, b_sbo) : make_desc(smem_u32(&Bs[0][0][0]) + b_off, plane, 128);
        const uint32_t idesc = make_idesc(M, N, amn, bmn);
        for (int w = 0; w < 2; ++w) {
            const long long t0 = clock64();
            const uint32_t m1 = nsets > 1 ? 64u : 0u, m2 = nsets > 2 ? 128u : 0u;
#pragma unroll 1
            for (int r = 0; r < reps; r += 4) {
                mma_f16(tmem, da, db, idesc, 1);
                mma_f16(tmem + m1, da, db, idesc, 1);
                mma_f16(tmem + m2, da, db, idesc, 1);
                mma_f16(tmem + m1 + m2, da, db, idesc, 1);
            }
            mma_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), w & 1);
            out[w] = clock64() - t0;
        }
    }
    __syncthreads();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

int main() {
    {
        long long* dout; CK(cudaMalloc(&dout, 16));
        struct V { const char* name; int M, N, amn, bmn, b_sbo, b_off, nsets; };
        const V vs[] = {
            {"K-major  M128 N16  (forward layer 2)", 128, 16, 0, 0, 0, 0, 1},
            {"K-major  M128 N64  (forward layer 3)", 128, 64, 0, 0, 0, 0, 1},
            {"K-major  M64  N24", 64, 24, 0, 0, 0, 0, 1},
            {"MN-major M64  N24  B blocks 16 B apart (wgrad)", 64, 24, 1, 1, 16, 0, 1},
            {"MN-major M64  N24  same, 4 accumulators round robin", 64, 24, 1, 1, 16, 0, 4},
            {"MN-major M64  N24  B blocks a plane apart", 64, 24, 1, 1, NPOS * 16, 0, 1},
            {"MN-major M64  N24  B start + 16 B (unaligned core matrices)", 64, 24, 1, 1, NPOS * 16, 16, 1},
            {"MN-major A only M64 N24", 64, 24, 1, 0, 0, 0, 1},
            {"MN-major B only M64 N24 (planes)", 64, 24, 0, 1, NPOS * 16, 0, 1},
            {"MN-major M128 N24 (planes)", 128, 32, 1, 1, NPOS * 16, 0, 1},
            {"MN-major M128 N48 B 16 B apart", 128, 48, 1, 1, 16, 0, 1},
            {"MN-major M64  N8", 64, 8, 1, 1, 16, 0, 1},
            {"MN-major M64  N48 B 16 B apart", 64, 48, 1, 1, 16, 0, 1},
        };
        for (const V& v : vs) {
            probe_time<<<1, 128>>>(v.M, v.N, v.amn, v.bmn, v.b_sbo, v.b_off, v.nsets, 400, dout);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: CUDA error %s\n", v.name, cudaGetErrorString(e)); return 1; }
            long long h[2]; CK(cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost));
            printf("%-62s %6.1f cycles per MMA\n", v.name, h[1] / 400.0);
        }
    }

    std::vector<__half> ha(AG * NPOS * 8), hb(NPOS * 8);
    std::vector<float> fa(ha.size()), fb(hb.size());
    srand(7);
    for (size_t i = 0; i < ha.size(); ++i) { fa[i] = (rand() % 2001 - 1000) / 512.f; ha[i] = __float2half(fa[i]); fa[i] = __half2float(ha[i]); }
    for (size_t i = 0; i < hb.size(); ++i) { fb[i] = (rand() % 2001 - 1000) / 512.f; hb[i] = __float2half(fb[i]); fb[i] = __half2float(hb[i]); }
    __half *da, *db; float* dd; int* derr;
    CK(cudaMalloc(&da, ha.size() * 2)); CK(cudaMalloc(&db, hb.size() * 2)); CK(cudaMalloc(&dd, 128 * NCOL * 4)); CK(cudaMalloc(&derr, 4));
    CK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
    for (int M : {64}) for (int variant : {0, 1}) for (int kstart : {0, 5}) {
        const int nb = 3;
        CK(cudaMemset(dd, 0xff, 128 * NCOL * 4)); CK(cudaMemset(derr, 0, 4));
        probe<<<1, 128>>>(da, db, dd, variant, M, nb, kstart, derr);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("M %d variant %d: CUDA error %s\n", M, variant, cudaGetErrorString(e)); return 1; }
        int herr; CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
        std::vector<float> hd(128 * NCOL);
        CK(cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost));
        // expected: row m = 8 blk + e  <-  A[blk][kstart + k][e];  column n = 8 j + e  <-  B[kstart + k + j][e]
        const int rows = M;
        std::vector<double> ex(rows * 8 * nb);
        for (int m = 0; m < rows; ++m) for (int n = 0; n < 8 * nb; ++n) {
            double s = 0;
            const int blk = m / 8;
            for (int k = 0; k < 16; ++k)
                s += (blk < AG ? (double)fa[((blk * NPOS) + kstart + k) * 8 + m % 8] : 0.0) * fb[(kstart + k + n / 8) * 8 + n % 8];
            ex[m * 8 * nb + n] = s;
        }
        int matched = 0; char map[1024]; int pos = 0;
        int lane_of_row[128];
        for (int m = 0; m < rows; ++m) {
            lane_of_row[m] = -1;
            for (int l = 0; l < 128; ++l) {
                bool ok = true;
                for (int n = 0; n < 8 * nb && ok; ++n) ok = fabs(hd[l * NCOL + n] - ex[m * 8 * nb + n]) < 1e-2;
                if (ok) { lane_of_row[m] = l; break; }
            }
            matched += lane_of_row[m] >= 0;
        }
        for (int m = 0; m < rows && pos < 900; m += 8) pos += snprintf(map + pos, 1024 - pos, "%d->%d ", m, lane_of_row[m]);
        printf("M %3d variant %d kstart %d timeout %d: %d of %d rows found;  row->lane: %s\n", M, variant, kstart, herr, matched, rows > AG * 8 ? AG * 8 : rows, map);
    }
    return 0;
}
