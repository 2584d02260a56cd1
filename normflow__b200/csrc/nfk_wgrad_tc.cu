// nfk_wgrad_tc.cu -- weight gradient of a 2-D 3x3 circular convolution with 8 input channels on the
// tensor cores (tcgen05, accumulators in TMEM).  Adjoint of ConvAct's layers (reference
// src/nn/scalar/modules.py:131-145) with respect to their weights and biases:
//
//     gw[co][ci][kh][kw] = sum_{b, r, c} gpre[b][co][r][c] * in[b][ci][r + kh - 1][c + kw - 1]     (periodic)
//     gbias[co]          = sum_{b, r, c} gpre[b][co][r][c]
//
// As a GEMM the reduction index is the SITE: D[m][n] += sum_k A[m][k] B[n][k] with k = sites of a tile,
// m = (tap, ci) (72 rows) plus one row of ones (m = 72, which makes D[72][n] the bias gradient), n = co.
// The CUDA-core kernel (nfk_conv.cu) does this with one warp per tap and 56 register accumulators per
// lane at ~10 % of the fp32 peak; here 16 producer warps stream the raw rows of the next tile into shared
// memory (cp.async) while they lay the current tile out as im2col operands, and one thread issues the MMAs.
//
//   * Precision.  Both operands vary per site and gradients span many orders of magnitude (those of a
//     batch-mean loss are ~1/B), so the operands are TF32 pairs, which keep float32's exponent range: v =
//     hi + lo with hi = the top 11 significant bits of v (exact in tf32) and lo = v - hi (exact in float32,
//     read by the tensor core to 11 bits).  D += A_hi B_hi + A_hi B_lo + A_lo B_hi, the lo*lo term (2^-22
//     relative) is dropped -- no scaling pass, no range assumptions.  The fp32 accumulator in TMEM is
//     drained into registers every few tiles (plain round-to-nearest additions) so that no long
//     accumulation chain builds up inside the tensor core.
//   * Shared-memory operand layout (K-major, SWIZZLE_NONE canonical, see nfk_tc.cuh): one k-group of 4
//     sites is 128 rows x 16 bytes; rows 0..72 are A (x patches + ones), rows 96..127 are B (gpre), so ONE
//     buffer serves both descriptors (SBO = 128, LBO = 2048); an MMA (K = 8) reads two k-groups.  Rows
//     that are never written only feed accumulator rows / columns that are never read.
//   * SPARSE (g_parity >= 0): gpre comes from a checkerboard coupling and vanishes off one partition;
//     only those sites enter the tile (k runs over the active columns of a row).
//   * A tile is R consecutive rows of one sample (R * ncols <= 64 sites, ncols = active columns per row, a
//     multiple of 8; rows past the lattice are zero-filled); tiles are dealt round-robin to one persistent
//     CTA per SM.
//
// Needs: Ci == 8, Co <= 32, 16-byte aligned tensors, L1 % 8 == 0 (SPARSE: L1 % 16 == 0 and even L0) and at
// most 64 (active) columns per row.  Otherwise NFK_EUNSUPPORTED (the caller falls back to the
// CUDA-core kernels).
#include <stdlib.h>

#include "nfk_common.cuh"
#include "nfk_tc.cuh"

#define NFK_STREAM(s) reinterpret_cast<cudaStream_t>(s)

namespace {

using namespace nfk;

constexpr int kProducerWarps = 16;
constexpr int kProducerThreads = kProducerWarps * 32;       // 512
constexpr int kDrainWarps = 4;                              // producer warps 0..3 also drain TMEM (lane quarter = warp)
constexpr int kMmaWarp = kProducerWarps;                    // 16
constexpr int kThreads = (kMmaWarp + 1) * 32;               // 544
constexpr int kRowsPerGroup = 128;                          // 16-byte records per k-group
constexpr int kBRow0 = 96;                                  // first B (gpre) row
constexpr int kOnesRow = 72;
constexpr int kGroupBytes = kRowsPerGroup * 16;             // LBO
constexpr int kMaxKT = 64;                                  // sites per tile
constexpr int kSlots = 2;                                   // im2col tasks (8 sites of one operand row) per producer thread
constexpr int kCpSlots = 4;                                 // cp.async chunks per producer thread and tile
constexpr int kTmemCols = 32;                               // D: 128 lanes x 32 columns
constexpr int kRawStages = 3;                                // raw-row stages in flight (cp.async)
constexpr int kBarFull0 = 1, kBarDrained = 3, kBarProducers = 4;   // named barriers (0 is __syncthreads)

__device__ long long g_wg_trace[2048];      // debug: phase time stamps of CTA 0 (NFK_WGRAD_TRACE=1)

struct WgTcArgs {
    int trace;
    const float* in;        // [B][8][L0][L1]
    const float* g;         // [B][Co][L0][L1]
    float* gw;              // [Co][8][3][3], accumulated into
    float* gbias;           // [Co] or null
    int L0, L1, Co;
    long long B;
    int g_parity;           // -1: dense
    int R, ncols, ncols8, KT;
    int raw_bytes;          // one raw stage: in rows r0-1 .. r0+R (8 channels) + gpre rows r0 .. r0+R-1 (Co channels)
    int drain_every;        // tiles between accumulator drains
    int row_stride;         // floats per staged row: [4 wrapped | L1 | 4 wrapped]
    int xch_stride, gch_stride;   // floats per staged channel, = 4 mod 32 so that the eight channels a quarter-warp
                                  // reads at the same (row, column) fall into different bank groups
};

// eight values -> two 16-byte records (sites 0..3 and 4..7) of their hi parts and two of their lo parts;
// consecutive k-groups are kGroupBytes apart
__device__ __forceinline__ void split_store(const float (&v)[8], unsigned char* hi_rec, unsigned char* lo_rec) {
    float h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        h[i] = __uint_as_float(__float_as_uint(v[i]) & 0xFFFFE000u);
        l[i] = v[i] - h[i];
    }
    *reinterpret_cast<float4*>(hi_rec) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(hi_rec + kGroupBytes) = make_float4(h[4], h[5], h[6], h[7]);
    *reinterpret_cast<float4*>(lo_rec) = make_float4(l[0], l[1], l[2], l[3]);
    *reinterpret_cast<float4*>(lo_rec + kGroupBytes) = make_float4(l[4], l[5], l[6], l[7]);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool live) {
    const int sz = live ? 16 : 0;                    // src-size 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(sz) : "memory");
}

// N aligned groups of four from a staged row (the row carries one wrapped group on either side)
template <int N>
__device__ __forceinline__ void load_row(const float* p, float (&v)[4 * N]) {
#pragma unroll
    for (int q = 0; q < N; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(p + 4 * q);
        v[4 * q + 0] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
}

template <bool SPARSE>
__global__ void __launch_bounds__(kThreads, 1) wgrad2d_tc_kernel(WgTcArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bars[3];            // empty[0], empty[1], accumulators complete
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L0 = a.L0, L1 = a.L1, Co = a.Co, R = a.R, KT = a.KT;
    const int ngroups = KT >> 3, gpr = a.ncols8 >> 3;                 // 8-site groups per tile / per row
    const int part_bytes = 2 * ngroups * kGroupBytes;                 // one precision part of one buffer (k-groups of 4)
    const int strips = (L0 + R - 1) / R;
    const long long total = a.B * strips;
    const long long n_tiles = total > blockIdx.x ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    // operand buffer s: [hi part | lo part]
    auto part = [&](int s, int lo) { return smem + (size_t)(2 * s + lo) * part_bytes; };
    // drain points: every `drain_every` tiles and after the last one (a running counter, no division)
    struct DrainClock {
        int every, since;
        __device__ bool tick(bool last) {                                 // call once per tile, in order
            if (++since == every || last) { since = 0; return true; }
            return false;
        }
    };

    if (tid == 0) {
        for (int i = 0; i < 3; ++i) tc::mbar_init(tc::smem_u32(bars + i), 1);
        tc::fence_mbar_init();
    }
    if (warp == kMmaWarp) tc::tmem_alloc(tc::smem_u32(&tmem_slot), kTmemCols);
    // the row of ones (bias gradient): hi = 1, lo = 0, every k-group of both buffers
    for (int e = tid; e < 4 * ngroups; e += kThreads) {
        const int s = e / (2 * ngroups), kg = e - s * 2 * ngroups;
        *reinterpret_cast<float4*>(part(s, 0) + kg * kGroupBytes + kOnesRow * 16) = make_float4(1.f, 1.f, 1.f, 1.f);
        *reinterpret_cast<float4*>(part(s, 1) + kg * kGroupBytes + kOnesRow * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;

    if (warp < kProducerWarps) {
        // =============================== producers ============================================
        // raw fp32 rows of the NEXT tile stream in with cp.async while the current tile's rows (already in
        // shared memory) are split into fp16 pairs and laid out as MMA operands.  A thread's share of the
        // work is the same for every tile, so it is decoded once.
        const int a_tasks = 8 * 9 * ngroups, g_tasks = Co * ngroups;      // one (channel, tap) / one gpre channel x 8 sites
        const int x_lines = 8 * (R + 2), g_lines = Co * R;
        const int RS = a.row_stride;
        const int gr_off = 8 * a.xch_stride;                              // gpre rows follow the input rows (floats)
        unsigned char* raw0 = smem + (size_t)4 * part_bytes;

        // ---- im2col task slots: kind 0 = input patch row of one tap, 1 = gpre row, -1 = none
        int kind[kSlots], src[kSlots], dst[kSlots], cc0s[kSlots], jrow[kSlots], kws[kSlots];
#pragma unroll
        for (int k = 0; k < kSlots; ++k) {
            const int t = tid + k * kProducerThreads;
            kind[k] = -1; src[k] = dst[k] = cc0s[k] = jrow[k] = kws[k] = 0;
            if (t < a_tasks) {
                const int ci = t & 7, rest = t >> 3;
                const int kg = rest % ngroups, tap = rest / ngroups;
                const int j = kg / gpr;
                kind[k] = 0;
                cc0s[k] = (kg - j * gpr) * 8;
                jrow[k] = j;
                kws[k] = tap % 3;
                src[k] = ci * a.xch_stride + (j + tap / 3) * RS + 4;         // column 0 of the staged row
                dst[k] = 2 * kg * kGroupBytes + (tap * 8 + ci) * 16;
            } else if (t < a_tasks + g_tasks) {
                const int tg = t - a_tasks;
                const int co = tg % Co, kg = tg / Co;
                const int j = kg / gpr;
                kind[k] = 1;
                cc0s[k] = (kg - j * gpr) * 8;
                jrow[k] = j;
                src[k] = gr_off + co * a.gch_stride + j * RS + 4;
                dst[k] = 2 * kg * kGroupBytes + (kBRow0 + co) * 16;
            }
        }

        // ---- cp.async slots: the 16-byte chunks of every staged line, spread evenly over the producers.
        // Chunk 0 of a line is the wrapped group in front (columns L1-4 .. L1-1), the last one the wrapped
        // group behind (columns 0 .. 3).
        const int cpl = (L1 >> 2) + 2;                                    // chunks per line
        const int n_chunks = (x_lines + g_lines) * cpl;
        int cp_dst[kCpSlots], cp_ch[kCpSlots], cp_row[kCpSlots], cp_col[kCpSlots];      // cp_ch < 0: none; >= 8: gpre channel + 8
#pragma unroll
        // full rounds go to every thread; the remainder goes to the HIGHEST thread ids, which are the ones
        // with fewer im2col tasks (those are dealt from thread 0 upwards)
        const int full_rounds = n_chunks / kProducerThreads, rem_chunks = n_chunks - full_rounds * kProducerThreads;
        for (int k = 0; k < kCpSlots; ++k) {
            int c = -1;
            if (k < full_rounds) c = tid + k * kProducerThreads;
            else if (k == full_rounds && kProducerThreads - 1 - tid < rem_chunks)
                c = full_rounds * kProducerThreads + (kProducerThreads - 1 - tid);
            cp_dst[k] = cp_row[k] = cp_col[k] = 0;
            cp_ch[k] = -1;
            if (c >= 0) {
                const int line = c / cpl, q = c - line * cpl;
                cp_col[k] = q == 0 ? L1 - 4 : (q == cpl - 1 ? 0 : 4 * (q - 1));
                if (line < x_lines) {
                    const int ci = line / (R + 2), jj = line - ci * (R + 2);
                    cp_ch[k] = ci;
                    cp_row[k] = jj - 1;                                   // lattice row = r0 + cp_row (wrapped)
                    cp_dst[k] = (ci * a.xch_stride + jj * RS + 4 * q) * 4;
                } else {
                    const int lg = line - x_lines;
                    const int co = lg / R, j = lg - co * R;
                    cp_ch[k] = 8 + co;
                    cp_row[k] = j;
                    cp_dst[k] = (gr_off + co * a.gch_stride + j * RS + 4 * q) * 4;
                }
            }
        }

        // a tile cursor: (sample, strip) of tile i, advanced without divisions
        struct Cursor { long long b; int strip; };
        const int step_b = (int)(gridDim.x / strips), step_s = (int)(gridDim.x % strips);
        auto advance = [&](Cursor& c) {
            c.b += step_b;
            c.strip += step_s;
            if (c.strip >= strips) { c.strip -= strips; ++c.b; }
        };
        auto issue = [&](long long i, const Cursor& cur, int stg) {
            if (i < n_tiles) {
                const int r0 = cur.strip * R;
                const uint32_t raw = tc::smem_u32(raw0 + (size_t)stg * a.raw_bytes);
                const float* in_b = a.in + cur.b * 8 * (long long)L0 * L1;
                const float* g_b = a.g + cur.b * Co * (long long)L0 * L1;
#pragma unroll
                for (int k = 0; k < kCpSlots; ++k) {
                    if (cp_ch[k] < 0) continue;
                    int rr = r0 + cp_row[k];
                    if (cp_ch[k] < 8) {
                        rr = rr < 0 ? rr + L0 : rr;
                        while (rr >= L0) rr -= L0;
                        cp_async16(raw + cp_dst[k], in_b + ((long long)cp_ch[k] * L0 + rr) * L1 + cp_col[k], true);
                    } else {
                        const bool live = rr < L0;                        // rows past the lattice: zero-filled
                        cp_async16(raw + cp_dst[k], g_b + ((long long)(cp_ch[k] - 8) * L0 + (live ? rr : 0)) * L1 + cp_col[k],
                                   live);
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");          // (an empty group keeps the count in step)
        };

        float acc[32];                                                    // drain warps: partial sums of my TMEM lane
#pragma unroll
        for (int n = 0; n < 32; ++n) acc[n] = 0.f;
        uint32_t drain_phase = 0;
        auto drain = [&]() {
            tc::mbar_wait(tc::smem_u32(bars + 2), drain_phase);
            drain_phase ^= 1u;
            tc::fence_after_sync();
            const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
            float d[16];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                tc::tmem_ld16(lane_addr + 16 * h, d);
                tc::tmem_ld_wait();
#pragma unroll
                for (int n = 0; n < 16; ++n) acc[16 * h + n] += d[n];
            }
            tc::fence_before_sync();
            tc::bar_arrive(kBarDrained, kDrainWarps * 32 + 32);
        };

        Cursor cur{(long long)(blockIdx.x / strips), (int)(blockIdx.x % strips)}, ahead = cur;
        issue(0, ahead, 0);
        advance(ahead);
        issue(1, ahead, 1 % kRawStages);
        advance(ahead);
        DrainClock clock{a.drain_every, 0};
        int tn = 0;
        auto stamp = [&]() {
            if (a.trace && blockIdx.x == 0 && tid == 160 && tn < 1000) g_wg_trace[tn++] = clock64();
        };
        bool drain_due = false;                                           // tile i - 1 ended a drain interval
        int stage = 0, stage_ahead = 2 % kRawStages;
        for (long long i = 0; i < n_tiles; ++i, advance(cur)) {
            const int s = (int)(i & 1);
            if (warp < kDrainWarps && drain_due) drain();
            drain_due = clock.tick(i + 1 == n_tiles);
            stamp();
            asm volatile("cp.async.wait_group 1;" ::: "memory");          // all but the newest group: tile i is here
            stamp();
            tc::bar_sync(kBarProducers, kProducerThreads);    // ... for everyone, and the stage of tile i - 1 is free
            stamp();
            issue(i + 2, ahead, stage_ahead);
            advance(ahead);
            stage_ahead = stage_ahead + 1 == kRawStages ? 0 : stage_ahead + 1;
            stamp();
            tc::mbar_wait(tc::smem_u32(bars + s), (uint32_t)(((i >> 1) & 1) ^ 1));     // MMAs of tile i - 2 done
            stamp();
            const int r0 = cur.strip * R;
            const float* raw = reinterpret_cast<const float*>(raw0 + (size_t)stage * a.raw_bytes);
            stage = stage + 1 == kRawStages ? 0 : stage + 1;
            unsigned char* hi = part(s, 0);
            unsigned char* lo = part(s, 1);
#pragma unroll
            for (int k = 0; k < kSlots; ++k) {
                if (kind[k] < 0) continue;
                const int par = SPARSE ? ((a.g_parity + r0 + jrow[k]) & 1) : 0;
                const float* row = raw + src[k];
                const int cc0 = cc0s[k];
                if (kind[k] == 0) {
                    // eight consecutive active sites of one (channel, tap): columns c0 + o + (SPARSE ? 2 e : e),
                    // o = kw - 1 (+ the row's parity offset), read as aligned groups of four starting at or below
                    const int o = kws[k] - 1 + par;                       // -1 .. 2, the same for the whole warp
                    const float* base = row + (SPARSE ? 2 * cc0 : cc0) + (o < 0 ? -4 : 0);
                    constexpr int NV = SPARSE ? 5 : 3;
                    float v[4 * NV];
                    load_row<NV>(base, v);
                    float vals[8];
                    switch (o) {
                        case -1:
#pragma unroll
                            for (int e = 0; e < 8; ++e) vals[e] = v[3 + (SPARSE ? 2 * e : e)];
                            break;
                        case 0:
#pragma unroll
                            for (int e = 0; e < 8; ++e) vals[e] = v[(SPARSE ? 2 * e : e)];
                            break;
                        case 1:
#pragma unroll
                            for (int e = 0; e < 8; ++e) vals[e] = v[1 + (SPARSE ? 2 * e : e)];
                            break;
                        default:
#pragma unroll
                            for (int e = 0; e < 8; ++e) vals[e] = v[2 + (SPARSE ? 2 * e : e)];
                            break;
                    }
                    split_store(vals, hi + dst[k], lo + dst[k]);
                } else {
                    constexpr int NV = SPARSE ? 4 : 2;
                    float v[4 * NV];
                    load_row<NV>(row + (SPARSE ? 2 * cc0 : cc0), v);
                    float vals[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) vals[e] = SPARSE ? (par ? v[2 * e + 1] : v[2 * e]) : v[e];
                    split_store(vals, hi + dst[k], lo + dst[k]);
                }
            }
            stamp();
            tc::fence_async_smem();                                       // my records -> visible to the tensor core
            tc::bar_arrive(kBarFull0 + s, kProducerThreads + 32);
            stamp();
        }
        if (warp < kDrainWarps && n_tiles > 0) {
            drain();                                                      // the last tile is always a drain point
            const int m = warp * 32 + lane;
            if (m < 72) {
                const int tap = m >> 3, ci = m & 7;
#pragma unroll
                for (int n = 0; n < 32; ++n)
                    if (n < Co) atomicAdd(a.gw + ((long long)n * 8 + ci) * 9 + tap, acc[n]);
            } else if (m == kOnesRow && a.gbias) {
#pragma unroll
                for (int n = 0; n < 32; ++n)
                    if (n < Co) atomicAdd(a.gbias + n, acc[n]);
            }
        }
    } else {
        // =============================== MMA warp ==============================================
        const bool lead = tc::elect_one();
        const uint32_t idesc = tc::make_idesc(2, 128, 32);               // tf32 operands, K = 8 per instruction
        DrainClock clock{a.drain_every, 0};
        constexpr uint64_t kStepDesc = (2 * kGroupBytes) >> 4;            // never carries out of the 14-bit address field
        uint64_t d_a_hi[2], d_a_lo[2], d_b_hi[2], d_b_lo[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const uint32_t hi = tc::smem_u32(part(s, 0)), lo = tc::smem_u32(part(s, 1));
            d_a_hi[s] = tc::make_desc(hi, kGroupBytes, 128);
            d_a_lo[s] = tc::make_desc(lo, kGroupBytes, 128);
            d_b_hi[s] = tc::make_desc(hi + kBRow0 * 16, kGroupBytes, 128);
            d_b_lo[s] = tc::make_desc(lo + kBRow0 * 16, kGroupBytes, 128);
        }
        uint32_t acc_on = 0;
        for (long long i = 0; i < n_tiles; ++i) {
            const int s = (int)(i & 1);
            if (a.trace && blockIdx.x == 0 && lead && i < 250) g_wg_trace[1024 + 4 * i] = clock64();
            tc::bar_sync(kBarFull0 + s, kProducerThreads + 32);           // tile i is in buffer s
            tc::fence_after_sync();
            if (a.trace && blockIdx.x == 0 && lead && i < 250) g_wg_trace[1024 + 4 * i + 1] = clock64();
            if (lead) {
                // descriptors of k-step 0; a k-step (two k-groups) further is +2 * kGroupBytes / 16 in the address field
                uint64_t a_hi = d_a_hi[s], a_lo = d_a_lo[s], b_hi = d_b_hi[s], b_lo = d_b_lo[s];
                for (int ks = 0; ks < (KT >> 3); ++ks) {
                    tc::mma_tf32(tmem, a_hi, b_hi, idesc, acc_on);
                    tc::mma_tf32(tmem, a_hi, b_lo, idesc, 1u);
                    tc::mma_tf32(tmem, a_lo, b_hi, idesc, 1u);
                    acc_on = 1u;
                    a_hi += kStepDesc; a_lo += kStepDesc; b_hi += kStepDesc; b_lo += kStepDesc;
                }
                tc::mma_commit(tc::smem_u32(bars + s));                   // buffer s free when these complete
                if (a.trace && blockIdx.x == 0 && i < 250) g_wg_trace[1024 + 4 * i + 2] = clock64();
            }
            __syncwarp();
            if (clock.tick(i + 1 == n_tiles)) {
                if (lead) tc::mma_commit(tc::smem_u32(bars + 2));
                __syncwarp();
                tc::bar_sync(kBarDrained, kDrainWarps * 32 + 32);         // registers hold the partial sums
                tc::fence_after_sync();
                acc_on = 0u;
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) tc::tmem_dealloc(tmem, kTmemCols);
}

template <bool SPARSE>
int launch(WgTcArgs a, cudaStream_t st) {
    const size_t smem = (size_t)4 * (a.KT >> 2) * kGroupBytes + kRawStages * (size_t)a.raw_bytes;   // operands + raw stages
    if (smem > 226 * 1024) return NFK_EUNSUPPORTED;
    if (ensure_dynamic_smem<wgrad2d_tc_kernel<SPARSE>>(226 * 1024) != NFK_OK) return NFK_ECUDA;
    static const int sm_count = [] {
        int dev = 0, n = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        return n > 0 ? n : 148;
    }();
    const long long tiles = a.B * ((a.L0 + a.R - 1) / a.R);
    const long long grid = tiles < sm_count ? tiles : sm_count;
    wgrad2d_tc_kernel<SPARSE><<<(unsigned)grid, kThreads, smem, st>>>(a);
    return check_launch();
}

}  // namespace

extern "C" int nfk_conv2d_wgrad_tc(const float* in, const float* gpre, int g_parity, float* gw, float* gbias,
                                   int L0, int L1, int Ci, int Co, int64_t B, void* stream) {
    if (!in || !gpre || !gw || L0 < 1 || L1 < 1 || Co < 1 || g_parity < -1 || g_parity > 1) return NFK_EINVAL;
    if (B <= 0) return NFK_OK;
    const bool sparse = g_parity >= 0;
    if (Ci != 8 || Co > 32 || L0 < 2) return NFK_EUNSUPPORTED;
    if (sparse && L0 % 2 != 0) return NFK_EUNSUPPORTED;
    if (((uintptr_t)in | (uintptr_t)gpre) % 16 != 0) return NFK_EUNSUPPORTED;
    WgTcArgs a;
    a.in = in; a.g = gpre; a.gw = gw; a.gbias = gbias;
    a.L0 = L0; a.L1 = L1; a.Co = Co; a.B = B; a.g_parity = g_parity;
    a.ncols = sparse ? L1 / 2 : L1;
    if (L1 % (sparse ? 16 : 8) != 0) return NFK_EUNSUPPORTED;             // whole groups of 8 (active) sites per row
    a.ncols8 = a.ncols;
    if (a.ncols > kMaxKT) return NFK_EUNSUPPORTED;
    int R = kMaxKT / a.ncols;                                             // R * ncols is a multiple of 8 = one MMA's K
    if (R > L0) R = L0;
    a.R = R;
    a.KT = R * a.ncols;
    auto bank_spread = [](int floats) { return floats + ((4 - floats % 32) + 32) % 32; };      // -> 4 mod 32
    a.row_stride = L1 + 8;
    a.xch_stride = bank_spread((R + 2) * a.row_stride);
    a.gch_stride = bank_spread(R * a.row_stride);
    a.raw_bytes = (8 * a.xch_stride + Co * a.gch_stride) * (int)sizeof(float);
    {   // per-thread slot capacity of the producers
        const int groups = a.KT / 8;
        const int tasks = (72 + Co) * groups;
        const int chunks = (8 * (R + 2) + Co * R) * (L1 / 4 + 2);
        if (tasks > kSlots * kProducerThreads || chunks > kCpSlots * kProducerThreads) return NFK_EUNSUPPORTED;
    }
    a.trace = getenv("NFK_WGRAD_TRACE") ? 1 : 0;
    a.drain_every = 4;
    if (const char* e = getenv("NFK_WGRAD_DRAIN")) {                      // tuning knob
        const int v = atoi(e);
        if (v >= 1 && v <= 1024) a.drain_every = v;
    }
    cudaStream_t st = NFK_STREAM(stream);
    return sparse ? launch<true>(a, st) : launch<false>(a, st);
}

extern "C" int nfk_debug_wgrad_trace(long long* host_out, int n) {
    return cudaMemcpyFromSymbol(host_out, g_wg_trace, sizeof(long long) * (size_t)(n < 2048 ? n : 2048)) == cudaSuccess
               ? NFK_OK : NFK_ECUDA;
}
