// nfk_conv.cu -- circular "same" convolution layers of the ConvAct conditioner
// (nn/scalar/modules.py:131-145, convNd.py:84-127), any lattice dimension 1..4,
// fp32 on the CUDA cores (the reference's channel widths, C = 1..8 -> P <= 28, are
// far too thin for a tcgen05 tile).
//
//   forward / data-gradient : one thread per output site, CO_BLK output channels in
//       registers, weights for the channel block staged in shared memory and read
//       as warp-wide broadcasts; periodic neighbours by index arithmetic (no padded
//       copy of the input is ever made).
//   weight-gradient         : threads walk sites, a (CO_B x CI_B x T_B) block of the
//       weight gradient lives in registers; one shuffle/shared reduction per CTA at
//       the end, then atomics.

#include "nfk_common.cuh"

using namespace nfk;

#define NFK_STREAM(s) reinterpret_cast<cudaStream_t>(s)

struct ConvArgs {
    const float* in;
    const float* w;
    int w_transposed;
    const float* bias;
    const uint8_t* in_mask;
    int in_keep;
    int act;
    const float* dact_from;
    int dact_kind;
    float* out;
    Lat lat;
    int ksize, T, Ci, Co, V, tiles;
};

template <int CO>
__global__ void __launch_bounds__(128) conv_fwd_kernel(ConvArgs a) {
    extern __shared__ float wt[];                 // [Ci*T][CO]
    const int co0 = blockIdx.y * CO;
    for (int i = threadIdx.x; i < a.Ci * a.T * CO; i += blockDim.x) {
        const int co = i % CO, r = i / CO;        // r = ci*T + t
        float v = 0.f;
        if (co0 + co < a.Co) {
            if (!a.w_transposed) {
                v = a.w[(int64_t)(co0 + co) * a.Ci * a.T + r];
            } else {                              // forward weight is [Ci][Co][T]; flip the taps
                const int ci = r / a.T, t = r % a.T;
                v = a.w[((int64_t)ci * a.Co + co0 + co) * a.T + (a.T - 1 - t)];
            }
        }
        wt[i] = v;
    }
    __syncthreads();
    const int64_t b = blockIdx.x / a.tiles;
    const int s = (int)(blockIdx.x % a.tiles) * blockDim.x + threadIdx.x;
    if (s >= a.V) return;
    float acc[CO];
#pragma unroll
    for (int co = 0; co < CO; ++co) acc[co] = (a.bias && co0 + co < a.Co) ? a.bias[co0 + co] : 0.f;
    conv_site<CO>(a.in + b * a.Ci * (int64_t)a.V, wt, a.in_mask, a.in_keep, a.lat, s, a.Ci, a.T, a.ksize,
                  a.V, acc);
#pragma unroll
    for (int co = 0; co < CO; ++co) {
        if (co0 + co >= a.Co) break;
        const int64_t o = (b * a.Co + co0 + co) * (int64_t)a.V + s;
        float v = act_apply(a.act, acc[co]);
        if (a.dact_from) v *= act_grad_from_post(a.dact_kind, a.dact_from[o]);
        a.out[o] = v;
    }
}

// ---------------------------------------------------------------- 2-D 3x3 layers, shared-memory tiled
// ConvAct layers on 2-D lattices (forward and data-gradient form) for 8 output channels at a time:
// a CTA owns a strip of R rows of one sample; input channels are staged through shared memory in
// chunks of 8 (rows with their periodic halo); a thread keeps 4 consecutive columns x 8 output
// channels in registers and, per input channel, reads its 3 x 6 window once (18 scalar loads)
// plus the 72 weights as warp-wide broadcasts for 288 FMAs.  Replaces the per-site global-memory
// gathers of the generic kernel (2-3x on the 8->8 / 28->8 layers of the training backward).
struct Conv2dArgs {
    const float* in;
    const float* w;
    int w_transposed;
    const float* bias;
    const uint8_t* in_mask;
    int in_keep, act;
    const float* dact_from;
    int dact_kind;
    float* out;
    int L0, L1, R, Ci, Co, strips;
    int in_parity;               // >= 0: `in` vanishes off that checkerboard partition (zero products are skipped)
};

constexpr int kC2Chunk = 8;      // input channels per shared-memory stage

// NC columns x 8 output channels += 3 x (NC + 2) window (x) 9 x 8 weights of one input channel.  SPARSE: the window
// entry (dr, k + dc) is known to be zero unless (dr + dc + k + A0) is even, and its products are left out.
template <bool SPARSE, int A0, int NC, int NCO>
__device__ __forceinline__ void conv2d_tile_fma(const float (&win)[3][NC + 2], const float* __restrict__ w_ci,
                                                float (&acc)[NC][NCO]) {
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        float wv[NCO];
        if (NCO == 8) {
            const float4 w0 = *reinterpret_cast<const float4*>(w_ci + t * 8);
            const float4 w1 = *reinterpret_cast<const float4*>(w_ci + t * 8 + 4);
            wv[0] = w0.x; wv[1 % NCO] = w0.y; wv[2 % NCO] = w0.z; wv[3 % NCO] = w0.w;
            wv[4 % NCO] = w1.x; wv[5 % NCO] = w1.y; wv[6 % NCO] = w1.z; wv[7 % NCO] = w1.w;
        } else {
#pragma unroll
            for (int co = 0; co < NCO; ++co) wv[co] = w_ci[t * 8 + co];
        }
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            if (SPARSE && ((t / 3 + t % 3 + k + A0) & 1)) continue;
#pragma unroll
            for (int co = 0; co < NCO; ++co) acc[k][co] = fmaf(win[t / 3][k + t % 3], wv[co], acc[k][co]);
        }
    }
}

// L1C, RC > 0: row length / rows per strip fixed at compile time (shared-memory addresses become base + immediate).
// NC: columns per thread (4 or 8; with 8 every weight fetched from shared memory feeds twice the FMAs).
// NCO: output channels per CTA (8, or 1 for the thin 8 -> 1 data gradient of the first conditioner layer).
template <int L1C, int RC, int NC, int NCO>
__global__ void __launch_bounds__(256) conv2d_tile_kernel(Conv2dArgs a) {
    extern __shared__ __align__(16) float sm2[];
    const int L0 = a.L0, L1 = L1C > 0 ? L1C : a.L1, R = RC > 0 ? RC : a.R;
    const int LW = L1 + 8;                                             // interior at +4 (16-byte aligned), halos at +3, +4+L1
    float* in_s = sm2;                                                 // [kC2Chunk][R + 2][LW]
    float* w_s = sm2 + kC2Chunk * (R + 2) * LW;                        // [kC2Chunk][9][8]
    const int tid = threadIdx.x, nq = L1 >> 2, npr = L1 / NC;          // float4 groups / threads per row
    const long long b = blockIdx.x / a.strips;
    const int r0 = (int)(blockIdx.x % a.strips) * R;
    const int rows = L0 - r0 < R ? L0 - r0 : R;
    const int co0 = blockIdx.y * NCO;
    const int V = L0 * L1;
    // my row of the strip and my NC columns.  The rows a warp spans are two apart (within groups of 2 * rows-per-warp
    // rows the even ones come first), so that a warp sees ONE row parity and the sparse variants do not diverge.
    const int jj = tid / npr, c0 = (tid - jj * npr) * NC;
    const int rpw = (npr <= 32 && 32 % npr == 0) ? 32 / npr : 1;
    const int j = (rpw > 1 && R % (2 * rpw) == 0) ? (jj / (2 * rpw)) * (2 * rpw) + (jj % rpw) * 2 + (jj / rpw) % 2 : jj;
    const bool live = j < rows;
    float acc[NC][NCO];
#pragma unroll
    for (int k = 0; k < NC; ++k)
#pragma unroll
        for (int co = 0; co < NCO; ++co) acc[k][co] = (a.bias && co0 + co < a.Co) ? __ldg(a.bias + co0 + co) : 0.f;
    const float* in_b = a.in + b * (long long)a.Ci * V;
    for (int ci0 = 0; ci0 < a.Ci; ci0 += kC2Chunk) {
        const int nci = a.Ci - ci0 < kC2Chunk ? a.Ci - ci0 : kC2Chunk;
        __syncthreads();                                               // previous chunk consumed
        for (int e = tid; e < nci * (rows + 2) * (nq + 1); e += blockDim.x) {
            const int q = e % (nq + 1), jj = (e / (nq + 1)) % (rows + 2), ci = e / ((nq + 1) * (rows + 2));
            int r = r0 - 1 + jj;
            r = r < 0 ? r + L0 : (r >= L0 ? r - L0 : r);
            const float* src = in_b + (long long)(ci0 + ci) * V + r * L1;
            const uint8_t* msk = a.in_mask ? a.in_mask + r * L1 : nullptr;
            float* dst = in_s + (ci * (R + 2) + jj) * LW;
            if (q < nq) {
                float4 v = __ldg(reinterpret_cast<const float4*>(src + 4 * q));
                if (msk) {
                    const uint32_t m = __ldg(reinterpret_cast<const uint32_t*>(msk + 4 * q));
                    if ((int)(m & 0xFF) != a.in_keep) v.x = 0.f;
                    if ((int)((m >> 8) & 0xFF) != a.in_keep) v.y = 0.f;
                    if ((int)((m >> 16) & 0xFF) != a.in_keep) v.z = 0.f;
                    if ((int)(m >> 24) != a.in_keep) v.w = 0.f;
                }
                *reinterpret_cast<float4*>(dst + 4 + 4 * q) = v;
            } else {                                                   // the two periodic halo columns
                float vl = __ldg(src + L1 - 1), vr = __ldg(src);
                if (msk) {
                    if ((int)__ldg(msk + L1 - 1) != a.in_keep) vl = 0.f;
                    if ((int)__ldg(msk) != a.in_keep) vr = 0.f;
                }
                dst[3] = vl;
                dst[4 + L1] = vr;
            }
        }
        for (int e = tid; e < nci * 72; e += blockDim.x) {
            const int co = e & 7, t = (e >> 3) % 9, ci = e / 72;
            float v = 0.f;
            if (co0 + co < a.Co) {
                if (!a.w_transposed) v = __ldg(a.w + ((long long)(co0 + co) * a.Ci + ci0 + ci) * 9 + t);
                else v = __ldg(a.w + ((long long)(ci0 + ci) * a.Co + co0 + co) * 9 + (8 - t));   // [Ci][Co][taps], flipped
            }
            w_s[e] = v;
        }
        __syncthreads();
        if (live) {
            for (int ci = 0; ci < nci; ++ci) {
                float win[3][NC + 2];
#pragma unroll
                for (int dr = 0; dr < 3; ++dr) {
                    // columns c0 - 1 .. c0 + NC as aligned groups of four: [c0-4, c0), NC/4 interior groups, [c0+NC, c0+NC+4)
                    const float* p = in_s + (ci * (R + 2) + j + dr) * LW + 4 + c0;
                    win[dr][0] = reinterpret_cast<const float4*>(p - 4)->w;
#pragma unroll
                    for (int q = 0; q < NC / 4; ++q) {
                        const float4 v = *reinterpret_cast<const float4*>(p + 4 * q);
                        win[dr][1 + 4 * q] = v.x; win[dr][2 + 4 * q] = v.y; win[dr][3 + 4 * q] = v.z; win[dr][4 + 4 * q] = v.w;
                    }
                    win[dr][NC + 1] = p[NC];
                }
                // input site (r0 + j + dr - 1, c0 + k + dc - 1) is on the partition iff dr + dc + k + a0 is even
                const float* w_ci = w_s + ci * 72;
                if (a.in_parity < 0) conv2d_tile_fma<false, 0, NC, NCO>(win, w_ci, acc);
                else if ((r0 + j + a.in_parity) & 1) conv2d_tile_fma<true, 1, NC, NCO>(win, w_ci, acc);
                else conv2d_tile_fma<true, 0, NC, NCO>(win, w_ci, acc);
            }
        }
    }
    if (!live) return;
#pragma unroll
    for (int co = 0; co < NCO; ++co) {
        if (co0 + co >= a.Co) break;
        const long long o = ((b * a.Co + co0 + co) * (long long)L0 + r0 + j) * L1 + c0;
#pragma unroll
        for (int q = 0; q < NC / 4; ++q) {
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = act_apply(a.act, acc[4 * q + k][co]);
            if (a.dact_from) {
                const float4 h = __ldg(reinterpret_cast<const float4*>(a.dact_from + o + 4 * q));
                v[0] *= act_grad_from_post(a.dact_kind, h.x);
                v[1] *= act_grad_from_post(a.dact_kind, h.y);
                v[2] *= act_grad_from_post(a.dact_kind, h.z);
                v[3] *= act_grad_from_post(a.dact_kind, h.w);
            }
            *reinterpret_cast<float4*>(a.out + o + 4 * q) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// eligible: 2-D, 3x3, rows of whole float4s, at least 4 output channels (thinner layers are
// load-bound either way), 16-byte aligned tensors
static int conv2d_tile_launch(const float* in, const float* w, int w_transposed, const float* bias,
                              const uint8_t* in_mask, int in_keep, int act, const float* dact_from, int dact_kind,
                              float* out, int L0, int L1, int Ci, int Co, int64_t B, int in_parity, cudaStream_t st) {
    // 4 columns per thread: with 8 every weight fetch feeds twice the FMAs, but the 64 accumulators cost a
    // third of the resident warps and the layer ran 20 % slower (measured, 64x64)
    constexpr int NC = 4;
    const int npr = L1 / NC;
    if (npr > 256) return NFK_EUNSUPPORTED;
    int R = 256 / npr;
    if (R > L0) R = L0;
    if (R > 32) R = 32;
    Conv2dArgs a;
    a.in = in; a.w = w; a.w_transposed = w_transposed; a.bias = bias; a.in_mask = in_mask; a.in_keep = in_keep;
    a.act = act; a.dact_from = dact_from; a.dact_kind = dact_kind; a.out = out;
    a.L0 = L0; a.L1 = L1; a.R = R; a.Ci = Ci; a.Co = Co; a.strips = (L0 + R - 1) / R;
    a.in_parity = (in_parity >= 0 && L0 % 2 == 0 && L1 % 2 == 0 && in_mask == nullptr) ? in_parity : -1;
    const size_t smem = (size_t)(kC2Chunk * (R + 2) * (L1 + 8) + kC2Chunk * 72) * sizeof(float);
    if (smem > 160 * 1024) return NFK_EUNSUPPORTED;
    const int threads = (R * npr + 31) / 32 * 32;
    if (Co == 1) {                                                        // thin layer: one output channel per CTA
        const dim3 grid1((unsigned)(B * a.strips), 1u);
        if (L1 == 64 && R == 16) {
            if (ensure_dynamic_smem<conv2d_tile_kernel<64, 16, NC, 1>>(160 * 1024) != NFK_OK) return NFK_ECUDA;
            conv2d_tile_kernel<64, 16, NC, 1><<<grid1, threads, smem, st>>>(a);
        } else {
            if (ensure_dynamic_smem<conv2d_tile_kernel<0, 0, NC, 1>>(160 * 1024) != NFK_OK) return NFK_ECUDA;
            conv2d_tile_kernel<0, 0, NC, 1><<<grid1, threads, smem, st>>>(a);
        }
        return check_launch();
    }
    const dim3 grid((unsigned)(B * a.strips), (unsigned)((Co + 7) / 8));
    if (L1 == 64 && R == 16) {                                            // the benchmark geometry: compile-time strides
        if (ensure_dynamic_smem<conv2d_tile_kernel<64, 16, NC, 8>>(160 * 1024) != NFK_OK) return NFK_ECUDA;
        conv2d_tile_kernel<64, 16, NC, 8><<<grid, threads, smem, st>>>(a);
    } else if (L1 == 32 && R == 32) {
        if (ensure_dynamic_smem<conv2d_tile_kernel<32, 32, NC, 8>>(160 * 1024) != NFK_OK) return NFK_ECUDA;
        conv2d_tile_kernel<32, 32, NC, 8><<<grid, threads, smem, st>>>(a);
    } else {
        if (ensure_dynamic_smem<conv2d_tile_kernel<0, 0, NC, 8>>(160 * 1024) != NFK_OK) return NFK_ECUDA;
        conv2d_tile_kernel<0, 0, NC, 8><<<grid, threads, smem, st>>>(a);
    }
    return check_launch();
}

static int conv_fwd_impl(const float* in, const float* w, int w_transposed, const float* bias,
                         const uint8_t* in_mask, int in_keep, int act, const float* dact_from, int dact_kind,
                         float* out, nfk_lattice lat, int ksize, int Ci, int Co, int64_t B, int in_parity,
                         void* stream);

extern "C" int nfk_conv_circ_fwd(const float* in, const float* w, int w_transposed, const float* bias,
                                 const uint8_t* in_mask, int in_keep,
                                 int act, const float* dact_from, int dact_kind,
                                 float* out, nfk_lattice lat, int ksize,
                                 int Ci, int Co, int64_t B, void* stream) {
    return conv_fwd_impl(in, w, w_transposed, bias, in_mask, in_keep, act, dact_from, dact_kind, out, lat, ksize,
                         Ci, Co, B, -1, stream);
}

extern "C" int nfk_conv_circ_fwd_cb(const float* in, int in_parity, const float* w, int w_transposed,
                                    const float* bias, int act, const float* dact_from, int dact_kind,
                                    float* out, nfk_lattice lat, int ksize, int Ci, int Co, int64_t B,
                                    void* stream) {
    if (in_parity != 0 && in_parity != 1) return NFK_EINVAL;
    return conv_fwd_impl(in, w, w_transposed, bias, nullptr, 0, act, dact_from, dact_kind, out, lat, ksize,
                         Ci, Co, B, in_parity, stream);
}

static int conv_fwd_impl(const float* in, const float* w, int w_transposed, const float* bias,
                         const uint8_t* in_mask, int in_keep, int act, const float* dact_from, int dact_kind,
                         float* out, nfk_lattice lat, int ksize, int Ci, int Co, int64_t B, int in_parity,
                         void* stream) {
    if (!in || !w || !out || !lat_ok(lat) || ksize < 1 || ksize % 2 == 0 || Ci < 1 || Co < 1) return NFK_EINVAL;
    if (B <= 0) return NFK_OK;
    if (lat.ndim == 2 && ksize == 3 && (Co >= 4 || (Co == 1 && Ci >= 4)) && lat.shape[1] % 4 == 0 && lat.shape[0] >= 2 && lat.shape[1] >= 4 &&
        B * (int64_t)((lat.shape[0] + 0) ) < (int64_t(1) << 31) && ((uintptr_t)in % 16) == 0 &&
        ((uintptr_t)out % 16) == 0 && (!dact_from || ((uintptr_t)dact_from % 16) == 0) &&
        (!in_mask || ((uintptr_t)in_mask % 4) == 0)) {
        const int rc = conv2d_tile_launch(in, w, w_transposed, bias, in_mask, in_keep, act, dact_from, dact_kind, out,
                                          lat.shape[0], lat.shape[1], Ci, Co, B, in_parity, NFK_STREAM(stream));
        if (rc != NFK_EUNSUPPORTED) return rc;
    }
    ConvArgs a;
    a.in = in; a.w = w; a.w_transposed = w_transposed; a.bias = bias; a.in_mask = in_mask; a.in_keep = in_keep;
    a.act = act; a.dact_from = dact_from; a.dact_kind = dact_kind; a.out = out;
    a.lat = to_lat(lat); a.ksize = ksize; a.Ci = Ci; a.Co = Co;
    a.V = (int)lat_volume(lat);
    int T = 1;
    for (int d = 0; d < lat.ndim; ++d) T *= ksize;
    a.T = T;
    const int threads = a.V >= 128 ? 128 : ((a.V + 31) / 32) * 32;
    a.tiles = (a.V + threads - 1) / threads;
    const int CO = Co >= 8 ? 8 : (Co >= 4 ? 4 : (Co >= 2 ? 2 : 1));
    const dim3 grid((unsigned)(B * a.tiles), (unsigned)((Co + CO - 1) / CO));
    const size_t smem = (size_t)Ci * T * CO * sizeof(float);
    if (smem > 200 * 1024) return NFK_EUNSUPPORTED;
    cudaStream_t st = NFK_STREAM(stream);
#define LAUNCH(N)                                                                                     \
    {                                                                                                 \
        if (smem > 48 * 1024 && ensure_dynamic_smem<conv_fwd_kernel<N>>(200 * 1024) != NFK_OK)        \
            return NFK_ECUDA;                                                                         \
        conv_fwd_kernel<N><<<grid, threads, smem, st>>>(a);                                           \
    }
    switch (CO) {
        case 8: LAUNCH(8) break;
        case 4: LAUNCH(4) break;
        case 2: LAUNCH(2) break;
        default: LAUNCH(1) break;
    }
#undef LAUNCH
    return check_launch();
}

// ---------------------------------------------------------------- weight gradient
struct ConvWArgs {
    const float* in;
    const uint8_t* in_mask;
    int in_keep;
    const float* gpre;
    float* gw;
    float* gbias;
    Lat lat;
    int ksize, T, Ci, Co, V;
    int n_tb, n_cib;          // tap blocks, ci blocks (grid.y = co blocks * n_cib * n_tb)
    int64_t BV;
};

template <int CO_B, int CI_B, int T_B>
__global__ void __launch_bounds__(256) conv_bwd_weight_kernel(ConvWArgs a) {
    constexpr int NACC = CO_B * CI_B * T_B;
    __shared__ float red[8][NACC + CO_B];
    int y = blockIdx.y;
    const int tb = y % a.n_tb; y /= a.n_tb;
    const int cib = y % a.n_cib; y /= a.n_cib;
    const int co0 = y * CO_B, ci0 = cib * CI_B, t0 = tb * T_B;
    const bool do_bias = a.gbias != nullptr && cib == 0 && tb == 0;

    float acc[NACC];
    float accb[CO_B];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < CO_B; ++i) accb[i] = 0.f;

    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < a.BV;
         g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = g / a.V;
        const int s = (int)(g % a.V);
        float gv[CO_B];
#pragma unroll
        for (int co = 0; co < CO_B; ++co)
            gv[co] = (co0 + co < a.Co) ? __ldg(a.gpre + (b * a.Co + co0 + co) * (int64_t)a.V + s) : 0.f;
        if (do_bias) {
#pragma unroll
            for (int co = 0; co < CO_B; ++co) accb[co] += gv[co];
        }
        int c[4];
        site_coords(a.lat, s, c);
        const float* in_b = a.in + b * a.Ci * (int64_t)a.V;
#pragma unroll
        for (int t = 0; t < T_B; ++t) {
            if (t0 + t >= a.T) break;
            const int n = tap_neighbor(a.lat, s, c, t0 + t, a.ksize);
            const bool keep = !a.in_mask || __ldg(a.in_mask + n) == (uint8_t)a.in_keep;
#pragma unroll
            for (int ci = 0; ci < CI_B; ++ci) {
                const float v = (keep && ci0 + ci < a.Ci) ? __ldg(in_b + (int64_t)(ci0 + ci) * a.V + n) : 0.f;
#pragma unroll
                for (int co = 0; co < CO_B; ++co)
                    acc[(co * CI_B + ci) * T_B + t] = fmaf(gv[co], v, acc[(co * CI_B + ci) * T_B + t]);
            }
        }
    }
    // CTA reduction: shuffle within warps, shared across the (<= 8) warps, then atomics
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        const float v = warp_sum(acc[i]);
        if (lane == 0) red[wid][i] = v;
    }
#pragma unroll
    for (int i = 0; i < CO_B; ++i) {
        const float v = warp_sum(accb[i]);
        if (lane == 0) red[wid][NACC + i] = v;
    }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < NACC + CO_B; i += blockDim.x) {
        float v = 0.f;
        for (int w = 0; w < nw; ++w) v += red[w][i];
        if (i < NACC) {
            const int t = i % T_B, ci = (i / T_B) % CI_B, co = i / (T_B * CI_B);
            if (co0 + co < a.Co && ci0 + ci < a.Ci && t0 + t < a.T)
                atomicAdd(a.gw + ((int64_t)(co0 + co) * a.Ci + ci0 + ci) * a.T + t0 + t, v);
        } else if (do_bias && co0 + (i - NACC) < a.Co) {
            atomicAdd(a.gbias + co0 + (i - NACC), v);
        }
    }
}

// The same arithmetic with the tap blocks dealt to the WARPS of a CTA instead of to grid.y: every warp of a CTA
// walks the same sites, so the d loss / d pre-activation and the input values of a site come from HBM / L2 once per
// block of output channels (and are L1 hits for the other warps) instead of once per (output block, tap block) --
// 63 passes over the data for a 3-D 8 -> 28 layer became 7 (the kernel was bound by those re-reads: 74 % of a 3-D
// training step).  A warp owns its accumulators outright: warp reduction, then atomics.
template <int CO_B, int CI_B, int T_B>
__global__ void __launch_bounds__(32 * 14) conv_bwd_weight_warptap_kernel(ConvWArgs a, int tb_per_cta) {
    constexpr int NACC = CO_B * CI_B * T_B;
    int y = blockIdx.y;
    const int tgroup = y % a.n_tb; y /= a.n_tb;               // here n_tb = tap-block groups per CTA row
    const int cib = y % a.n_cib; y /= a.n_cib;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int tb = tgroup * tb_per_cta + wid;
    const int co0 = y * CO_B, ci0 = cib * CI_B, t0 = tb * T_B;
    if (t0 >= a.T) return;
    const bool do_bias = a.gbias != nullptr && cib == 0 && tb == 0;

    float acc[NACC];
    float accb[CO_B];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < CO_B; ++i) accb[i] = 0.f;

    for (int64_t g = (int64_t)blockIdx.x * 32 + lane; g < a.BV; g += (int64_t)gridDim.x * 32) {
        const int64_t b = g / a.V;
        const int s = (int)(g - b * a.V);
        float gv[CO_B];
#pragma unroll
        for (int co = 0; co < CO_B; ++co)
            gv[co] = (co0 + co < a.Co) ? __ldg(a.gpre + (b * a.Co + co0 + co) * (int64_t)a.V + s) : 0.f;
        if (do_bias) {
#pragma unroll
            for (int co = 0; co < CO_B; ++co) accb[co] += gv[co];
        }
        int c[4];
        site_coords(a.lat, s, c);
        const float* in_b = a.in + b * a.Ci * (int64_t)a.V;
#pragma unroll
        for (int t = 0; t < T_B; ++t) {
            if (t0 + t >= a.T) break;
            const int n = tap_neighbor(a.lat, s, c, t0 + t, a.ksize);
            const bool keep = !a.in_mask || __ldg(a.in_mask + n) == (uint8_t)a.in_keep;
#pragma unroll
            for (int ci = 0; ci < CI_B; ++ci) {
                const float v = (keep && ci0 + ci < a.Ci) ? __ldg(in_b + (int64_t)(ci0 + ci) * a.V + n) : 0.f;
#pragma unroll
                for (int co = 0; co < CO_B; ++co)
                    acc[(co * CI_B + ci) * T_B + t] = fmaf(gv[co], v, acc[(co * CI_B + ci) * T_B + t]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        const float v = warp_sum(acc[i]);
        const int t = i % T_B, ci = (i / T_B) % CI_B, co = i / (T_B * CI_B);
        if (lane == 0 && co0 + co < a.Co && ci0 + ci < a.Ci && t0 + t < a.T)
            atomicAdd(a.gw + ((int64_t)(co0 + co) * a.Ci + ci0 + ci) * a.T + t0 + t, v);
    }
    if (do_bias) {
#pragma unroll
        for (int i = 0; i < CO_B; ++i) {
            const float v = warp_sum(accb[i]);
            if (lane == 0 && co0 + i < a.Co) atomicAdd(a.gbias + co0 + i, v);
        }
    }
}

// ---------------------------------------------------------------- weight gradient, 3-D / 4-D, 8 input channels
// gw[co][ci][t] += sum_{b, s} gpre[b][co][s] in[b][ci][nbr(s, t)]  for the ConvNd / Conv4d layers (convNd.py:84-127,
// adjoint) with 3^D taps.  The generic kernel above spends ~100 index instructions per (site, tap) on periodic
// neighbours and re-reads the data once per (output block, tap block); it was 74 % of a 3-D training step.  Here a
// persistent CTA stages a strip of one (y, x) plane of a sample -- the input with its one-site halo in every
// direction (3^(D-2) planes x 8 channels, periodic wrap resolved while loading) and the strip of gpre for a block of
// output channels -- in shared memory; THREAD (t, ci) owns tap t and input channel ci for good, so its input address
// is a constant offset plus the site index, keeps CO_B accumulators per output block in registers across all the
// units it walks, and issues one atomic per weight at the end.  Per four sites: 4 LDS (conflict-free: the channel
// plane stride is 4 mod 32, taps of a warp differ by 1) + CO_B broadcast LDS.128 + 4 CO_B FMAs.
__device__ __forceinline__ void wg_cp_async4(float* dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void wg_cp_async16(float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

struct WgNdArgs {
    const float* in;        // [B][8][V]
    const float* gpre;      // [B][Co][V]
    float* gw;              // [Co][8][T]
    float* gbias;           // [Co] or NULL
    int D, Co, T, nplanes;  // nplanes = 3^(D-2)
    int O0, O1, Y, X;       // lattice extents, right-aligned (O0 = O1 = 1 in 2-D, O0 = 1 in 3-D)
    int R;                  // rows of a strip (divides Y)
    long long B;
};

template <int CO_B, int NCB>
__global__ void __launch_bounds__(672) conv_wgrad_nd_tile_kernel(const WgNdArgs a) {
    extern __shared__ __align__(16) float wsm[];
    const int XS = a.X + 2, PSZ = (a.R + 2) * XS;            // padded row / channel-plane size of the input strip
    float* in_s = wsm;                                        // [nplanes][8][R + 2][X + 2]
    float* g_s = wsm + a.nplanes * 8 * PSZ;                   // [CO_B][R][X]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5;
    const int t = tid >> 3, ci = tid & 7;
    const bool worker = t < a.T;
    const int pt = t / 9, ky = (t / 3) % 3, kx = t % 3;
    const int base = (pt * 8 + ci) * PSZ + ky * XS + kx;
    const int V = a.O0 * a.O1 * a.Y * a.X, PL = a.Y * a.X;
    const int strips = a.Y / a.R;
    const long long units = a.B * a.O0 * a.O1 * strips;
    float acc[NCB][CO_B];
    float bsum[NCB][2];                  // bias sums of channels wid and wid + nwarps of a block (CO_B <= 2 nwarps)
#pragma unroll
    for (int cb = 0; cb < NCB; ++cb) {
        bsum[cb][0] = bsum[cb][1] = 0.f;
#pragma unroll
        for (int c = 0; c < CO_B; ++c) acc[cb][c] = 0.f;
    }
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
        long long rem = u;
        const int strip = (int)(rem % strips); rem /= strips;
        const int o1 = (int)(rem % a.O1); rem /= a.O1;
        const int o0 = (int)(rem % a.O0);
        const long long b = rem / a.O0;
        const int y0 = strip * a.R;
        __syncthreads();                                      // the previous unit's readers are done
        // ---- input strip with halo: in_s[p][c][j][i] = in[b][c][o + dp][y0 + j - 1][i - 1] (periodic)
        const int total_in = a.nplanes * 8 * PSZ;
        for (int e = tid; e < total_in; e += blockDim.x) {
            const int i = e % XS;
            int r = e / XS;
            const int j = r % (a.R + 2); r /= (a.R + 2);
            const int c = r & 7, p = r >> 3;
            int q0 = o0, q1 = o1;
            if (a.D == 4) { q0 += p / 3 - 1; q1 += p % 3 - 1; }
            else if (a.D == 3) { q1 += p - 1; }
            q0 += q0 < 0 ? a.O0 : 0; q0 -= q0 >= a.O0 ? a.O0 : 0;
            q1 += q1 < 0 ? a.O1 : 0; q1 -= q1 >= a.O1 ? a.O1 : 0;
            int yy = y0 + j - 1, xx = i - 1;
            yy += yy < 0 ? a.Y : 0; yy -= yy >= a.Y ? a.Y : 0;
            xx += xx < 0 ? a.X : 0; xx -= xx >= a.X ? a.X : 0;
            // asynchronous 4-byte copies: with seven warps per SM a register-staged load would pay the memory latency
            // once per element (the kernel was bound by exactly that: 200 k of its 560 k cycles per unit)
            wg_cp_async4(in_s + e, a.in + (b * 8 + c) * (long long)V + ((long long)(q0 * a.O1 + q1) * a.Y + yy) * a.X + xx);
        }
        const long long plane0 = ((long long)(o0 * a.O1 + o1) * a.Y + y0) * a.X;       // first site of the strip
#pragma unroll
        for (int cb = 0; cb < NCB; ++cb) {
            if (cb * CO_B >= a.Co) break;
            if (cb > 0) __syncthreads();                      // g_s readers of the previous block are done
            const int rowq = (a.R * a.X) >> 2, total_g4 = CO_B * rowq;      // 16-byte chunks
            for (int e = tid; e < total_g4; e += blockDim.x) {
                const int c = e / rowq, k = (e - c * rowq) << 2;
                const int co = cb * CO_B + c;
                float* dst = g_s + c * a.R * a.X + k;
                if (co < a.Co) wg_cp_async16(dst, a.gpre + (b * a.Co + co) * (long long)V + plane0 + k);
                else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();
            if (a.gbias) {                                    // bias gradient: warp w sums channel w (+ nwarps, ...)
#pragma unroll
                for (int k2 = 0; k2 < 2; ++k2) {
                    const int c = wid + k2 * nwarps;
                    if (c < CO_B) {
                        float v = 0.f;
                        for (int k = lane; k < a.R * a.X; k += 32) v += g_s[c * a.R * a.X + k];
                        bsum[cb][k2] += warp_sum(v);
                    }
                }
            }
            if (worker) {
                const int RX = a.R * a.X;
                for (int y = 0; y < a.R; ++y) {
                    const float* ip = in_s + base + y * XS;
                    const float* gp = g_s + y * a.X;
#pragma unroll 2
                    for (int x4 = 0; x4 < a.X; x4 += 4) {
                        const float i0 = ip[x4], i1 = ip[x4 + 1], i2 = ip[x4 + 2], i3 = ip[x4 + 3];
#pragma unroll
                        for (int c = 0; c < CO_B; ++c) {
                            const float4 gv = *reinterpret_cast<const float4*>(gp + c * RX + x4);
                            acc[cb][c] = fmaf(gv.x, i0, fmaf(gv.y, i1, fmaf(gv.z, i2, fmaf(gv.w, i3, acc[cb][c]))));
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int cb = 0; cb < NCB; ++cb) {
#pragma unroll
        for (int c = 0; c < CO_B; ++c) {
            const int co = cb * CO_B + c;
            if (worker && co < a.Co) atomicAdd(a.gw + ((long long)co * 8 + ci) * a.T + t, acc[cb][c]);
        }
#pragma unroll
        for (int k2 = 0; k2 < 2; ++k2) {
            const int c = wid + k2 * nwarps;
            if (a.gbias && lane == 0 && c < CO_B && cb * CO_B + c < a.Co) atomicAdd(a.gbias + cb * CO_B + c, bsum[cb][k2]);
        }
    }
}

// First layer of a conditioner (one input channel, optionally seen through the checkerboard mask; Co = 8 per block):
// the same staging, THREAD (t, co) owns weight (co, t).  Per four sites: 4 LDS of the input (the eight lanes of a tap
// read the same words) + one LDS.128 of gpre (channel stride padded by 4 floats: the eight channels of a tap sit on
// different banks) + 4 FMAs.
struct WgNd1Args {
    const float* in;        // [B][1][V]
    const uint8_t* in_mask; // [V] or NULL
    int in_keep;
    const float* gpre;      // [B][Co][V]
    float* gw;              // [Co][1][T]
    float* gbias;
    int D, Co, T, nplanes;
    int O0, O1, Y, X, R;
    long long B;
};

__global__ void __launch_bounds__(672) conv_wgrad_nd_first_kernel(const WgNd1Args a) {
    extern __shared__ __align__(16) float wsm[];
    const int XS = a.X + 2, PSZ = (a.R + 2) * XS, RX = a.R * a.X, GS = RX + 4;
    float* in_s = wsm;                                        // [nplanes][R + 2][X + 2]
    float* g_s = wsm + ((a.nplanes * PSZ + 3) & ~3);          // [8][R X + 4]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5;
    const int t = tid >> 3, co = tid & 7;
    const bool worker = t < a.T;
    const int pt = t / 9, ky = (t / 3) % 3, kx = t % 3;
    const int base = pt * PSZ + ky * XS + kx;
    const int V = a.O0 * a.O1 * a.Y * a.X;
    const int strips = a.Y / a.R;
    const long long units = a.B * a.O0 * a.O1 * strips;
    const int nblocks = (a.Co + 7) >> 3;
    for (int cb = 0; cb < nblocks; ++cb) {                    // (wide first layers: one pass over the data per 8 channels)
        float acc = 0.f, bsum = 0.f;
        for (long long u = blockIdx.x; u < units; u += gridDim.x) {
            long long rem = u;
            const int strip = (int)(rem % strips); rem /= strips;
            const int o1 = (int)(rem % a.O1); rem /= a.O1;
            const int o0 = (int)(rem % a.O0);
            const long long b = rem / a.O0;
            const int y0 = strip * a.R;
            __syncthreads();
            for (int e = tid; e < a.nplanes * PSZ; e += blockDim.x) {
                const int i = e % XS;
                int r = e / XS;
                const int j = r % (a.R + 2), p = r / (a.R + 2);
                int q0 = o0, q1 = o1;
                if (a.D == 4) { q0 += p / 3 - 1; q1 += p % 3 - 1; }
                else if (a.D == 3) { q1 += p - 1; }
                q0 += q0 < 0 ? a.O0 : 0; q0 -= q0 >= a.O0 ? a.O0 : 0;
                q1 += q1 < 0 ? a.O1 : 0; q1 -= q1 >= a.O1 ? a.O1 : 0;
                int yy = y0 + j - 1, xx = i - 1;
                yy += yy < 0 ? a.Y : 0; yy -= yy >= a.Y ? a.Y : 0;
                xx += xx < 0 ? a.X : 0; xx -= xx >= a.X ? a.X : 0;
                const int n = ((q0 * a.O1 + q1) * a.Y + yy) * a.X + xx;
                const bool keep = !a.in_mask || __ldg(a.in_mask + n) == (uint8_t)a.in_keep;
                if (keep) wg_cp_async4(in_s + e, a.in + b * (long long)V + n);
                else in_s[e] = 0.f;
            }
            const long long plane0 = ((long long)(o0 * a.O1 + o1) * a.Y + y0) * a.X;
            const int rowq = RX >> 2;
            for (int e = tid; e < 8 * rowq; e += blockDim.x) {
                const int c = e / rowq, k = (e - c * rowq) << 2;
                const int cc = cb * 8 + c;
                float* dst = g_s + c * GS + k;
                if (cc < a.Co) wg_cp_async16(dst, a.gpre + (b * a.Co + cc) * (long long)V + plane0 + k);
                else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();
            if (a.gbias && wid < 8) {                         // nwarps >= 7: channel 7 falls to warp 0 as well
                for (int c = wid; c < 8; c += nwarps) {
                    float v = 0.f;
                    for (int k = lane; k < RX; k += 32) v += g_s[c * GS + k];
                    v = warp_sum(v);
                    if (c == wid) bsum += v;
                    else if (lane == 0 && cb * 8 + c < a.Co) atomicAdd(a.gbias + cb * 8 + c, v);
                }
            }
            if (worker) {
                for (int y = 0; y < a.R; ++y) {
                    const float* ip = in_s + base + y * XS;
                    const float* gp = g_s + co * GS + y * a.X;
#pragma unroll 2
                    for (int x4 = 0; x4 < a.X; x4 += 4) {
                        const float4 gv = *reinterpret_cast<const float4*>(gp + x4);
                        acc = fmaf(gv.x, ip[x4], fmaf(gv.y, ip[x4 + 1], fmaf(gv.z, ip[x4 + 2], fmaf(gv.w, ip[x4 + 3], acc))));
                    }
                }
            }
        }
        if (worker && cb * 8 + co < a.Co) atomicAdd(a.gw + (long long)(cb * 8 + co) * a.T + t, acc);
        if (a.gbias && lane == 0 && wid < 8 && cb * 8 + wid < a.Co) atomicAdd(a.gbias + cb * 8 + wid, bsum);
    }
}

static int wgrad_nd_first_launch(WgNd1Args a, cudaStream_t st) {
    int dev = 0, sm = 148, max_smem = 227 * 1024;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    auto need = [&](int R) {
        return ((((size_t)a.nplanes * (R + 2) * (a.X + 2) + 3) & ~(size_t)3) + (size_t)8 * (R * a.X + 4)) * 4;
    };
    int best = 0;
    for (int R = 1; R <= a.Y; ++R)
        if (a.Y % R == 0 && need(R) <= ((size_t)max_smem + 1024) / 2 - 1024) best = R;
    if (best == 0) return NFK_EUNSUPPORTED;
    a.R = best;
    if (ensure_dynamic_smem<conv_wgrad_nd_first_kernel>(max_smem) != NFK_OK) return NFK_ECUDA;
    const long long units = a.B * a.O0 * a.O1 * (a.Y / best);
    const int threads = (a.T * 8 + 31) / 32 * 32;
    const long long ctas = (long long)sm * (threads <= 512 ? 2 : 1);
    const long long grid = units < ctas ? units : ctas;
    conv_wgrad_nd_first_kernel<<<(unsigned)grid, threads, need(best), st>>>(a);
    return check_launch();
}

template <int CO_B, int NCB>
static int wgrad_nd_tile_launch(WgNdArgs a, cudaStream_t st) {
    int dev = 0, sm = 148, max_smem = 227 * 1024;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    // the tallest strip that still lets two CTAs share an SM (seven warps alone cannot hide the shared-memory
    // latency of the inner loop); the tallest that fits at all otherwise
    int best = 0, best2 = 0;
    for (int R = 1; R <= a.Y; ++R) {
        if (a.Y % R) continue;
        const size_t need = ((size_t)a.nplanes * 8 * (R + 2) * (a.X + 2) + (size_t)CO_B * R * a.X) * 4;
        if (need <= (size_t)max_smem - 1024) best = R;
        if (need <= ((size_t)max_smem + 1024) / 2 - 1024 && a.T * 8 <= 512) best2 = R;
    }
    if (best2 >= 4) best = best2;
    if (best == 0) return NFK_EUNSUPPORTED;
    a.R = best;
    const size_t smem = ((size_t)a.nplanes * 8 * (best + 2) * (a.X + 2) + (size_t)CO_B * best * a.X) * 4;
    if (ensure_dynamic_smem<conv_wgrad_nd_tile_kernel<CO_B, NCB>>(max_smem) != NFK_OK) return NFK_ECUDA;
    const long long units = a.B * a.O0 * a.O1 * (a.Y / best);
    const int threads = (a.T * 8 + 31) / 32 * 32;
    const long long ctas = (long long)sm * (best == best2 ? 2 : 1);
    const long long grid = units < ctas ? units : ctas;
    conv_wgrad_nd_tile_kernel<CO_B, NCB><<<(unsigned)grid, threads, smem, st>>>(a);
    return check_launch();
}

// ---------------------------------------------------------------- weight gradient, 2-D 3x3
// The ConvAct layers of the BASELINE configs: 2-D lattice, 3x3 taps, Ci <= 8.  A persistent CTA
// (blockIdx.y = block of CO_B output channels) walks samples in strips of R rows held in shared
// memory -- the input strip with its periodic halo, the strip of d loss / d pre-activation -- and
// each of its nine warps owns ONE tap: lanes walk 32 consecutive columns, a lane keeps the
// CO_B x CI block of the weight gradient of its tap in registers (56 FMAs per 15 shared loads,
// all conflict-free), summed over every site the lane visits in the whole launch; one shuffle
// reduction and one atomicAdd per (CTA, weight) at the very end.
struct Wgrad2dArgs {
    const float* in;
    const uint8_t* in_mask;
    int in_keep;
    const float* gpre;
    float* gw;
    float* gbias;
    int L0, L1, R, Ci, Co;
    long long B;
    int g_parity;             // SPARSE: gpre vanishes on sites with (row + col) % 2 != g_parity
    int G;                    // small-lattice kernel: samples staged together per pipeline step
};

// SPARSE: the gradient handed in is that of a checkerboard coupling's conditioner output, which
// is non-zero on ONE partition only; lanes then walk the sites of that partition (every other
// column, the offset alternating with the row), halving the FMAs at the price of 2-way bank
// conflicts on the shared loads.
template <int CI, int CO_B, bool SPARSE>
__global__ void __launch_bounds__(288, 2) conv2d_wgrad_kernel(Wgrad2dArgs a) {
    extern __shared__ float sm[];
    const int L0 = a.L0, L1 = a.L1, R = a.R, LW = L1 + 2;
    float* in_s = sm;                               // [CI][R + 2][LW]
    float* g_s = sm + CI * (R + 2) * LW;            // [CO_B][R][L1]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kh = warp / 3, kw = warp % 3;         // this warp's tap
    const int co0 = blockIdx.y * CO_B;
    const int V = L0 * L1;
    float acc[CO_B][CI];
    float accb[CO_B];
#pragma unroll
    for (int co = 0; co < CO_B; ++co) {
        accb[co] = 0.f;
#pragma unroll
        for (int ci = 0; ci < CI; ++ci) acc[co][ci] = 0.f;
    }
    for (long long b = blockIdx.x; b < a.B; b += gridDim.x) {
        const float* in_b = a.in + b * (long long)CI * V;
        const float* g_b = a.gpre + (b * a.Co + co0) * (long long)V;
        for (int r0 = 0; r0 < L0; r0 += R) {
            const int rows = L0 - r0 < R ? L0 - r0 : R;
            __syncthreads();                        // previous strip consumed
            if (L1 < 32) {
                // short rows: every element is one item of a flat list spread over the CTA
                for (int it = tid; it < CI * (rows + 2) * LW; it += 288) {
                    const int p = it / LW, k = it - p * LW;
                    const int ci = p / (rows + 2), j = p - ci * (rows + 2);
                    int r = r0 - 1 + j;
                    r = r < 0 ? r + L0 : (r >= L0 ? r - L0 : r);
                    int c = k - 1;
                    c = c < 0 ? c + L1 : (c >= L1 ? c - L1 : c);
                    float v = __ldg(in_b + ci * V + r * L1 + c);
                    if (a.in_mask && __ldg(a.in_mask + r * L1 + c) != (uint8_t)a.in_keep) v = 0.f;
                    in_s[(ci * (R + 2) + j) * LW + k] = v;
                }
                for (int it = tid; it < CO_B * rows * L1; it += 288) {
                    const int p = it / L1, c = it - p * L1;
                    const int co = p / rows, j = p - co * rows;
                    g_s[(co * R + j) * L1 + c] = co0 + co < a.Co ? __ldg(g_b + (long long)co * V + (r0 + j) * L1 + c) : 0.f;
                }
            } else {
                // staging: one warp per (channel, row) line, lanes over the columns
                for (int p = warp; p < CI * (rows + 2); p += 9) {
                    const int ci = p / (rows + 2), j = p - ci * (rows + 2);
                    int r = r0 - 1 + j;
                    r = r < 0 ? r + L0 : (r >= L0 ? r - L0 : r);
                    const float* src = in_b + ci * V + r * L1;
                    const uint8_t* msk = a.in_mask ? a.in_mask + r * L1 : nullptr;
                    float* dst = in_s + (ci * (R + 2) + j) * LW;
                    for (int k = lane; k < LW; k += 32) {
                        int c = k - 1;
                        c = c < 0 ? c + L1 : (c >= L1 ? c - L1 : c);
                        float v = __ldg(src + c);
                        if (msk && __ldg(msk + c) != (uint8_t)a.in_keep) v = 0.f;
                        dst[k] = v;
                    }
                }
                for (int p = warp; p < CO_B * rows; p += 9) {
                    const int co = p / rows, j = p - co * rows;
                    const float* src = g_b + (long long)co * V + (r0 + j) * L1;
                    float* dst = g_s + (co * R + j) * L1;
                    const bool live = co0 + co < a.Co;
                    for (int c = lane; c < L1; c += 32) dst[c] = live ? __ldg(src + c) : 0.f;
                }
            }
            __syncthreads();
            auto site = [&](int j, int cc) {
                const int c = SPARSE ? 2 * cc + ((a.g_parity + r0 + j) & 1) : cc;
                if (SPARSE && c >= L1) return;
                float gv[CO_B], xv[CI];
#pragma unroll
                for (int co = 0; co < CO_B; ++co) gv[co] = g_s[(co * R + j) * L1 + c];
#pragma unroll
                for (int ci = 0; ci < CI; ++ci) xv[ci] = in_s[(ci * (R + 2) + j + kh) * LW + c + kw];
#pragma unroll
                for (int co = 0; co < CO_B; ++co) {
#pragma unroll
                    for (int ci = 0; ci < CI; ++ci) acc[co][ci] = fmaf(gv[co], xv[ci], acc[co][ci]);
                }
                if (warp == 4) {                // the centre-tap warp also sums the bias gradient
#pragma unroll
                    for (int co = 0; co < CO_B; ++co) accb[co] += gv[co];
                }
            };
            const int ncols = SPARSE ? (L1 + 1) / 2 : L1;
            if (L1 < 32) {                      // short rows: lanes walk the flattened (row, column) sites
                for (int t = lane; t < rows * ncols; t += 32) site(t / ncols, t % ncols);
            } else {
                for (int j = 0; j < rows; ++j)
                    for (int cc = lane; cc < ncols; cc += 32) site(j, cc);
            }
        }
    }
#pragma unroll
    for (int co = 0; co < CO_B; ++co) {
#pragma unroll
        for (int ci = 0; ci < CI; ++ci) {
            const float v = warp_sum(acc[co][ci]);
            if (lane == 0 && co0 + co < a.Co)
                atomicAdd(a.gw + ((long long)(co0 + co) * CI + ci) * 9 + kh * 3 + kw, v);
        }
        if (warp == 4 && a.gbias) {
            const float v = warp_sum(accb[co]);
            if (lane == 0 && co0 + co < a.Co) atomicAdd(a.gbias + co0 + co, v);
        }
    }
}

template <int CI, int CO_B, bool SPARSE>
static int wgrad2d_launch(Wgrad2dArgs a, cudaStream_t st) {
    const int LW = a.L1 + 2;
    int R = a.L0 < 16 ? a.L0 : 16;
    auto bytes = [&](int r) { return (size_t)(CI * (r + 2) * LW + CO_B * r * a.L1) * sizeof(float); };
    while (R > 1 && bytes(R) > 72 * 1024) R /= 2;
    if (bytes(R) > 200 * 1024) return NFK_EUNSUPPORTED;
    a.R = R;
    if (ensure_dynamic_smem<conv2d_wgrad_kernel<CI, CO_B, SPARSE>>(200 * 1024) != NFK_OK) return NFK_ECUDA;
    const int ncb = (a.Co + CO_B - 1) / CO_B;
    long long gx = (148LL * 2 + ncb - 1) / ncb;            // two CTAs per SM in all
    if (gx > a.B) gx = a.B;
    if (gx < 1) gx = 1;
    conv2d_wgrad_kernel<CI, CO_B, SPARSE><<<dim3((unsigned)gx, ncb), 288, bytes(R), st>>>(a);
    return check_launch();
}

// The same kernel with the strips streamed by cp.async into TWO shared-memory buffers: while the
// warps accumulate strip i, the copies of strip i+1 are in flight (the synchronous version spends
// more than half of its time staging).  Needs 16-byte aligned rows (L1 % 4 == 0) and no input mask.
//   in_s row: [.. 3 pad | col -1 | cols 0..L1-1 (16-byte aligned) | col L1 | pad ..], stride L1 + 8
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool live) {
    const int sz = live ? 16 : 0;                    // src-size 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst), "l"(src) : "memory");
}

// L1C, RC > 0: row length and rows per strip fixed at compile time, which turns every shared-memory address
// of the inner loop into base + immediate (with run-time strides the kernel spent 40 % of its instructions on
// integer address arithmetic and spilled its accumulators; ncu, 64x64).
template <int CI, int CO_B, bool SPARSE, int L1C = 0, int RC = 0>
__global__ void __launch_bounds__(288, 2) conv2d_wgrad_async_kernel(Wgrad2dArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int L0 = a.L0, L1 = L1C > 0 ? L1C : a.L1, R = RC > 0 ? RC : a.R, LW = L1 + 8;
    const int in_floats = CI * (R + 2) * LW, buf_floats = in_floats + CO_B * R * L1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kh = warp / 3, kw = warp % 3;
    const int co0 = blockIdx.y * CO_B;
    const int V = L0 * L1, nq = L1 >> 2;
    const int strips = (L0 + R - 1) / R;
    // One input channel: a tap per warp would mean 9 shared-memory loads per 8 FMAs.  Instead every warp takes
    // rows of the strip and keeps all 9 taps (8 x 9 accumulators per lane): 17 loads per 72 FMAs.
    constexpr bool kRowWarps = (CI == 1 && !SPARSE);
    float acc[CO_B][kRowWarps ? 9 : CI];
    float accb[CO_B];
#pragma unroll
    for (int co = 0; co < CO_B; ++co) {
        accb[co] = 0.f;
#pragma unroll
        for (int ci = 0; ci < (kRowWarps ? 9 : CI); ++ci) acc[co][ci] = 0.f;
    }
    const long long n_mine = a.B > blockIdx.x ? (a.B - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long units = n_mine * strips;

    auto stage = [&](long long u, float* buf) {
        const long long b = blockIdx.x + (u / strips) * gridDim.x;
        const int r0 = (int)(u % strips) * R;
        const int rows = L0 - r0 < R ? L0 - r0 : R;
        const float* in_b = a.in + b * (long long)CI * V;
        const float* g_b = a.gpre + (b * a.Co + co0) * (long long)V;
        const uint32_t in_u = (uint32_t)__cvta_generic_to_shared(buf), g_u = in_u + in_floats * 4;
        for (int p = warp; p < CI * (rows + 2); p += 9) {
            const int ci = p / (rows + 2), j = p - ci * (rows + 2);
            int r = r0 - 1 + j;
            r = r < 0 ? r + L0 : (r >= L0 ? r - L0 : r);
            const float* src = in_b + ci * V + r * L1;
            const uint32_t dst = in_u + (uint32_t)((ci * (R + 2) + j) * LW) * 4;
            for (int q = lane; q < nq; q += 32) cp_async16(dst + 16 + q * 16, src + q * 4, true);
            if (lane == 0) cp_async4(dst + 12, src + L1 - 1);             // column -1
            if (lane == 1) cp_async4(dst + 16 + L1 * 4, src);             // column L1
        }
        for (int p = warp; p < CO_B * rows; p += 9) {
            const int co = p / rows, j = p - co * rows;
            const bool live = co0 + co < a.Co;
            const float* src = g_b + (long long)(live ? co : 0) * V + (r0 + j) * L1;
            const uint32_t dst = g_u + (uint32_t)((co * R + j) * L1) * 4;
            for (int q = lane; q < nq; q += 32) cp_async16(dst + q * 16, src + q * 4, live);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    if (units > 0) stage(0, sm);
    for (long long u = 0; u < units; ++u) {
        float* buf = sm + (u & 1) * buf_floats;
        if (u + 1 < units) {
            stage(u + 1, sm + ((u + 1) & 1) * buf_floats);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();                                                  // strip u has landed for everyone
        const float* in_s = buf;
        const float* g_s = buf + in_floats;
        const int r0 = (int)(u % strips) * R;
        const int rows = L0 - r0 < R ? L0 - r0 : R;
        if constexpr (kRowWarps) {
            for (int j = warp; j < rows; j += 9)
                for (int c = lane; c < L1; c += 32) {
                    float gv[CO_B], xv[9];
#pragma unroll
                    for (int co = 0; co < CO_B; ++co) gv[co] = g_s[(co * R + j) * L1 + c];
#pragma unroll
                    for (int t = 0; t < 9; ++t) xv[t] = in_s[(j + t / 3) * LW + 3 + c + t % 3];
#pragma unroll
                    for (int co = 0; co < CO_B; ++co) {
#pragma unroll
                        for (int t = 0; t < 9; ++t) acc[co][t] = fmaf(gv[co], xv[t], acc[co][t]);
                        accb[co] += gv[co];
                    }
                }
        } else {
            for (int j = 0; j < rows; ++j)
                for (int cc = lane; cc < (SPARSE ? L1 / 2 : L1); cc += 32) {
                    const int c = SPARSE ? 2 * cc + ((a.g_parity + r0 + j) & 1) : cc;
                    float gv[CO_B], xv[CI];
#pragma unroll
                    for (int co = 0; co < CO_B; ++co) gv[co] = g_s[(co * R + j) * L1 + c];
#pragma unroll
                    for (int ci = 0; ci < CI; ++ci) xv[ci] = in_s[(ci * (R + 2) + j + kh) * LW + 3 + c + kw];
#pragma unroll
                    for (int co = 0; co < CO_B; ++co) {
#pragma unroll
                        for (int ci = 0; ci < CI; ++ci) acc[co][ci] = fmaf(gv[co], xv[ci], acc[co][ci]);
                    }
                    if (warp == 4) {
#pragma unroll
                        for (int co = 0; co < CO_B; ++co) accb[co] += gv[co];
                    }
                }
        }
        __syncthreads();                                                  // buffer free for the copies of strip u + 2
    }
    if constexpr (kRowWarps) {
#pragma unroll
        for (int co = 0; co < CO_B; ++co) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float v = warp_sum(acc[co][t]);
                if (lane == 0 && co0 + co < a.Co) atomicAdd(a.gw + (long long)(co0 + co) * 9 + t, v);
            }
            if (a.gbias) {
                const float v = warp_sum(accb[co]);
                if (lane == 0 && co0 + co < a.Co) atomicAdd(a.gbias + co0 + co, v);
            }
        }
    } else {
#pragma unroll
        for (int co = 0; co < CO_B; ++co) {
#pragma unroll
            for (int ci = 0; ci < CI; ++ci) {
                const float v = warp_sum(acc[co][ci]);
                if (lane == 0 && co0 + co < a.Co)
                    atomicAdd(a.gw + ((long long)(co0 + co) * CI + ci) * 9 + kh * 3 + kw, v);
            }
            if (warp == 4 && a.gbias) {
                const float v = warp_sum(accb[co]);
                if (lane == 0 && co0 + co < a.Co) atomicAdd(a.gbias + co0 + co, v);
            }
        }
    }
}

template <int CI, int CO_B, bool SPARSE>
static int wgrad2d_async_launch(Wgrad2dArgs a, cudaStream_t st) {
    const int LW = a.L1 + 8;
    auto bytes = [&](int r) { return 2 * (size_t)(CI * (r + 2) * LW + CO_B * r * a.L1) * sizeof(float); };
    int R = a.L0 < 16 ? a.L0 : 16;
    while (R > 1 && bytes(R) > 100 * 1024) R /= 2;                        // two CTAs per SM
    if (bytes(R) > 100 * 1024) return NFK_EUNSUPPORTED;
    a.R = R;
    const int ncb = (a.Co + CO_B - 1) / CO_B;
    long long gx = (148LL * 2 + ncb - 1) / ncb;
    if (gx > a.B) gx = a.B;
    if (gx < 1) gx = 1;
    const dim3 grid((unsigned)gx, ncb);
    if (a.L1 == 64 && R == 8) {                                           // the benchmark geometry: compile-time strides
        if (ensure_dynamic_smem<conv2d_wgrad_async_kernel<CI, CO_B, SPARSE, 64, 8>>(100 * 1024) != NFK_OK) return NFK_ECUDA;
        conv2d_wgrad_async_kernel<CI, CO_B, SPARSE, 64, 8><<<grid, 288, bytes(R), st>>>(a);
    } else if (a.L1 == 64 && R == 16) {                                   // (one input channel)
        if (ensure_dynamic_smem<conv2d_wgrad_async_kernel<CI, CO_B, SPARSE, 64, 16>>(100 * 1024) != NFK_OK) return NFK_ECUDA;
        conv2d_wgrad_async_kernel<CI, CO_B, SPARSE, 64, 16><<<grid, 288, bytes(R), st>>>(a);
    } else if (a.L1 == 32 && R == 16) {
        if (ensure_dynamic_smem<conv2d_wgrad_async_kernel<CI, CO_B, SPARSE, 32, 16>>(100 * 1024) != NFK_OK) return NFK_ECUDA;
        conv2d_wgrad_async_kernel<CI, CO_B, SPARSE, 32, 16><<<grid, 288, bytes(R), st>>>(a);
    } else {
        if (ensure_dynamic_smem<conv2d_wgrad_async_kernel<CI, CO_B, SPARSE>>(100 * 1024) != NFK_OK) return NFK_ECUDA;
        conv2d_wgrad_async_kernel<CI, CO_B, SPARSE><<<grid, 288, bytes(R), st>>>(a);
    }
    return check_launch();
}

// Small lattices (a whole sample is one short strip, rows shorter than a warp): the kernel above
// leaves most lanes idle (lanes map to columns) and pays two barriers per sample.  Here G samples are
// staged per pipeline step, every copy is one item of a flat list spread over the CTA, and the lanes of
// a tap's warp walk the flattened (sample, row, column) sites of the group.
template <int CI, int CO_B, bool SPARSE>
__global__ void __launch_bounds__(288, 2) conv2d_wgrad_small_kernel(Wgrad2dArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int L0 = a.L0, L1 = a.L1, LW = L1 + 8, G = a.G;
    const int in_floats = CI * (L0 + 2) * LW, slot_floats = in_floats + CO_B * L0 * L1;
    const int buf_floats = G * slot_floats;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kh = warp / 3, kw = warp % 3;
    const int co0 = blockIdx.y * CO_B;
    const int V = L0 * L1, nq = L1 >> 2;
    float acc[CO_B][CI];
    float accb[CO_B];
#pragma unroll
    for (int co = 0; co < CO_B; ++co) {
        accb[co] = 0.f;
#pragma unroll
        for (int ci = 0; ci < CI; ++ci) acc[co][ci] = 0.f;
    }
    const long long n_mine = a.B > blockIdx.x ? (a.B - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long units = (n_mine + G - 1) / G;
    const int in_lines = CI * (L0 + 2), g_lines = CO_B * L0;
    const int in_items = in_lines * (nq + 2), g_items = g_lines * nq;

    auto stage = [&](long long u, float* buf) {
        const long long k0 = u * G;
        const int ns = (int)(n_mine - k0 < G ? n_mine - k0 : G);
        for (int it = tid; it < ns * in_items; it += 288) {
            const int s = it / in_items, r1 = it - s * in_items;
            const int line = r1 / (nq + 2), q = r1 - line * (nq + 2);
            const int ci = line / (L0 + 2), j = line - ci * (L0 + 2);
            int r = j - 1;
            r = r < 0 ? r + L0 : (r >= L0 ? r - L0 : r);
            const long long b = blockIdx.x + (k0 + s) * gridDim.x;
            const float* src = a.in + (b * CI + ci) * (long long)V + r * L1;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(buf + s * slot_floats + line * LW);
            if (q < nq) cp_async16(dst + 16 + q * 16, src + q * 4, true);
            else if (q == nq) cp_async4(dst + 12, src + L1 - 1);          // column -1
            else cp_async4(dst + 16 + L1 * 4, src);                       // column L1
        }
        for (int it = tid; it < ns * g_items; it += 288) {
            const int s = it / g_items, r1 = it - s * g_items;
            const int line = r1 / nq, q = r1 - line * nq;
            const int co = line / L0, j = line - co * L0;
            const bool live = co0 + co < a.Co;
            const long long b = blockIdx.x + (k0 + s) * gridDim.x;
            const float* src = a.gpre + (b * a.Co + co0 + (live ? co : 0)) * (long long)V + j * L1;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(buf + s * slot_floats + in_floats + line * L1);
            cp_async16(dst + q * 16, src + q * 4, live);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int ncols = SPARSE ? L1 / 2 : L1;
    const int per_sample = L0 * ncols;
    if (units > 0) stage(0, sm);
    for (long long u = 0; u < units; ++u) {
        float* buf = sm + (u & 1) * buf_floats;
        if (u + 1 < units) {
            stage(u + 1, sm + ((u + 1) & 1) * buf_floats);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const long long k0 = u * G;
        const int ns = (int)(n_mine - k0 < G ? n_mine - k0 : G);
        for (int t = lane; t < ns * per_sample; t += 32) {
            const int s = t / per_sample, r1 = t - s * per_sample;
            const int j = r1 / ncols, cc = r1 - j * ncols;
            const int c = SPARSE ? 2 * cc + ((a.g_parity + j) & 1) : cc;
            const float* in_s = buf + s * slot_floats;
            const float* g_s = in_s + in_floats;
            float gv[CO_B], xv[CI];
#pragma unroll
            for (int co = 0; co < CO_B; ++co) gv[co] = g_s[(co * L0 + j) * L1 + c];
#pragma unroll
            for (int ci = 0; ci < CI; ++ci) xv[ci] = in_s[(ci * (L0 + 2) + j + kh) * LW + 3 + c + kw];
#pragma unroll
            for (int co = 0; co < CO_B; ++co) {
#pragma unroll
                for (int ci = 0; ci < CI; ++ci) acc[co][ci] = fmaf(gv[co], xv[ci], acc[co][ci]);
            }
            if (warp == 4) {
#pragma unroll
                for (int co = 0; co < CO_B; ++co) accb[co] += gv[co];
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int co = 0; co < CO_B; ++co) {
#pragma unroll
        for (int ci = 0; ci < CI; ++ci) {
            const float v = warp_sum(acc[co][ci]);
            if (lane == 0 && co0 + co < a.Co)
                atomicAdd(a.gw + ((long long)(co0 + co) * CI + ci) * 9 + kh * 3 + kw, v);
        }
        if (warp == 4 && a.gbias) {
            const float v = warp_sum(accb[co]);
            if (lane == 0 && co0 + co < a.Co) atomicAdd(a.gbias + co0 + co, v);
        }
    }
}

// applies when one sample is at most 512 sites and its rows are shorter than a warp
static bool wgrad2d_small_ok(const Wgrad2dArgs& a) { return a.L1 < 32 && a.L0 * a.L1 <= 512; }

template <int CI, int CO_B, bool SPARSE>
static int wgrad2d_small_launch(Wgrad2dArgs a, cudaStream_t st) {
    const int LW = a.L1 + 8;
    const size_t slot = (size_t)(CI * (a.L0 + 2) * LW + CO_B * a.L0 * a.L1) * sizeof(float);
    int G = (int)(1024 / (a.L0 * a.L1));                                  // ~1024 sites per step
    if (G < 1) G = 1;
    if (G > 16) G = 16;
    while (G > 1 && 2 * G * slot > 100 * 1024) --G;
    if (2 * G * slot > 100 * 1024) return NFK_EUNSUPPORTED;
    a.G = G;
    if (ensure_dynamic_smem<conv2d_wgrad_small_kernel<CI, CO_B, SPARSE>>(100 * 1024) != NFK_OK) return NFK_ECUDA;
    const int ncb = (a.Co + CO_B - 1) / CO_B;
    long long gx = (148LL * 2 + ncb - 1) / ncb;
    const long long groups = (a.B + G - 1) / G;
    if (gx > groups) gx = groups;
    if (gx < 1) gx = 1;
    conv2d_wgrad_small_kernel<CI, CO_B, SPARSE><<<dim3((unsigned)gx, ncb), 288, 2 * G * slot, st>>>(a);
    return check_launch();
}

static int conv_bwd_weight_impl(const float* in, const uint8_t* in_mask, int in_keep, const float* gpre,
                                int g_parity, float* gw, float* gbias, nfk_lattice lat, int ksize, int Ci, int Co,
                                int64_t B, void* stream);

extern "C" int nfk_conv_circ_bwd_weight(const float* in, const uint8_t* in_mask, int in_keep,
                                        const float* gpre, float* gw, float* gbias,
                                        nfk_lattice lat, int ksize, int Ci, int Co, int64_t B, void* stream) {
    return conv_bwd_weight_impl(in, in_mask, in_keep, gpre, -1, gw, gbias, lat, ksize, Ci, Co, B, stream);
}
extern "C" int nfk_conv_circ_bwd_weight_cb(const float* in, const float* gpre, int g_parity, float* gw, float* gbias,
                                           nfk_lattice lat, int ksize, int Ci, int Co, int64_t B, void* stream) {
    if (g_parity != 0 && g_parity != 1) return NFK_EINVAL;
    return conv_bwd_weight_impl(in, nullptr, 0, gpre, g_parity, gw, gbias, lat, ksize, Ci, Co, B, stream);
}

static int conv_bwd_weight_impl(const float* in, const uint8_t* in_mask, int in_keep, const float* gpre,
                                int g_parity, float* gw, float* gbias, nfk_lattice lat, int ksize, int Ci, int Co,
                                int64_t B, void* stream) {
    if (!in || !gpre || !gw || !lat_ok(lat) || ksize < 1 || ksize % 2 == 0 || Ci < 1 || Co < 1) return NFK_EINVAL;
    if (B <= 0) return NFK_OK;
    if (lat.ndim == 2 && ksize == 3 && (Ci == 1 || Ci == 8) && lat.shape[0] >= 2 && lat.shape[1] >= 2) {
        Wgrad2dArgs w;
        w.in = in; w.in_mask = in_mask; w.in_keep = in_keep; w.gpre = gpre; w.gw = gw; w.gbias = gbias;
        w.L0 = lat.shape[0]; w.L1 = lat.shape[1]; w.R = 0; w.Ci = Ci; w.Co = Co; w.B = B;
        w.g_parity = g_parity;
        // the checkerboard shortcut needs a consistent wrap (even sides); otherwise sum densely
        const bool sparse = g_parity >= 0 && lat.shape[0] % 2 == 0 && lat.shape[1] % 2 == 0;
        int rc = NFK_EUNSUPPORTED;
        const bool async_ok = Ci == 8 && !in_mask && lat.shape[1] % 4 == 0 && ((uintptr_t)in % 16) == 0 &&
                              ((uintptr_t)gpre % 16) == 0;
        if (async_ok && wgrad2d_small_ok(w)) {                // small lattices: groups of samples per step
            if (Co <= 8) rc = sparse ? wgrad2d_small_launch<8, 8, true>(w, NFK_STREAM(stream))
                                     : wgrad2d_small_launch<8, 8, false>(w, NFK_STREAM(stream));
            else rc = sparse ? wgrad2d_small_launch<8, 7, true>(w, NFK_STREAM(stream))
                             : wgrad2d_small_launch<8, 7, false>(w, NFK_STREAM(stream));
            if (rc != NFK_EUNSUPPORTED) return rc;
        }
        if (async_ok) {                                       // strips streamed with cp.async, double-buffered
            if (Co <= 8) rc = sparse ? wgrad2d_async_launch<8, 8, true>(w, NFK_STREAM(stream))
                                     : wgrad2d_async_launch<8, 8, false>(w, NFK_STREAM(stream));
            else rc = sparse ? wgrad2d_async_launch<8, 7, true>(w, NFK_STREAM(stream))
                             : wgrad2d_async_launch<8, 7, false>(w, NFK_STREAM(stream));
            if (rc != NFK_EUNSUPPORTED) return rc;
        }
        // one input channel without a mask (the caller has applied Mask.split): the streamed kernel again
        if (Ci == 1 && !in_mask && Co <= 8 && lat.shape[1] % 4 == 0 && ((uintptr_t)in % 16) == 0 &&
            ((uintptr_t)gpre % 16) == 0) {
            rc = wgrad2d_small_ok(w) ? wgrad2d_small_launch<1, 8, false>(w, NFK_STREAM(stream))
                                     : wgrad2d_async_launch<1, 8, false>(w, NFK_STREAM(stream));
            if (rc != NFK_EUNSUPPORTED) return rc;
        }
        if (Ci == 1) rc = wgrad2d_launch<1, 8, false>(w, NFK_STREAM(stream));
        else if (Co <= 8) rc = sparse ? wgrad2d_launch<8, 8, true>(w, NFK_STREAM(stream))
                                      : wgrad2d_launch<8, 8, false>(w, NFK_STREAM(stream));
        else rc = sparse ? wgrad2d_launch<8, 7, true>(w, NFK_STREAM(stream))
                         : wgrad2d_launch<8, 7, false>(w, NFK_STREAM(stream));
        if (rc != NFK_EUNSUPPORTED) return rc;
    }
    ConvWArgs a;
    a.in = in; a.in_mask = in_mask; a.in_keep = in_keep; a.gpre = gpre; a.gw = gw; a.gbias = gbias;
    a.lat = to_lat(lat); a.ksize = ksize; a.Ci = Ci; a.Co = Co; a.V = (int)lat_volume(lat);
    int T = 1;
    for (int d = 0; d < lat.ndim; ++d) T *= ksize;
    a.T = T;
    a.BV = B * (int64_t)a.V;
    constexpr int TB = 3;
    a.n_tb = (T + TB - 1) / TB;
    cudaStream_t st = NFK_STREAM(stream);
    if ((lat.ndim == 3 || lat.ndim == 4) && ksize == 3 && Ci == 8 && Co <= 28 && !in_mask &&
        lat.shape[lat.ndim - 1] % 4 == 0 && ((uintptr_t)gpre % 16) == 0) {
        WgNdArgs t{};
        t.in = in; t.gpre = gpre; t.gw = gw; t.gbias = gbias;
        t.D = lat.ndim; t.Co = Co; t.T = T; t.nplanes = T / 9; t.B = B;
        t.O0 = lat.ndim == 4 ? lat.shape[0] : 1;
        t.O1 = lat.shape[lat.ndim - 3];
        t.Y = lat.shape[lat.ndim - 2];
        t.X = lat.shape[lat.ndim - 1];
        const int rc = Co <= 8 ? wgrad_nd_tile_launch<8, 1>(t, st) : wgrad_nd_tile_launch<7, 4>(t, st);
        if (rc != NFK_EUNSUPPORTED) return rc;
    }
    if ((lat.ndim == 3 || lat.ndim == 4) && ksize == 3 && Ci == 1 && lat.shape[lat.ndim - 1] % 4 == 0 &&
        ((uintptr_t)gpre % 16) == 0) {
        WgNd1Args t{};
        t.in = in; t.in_mask = in_mask; t.in_keep = in_keep; t.gpre = gpre; t.gw = gw; t.gbias = gbias;
        t.D = lat.ndim; t.Co = Co; t.T = T; t.nplanes = T / 9; t.B = B;
        t.O0 = lat.ndim == 4 ? lat.shape[0] : 1;
        t.O1 = lat.shape[lat.ndim - 3];
        t.Y = lat.shape[lat.ndim - 2];
        t.X = lat.shape[lat.ndim - 1];
        const int rc = wgrad_nd_first_launch(t, st);
        if (rc != NFK_EUNSUPPORTED) return rc;
    }
    if (T >= 27 && Ci >= 8 && a.BV >= 32 * 148) {
        // 3-D / 4-D layers with 8+ input channels: tap blocks dealt to the warps of a CTA (data read once per block
        // of output channels).  9 tap blocks of 3 (27 taps) per CTA; 81 taps take three such rows of CTAs.
        const int tb_per_cta = a.n_tb < 9 ? a.n_tb : 9;
        const int tgroups = (a.n_tb + tb_per_cta - 1) / tb_per_cta;
        a.n_tb = tgroups;
        a.n_cib = (Ci + 7) / 8;
        int64_t want32 = (a.BV + 31) / 32;
        const int gx32 = (int)(want32 < 148 * 2 ? want32 : 148 * 2);
        const dim3 grid(gx32, ((Co + 3) / 4) * a.n_cib * tgroups);
        conv_bwd_weight_warptap_kernel<4, 8, TB><<<grid, 32 * tb_per_cta, 0, st>>>(a, tb_per_cta);
        return check_launch();
    }
    // enough CTAs to fill 148 SMs a few times over, never more than the work
    int64_t want = (a.BV + 255) / 256;
    const int gx = (int)(want < 148 * 4 ? want : 148 * 4);
    if (Ci == 1) {
        a.n_cib = 1;
        const dim3 grid(gx, ((Co + 7) / 8) * a.n_cib * a.n_tb);
        conv_bwd_weight_kernel<8, 1, TB><<<grid, 256, 0, st>>>(a);
    } else if (Ci <= 4) {
        a.n_cib = 1;
        const dim3 grid(gx, ((Co + 3) / 4) * a.n_cib * a.n_tb);
        conv_bwd_weight_kernel<4, 4, TB><<<grid, 256, 0, st>>>(a);
    } else {
        a.n_cib = (Ci + 7) / 8;
        const dim3 grid(gx, ((Co + 3) / 4) * a.n_cib * a.n_tb);
        conv_bwd_weight_kernel<4, 8, TB><<<grid, 256, 0, st>>>(a);
    }
    return check_launch();
}
