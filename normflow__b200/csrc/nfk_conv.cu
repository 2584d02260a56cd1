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

extern "C" int nfk_conv_circ_fwd(const float* in, const float* w, int w_transposed, const float* bias,
                                 const uint8_t* in_mask, int in_keep,
                                 int act, const float* dact_from, int dact_kind,
                                 float* out, nfk_lattice lat, int ksize,
                                 int Ci, int Co, int64_t B, void* stream) {
    if (!in || !w || !out || !lat_ok(lat) || ksize < 1 || ksize % 2 == 0 || Ci < 1 || Co < 1) return NFK_EINVAL;
    if (B <= 0) return NFK_OK;
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
        if (smem > 48 * 1024)                                                                         \
            cudaFuncSetAttribute(conv_fwd_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
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

extern "C" int nfk_conv_circ_bwd_weight(const float* in, const uint8_t* in_mask, int in_keep,
                                        const float* gpre, float* gw, float* gbias,
                                        nfk_lattice lat, int ksize, int Ci, int Co, int64_t B, void* stream) {
    if (!in || !gpre || !gw || !lat_ok(lat) || ksize < 1 || ksize % 2 == 0 || Ci < 1 || Co < 1) return NFK_EINVAL;
    if (B <= 0) return NFK_OK;
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
