// nfk_convnd_tc.cu -- one atomic coupling step on a 2-D, 3-D or 4-D lattice with its ConvAct(1 -> 8 -> 8 -> P)
// conditioner on the tensor cores (Coupling_.forward's step k, couplings_.py:56-64; conditioner modules.py:131-145
// with ConvNd / Conv4d, convNd.py:84-127; transforms couplings_.py:123-139, 178-262).
//
// The 2-D kernel (nfk_fused_tc.cu) keeps all three layers of a strip on chip.  In 3-D / 4-D the halo of a tile
// is two (hyper)planes per layer, so the layers run as three kernels that hand the hidden layers over as fp16-PAIR
// RECORDS in HBM (one site = 8 channels = 16 bytes hi + 16 bytes lo: the same 32 bytes per site as float32,
// already in the tensor core's operand format):
//
//   layer 1 (1 -> 8, CUDA cores)   x (frozen partition) -> h1 records.  A thread owns a pair of neighbouring
//                                  sites: one of them sees the even taps, the other the odd ones (the input
//                                  vanishes on the active partition), so half of the 3^D taps are issued.
//   layer 2 (8 -> 8, tcgen05)      h1 records -> h2 records, bias + tanh + fp16 split in the TMEM epilogue.
//   layer 3 (8 -> P, tcgen05)      h2 records -> P conditioner channels per site in TMEM -> affine / RQ-spline
//                                  transform in registers -> y, log|det J|.  The (B, P, *L) tensor never exists.
//
// Layers 2 and 3 are implicit GEMMs by shifted descriptors, as in 2-D: a tile of the lattice with its one-site
// halo (periodic wrap resolved while loading) sits in shared memory as a padded box in LINEAR order, so that 128
// consecutive records are a valid K-major no-swizzle operand and a convolution tap is the same tile with the start
// address moved by the tap's linear offset -- 27 / 81 MMAs (K = 16 = 8 channels hi | lo, N = [w_hi | w_lo]) per
// tile of 128 box positions, fp32 accumulation in TMEM.  Rows that fall on the box's halo are computed and
// dropped.  A persistent CTA per SM walks (sample, tile) units: cp.async box load -> MMAs of every M tile, each
// committed to its own mbarrier -> epilogue warps drain the TMEM accumulator ring behind the MMA stream.
//
// The same file holds the BACKWARD of those conditioner layers on the tensor cores (autograd of modules.py:131-145):
//   data gradient    (nfk_convnd_dgrad)   the hidden-layer kernel on transposed, mirrored weights (MODE 2), its input
//                                         d loss / d pre-activation packed into records scaled by a power of two from its
//                                         device-side max (fp16 pairs have float16's exponent range)
//   weight gradient  (nfk_convnd_wgrad)   a GEMM over the SITES whose operands are the records read as MN-major matrices
//   nfk_convnd_layer_bwd                  both, behind one reduction + packing of the gradient

#include <stdio.h>
#include <stdlib.h>

#include "nfk_common.cuh"
#include "nfk_fused_tc.cuh"

using namespace nfk;

#define NFK_STREAM(s) reinterpret_cast<cudaStream_t>(s)

namespace {

constexpr int kNdEpiWarps = 8;       // epilogue warps: two per TMEM lane quarter, alternating M tiles
#ifndef NFK_ND_ISSUERS
#define NFK_ND_ISSUERS 2
#endif
constexpr int kNdIssuers = NFK_ND_ISSUERS;   // MMA-issuing warps (one elected thread each; M tiles dealt round robin): the
                                     // issuing thread's ~11 instructions per MMA, not the tensor pipe, pace one issuer
constexpr int kNdThreads = 32 * (kNdEpiWarps + kNdIssuers);
constexpr int kNdMaxSlots = 32;      // TMEM accumulator ring (512 columns / 16)
constexpr int kNdTmemCols = 512;     // one CTA per SM (enforced through the shared-memory request)
constexpr uint32_t kNdMinSmem = 120 * 1024;

// Geometry.  Every array has four entries: the lattice's D axes occupy the LAST D of them, the leading ones are
// dummies (extent 1, stride 0), so that device loops have a fixed trip count of four and constant indices.
struct NdGeom {
    int D, taps, ngroups;        // ngroups = 3^(D-2): groups of the nine taps of the two innermost axes
    int L[4], T[4], ntile[4], box[4], bstride[4];
    int lo[4], hi[4];            // interior box coordinates (1 .. T; 0 .. 0 for a dummy axis)
    int off[4];                  // lattice coordinate = origin + box coordinate - off
    int gstride[4];              // lattice strides in sites (unpadded field x, y)
    int pstride[4];              // strides of the PADDED record arrays (extent L + 2 per axis)
    uint32_t magic_box[4], magic_ntile[4];
    int nbox, first, span, nt, tiles_per_sample;
    int V, Vp;                   // sites per sample; padded record positions per sample and plane
    int split, nruns, run_rec;   // the box as `nruns` contiguous runs of `run_rec` records of the padded array
    int nslots, bdup, nchunk;
    int G;                       // input channel groups of 8 (hidden width / 8)
    int npass;                   // hidden layer: passes over the output channels (32 per pass at most)
    uint32_t comp_bytes;         // hi plane -> lo plane of the box in shared memory (planes: [group][hi | lo])
    uint32_t b_bytes;            // B operand image of one pass
    // Second hidden layer, stored for a COMPACT last layer: padded to the odd extent L + 3 on every axis (all strides odd:
    // the parity of a linear position is the parity of its coordinate sum) and split by that parity into two arrays,
    // so that the last layer's M tiles enumerate the ACTIVE sites only, as the 2-D kernel does (nfk_fused_tc.cuh)
    int estride[4];              // strides of that padded array
    int Ep, Eh;                  // positions per sample and plane pair; records per parity plane (Ep / 2 + 1)
    int out_split;               // hidden layer (MODE 0): write the records in that form
    int compact;                 // last layer (MODE 1): read them in that form
    int prefetch;                // pull the next unit's box into L2 while this one is computed (NFK_ND_PREFETCH=0: off)
    // Data gradient of a checkerboard-sparse input (MODE 2): the records exist on ONE parity plane of the odd-extent array
    // only (sites with coordinate sum % 2 == gpar); M tiles enumerate the output positions of one box parity at a time
    // (nt = 2 x tiles of one parity), and a tap is issued only where it lands on the populated plane -- half the MMAs
    int sparse, gpar;
    int last;                    // compact: linear box index of the last interior position (first: the first)
    uint32_t pb_bytes;           // compact: bytes of one parity plane of the box in shared memory
    uint32_t off_a, off_b, off_tab, off_bar, smem_bytes;
    int mask_parity, active_val;
};

struct NdArgs {
    const uint4* in_rec;         // [B][G][2][Vp]: per channel group a hi and a lo plane of 16-byte records, halo images included
    uint4* out_rec;              // hidden layer: the same layout
    const __half* bimg;          // B operand image prepared by nd_prep_weights_kernel
    const float* bias;           // [Co] or NULL
    const float* x;              // final layer: the field, its transform and the per-sample log-Jacobian
    float* y;
    float* log_out;
    float* save;                 // training forward: this layer's output, float32 channel-major [B][save_ch][V] (or NULL):
    int save_ch;                 // hidden layer: post-activation h (H channels); last layer: conditioner output (P channels,
                                 // written at the active sites only)
    const float* act;            // data gradient (MODE 2): post-activation output h of the layer below, [B][save_ch][V]
                                 // (multiplies by tanh' = 1 - h^2), or NULL
    const float* amax;           // data gradient: max |input| on the device; the records carry input * nd_grad_scale(amax)
    long long B;
    NdGeom g;
    RqsCfg cfg;
};

// Power of two that brings the largest magnitude of a gradient tensor to [4096, 8192): fp16 pairs have float16's
// exponent range, gradients of a mean over the batch sit far below it.
__host__ __device__ __forceinline__ float nd_grad_scale(float amax) {
    if (!(amax > 0.f) || amax > 3.0e38f) return 1.f;
    int e;
    frexpf(amax, &e);                                   // amax = m 2^e, m in [0.5, 1)
    e = 13 - e;
    return ldexpf(1.f, e > 100 ? 100 : e);
}

// n / d by the precomputed magic ceil(2^32 / d) (exact for n d < 2^32); d == 1 has no 32-bit magic
__device__ __forceinline__ int nd_div(int n, int d, uint32_t magic) { return d == 1 ? n : tc_div(n, magic); }

__device__ __forceinline__ int nd_wrap1(int c, int L) {     // c in [-1, L]
    c += c < 0 ? L : 0;
    c -= c >= L ? L : 0;
    return c;
}

// weights [Co][Ci][taps] (standard (Co, Ci, *k) layout) -> B operand images, one per pass over the output channels:
// [pass][tap][group of 8 input channels][N2][8] fp16; rows [0, NH) the hi parts of output channels pass * NH + n,
// rows [NH, 2 NH) their lo parts times 2^11 (zero rows beyond Co).  Both K groups of an MMA (activations hi | lo)
// read the same rows (descriptor LBO = 0), or `bdup` = 2 copies of them.
__global__ void nd_prep_weights_kernel(const float* w, int Co, int Ci, int taps, int NH, int npass, int bdup, __half* img) {
    const int N2 = 2 * NH, G = Ci / 8;
    const long long total = (long long)npass * taps * G * N2 * 8;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(e & 7);
        long long r = e >> 3;
        const int n = (int)(r % N2); r /= N2;
        const int gi = (int)(r % G); r /= G;
        const int t = (int)(r % taps);
        const int pass = (int)(r / taps);
        const int p = pass * NH + (n < NH ? n : n - NH);
        __half val = __float2half_rn(0.f);
        if (p < Co) {
            const float v = w[((long long)p * Ci + gi * 8 + ci) * taps + t];
            const __half hi = __float2half_rn(v);
            val = n < NH ? hi : __float2half_rn((v - __half2float(hi)) * kLoScale);
        }
        for (int kg = 0; kg < bdup; ++kg)
            img[((((long long)pass * taps + t) * G + gi) * bdup + kg) * N2 * 8 + n * 8 + ci] = val;
    }
}

// 8 channel values of one site -> the hi / lo fp16 records
__device__ __forceinline__ void nd_records(const float (&v)[8], uint4& hi, uint4& lo) {
    float l[8], h[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) h[c] = tc_split(v[c], l[c]);
    hi = make_uint4(tc_pack(h[0], h[1]), tc_pack(h[2], h[3]), tc_pack(h[4], h[5]), tc_pack(h[6], h[7]));
    lo = make_uint4(tc_pack(l[0], l[1]), tc_pack(l[2], l[3]), tc_pack(l[4], l[5]), tc_pack(l[6], l[7]));
}

// Writes the records of lattice site c into a padded record array (position c + 1 on every axis) together with
// its periodic images: a site on a face of the lattice also lives in the opposite halo (coordinate 0 -> L + 1,
// L - 1 -> 0), a site on an edge / corner in every combination of them -- the consumer's tile + halo is then a
// plain sub-box of the array that a bulk copy can fetch.  Axes with pstride 0 are dummies.
template <int ND>
__device__ __forceinline__ void nd_store_site(uint4* hi_plane, int Vp, const int (&c)[ND], const int (&L)[ND],
                                              const int (&ps)[ND], const uint4& hi, const uint4& lo) {
    int diff[ND];                       // image position - main position along an axis that has an image
    int hasmask = 0, base = 0;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        const int m = (c[d] + 1) * ps[d];
        const bool has = ps[d] != 0 && (c[d] == 0 || c[d] == L[d] - 1);
        diff[d] = (c[d] == 0 ? (L[d] + 1) * ps[d] : 0) - m;
        hasmask |= has ? 1 << d : 0;
        base += m;
    }
    hi_plane[base] = hi;
    hi_plane[Vp + base] = lo;
    for (int mask = hasmask; mask; mask = (mask - 1) & hasmask) {      // the non-empty subsets of `hasmask`
        int o = base;
#pragma unroll
        for (int d = 0; d < ND; ++d) o += (mask >> d) & 1 ? diff[d] : 0;
        hi_plane[o] = hi;
        hi_plane[Vp + o] = lo;
    }
}

// The same for the parity-split layout (extent L + 3 per axis, strides es): position P goes to parity plane P & 1 at
// index P >> 1; planes of one (sample, channel group): [hi even][hi odd][lo even][lo odd], Eh records each.
template <int ND>
__device__ __forceinline__ void nd_store_site_split(uint4* base, int Eh, const int (&c)[ND], const int (&L)[ND],
                                                    const int (&es)[ND], const uint4& hi, const uint4& lo) {
    int diff[ND];
    int hasmask = 0, P = 0;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        const int m = (c[d] + 1) * es[d];
        const bool has = es[d] != 0 && (c[d] == 0 || c[d] == L[d] - 1);
        diff[d] = (c[d] == 0 ? (L[d] + 1) * es[d] : 0) - m;
        hasmask |= has ? 1 << d : 0;
        P += m;
    }
    base[(P & 1) * Eh + (P >> 1)] = hi;
    base[(2 + (P & 1)) * Eh + (P >> 1)] = lo;
    for (int mask = hasmask; mask; mask = (mask - 1) & hasmask) {
        int o = P;
#pragma unroll
        for (int d = 0; d < ND; ++d) o += (mask >> d) & 1 ? diff[d] : 0;
        base[(o & 1) * Eh + (o >> 1)] = hi;
        base[(2 + (o & 1)) * Eh + (o >> 1)] = lo;
    }
}

// Data gradient: the same image for the TRANSPOSED layer with mirrored taps -- output rows are the layer's Ci input
// channels, the K groups its Co output channels (zero beyond Co): d/d in[ci](s) = sum_{co, t} w[co][ci][t] g[co](s - t).
__global__ void nd_prep_weights_t_kernel(const float* w, int Co, int Ci, int taps, int NH, int npass, int bdup, __half* img) {
    const int N2 = 2 * NH, G = (Co + 7) / 8;
    const long long total = (long long)npass * taps * G * N2 * 8;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int k8 = (int)(e & 7);
        long long r = e >> 3;
        const int n = (int)(r % N2); r /= N2;
        const int gi = (int)(r % G); r /= G;
        const int t = (int)(r % taps);
        const int pass = (int)(r / taps);
        const int ci = pass * NH + (n < NH ? n : n - NH), co = gi * 8 + k8;
        __half val = __float2half_rn(0.f);
        if (ci < Ci && co < Co) {
            const float v = w[((long long)co * Ci + ci) * taps + (taps - 1 - t)];
            const __half hi = __float2half_rn(v);
            val = n < NH ? hi : __float2half_rn((v - __half2float(hi)) * kLoScale);
        }
        for (int kg = 0; kg < bdup; ++kg)
            img[((((long long)pass * taps + t) * G + gi) * bdup + kg) * N2 * 8 + n * 8 + k8] = val;
    }
}

// max |v| of a float32 array -> *amax (bit pattern of a non-negative float: unsigned order = float order)
__global__ void __launch_bounds__(256) nd_amax_kernel(const float* __restrict__ v, long long n, unsigned* amax) {
    float m = 0.f;
    const long long n4 = n >> 2;
    const float4* v4 = reinterpret_cast<const float4*>(v);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 q = NFK_LDG(v4 + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(q.x), fabsf(q.y))), fmaxf(fabsf(q.z), fabsf(q.w)));
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) m = fmaxf(m, fabsf(NFK_LDG(v + (n4 << 2) + threadIdx.x)));
    if (!(m == m)) m = __int_as_float(0x7f800000);      // NaN -> inf: the scale falls back to 1, the NaN propagates
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) m = fmaxf(m, sm[i]);
        atomicMax(amax, __float_as_uint(m));
    }
}

// The same for records that exist on ONE parity plane only (checkerboard-sparse gradients): planes [hi][lo], Eh records each.
template <int ND>
__device__ __forceinline__ void nd_store_site_plane(uint4* base, int Eh, const int (&c)[ND], const int (&L)[ND],
                                                    const int (&es)[ND], const uint4& hi, const uint4& lo) {
    int diff[ND];
    int hasmask = 0, P = 0;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        const int m = (c[d] + 1) * es[d];
        const bool has = es[d] != 0 && (c[d] == 0 || c[d] == L[d] - 1);
        diff[d] = (c[d] == 0 ? (L[d] + 1) * es[d] : 0) - m;
        hasmask |= has ? 1 << d : 0;
        P += m;
    }
    base[P >> 1] = hi;
    base[Eh + (P >> 1)] = lo;
    for (int mask = hasmask; mask; mask = (mask - 1) & hasmask) {      // (images differ by even amounts: same plane)
        int o = P;
#pragma unroll
        for (int d = 0; d < ND; ++d) o += (mask >> d) & 1 ? diff[d] : 0;
        base[o >> 1] = hi;
        base[Eh + (o >> 1)] = lo;
    }
}

// ------------------------------------------------------------------------------------------- layer 1
struct NdLat {
    int L[4];                    // the D lattice extents first (unlike NdGeom), the rest 1
    int gstride[4];
    int pstride[4];
    uint32_t magic_L[4];
    int V, Vp;
};

// h1 = tanh(conv(x on the frozen partition)) as records.  A thread owns the sites (.., 2i) and (.., 2i + 1) of
// the innermost axis: the frozen one of the two sees the taps with an even number of unit steps, the active one
// those with an odd number (all other inputs are masked to zero), so each tap is issued once per pair.
template <int D>
__global__ void __launch_bounds__(256) nd_layer1_kernel(const float* __restrict__ x, const float* __restrict__ w1,
                                                        const float* __restrict__ b1, uint4* __restrict__ out_rec,
                                                        float* __restrict__ save_h1, float* __restrict__ y_frozen,
                                                        const NdLat lat, int mask_parity, int active_val, long long B) {
    constexpr int TAPS = D == 2 ? 9 : (D == 3 ? 27 : 81);
    __shared__ __align__(16) float ws[TAPS * 8];
    __shared__ __align__(16) float bs[8];
    const int gidx = blockIdx.y, G = gridDim.y;              // this block's group of 8 output channels
    for (int e = threadIdx.x; e < TAPS * 8; e += blockDim.x)
        ws[e] = kTwoLog2e * NFK_LDG(w1 + (gidx * 8 + (e & 7)) * TAPS + (e >> 3));
    if (threadIdx.x < 8) bs[threadIdx.x] = b1 ? kTwoLog2e * NFK_LDG(b1 + gidx * 8 + threadIdx.x) : 0.f;
    __syncthreads();
    const int pairs = lat.V >> 1;
    const int bps = (pairs + 255) >> 8;
    const long long b = blockIdx.x / bps;
    const int pi = (int)(blockIdx.x - b * bps) * 256 + threadIdx.x;
    if (b >= B || pi >= pairs) return;
    int Ld[D], gs[D], ps[D], c[D];
#pragma unroll
    for (int d = 0; d < D; ++d) { Ld[d] = lat.L[d]; gs[d] = lat.gstride[d]; ps[d] = lat.pstride[d]; }
    int rem = 2 * pi, csum = 0;
#pragma unroll
    for (int d = D - 1; d >= 0; --d) {
        const int q = nd_div(rem, Ld[d], lat.magic_L[d]);
        c[d] = rem - q * Ld[d];
        rem = q;
        csum += c[d];
    }
    // neighbour index contributions of the outer axes (slots 0..2 <-> the D - 1 outer lattice axes, right aligned)
    int idx[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int d = j - (4 - D);                     // lattice axis of slot j (negative: no such axis)
            idx[j][k] = d >= 0 ? nd_wrap1(c[d >= 0 ? d : 0] + k - 1, Ld[d >= 0 ? d : 0]) * gs[d >= 0 ? d : 0] : 0;
        }
    }
    int xi[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int v = c[D - 1] + k - 1;                          // in [-1, L]; the pair starts at an even coordinate
        v += v < 0 ? Ld[D - 1] : 0;
        v -= v >= Ld[D - 1] ? Ld[D - 1] : 0;
        xi[k] = v;
    }
    const bool f_first = ((1 - mask_parity + csum) & 1) != active_val;     // the pair's first site is a frozen one
    int xF[3], xA[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        xF[k] = f_first ? xi[k] : xi[k + 1];
        xA[k] = f_first ? xi[k + 1] : xi[k];
    }
    float accF[8], accA[8];
    {
        const float4 t0 = reinterpret_cast<const float4*>(bs)[0], t1 = reinterpret_cast<const float4*>(bs)[1];
        const float bv[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
        for (int co = 0; co < 8; ++co) accF[co] = accA[co] = bv[co];
    }
    const float* xb = x + b * (long long)lat.V;
#pragma unroll
    for (int k0 = 0; k0 < (D >= 4 ? 3 : 1); ++k0) {
#pragma unroll
        for (int k1 = 0; k1 < (D >= 3 ? 3 : 1); ++k1) {
#pragma unroll
            for (int k2 = 0; k2 < 3; ++k2) {
#pragma unroll
                for (int kl = 0; kl < 3; ++kl) {
                    const int t = ((k0 * (D >= 3 ? 3 : 1) + k1) * 3 + k2) * 3 + kl;
                    const int steps = (D >= 4 && k0 != 1) + (D >= 3 && k1 != 1) + (k2 != 1) + (kl != 1);
                    const int base = (D >= 4 ? idx[0][k0] : 0) + (D >= 3 ? idx[1][k1] : 0) + idx[2][k2];
                    const bool even = (steps & 1) == 0;
                    const float v = NFK_LDG(xb + base + (even ? xF[kl] : xA[kl]));
                    const float4 w0 = reinterpret_cast<const float4*>(ws + t * 8)[0];
                    const float4 w1v = reinterpret_cast<const float4*>(ws + t * 8)[1];
                    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1v.x, w1v.y, w1v.z, w1v.w};
                    if (even) {
#pragma unroll
                        for (int co = 0; co < 8; ++co) accF[co] = fmaf(v, wv[co], accF[co]);
                    } else {
#pragma unroll
                        for (int co = 0; co < 8; ++co) accA[co] = fmaf(v, wv[co], accA[co]);
                    }
                }
            }
        }
    }
    if (y_frozen && gidx == 0) {
        // a compact last layer visits the active sites only: the frozen site of the pair is passed through here
        const int s = (D >= 4 ? idx[0][1] : 0) + (D >= 3 ? idx[1][1] : 0) + idx[2][1] + xF[1];
        y_frozen[b * (long long)lat.V + s] = NFK_LDG(xb + s);
    }
    float vF[8], vA[8];
#pragma unroll
    for (int co = 0; co < 8; ++co) {
        vF[co] = tanh_from_scaled(accF[co]);
        vA[co] = tanh_from_scaled(accA[co]);
    }
    uint4 hF, lF, hA, lA;
    nd_records(vF, hF, lF);
    nd_records(vA, hA, lA);
    uint4* ob = out_rec + (b * G + gidx) * 2LL * lat.Vp;
    int cF[D], cA[D];
#pragma unroll
    for (int d = 0; d < D; ++d) cF[d] = cA[d] = c[d];
    cF[D - 1] += f_first ? 0 : 1;
    cA[D - 1] += f_first ? 1 : 0;
    nd_store_site<D>(ob, lat.Vp, cF, Ld, ps, hF, lF);
    nd_store_site<D>(ob, lat.Vp, cA, Ld, ps, hA, lA);
    if (save_h1) {                                            // training forward: h1 as float32 [B][H][V] as well
        float* sp = save_h1 + (b * G + gidx) * 8LL * lat.V + 2 * pi;
#pragma unroll
        for (int co = 0; co < 8; ++co) {
            const float2 v = f_first ? make_float2(vF[co], vA[co]) : make_float2(vA[co], vF[co]);
            *reinterpret_cast<float2*>(sp + (long long)co * lat.V) = v;
        }
    }
}

// Data gradient input: float32 channel-major g [B][C][V] -> scaled fp16-pair records with their periodic images
// ([B][ceil(C/8)][2][Vp], zero channels beyond C).  A thread owns one site of one channel group.
template <int D>
__global__ void __launch_bounds__(256) nd_pack_records_kernel(const float* __restrict__ gsrc, int C, const float* __restrict__ amax,
                                                              uint4* __restrict__ out_rec, const NdLat lat, long long B) {
    const int gidx = blockIdx.y, G = gridDim.y;
    const int bps = (lat.V + 255) >> 8;
    const long long b = blockIdx.x / bps;
    const int site = (int)(blockIdx.x - b * bps) * 256 + threadIdx.x;
    if (b >= B || site >= lat.V) return;
    const float scale = amax ? nd_grad_scale(NFK_LDG(amax)) : 1.f;
    int Ld[D], ps[D], c[D];
#pragma unroll
    for (int d = 0; d < D; ++d) { Ld[d] = lat.L[d]; ps[d] = lat.pstride[d]; }
    int rem = site;
#pragma unroll
    for (int d = D - 1; d >= 0; --d) {
        const int q = nd_div(rem, Ld[d], lat.magic_L[d]);
        c[d] = rem - q * Ld[d];
        rem = q;
    }
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int ch = gidx * 8 + k;
        v[k] = ch < C ? scale * NFK_LDG(gsrc + (b * C + ch) * (long long)lat.V + site) : 0.f;
    }
    uint4 hi, lo;
    nd_records(v, hi, lo);
    nd_store_site<D>(out_rec + (b * G + gidx) * 2LL * lat.Vp, lat.Vp, c, Ld, ps, hi, lo);
}

uint32_t nd_magic(int d);
// host side of nd_pack_records_kernel (L / pstride in NdGeom's right-aligned slots)
int nd_pack(const float* src, int C, const float* amax, uint4* rec, int D, const int (&L)[4], const int (&pstride)[4], int V, int Vp,
            int64_t B, cudaStream_t st) {
    NdLat nl{};
    for (int d = 0; d < 4; ++d) {
        const int j = d + 4 - D;
        nl.L[d] = d < D ? L[j] : 1;
        nl.pstride[d] = d < D ? pstride[j] : 0;
        nl.magic_L[d] = nd_magic(nl.L[d]);
    }
    nl.V = V;
    nl.Vp = Vp;
    const long long blocks = B * ((V + 255) / 256);
    if (blocks >= (1LL << 31)) return NFK_EUNSUPPORTED;
    const dim3 gridp((unsigned)blocks, (unsigned)((C + 7) / 8));
    switch (D) {
        case 2: nd_pack_records_kernel<2><<<gridp, 256, 0, st>>>(src, C, amax, rec, nl, B); break;
        case 3: nd_pack_records_kernel<3><<<gridp, 256, 0, st>>>(src, C, amax, rec, nl, B); break;
        default: nd_pack_records_kernel<4><<<gridp, 256, 0, st>>>(src, C, amax, rec, nl, B); break;
    }
    return check_launch();
}

// Checkerboard-sparse gradient (non-zero on the sites whose coordinate sum % 2 == gpar only): records of those sites
// on one parity plane of the odd-extent array ([B][ceil(C/8)][2][Eh]).  A thread owns a pair of sites of the innermost axis.
template <int D>
__global__ void __launch_bounds__(256) nd_pack_active_kernel(const float* __restrict__ gsrc, int C, const float* __restrict__ amax,
                                                             uint4* __restrict__ out_rec, const NdLat lat, int Eh, int gpar,
                                                             long long B) {
    const int gidx = blockIdx.y, G = gridDim.y;
    const int pairs = lat.V >> 1;
    const int bps = (pairs + 255) >> 8;
    const long long b = blockIdx.x / bps;
    const int pi = (int)(blockIdx.x - b * bps) * 256 + threadIdx.x;
    if (b >= B || pi >= pairs) return;
    const float scale = amax ? nd_grad_scale(NFK_LDG(amax)) : 1.f;
    int Ld[D], es[D], c[D];
#pragma unroll
    for (int d = 0; d < D; ++d) { Ld[d] = lat.L[d]; es[d] = lat.pstride[d]; }       // (pstride carries the odd-extent strides here)
    int rem = 2 * pi, csum = 0;
#pragma unroll
    for (int d = D - 1; d >= 0; --d) {
        const int q = nd_div(rem, Ld[d], lat.magic_L[d]);
        c[d] = rem - q * Ld[d];
        rem = q;
        csum += c[d];
    }
    const int add = (csum ^ gpar) & 1;                  // the pair's populated site
    c[D - 1] += add;
    const int site = 2 * pi + add;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int ch = gidx * 8 + k;
        v[k] = ch < C ? scale * NFK_LDG(gsrc + (b * C + ch) * (long long)lat.V + site) : 0.f;
    }
    uint4 hi, lo;
    nd_records(v, hi, lo);
    nd_store_site_plane<D>(out_rec + (b * G + gidx) * 2LL * Eh, Eh, c, Ld, es, hi, lo);
}

int nd_pack_active(const float* src, int C, const float* amax, uint4* rec, int D, const int (&L)[4], const int (&estride)[4], int V,
                   int Eh, int gpar, int64_t B, cudaStream_t st) {
    NdLat nl{};
    for (int d = 0; d < 4; ++d) {
        const int j = d + 4 - D;
        nl.L[d] = d < D ? L[j] : 1;
        nl.pstride[d] = d < D ? estride[j] : 0;
        nl.magic_L[d] = nd_magic(nl.L[d]);
    }
    nl.V = V;
    const long long blocks = B * ((V / 2 + 255) / 256);
    if (blocks >= (1LL << 31)) return NFK_EUNSUPPORTED;
    const dim3 gridp((unsigned)blocks, (unsigned)((C + 7) / 8));
    switch (D) {
        case 2: nd_pack_active_kernel<2><<<gridp, 256, 0, st>>>(src, C, amax, rec, nl, Eh, gpar, B); break;
        case 3: nd_pack_active_kernel<3><<<gridp, 256, 0, st>>>(src, C, amax, rec, nl, Eh, gpar, B); break;
        default: nd_pack_active_kernel<4><<<gridp, 256, 0, st>>>(src, C, amax, rec, nl, Eh, gpar, B); break;
    }
    return check_launch();
}

// ------------------------------------------------------------------------------------------- layers 2 and 3
__device__ __forceinline__ void nd_bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void nd_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void nd_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}

// The nine taps of one (outer tap group, channel group) step of the ACTIVE-SITES-ONLY last layer; WANT = parity of the
// step's base offset.  `ad` already points at row shift bh of plane 0.
template <int WANT>
__device__ __forceinline__ void nd_issue9_compact(uint32_t acc, uint64_t ad, uint64_t bdg, int h2, int pbrec, int tap_step,
                                                  uint32_t idesc, int cnt) {
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        constexpr int kNone = 0;
        const int dy = i / 3 - 1, dx = i % 3 - 1;
        const int e = WANT + dy + dx;                         // in [-2, 3]
        const int sh = (e + 4) / 2 - 2 + kNone;               // floor(e / 2)
        const int pl = (e + 4) & 1;
        const int dl = dy * h2 + sh + pl * pbrec;
        tc::mma_f16(acc, tc::desc_advance(ad, dl), tc::desc_advance(bdg, i * tap_step), idesc, (cnt | i) != 0);
    }
}

// MODE 0: hidden layer (H -> OC of the H output channels per pass, tanh) -> records; the template argument K
// carries OC (8, 16 or 32).  MODE 1: last layer (H -> P) + transform of the field.  MODE 2: data gradient of a layer
// (records of d loss / d pre-activation, transposed mirrored weights) -> float32 [B][Ci][V], times 1 - h^2.
// No CTA-wide barrier inside the unit loop: one thread of the last warp fetches the box of a unit with bulk
// copies (the tile and its halo are contiguous runs of the padded record arrays, one pair of planes per group of
// 8 input channels), issues the MMAs of its M tiles -- taps x channel groups, K = 16 = 8 channels hi | lo each --
// into a ring of TMEM accumulator slots and commits each tile to an mbarrier; the epilogue warps drain the ring.
template <int MODE, int KIND, int K, int INV>
__global__ void __launch_bounds__(kNdThreads, 1) convnd_tc_kernel(const NdArgs a) {
    constexpr int P = MODE != 1 ? K : (KIND == 0 ? 2 : 3 * K - 2);
    constexpr int NH = (P + 7) / 8 * 8, N2 = 2 * NH;
    extern __shared__ __align__(128) uint8_t smem[];
    const NdGeom& g = a.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // only the last layer reads parity-split records: for the other modes the per-tap test in the MMA issue loop compiles
    // away (a predicated branch per MMA on the issuing thread is measurable: see the sparse data gradient)
    const bool kcompact = MODE == 1 && g.compact;

    uint8_t* A = smem + g.off_a;
    uint8_t* Bs = smem + g.off_b;
    float* bias_s = reinterpret_cast<float*>(smem + g.off_tab);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + g.off_bar);
    uint64_t* empty = full + kNdMaxSlots;
    uint64_t* loaded = empty + kNdMaxSlots;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(loaded + 1);

    // ---- one-time set-up ---------------------------------------------------------------------------
    {
        const int nbias = MODE != 1 ? g.npass * NH : NH;
        for (int c = tid; c < nbias; c += kNdThreads) {
            const float v = (a.bias && (MODE != 1 || c < P)) ? NFK_LDG(a.bias + c) : 0.f;
            bias_s[c] = MODE == 0 ? kTwoLog2e * v : v;
        }
        if (tid == 0) {
            for (int i = 0; i < kNdMaxSlots; ++i) {
                tc::mbar_init(tc::smem_u32(full + i), 1);
                tc::mbar_init(tc::smem_u32(empty + i), 4);
            }
            tc::mbar_init(tc::smem_u32(loaded), 1);
            tc::fence_mbar_init();
        }
        if (warp == kNdEpiWarps) tc::tmem_alloc(tc::smem_u32(tmem_slot), kNdTmemCols);
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    }
    const uint32_t tmem = *tmem_slot;
    const long long per_pass = a.B * g.tiles_per_sample;
    const long long nunits = per_pass * g.npass;
    const int nslots = g.nslots;
    const int cols_per_slot = g.nchunk * N2;
    const int G = g.G;

    // pass over the output channels, sample and tile origin (lattice coordinates) of a unit
    auto unit_origin = [&](long long unit, int& pass, long long& b, int (&org)[4]) {
        pass = (int)(unit / per_pass);
        unit -= pass * per_pass;
        b = unit / g.tiles_per_sample;
        int trem = (int)(unit - b * g.tiles_per_sample);
#pragma unroll
        for (int d = 3; d >= 0; --d) {
            const int q = nd_div(trem, g.ntile[d], g.magic_ntile[d]);
            org[d] = (trem - q * g.ntile[d]) * g.T[d];
            trem = q;
        }
    };

    // compact last layer: parity of the linear box position of the unit's ACTIVE sites (all box strides are odd, so it is
    // the parity of the box-coordinate sum; lattice coordinate = origin + box coordinate - 1 on each of the D axes)
    auto unit_parity = [&](const int (&org)[4]) {
        return (g.active_val ^ (1 - g.mask_parity) ^ (org[0] + org[1] + org[2] + org[3]) ^ g.D) & 1;
    };

    if (warp >= kNdEpiWarps) {
        // =============================== loads + MMA issue (one elected thread per issuer warp) ===========
        const int iw = warp - kNdEpiWarps;                            // issuer 0 also fetches the boxes
        if (tc::elect_one()) {
            const uint32_t a_base = tc::smem_u32(A);
            const uint32_t b_base = tc::smem_u32(Bs);
            // (compact: the hi -> lo distance is two parity planes; plane 1 follows plane 0)
            const bool sparse = MODE == 2 && g.sparse;
            const uint64_t a_desc = tc::make_desc(a_base, sparse ? g.pb_bytes : (kcompact ? 2u * g.pb_bytes : g.comp_bytes), 128);
            const int pbrec = (int)(g.pb_bytes >> 4);
            const uint64_t b_desc = tc::make_desc(b_base, g.bdup == 2 ? N2 * 16 : 0, 128);
            const uint32_t idesc = tc::make_idesc(0, 128, N2);
            const uint32_t loaded_bar = tc::smem_u32(loaded);
            const int s2 = g.bstride[2];
            // steps (nine taps of one channel group) per accumulation chain
            const int chain_len = (g.ngroups * G + g.nchunk - 1) / g.nchunk;
            const int bstep = g.bdup * N2;                           // B rows per (tap, channel group)
            const int gstep = (int)((sparse ? 2 * g.pb_bytes : (kcompact ? 4 * g.pb_bytes : 2 * g.comp_bytes)) >> 4);   // records between two channel groups' planes
            int slot = 0;
            uint32_t ring_phase = 0, load_phase = 0;
            int last_slot[kNdIssuers], cur_pass = -1;                // last tile of every issuer in the previous unit
            uint32_t last_phase[kNdIssuers];
#pragma unroll
            for (int w = 0; w < kNdIssuers; ++w) { last_slot[w] = -1; last_phase[w] = 0; }
            for (long long unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
                long long b;
                int org[4], pass;
                unit_origin(unit, pass, b, org);
                const int pib = unit_parity(org);
                const int cfirst = (g.first + ((pib ^ g.first) & 1)) >> 1;
                // sparse: box parity of the populated (non-zero) positions of this unit
                const int abox = (g.gpar ^ org[0] ^ org[1] ^ org[2] ^ org[3] ^ g.D) & 1;
                if (iw == 0) {
                    // box (and weights) may be overwritten once every MMA of the previous unit has completed
#pragma unroll
                    for (int w = 0; w < kNdIssuers; ++w)
                        if (last_slot[w] >= 0) tc::mbar_wait(tc::smem_u32(full + last_slot[w]), last_phase[w]);
                    const uint32_t run_bytes = (uint32_t)g.run_rec * 16;
                    uint32_t tx = sparse ? 2u * G * (uint32_t)((g.nbox + 1 - abox) >> 1) * 16u
                                         : (kcompact ? 2u * G * (uint32_t)g.nbox * 16u : 2u * G * g.nruns * run_bytes);
                    if (pass != cur_pass) tx += g.b_bytes;
                    nd_expect_tx(loaded_bar, tx);
                    if (pass != cur_pass) {
                        nd_bulk_load(b_base, reinterpret_cast<const uint8_t*>(a.bimg) + (size_t)pass * g.b_bytes,
                                     g.b_bytes, loaded_bar);
                        cur_pass = pass;
                    }
                    for (int k = 0; k < g.nruns; ++k) {
                        int rem = k, so = 0;
#pragma unroll
                        for (int d = 3; d >= 0; --d) {
                            const int st = (kcompact || sparse) ? g.estride[d] : g.pstride[d];
                            if (d < g.split) {
                                const int q = nd_div(rem, g.box[d], g.magic_box[d]);
                                so += (org[d] + rem - q * g.box[d]) * st;
                                rem = q;
                            } else {
                                so += org[d] * st;         // (d == split; the axes behind it are whole: origin 0)
                            }
                        }
                        if (sparse) {
                            // the run's positions of box parity abox, from the one populated plane of the source
                            const int a0 = k * g.run_rec;
                            const int pq = a0 + ((abox ^ a0) & 1);
                            if (pq < a0 + g.run_rec) {
                                const int n = (a0 + g.run_rec - pq + 1) >> 1;
                                const int Pq = so + (pq - a0);
                                for (int gi = 0; gi < G; ++gi) {
                                    const uint4* src = a.in_rec + (b * G + gi) * 2LL * g.Eh + (Pq >> 1);
                                    const uint32_t dst = a_base + (uint32_t)gi * 2u * g.pb_bytes + (pq >> 1) * 16;
                                    nd_bulk_load(dst, src, n * 16, loaded_bar);
                                    nd_bulk_load(dst + g.pb_bytes, src + g.Eh, n * 16, loaded_bar);
                                }
                            }
                        } else if (!kcompact) {
                            for (int gi = 0; gi < G; ++gi) {
                                const uint4* src = a.in_rec + (b * G + gi) * 2LL * g.Vp + so;
                                const uint32_t dst = a_base + (uint32_t)gi * 2u * g.comp_bytes + k * run_bytes;
                                nd_bulk_load(dst, src, run_bytes, loaded_bar);
                                nd_bulk_load(dst + g.comp_bytes, src + g.Vp, run_bytes, loaded_bar);
                            }
                        } else {
                            // parity-split box: the run's positions of box parity q sit in source plane (so + ...) & 1
                            const int a0 = k * g.run_rec;
#pragma unroll
                            for (int q = 0; q < 2; ++q) {
                                const int pq = a0 + ((q ^ a0) & 1);
                                if (pq >= a0 + g.run_rec) continue;
                                const int n = (a0 + g.run_rec - pq + 1) >> 1;
                                const int Pq = so + (pq - a0);
                                for (int gi = 0; gi < G; ++gi) {
                                    const uint4* src = a.in_rec + (b * G + gi) * 4LL * g.Eh + (Pq & 1) * g.Eh + (Pq >> 1);
                                    const uint32_t dst = a_base + (uint32_t)gi * 4u * g.pb_bytes + q * g.pb_bytes + (pq >> 1) * 16;
                                    nd_bulk_load(dst, src, n * 16, loaded_bar);
                                    nd_bulk_load(dst + 2u * g.pb_bytes, src + 2LL * g.Eh, n * 16, loaded_bar);
                                }
                            }
                        }
                    }
                }
                if (iw == kNdIssuers - 1 && g.prefetch && unit + gridDim.x < nunits) {
                    // the single box buffer leaves the load of a unit exposed: pull the NEXT unit's runs into L2 meanwhile
                    long long nb;
                    int norg[4], npass;
                    unit_origin(unit + gridDim.x, npass, nb, norg);
                    for (int k = 0; k < g.nruns; ++k) {
                        int rem = k, so = 0;
#pragma unroll
                        for (int d = 3; d >= 0; --d) {
                            const int st = (kcompact || sparse) ? g.estride[d] : g.pstride[d];
                            if (d < g.split) {
                                const int q = nd_div(rem, g.box[d], g.magic_box[d]);
                                so += (norg[d] + rem - q * g.box[d]) * st;
                                rem = q;
                            } else {
                                so += norg[d] * st;
                            }
                        }
                        for (int gi = 0; gi < G; ++gi) {
                            if (sparse) {
                                const uint4* src = a.in_rec + (nb * G + gi) * 2LL * g.Eh + (so >> 1);
                                const uint32_t n16 = (uint32_t)((g.run_rec + 1) >> 1) * 16;
                                nd_prefetch_l2(src, n16);
                                nd_prefetch_l2(src + g.Eh, n16);
                            } else if (!kcompact) {
                                const uint4* src = a.in_rec + (nb * G + gi) * 2LL * g.Vp + so;
                                nd_prefetch_l2(src, (uint32_t)g.run_rec * 16);
                                nd_prefetch_l2(src + g.Vp, (uint32_t)g.run_rec * 16);
                            } else {
                                const uint4* src = a.in_rec + (nb * G + gi) * 4LL * g.Eh + (so >> 1);
                                const uint32_t n16 = (uint32_t)((g.run_rec + 1) >> 1) * 16;
#pragma unroll
                                for (int pl = 0; pl < 4; ++pl) nd_prefetch_l2(src + (long long)pl * g.Eh, n16);
                            }
                        }
                    }
                }
                tc::mbar_wait(loaded_bar, load_phase);
                load_phase ^= 1u;
                tc::fence_after_sync();
                for (int m = 0; m < g.nt; ++m) {
                    // a TMEM slot belongs to one issuer (and one epilogue half) for good: its mbarrier phases are
                    // then walked by one thread in order, which is what a parity wait can tell apart
                    const int mine = slot % kNdIssuers;
                    last_slot[mine] = slot;
                    last_phase[mine] = ring_phase;
                    if (mine != iw) {                                  // another issuer's tile
                        if (++slot == nslots) { slot = 0; ring_phase ^= 1u; }
                        continue;
                    }
                    tc::mbar_wait(tc::smem_u32(empty + slot), ring_phase ^ 1u);
                    tc::fence_after_sync();
                    // The tensor core adds into its fp32 accumulator with truncation, so the error of a long
                    // accumulation chain grows linearly with its length: the taps x channel groups MMAs of a tile
                    // are cut into `nchunk` chains, each in its own TMEM columns, which the epilogue adds up with
                    // round-to-nearest
                    // compact: tile rows are the active positions c (box position 2 c + pib); a tap with linear offset
                    // d reads parity plane (pib + d) & 1 at row c + ((pib + d) >> 1) -- the same shift for every row
                    // sparse: tiles 0 .. nt/2 - 1 are the output positions of box parity 0, the rest those of parity 1
                    const int rp = sparse ? (m >= (g.nt >> 1) ? 1 : 0) : 0;
                    const int ms = sparse ? m - rp * (g.nt >> 1) : m;
                    const int cfs = (g.first + ((rp ^ g.first) & 1)) >> 1;
                    const uint64_t ad0 = tc::desc_advance(a_desc, (sparse ? cfs : (kcompact ? cfirst : g.first)) + ms * 128);
                    uint64_t bd = b_desc;
                    uint32_t acc = tmem + slot * cols_per_slot;
                    int cnt = 0;
                    uint32_t accf = 0;                                 // sparse: 0 for the first MMA of a chain
                    if (MODE == 2 && sparse) {
                        for (int o = 0; o < g.ngroups; ++o) {
                            int od;
                            if (g.ngroups == 9) od = (o / 3 - 1) * g.bstride[0] + (o % 3 - 1) * g.bstride[1];
                            else if (g.ngroups == 3) od = (o - 1) * g.bstride[1];
                            else od = 0;
                            uint64_t ad = ad0;
                            uint64_t bdg = bd;
                            // (all box strides odd: the parity of a tap's offset is that of its number of unit steps, so the taps
                            // that land on the populated plane are the five "even" or the four "odd" ones of each group of nine)
                            const int want = (abox ^ rp ^ od) & 1;
#pragma unroll 1
                            for (int gi = 0; gi < G; ++gi) {
#pragma unroll
                                for (int i = 0; i < 9; ++i) {
                                    const int dl = (i / 3 - 1) * s2 + (i % 3 - 1);
                                    if ((((i / 3) + (i % 3)) & 1) == want) {
                                        tc::mma_f16(acc, tc::desc_advance(ad, (rp + od + dl) >> 1), tc::desc_advance(bdg, i * G * bstep),
                                                    idesc, accf);
                                        accf = 1u;
                                    }
                                }
                                if (++cnt == chain_len) { cnt = 0; acc += N2; accf = 0; }
                                ad = tc::desc_advance(ad, gstep);
                                bdg = tc::desc_advance(bdg, bstep);
                            }
                            bd = tc::desc_advance(bd, 9 * G * bstep);
                        }
                    } else {
                        for (int o = 0; o < g.ngroups; ++o) {
                            // offset of the outer taps of this group: o enumerates (k0, k1) of the two outer axes
                            int od;
                            if (g.ngroups == 9) od = (o / 3 - 1) * g.bstride[0] + (o % 3 - 1) * g.bstride[1];
                            else if (g.ngroups == 3) od = (o - 1) * g.bstride[1];
                            else od = 0;
                            uint64_t ad = tc::desc_advance(ad0, od);
                            uint64_t bdg = bd;
                            // one compact copy of the nine unrolled taps, looped over the channel groups; a chain is a
                            // whole number of such steps
                            if (kcompact) {
                                // Active rows only: tap (dy, dx) of this group reads parity plane (base + dy s2 + dx) & 1 at row
                                // shift (base + dy s2 + dx) >> 1, base = pib + od.  s2 is odd (= 2 h2 + 1), so with base =
                                // 2 bh + want the shift is bh + dy h2 + ((want + dy + dx) >> 1) and the plane (want + dy + dx) & 1:
                                // compile-time constants per tap once `want` is known -- nothing per MMA on the issuing thread
                                // but the descriptor add.
                                const int base = pib + od, want = base & 1, bh = base >> 1, h2 = s2 >> 1;
                                const uint64_t adc = tc::desc_advance(ad0, bh);
                                uint64_t adg = adc;
#pragma unroll 1
                                for (int gi = 0; gi < G; ++gi) {
                                    if (want) nd_issue9_compact<1>(acc, adg, bdg, h2, pbrec, G * bstep, idesc, cnt);
                                    else nd_issue9_compact<0>(acc, adg, bdg, h2, pbrec, G * bstep, idesc, cnt);
                                    if (++cnt == chain_len) { cnt = 0; acc += N2; }
                                    adg = tc::desc_advance(adg, gstep);
                                    bdg = tc::desc_advance(bdg, bstep);
                                }
                            } else {
#pragma unroll 1
                                for (int gi = 0; gi < G; ++gi) {
#pragma unroll
                                    for (int i = 0; i < 9; ++i) {
                                        const int dl = (i / 3 - 1) * s2 + (i % 3 - 1);
                                        tc::mma_f16(acc, tc::desc_advance(ad, dl), tc::desc_advance(bdg, i * G * bstep), idesc,
                                                    (cnt | i) != 0);
                                    }
                                    if (++cnt == chain_len) { cnt = 0; acc += N2; }
                                    ad = tc::desc_advance(ad, gstep);
                                    bdg = tc::desc_advance(bdg, bstep);
                                }
                            }
                            bd = tc::desc_advance(bd, 9 * G * bstep);
                        }
                    }
                    tc::mma_commit(tc::smem_u32(full + slot));
                    if (++slot == nslots) { slot = 0; ring_phase ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else {
        // =============================== epilogue =======================================================
        const int quarter = warp & 3, half = warp >> 2;            // TMEM lanes 32 quarter .. +31; tiles of my parity
        const float inv_scale = MODE == 2 ? 1.f / nd_grad_scale(NFK_LDG(a.amax)) : 1.f;
        const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
        int slot = 0;
        uint32_t ring_phase = 0;
        for (long long unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
            long long b;
            int org[4], pass;
            unit_origin(unit, pass, b, org);
            const int pib = unit_parity(org);
            const int cfirst = (g.first + ((pib ^ g.first) & 1)) >> 1;
            float lsum = 0.f;
            for (int m = 0; m < g.nt; ++m) {
                if ((slot & 1) != half) {                          // the other warp of my quarter owns this slot
                    if (++slot == nslots) { slot = 0; ring_phase ^= 1u; }
                    continue;
                }
                // data gradient: the tanh outputs this row's result is multiplied by are fetched BEFORE waiting for the
                // tile's MMAs (eight dependent global loads per row were a fifth of the kernel's stall samples)
                float hv[MODE == 2 ? NH : 1];
                if (MODE == 2 && a.act) {
                    const int r0 = m * 128 + quarter * 32 + lane;
                    int rem0;
                    bool ok0 = true;
                    if (g.sparse) {
                        const int rp = m >= (g.nt >> 1) ? 1 : 0;
                        const int cfs = (g.first + ((rp ^ g.first) & 1)) >> 1;
                        rem0 = 2 * (cfs + (r0 - rp * (g.nt >> 1) * 128)) + rp;
                        ok0 = rem0 <= g.last;
                    } else {
                        ok0 = r0 < g.span;
                        rem0 = g.first + r0;
                    }
                    int site0 = 0;
#pragma unroll
                    for (int d = 3; d >= 0; --d) {
                        const int q = nd_div(rem0, g.box[d], g.magic_box[d]);
                        const int i = rem0 - q * g.box[d];
                        rem0 = q;
                        ok0 = ok0 && i >= g.lo[d] && i <= g.hi[d];
                        site0 += (org[d] + i - g.off[d]) * g.gstride[d];
                    }
#pragma unroll
                    for (int c = 0; c < NH; ++c)
                        hv[c] = ok0 ? NFK_LDG(a.act + (b * a.save_ch + pass * NH + c) * (long long)g.V + site0) : 0.f;
                }
                tc::mbar_wait(tc::smem_u32(full + slot), ring_phase);
                tc::fence_after_sync();
                float hi[NH], lo[NH];
#pragma unroll
                for (int c = 0; c < NH; ++c) hi[c] = lo[c] = 0.f;
                for (int ch = 0; ch < g.nchunk; ++ch) {
                    const uint32_t col = lane_addr + slot * cols_per_slot + ch * N2;
                    if (NH == 8) {
                        // two chains per round trip to TMEM (a load + wait is ~150 cycles; the narrow layers' tiles are
                        // a few hundred cycles of MMAs, so the epilogue paced them)
                        float acc[16], acc2[16];
                        const bool two = ch + 1 < g.nchunk;
                        tc::tmem_ld16(col, acc);
                        if (two) tc::tmem_ld16(col + N2, acc2);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 8; ++c) { hi[c] += acc[c]; lo[c] += acc[8 + c]; }
                        if (two) {
#pragma unroll
                            for (int c = 0; c < 8; ++c) { hi[c] += acc2[c]; lo[c] += acc2[8 + c]; }
                            ++ch;
                        }
                    } else {
#pragma unroll
                        for (int q8 = 0; q8 < NH / 8; ++q8) {
                            float h8[8], l8[8];
                            tc::tmem_ld8(col + q8 * 8, h8);
                            tc::tmem_ld8(col + NH + q8 * 8, l8);
                            tc::tmem_ld_wait();
#pragma unroll
                            for (int c = 0; c < 8; ++c) { hi[q8 * 8 + c] += h8[c]; lo[q8 * 8 + c] += l8[c]; }
                        }
                    }
                }
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(tc::smem_u32(empty + slot));     // the accumulator slot is free again
                if (++slot == nslots) { slot = 0; ring_phase ^= 1u; }
                // which site is this row?
                const int r = m * 128 + quarter * 32 + lane;
                int rem;
                if (MODE == 2 && g.sparse) {
                    const int rp = m >= (g.nt >> 1) ? 1 : 0;
                    const int cfs = (g.first + ((rp ^ g.first) & 1)) >> 1;
                    rem = 2 * (cfs + (r - rp * (g.nt >> 1) * 128)) + rp;
                    if (rem > g.last) continue;
                } else if (kcompact) {
                    rem = 2 * (cfirst + r) + pib;
                    if (rem > g.last) continue;
                } else {
                    if (r >= g.span) continue;
                    rem = g.first + r;
                }
                int site = 0, csum = 0;
                int cc[4];
                bool interior = true;
#pragma unroll
                for (int d = 3; d >= 0; --d) {
                    const int q = nd_div(rem, g.box[d], g.magic_box[d]);
                    const int i = rem - q * g.box[d];
                    rem = q;
                    interior = interior && i >= g.lo[d] && i <= g.hi[d];
                    cc[d] = org[d] + i - g.off[d];
                    site += cc[d] * g.gstride[d];
                    csum += cc[d];
                }
                if (!interior) continue;
                if (MODE == 0) {
                    const int Gout = g.npass * (NH / 8);            // channel groups of the output array
#pragma unroll
                    for (int q8 = 0; q8 < NH / 8; ++q8) {
                        float v[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            v[c] = tanh_from_scaled(fmaf(lo[q8 * 8 + c], kTwoLog2e / kLoScale,
                                                         fmaf(hi[q8 * 8 + c], kTwoLog2e, bias_s[pass * NH + q8 * 8 + c])));
                        uint4 rh, rl;
                        nd_records(v, rh, rl);
                        if (g.out_split)
                            nd_store_site_split<4>(a.out_rec + (b * Gout + pass * (NH / 8) + q8) * 4LL * g.Eh, g.Eh, cc,
                                                   g.L, g.estride, rh, rl);
                        else
                            nd_store_site<4>(a.out_rec + (b * Gout + pass * (NH / 8) + q8) * 2LL * g.Vp, g.Vp, cc, g.L,
                                             g.pstride, rh, rl);
                        if (a.save) {
                            float* sp = a.save + (b * a.save_ch + pass * NH + q8 * 8) * (long long)g.V + site;
#pragma unroll
                            for (int c = 0; c < 8; ++c) sp[(long long)c * g.V] = v[c];
                        }
                    }
                } else if (MODE == 2) {
                    // data gradient: undo the input scale, multiply by tanh'(pre-activation) = 1 - h^2 of the layer below
#pragma unroll
                    for (int c = 0; c < NH; ++c) {
                        const long long o = (b * a.save_ch + pass * NH + c) * (long long)g.V + site;
                        float v = fmaf(lo[c], 1.f / kLoScale, hi[c]) * inv_scale;
                        if (a.act) v *= fmaf(-hv[c], hv[c], 1.f);
                        a.save[o] = v;
                    }
                } else {
                    const float xv = NFK_LDG(a.x + b * (long long)g.V + site);
                    float out = xv;
                    if (((1 - g.mask_parity + csum) & 1) == g.active_val) {
                        float prm[NH];
#pragma unroll
                        for (int c = 0; c < NH; ++c) prm[c] = fmaf(lo[c], 1.f / kLoScale, hi[c]) + bias_s[c];
                        if (a.save) {
                            float* sp = a.save + b * a.save_ch * (long long)g.V + site;
#pragma unroll
                            for (int c = 0; c < P; ++c) sp[(long long)c * g.V] = prm[c];
                        }
                        float l;
                        if (KIND == 0) {
                            const float t = prm[0], sc = fabsf(prm[1]);
                            if (!INV) { out = fmaf(xv, fast_ex2(-sc * kInvLn2), t); l = -sc; }
                            else { out = (xv - t) * fast_ex2(sc * kInvLn2); l = sc; }
                        } else {
                            tc_rqs<K, INV, NH>(prm, a.cfg, xv, out, l);
                        }
                        lsum += l;
                    }
                    a.y[b * (long long)g.V + site] = out;
                }
            }
            if (MODE == 1) {
                lsum = warp_sum(lsum);
                if (lane == 0) atomicAdd(a.log_out + b, lsum);
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == kNdEpiWarps) tc::tmem_dealloc(tmem, kNdTmemCols);
}

// ------------------------------------------------------------------------------------------- host side
uint32_t nd_magic(int d) { return (uint32_t)((0x100000000ULL + d - 1) / d); }

// cycles of one M = 128 MMA by accumulator width (measured, scratch/tc_probe.cu)
float nd_mma_cycles(int N2) { return N2 <= 32 ? 42.f : N2 <= 64 ? 52.f : 68.f; }

// Fills the lattice part of the geometry: the D axes right-aligned in the four slots.
void nd_lattice(NdGeom& g, const nfk_lattice& lat) {
    const int D = lat.ndim, r0 = 4 - D;
    g.D = D;
    g.taps = 1;
    for (int d = 0; d < D; ++d) g.taps *= 3;
    g.ngroups = g.taps / 9;
    for (int j = 0; j < 4; ++j) g.L[j] = j >= r0 ? lat.shape[j - r0] : 1;
    int gs = 1, ps = 1;
    for (int j = 3; j >= 0; --j) {
        const bool real = j >= r0;
        g.gstride[j] = real ? gs : 0;
        g.pstride[j] = real ? ps : 0;
        if (real) { gs *= g.L[j]; ps *= g.L[j] + 2; }
        g.off[j] = real ? 1 : 0;
    }
    g.V = gs;
    g.Vp = ps;
    int es = 1;
    for (int j = 3; j >= 0; --j) {
        const bool real = j >= r0;
        g.estride[j] = real ? es : 0;
        if (real) es *= g.L[j] + 3;
    }
    g.Ep = es;
    g.Eh = es / 2 + 2;
}

// Chooses the tile of a (sample, tile) unit: trailing axes whole, one axis cut into divisors, leading axes one
// site thick; minimises modelled cycles per output site subject to the shared-memory budget.
bool nd_plan(NdGeom& g, int N2, int G, int npass, int bdup, uint32_t budget, bool compact = false, bool sparse = false) {
    const int r0 = 4 - g.D;
    g.compact = compact ? 1 : 0;
    g.sparse = sparse ? 1 : 0;
    const bool odd = compact || sparse;
    { const char* e = getenv("NFK_ND_PREFETCH"); g.prefetch = !(e && e[0] == '0'); }
    g.bdup = bdup;
    g.G = G;
    g.npass = npass;
    g.b_bytes = (uint32_t)g.taps * G * bdup * N2 * 16;
    auto align = [](uint32_t v) { return (v + 127u) & ~127u; };
    float best = 1e30f;
    NdGeom bg = g;
    bool found = false;
    for (int split = r0; split < 4; ++split) {
        for (int t = 1; t <= g.L[split]; ++t) {
            if (g.L[split] % t) continue;
            NdGeom c = g;
            int outputs = 1;
            for (int j = 0; j < 4; ++j) {
                const bool real = j >= r0;
                c.T[j] = !real ? 1 : (j < split ? 1 : (j == split ? t : g.L[j]));
                c.ntile[j] = g.L[j] / c.T[j];
                // compact: every box extent odd (T + 2 for odd T, T + 3 for even T), so that every box stride is odd
                c.box[j] = real ? c.T[j] + 2 + ((odd && c.T[j] % 2 == 0) ? 1 : 0) : 1;
                c.lo[j] = real ? 1 : 0;
                c.hi[j] = real ? c.T[j] : 0;
                outputs *= c.T[j];
            }
            c.bstride[3] = 1;
            for (int j = 2; j >= 0; --j) c.bstride[j] = c.bstride[j + 1] * c.box[j + 1];
            const long long nbox = (long long)c.bstride[0] * c.box[0];
            if (nbox > 8000) continue;
            c.nbox = (int)nbox;
            c.first = 0;
            c.span = 1;
            c.tiles_per_sample = 1;
            for (int j = 0; j < 4; ++j) {
                c.first += c.lo[j] * c.bstride[j];
                c.span += (c.hi[j] - c.lo[j]) * c.bstride[j];
                c.tiles_per_sample *= c.ntile[j];
                c.magic_box[j] = nd_magic(c.box[j]);
                c.magic_ntile[j] = nd_magic(c.ntile[j]);
            }
            c.last = c.first + c.span - 1;
            c.nt = odd ? ((c.last - c.first) / 2 + 1 + 127) / 128 : (c.span + 127) / 128;
            if (sparse) c.nt *= 2;                               // the output positions of box parity 0, then those of parity 1
            c.split = split;
            c.run_rec = c.bstride[split] * c.box[split];
            c.nruns = c.nbox / c.run_rec;
            c.comp_bytes = align((uint32_t)(c.nbox + 128) * 16);
            c.pb_bytes = align((uint32_t)((c.nbox + 1) / 2 + 136) * 16);
            const uint32_t a_bytes = sparse ? 2 * G * c.pb_bytes : (compact ? 4 * G * c.pb_bytes : 2 * G * c.comp_bytes);
            uint32_t off = 0;
            if ((long long)a_bytes + c.b_bytes > (long long)budget) continue;
            c.off_a = off; off += a_bytes;
            c.off_b = off; off = align(off + c.b_bytes);
            c.off_tab = off; off = align(off + 64 * 4);
            c.off_bar = off; off = align(off + (2 * kNdMaxSlots + 1) * 8 + 64);
            c.smem_bytes = off < kNdMinSmem ? kNdMinSmem : off;      // > half an SM: one CTA per SM owns all of TMEM
            if (c.smem_bytes > budget) continue;
            // Accumulation chains, in steps of nine MMAs (the taps of the two innermost axes for one channel group):
            // as many chains as still leave six accumulator slots in TMEM, and never more than three steps per chain
            // (measured: 27 MMAs are within the parity contract, 81 are not; 16^4 with nine chains and three slots ran
            // 10 % slower than with fewer chains and more slots).
            {
                const int steps = c.ngroups * G;
                const int by_slots = kNdTmemCols / (6 * N2), by_len = (steps + 2) / 3;
                c.nchunk = by_slots > by_len ? by_slots : by_len;
                if (c.nchunk > steps) c.nchunk = steps;
                if (c.nchunk * N2 > 256) c.nchunk = 256 / N2 > 0 ? 256 / N2 : 1;
                if (const char* e = getenv("NFK_ND_CHUNKS")) {
                    const int v = atoi(e);
                    if (v >= 1 && v * N2 <= kNdTmemCols && v <= steps) c.nchunk = v;
                }
                // (the kernel starts a chain every ceil(steps / nchunk) steps: make nchunk the number of chains that gives)
                const int len = (steps + c.nchunk - 1) / c.nchunk;
                c.nchunk = (steps + len - 1) / len;
            }
            c.nslots = kNdTmemCols / (N2 * c.nchunk);
            if (c.nslots > kNdMaxSlots) c.nslots = kNdMaxSlots;
            if (c.nslots < 2) continue;      // (slot -> issuer and slot -> epilogue half are fixed maps: any count works)
            const float eff = (float)outputs / (c.nt * 128.f) * (compact ? 0.5f : 1.f);
            const float cost = (sparse ? 0.5f : 1.f) * c.taps * G * nd_mma_cycles(N2) / (128.f * eff) + 0.15f * G * (float)c.nbox / outputs +
                               2500.f / outputs;
            if (cost < best) { best = cost; bg = c; found = true; }
        }
    }
    if (found) g = bg;
    return found;
}

struct NdProps { int sm_count, max_smem; };
const NdProps& nd_props() {
    static const NdProps props = [] {
        NdProps p{0, 0};
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&p.sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&p.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (p.sm_count <= 0) p.sm_count = 148;
        if (p.max_smem <= 0) p.max_smem = 227 * 1024;
        return p;
    }();
    return props;
}

bool nd_compact() {
    const char* e = getenv("NFK_ND_COMPACT");          // 0: the last layer evaluates every site (A/B)
    return !(e && e[0] == '0');
}

int nd_bdup() {
    const char* e = getenv("NFK_ND_BDUP");             // 2: duplicate the weight rows for the second K group
    return (e && e[0] == '2') ? 2 : 1;
}

template <int MODE, int KIND, int K, int INV>
int nd_launch(NdArgs a, cudaStream_t st) {
    const NdProps& pr = nd_props();
    if (ensure_dynamic_smem<convnd_tc_kernel<MODE, KIND, K, INV>>(pr.max_smem) != NFK_OK) return NFK_ECUDA;
    long long grid = pr.sm_count;
    const long long nunits = a.B * a.g.tiles_per_sample * a.g.npass;
    if (grid > nunits) grid = nunits;
    convnd_tc_kernel<MODE, KIND, K, INV><<<(unsigned)grid, kNdThreads, a.g.smem_bytes, st>>>(a);
    return check_launch();
}

template <int KIND, int K>
int nd_launch_final(const NdArgs& a, int inverse, cudaStream_t st) {
    return inverse ? nd_launch<1, KIND, K, 1>(a, st) : nd_launch<1, KIND, K, 0>(a, st);
}

int nd_hi_cols(int kind, int n_knots) {
    const int P = kind == 0 ? 2 : 3 * n_knots - 2;
    return (P + 7) / 8 * 8;
}

bool nd_lattice_ok(const nfk_lattice& lat) {
    if (lat.ndim < 2 || lat.ndim > 4) return false;
    long long v = 1;
    for (int d = 0; d < lat.ndim; ++d) {
        if (lat.shape[d] < 2 || (lat.shape[d] & 1)) return false;     // the checkerboard must wrap consistently
        v *= lat.shape[d] + 2;
    }
    return v < (1LL << 27);
}

bool nd_knots_ok(int kind, int n_knots) {
    return kind == 0 || n_knots == 4 || n_knots == 5 || n_knots == 6 || n_knots == 8 || n_knots == 10;
}

// hidden layers: output channels per pass (32 at most: 64 accumulator columns)
int nd_oc(int H) { return H < 32 ? H : 32; }

struct NdWorkspace {
    long long rec_bytes, rec2_bytes, img2_bytes, img3_bytes, total;
};
NdWorkspace nd_workspace(const NdGeom& g, int H, int kind, int n_knots, long long B, int bdup, bool compact) {
    auto al = [](long long v) { return (v + 255) / 256 * 256; };
    const int G = H / 8;
    NdWorkspace w;
    w.rec_bytes = al(B * G * g.Vp * 32LL);
    w.rec2_bytes = compact ? al(B * G * 4LL * g.Eh * 16LL) : w.rec_bytes;
    w.img2_bytes = al((long long)(H / nd_oc(H)) * g.taps * G * bdup * 2 * nd_oc(H) * 16);
    w.img3_bytes = al((long long)g.taps * G * bdup * 2 * nd_hi_cols(kind, n_knots) * 16);
    w.total = w.rec_bytes + w.rec2_bytes + w.img2_bytes + w.img3_bytes;
    return w;
}

bool nd_width_ok(int H) { return H == 8 || H == 16 || H == 32 || H == 64; }

}  // namespace

/* Workspace (bytes) nfk_fusednd_step needs for this problem, or a negative NFK_E* code when the step is outside
 * what the kernels cover (the caller then runs the layer-by-layer kernels). */
extern "C" int64_t nfk_fusednd_workspace(nfk_lattice lat, int H, int kind, int n_knots, int64_t B) {
    if (!nd_lattice_ok(lat) || !nd_width_ok(H) || (kind != 0 && kind != 1) || !nd_knots_ok(kind, n_knots)) return NFK_EUNSUPPORTED;
    NdGeom g{};
    nd_lattice(g, lat);
    const int bdup = nd_bdup();
    const uint32_t budget = (uint32_t)nd_props().max_smem;
    NdGeom g2 = g, g3 = g;
    bool compact = nd_compact();
    if (!nd_plan(g2, 2 * nd_oc(H), H / 8, H / nd_oc(H), bdup, budget)) return NFK_EUNSUPPORTED;
    if (compact && !nd_plan(g3, 2 * nd_hi_cols(kind, n_knots), H / 8, 1, bdup, budget, true)) compact = false;
    if (!compact && !nd_plan(g3, 2 * nd_hi_cols(kind, n_knots), H / 8, 1, bdup, budget)) return NFK_EUNSUPPORTED;
    return nd_workspace(g, H, kind, n_knots, B > 0 ? B : 1, bdup, compact).total;
}

static int nd_step_impl(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                        const float* w3, const float* b3, int H, int kind, nfk_rqs_params prm,
                        nfk_lattice lat, int mask_parity, int parity, int inverse,
                        const float* log_in, float* y, float* log_out, int64_t B,
                        void* workspace, int64_t workspace_bytes, void* stream,
                        float* save_h1, float* save_h2, float* save_out) {
    if (!x || !w1 || !w2 || !w3 || !y || !log_out || !workspace || x == y) return NFK_EINVAL;
    if (!nd_width_ok(H) || (kind != 0 && kind != 1) || !nd_lattice_ok(lat)) return NFK_EUNSUPPORTED;
    if (!nd_knots_ok(kind, prm.n_knots)) return NFK_EUNSUPPORTED;
    if (B <= 0) return NFK_OK;
    if (kind == 1) {
        if (!(prm.xlim1 > prm.xlim0) || !(prm.ylim1 > prm.ylim0)) return NFK_EINVAL;
        if ((prm.extrap_left != NFK_EXTRAP_NONE && prm.extrap_left != NFK_EXTRAP_LINEAR) ||
            (prm.extrap_right != NFK_EXTRAP_NONE && prm.extrap_right != NFK_EXTRAP_LINEAR)) return NFK_EINVAL;
    }
    cudaStream_t st = NFK_STREAM(stream);
    const int D = lat.ndim, G = H / 8, OC = nd_oc(H), npass = H / OC;
    const int bdup = nd_bdup();
    const uint32_t budget = (uint32_t)nd_props().max_smem;
    NdGeom g{};
    nd_lattice(g, lat);
    g.mask_parity = mask_parity;
    g.active_val = parity == 0 ? 1 : 0;
    NdGeom g2 = g, g3 = g;
    const int NH3 = nd_hi_cols(kind, prm.n_knots);
    bool compact = nd_compact();
    if (!nd_plan(g2, 2 * OC, G, npass, bdup, budget)) return NFK_EUNSUPPORTED;
    if (compact && !nd_plan(g3, 2 * NH3, G, 1, bdup, budget, true)) compact = false;
    if (!compact && !nd_plan(g3, 2 * NH3, G, 1, bdup, budget)) return NFK_EUNSUPPORTED;
    g2.out_split = compact ? 1 : 0;
    if (getenv("NFK_ND_DEBUG")) {                      // tile plan of the two tensor-core layers (tuning aid)
        for (const NdGeom* q : {&g2, &g3})
            fprintf(stderr, "nfk_fusednd: T = %d %d %d %d  box %d  tiles/unit %d (span %d)  runs %d x %d rec  G %d  passes %d  "
                    "chains %d  slots %d  smem %u B%s\n", q->T[0], q->T[1], q->T[2], q->T[3], q->nbox, q->nt, q->span, q->nruns,
                    q->run_rec, q->G, q->npass, q->nchunk, q->nslots, q->smem_bytes, q->compact ? "  (active sites only)" : "");
    }
    const NdWorkspace ws = nd_workspace(g, H, kind, prm.n_knots, B, bdup, compact);
    if (workspace_bytes < ws.total || ((uintptr_t)workspace % 256) != 0) return NFK_EINVAL;
    uint8_t* wsp = static_cast<uint8_t*>(workspace);
    uint4* h1 = reinterpret_cast<uint4*>(wsp);
    uint4* h2 = reinterpret_cast<uint4*>(wsp + ws.rec_bytes);
    __half* img2 = reinterpret_cast<__half*>(wsp + ws.rec_bytes + ws.rec2_bytes);
    __half* img3 = reinterpret_cast<__half*>(wsp + ws.rec_bytes + ws.rec2_bytes + ws.img2_bytes);
    const int P = kind == 0 ? 2 : 3 * prm.n_knots - 2;

    nd_prep_weights_kernel<<<64, 256, 0, st>>>(w2, H, H, g.taps, OC, npass, bdup, img2);
    if (int e = check_launch()) return e;
    nd_prep_weights_kernel<<<64, 256, 0, st>>>(w3, P, H, g.taps, NH3, 1, bdup, img3);
    if (int e = check_launch()) return e;

    NdLat nl{};
    for (int d = 0; d < 4; ++d) {
        const int j = d + 4 - D;                      // NdLat keeps the lattice axes first
        nl.L[d] = d < D ? g.L[j] : 1;
        nl.gstride[d] = d < D ? g.gstride[j] : 0;
        nl.pstride[d] = d < D ? g.pstride[j] : 0;
        nl.magic_L[d] = nd_magic(nl.L[d]);
    }
    nl.V = g.V;
    nl.Vp = g.Vp;
    const int pairs = nl.V / 2;
    const long long blocks = B * ((pairs + 255) / 256);
    if (blocks >= (1LL << 31)) return NFK_EUNSUPPORTED;
    const dim3 grid1((unsigned)blocks, (unsigned)G);
    float* yf = compact ? y : nullptr;            // (the full last layer writes every site itself)
    switch (D) {
        case 2: nd_layer1_kernel<2><<<grid1, 256, 0, st>>>(x, w1, b1, h1, save_h1, yf, nl, mask_parity, g.active_val, B); break;
        case 3: nd_layer1_kernel<3><<<grid1, 256, 0, st>>>(x, w1, b1, h1, save_h1, yf, nl, mask_parity, g.active_val, B); break;
        default: nd_layer1_kernel<4><<<grid1, 256, 0, st>>>(x, w1, b1, h1, save_h1, yf, nl, mask_parity, g.active_val, B); break;
    }
    if (int e = check_launch()) return e;

    NdArgs a2{};
    a2.in_rec = h1; a2.out_rec = h2; a2.bimg = img2; a2.bias = b2; a2.B = B; a2.g = g2;
    a2.save = save_h2; a2.save_ch = H;
    a2.cfg = RqsCfg{0.f, 1.f, 0.f, 1.f, 0, 0};
    int e2;
    if (OC == 8) e2 = nd_launch<0, 0, 8, 0>(a2, st);
    else if (OC == 16) e2 = nd_launch<0, 0, 16, 0>(a2, st);
    else e2 = nd_launch<0, 0, 32, 0>(a2, st);
    if (e2) return e2;

    init_log_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(log_in, log_out, B);
    if (int e = check_launch()) return e;

    NdArgs a3{};
    a3.in_rec = h2; a3.bimg = img3; a3.bias = b3; a3.x = x; a3.y = y; a3.log_out = log_out; a3.B = B; a3.g = g3;
    a3.save = save_out; a3.save_ch = P;
    a3.cfg = RqsCfg{0.f, 1.f, 0.f, 1.f, 0, 0};
    if (kind == 0) return nd_launch_final<0, 2>(a3, inverse, st);
    a3.cfg = RqsCfg{prm.xlim0, prm.xlim1 - prm.xlim0, prm.ylim0, prm.ylim1 - prm.ylim0, prm.extrap_left,
                    prm.extrap_right};
    switch (prm.n_knots) {
        case 4: return nd_launch_final<1, 4>(a3, inverse, st);
        case 5: return nd_launch_final<1, 5>(a3, inverse, st);
        case 6: return nd_launch_final<1, 6>(a3, inverse, st);
        case 8: return nd_launch_final<1, 8>(a3, inverse, st);
        case 10: return nd_launch_final<1, 10>(a3, inverse, st);
        default: return NFK_EUNSUPPORTED;
    }
}

extern "C" int nfk_fusednd_step(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                                const float* w3, const float* b3, int H, int kind, nfk_rqs_params prm,
                                nfk_lattice lat, int mask_parity, int parity, int inverse,
                                const float* log_in, float* y, float* log_out, int64_t B,
                                void* workspace, int64_t workspace_bytes, void* stream) {
    return nd_step_impl(x, w1, b1, w2, b2, w3, b3, H, kind, prm, lat, mask_parity, parity, inverse, log_in, y, log_out, B,
                        workspace, workspace_bytes, stream, nullptr, nullptr, nullptr);
}

/* The same step as the forward pass of TRAINING: additionally stores the post-activation hidden layers
 * h1, h2 [B][H][V] and the conditioner output out [B][P][V] (defined at the active sites only) in float32 for the
 * gradient kernels (nfk_rqs_bwd / nfk_affine_bwd, nfk_conv_circ_fwd with transposed weights, nfk_conv_circ_bwd_weight). */
extern "C" int nfk_fusednd_step_train(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                                      const float* w3, const float* b3, int H, int kind, nfk_rqs_params prm,
                                      nfk_lattice lat, int mask_parity, int parity,
                                      const float* log_in, float* y, float* log_out,
                                      float* h1, float* h2, float* out, int64_t B,
                                      void* workspace, int64_t workspace_bytes, void* stream) {
    if (!h1 || !h2 || !out) return NFK_EINVAL;
    return nd_step_impl(x, w1, b1, w2, b2, w3, b3, H, kind, prm, lat, mask_parity, parity, 0, log_in, y, log_out, B,
                        workspace, workspace_bytes, stream, h1, h2, out);
}

// ------------------------------------------------------------------------------------------- data gradient
namespace {

struct NdDgradPlan {
    NdGeom g;
    int OC, npass, G, bdup;
    long long img_bytes, rec_bytes, total;
};

int nd_dgrad_plan(NdDgradPlan& p, nfk_lattice lat, int Co, int Ci, long long B, bool sparse = false) {
    if (!nd_width_ok(Ci) || Co < 1 || Co > 64 || !nd_lattice_ok(lat)) return NFK_EUNSUPPORTED;
    auto al = [](long long v) { return (v + 255) / 256 * 256; };
    p.OC = nd_oc(Ci);
    p.npass = Ci / p.OC;
    p.G = (Co + 7) / 8;
    p.bdup = nd_bdup();
    p.g = NdGeom{};
    nd_lattice(p.g, lat);
    if (!nd_plan(p.g, 2 * p.OC, p.G, p.npass, p.bdup, (uint32_t)nd_props().max_smem, false, sparse)) return NFK_EUNSUPPORTED;
    p.img_bytes = al((long long)p.npass * p.g.taps * p.G * p.bdup * 2 * p.OC * 16);
    p.rec_bytes = sparse ? al((B > 0 ? B : 1) * p.G * 2LL * p.g.Eh * 16LL) : al((B > 0 ? B : 1) * p.G * p.g.Vp * 32LL);
    p.total = 256 + p.img_bytes + p.rec_bytes;
    return NFK_OK;
}

// the checkerboard-sparse form where it can be planned (NFK_DGRAD_SPARSE=0: never), the dense one otherwise
int nd_dgrad_plan_for(NdDgradPlan& p, nfk_lattice lat, int Co, int Ci, long long B, int g_parity) {
    const char* e = getenv("NFK_DGRAD_SPARSE");
    if (g_parity >= 0 && !(e && e[0] == '0') && nd_dgrad_plan(p, lat, Co, Ci, B, true) == NFK_OK) {
        p.g.gpar = g_parity & 1;
        return NFK_OK;
    }
    return nd_dgrad_plan(p, lat, Co, Ci, B, false);
}

}  // namespace

extern "C" int64_t nfk_convnd_dgrad_workspace(nfk_lattice lat, int Co, int Ci, int64_t B) {
    NdDgradPlan p, q;
    if (int e = nd_dgrad_plan(p, lat, Co, Ci, B)) return e;
    if (nd_dgrad_plan(q, lat, Co, Ci, B, true) == NFK_OK && q.total > p.total) return q.total;      // (either form fits)
    return p.total;
}

/* Data gradient of one circular 3^D convolution layer of a ConvAct stack on the tensor cores
 * (autograd of modules.py:131-145): gin[b][ci][s] = act'(h[b][ci][s]) * sum_{co, t} w[co][ci][t] gpre[b][co][s - t],
 * act' = 1 - h^2 for the tanh layer below (h = its post-activation output) or 1 when `h` is NULL.
 * gpre [B][Co][V], w [Co][Ci][3^D], h / gin [B][Ci][V]; Ci in {8, 16, 32, 64}, Co <= 64, even extents, 2-D .. 4-D.
 * The input is packed into fp16-pair records scaled by a power of two taken from its largest magnitude (found on the
 * device), so the result has float32-level accuracy at any gradient magnitude. */
namespace {

// max |g| -> *amax (device), then g -> padded fp16-pair records scaled by nd_grad_scale(*amax); shared by the data and the
// weight gradient of a layer
int nd_amax(const float* gpre, long long n, unsigned* amax, cudaStream_t st) {
    if (cudaMemsetAsync(amax, 0, 4, st) != cudaSuccess) return NFK_ECUDA;
    long long ablocks = (n / 4 + 255) / 256;
    if (ablocks > 148 * 16) ablocks = 148 * 16;
    if (ablocks < 1) ablocks = 1;
    nd_amax_kernel<<<(unsigned)ablocks, 256, 0, st>>>(gpre, n, amax);
    return check_launch();
}
int nd_pack_gradient(const float* gpre, int Co, const nfk_lattice& lat, const int (&L)[4], const int (&pstride)[4], int V, int Vp,
                     int64_t B, unsigned* amax, uint4* rec, cudaStream_t st) {
    if (int e = nd_amax(gpre, (long long)B * Co * V, amax, st)) return e;
    return nd_pack(gpre, Co, reinterpret_cast<const float*>(amax), rec, lat.ndim, L, pstride, V, Vp, B, st);
}

int nd_dgrad_run(const NdDgradPlan& p, const uint4* rec, const float* amax, const float* w, const float* h, float* gin,
                 int Co, int Ci, int64_t B, __half* img, cudaStream_t st) {
    const NdGeom& g = p.g;
    if (getenv("NFK_ND_DEBUG"))
        fprintf(stderr, "nfk_convnd_dgrad: T = %d %d %d %d  box %d  tiles/unit %d  G %d  passes %d  chains %d  slots %d  smem %u B%s\n",
                g.T[0], g.T[1], g.T[2], g.T[3], g.nbox, g.nt, g.G, g.npass, g.nchunk, g.nslots, g.smem_bytes,
                g.sparse ? "  (checkerboard-sparse input)" : "");
    nd_prep_weights_t_kernel<<<64, 256, 0, st>>>(w, Co, Ci, g.taps, p.OC, p.npass, p.bdup, img);
    if (int e = check_launch()) return e;
    NdArgs a{};
    a.in_rec = rec; a.bimg = img; a.B = B; a.g = g;
    a.save = gin; a.save_ch = Ci; a.act = h; a.amax = amax;
    a.cfg = RqsCfg{0.f, 1.f, 0.f, 1.f, 0, 0};
    if (p.OC == 8) return nd_launch<2, 0, 8, 0>(a, st);
    if (p.OC == 16) return nd_launch<2, 0, 16, 0>(a, st);
    return nd_launch<2, 0, 32, 0>(a, st);
}

}  // namespace

extern "C" int nfk_convnd_dgrad(const float* gpre, int g_parity, const float* w, const float* h, float* gin, int Co, int Ci,
                                nfk_lattice lat, int64_t B, void* workspace, int64_t workspace_bytes, void* stream) {
    if (!gpre || !w || !gin || !workspace) return NFK_EINVAL;
    NdDgradPlan p;
    if (int e = nd_dgrad_plan_for(p, lat, Co, Ci, B, g_parity)) return e;
    if (B <= 0) return NFK_OK;
    if (workspace_bytes < p.total || ((uintptr_t)workspace % 256) != 0) return NFK_EINVAL;
    cudaStream_t st = NFK_STREAM(stream);
    uint8_t* wsp = static_cast<uint8_t*>(workspace);
    unsigned* amax = reinterpret_cast<unsigned*>(wsp);
    __half* img = reinterpret_cast<__half*>(wsp + 256);
    uint4* rec = reinterpret_cast<uint4*>(wsp + 256 + p.img_bytes);
    const float* am = reinterpret_cast<const float*>(amax);
    if (p.g.sparse) {
        if (int e = nd_amax(gpre, (long long)B * Co * p.g.V, amax, st)) return e;
        if (int e = nd_pack_active(gpre, Co, am, rec, lat.ndim, p.g.L, p.g.estride, p.g.V, p.g.Eh, p.g.gpar, B, st)) return e;
    } else {
        if (int e = nd_pack_gradient(gpre, Co, lat, p.g.L, p.g.pstride, p.g.V, p.g.Vp, B, amax, rec, st)) return e;
    }
    return nd_dgrad_run(p, rec, am, w, h, gin, Co, Ci, B, img, st);
}

// ------------------------------------------------------------------------------------------- weight gradient
// gw[co][ci][t] = sum_{b, s} gpre[b][co][s] h[b][ci][s + t - 1]  (adjoint of convNd.py:84-127) as tensor-core GEMMs with
// the SITES as the K dimension, read straight from the site-major records without an im2col: eight consecutive records
// (8 sites x 8 channels x fp16) are one core matrix of an MN-MAJOR no-swizzle operand (checked on hardware,
// scratch/mn_probe.cu: SBO = stride between MN blocks, LBO = stride between K blocks of 8 sites; an M = 64 accumulator
// keeps row i in TMEM lane 32 (i / 16) + i % 16).
//   A (M = 64): the planes [channel group of gpre][hi | lo] of a tile, block stride = plane stride
//   B (N = 24): the three taps of the innermost axis = the SAME records of h started one record apart (block stride =
//               16 bytes, overlapping core matrices); the taps of the outer axes move the start by box strides
//   D (64 x 24 per (outer tap group, hi | lo of h)): TMEM columns, accumulated over the K steps (16 sites each) of a
//               chain of units, then added into float32 accumulators in shared memory by the epilogue warps (the tensor
//               core truncates when it accumulates: chains stay at ~32 steps), atomics into gw once per CTA and pass.
// The bias gradient is one more MMA per K step against a record of ones.
namespace {

constexpr int kWgIssuers = 4;            // MMA-issuing warps: the outer tap groups are dealt among them (one thread's issue
                                         // loop, ~50 cycles per MMA, is slower than an M64 x N24 x K16 MMA)
constexpr int kWgThreads = 32 * (4 + kWgIssuers);   // + 4 epilogue warps (TMEM lane quarters = gpre channel groups)
constexpr int kWgGroupCols = 48;         // accumulator columns of one outer tap group: [dx][ci] for h hi, then for h lo
constexpr int kWgMaxGroups = 9;          // outer tap groups per pass (432 columns + 8 for the bias)
constexpr int kWgChainSteps = 64;        // K steps accumulated in TMEM before the float32 flush

struct WgGeom {
    int D, taps, ngroups, npass, gpp;
    int L[4], T[4], ntile[4], box[4], bstride[4], pstride[4];     // dummy axes: extent 1, strides 0
    uint32_t magic_ntile[4], magic_box[4];
    int nbox, tiles_per_sample, V, Vp, split, nruns, run_rec;
    int rows, row_len, tile_sites;
    int Gg, Co;
    int chain_units;
    int acc_stride;                      // floats per row of the shared accumulators (odd: the 16 rows of a warp on different banks)
    int nsets;                           // 2: two accumulator column sets in TMEM (the flush of a chain overlaps the next chain)
    int toff[27];                        // box offset of an outer tap group
    uint32_t hplane_bytes, gplane_bytes, buf_bytes, off_acc, off_ones, off_bar, smem_bytes;
};

struct WgArgs {
    const uint4* h_rec;                  // [B][1][2][Vp] padded records of the layer's input (8 channels)
    const uint4* g_rec;                  // [B][Gg][2][Vp] padded records of d loss / d pre-activation, scaled
    const float* amax;
    float* gw;                           // [Co][8][taps], accumulated into
    float* gbias;                        // [Co] or NULL
    long long B;
    WgGeom g;
};

#ifndef NFK_WG_M
#define NFK_WG_M 64
#endif
constexpr int kWgM = NFK_WG_M;           // 128: rows 64 .. 127 are junk (the planes that follow in shared memory); A/B of the MMA cost
__device__ __forceinline__ uint32_t wg_idesc(int N) {      // f16 x f16 -> f32, A and B MN-major
    return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kWgM >> 4) << 24);
}

// Use n (0, 1, ..) of a barrier is waited for with parity n & 1; every waiter sees the completions of its barrier in
// order and never falls two behind (an issuer cannot start a chain in an accumulator set before that set is flushed).
__global__ void __launch_bounds__(kWgThreads, 1) convnd_wgrad_tc_kernel(const WgArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const WgGeom& g = a.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* acc_s = reinterpret_cast<float*>(smem + g.off_acc);
    uint64_t* loaded = reinterpret_cast<uint64_t*>(smem + g.off_bar);       // [2] box of a unit has arrived
    uint64_t* mma_done = loaded + 2;                                        // [2] the MMAs reading a buffer have completed
    uint64_t* chain_done = mma_done + 2;                                    // [2] the MMAs of a chain (accumulator set) have completed
    uint64_t* flushed = chain_done + 2;                                     // [2] the epilogue has drained an accumulator set
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(flushed + 2);
    const int AS = g.acc_stride;

    for (int i = tid; i < 64 * AS; i += kWgThreads) acc_s[i] = 0.f;
    if (tid < 64) reinterpret_cast<__half*>(smem + g.off_ones)[tid] = __float2half_rn(1.f);     // 8 sites x 8 "channels" of ones
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(tc::smem_u32(loaded + i), 1);
            tc::mbar_init(tc::smem_u32(mma_done + i), kWgIssuers);
            tc::mbar_init(tc::smem_u32(chain_done + i), kWgIssuers);
            tc::mbar_init(tc::smem_u32(flushed + i), 4);
        }
        tc::fence_mbar_init();
    }
    if (warp == 4) tc::tmem_alloc(tc::smem_u32(tmem_slot), 512);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const long long nunits = a.B * g.tiles_per_sample;
    const long long mine = nunits > blockIdx.x ? (nunits - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long chains = (mine + g.chain_units - 1) / g.chain_units;
    const float inv_scale = 1.f / nd_grad_scale(NFK_LDG(a.amax));
    const int set_cols = g.nsets == 2 ? 256 : 0;

    for (int pass = 0; pass < g.npass; ++pass) {
        const int g0 = pass * g.gpp;
        const int ng = g.ngroups - g0 < g.gpp ? g.ngroups - g0 : g.gpp;
        const bool with_bias = pass == 0 && a.gbias != nullptr;
        const int ncols = ng * kWgGroupCols + (with_bias ? 8 : 0);
        if (warp >= 4) {
            const int iw = warp - 4;
            if (tc::elect_one()) {
                const uint32_t base = tc::smem_u32(smem);
                const uint32_t id24 = wg_idesc(24), id8 = wg_idesc(8);
                const uint64_t o_desc = tc::make_desc(tc::smem_u32(smem + g.off_ones), 0, 0);
                const bool bias_here = with_bias && iw == kWgIssuers - 1;
                // this issuer's outer tap groups (at most three of the nine of a pass), their box offsets and accumulator columns
                int n_mine = 0;
                uint32_t off_mine[3] = {0, 0, 0}, col_mine[3] = {0, 0, 0};
                for (int tg = iw; tg < ng && n_mine < 3; tg += kWgIssuers) {
                    off_mine[n_mine] = (uint32_t)g.toff[g0 + tg];
                    col_mine[n_mine] = (uint32_t)(tg * kWgGroupCols);
                    ++n_mine;
                }
                auto load_unit = [&](long long k, int buf) {
                    const long long unit = blockIdx.x + k * gridDim.x;
                    const long long b = unit / g.tiles_per_sample;
                    int trem = (int)(unit - b * g.tiles_per_sample);
                    int org[4];
#pragma unroll
                    for (int d = 3; d >= 0; --d) {
                        const int q = nd_div(trem, g.ntile[d], g.magic_ntile[d]);
                        org[d] = (trem - q * g.ntile[d]) * g.T[d];
                        trem = q;
                    }
                    const uint32_t bar = tc::smem_u32(loaded + buf);
                    const uint32_t hdst = base + buf * g.buf_bytes, gdst = hdst + 2 * g.hplane_bytes;
                    nd_expect_tx(bar, 2u * (uint32_t)g.nbox * 16u + 2u * g.Gg * (uint32_t)g.tile_sites * 16u);
                    // the input's tile + halo: contiguous runs of the padded array
                    const uint32_t run_bytes = (uint32_t)g.run_rec * 16;
                    for (int k2 = 0; k2 < g.nruns; ++k2) {
                        int rem = k2, so = 0;
#pragma unroll
                        for (int d = 3; d >= 0; --d) {
                            if (d < g.split) {
                                const int q = nd_div(rem, g.box[d], g.magic_box[d]);
                                so += (org[d] + rem - q * g.box[d]) * g.pstride[d];
                                rem = q;
                            } else {
                                so += org[d] * g.pstride[d];
                            }
                        }
                        const uint4* src = a.h_rec + b * 2LL * g.Vp + so;
                        nd_bulk_load(hdst + k2 * run_bytes, src, run_bytes, bar);
                        nd_bulk_load(hdst + g.hplane_bytes + k2 * run_bytes, src + g.Vp, run_bytes, bar);
                    }
                    // the gradient's interior rows, compact
                    const uint32_t row_bytes = (uint32_t)g.row_len * 16;
                    uint32_t dst = gdst;
                    for (int c0 = 0; c0 < g.T[0]; ++c0)
                        for (int c1 = 0; c1 < g.T[1]; ++c1)
                            for (int c2 = 0; c2 < g.T[2]; ++c2) {
                                const int so = (org[0] + c0 + 1) * g.pstride[0] + (org[1] + c1 + 1) * g.pstride[1] +
                                               (org[2] + c2 + 1) * g.pstride[2] + g.pstride[3];
                                const uint4* src = a.g_rec + b * 2LL * g.Gg * g.Vp + so;
                                for (int p = 0; p < 2 * g.Gg; ++p)
                                    nd_bulk_load(dst + p * g.gplane_bytes, src + (long long)p * g.Vp, row_bytes, bar);
                                dst += row_bytes;
                            }
                };
                const long long ku0 = pass * mine;                 // units walked before this pass
                if (iw == 0 && mine > 0) load_unit(0, (int)(ku0 & 1));
                for (long long k = 0; k < mine; ++k) {
                    const long long ku = ku0 + k;
                    const int buf = (int)(ku & 1);
                    if (iw == 0 && k + 1 < mine) {
                        // the other buffer was read by unit ku - 1
                        if (ku >= 1) tc::mbar_wait(tc::smem_u32(mma_done + (buf ^ 1)), (uint32_t)(((ku - 1) >> 1) & 1));
                        load_unit(k + 1, buf ^ 1);
                    }
                    tc::mbar_wait(tc::smem_u32(loaded + buf), (uint32_t)((ku >> 1) & 1));
                    const bool chain_start = (k % g.chain_units) == 0;
                    const long long cn = pass * chains + k / g.chain_units;
                    const int set = (int)(cn % g.nsets);
                    if (chain_start && cn >= g.nsets)
                        tc::mbar_wait(tc::smem_u32(flushed + set), (uint32_t)((cn / g.nsets - 1) & 1));
                    tc::fence_after_sync();
                    const uint32_t hbase = base + buf * g.buf_bytes, gbase = hbase + 2 * g.hplane_bytes;
                    // descriptors as (constant high word, running low word): every start address stays inside shared
                    // memory, so the 14-bit address field never carries into the stride fields
                    const uint64_t a_desc = tc::make_desc(gbase, 128, g.gplane_bytes);
                    const uint64_t bh_desc = tc::make_desc(hbase, 128, 16);
                    const uint64_t a_hi = a_desc & 0xFFFFFFFF00000000ULL, b_hi = bh_desc & 0xFFFFFFFF00000000ULL;
                    const uint32_t lo_plane = g.hplane_bytes >> 4;
                    const uint32_t acc0 = tmem + set * set_cols;
                    uint32_t accf = chain_start ? 0u : 1u;
                    uint32_t a_lo = (uint32_t)a_desc;
                    const int nsteps = g.row_len >> 4;
                    for (int c0 = 0; c0 < g.T[0]; ++c0)
                        for (int c1 = 0; c1 < g.T[1]; ++c1)
                            for (int c2 = 0; c2 < g.T[2]; ++c2) {
                                const uint32_t b_row = (uint32_t)bh_desc + (c0 + 1) * g.bstride[0] + (c1 + 1) * g.bstride[1] +
                                                       (c2 + 1) * g.bstride[2];
                                uint32_t b0 = b_row + off_mine[0], b1 = b_row + off_mine[1], b2 = b_row + off_mine[2];
#pragma unroll 1
                                for (int i = 0; i < nsteps; ++i) {
                                    const uint64_t ad = a_hi | a_lo;
                                    if (n_mine > 0) {
                                        tc::mma_f16(acc0 + col_mine[0], ad, b_hi | b0, id24, accf);
                                        tc::mma_f16(acc0 + col_mine[0] + 24, ad, b_hi | (b0 + lo_plane), id24, accf);
                                    }
                                    if (n_mine > 1) {
                                        tc::mma_f16(acc0 + col_mine[1], ad, b_hi | b1, id24, accf);
                                        tc::mma_f16(acc0 + col_mine[1] + 24, ad, b_hi | (b1 + lo_plane), id24, accf);
                                    }
                                    if (n_mine > 2) {
                                        tc::mma_f16(acc0 + col_mine[2], ad, b_hi | b2, id24, accf);
                                        tc::mma_f16(acc0 + col_mine[2] + 24, ad, b_hi | (b2 + lo_plane), id24, accf);
                                    }
                                    if (bias_here) tc::mma_f16(acc0 + ng * kWgGroupCols, ad, o_desc, id8, accf);
                                    accf = 1u;
                                    a_lo += 16; b0 += 16; b1 += 16; b2 += 16;
                                }
                            }
                    tc::mma_commit(tc::smem_u32(mma_done + buf));
                    if ((k + 1) % g.chain_units == 0 || k + 1 == mine) tc::mma_commit(tc::smem_u32(chain_done + set));
                }
            }
            __syncwarp();
        } else {
            // epilogue: warp w owns TMEM lanes 32 w .. 32 w + 15 = rows 16 w .. 16 w + 15 (channel group w: 8 hi rows, 8 lo rows)
            const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
            // M = 64: row i of D lives in lane 32 (i / 16) + i % 16; M = 128: in lane i (rows 64 .. 127 are not ours)
            const bool valid = kWgM == 64 ? lane < 16 : warp < 2;
            float* row = acc_s + (kWgM == 64 ? warp * 16 + (lane & 15) : (warp & 1) * 32 + lane) * AS;
            for (long long c = 0; c < chains; ++c) {
                const long long cn = pass * chains + c;
                const int set = (int)(cn % g.nsets);
                tc::mbar_wait(tc::smem_u32(chain_done + set), (uint32_t)((cn / g.nsets) & 1));
                tc::fence_after_sync();
                const uint32_t col0 = lane_addr + set * set_cols;
                for (int c32 = 0; c32 < ncols; c32 += 32) {
                    float v[4][8];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (c32 + q * 8 < ncols) tc::tmem_ld8(col0 + c32 + q * 8, v[q]);
                    tc::tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (c32 + q * 8 < ncols) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) row[c32 + q * 8 + j] += v[q][j];
                            }
                    }
                }
                tc::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(tc::smem_u32(flushed + set));
            }
        }
        // ---- end of the pass: shared accumulators -> gw (all four hi / lo products), cleared for the next pass
        __syncthreads();
        const int per_co = ng * 24;
        for (int e = tid; e < g.Co * per_co; e += kWgThreads) {
            const int co = e / per_co, rest = e - co * per_co;
            const int tg = rest / 24, dc = rest - tg * 24;                      // dc = dx * 8 + ci
            const float* rh = acc_s + ((co >> 3) * 16 + (co & 7)) * AS + tg * kWgGroupCols;
            const float* rl = rh + 8 * AS;
            const float v = (rh[dc] + (rh[24 + dc] + rl[dc])) + rl[24 + dc];
            const int t = (g0 + tg) * 3 + (dc >> 3);
            atomicAdd(a.gw + ((long long)co * 8 + (dc & 7)) * g.taps + t, v * inv_scale);
        }
        if (with_bias)
            for (int co = tid; co < g.Co; co += kWgThreads) {
                const float* rh = acc_s + ((co >> 3) * 16 + (co & 7)) * AS + ng * kWgGroupCols;
                atomicAdd(a.gbias + co, (rh[0] + rh[8 * AS]) * inv_scale);
            }
        __syncthreads();
        if (pass + 1 < g.npass) {
            for (int i = tid; i < 64 * AS; i += kWgThreads) acc_s[i] = 0.f;
            __syncthreads();
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 4) tc::tmem_dealloc(tmem, 512);
}

// Tile of a unit: innermost axis whole (a multiple of 16 sites: K steps do not straddle rows), the axes before it as
// in nd_plan; the largest tile whose two buffers fit next to the accumulators.
bool wg_plan(WgGeom& g, const nfk_lattice& lat, int Co, uint32_t budget) {
    const int D = lat.ndim, r0 = 4 - D;
    g = WgGeom{};
    g.D = D;
    g.Co = Co;
    g.Gg = (Co + 7) / 8;
    g.taps = 1;
    for (int d = 0; d < D; ++d) g.taps *= 3;
    g.ngroups = g.taps / 3;
    g.gpp = g.ngroups < kWgMaxGroups ? g.ngroups : kWgMaxGroups;
    g.npass = (g.ngroups + g.gpp - 1) / g.gpp;
    int ps = 1;
    g.V = 1;
    for (int j = 3; j >= 0; --j) {
        const bool real = j >= r0;
        g.L[j] = real ? lat.shape[j - r0] : 1;
        g.pstride[j] = real ? ps : 0;
        if (real) { ps *= g.L[j] + 2; g.V *= g.L[j]; }
    }
    g.Vp = ps;
    if (g.L[3] % 16 != 0 || g.Gg > 4) return false;
    auto align = [](uint32_t v) { return (v + 127u) & ~127u; };
    bool found = false;
    WgGeom best{};
    for (int split = r0; split <= 2; ++split) {
        for (int t = 1; t <= g.L[split]; ++t) {
            if (g.L[split] % t) continue;
            WgGeom c = g;
            int bs = 1;
            c.tiles_per_sample = 1;
            c.rows = 1;
            for (int j = 3; j >= 0; --j) {
                const bool real = j >= r0;
                c.T[j] = !real ? 1 : (j < split ? 1 : (j == split ? t : g.L[j]));
                c.ntile[j] = g.L[j] / c.T[j];
                c.box[j] = real ? c.T[j] + 2 : 1;
                c.bstride[j] = real ? bs : 0;
                bs *= c.box[j];
                c.tiles_per_sample *= c.ntile[j];
                if (j < 3) c.rows *= c.T[j];
                c.magic_ntile[j] = nd_magic(c.ntile[j]);
                c.magic_box[j] = nd_magic(c.box[j]);
            }
            c.nbox = bs;
            c.split = split;
            c.run_rec = c.bstride[split] * c.box[split];
            c.nruns = c.nbox / c.run_rec;
            c.row_len = g.L[3];
            c.tile_sites = c.rows * c.row_len;
            c.hplane_bytes = align((uint32_t)(c.nbox + 8) * 16);
            c.gplane_bytes = align((uint32_t)c.tile_sites * 16);
            if (c.gplane_bytes / 16 > 0x3FFF || c.nbox > 0x3FFF) continue;
            c.buf_bytes = 2 * c.hplane_bytes + 2 * c.Gg * c.gplane_bytes;
            uint32_t off = 2 * c.buf_bytes;
            const int ncols_max = c.gpp * kWgGroupCols + 8;
            c.acc_stride = ncols_max + 1;
            c.nsets = 2 * ncols_max <= 512 && ncols_max <= 256 ? 2 : 1;
            c.off_acc = off; off = align(off + 64 * c.acc_stride * 4);
            c.off_ones = off; off += 128;
            c.off_bar = off; off = align(off + 8 * 8 + 16);
            c.smem_bytes = off < kNdMinSmem ? kNdMinSmem : off;
            if (c.smem_bytes > budget) continue;
            // the M = 64 operand always spans eight planes: the ones beyond 2 Gg must still lie inside the allocation
            if (c.buf_bytes + 2 * c.hplane_bytes + (kWgM / 8) * c.gplane_bytes > c.smem_bytes) continue;
            const int ksteps = c.tile_sites / 16;
            c.chain_units = ksteps >= kWgChainSteps ? 1 : kWgChainSteps / ksteps;
            for (int tg = 0; tg < c.ngroups; ++tg) {
                int rem = tg, o = 0;
                for (int j = 2; j >= r0; --j) { o += (rem % 3 - 1) * c.bstride[j]; rem /= 3; }
                c.toff[tg] = o;
            }
            if (!found || c.tile_sites > best.tile_sites) { best = c; found = true; }
        }
    }
    if (found) g = best;
    return found;
}

struct NdWgradPlan {
    WgGeom g;
    long long hrec_bytes, grec_bytes, total;
};

int nd_wgrad_plan(NdWgradPlan& p, nfk_lattice lat, int Co, int Ci, long long B) {
    if (Ci != 8 || Co < 1 || Co > 32 || !nd_lattice_ok(lat)) return NFK_EUNSUPPORTED;
    if (!wg_plan(p.g, lat, Co, (uint32_t)nd_props().max_smem)) return NFK_EUNSUPPORTED;
    auto al = [](long long v) { return (v + 255) / 256 * 256; };
    const long long b = B > 0 ? B : 1;
    p.hrec_bytes = al(b * p.g.Vp * 32LL);
    p.grec_bytes = al(b * p.g.Gg * p.g.Vp * 32LL);
    p.total = 256 + p.hrec_bytes + p.grec_bytes;
    return NFK_OK;
}

}  // namespace

extern "C" int64_t nfk_convnd_wgrad_workspace(nfk_lattice lat, int Co, int Ci, int64_t B) {
    NdWgradPlan p;
    if (int e = nd_wgrad_plan(p, lat, Co, Ci, B)) return e;
    return p.total;
}

/* Weight (and bias) gradient of one circular 3^D convolution layer with 8 input channels on the tensor cores:
 *     gw[co][ci][t] += sum_{b, s} gpre[b][co][s] * h[b][ci][s + t - 1],   gb[co] += sum_{b, s} gpre[b][co][s]
 * h [B][8][V] the layer's input, gpre [B][Co][V] d loss / d (its pre-activation output), Co <= 32; gw [Co][8][3^D] and gb
 * [Co] (or NULL) are ACCUMULATED into (zero them first).  2-D .. 4-D lattices with even extents and an innermost extent
 * that is a multiple of 16; NFK_EUNSUPPORTED otherwise (the caller then uses nfk_conv_circ_bwd_weight).            */
namespace {

int nd_wgrad_run(const NdWgradPlan& p, const uint4* hrec, const uint4* grec, const float* amax, float* gw, float* gb,
                 int64_t B, cudaStream_t st) {
    const WgGeom& g = p.g;
    if (getenv("NFK_ND_DEBUG"))
        fprintf(stderr, "nfk_convnd_wgrad: T = %d %d %d %d  box %d  rows %d x %d  runs %d x %d rec  groups %d in %d passes  chain %d units  smem %u B\n",
                g.T[0], g.T[1], g.T[2], g.T[3], g.nbox, g.rows, g.row_len, g.nruns, g.run_rec, g.ngroups, g.npass, g.chain_units, g.smem_bytes);
    WgArgs a{};
    a.h_rec = hrec; a.g_rec = grec; a.amax = amax; a.gw = gw; a.gbias = gb; a.B = B; a.g = g;
    const NdProps& pr = nd_props();
    if (ensure_dynamic_smem<convnd_wgrad_tc_kernel>(pr.max_smem) != NFK_OK) return NFK_ECUDA;
    long long grid = pr.sm_count;
    const long long nunits = B * g.tiles_per_sample;
    if (grid > nunits) grid = nunits;
    convnd_wgrad_tc_kernel<<<(unsigned)grid, kWgThreads, g.smem_bytes, st>>>(a);
    return check_launch();
}

}  // namespace

extern "C" int nfk_convnd_wgrad(const float* h, const float* gpre, float* gw, float* gb, int Co, int Ci,
                                nfk_lattice lat, int64_t B, void* workspace, int64_t workspace_bytes, void* stream) {
    if (!h || !gpre || !gw || !workspace) return NFK_EINVAL;
    NdWgradPlan p;
    if (int e = nd_wgrad_plan(p, lat, Co, Ci, B)) return e;
    if (B <= 0) return NFK_OK;
    if (workspace_bytes < p.total || ((uintptr_t)workspace % 256) != 0) return NFK_EINVAL;
    cudaStream_t st = NFK_STREAM(stream);
    const WgGeom& g = p.g;
    uint8_t* wsp = static_cast<uint8_t*>(workspace);
    unsigned* amax = reinterpret_cast<unsigned*>(wsp);
    uint4* hrec = reinterpret_cast<uint4*>(wsp + 256);
    uint4* grec = reinterpret_cast<uint4*>(wsp + 256 + p.hrec_bytes);
    if (int e = nd_pack_gradient(gpre, Co, lat, g.L, g.pstride, g.V, g.Vp, B, amax, grec, st)) return e;
    if (int e = nd_pack(h, 8, nullptr, hrec, lat.ndim, g.L, g.pstride, g.V, g.Vp, B, st)) return e;
    return nd_wgrad_run(p, hrec, grec, reinterpret_cast<const float*>(amax), gw, gb, B, st);
}

// ------------------------------------------------------------------------------------------- one layer, both gradients
namespace {
struct NdLayerBwdPlan {
    NdDgradPlan d;
    NdWgradPlan w;
    long long total;
};
// g_parity >= 0: the data gradient takes the checkerboard-sparse form (its own, smaller records); -2: size for either
int nd_layer_bwd_plan(NdLayerBwdPlan& p, nfk_lattice lat, int Co, int Ci, long long B, int g_parity) {
    if (int e = nd_dgrad_plan_for(p.d, lat, Co, Ci, B, g_parity == -2 ? -1 : g_parity)) return e;
    if (int e = nd_wgrad_plan(p.w, lat, Co, Ci, B)) return e;
    // dense gradient records are shared by both kernels; a sparse data gradient reads its own
    p.total = 256 + p.d.img_bytes + p.w.hrec_bytes + p.w.grec_bytes + (p.d.g.sparse ? p.d.rec_bytes : 0);
    if (g_parity == -2) {
        NdDgradPlan q;
        if (nd_dgrad_plan(q, lat, Co, Ci, B, true) == NFK_OK) p.total += q.rec_bytes;
    }
    return NFK_OK;
}
}  // namespace

extern "C" int64_t nfk_convnd_layer_bwd_workspace(nfk_lattice lat, int Co, int Ci, int64_t B) {
    NdLayerBwdPlan p;
    if (int e = nd_layer_bwd_plan(p, lat, Co, Ci, B, -2)) return e;
    return p.total;
}

/* nfk_convnd_wgrad and nfk_convnd_dgrad of ONE layer in one call: d loss / d pre-activation is reduced (max |g|) and packed
 * into records once for both kernels.  h_in [B][8][V] the layer's input (also the tensor whose tanh' the data gradient is
 * multiplied by when `act_below` != 0), gpre [B][Co][V]; gw / gb accumulated into, gin [B][8][V] written.  Applies where
 * both kernels do (8 input channels, Co <= 32, innermost extent a multiple of 16); NFK_EUNSUPPORTED otherwise.          */
extern "C" int nfk_convnd_layer_bwd(const float* h_in, const float* gpre, int g_parity, const float* w, int act_below,
                                    float* gin, float* gw, float* gb, int Co, int Ci, nfk_lattice lat, int64_t B,
                                    void* workspace, int64_t workspace_bytes, void* stream) {
    if (!h_in || !gpre || !w || !gin || !gw || !workspace) return NFK_EINVAL;
    NdLayerBwdPlan p;
    if (int e = nd_layer_bwd_plan(p, lat, Co, Ci, B, g_parity < 0 ? -1 : g_parity)) return e;
    if (B <= 0) return NFK_OK;
    if (workspace_bytes < p.total || ((uintptr_t)workspace % 256) != 0) return NFK_EINVAL;
    cudaStream_t st = NFK_STREAM(stream);
    const WgGeom& g = p.w.g;
    uint8_t* wsp = static_cast<uint8_t*>(workspace);
    unsigned* amax = reinterpret_cast<unsigned*>(wsp);
    __half* img = reinterpret_cast<__half*>(wsp + 256);
    uint4* hrec = reinterpret_cast<uint4*>(wsp + 256 + p.d.img_bytes);
    uint4* grec = reinterpret_cast<uint4*>(wsp + 256 + p.d.img_bytes + p.w.hrec_bytes);
    const float* am = reinterpret_cast<const float*>(amax);
    if (int e = nd_pack_gradient(gpre, Co, lat, g.L, g.pstride, g.V, g.Vp, B, amax, grec, st)) return e;
    if (int e = nd_pack(h_in, 8, nullptr, hrec, lat.ndim, g.L, g.pstride, g.V, g.Vp, B, st)) return e;
    if (int e = nd_wgrad_run(p.w, hrec, grec, am, gw, gb, B, st)) return e;
    const uint4* drec = grec;
    if (p.d.g.sparse) {
        uint4* srec = reinterpret_cast<uint4*>(wsp + 256 + p.d.img_bytes + p.w.hrec_bytes + p.w.grec_bytes);
        if (int e = nd_pack_active(gpre, Co, am, srec, lat.ndim, p.d.g.L, p.d.g.estride, p.d.g.V, p.d.g.Eh, p.d.g.gpar, B, st)) return e;
        drec = srec;
    }
    return nd_dgrad_run(p.d, drec, am, w, act_below ? h_in : nullptr, gin, Co, Ci, B, img, st);
}
