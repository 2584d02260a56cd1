// nfk_fused.cu -- launch side of the fused 2-D coupling step (see nfk_fused.cuh).
// One CTA per sample walks the lattice in strips of R rows; the per-sample log|det J| is
// reduced once at the end (warp shuffles + shared memory), so it is deterministic.

#include <stdlib.h>

#include "nfk_common.cuh"
#include "nfk_fused.cuh"

using namespace nfk;

#define NFK_STREAM(s) reinterpret_cast<cudaStream_t>(s)

struct FusedArgs {
    const float *x, *w1, *b1, *w2, *b2, *w3, *b3, *log_in;
    float *y, *log_out;
    FusedGeom g;
    FusedXform xf;
};

template <int KIND, int K>
__global__ void __launch_bounds__(288, 2) fused2d_kernel(FusedArgs a) {
    constexpr int P = KIND == 0 ? 2 : 3 * K - 2;
    constexpr int PP = (P + 3) / 4 * 4;
    extern __shared__ __align__(16) float smem[];
    const FusedGeom g = a.g;
    const FusedSmem m = fused_carve<PP>(smem, g.L1, g.R);
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < fused_weight_elems<PP>(); e += nt)
        fused_load_weight<P, PP>(m, a.w1, a.w2, a.w3, a.b1, a.b2, a.b3, e);
    const int64_t V = (int64_t)g.L0 * g.L1;
    const float* xb = a.x + blockIdx.x * V;
    float* yb = a.y + blockIdx.x * V;
    float lacc = 0.f;
    for (int r0 = 0; r0 < g.L0; r0 += g.R) {
        const int rows = g.L0 - r0 < g.R ? g.L0 - r0 : g.R;
        __syncthreads();                                   // weights visible / previous strip consumed
        for (int e = tid; e < (rows + 6) * g.WS; e += nt) fused_load_x(g, m, xb, r0, rows, e);
        __syncthreads();
        for (int it = tid; it < (rows + 4) * g.ncg; it += nt)
            fused_hidden_item<1>(g, m.xf, g.R + 6, m.w1s, m.b1s, m.h1s, g.R + 4, rows + 4, it);
        __syncthreads();
        for (int it = tid; it < (rows + 2) * g.ncg; it += nt)
            fused_hidden_item<kFH>(g, m.h1s, g.R + 4, m.w2s, m.b2s, m.h2s, g.R + 2, rows + 2, it);
        __syncthreads();
        for (int it = tid; it < rows * g.ncg; it += nt)
            lacc += fused_out_item<KIND, K>(g, m, a.xf, r0, rows, yb, it);
    }
    lacc = block_sum(lacc);
    if (tid == 0 && a.log_out) a.log_out[blockIdx.x] = (a.log_in ? a.log_in[blockIdx.x] : 0.f) + lacc;
}

template <int KIND, int K>
static int fused_launch(FusedArgs a, int64_t B, cudaStream_t st) {
    constexpr int P = KIND == 0 ? 2 : 3 * K - 2;
    constexpr int PP = (P + 3) / 4 * 4;
    int R = 16;
    if (R > a.g.L0) R = a.g.L0;
    while (R > 1 && fused_smem_floats<PP>(a.g.L1, R) * sizeof(float) > 100 * 1024) R /= 2;
    const size_t smem = fused_smem_floats<PP>(a.g.L1, R) * sizeof(float);
    if (smem > 200 * 1024) return NFK_EUNSUPPORTED;
    a.g.R = R;
    int threads = ((R + 2) * a.g.ncg + 31) / 32 * 32;
    if (threads > 288) threads = 288;
    if (threads < 32) threads = 32;
    if (ensure_dynamic_smem<fused2d_kernel<KIND, K>>(200 * 1024) != NFK_OK) return NFK_ECUDA;
    fused2d_kernel<KIND, K><<<(unsigned)B, threads, smem, st>>>(a);
    return check_launch();
}

#define NFK_FUSED_K(X) X(4) X(5) X(6) X(8) X(10) X(12) X(16)

namespace nfk {
int fused2d_tc_step(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                    const float* w3, const float* b3, int kind, const nfk_rqs_params& prm, int mask_parity,
                    int parity, int inverse, const float* log_in, float* y, float* log_out, int L0, int L1,
                    int64_t B, cudaStream_t st, float* save_h1, float* save_h2, float* save_out);
}

// NFK_FUSED_TC=0 in the environment keeps the CUDA-core kernel (A/B timing and debugging)
static bool tc_enabled() {
    const char* e = getenv("NFK_FUSED_TC");      // read per call: tests flip it at run time
    return !(e && e[0] == '0');
}

extern "C" int nfk_fused2d_step(const float* x, const float* w1, const float* b1, const float* w2,
                                const float* b2, const float* w3, const float* b3, int H, int kind,
                                nfk_rqs_params prm, int mask_parity, int parity, int inverse,
                                const float* log_in, float* y, float* log_out,
                                int L0, int L1, int64_t B, void* stream) {
    if (!x || !w1 || !w2 || !w3 || !y || x == y) return NFK_EINVAL;
    if (H != kFH || L0 < 1 || L1 < 1 || (kind != 0 && kind != 1)) return NFK_EUNSUPPORTED;
    if (B <= 0) return NFK_OK;
    if (kind == 1) {
        if (prm.n_knots < 2 || !(prm.xlim1 > prm.xlim0) || !(prm.ylim1 > prm.ylim0)) return NFK_EINVAL;
        if ((prm.extrap_left != NFK_EXTRAP_NONE && prm.extrap_left != NFK_EXTRAP_LINEAR) ||
            (prm.extrap_right != NFK_EXTRAP_NONE && prm.extrap_right != NFK_EXTRAP_LINEAR)) return NFK_EINVAL;
    }
    if (tc_enabled()) {        // conditioner on the tensor cores when the geometry allows it
        const int rc = fused2d_tc_step(x, w1, b1, w2, b2, w3, b3, kind, prm, mask_parity, parity, inverse, log_in, y,
                                       log_out, L0, L1, B, NFK_STREAM(stream), nullptr, nullptr, nullptr);
        if (rc != NFK_EUNSUPPORTED) return rc;
    }
    // CUDA-core kernel: rows are walked in groups of four columns with 128-bit accesses
    if (L1 < 4 || L1 % 4 != 0) return NFK_EUNSUPPORTED;
    if (((uintptr_t)y % 16) != 0) return NFK_EINVAL;
    FusedArgs a;
    a.x = x; a.w1 = w1; a.b1 = b1; a.w2 = w2; a.b2 = b2; a.w3 = w3; a.b3 = b3;
    a.log_in = log_in; a.y = y; a.log_out = log_out;
    a.g.L0 = L0; a.g.L1 = L1; a.g.WS = L1 + 4; a.g.ncg = L1 / 4; a.g.R = 0;
    a.g.mask_parity = mask_parity; a.g.active_val = parity == 0 ? 1 : 0;
    a.xf.inverse = inverse;
    a.xf.cfg = RqsCfg{0.f, 1.f, 0.f, 1.f, 0, 0};
    cudaStream_t st = NFK_STREAM(stream);
    if (kind == 0) return fused_launch<0, 2>(a, B, st);
    if (prm.n_knots < 2 || !(prm.xlim1 > prm.xlim0) || !(prm.ylim1 > prm.ylim0)) return NFK_EINVAL;
    if ((prm.extrap_left != NFK_EXTRAP_NONE && prm.extrap_left != NFK_EXTRAP_LINEAR) ||
        (prm.extrap_right != NFK_EXTRAP_NONE && prm.extrap_right != NFK_EXTRAP_LINEAR)) return NFK_EINVAL;
    a.xf.cfg = RqsCfg{prm.xlim0, prm.xlim1 - prm.xlim0, prm.ylim0, prm.ylim1 - prm.ylim0,
                      prm.extrap_left, prm.extrap_right};
    switch (prm.n_knots) {
#define X(KK) case KK: return fused_launch<1, KK>(a, B, st);
        NFK_FUSED_K(X)
#undef X
        default: return NFK_EUNSUPPORTED;
    }
}

// Training forward of the same step: additionally stores the hidden layers and the conditioner
// output (channel-major fp32) that nfk_rqs_bwd / nfk_affine_bwd and the convolution gradient
// kernels consume.  Tensor-core kernel only: NFK_EUNSUPPORTED outside its geometry.
extern "C" int nfk_fused2d_step_train(const float* x, const float* w1, const float* b1, const float* w2,
                                      const float* b2, const float* w3, const float* b3, int H, int kind,
                                      nfk_rqs_params prm, int mask_parity, int parity,
                                      const float* log_in, float* y, float* log_out,
                                      float* h1, float* h2, float* out,
                                      int L0, int L1, int64_t B, void* stream) {
    if (!x || !w1 || !w2 || !w3 || !y || !h1 || !h2 || !out || x == y) return NFK_EINVAL;
    if (H != kFH || L0 < 1 || L1 < 1 || (kind != 0 && kind != 1)) return NFK_EUNSUPPORTED;
    if (B <= 0) return NFK_OK;
    if (kind == 1) {
        if (prm.n_knots < 2 || !(prm.xlim1 > prm.xlim0) || !(prm.ylim1 > prm.ylim0)) return NFK_EINVAL;
        if ((prm.extrap_left != NFK_EXTRAP_NONE && prm.extrap_left != NFK_EXTRAP_LINEAR) ||
            (prm.extrap_right != NFK_EXTRAP_NONE && prm.extrap_right != NFK_EXTRAP_LINEAR)) return NFK_EINVAL;
    }
    return fused2d_tc_step(x, w1, b1, w2, b2, w3, b3, kind, prm, mask_parity, parity, 0, log_in, y, log_out, L0, L1,
                           B, NFK_STREAM(stream), h1, h2, out);
}
