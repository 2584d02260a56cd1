// nfk_knots.cu -- SplineNet.make_spline (reference src/nn/scalar/modules.py:369-391) for one shared
// 1-D spline: raw parameters -> knot table, in ONE launch instead of ~25 tiny softmax / cumsum /
// cat launches (the parameter-only part of DistConvertor_, MeanFieldNet_ and IPSD; it dominated
// the launch-bound configurations).
//
//   table[0] = knots_x    = xlo + xw * sum_{i<j} p_i          p = softmax(weights_x)   (K-1 terms)
//   table[1] = knots_y    = ylo + yw * sum_{i<j} q_i          q = softmax(weights_y)
//   table[2] = knots_d    = softplus_{beta=ln2}(weights_d)    or, smooth (weights_d == NULL), the mean
//                           of the adjacent bin slopes s_i = (yw q_i)/(xw p_i), end slopes at the ends
//   table[3] = xlo + xw - knots_x, accumulated from the right end (sum_{i>=j} p_i), exact 0 at the end
//   table[4] = ylo + yw - knots_y, likewise
//
// K is a few tens at most: one warp, double precision (exponentials strided over the lanes, running
// sums on lane 0).
#include "nfk_common.cuh"
#include "nfk_knots.cuh"

#define NFK_STREAM(s) reinterpret_cast<cudaStream_t>(s)

namespace nfk {

struct KnotWarpSync {
    __device__ void operator()() const { __syncwarp(); }
};

__global__ void knots_fwd_kernel(KnotArgs a, float* __restrict__ table) {
    __shared__ KnotScratch scratch;
    knots_fwd_body(a, table, scratch, (int)threadIdx.x, 32, KnotWarpSync{});
}

__global__ void knots_bwd_kernel(KnotArgs a, const float* __restrict__ g, float* __restrict__ gwx,
                                 float* __restrict__ gwy, float* __restrict__ gwd) {
    __shared__ KnotScratch scratch;
    knots_bwd_body(a, g, gwx, gwy, gwd, scratch, (int)threadIdx.x, 32, KnotWarpSync{});
}

}  // namespace nfk

using namespace nfk;

static bool knots_ok(const float* wx, const float* wy, int K, float xw, float yw) {
    return wx && wy && K >= 2 && K <= kMaxKnots && xw > 0.f && yw > 0.f;
}

extern "C" int nfk_knots_fwd(const float* wx, const float* wy, const float* wd, int K, float xlo, float xw,
                             float ylo, float yw, float* table, void* stream) {
    if (!knots_ok(wx, wy, K, xw, yw) || !table) return NFK_EINVAL;
    KnotArgs a{wx, wy, wd, K, xlo, xw, ylo, yw};
    knots_fwd_kernel<<<1, 32, 0, NFK_STREAM(stream)>>>(a, table);
    return check_launch();
}

extern "C" int nfk_knots_bwd(const float* wx, const float* wy, const float* wd, int K, float xlo, float xw,
                             float ylo, float yw, const float* g_table, float* gwx, float* gwy, float* gwd,
                             void* stream) {
    if (!knots_ok(wx, wy, K, xw, yw) || !g_table || !gwx || !gwy || (wd && !gwd)) return NFK_EINVAL;
    KnotArgs a{wx, wy, wd, K, xlo, xw, ylo, yw};
    knots_bwd_kernel<<<1, 32, 0, NFK_STREAM(stream)>>>(a, g_table, gwx, gwy, gwd);
    return check_launch();
}
