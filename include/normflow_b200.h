/*
 * normflow_b200.h -- C ABI of libnormflow_b200.so
 *
 * Hand-written sm_100a CUDA kernels for the data-parallel hot path of
 * jkomijani/normflow_ (Model.fit / posterior.sample / mcmc.sample inner loop).
 *
 * The reference is 100 % Python on PyTorch ops and defines NO FFI / plugin
 * interface; the drop-in boundary is its duck-typed Python
 * protocol, which normflow__b200/ mirrors.  This header is the native layer under
 * that mirror: every entry point cites the reference call site (file:line under
 * /root/reference/src) whose ATen operator sequence it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - All data pointers are DEVICE pointers to contiguous arrays, float32 unless
 *     stated (uint8 masks, int64 indices, float64 Metropolis inputs).
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued on it and
 *     the call returns immediately (no host sync, no allocation; CUDA-graph
 *     capturable).  The caller owns every buffer.
 *   - Return value: 0 on success, a negative NFK_E* code otherwise
 *     (nfk_strerror).  Nothing throws.
 *   - Field layout: x[B][V] with V = prod(lattice shape), row-major sites.
 *     Conditioner output: out[B][P][V] (channel-major per sample, NCHW-like).
 *   - mask: uint8[V]; mask[s] = 1 <=> site s belongs to partition 0
 *     (Mask._mask, mask/mask.py:23).  A coupling step with `parity` p updates the
 *     sites with mask[s] == (p == 0 ? 1 : 0)  (couplings_.py:56-64).
 *   - frozen_mode: what the step writes at the frozen (non-updated) sites:
 *       NFK_FROZEN_ZERO  (0)  y = 0      -- the reference's x_active -> fx_active
 *                                           dataflow (atomic_forward)
 *       NFK_FROZEN_COPY  (1)  y = x      -- full-field update (split/cat fused away)
 *   - log_in may be NULL (treated as zeros); log_out[b] = log_in[b] + sum_sites(...).
 */
#ifndef NORMFLOW_B200_H
#define NORMFLOW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NFK_OK            0
#define NFK_EINVAL       -1   /* bad argument (null pointer, unsupported size)  */
#define NFK_EUNSUPPORTED -2   /* combination not implemented                     */
#define NFK_ECUDA        -3   /* CUDA runtime reported an error at launch        */

#define NFK_FROZEN_ZERO 0
#define NFK_FROZEN_COPY 1

/* spline extrapolation modes (lib/spline/spline.py:415-432) */
#define NFK_EXTRAP_NONE   0   /* first / last segment extended                   */
#define NFK_EXTRAP_LINEAR 1   /* straight line with the end-knot derivative      */
#define NFK_EXTRAP_ANTI   2   /* point mirror about the end knot                 */
#define NFK_EXTRAP_PERIODIC 3 /* even mirror image about the end knot (spline.py:502-508, 518-524): shared 1-D
                                 knots only, forward direction only (the map is not monotone); the slope
                                 changes sign there, so log_out is NaN for a sample with such a point    */

/* activations of the ConvAct conditioner (nn/scalar/modules.py:43-54) */
#define NFK_ACT_NONE       0
#define NFK_ACT_TANH       1
#define NFK_ACT_RELU       2
#define NFK_ACT_LEAKY_RELU 3
#define NFK_ACT_SOFTPLUS   4
#define NFK_ACT_ABS        5

#define NFK_MAX_DIM   4
#define NFK_MAX_KNOTS 64

/* Lattice geometry: ndim in 1..4, shape[d] = extent (unused entries = 1). */
typedef struct {
    int32_t ndim;
    int32_t shape[NFK_MAX_DIM];
} nfk_lattice;

/* Parameters of a rational-quadratic coupling spline
 * (RQSplineCoupling_.__init__, nn/scalar/couplings_.py:159-176). */
typedef struct {
    int32_t n_knots;        /* K; conditioner emits P = 3K-2 channels            */
    float   xlim0, xlim1;   /* xlim                                              */
    float   ylim0, ylim1;   /* ylim                                              */
    int32_t extrap_left;    /* NFK_EXTRAP_NONE | NFK_EXTRAP_LINEAR               */
    int32_t extrap_right;
} nfk_rqs_params;

const char* nfk_strerror(int code);
int nfk_version(void);
/* number of kernel launches issued by this library since load (bench "gpu_launches") */
uint64_t nfk_launch_count(void);

/* ---------------------------------------------------------------- masks ----
 * EvenOddMask.make_mask (mask/mask.py:53-61): mask[ind] = (1 - parity + sum(ind)
 * [- ind[exclude_mu]]) mod 2; exclude_mu < 0 means none.
 * AlongAxesEvenOddMask.make_mask (mask/mask.py:64-72): (1 - parity + ind[mu]) mod 2.
 * Integer work; bit-exact with the reference.                                   */
int nfk_mask_evenodd(uint8_t* mask, nfk_lattice lat, int parity, int exclude_mu, void* stream);
int nfk_mask_alongaxis(uint8_t* mask, nfk_lattice lat, int parity, int mu, void* stream);
/* Mask.split / purify (mask/mask.py:30-37): y[b][s] = mask[s]==keep ? x[b][s] : 0 */
int nfk_mask_select(const float* x, const uint8_t* mask, int keep, float* y,
                    int64_t B, int64_t V, void* stream);

/* ---------------------------------------------------------------- prior ----
 * NormalPrior.sample_ (prior/prior.py:26-28, 92-104): x = loc + scale * N(0,1)
 * (Philox4x32-10 counter RNG + Box-Muller, stream (seed, offset)), and
 * Prior.log_prob (prior/prior.py:30-36): logr[b] = sum_s( -(x-loc)^2/(2 scale^2)
 * - log scale - log sqrt(2 pi) ).  loc / scale: float[V] or NULL (0 / 1).       */
int nfk_prior_normal_sample(float* x, float* logr, int64_t B, int64_t V,
                            const float* loc, const float* scale,
                            uint64_t seed, uint64_t offset, void* stream);
/* The same draw with the generator state in device memory: state[0] = Philox key (seed),
 * state[1] = stream offset, advanced by one on the stream after the draw -- the form that can
 * be captured in a CUDA graph (every replay draws a fresh batch).                          */
int nfk_prior_normal_sample_dev(float* x, float* logr, int64_t B, int64_t V,
                                const float* loc, const float* scale,
                                uint64_t* state, void* stream);
int nfk_prior_normal_logprob(const float* x, float* logr, int64_t B, int64_t V,
                             const float* loc, const float* scale, void* stream);

/* -------------------------------------------------------------- couplings --
 * AffineCoupling_.atomic_forward / atomic_backward (couplings_.py:123-139):
 *   out[B][2][V] -> t, s = |s| at active sites;
 *   fwd: y = t + x e^{-s}, log_out = log_in - sum s ; inv: y = (x - t) e^{s}, + sum s.
 * _bwd is the vector-Jacobian product of _fwd: given gy[B][V] and glog[B]
 * (NULL = 0) it writes gx[B][V] and gout[B][2][V] (zeros at frozen sites).      */
int nfk_affine_fwd(const float* x, const float* out, const uint8_t* mask, int parity,
                   int frozen_mode, const float* log_in, float* y, float* log_out,
                   int64_t B, int64_t V, void* stream);
int nfk_affine_inv(const float* x, const float* out, const uint8_t* mask, int parity,
                   int frozen_mode, const float* log_in, float* y, float* log_out,
                   int64_t B, int64_t V, void* stream);
int nfk_affine_bwd(const float* x, const float* out, const uint8_t* mask, int parity,
                   int frozen_mode, const float* gy, const float* glog,
                   float* gx, float* gout, int64_t B, int64_t V, void* stream);

/* ShiftCoupling_.atomic_forward / atomic_backward (couplings_.py:110-116):
 *   out[B][1][V]; y = x + sign * t at active sites (sign = +1 fwd, -1 inv).     */
int nfk_shift_apply(const float* x, const float* out, const uint8_t* mask, int parity,
                    int frozen_mode, float sign, float* y, int64_t B, int64_t V, void* stream);

/* RQSplineCoupling_.atomic_forward / atomic_backward + make_spline
 * (couplings_.py:178-262) + RQSpline (lib/spline/spline.py:39-123, 154-287,
 * 458-486): out[B][3K-2][V] -> knots by softmax/cumsum/softplus(beta=ln2), bin
 * search (searchsorted right=False + clamp), Pade[2,2] value and derivative,
 * log|dy/dx| summed over active sites.  _inv uses the numerically stable root.
 * _bwd: VJP of _fwd (gx[B][V], gout[B][3K-2][V]).                               */
int nfk_rqs_fwd(const float* x, const float* out, const uint8_t* mask, int parity,
                int frozen_mode, nfk_rqs_params prm, const float* log_in,
                float* y, float* log_out, int64_t B, int64_t V, void* stream);
int nfk_rqs_inv(const float* x, const float* out, const uint8_t* mask, int parity,
                int frozen_mode, nfk_rqs_params prm, const float* log_in,
                float* y, float* log_out, int64_t B, int64_t V, void* stream);
int nfk_rqs_bwd(const float* x, const float* out, const uint8_t* mask, int parity,
                int frozen_mode, nfk_rqs_params prm, const float* gy, const float* glog,
                float* gx, float* gout, int64_t B, int64_t V, void* stream);

/* ------------------------------------------------- pointwise spline chain --
 * SplineNet_ / DistConvertor_ (nn/scalar/modules_.py:93-114, 277-302, 333-383):
 * one shared 1-D spline with explicit knots applied to every element.
 *   knots: float[5][K] (device) = kx | ky | kd | cx | cy, where cx = x_hi - kx and
 *          cy = y_hi - ky are the same knots measured from the upper end (read only
 *          when logistic != 0; the first 3K floats suffice otherwise).
 * logistic != 0 wraps the spline as Expit_ -> spline -> Logit_ (the DistConvertor_
 * chain, x_hi = y_hi = 1) evaluated in complement form -- every point of (0,1) is
 * carried as (s, 1-s) and every knot as (k, 1-k) -- so the tails keep fp32 relative
 * accuracy.  inverse != 0 evaluates ModuleList_.backward of the chain.
 * _bwd: VJP of the forward direction: gx[B][V] and gknots float[5][K] (same layout;
 * ACCUMULATED with atomics: zero it first).                                     */
int nfk_spline1d_fwd(const float* x, const float* knots, int K, int extrap_left, int extrap_right,
                     int logistic, int inverse, const float* log_in, float* y, float* log_out,
                     int64_t B, int64_t V, void* stream);
int nfk_spline1d_bwd(const float* x, const float* knots, int K, int extrap_left, int extrap_right,
                     int logistic, const float* gy, const float* glog,
                     float* gx, float* gknots, int64_t B, int64_t V, void* stream);
/* Expit_ / Logit_ alone (modules_.py:93-114); which = 0 expit, 1 logit.         */
int nfk_logistic_fwd(const float* x, int which, const float* log_in, float* y, float* log_out,
                     int64_t B, int64_t V, void* stream);
int nfk_logistic_bwd(const float* x, int which, const float* gy, const float* glog, float* gx,
                     int64_t B, int64_t V, void* stream);

/* --------------------------------------------------------------- action ----
 * ScalarPhi4Action.action (action/scalar_action.py:38-46):
 *   S[b] = sum_x (w2 phi^2 + w4 phi^4) - w0 sum_mu sum_x phi(x) phi(x - mu)  (periodic)
 * _bwd: gphi = gS[b] (2 w2 phi + 4 w4 phi^3 - w0 sum_mu (phi(x+mu) + phi(x-mu))). */
int nfk_phi4_action_fwd(const float* phi, nfk_lattice lat, float w0, float w2, float w4,
                        float* S, int64_t B, void* stream);
int nfk_phi4_action_bwd(const float* phi, nfk_lattice lat, float w0, float w2, float w4,
                        const float* gS, float* gphi, int64_t B, void* stream);

/* ------------------------------------------------------------ conditioner --
 * One layer of ConvAct (nn/scalar/modules.py:131-145): Conv{1,2,3}d / Conv4d
 * (convNd.py:84-127) with padding='same', padding_mode='circular', odd kernel
 * size ksize in every direction, fused activation.
 *   in[B][Ci][V], w[Co][Ci][ksize^ndim] (standard layout), bias[Co] or NULL,
 *   out[B][Co][V] = act(bias + sum w * in(shifted, periodic)).
 * in_mask != NULL: the input is read as (in_mask[s]==in_keep ? in : 0), i.e.
 * Mask.split fused into the first layer (couplings_.py:88-89).
 * dact_from != NULL (backward use): out is multiplied by act'(.) evaluated from
 * the saved post-activation tensor dact_from[B][Co][V] with activation dact_kind.
 * w_transposed != 0 (backward use, data gradient): `w` is the FORWARD layer's
 * weight w[Ci][Co][taps] and is read transposed with every tap flipped, so the
 * same kernel computes d loss / d input from d loss / d pre-activation.          */
int nfk_conv_circ_fwd(const float* in, const float* w, int w_transposed, const float* bias,
                      const uint8_t* in_mask, int in_keep,
                      int act, const float* dact_from, int dact_kind,
                      float* out, nfk_lattice lat, int ksize,
                      int Ci, int Co, int64_t B, void* stream);
/* The same layer for an input known to vanish off one checkerboard partition (in_parity: the parity of
 * row + column of its non-zero sites) -- the data gradient of a checkerboard coupling's conditioner, whose
 * incoming gradient lives on the updated partition only (couplings_.py:56-64).  Same result as
 * nfk_conv_circ_fwd; on 2-D 3x3 layers the products with the known zeros are not computed.          */
int nfk_conv_circ_fwd_cb(const float* in, int in_parity, const float* w, int w_transposed,
                         const float* bias, int act, const float* dact_from, int dact_kind,
                         float* out, nfk_lattice lat, int ksize, int Ci, int Co, int64_t B, void* stream);
/* gw[Co][Ci][ksize^ndim] += sum_{b,s} gpre[b][co][s] * in[b][ci][s + tap]  and
 * gbias[Co] += sum gpre (gbias may be NULL).  ACCUMULATES with atomics.         */
int nfk_conv_circ_bwd_weight(const float* in, const uint8_t* in_mask, int in_keep,
                             const float* gpre, float* gw, float* gbias,
                             nfk_lattice lat, int ksize, int Ci, int Co, int64_t B, void* stream);

/* The same weight gradient when gpre is the gradient of a CHECKERBOARD coupling's conditioner
 * output (couplings_.py:124,179 followed by purify, :126-127,186-187): it vanishes on every site
 * with (sum of coordinates) % 2 != g_parity, so only the other half of the sites is visited.
 * No input mask.  ACCUMULATES with atomics like nfk_conv_circ_bwd_weight.                */
int nfk_conv_circ_bwd_weight_cb(const float* in, const float* gpre, int g_parity, float* gw, float* gbias,
                                nfk_lattice lat, int ksize, int Ci, int Co, int64_t B, void* stream);

/* ----------------------------------------------------------------- mcmc ----
 * Metropolis.calc_accept_status + calc_accept_indices (mcmc/mcmc.py:304-328),
 * sequential independence-Metropolis scan on the device (one warp):
 *   accept[i] = log_u[i] < ref - l[i];  on accept ref <- l[i]
 * logq, logp float32[B] (l = logq - logp formed in float64); log_u float64[B] =
 * np.log(np.random.rand(B)) drawn and logged on the HOST exactly as the reference
 * does (mcmc.py:312), so decisions match it bit for bit given the same seed;
 * ref_inout float64[2] = {ref, has_ref} is the chain state carried across calls
 * (MCMCSampler._ref['logqp'], mcmc.py:21,77-81); accept uint8[B]; idx int64[B] =
 * index of the last accepted proposal <= i, or -1 while none has been accepted
 * yet (the caller substitutes the previous call's last sample, mcmc.py:67-68);
 * n_accept int64[1] (may be NULL) = number of acceptances.                       */
int nfk_metropolis_scan(const float* logq, const float* logp, const double* log_u,
                        double* ref_inout, uint8_t* accept, int64_t* idx, int64_t* n_accept,
                        int64_t B, void* stream);
/* MCMCSampler.estimate_accept_rate (mcmc/mcmc.py:117-124; Resampler('shuffling'),
 * lib/stats/resampler.py:62-64): acceptance rates of R chains built from R permutations
 * of the same N values, one warp per chain, all chains concurrently.
 * logqp float64[N] (device); perm int64[R][N] (NULL: identity); log_u float64[R][N] =
 * np.log(np.random.rand(N)) per chain, drawn on the host in the reference's order;
 * rates float64[R] = mean accept flag of chain r (its first proposal is accepted).  */
int nfk_metropolis_rates(const double* logqp, const int64_t* perm, const double* log_u,
                         double* rates, int64_t N, int64_t R, void* stream);
/* index_select(0, idx) (mcmc.py:73-75): dst[i][:] = idx[i] >= 0 ? src[idx[i]][:] :
 * prev[:]  (prev may be NULL when no idx is negative).                          */
int nfk_gather_rows(const float* src, const int64_t* idx, const float* prev, float* dst,
                    int64_t B, int64_t row_elems, void* stream);

/* Weight (and bias) gradient of a 2-D 3x3 circular convolution with 8 input channels on the tensor cores
 * (tcgen05): gw[Co][8][3][3] += sum over batch and sites of gpre x shifted in, gbias[Co] += sum gpre
 * (adjoint of ConvAct's layers, modules.py:131-145; same contract as nfk_conv_circ_bwd_weight[_cb], the
 * caller zeroes gw / gbias).  g_parity -1: gpre dense; 0 / 1: gpre vanishes off that checkerboard
 * partition and only its sites are visited.  Operands are split into tf32 pairs (float32 exponent range,
 * relative accuracy 2^-22), accumulation is fp32.  NFK_EUNSUPPORTED unless Ci == 8, Co <= 32, 16-byte
 * aligned inputs, L1 % 8 == 0 (checkerboard form: L1 % 16 == 0, even L0) and at most 64 (active) columns
 * per row; the caller then uses nfk_conv_circ_bwd_weight[_cb].                                       */
int nfk_conv2d_wgrad_tc(const float* in, const float* gpre, int g_parity, float* gw, float* gbias,
                        int L0, int L1, int Ci, int Co, int64_t B, void* stream);

/* -------------------------------------------------------- fused 2-D step ---
 * One whole atomic coupling step on a 2-D lattice with a ConvAct(1 -> H -> H -> P)
 * conditioner (3x3 circular convolutions, tanh, tanh, none; optional biases) in ONE
 * kernel: Coupling_.forward's step k (couplings_.py:56-64) = Mask.split + conditioner
 * (modules.py:131-145) + atomic_forward/backward (couplings_.py:123-139, 178-200).  The
 * (B,P,L0,L1) conditioner output never reaches memory and its last layer is evaluated
 * at the active sites only.  Evaluation only (sampling / log_prob); see
 * nfk_fused2d_step_train for the training forward.  Two kernels sit behind it: the
 * tcgen05 kernel (conditioner layers 2 and 3 as fp16-pair implicit GEMMs; even L0 and
 * L1 <= 160, n_knots in {4,5,6,8,10}) and a CUDA-core kernel (L1 % 4 == 0, n_knots <= 16).
 *   x, y: [B][L0][L1] (y != x), full-field semantics (frozen sites copied).
 *   w1[H][1][3][3], w2[H][H][3][3], w3[P][H][3][3]; b1, b2, b3 may be NULL.  H == 8.
 *   kind 0: affine (P = 2); kind 1: RQ spline (P = 3K-2, prm as in nfk_rqs_fwd).
 *   mask_parity: the `parity` of EvenOddMask.make_mask; parity: the step's partition.
 *   inverse != 0: the atomic_backward direction.                                   */
int nfk_fused2d_step(const float* x, const float* w1, const float* b1, const float* w2,
                     const float* b2, const float* w3, const float* b3, int H, int kind,
                     nfk_rqs_params prm, int mask_parity, int parity, int inverse,
                     const float* log_in, float* y, float* log_out,
                     int L0, int L1, int64_t B, void* stream);

/* The same step as the forward pass of TRAINING (Fitter.step, _normflowcore.py:275-294): besides y
 * and log_out it stores what the gradient kernels need -- the post-activation hidden layers
 * h1, h2 [B][H][L0][L1] and the conditioner output out[B][P][L0][L1] (defined at the active
 * sites only) -- so that autograd can run nfk_rqs_bwd / nfk_affine_bwd and the convolution
 * gradient kernels without re-evaluating the conditioner.  Forward direction only.  Runs on the
 * tensor-core kernel; returns NFK_EUNSUPPORTED outside its geometry (even L0, L1 <= 160;
 * n_knots in {4, 5, 6, 8, 10}), in which case the caller evaluates the layers one by one.   */
int nfk_fused2d_step_train(const float* x, const float* w1, const float* b1, const float* w2,
                           const float* b2, const float* w3, const float* b3, int H, int kind,
                           nfk_rqs_params prm, int mask_parity, int parity,
                           const float* log_in, float* y, float* log_out,
                           float* h1, float* h2, float* out,
                           int L0, int L1, int64_t B, void* stream);

/* ------------------------------------------------ fused N-D step (2-D .. 4-D) ---
 * The same atomic coupling step (couplings_.py:56-64) on a lattice of 2 to 4 dimensions -- ConvAct(1 -> H -> H -> P)
 * conditioner with 3^D-tap circular convolutions (modules.py:131-145; ConvNd / Conv4d, convNd.py:84-127), tanh,
 * tanh, none -- with layers 2 and 3 of the conditioner on the tensor cores (tcgen05 fp16-pair implicit GEMM) and
 * the affine / RQ-spline transform fused into the last layer's epilogue; the (B, P, *L) conditioner output is
 * never formed.  Evaluation only (sampling / log_prob).  The hidden layers pass between the three kernels as
 * fp16-pair records in `workspace` (4 H bytes per site each).
 *   x, y: [B][*lat.shape] (y != x), full-field semantics;  w1[H][1][3^D], w2[H][H][3^D], w3[P][H][3^D] in the
 *   standard (Co, Ci, *k) layout (Conv4d: its (Co, Ci, k0, k, k, k) view); b1, b2, b3 may be NULL;
 *   hidden width H in {8, 16, 32, 64} (channel groups of 8: one MMA per tap and group; H = 64 runs layer 2 in two
 *   passes of 32 output channels).
 *   kind / prm / mask_parity / parity / inverse / log_in / log_out as for nfk_fused2d_step.
 *   workspace: device buffer of at least nfk_fusednd_workspace(...) bytes, 256-byte aligned.
 * NFK_EUNSUPPORTED (from either entry) unless every lattice extent is even and >= 2, n_knots in {4,5,6,8,10}
 * and a tile of the lattice fits shared memory; the caller then evaluates the layers one by one.           */
int64_t nfk_fusednd_workspace(nfk_lattice lat, int H, int kind, int n_knots, int64_t B);
int nfk_fusednd_step(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                     const float* w3, const float* b3, int H, int kind, nfk_rqs_params prm,
                     nfk_lattice lat, int mask_parity, int parity, int inverse,
                     const float* log_in, float* y, float* log_out, int64_t B,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* The N-D step as the forward pass of TRAINING (Fitter.step, _normflowcore.py:275-294): additionally stores the
 * post-activation hidden layers h1, h2 [B][H][V] and the conditioner output out [B][P][V] (defined at the active
 * sites only) in float32 channel-major layout for the gradient kernels.  Forward direction only.                */
int nfk_fusednd_step_train(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                           const float* w3, const float* b3, int H, int kind, nfk_rqs_params prm,
                           nfk_lattice lat, int mask_parity, int parity,
                           const float* log_in, float* y, float* log_out,
                           float* h1, float* h2, float* out, int64_t B,
                           void* workspace, int64_t workspace_bytes, void* stream);

/* Data gradient of one circular 3^D convolution layer of a ConvAct stack on the tensor cores -- what autograd runs
 * for modules.py:131-145 in Fitter.step (_normflowcore.py:288):
 *     gin[b][ci][s] = act'(h[b][ci][s]) * sum_{co, t} w[co][ci][t] * gpre[b][co][s - t]
 * with act' = 1 - h^2 (h = post-activation output of the tanh layer below, [B][Ci][V]) or 1 when h is NULL.
 * gpre [B][Co][V] = d loss / d (this layer's pre-activation output), w [Co][Ci][3^D], gin [B][Ci][V].
 * Ci in {8, 16, 32, 64}, 1 <= Co <= 64, 2-D .. 4-D lattices with even extents; NFK_EUNSUPPORTED otherwise (the
 * caller then uses nfk_conv_circ_fwd with transposed weights).  Operands are fp16 pairs scaled by a power of two
 * found from max |gpre| on the device.  g_parity >= 0 (as in nfk_conv_circ_fwd_cb): gpre vanishes on the sites whose
 * coordinate sum % 2 != g_parity -- the gradient of a checkerboard coupling's conditioner output -- and only the MMAs
 * that land on populated sites are issued (half of them); -1: dense.
 * `workspace`: nfk_convnd_dgrad_workspace bytes (covers either form), 256-byte aligned.                          */
int64_t nfk_convnd_dgrad_workspace(nfk_lattice lat, int Co, int Ci, int64_t B);
int nfk_convnd_dgrad(const float* gpre, int g_parity, const float* w, const float* h, float* gin, int Co, int Ci,
                     nfk_lattice lat, int64_t B, void* workspace, int64_t workspace_bytes, void* stream);

/* Weight and bias gradient of one circular 3^D convolution layer with 8 input channels on the tensor cores (autograd
 * of convNd.py:84-127 / modules.py:131-145 in Fitter.step):
 *     gw[co][ci][t] += sum_{b, s} gpre[b][co][s] * h[b][ci][s + t - 1],     gb[co] += sum_{b, s} gpre[b][co][s]
 * h [B][8][V] the layer's input, gpre [B][Co][V], Co <= 32; gw [Co][8][3^D] and gb [Co] (or NULL) are ACCUMULATED into.
 * The sites are the GEMM's K dimension, read from site-major fp16-pair records as MN-major operands (no im2col).
 * 2-D .. 4-D, even extents, innermost extent a multiple of 16; NFK_EUNSUPPORTED otherwise (use
 * nfk_conv_circ_bwd_weight).  `workspace`: nfk_convnd_wgrad_workspace bytes, 256-byte aligned.                    */
int64_t nfk_convnd_wgrad_workspace(nfk_lattice lat, int Co, int Ci, int64_t B);
int nfk_convnd_wgrad(const float* h, const float* gpre, float* gw, float* gb, int Co, int Ci,
                     nfk_lattice lat, int64_t B, void* workspace, int64_t workspace_bytes, void* stream);

/* Both gradients of ONE layer in one call (the per-layer step of autograd through modules.py:131-145): what
 * nfk_convnd_wgrad and nfk_convnd_dgrad compute, with d loss / d pre-activation reduced (max |g|) and packed into records
 * once for both kernels.  h_in [B][8][V] is the layer's input -- and, when act_below != 0, the tanh output whose derivative
 * multiplies the data gradient; gw / gb are accumulated into, gin [B][8][V] is written.  Applies where both kernels do
 * (8 input channels, Co <= 32, innermost extent a multiple of 16); NFK_EUNSUPPORTED otherwise.  g_parity as in
 * nfk_convnd_dgrad (the weight gradient reads the dense records either way).                                        */
int64_t nfk_convnd_layer_bwd_workspace(nfk_lattice lat, int Co, int Ci, int64_t B);
int nfk_convnd_layer_bwd(const float* h_in, const float* gpre, int g_parity, const float* w, int act_below,
                         float* gin, float* gw, float* gb, int Co, int Ci, nfk_lattice lat, int64_t B,
                         void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------- PSD block (spectral part) ---
 * PSDBlock_ / FFTNet_ (psd_.py:25-40, fftflow_.py:121-131,167-180): the real-to-complex
 * and complex-to-real transforms are cuFFT calls made by the host package; these entries
 * are everything between them.  Spectra are complex64 half-spectra [B][Kc] passed as
 * float pairs, Kc = prod(L[:-1]) * last_half, last_half = L[-1]/2 + 1.
 *
 * nfk_psd_weights_fwd: w[k] = ipsd[k]^(-1/2) (inverse != 0: ^(+1/2), the FFTNet_.backward
 *   direction) and logj[0] = sum_k m_k log w[k] with m_k = 2 - [col == 0] - [col ==
 *   last_half-1], col = k % last_half (FFTNet_.log_jacobian, fftflow_.py:167-178).
 * nfk_psd_weights_bwd: g_ipsd from gw[Kc] (may be NULL) and glogj[1] (may be NULL).     */
int nfk_psd_weights_fwd(const float* ipsd, int64_t Kc, int last_half, int inverse,
                        float* w, float* logj, void* stream);
int nfk_psd_weights_bwd(const float* ipsd, const float* w, const float* gw, const float* glogj,
                        int64_t Kc, int last_half, int inverse, float* g_ipsd, void* stream);
/* nfk_psd_scale: Y[b][k] = X[b][k] * w[k] (Y may alias X).  With zero_mode != NULL the k = 0
 *   element becomes (zero_scale * zero_mode[b], 0) instead: PSDBlock_.forward subtracts the
 *   sample mean before the transform and adds the mean-field net's output afterwards
 *   (psd_.py:28-32), i.e. it replaces the zero mode by V * y_mf[b].
 * nfk_psd_scale_bwd: gX = gY * w (gX may be NULL or alias gY; 0 at a replaced zero mode),
 *   g_zero[b] = zero_scale * Re gY[b][0] (required iff replace_zero), and, when gw != NULL,
 *   gw[k] = sum_b Re(conj(X[b][k]) gY[b][k]) through the workspace gw_part
 *   [nfk_psd_chunks(B, Kc)][Kc] (deterministic two-stage sum, no atomics).               */
int nfk_psd_scale(const float* X, const float* w, const float* zero_mode, float zero_scale,
                  float* Y, int64_t B, int64_t Kc, void* stream);
int nfk_psd_scale_bwd(const float* X, const float* gY, const float* w, int replace_zero,
                      float zero_scale, float* gX, float* gw, float* gw_part, float* g_zero,
                      int64_t B, int64_t Kc, void* stream);
int nfk_psd_chunks(int64_t B, int64_t Kc);
/* MeanFieldNet_ applied to a whole field (meanfield_.py:26-32, 42-48):
 * nfk_sample_mean: mean[b] = scale * sum_v x[b][v]   (scale = 1/V for the mean; the same
 *   kernel with scale = 1 is the adjoint of nfk_sample_shift with respect to delta);
 * nfk_sample_shift: y[b][v] = x[b][v] + delta[b]     (y may alias x).                    */
int nfk_sample_mean(const float* x, int64_t B, int64_t V, float scale, float* mean, void* stream);
int nfk_sample_shift(const float* x, const float* delta, float* y, int64_t B, int64_t V, void* stream);

/* ---------------------------------------------------------------- knot table ---
 * SplineNet.make_spline for ONE shared spline (modules.py:369-391): K-1 raw widths wx, K-1 raw
 * heights wy and K raw derivatives wd (NULL: smooth derivatives, spline.py:125-152) -> table
 * float32[5][K] = knots_x | knots_y | knots_d | (xlo+xw) - knots_x | (ylo+yw) - knots_y, the
 * last two accumulated from the right end (what nfk_spline1d_* takes as its both-ends table;
 * the first three rows are the reference's knots).  2 <= K <= 256.  _bwd: gradients of the raw
 * parameters from g_table[5][K] (gwd required iff wd != NULL).                              */
int nfk_knots_fwd(const float* wx, const float* wy, const float* wd, int K, float xlo, float xw,
                  float ylo, float yw, float* table, void* stream);
int nfk_knots_bwd(const float* wx, const float* wy, const float* wd, int K, float xlo, float xw,
                  float ylo, float yw, const float* g_table, float* gwx, float* gwy, float* gwd,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NORMFLOW_B200_H */
