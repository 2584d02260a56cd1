// nfk_psd.cuh -- per-mode arithmetic of the spectral weights (FFTNet_, reference fftflow_.py:121-131,167-180),
// host/device so that the CPU test harness runs the kernel's own code.
#pragma once
#include <math.h>
#include <stdint.h>
#include "nfk_math.cuh"

namespace nfk {

// multiplicity of half-spectrum mode k in the full spectrum: every mode counts twice (k and -k) except the
// self-conjugate planes at the two ends of the last axis (fftflow_.py:172-178)
NFK_HD float psd_mult(int64_t k, int Lh) {
    const int col = (int)(k % Lh);
    return 2.f - (col == 0 ? 1.f : 0.f) - (col == Lh - 1 ? 1.f : 0.f);
}

// w = ipsd^(-1/2) for the forward map, ipsd^(+1/2) for the inverse one
NFK_HD float psd_weight(float ipsd, int inverse) { return inverse ? sqrtf(ipsd) : 1.f / sqrtf(ipsd); }

// mode k's term of sum_k m_k log ipsd_k (the caller multiplies the sum by -+1/2)
NFK_HD double psd_logj_term(int64_t k, int Lh, float ipsd) { return (double)psd_mult(k, Lh) * (double)logf(ipsd); }

// d/d ipsd_k of  sum_k gw_k w_k + glogj * logj:   sign/2 * (gw w + glogj m) / ipsd   (sign -1 forward, +1 inverse)
NFK_HD float psd_weight_grad(int64_t k, int Lh, float ipsd, float w, float gw, float glogj, int inverse) {
    const float v = 0.5f * (gw * w + glogj * psd_mult(k, Lh)) / ipsd;
    return inverse ? v : -v;
}

}  // namespace nfk
