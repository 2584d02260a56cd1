// nfk_fused.cuh -- one whole atomic coupling step on a 2-D lattice, conditioner included.
//
// Reference dataflow (couplings_.py:123-130, 178-188 with modules.py:131-145): three
// circular 3x3 convolutions (1 -> H -> H -> P, tanh, tanh, none) write a dense
// (B, P, L0, L1) tensor to memory -- 7.5 GB per step at B = 16384, 64x64, P = 28 -- that
// the spline then reads back once.  Here a CTA keeps a strip of the lattice on chip:
//
//   phase 0  x rows r0-3 .. r0+R+2 -> shared memory (periodic wrap by index; masked copy
//            xf = frozen partition only, Mask.split fused away)
//   phase 1  h1 = tanh(conv(xf))         rows r0-2 .. r0+R+1     (H channels)
//   phase 2  h2 = tanh(conv(h1))         rows r0-1 .. r0+R       (H channels)
//   phase 3  out = conv(h2) ONLY at the active sites of rows r0 .. r0+R-1, kept in
//            registers (P accumulators per site), handed straight to the spline / affine
//            transform; y (active: transformed, frozen: copied) is the only HBM write.
//
// Work is cut into items = (row, group of 4 columns); an item keeps a 4-column x H (or
// 2-site x P) accumulator tile in registers, reads its 3 x 6 input window with 128-bit
// shared loads and the weights as warp-wide broadcasts.  Each phase is a plain function of
// its item index so the host harness (tests/cpu_harness) can run the same code.
#pragma once

#include "nfk_ops.cuh"

namespace nfk {

constexpr int kFH = 8;     // hidden width the kernel is specialised for

// fast, accurate tanh: 1 - 2/(e^{2x} + 1); abs error ~1e-7 (MUFU.EX2 + MUFU.RCP on the device)
NFK_HD float tanh_fast(float v) {
#if defined(__CUDA_ARCH__)
    const float e = __expf(2.f * v);
    return 1.f - __fdividef(2.f, e + 1.f);
#else
    const float e = expf(2.f * v);
    return 1.f - 2.f / (e + 1.f);
#endif
}

struct FusedGeom {
    int L0, L1;          // lattice rows, columns (L1 % 4 == 0)
    int R;               // output rows per strip
    int WS;              // shared row stride = L1 + 4 (index 0 = column -1, 1..L1, L1+1 = column L1)
    int ncg;             // column groups = L1 / 4
    int mask_parity;     // EvenOddMask(parity=...)
    int active_val;      // mask value of the partition being updated
};

// mask bit of site (r, c): (1 - parity + r + c) mod 2   (mask.py:55-58)
NFK_HD int fused_mask_bit(const FusedGeom& g, int r, int c) { return (1 - g.mask_parity + r + c) & 1; }

struct FusedSmem {
    float *w1s, *w2s, *w3s;   // [tap][H], [(ci*9+tap)][H], [(ci*9+tap)][PP]
    float *b1s, *b2s, *b3s;   // biases (zeros when absent)
    float *xs, *xf;           // [R+6][WS] field strip, and its frozen-partition copy
    float *h1s, *h2s;         // [H][R+4][WS], [H][R+2][WS]
};

NFK_HD int wrap(int v, int L) {
    v %= L;
    return v < 0 ? v + L : v;
}

// ---- phase 0: one element of the [R+6][WS] strip
NFK_HD void fused_load_x(const FusedGeom& g, const FusedSmem& m, const float* xb, int r0, int rows, int e) {
    const int j = e / g.WS, k = e % g.WS;
    if (j >= rows + 6) return;
    float v = 0.f, vf = 0.f;
    if (k <= g.L1 + 1) {
        const int r = wrap(r0 - 3 + j, g.L0), c = wrap(k - 1, g.L1);
        v = NFK_LDG(xb + r * g.L1 + c);
        vf = fused_mask_bit(g, r, c) == g.active_val ? 0.f : v;
    }
    m.xs[e] = v;
    m.xf[e] = vf;
}

// 3 x 6 window of a [rows][WS] plane starting at row i, padded index k0 (k0 % 4 == 0)
NFK_HD void load_window(const float* plane, int WS, int i, int k0, float (&w)[3][6]) {
#pragma unroll
    for (int dr = 0; dr < 3; ++dr) {
        const float* p = plane + (i + dr) * WS + k0;
#if defined(__CUDA_ARCH__)
        const float4 a = *reinterpret_cast<const float4*>(p);
        const float2 b = *reinterpret_cast<const float2*>(p + 4);
        w[dr][0] = a.x; w[dr][1] = a.y; w[dr][2] = a.z; w[dr][3] = a.w; w[dr][4] = b.x; w[dr][5] = b.y;
#else
        for (int k = 0; k < 6; ++k) w[dr][k] = p[k];
#endif
    }
}

// writes 4 values of row i (columns c0..c0+3) of a plane plus the wrap halos
NFK_HD void store_row4(float* plane, const FusedGeom& g, int i, int c0, const float (&v)[4]) {
    float* p = plane + i * g.WS + 1 + c0;
#pragma unroll
    for (int k = 0; k < 4; ++k) p[k] = v[k];
    if (c0 == 0) plane[i * g.WS + g.L1 + 1] = v[0];            // column L1  == column 0
    if (c0 == g.L1 - 4) plane[i * g.WS] = v[3];                // column -1  == column L1-1
}

// ---- phases 1 and 2: 4 columns x H channels of one row of a hidden layer
//   CI = 1: input plane xf, weights w1s;  CI = H: input planes h1s, weights w2s
template <int CI>
NFK_HD void fused_hidden_item(const FusedGeom& g, const float* in, int in_rows, const float* ws,
                              const float* bs, float* out, int out_rows, int out_valid, int item) {
    // in_rows / out_rows: allocated rows of one channel plane; out_valid: rows to compute
    const int i = item / g.ncg, c0 = (item % g.ncg) * 4;
    if (i >= out_valid) return;
    float acc[4][kFH];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int co = 0; co < kFH; ++co) acc[k][co] = bs[co];
#pragma unroll 1
    for (int ci = 0; ci < CI; ++ci) {
        float w[3][6];
        load_window(in + ci * in_rows * g.WS, g.WS, i, c0, w);
#pragma unroll
        for (int dr = 0; dr < 3; ++dr)
#pragma unroll
            for (int dc = 0; dc < 3; ++dc) {
                const float* wr = ws + (ci * 9 + dr * 3 + dc) * kFH;
#pragma unroll
                for (int co = 0; co < kFH; ++co) {
                    const float wv = wr[co];
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[k][co] = fmaf(w[dr][k + dc], wv, acc[k][co]);
                }
            }
    }
#pragma unroll
    for (int co = 0; co < kFH; ++co) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = tanh_fast(acc[k][co]);
        store_row4(out + co * out_rows * g.WS, g, i, c0, v);
    }
}

// channels held in registers (see the Ld interface in nfk_math.cuh)
template <int P>
struct RegLoad {
    const float* a;
    NFK_HD float operator()(int c) const { return a[c]; }
    NFK_HD float pick(int c0, int n, int j) const {
        float r = a[c0];
#pragma unroll
        for (int k = 1; k < 32; ++k)
            if (k < n) r = (k == j) ? a[c0 + k] : r;
        return r;
    }
};

struct FusedXform {
    RqsCfg cfg;
    int inverse;
};

// ---- phase 3: last conv layer at the two active sites of a 4-column group + transform.
// KIND 0: affine (P = 2), KIND 1: RQ spline with K knots (P = 3K-2).  PP = P rounded up to 4.
// Returns the group's contribution to log|det J|; writes y[r][c0..c0+3].
template <int KIND, int K>
NFK_HD float fused_out_item(const FusedGeom& g, const FusedSmem& m, const FusedXform& xf, int r0, int rows,
                            float* yb, int item) {
    constexpr int P = KIND == 0 ? 2 : 3 * K - 2;
    constexpr int PP = (P + 3) / 4 * 4;
    const int i = item / g.ncg, c0 = (item % g.ncg) * 4;
    if (i >= rows) return 0.f;
    const int r = r0 + i;                                     // lattice row (r0 + i < L0)
    // first active column of the group: mask(r, c0 + a0) == active_val
    const int a0 = (fused_mask_bit(g, r, c0) == g.active_val) ? 0 : 1;
    float acc[2][PP];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int p = 0; p < PP; ++p) acc[s][p] = m.b3s[p];
#pragma unroll 1
    for (int ci = 0; ci < kFH; ++ci) {
        float w[3][6];
        load_window(m.h2s + ci * (g.R + 2) * g.WS, g.WS, i, c0, w);
        float v[3][5];                                        // columns c0+a0-1 .. c0+a0+3
#pragma unroll
        for (int dr = 0; dr < 3; ++dr)
#pragma unroll
            for (int k = 0; k < 5; ++k) v[dr][k] = a0 ? w[dr][k + 1] : w[dr][k];
#pragma unroll
        for (int dr = 0; dr < 3; ++dr)
#pragma unroll
            for (int dc = 0; dc < 3; ++dc) {
                const float* wr = m.w3s + (ci * 9 + dr * 3 + dc) * PP;
#pragma unroll
                for (int p = 0; p < PP; ++p) {
                    const float wv = wr[p];
                    acc[0][p] = fmaf(v[dr][dc], wv, acc[0][p]);
                    acc[1][p] = fmaf(v[dr][dc + 2], wv, acc[1][p]);
                }
            }
    }
    // the field itself: row i+3 of the strip, columns c0..c0+3
    float yv[4];
    const float* xrow = m.xs + (i + 3) * g.WS + 1 + c0;
#pragma unroll
    for (int k = 0; k < 4; ++k) yv[k] = xrow[k];
    float lg = 0.f;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const float xv = a0 ? (s ? yv[3] : yv[1]) : (s ? yv[2] : yv[0]);
        float out, l;
        if (KIND == 0) {
            const float t = acc[s][0], sc = fabsf(acc[s][1]);
            if (!xf.inverse) { out = t + xv * expf(-sc); l = -sc; }
            else { out = (xv - t) * expf(sc); l = sc; }
        } else {
            const RegLoad<PP> ld{acc[s]};
            if (!xf.inverse) rqs_site_forward<K>(ld, xf.cfg, xv, out, l);
            else rqs_site_inverse<K>(ld, xf.cfg, xv, out, l);
        }
        lg += l;
        if (a0) { if (s) yv[3] = out; else yv[1] = out; }
        else { if (s) yv[2] = out; else yv[0] = out; }
    }
    float* yrow = yb + r * g.L1 + c0;
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<float4*>(yrow) = make_float4(yv[0], yv[1], yv[2], yv[3]);
#else
    for (int k = 0; k < 4; ++k) yrow[k] = yv[k];
#endif
    return lg;
}

// shared-memory floats needed for a geometry (weights + strips)
template <int PP>
NFK_HD int fused_smem_floats(int L1, int R) {
    const int WS = L1 + 4;
    return 9 * kFH + 9 * kFH * kFH + 9 * kFH * PP + 2 * kFH + PP      // weights and biases
           + 2 * (R + 6) * WS + kFH * (R + 4) * WS + kFH * (R + 2) * WS;
}

template <int PP>
NFK_HD FusedSmem fused_carve(float* base, int L1, int R) {
    const int WS = L1 + 4;
    FusedSmem m;
    float* p = base;                              // strips first: 16-byte aligned rows
    m.xs = p; p += (R + 6) * WS;
    m.xf = p; p += (R + 6) * WS;
    m.h1s = p; p += kFH * (R + 4) * WS;
    m.h2s = p; p += kFH * (R + 2) * WS;
    m.w1s = p; p += 9 * kFH;
    m.w2s = p; p += 9 * kFH * kFH;
    m.w3s = p; p += 9 * kFH * PP;
    m.b1s = p; p += kFH;
    m.b2s = p; p += kFH;
    m.b3s = p; p += PP;
    return m;
}

// weights into the shared layouts ([tap][co] / [(ci*9+tap)][co]); e walks all of them
template <int P, int PP>
NFK_HD void fused_load_weight(const FusedSmem& m, const float* w1, const float* w2, const float* w3,
                              const float* b1, const float* b2, const float* b3, int e) {
    const int n1 = 9 * kFH, n2 = 9 * kFH * kFH, n3 = 9 * kFH * PP;
    if (e < n1) {
        const int tap = e / kFH, co = e % kFH;
        m.w1s[e] = NFK_LDG(w1 + co * 9 + tap);
    } else if (e < n1 + n2) {
        const int q = e - n1, co = q % kFH, rt = q / kFH, ci = rt / 9, tap = rt % 9;
        m.w2s[q] = NFK_LDG(w2 + (co * kFH + ci) * 9 + tap);
    } else if (e < n1 + n2 + n3) {
        const int q = e - n1 - n2, p = q % PP, rt = q / PP, ci = rt / 9, tap = rt % 9;
        m.w3s[q] = p < P ? NFK_LDG(w3 + (p * kFH + ci) * 9 + tap) : 0.f;
    } else {
        const int q = e - n1 - n2 - n3;
        if (q < kFH) m.b1s[q] = b1 ? NFK_LDG(b1 + q) : 0.f;
        else if (q < 2 * kFH) m.b2s[q - kFH] = b2 ? NFK_LDG(b2 + q - kFH) : 0.f;
        else if (q < 2 * kFH + PP) m.b3s[q - 2 * kFH] = (b3 && q - 2 * kFH < P) ? NFK_LDG(b3 + q - 2 * kFH) : 0.f;
    }
}
template <int PP>
NFK_HD int fused_weight_elems() { return 9 * kFH + 9 * kFH * kFH + 9 * kFH * PP + 2 * kFH + PP; }

}  // namespace nfk
