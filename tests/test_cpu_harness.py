"""fp32 per-site arithmetic of the CUDA kernels, run on the host through
tests/cpu_harness (the same nfk_ops.cuh functors, compiled with g++), against the
golden fixtures made from the reference.  Catches arithmetic / indexing mistakes
before any GPU time is spent; the `-m gpu` tests repeat these through the real
C-ABI on the device.  Tolerance: |d| <= 1e-5 * max(|ref|, 1) (north_star)."""

import ctypes

import numpy as np
import pytest

from conftest import load_golden
from cpu_harness import load, lattice, RqsParams

H = load()
F = ctypes.POINTER(ctypes.c_float)
U8 = ctypes.POINTER(ctypes.c_uint8)
I64 = ctypes.c_int64


def fp(a):
    return None if a is None else a.ctypes.data_as(F)


def up(a):
    return a.ctypes.data_as(U8)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def close(got, ref, tol=1e-5):
    ref = np.asarray(ref, dtype=np.float64)
    err = np.abs(np.asarray(got, dtype=np.float64) - ref)
    bound = tol * np.maximum(np.abs(ref), 1.0)
    assert np.all(err <= bound), f"max excess {np.max(err / bound):.3g} (abs err {err.max():.3g})"


def close_grad(got, ref, tol=1e-5):
    """Gradients: error measured against the scale of the whole gradient tensor (the
    softmax chain sums terms of mixed sign, so single small components carry the rounding
    of the large ones)."""
    ref = np.asarray(ref, dtype=np.float64)
    err = np.abs(np.asarray(got, dtype=np.float64) - ref).max()
    assert err <= tol * max(np.abs(ref).max(), 1.0), f"abs err {err:.3g} vs scale {np.abs(ref).max():.3g}"


def test_masks_bit_exact():
    g = load_golden("masks")
    for key in g.files:
        if not key.endswith("_meta"):
            continue
        tag, meta = key[:-5], g[key]
        parity, mu, shape = int(meta[0]), int(meta[1]), tuple(int(v) for v in meta[2:])
        m = np.empty(shape, dtype=np.uint8)
        if tag.startswith("eo_"):
            H.cpu_mask_evenodd(up(m), lattice(shape), parity, mu)
        else:
            H.cpu_mask_alongaxis(up(m), lattice(shape), parity, mu)
        assert np.array_equal(m, g[tag + "_mask"])


def test_action():
    g = load_golden("action")
    for i in range(9):
        cfgs = f32(g[f"a{i}_cfgs"])
        B, shape = cfgs.shape[0], cfgs.shape[1:]
        S = np.empty(B, dtype=np.float32)
        w0, w2, w4 = 0.67, 0.5 * (-4 * 0.67 + 2 * 0.67 * len(shape)), 0.5
        H.cpu_phi4_action_fwd(fp(cfgs), lattice(shape), ctypes.c_float(w0), ctypes.c_float(w2),
                              ctypes.c_float(w4), fp(S), I64(B))
        close(S, g[f"a{i}_S"])
        gS, gphi = f32(g[f"a{i}_gS"]), np.empty_like(cfgs)
        H.cpu_phi4_action_bwd(fp(cfgs), lattice(shape), ctypes.c_float(w0), ctypes.c_float(w2),
                              ctypes.c_float(w4), fp(gS), fp(gphi), I64(B))
        close(gphi, g[f"a{i}_gcfgs"])


def test_prior_logprob_and_sampler_statistics():
    g = load_golden("prior")
    x = f32(g["gen_x"])
    B, V = x.shape[0], x[0].size
    logr = np.empty(B, dtype=np.float32)
    H.cpu_prior_normal_logprob(fp(x), fp(logr), I64(B), I64(V), fp(f32(g["gen_loc"])), fp(f32(g["gen_scale"])))
    close(logr, g["gen_logr"])
    H.cpu_prior_normal_logprob(fp(f32(g["std_x"])), fp(logr), I64(B), I64(V), None, None)
    close(logr, g["std_logr"])
    # sampler: moments, reproducibility, stream separation, fused log-prob
    B, V = 64, 1030
    xs, lr = np.empty((B, V), dtype=np.float32), np.empty(B, dtype=np.float32)
    H.cpu_prior_normal_sample(fp(xs), fp(lr), I64(B), I64(V), None, None, ctypes.c_uint64(1234), ctypes.c_uint64(0))
    n = xs.size
    assert abs(xs.mean()) < 4 / np.sqrt(n) and abs(xs.var() - 1) < 4 * np.sqrt(2 / n)
    assert abs(np.mean(xs ** 3)) < 0.05 and abs(np.mean(xs ** 4) - 3) < 0.1
    ref = (-0.5 * xs.astype(np.float64) ** 2 - 0.5 * np.log(2 * np.pi)).sum(1)
    close(lr, ref)
    xs2 = np.empty_like(xs)
    H.cpu_prior_normal_sample(fp(xs2), None, I64(B), I64(V), None, None, ctypes.c_uint64(1234), ctypes.c_uint64(0))
    assert np.array_equal(xs, xs2)
    H.cpu_prior_normal_sample(fp(xs2), None, I64(B), I64(V), None, None, ctypes.c_uint64(1234), ctypes.c_uint64(1))
    assert abs(np.corrcoef(xs.ravel(), xs2.ravel())[0, 1]) < 0.02
    # no correlation between neighbouring samples / sites
    assert abs(np.corrcoef(xs[:-1].ravel(), xs[1:].ravel())[0, 1]) < 0.02
    assert abs(np.corrcoef(xs[:, :-1].ravel(), xs[:, 1:].ravel())[0, 1]) < 0.02


def test_affine_kernel():
    g = load_golden("affine_kernel")
    mask = np.ascontiguousarray(g["mask"])
    B, V = g["x"].shape[0], mask.size
    out = f32(g["out"])
    for parity in (0, 1):
        xa = f32(g["x"] * (mask if parity == 0 else 1 - mask))
        y, logJ = np.empty_like(xa), np.empty(B, dtype=np.float32)
        H.cpu_affine_fwd(fp(xa), fp(out), up(mask), parity, 0, None, fp(y), fp(logJ), I64(B), I64(V))
        close(y, g[f"p{parity}_fx"])
        close(logJ, g[f"p{parity}_logJ"])
        xi, li = np.empty_like(xa), np.empty(B, dtype=np.float32)
        H.cpu_affine_inv(fp(y), fp(out), up(mask), parity, 0, None, fp(xi), fp(li), I64(B), I64(V))
        close(xi, g[f"p{parity}_xinv"])
        close(li, g[f"p{parity}_loginv"])
        gx, gout = np.empty_like(xa), np.empty_like(out)
        # the reference's  t + x_active * exp(-s)  has t = s = 0 at frozen sites, i.e. it hands
        # x_active through there: its VJP is the FROZEN_COPY mode of the kernel
        H.cpu_affine_bwd(fp(xa), fp(out), up(mask), parity, 1, fp(f32(g["r"])), fp(f32(g["c"])),
                         fp(gx), fp(gout), I64(B), I64(V))
        close(gx, g[f"p{parity}_gx"])
        close(gout, g[f"p{parity}_gout"])
        # full-field mode: frozen sites pass through
        xfull = f32(g["x"])
        H.cpu_affine_fwd(fp(xfull), fp(out), up(mask), parity, 1, None, fp(y), fp(logJ), I64(B), I64(V))
        act = (mask if parity == 0 else 1 - mask).astype(bool)
        assert np.array_equal(y[:, ~act], xfull[:, ~act])
        close(y[:, act], g[f"p{parity}_fx"][:, act])


@pytest.mark.parametrize("tag,left,right", [("lin", 1, 1), ("none", 0, 0), ("mixed", 1, 0)])
def test_rqs_kernel(tag, left, right):
    g = load_golden("rqs_kernel")
    mask = np.ascontiguousarray(g[f"{tag}_mask"])
    B, V = g[f"{tag}_x"].shape[0], mask.size
    out = f32(g[f"{tag}_out"])
    prm = RqsParams(10, -5.0, 5.0, -4.0, 6.0, left, right)
    for parity in (0, 1):
        m = (mask if parity == 0 else 1 - mask)
        xa = f32(g[f"{tag}_x"] * m)
        y, logJ = np.empty_like(xa), np.empty(B, dtype=np.float32)
        assert H.cpu_rqs_fwd(fp(xa), fp(out), up(mask), parity, 0, prm, None, fp(y), fp(logJ), I64(B), I64(V)) == 0
        close(y, g[f"{tag}_p{parity}_fx"])
        close(logJ, g[f"{tag}_p{parity}_logJ"])
        gx, gout = np.empty_like(xa), np.empty_like(out)
        H.cpu_rqs_bwd(fp(xa), fp(out), up(mask), parity, 0, prm, fp(f32(g[f"{tag}_r"])), fp(f32(g[f"{tag}_c"])),
                      fp(gx), fp(gout), I64(B), I64(V))
        close_grad(gx, g[f"{tag}_p{parity}_gx"])
        close_grad(gout, g[f"{tag}_p{parity}_gout"])
        # inverse on the forward image: back to x (active sites), log-Jacobians cancel
        xi, li = np.empty_like(xa), np.empty(B, dtype=np.float32)
        H.cpu_rqs_inv(fp(f32(g[f"{tag}_p{parity}_fx"])), fp(out), up(mask), parity, 0, prm, None,
                      fp(xi), fp(li), I64(B), I64(V))
        if tag == "lin":    # monotone everywhere: exact inverse, log-Jacobians cancel
            close(xi, xa.astype(np.float64), tol=3e-5)
            close(li, -g[f"{tag}_p{parity}_logJ"], tol=3e-5)
        else:               # an end segment extended without 'linear' is not monotone outside xlim
            inside = (np.abs(xa) < 5.0) & m.astype(bool)
            close(xi[inside], xa[inside].astype(np.float64), tol=3e-5)


def _knots(wx, wy, wd, lim, both_ends=True):
    """Knot table exactly as the product builds it: the knot kernel's own arithmetic
    (nfk_knots.cuh) -- float32 table, cumulative softmax sums from both ends."""
    wx, wy = f32(wx), f32(wy)
    wd = None if wd is None else f32(wd)
    K = len(wx) + 1
    table = np.empty((5, K), dtype=np.float32)
    H.cpu_knots_fwd(fp(wx), fp(wy), fp(wd), K, ctypes.c_float(lim[0]), ctypes.c_float(lim[1] - lim[0]),
                    ctypes.c_float(lim[0]), ctypes.c_float(lim[1] - lim[0]), fp(table))
    return table if both_ends else table[:3].copy()


def test_knot_table_against_oracle_and_central_differences():
    """nfk_knots.cuh forward vs the oracle's splinenet_knots (SplineNet.make_spline, modules.py:369-391)
    and its adjoint vs central differences of the oracle, smooth and free derivatives."""
    from oracle import nf_oracle as O
    rs = np.random.RandomState(2)
    for K, smooth, lim in [(2, True, (0.0, 1.0)), (10, False, (0.5, 1.0)), (12, True, (0.0, 1.0)), (50, False, (-2.0, 3.0))]:
        wx, wy = f32(rs.randn(K - 1)), f32(rs.randn(K - 1))
        wd = None if smooth else f32(rs.randn(K) * 2)
        table = _knots(wx, wy, wd, lim)

        def ref_table(wx_, wy_, wd_):
            kx, ky, kd = O.splinenet_knots(wx_, wy_, wd_, xlim=lim, ylim=lim)
            if kd is None:
                kd = O.smooth_derivatives(kx, ky, 0)
            return np.stack([kx, ky, kd, lim[1] - kx, lim[1] - ky])
        args64 = [wx.astype(np.float64), wy.astype(np.float64)] + ([] if smooth else [wd.astype(np.float64)])
        ref = ref_table(args64[0], args64[1], None if smooth else args64[2])
        close(table, ref, tol=3e-7)
        assert table[0, -1] == np.float32(lim[1]) and table[3, -1] == 0
        r = f32(rs.randn(5, K))
        gwx, gwy = np.empty(K - 1, dtype=np.float32), np.empty(K - 1, dtype=np.float32)
        gwd = None if smooth else np.empty(K, dtype=np.float32)
        H.cpu_knots_bwd(fp(wx), fp(wy), fp(wd), K, ctypes.c_float(lim[0]), ctypes.c_float(lim[1] - lim[0]),
                        ctypes.c_float(lim[0]), ctypes.c_float(lim[1] - lim[0]), fp(r), fp(gwx), fp(gwy), fp(gwd))
        for j, got in enumerate([gwx, gwy] + ([] if smooth else [gwd])):
            num = np.empty_like(got, dtype=np.float64)
            for i in range(len(got)):
                vals = []
                for sgn in (1, -1):
                    a = [v.copy() for v in args64]
                    a[j][i] += sgn * 1e-6
                    vals.append((ref_table(a[0], a[1], None if smooth else a[2]) * r).sum())
                num[i] = (vals[0] - vals[1]) / 2e-6
            close_grad(got, num, tol=2e-6)


@pytest.mark.parametrize("tag,symmetric", [("zd_sym", True), ("lat_sym_smooth", True), ("lat_asym", False)])
def test_distconvertor(tag, symmetric):
    g = load_golden("distconv")
    wd = g[f"{tag}_wd"] if f"{tag}_wd" in g.files else None
    lim = (0.5, 1.0) if symmetric else (0.0, 1.0)
    kn = _knots(g[f"{tag}_wx"], g[f"{tag}_wy"], wd, lim)
    K = kn.shape[1]
    x = f32(g[f"{tag}_x"])
    B, V = x.shape[0], x[0].size
    left = 2 if symmetric else 0
    y, logJ = np.empty_like(x), np.empty(B, dtype=np.float32)
    H.cpu_spline1d_fwd(fp(x), fp(kn), K, left, 0, 1, 0, None, fp(y), fp(logJ), I64(B), I64(V))
    close(y, g[f"{tag}_y"])
    close(logJ, g[f"{tag}_logJ"])
    # inverse: ModuleList_.backward(y, log0=logJ) -> (x, ~0)
    xb, lb = np.empty_like(x), np.empty(B, dtype=np.float32)
    H.cpu_spline1d_fwd(fp(f32(g[f"{tag}_y"])), fp(kn), K, left, 0, 1, 1, fp(f32(g[f"{tag}_logJ"])),
                       fp(xb), fp(lb), I64(B), I64(V))
    close(xb, g[f"{tag}_x"], tol=2e-5)
    close(lb, np.zeros(B), tol=3e-5)
    # gradient w.r.t. x (knot gradients are checked through autograd in the gpu tests)
    gx, gk = np.empty_like(x), np.zeros(5 * K, dtype=np.float32)
    H.cpu_spline1d_bwd(fp(x), fp(kn), K, left, 0, 1, fp(f32(g[f"{tag}_r"])), fp(f32(g[f"{tag}_c"])),
                       fp(gx), fp(gk), I64(B), I64(V))
    close_grad(gx, g[f"{tag}_gx"])


def test_distconvertor_tails_keep_relative_accuracy():
    """|x| up to 15: a naive fp32 expit->spline->logit loses all digits there."""
    from oracle import nf_oracle as O
    rs = np.random.RandomState(0)
    wx, wy, wd = rs.randn(9) * 0.5, rs.randn(9) * 0.5, rs.randn(10) * 0.5
    for symmetric in (True, False):
        lim = (0.5, 1.0) if symmetric else (0.0, 1.0)
        kn = _knots(wx, wy, wd, lim)
        x = f32(np.linspace(-15, 15, 601)[None, :])
        yr, lr = O.distconvertor(x.astype(np.float64), 0.0, (f32(wx).astype(np.float64), f32(wy).astype(np.float64),
                                                              f32(wd).astype(np.float64)), symmetric=symmetric)
        y, lj = np.empty_like(x), np.empty(1, dtype=np.float32)
        H.cpu_spline1d_fwd(fp(x), fp(kn), 10, 2 if symmetric else 0, 0, 1, 0, None,
                           fp(y), fp(lj), I64(1), I64(x.size))
        close(y, yr, tol=2e-5)
        close(lj, lr, tol=2e-5)


def test_spline1d_plain_modes():
    """Non-logistic shared-knot spline, every extrapolation mode, vs spline golden."""
    g = load_golden("spline")
    kx = g["s1_kx"]
    kn = f32(np.stack([g["s1_kx"], g["s1_ky"], g["s1_kd"][len(kx) - 1:]]))
    x = f32(g["s1_x"][None, :])
    y, lj = np.empty_like(x), np.empty(1, dtype=np.float32)
    H.cpu_spline1d_fwd(fp(x), fp(kn), len(kx), 2, 0, 0, 0, None, fp(y), fp(lj), I64(1), I64(x.size))
    close(y[0], g["s1_y"], tol=2e-5)
    close(lj, np.log(g["s1_g"]).sum(), tol=2e-5)
    xi, li = np.empty_like(x), np.empty(1, dtype=np.float32)
    H.cpu_spline1d_fwd(fp(f32(g["s1_y"][None, :])), fp(kn), len(kx), 2, 0, 0, 1, None,
                       fp(xi), fp(li), I64(1), I64(x.size))
    inside = g["s1_x"] > 2 * kx[0] - kx[-1]   # beyond the mirrored range the end segment is extended
    close(xi[0][inside], g["s1_x"][inside], tol=5e-5)


def test_logistic_ops():
    x = f32(np.linspace(-12, 12, 97)[None, :])
    y, lj = np.empty_like(x), np.empty(1, dtype=np.float32)
    H.cpu_logistic_fwd(fp(x), 0, None, fp(y), fp(lj), I64(1), I64(x.size))
    xd = x.astype(np.float64)
    close(y, 1 / (1 + np.exp(-xd)))
    close(lj, (-xd + 2 * np.log(1 / (1 + np.exp(-xd)))).sum())
    p = f32(np.linspace(0.01, 0.99, 50)[None, :])
    H_y, H_l = np.empty_like(p), np.empty(1, dtype=np.float32)
    H.cpu_logistic_fwd(fp(p), 1, None, fp(H_y), fp(H_l), I64(1), I64(p.size))
    pd = p.astype(np.float64)
    close(H_y, np.log(pd / (1 - pd)))
    close(H_l, -np.log(pd * (1 - pd)).sum())


@pytest.mark.parametrize("tag", ["d1", "d2", "d3", "d4"])
def test_conv_layers(tag):
    g = load_golden("conv")
    n = len(g[f"{tag}_hidden"]) + 1
    h = f32(g[f"{tag}_x"])
    B, shape = h.shape[0], h.shape[2:]
    for i in range(n):
        w = f32(g[f"{tag}_w{i}"])
        b = f32(g[f"{tag}_b{i}"]) if f"{tag}_b{i}" in g.files else None
        Co, Ci = w.shape[:2]
        out = np.empty((B, Co) + shape, dtype=np.float32)
        H.cpu_conv_circ_fwd(fp(h), fp(w), 0, fp(b), None, 0, 1 if i < n - 1 else 0, None, 0, fp(out),
                            lattice(shape), 3, Ci, Co, I64(B))
        h = out
    close(h, g[f"{tag}_out"])


def test_whole_rqs_stack_through_harness():
    """Config-3-style stack (8x8 here): conv -> rqs, four steps, full-field mode."""
    g = load_golden("cpl_rqs_2d")
    mask = np.ascontiguousarray(g["mask"])
    x = f32(g["x"])
    B, shape, V = x.shape[0], x.shape[1:], mask.size
    log = np.zeros(B, dtype=np.float32)
    prm = RqsParams(10, -5.0, 5.0, -5.0, 5.0, 1, 1)
    for k in range(4):
        p = k % 2
        h = x.reshape((B, 1) + shape)
        for i in range(3):
            w = f32(g[f"blk0_step{k}_w{i}"])
            Co, Ci = w.shape[:2]
            out = np.empty((B, Co) + shape, dtype=np.float32)
            # first layer reads the frozen partition only: mask.split fused in
            H.cpu_conv_circ_fwd(fp(h), fp(w), 0, None, up(mask) if i == 0 else None, 0 if p == 0 else 1,
                                1 if i < 2 else 0, None, 0, fp(out), lattice(shape), 3, Ci, Co, I64(B))
            h = out
        if k < 4:
            close(h, g[f"blk0_step{k}_out"], tol=2e-5) if k == 0 else None
        y, lo = np.empty_like(x), np.empty(B, dtype=np.float32)
        H.cpu_rqs_fwd(fp(x), fp(h), up(mask), p, 1, prm, fp(log), fp(y), fp(lo), I64(B), I64(V))
        x, log = y, lo
    close(x, g["y"])
    close(log, g["logJ"])


@pytest.mark.parametrize("name,left,right", [("perleft", 3, 0), ("perright", 0, 3), ("perboth", 3, 3),
                                             ("perleft_antiright", 3, 2)])
def test_spline1d_periodic_mode(name, left, right):
    """kExtrapPeriodic in the kernels' functor: values against the reference fixture; the derivative through
    the VJP (negative on the mirror image, where the forward's log-derivative is NaN like the reference's)."""
    g = load_golden("spline_extra")
    kn = f32(np.stack([g["per_kx"], g["per_ky"], g["per_kd"]]))
    K = kn.shape[1]
    x = f32(g["per_x"][None, :])
    # beyond the mirror image the reference extends the outermost mirrored segment; the kernel extends the
    # matching in-range segment -- the same function -- unless the far side has its own rule: compare where
    # at most one reflection applies
    lo, hi = g["per_kx"][0], g["per_kx"][-1]
    ok = (x[0] > 2 * lo - hi) & (x[0] < 2 * hi - lo)
    y, lj = np.empty_like(x), np.empty(1, dtype=np.float32)
    H.cpu_spline1d_fwd(fp(x), fp(kn), K, left, right, 0, 0, None, fp(y), fp(lj), I64(1), I64(x.size))
    close(y[0][ok], g[f"{name}_y"][ok], tol=2e-5)
    assert np.isnan(lj[0])
    gx, gk = np.empty_like(x), np.zeros(5 * K, dtype=np.float32)
    ones, zero = f32(np.ones_like(x)), f32(np.zeros(1))
    H.cpu_spline1d_bwd(fp(x), fp(kn), K, left, right, 0, fp(ones), fp(zero), fp(gx), fp(gk), I64(1), I64(x.size))
    close(gx[0][ok], g[f"{name}_g"][ok], tol=5e-5)
    assert (gx[0][ok] < 0).any()


@pytest.mark.parametrize("logistic,left,right,lim", [(1, 2, 0, (0.5, 1.0)), (1, 0, 0, (0.0, 1.0)),
                                                      (0, 1, 1, (-2.0, 3.0)), (0, 2, 0, (0.0, 2.0)),
                                                      (0, 0, 2, (-1.0, 1.0))])
def test_spline1d_knot_gradients(logistic, left, right, lim):
    """VJP w.r.t. the explicit knots vs a complex-step derivative of the oracle.  For the
    logistic chain the kernel reports gradients for the knots measured from either end
    (k and hi - k are the same degree of freedom): fold them as g_k - g_c."""
    from oracle import nf_oracle as O
    rs = np.random.RandomState(11)
    K = 7
    wx, wy, wd = rs.randn(K - 1) * 0.6, rs.randn(K - 1) * 0.6, rs.randn(K) * 0.5
    kn = _knots(wx, wy, wd, lim, both_ends=bool(logistic))
    B, V = 5, 9
    x = f32(rs.randn(B, V) * (2.0 if logistic else 1.5) + (0 if logistic else 0.5 * (lim[0] + lim[1])))
    r, c = f32(rs.randn(B, V)), f32(rs.randn(B))
    names = {0: None, 1: 'linear', 2: 'anti'}
    extrap = {k: v for k, v in (('left', names[left]), ('right', names[right])) if v}

    def scalar(kx_, ky_, kd_):
        xx, log0 = x.astype(np.float64), 0.0
        if logistic:
            xx, log0 = O.expit_forward(xx, log0)
        xx, log0 = O.splinenet_forward(xx, log0, (kx_, ky_, kd_), extrap)
        if logistic:
            xx, log0 = O.logit_forward(xx, log0)
        return (xx * r).sum() + (log0 * c).sum()

    args = tuple(kn[i].astype(np.float64) for i in range(3))
    gx, gk = np.empty_like(x), np.zeros(5 * K, dtype=np.float32)
    H.cpu_spline1d_bwd(fp(x), fp(kn), K, left, right, logistic, fp(r), fp(c), fp(gx), fp(gk), I64(B), I64(V))
    ref = np.zeros(3 * K)
    for j in range(3):
        for i in range(K):
            if logistic and i in (0, K - 1) and j < 2:
                continue     # end knots are the constants xlim / ylim of DistConvertor_
            v = [np.zeros(K) for _ in range(3)]
            v[j][i] = 1.0
            ref[j * K + i] = O.directional_derivative(scalar, args, v)
    gk = gk.astype(np.float64)
    got = gk[:3 * K].copy()
    got[:K] -= gk[3 * K:4 * K]
    got[K:2 * K] -= gk[4 * K:5 * K]
    if logistic:
        for j in range(2):
            got[j * K] = got[j * K + K - 1] = 0.0
    close_grad(got, ref, tol=2e-5)


def test_distconvertor_narrow_top_bin_is_well_conditioned():
    """A narrow last bin with points inside it: with knots known only from the lower end a
    1-ulp knot error moves log J by ~4e-5; the two-ended table keeps it within tolerance."""
    g = load_golden("distconv")
    tag = "lat_asym"
    kn = _knots(g[f"{tag}_wx"], g[f"{tag}_wy"], g[f"{tag}_wd"], (0.0, 1.0))
    assert kn[0, -1] - kn[0, -2] < 0.03          # the narrow bin [0.9725, 1]
    x = f32(g[f"{tag}_x"])
    y, logJ = np.empty_like(x), np.empty(x.shape[0], dtype=np.float32)
    H.cpu_spline1d_fwd(fp(x), fp(kn), kn.shape[1], 0, 0, 1, 0, None, fp(y), fp(logJ), I64(x.shape[0]), I64(x[0].size))
    close(logJ, g[f"{tag}_logJ"])
    close(y, g[f"{tag}_y"])


@pytest.mark.parametrize("name,kind,P", [("cpl_rqs_2d", 1, 28), ("cpl_affine_2d", 0, 2)])
@pytest.mark.parametrize("R", [16, 4, 3])
def test_fused_2d_step_phases(name, kind, P, R):
    """The fused conditioner+transform kernel, phase by phase on the host: whole strips
    (R >= L0), several strips (R = 4) and a ragged last strip (R = 3), forward and inverse."""
    g = load_golden(name)
    mask = np.ascontiguousarray(g["mask"])
    x = f32(g["x"])
    B, (L0, L1) = x.shape[0], x.shape[1:]
    prm = RqsParams(10, -5.0, 5.0, -5.0, 5.0, 1, 1)
    log = np.zeros(B, dtype=np.float32)
    hist = []
    for k in range(4):
        w = [f32(g[f"blk0_step{k}_w{i}"]) for i in range(3)]
        y, lo = np.empty_like(x), np.empty(B, dtype=np.float32)
        assert H.cpu_fused2d_step(fp(x), fp(w[0]), None, fp(w[1]), None, fp(w[2]), None, 8, kind, prm, 0, k % 2, 0,
                                  fp(log), fp(y), fp(lo), L0, L1, I64(B), R) == 0
        hist.append((x, log, w))
        x, log = y, lo
        if k == 0:
            close(y * (mask if k % 2 == 0 else 1 - mask), g["blk0_step0_fx"], tol=2e-5)
    close(x, g["y"], tol=2e-5)
    close(log, g["logJ"], tol=2e-5)
    # inverse sweep restores the field and cancels the log-Jacobian
    for k in reversed(range(4)):
        xin, login, w = hist[k]
        xb, lb = np.empty_like(x), np.empty(B, dtype=np.float32)
        H.cpu_fused2d_step(fp(x), fp(w[0]), None, fp(w[1]), None, fp(w[2]), None, 8, kind, prm, 0, k % 2, 1,
                           fp(log), fp(xb), fp(lb), L0, L1, I64(B), R)
        x, log = xb, lb
    close(x, g["x"], tol=1e-4)
    close(log, np.zeros(B), tol=2e-4)


def test_fused_2d_step_with_bias_and_odd_mask_parity():
    """Biases, EvenOddMask(parity=1) and a non-square lattice against the unfused host ops."""
    rs = np.random.RandomState(21)
    B, L0, L1 = 2, 6, 12
    x = f32(rs.randn(B, L0, L1) * 1.5)
    w = [f32(rs.randn(8, 1, 3, 3) * 0.4), f32(rs.randn(8, 8, 3, 3) * 0.15), f32(rs.randn(28, 8, 3, 3) * 0.15)]
    b = [f32(rs.randn(8) * 0.2), f32(rs.randn(8) * 0.2), f32(rs.randn(28) * 0.2)]
    mask = np.empty((L0, L1), dtype=np.uint8)
    H.cpu_mask_evenodd(up(mask), lattice((L0, L1)), 1, -1)
    prm = RqsParams(10, -4.0, 4.0, -3.0, 5.0, 1, 1)
    for parity in (0, 1):
        h = x.reshape(B, 1, L0, L1)
        for i in range(3):
            Co, Ci = w[i].shape[:2]
            out = np.empty((B, Co, L0, L1), dtype=np.float32)
            H.cpu_conv_circ_fwd(fp(h), fp(w[i]), 0, fp(b[i]), up(mask) if i == 0 else None, 0 if parity == 0 else 1,
                                1 if i < 2 else 0, None, 0, fp(out), lattice((L0, L1)), 3, Ci, Co, I64(B))
            h = out
        yr, lr = np.empty_like(x), np.empty(B, dtype=np.float32)
        H.cpu_rqs_fwd(fp(x), fp(h), up(mask), parity, 1, prm, None, fp(yr), fp(lr), I64(B), I64(L0 * L1))
        y, lo = np.empty_like(x), np.empty(B, dtype=np.float32)
        assert H.cpu_fused2d_step(fp(x), fp(w[0]), fp(b[0]), fp(w[1]), fp(b[1]), fp(w[2]), fp(b[2]), 8, 1, prm, 1,
                                  parity, 0, None, fp(y), fp(lo), L0, L1, I64(B), 4) == 0
        close(y, yr.astype(np.float64), tol=2e-5)
        close(lo, lr.astype(np.float64), tol=2e-5)


@pytest.mark.parametrize("shape", [(8, 4), (6,), (4, 6, 3), (16, 9), (5, 1)])
@pytest.mark.parametrize("inverse", [0, 1])
def test_psd_weights_arithmetic(shape, inverse):
    """nfk_psd.cuh (the spectral-weight kernel's own code): w = ipsd^(-+1/2) and the log-Jacobian against the
    oracle's FFTNet_.log_jacobian (fftflow_.py:167-180); its adjoint against central differences of the oracle."""
    from oracle import nf_oracle as O
    rs = np.random.RandomState(len(shape) * 10 + inverse)
    ipsd = f32(rs.rand(*shape) * 3 + 0.05)
    Kc, Lh = ipsd.size, shape[-1]
    w, logj = np.empty_like(ipsd), np.empty(1, dtype=np.float32)
    H.cpu_psd_weights_fwd(fp(ipsd), I64(Kc), Lh, inverse, fp(w), fp(logj))

    def ref(v):
        wr = 1 / v ** 0.5
        lj = O.fftnet_log_jacobian(wr)
        return (1 / wr, -lj) if inverse else (wr, lj)
    wr, lr = ref(ipsd.astype(np.float64))
    close(w, wr, tol=2e-7)
    close(logj, lr, tol=1e-6)
    r, c = f32(rs.randn(*shape)), 0.7
    g = np.empty_like(ipsd)
    H.cpu_psd_weights_bwd(fp(ipsd), fp(w), fp(r), ctypes.c_float(c), I64(Kc), Lh, inverse, fp(g))
    num = np.empty(ipsd.shape)
    for idx in np.ndindex(*ipsd.shape):
        vals = []
        for sgn in (1, -1):
            v = ipsd.astype(np.float64).copy()
            v[idx] += sgn * 1e-6
            wv, lv = ref(v)
            vals.append((wv * r).sum() + c * lv)
        num[idx] = (vals[0] - vals[1]) / 2e-6
    close_grad(g, num, tol=5e-6)
