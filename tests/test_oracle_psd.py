"""Pin the oracle's PSDBlock_ / MeanFieldNet_ / FFTNet_ restatement against the reference's
own outputs (tests/golden/psd.npz, model_psd_affine.npz).  CPU only."""

import numpy as np
import pytest

from oracle import nf_oracle as O
from conftest import load_golden

ACTION = dict(kappa=0.67, m_sq=-4 * 0.67, lambd=0.5)

# tag -> (has a mean-field net, ignore_zeromode, FFTNet_ on its own)
CASES = {"ex2d": (True, True, False), "id3d": (False, False, False),
         "fft1d": (False, False, True), "fft2d": (False, True, True)}


def psd_params(g, tag):
    """(mf_weights | None, final_scale | None, ipsd spline weights, logy) of a psd.npz case."""
    def get(name):
        key = f"{tag}_w_{name}"
        return g[key] if key in g.files else None
    pre = "" if CASES[tag][2] else "fftnet_."
    ipsd_w = (get(pre + "ipsd_net.weights_x"), get(pre + "ipsd_net.weights_y"), get(pre + "ipsd_net.weights_d"))
    logy = get(pre + "ipsd_net.logy")
    mf_w = fs = None
    if CASES[tag][0]:
        mf_w = (get("mfnet_.dc_.1.weights_x"), get("mfnet_.dc_.1.weights_y"), get("mfnet_.dc_.1.weights_d"))
        fs = get("mfnet_.dc_.3._weight")
    return mf_w, fs, ipsd_w, logy


def oracle_run(g, tag, x, log0=0.0, inverse=False):
    mf_w, fs, ipsd_w, logy = psd_params(g, tag)
    lat = tuple(int(v) for v in g[f"{tag}_lat_shape"])
    k2 = O.lattice_k2(lat)
    ipsd = O.ipsd_forward(k2 / k2.max(), ipsd_w, logy, ignore_zeromode=CASES[tag][1])
    if CASES[tag][2]:
        return O.fftnet(x, log0, ipsd, inverse=inverse)
    return O.psdblock(x, log0, mf_weights=mf_w, mf_symmetric=True, mf_final_scale=fs, ipsd=ipsd,
                      inverse=inverse)


@pytest.mark.parametrize("tag", sorted(CASES))
def test_lattice_k2_and_ipsd_golden(tag):
    g = load_golden("psd")
    lat = tuple(int(v) for v in g[f"{tag}_lat_shape"])
    k2 = O.lattice_k2(lat)
    np.testing.assert_allclose(k2.max(), g[f"{tag}_max_lat_k2"], rtol=1e-14)
    np.testing.assert_allclose(k2 / k2.max(), g[f"{tag}_norm_lat_k2"], rtol=1e-13, atol=1e-15)
    _, _, ipsd_w, logy = psd_params(g, tag)
    ipsd = O.ipsd_forward(k2 / k2.max(), ipsd_w, logy, ignore_zeromode=CASES[tag][1])
    np.testing.assert_allclose(ipsd, g[f"{tag}_ipsd"], rtol=1e-11)


def test_lattice_k2_known_answer():
    """Docstring example of outer_lattice_k2 (fftflow_.py:337-343) uses k in (0, 1, 3 points);
    here the physical grid: k^2 = 4 sin^2(pi n / L) summed over axes."""
    k2 = O.lattice_k2((4, 4))
    assert k2.shape == (4, 3)
    np.testing.assert_allclose(k2[:, 0], [0, 2, 4, 2], atol=1e-15)
    np.testing.assert_allclose(k2[2, 2], 8.0, rtol=1e-15)


@pytest.mark.parametrize("tag", sorted(CASES))
def test_psd_forward_and_inverse_golden(tag):
    g = load_golden("psd")
    y, logJ = oracle_run(g, tag, g[f"{tag}_x"])
    np.testing.assert_allclose(y, g[f"{tag}_y"], rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(logJ * np.ones(len(y)), g[f"{tag}_logJ"], rtol=1e-10, atol=1e-10)
    xb, lb = oracle_run(g, tag, g[f"{tag}_y"], g[f"{tag}_logJ"], inverse=True)
    np.testing.assert_allclose(xb, g[f"{tag}_rt_x"], rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(lb, g[f"{tag}_rt_log"], atol=1e-7)
    yi, li = oracle_run(g, tag, g[f"{tag}_x"], inverse=True)
    np.testing.assert_allclose(yi, g[f"{tag}_inv_y"], rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(li * np.ones(len(y)), g[f"{tag}_inv_logJ"], rtol=1e-8, atol=1e-8)


def test_psd_hack_parts_golden():
    g = load_golden("psd")
    tag = "ex2d"
    x = g[f"{tag}_x"]
    mf_w, fs, ipsd_w, logy = psd_params(g, tag)
    x_mean = x.mean(axis=(1, 2)).reshape(-1, 1, 1)
    np.testing.assert_allclose(x_mean, g[f"{tag}_x_mean"], rtol=1e-13, atol=1e-15)
    y_mf, l_mf = O.meanfieldnet(x_mean, 0, np.prod(x.shape[1:]) ** 0.5, mf_w, symmetric=True, final_scale=fs)
    np.testing.assert_allclose(y_mf, g[f"{tag}_y_mf"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(l_mf, g[f"{tag}_logJ_mf"], rtol=1e-10, atol=1e-10)
    y_fft, l_fft = O.fftnet(x - x_mean, 0, g[f"{tag}_ipsd"])
    np.testing.assert_allclose(y_fft, g[f"{tag}_y_fft"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(l_fft, g[f"{tag}_logJ_fft"], rtol=1e-12)


@pytest.mark.parametrize("tag", ["ex2d", "fft2d"])
def test_psd_gradients_via_complex_step(tag):
    """Reference-autograd gradients of L = sum(y r) + sum(logJ c) w.r.t. x and every parameter,
    checked by complex-step differentiation of the oracle (np.fft is complex-analytic in x only
    through real linear maps, so x is perturbed with real central differences instead)."""
    g = load_golden("psd")
    r, c, x = g[f"{tag}_r"], g[f"{tag}_c"], g[f"{tag}_x"]

    def scalar(xv):
        y, logJ = oracle_run(g, tag, xv)
        return (y * r).sum() + (logJ * np.ones(len(xv)) * c).sum()

    gx = g[f"{tag}_gx"]
    rng = np.random.RandomState(3)
    for _ in range(6):
        idx = tuple(rng.randint(0, n) for n in x.shape)
        h = 1e-5
        xp, xm = x.copy(), x.copy()
        xp[idx] += h
        xm[idx] -= h
        np.testing.assert_allclose((scalar(xp) - scalar(xm)) / (2 * h), gx[idx], rtol=2e-6, atol=1e-8)
    # parameters: central differences on the stored weights
    names = [str(n) for n in g[f"{tag}_param_names"]]
    for name in names:
        key = f"{tag}_w_{name}"
        w0 = g[key].copy()
        ref = g[f"{tag}_grad_{name}"]
        gg = {k: g[k] for k in g.files}
        for i in range(len(w0)):
            h = 1e-6
            vals = []
            for sgn in (+1, -1):
                w = w0.copy()
                w[i] += sgn * h

                class _G:          # a view of the fixture with one weight replaced
                    files = g.files

                    def __getitem__(self, k, _w=w):
                        return _w if k == key else gg[k]
                y, logJ = oracle_run(_G(), tag, x)
                vals.append((y * r).sum() + (logJ * np.ones(len(x)) * c).sum())
            np.testing.assert_allclose((vals[0] - vals[1]) / (2 * h), ref[i], rtol=5e-5, atol=5e-7)


def test_model_psd_affine_golden():
    """The whole examples/scalar_affine.py net: every block's output, logq/logp/loss and the inverse."""
    g = load_golden("model_psd_affine")
    lat = tuple(int(v) for v in g["lat_shape"])
    x = g["x"]
    k2 = O.lattice_k2(lat)
    w = lambda n: g["w_" + n]
    ipsd = O.ipsd_forward(k2 / k2.max(), (w("0.fftnet_.ipsd_net.weights_x"), w("0.fftnet_.ipsd_net.weights_y"),
                                          w("0.fftnet_.ipsd_net.weights_d")), w("0.fftnet_.ipsd_net.logy"),
                          ignore_zeromode=True)
    mask = O.evenodd_mask(lat)
    steps = []
    for k in range(4):
        layers = [(w(f"2.nets.{k}.{2 * i}.weight"), None) for i in range(3)]
        steps.append(O.make_convact_step('affine', layers, ('tanh', 'tanh', None), mask))

    def flow(x, log0, inverse=False):
        blocks = [
            lambda v, l, inv: O.psdblock(v, l, mf_weights=(w("0.mfnet_.dc_.1.weights_x"), w("0.mfnet_.dc_.1.weights_y"), None),
                                         mf_symmetric=True, mf_final_scale=w("0.mfnet_.dc_.3._weight"), ipsd=ipsd,
                                         inverse=inv),
            lambda v, l, inv: O.distconvertor(v, l, (w("1.1.weights_x"), w("1.1.weights_y"), None), symmetric=True,
                                              inverse=inv),
            lambda v, l, inv: O.coupling_forward(v, l, mask, steps, inverse=inv),
            lambda v, l, inv: O.distconvertor(v, l, (w("3.1.weights_x"), w("3.1.weights_y"), None), symmetric=True,
                                              inverse=inv)]
        outs = []
        for blk in (reversed(blocks) if inverse else blocks):
            x, log0 = blk(x, log0, inverse)
            outs.append((x, log0))
        return x, log0, outs

    y, logJ, outs = flow(x, 0.0)
    for i, (yi, li) in enumerate(outs):
        np.testing.assert_allclose(yi, g[f"blk{i}_y"], rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(li, g[f"blk{i}_logJ"], rtol=1e-9, atol=1e-9)
    logq = O.normal_log_prob(x) - logJ
    logp = -O.phi4_action(y, **ACTION)
    np.testing.assert_allclose(logq, g["logq"], rtol=1e-9)
    np.testing.assert_allclose(logp, g["logp"], rtol=1e-9)
    np.testing.assert_allclose(O.kl_loss(logq, logp), g["loss"], rtol=1e-10)
    xb, lb, _ = flow(g["y"], g["blk3_logJ"], inverse=True)
    np.testing.assert_allclose(xb, g["inv_x"], rtol=1e-7, atol=1e-7)
    np.testing.assert_allclose(lb, g["inv_log"], atol=1e-6)
