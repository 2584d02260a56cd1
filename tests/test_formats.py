"""On-disk compatibility with the reference: a training snapshot and a weights
blob WRITTEN BY THE REFERENCE (tests/golden/ref_affine.E3.tar, ref_affine_blob.txt, made by
tests/golden/make_golden.py) load into this package's nets, the state_dict layouts agree key by key,
and snapshots written here have the reference's format.  The CPU part checks layout and I/O, the GPU
part that the loaded flow reproduces the reference's outputs and that training resumes."""

import os
import shutil

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden

import normflow__b200 as nf  # noqa: F401
from normflow__b200 import Model
from normflow__b200.action import ScalarPhi4Action
from normflow__b200.mask import EvenOddMask
from normflow__b200.nn import (ModuleList_, ConvAct, AffineCoupling_, DistConvertor_,
                               FFTNet_, MeanFieldNet_, PSDBlock_)
from normflow__b200.prior import NormalPrior

ACTION = dict(kappa=0.67, m_sq=-4 * 0.67, lambd=0.5)
SNAP = os.path.join(GOLDEN, "ref_affine.E3.tar")
BLOB = os.path.join(GOLDEN, "ref_affine_blob.txt")


def example_model(lat=(8, 8)):
    mf = MeanFieldNet_.build(knots_len=10, symmetric=True, final_scale=True, smooth=True)
    ff = FFTNet_.build(lat, knots_len=10, ignore_zeromode=True)
    conv = dict(in_channels=1, out_channels=2, hidden_sizes=[8, 8], kernel_size=3, padding_mode='circular',
                conv_dim=len(lat), acts=('tanh', 'tanh', None), bias=False)
    net_ = ModuleList_([
        PSDBlock_(mfnet_=mf, fftnet_=ff),
        DistConvertor_(12, symmetric=True, smooth=True),
        AffineCoupling_([ConvAct(**conv) for _ in range(4)], mask=EvenOddMask(shape=lat)),
        DistConvertor_(12, symmetric=True, smooth=True)])
    return Model(net_=net_, prior=NormalPrior(shape=lat), action=ScalarPhi4Action(**ACTION))


def _layout(sd):
    return [(k, tuple(v.shape)) for k, v in sd.items()]


def test_state_dict_layout_equals_reference_key_by_key():
    g = load_golden("snapshot_layout")
    ref = [(str(k), tuple(int(n) for n in s.split(",")) if s else ()) for k, s in zip(g["keys"], g["shapes"])]
    assert _layout(example_model().net_.state_dict()) == ref
    # 4-D: Conv4d keeps the reference's `_conv_lower_dim.weight` parameter (convNd.py:59-82)
    lat4 = (4, 4, 4, 4)
    conv4 = dict(in_channels=1, out_channels=2, hidden_sizes=[4], kernel_size=3, padding_mode='circular',
                 conv_dim=4, acts=('tanh', None), bias=True)
    net4 = ModuleList_([AffineCoupling_([ConvAct(**conv4) for _ in range(2)], mask=EvenOddMask(shape=lat4))])
    ref4 = [(str(k), tuple(int(n) for n in s.split(",")) if s else ()) for k, s in zip(g["keys4d"], g["shapes4d"])]
    assert _layout(net4.state_dict()) == ref4


def test_reference_snapshot_file_loads(tmp_path):
    snap = torch.load(SNAP, map_location="cpu")
    assert set(snap) == {"MODEL_STATE", "EPOCHS_RUN"} and snap["EPOCHS_RUN"] == 3
    model = example_model()
    path = tmp_path / "run.E3.tar"
    shutil.copy(SNAP, path)
    model.fit.checkpoint_dict.update(snapshot_path=str(path))
    model.fit._load_snapshot()
    assert model.fit.checkpoint_dict['epochs_run'] == 3
    sd = model.net_.state_dict()
    for k, v in snap["MODEL_STATE"].items():
        if torch.is_floating_point(v):
            assert sd[k].dtype == torch.float32                          # float64 on disk, float32 here
            np.testing.assert_array_equal(sd[k].cpu().numpy(), v.float().numpy())
        else:
            assert sd[k].dtype == v.dtype and torch.equal(sd[k].cpu(), v)    # uint8 masks bit-exact
    # and a snapshot written here has the reference's format and naming (path.rsplit('.', 2)[0] + .E<n>.tar)
    model.fit._save_snapshot(2)
    out = tmp_path / "run.E5.tar"
    assert out.exists()
    again = torch.load(out, map_location="cpu")
    assert set(again) == {"MODEL_STATE", "EPOCHS_RUN"} and again["EPOCHS_RUN"] == 5
    assert list(again["MODEL_STATE"].keys()) == list(snap["MODEL_STATE"].keys())


def test_reference_weights_blob_loads_and_round_trips():
    model = example_model()
    blob = open(BLOB).read()
    model.net_.set_weights_blob(blob)
    snap = torch.load(SNAP, map_location="cpu")["MODEL_STATE"]
    sd = model.net_.state_dict()
    for k, v in snap.items():
        np.testing.assert_array_equal(sd[k].cpu().numpy(), v.to(sd[k].dtype).numpy())
    other = example_model()
    other.net_.set_weights_blob(model.net_.get_weights_blob())
    for (k, a), (_, b) in zip(other.net_.state_dict().items(), sd.items()):
        assert torch.equal(a, b), k


@pytest.mark.gpu
def test_loaded_reference_snapshot_reproduces_reference_outputs(tmp_path):
    from test_gpu_parity import cu, close
    g = load_golden("snapshot_layout")
    model = example_model()
    model.device_handler.to("cuda")
    path = tmp_path / "run.E3.tar"
    shutil.copy(SNAP, path)
    model.fit.checkpoint_dict.update(snapshot_path=str(path))
    model.fit._load_snapshot()
    with torch.no_grad():
        y, logJ = model.net_(cu(g["x"]))
    close(y, g["y"])
    close(logJ, g["logJ"])
    # resume: the epoch counter continues from the snapshot and the next one is written as .E5.tar
    model.fit(n_epochs=2, batch_size=64, save_every=2,
              checkpoint_dict=dict(print_stride=10, snapshot_path=str(path)))
    assert (tmp_path / "run.E5.tar").exists()
    assert torch.load(tmp_path / "run.E5.tar", map_location="cpu")["EPOCHS_RUN"] == 5


@pytest.mark.gpu
def test_transfer_to_larger_lattice_keeps_the_conditioners():
    """ModuleList_.transfer (nn/_core.py:105-106; couplings_.py:97-103; fftflow_.py:187-212): lattice-size
    transfer learning -- convolution kernels are size independent, masks and momentum grids are rebuilt."""
    model = example_model((8, 8))
    model.device_handler.to("cuda")
    big = model.net_.transfer(scale_factor=1, shape=(16, 16), mask=EvenOddMask(shape=(16, 16)))
    assert big[0].fftnet_.lat_shape == (16, 16)
    assert tuple(big[2].mask._mask.shape) == (16, 16)
    for a, b in zip(model.net_[2].nets, big[2].nets):
        for pa, pb in zip(a.parameters(), b.parameters()):
            assert torch.equal(pa, pb) and pa.data_ptr() != pb.data_ptr()
    with torch.no_grad():
        x = torch.randn(3, 16, 16, device="cuda")
        y, logJ = big(x)
        xb, lb = big.backward(y, logJ)
    assert y.shape == (3, 16, 16) and float((xb - x).abs().max()) < 1e-4 and float(lb.abs().max()) < 1e-3
