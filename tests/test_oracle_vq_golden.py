"""Pins the VQ-GAN oracle (oracle/vq_oracle.py) against outputs of the reference's own VQModelTorch
(oracle/make_golden_vq.py -> tests/golden/vq_*.npz, bicubic_x4.npz)."""
import json

import numpy as np
import pytest
import torch

from oracle import vq_oracle as vo
from resshift_b200.vq_arch import random_vq_state_dict, vq_param_spec, vq_preset

TOL = 2e-4


@pytest.mark.parametrize("name", ["f4", "f8_face"])
def test_vq_param_inventory_matches_reference(golden_dir, name):
    gold = json.loads((golden_dir / "vq_keys.json").read_text())[name]
    assert [(k, list(s)) for k, s, _ in vq_param_spec(vq_preset(name))] == [(k, s) for k, s in gold]


@pytest.mark.parametrize("name,fname", [("tiny", "vq_tiny.npz"), ("f4", "vq_f4_64.npz"), ("f8_face", "vq_f8_face_128.npz")])
def test_vq_encode_decode(golden_dir, name, fname):
    g = np.load(golden_dir / fname)
    cfg = vq_preset(name)
    sd = random_vq_state_dict(cfg, 0)
    x, z = torch.from_numpy(g["x"]), torch.from_numpy(g["z"])
    assert np.abs(vo.vq_encode(x, sd, cfg).numpy() - g["enc"]).max() < TOL
    zq, idx = vo.quantize(z, sd)
    assert np.array_equal(idx.numpy(), g["idx"])
    assert np.abs(zq.numpy() - g["quant"]).max() < 1e-6
    assert np.abs(vo.vq_decode(z, sd, cfg).numpy() - g["dec"]).max() < TOL
    assert np.abs(vo.vq_decode(z, sd, cfg, force_not_quantize=True).numpy() - g["dec_nq"]).max() < TOL


def test_bicubic(golden_dir):
    g = np.load(golden_dir / "bicubic_x4.npz")
    y = torch.from_numpy(g["y"])
    assert np.abs(vo.bicubic_upsample(y, 4).numpy() - g["up"]).max() < 1e-6
    assert np.abs(vo.bicubic_upsample(y, 2).numpy() - g["up2"]).max() < 1e-6
