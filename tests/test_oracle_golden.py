"""Pins the oracle (oracle/*.py, CPU fp32 restatement) against outputs of the reference itself.

The fixtures in tests/golden/ were produced by oracle/make_golden.py, which imports and runs the
unmodified reference (UNetModelSwin, create_gaussian_diffusion, p_sample_loop_progressive).
The reference ships no tests / known-answer vectors of its own.
"""
import json

import numpy as np
import pytest
import torch

from oracle import diffusion_oracle as do
from oracle import unet_oracle as uo
from resshift_b200.arch import unet_param_spec
from resshift_b200.config import preset
from resshift_b200.weights import random_state_dict

TOL = 2e-4   # fp32 CPU vs fp32 CPU, different op order (oracle is functional, reference is nn.Module)


@pytest.mark.parametrize("name", ["realsr", "faceir", "inpaint"])
def test_param_inventory_matches_reference(golden_dir, name):
    gold = json.loads((golden_dir / f"unet_keys_{name}.json").read_text())
    ucfg, _ = preset(name)
    spec = unet_param_spec(ucfg)
    mine = sorted((n, list(s)) for n, s, _ in spec)
    ref = sorted((n, s) for n, s, _ in gold["entries"])
    assert mine == ref
    n_params = sum(int(np.prod(s)) for n, s, r in spec if not r.startswith("buf_"))
    assert n_params == gold["n_params"]


def _check_forward(golden_dir, preset_name, fname):
    g = np.load(golden_dir / fname)
    ucfg, _ = preset(preset_name)
    sd = random_state_dict(ucfg, 0)
    probes = {}
    from tests.golden_util import golden_inputs
    x, t, lq, mask = golden_inputs(g)
    out = uo.unet_forward(sd, ucfg, x, t, lq=lq, mask=mask, probes=probes)
    assert np.abs(out.numpy() - g["out"]).max() < TOL
    for k, v in probes.items():
        ref_sub = g[f"probe_sub/{k}"]
        got = v.reshape(-1)[::37].numpy()
        assert np.abs(got - ref_sub).max() < TOL * max(1.0, float(np.abs(ref_sub).max())), k


def test_unet_tiny_forward(golden_dir):
    _check_forward(golden_dir, "tiny", "unet_tiny.npz")


def test_unet_tiny_inpaint_forward(golden_dir):
    _check_forward(golden_dir, "tiny_inpaint", "unet_tiny_inpaint.npz")


def test_unet_tiny_faceir_forward(golden_dir):
    """Face-restoration topology (8 latent channels, three-stage feature extractor on a 512x512 LQ input)."""
    _check_forward(golden_dir, "tiny_faceir", "unet_tiny_faceir.npz")


def test_unet_realsr_forward(golden_dir):
    _check_forward(golden_dir, "realsr", "unet_realsr.npz")


@pytest.mark.parametrize("name,fname", [("faceir", "unet_faceir.npz"), ("inpaint", "unet_inpaint.npz"),
                                         ("realsr", "unet_realsr_64x128.npz"), ("tiny", "unet_tiny_128x64.npz")])
def test_unet_round2_fixtures(golden_dir, name, fname):
    """Full-width face-restoration / inpainting topologies and non-square latents, reference-generated."""
    _check_forward(golden_dir, name, fname)


@pytest.mark.parametrize("name,steps,T", [("realsr", None, 15), ("realsr_journal", None, 4), ("realsr_journal", 15, 15)])
def test_schedule_tables(golden_dir, name, steps, T):
    g = np.load(golden_dir / f"schedule_{name}_T{T}.npz")
    _, d = preset(name, steps)
    se = do.eta_schedule(d.steps, d.min_noise_level, d.etas_end, d.kappa, d.schedule_kwargs["power"])
    tabs = do.schedule_tables(se, d.kappa)
    for k in ("sqrt_etas", "etas", "coef1", "coef2", "log_var"):
        np.testing.assert_allclose(tabs[k], g[k], rtol=1e-12, atol=0)
    np.testing.assert_allclose(tabs["in_scale"], g["in_scale"], rtol=1e-6)


def test_schedule_known_answers():
    """SURVEY.md Appendix B constants (realsr, T=15)."""
    se = do.eta_schedule(15, 0.04, 0.99, 2.0, 0.3)
    tabs = do.schedule_tables(se, 2.0)
    np.testing.assert_allclose(se[[0, 1, 7, 14]], [0.02, 0.11717, 0.47586, 0.99], atol=1e-5)
    np.testing.assert_allclose(tabs["coef1"][[0, 1, 14]], [0.0, 0.02914, 0.84233], atol=1e-5)
    np.testing.assert_allclose(tabs["std"][[1, 14]], [0.03941, 0.72158], atol=1e-5)
    np.testing.assert_allclose(tabs["in_scale"][[0, 14]], [0.99920, 0.45082], atol=1e-5)
    assert abs(2.0 * se[-1] - 1.98) < 1e-12


def _check_loop(golden_dir, preset_name, steps, fname):
    g = np.load(golden_dir / fname)
    ucfg, d = preset(preset_name, steps)
    sd = random_state_dict(ucfg, 0)
    se = do.eta_schedule(d.steps, d.min_noise_level, d.etas_end, d.kappa, d.schedule_kwargs["power"])
    tabs = do.schedule_tables(se, d.kappa)
    from tests.golden_util import golden_loop_inputs
    y, noises = golden_loop_inputs(g)
    noises = list(noises)
    rec = []
    final = do.p_sample_loop(lambda x, t: uo.unet_forward(sd, ucfg, x, t, lq=y), y, noises, tabs, d.kappa, rec)
    for key in g.files:
        if key.startswith("pred_xstart/"):
            k = int(key.split("/")[1])
            assert np.abs(rec[k]["pred_xstart"].numpy() - g[key]).max() < 5 * TOL, key
            assert np.abs(rec[k]["sample"].numpy() - g[f"sample/{k}"]).max() < 5 * TOL, key
    assert np.abs(final.numpy() - g["final"]).max() < 5 * TOL


def test_loop_tiny(golden_dir):
    _check_loop(golden_dir, "tiny", 4, "loop_tiny_T4.npz")


def test_loop_realsr_15(golden_dir):
    _check_loop(golden_dir, "realsr", 15, "loop_realsr_T15.npz")


def test_loop_realsr_15_batch2(golden_dir):
    _check_loop(golden_dir, "realsr", 15, "loop_realsr_T15_b2.npz")
