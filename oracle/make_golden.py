"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference; the GPU box has no copy):

    python -m oracle.make_golden

What it does: puts /root/reference (+ the timm import shim in oracle/_shims) on sys.path,
builds the reference's own ``UNetModelSwin`` / ``create_gaussian_diffusion`` objects, loads
the deterministic synthetic weights of ``resshift_b200.weights.random_state_dict`` (strict),
and records small input/output vectors.  The reference has no tests or golden vectors of
its own, so these outputs of the reference itself are what pins the oracle.
Nothing here copies reference source; it only imports and executes it.
"""
from __future__ import annotations

import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("RESSHIFT_REFERENCE", "/root/reference"))
GOLD = ROOT / "tests" / "golden"


def _import_reference():
    sys.path.insert(0, str(ROOT / "oracle" / "_shims"))
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(ROOT))
    from models.unet import UNetModelSwin                      # noqa: E402  (reference)
    from models.script_util import create_gaussian_diffusion   # noqa: E402  (reference)
    import models.gaussian_diffusion as gd                     # noqa: E402  (reference)
    return UNetModelSwin, create_gaussian_diffusion, gd


def _sub(t: torch.Tensor, stride: int = 37) -> np.ndarray:
    return t.reshape(-1)[::stride].float().numpy().copy()


def _stats(t: torch.Tensor) -> np.ndarray:
    t = t.float()
    return np.array([t.mean().item(), t.std().item(), t.abs().max().item()], dtype=np.float64)


def main():
    from resshift_b200.config import preset
    from resshift_b200.weights import random_state_dict

    UNetModelSwin, create_gaussian_diffusion, gd = _import_reference()
    torch.set_grad_enabled(False)
    GOLD.mkdir(parents=True, exist_ok=True)

    def build(name, seed=0):
        ucfg, dcfg = preset(name)
        model = UNetModelSwin(**ucfg.to_kwargs()).eval()
        sd = random_state_dict(ucfg, seed)
        model.load_state_dict(sd, strict=True)                  # names AND shapes must match the reference
        return ucfg, dcfg, model, sd

    # 1. state_dict inventories of the shipped configs --------------------------------
    for name in ("realsr", "faceir", "inpaint"):
        ucfg, _ = preset(name)
        model = UNetModelSwin(**ucfg.to_kwargs())
        inv = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in model.state_dict().items()]
        nparam = sum(p.numel() for p in model.parameters())
        (GOLD / f"unet_keys_{name}.json").write_text(json.dumps({"n_params": nparam, "entries": inv}))
        print(name, len(inv), "entries", nparam, "params")

    # 2. single forward passes --------------------------------------------------------
    def forward_fixture(name, batch, tvals, fname, lq_hw=None, with_mask=False):
        ucfg, _, model, _ = build(name)
        g = torch.Generator().manual_seed(4321)
        x = torch.randn(batch, ucfg.in_channels, 64, 64, generator=g)
        hw = lq_hw or 64
        lq = torch.rand(batch, 3, hw, hw, generator=g) * 2 - 1
        mask = None
        if with_mask:
            mask = -torch.ones(batch, 1, hw, hw)
            mask[:, :, hw // 4: hw // 4 * 3, hw // 8: hw // 2] = 1.0
        t = torch.tensor(tvals, dtype=torch.long)
        probes = {}
        hooks = []
        for i, m in enumerate(model.input_blocks):
            hooks.append(m.register_forward_hook(lambda _m, _i, o, k=f"input_blocks.{i}": probes.__setitem__(k, o)))
        hooks.append(model.middle_block.register_forward_hook(lambda _m, _i, o: probes.__setitem__("middle_block", o)))
        for i, m in enumerate(model.output_blocks):
            hooks.append(m.register_forward_hook(lambda _m, _i, o, k=f"output_blocks.{i}": probes.__setitem__(k, o)))
        out = model(x, t, lq=lq, mask=mask) if with_mask else model(x, t, lq=lq)
        for h in hooks:
            h.remove()
        arrays = {"x": x.numpy(), "t": t.numpy(), "lq": lq.numpy(), "out": out.numpy()}
        if mask is not None:
            arrays["mask"] = mask.numpy()
        for k, v in probes.items():
            arrays[f"probe_stats/{k}"] = _stats(v)
            arrays[f"probe_sub/{k}"] = _sub(v)
        amax = max(float(v.abs().max()) for v in probes.values())
        np.savez_compressed(GOLD / fname, **arrays)
        print(fname, "out std %.4f" % out.std().item(), "max |act| %.2f" % amax)

    forward_fixture("tiny", 2, [3, 1], "unet_tiny.npz")
    forward_fixture("tiny_inpaint", 1, [2], "unet_tiny_inpaint.npz", lq_hw=256, with_mask=True)
    forward_fixture("realsr", 1, [14], "unet_realsr.npz")

    # 3. schedule tables --------------------------------------------------------------
    for name, steps in (("realsr", None), ("realsr_journal", None), ("realsr_journal", 15)):
        _, dcfg = preset(name, steps)
        diff = create_gaussian_diffusion(**dcfg.to_kwargs())
        T = diff.num_timesteps
        tt = torch.arange(T)
        one = torch.ones(T, 1)
        in_scale = diff._scale_input(one, tt)[:, 0].numpy()
        np.savez(GOLD / f"schedule_{name}_T{T}.npz",
                 sqrt_etas=diff.sqrt_etas, etas=diff.etas, coef1=diff.posterior_mean_coef1,
                 coef2=diff.posterior_mean_coef2, log_var=diff.posterior_log_variance_clipped,
                 in_scale=in_scale, kappa=np.array(diff.kappa))
        print("schedule", name, T, diff.sqrt_etas[:3])

    # 4. sampling trajectories through the reference's own p_sample_loop_progressive ----
    class IdentityAE(torch.nn.Module):
        """Stand-in first stage: the loop under test is the latent-space loop; the reference
        dereferences first_stage_model.parameters() unconditionally (gaussian_diffusion.py:502)."""
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1))
        def encode(self, x):
            return x
        def decode(self, x):
            return x

    def loop_fixture(name, steps, batch, fname, keep_steps):
        ucfg, dcfg, model, _ = build(name)
        if steps is not None:
            dcfg.steps = steps
        dcfg.sf = 1                                    # latent == LQ size: skips the bicubic pre-upsample only
        diff = create_gaussian_diffusion(**dcfg.to_kwargs())
        T = diff.num_timesteps
        g = torch.Generator().manual_seed(777)
        y = torch.rand(batch, 3, 64, 64, generator=g) * 2 - 1
        noises = [torch.randn(batch, 3, 64, 64, generator=g) for _ in range(T + 1)]
        queue = list(noises[1:])
        orig = gd.th.randn_like
        gd.th.randn_like = lambda ref: queue.pop(0)
        try:
            rec = list(diff.p_sample_loop_progressive(
                y, model, first_stage_model=IdentityAE(), noise=noises[0], noise_repeat=False,
                clip_denoised=False, denoised_fn=None, model_kwargs={"lq": y}, device="cpu"))
        finally:
            gd.th.randn_like = orig
        arrays = {"y": y.numpy(), "noises": torch.stack(noises).numpy(), "final": rec[-1]["sample"].numpy()}
        for k in keep_steps:                           # k = position in execution order (0 = t=T-1)
            arrays[f"pred_xstart/{k}"] = rec[k]["pred_xstart"].numpy()
            arrays[f"sample/{k}"] = rec[k]["sample"].numpy()
        np.savez_compressed(GOLD / fname, **arrays)
        print(fname, "final std %.4f" % rec[-1]["sample"].std().item())

    loop_fixture("tiny", 4, 2, "loop_tiny_T4.npz", keep_steps=[0, 1, 2, 3])
    loop_fixture("realsr", 15, 1, "loop_realsr_T15.npz", keep_steps=[0, 7, 14])


if __name__ == "__main__":
    main()
