"""Generate the VQ-GAN fixtures under tests/golden/ by running the UNMODIFIED reference (build container only):

    python -m oracle.make_golden_vq

Imports the reference's own ``VQModelTorch`` (ldm/models/autoencoder.py), loads the deterministic synthetic weights of
``resshift_b200.vq_arch.random_vq_state_dict`` strictly (names, shapes AND order are asserted against the reference's
``state_dict``) and records encode / decode outputs, the code indices, and torch's bicubic x4 (the pre-upsample of
``encode_first_stage``, models/gaussian_diffusion.py:503-504).  Nothing here copies reference source.
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


def main():
    sys.path.insert(0, str(ROOT / "oracle" / "_shims"))
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(ROOT))
    from ldm.models.autoencoder import VQModelTorch          # noqa: E402  (reference)
    import torch.nn.functional as F
    from resshift_b200.vq_arch import random_vq_state_dict, vq_param_spec, vq_preset

    torch.set_grad_enabled(False)
    GOLD.mkdir(parents=True, exist_ok=True)

    inv = {}
    for name in ("f4", "f8_face"):
        cfg = vq_preset(name)
        m = VQModelTorch(**cfg.to_kwargs())
        inv[name] = [[k, list(v.shape)] for k, v in m.state_dict().items()]
        assert [(k, tuple(s)) for k, s in inv[name]] == [(n, tuple(s)) for n, s, _ in vq_param_spec(cfg)]
    (GOLD / "vq_keys.json").write_text(json.dumps(inv))

    def fixture(name, batch, hw, fname, seed=0):
        cfg = vq_preset(name)
        model = VQModelTorch(**cfg.to_kwargs()).eval()
        sd = random_vq_state_dict(cfg, seed)
        model.load_state_dict(sd, strict=True)
        g = torch.Generator().manual_seed(2468)
        x = torch.rand(batch, 3, hw, hw, generator=g) * 2 - 1
        lat = hw // cfg.downscale
        z = torch.randn(batch, cfg.embed_dim, lat, lat, generator=g) * 0.6
        enc = model.encode(x)
        quant, _, info = model.quantize(z)
        dec = model.decode(z)
        dec_nq = model.decode(z, force_not_quantize=True)
        idx = info[2].view(batch, lat, lat)
        # margin between the best and the second-best code per position: small margins are where an fp16 path may flip
        emb = sd["quantize.embedding.weight"]
        flat = z.permute(0, 2, 3, 1).reshape(-1, cfg.embed_dim)
        d = (flat ** 2).sum(1, keepdim=True) + (emb ** 2).sum(1) - 2 * flat @ emb.t()
        top2 = torch.topk(d, 2, dim=1, largest=False).values
        np.savez_compressed(GOLD / fname, x=x.numpy(), z=z.numpy(), enc=enc.numpy(), dec=dec.numpy(), dec_nq=dec_nq.numpy(),
                            idx=idx.numpy().astype(np.int32), quant=quant.numpy(),
                            margin=(top2[:, 1] - top2[:, 0]).view(batch, lat, lat).numpy())
        print(fname, "enc std %.3f" % enc.std().item(), "dec std %.3f" % dec.std().item(), "codes used", idx.unique().numel())

    fixture("tiny", 2, 64, "vq_tiny.npz")
    fixture("f4", 1, 64, "vq_f4_64.npz")
    fixture("f8_face", 1, 128, "vq_f8_face_128.npz")

    g = torch.Generator().manual_seed(1357)
    y = torch.rand(2, 3, 16, 24, generator=g) * 2 - 1
    np.savez_compressed(GOLD / "bicubic_x4.npz", y=y.numpy(), up=F.interpolate(y, scale_factor=4, mode="bicubic").numpy(),
                        up2=F.interpolate(y, scale_factor=2, mode="bicubic").numpy())
    print("bicubic ok")


if __name__ == "__main__":
    main()
