"""ORACLE (test infrastructure, never shipped / never measured as the product).

CPU fp32 functional restatement of the VQ-GAN first stage around ResShift's denoising loop — ``VQModelTorch.encode`` /
``.decode`` (reference ldm/models/autoencoder.py:28-40) with ``Encoder`` / ``Decoder`` / ``ResnetBlock`` / ``AttnBlock``
(ldm/modules/diffusionmodules/model.py:452-660, 90-149, 152-203), ``VectorQuantizer2.forward``
(ldm/modules/vqvae/quantize.py:271-312) and the bicubic pre-upsample of ``encode_first_stage``
(models/gaussian_diffusion.py:500-515) — working directly on a reference-named ``state_dict``.
Pinned against outputs of the imported reference (``oracle/make_golden_vq.py`` -> ``tests/golden/vq_*.npz``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s baseline legs may import this module.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from resshift_b200.vq_arch import VQConfig, decoder_blocks, encoder_blocks

SD = Dict[str, torch.Tensor]


def _norm(x, sd: SD, name: str):
    """Normalize = GroupNorm(32, eps=1e-6, affine) — reference model.py:46-47."""
    return F.group_norm(x, 32, sd[f"{name}.weight"], sd[f"{name}.bias"], eps=1e-6)


def _swish(x):
    return x * torch.sigmoid(x)          # reference model.py:41-43


def _conv(x, sd: SD, name: str, stride: int = 1, padding: Optional[int] = None):
    w = sd[f"{name}.weight"]
    return F.conv2d(x, w, sd[f"{name}.bias"], stride=stride, padding=w.shape[-1] // 2 if padding is None else padding)


def resnet_block(x, sd: SD, p: str):
    """reference model.py:127-149 with temb = None."""
    h = _conv(_swish(_norm(x, sd, f"{p}.norm1")), sd, f"{p}.conv1")
    h = _conv(_swish(_norm(h, sd, f"{p}.norm2")), sd, f"{p}.conv2")
    if f"{p}.nin_shortcut.weight" in sd:
        x = _conv(x, sd, f"{p}.nin_shortcut")
    return x + h


def attn_block(x, sd: SD, p: str):
    """Single-head self-attention over all H*W positions — reference model.py:180-203."""
    h_ = _norm(x, sd, f"{p}.norm")
    q, k, v = (_conv(h_, sd, f"{p}.{n}") for n in ("q", "k", "v"))
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + _conv(h_, sd, f"{p}.proj_out")


def downsample(x, sd: SD, p: str):
    """pad (0,1,0,1) then 3x3 stride-2 conv without padding — reference model.py:78-87."""
    return _conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0), sd, f"{p}.conv", stride=2, padding=0)


def upsample(x, sd: SD, p: str):
    """nearest x2 then 3x3 conv — reference model.py:62-66."""
    return _conv(F.interpolate(x, scale_factor=2.0, mode="nearest"), sd, f"{p}.conv")


@torch.no_grad()
def encoder(x, sd: SD, cfg: VQConfig):
    """reference model.py:533-559."""
    h = _conv(x, sd, "encoder.conv_in")
    for i, blocks, down in encoder_blocks(cfg):
        for j in range(len(blocks)):
            h = resnet_block(h, sd, f"encoder.down.{i}.block.{j}")
        if down:
            h = downsample(h, sd, f"encoder.down.{i}.downsample")
    h = resnet_block(h, sd, "encoder.mid.block_1")
    h = attn_block(h, sd, "encoder.mid.attn_1")
    h = resnet_block(h, sd, "encoder.mid.block_2")
    return _conv(_swish(_norm(h, sd, "encoder.norm_out")), sd, "encoder.conv_out")


@torch.no_grad()
def decoder(z, sd: SD, cfg: VQConfig):
    """reference model.py:626-660."""
    h = _conv(z, sd, "decoder.conv_in")
    h = resnet_block(h, sd, "decoder.mid.block_1")
    h = attn_block(h, sd, "decoder.mid.attn_1")
    h = resnet_block(h, sd, "decoder.mid.block_2")
    for i, blocks, up in decoder_blocks(cfg):
        for j in range(len(blocks)):
            h = resnet_block(h, sd, f"decoder.up.{i}.block.{j}")
        if up:
            h = upsample(h, sd, f"decoder.up.{i}.upsample")
    return _conv(_swish(_norm(h, sd, "decoder.norm_out")), sd, "decoder.conv_out")


@torch.no_grad()
def quantize(z, sd: SD):
    """Nearest codebook entry (returns z_q and the indices) — reference quantize.py:271-284: the distance is evaluated
    as  z^2 + e^2 - 2 z.e  and argmin takes the first minimum."""
    emb = sd["quantize.embedding.weight"]
    zf = z.permute(0, 2, 3, 1).contiguous()
    flat = zf.view(-1, emb.shape[1])
    d = torch.sum(flat ** 2, dim=1, keepdim=True) + torch.sum(emb ** 2, dim=1) - 2 * torch.einsum("bd,dn->bn", flat, emb.t())
    idx = torch.argmin(d, dim=1)
    z_q = emb[idx].view(zf.shape).permute(0, 3, 1, 2).contiguous()
    return z_q, idx.view(z.shape[0], z.shape[2], z.shape[3])


@torch.no_grad()
def vq_encode(x, sd: SD, cfg: VQConfig):
    """VQModelTorch.encode — reference autoencoder.py:28-31."""
    return _conv(encoder(x, sd, cfg), sd, "quant_conv")


@torch.no_grad()
def vq_decode(h, sd: SD, cfg: VQConfig, force_not_quantize: bool = False, return_indices: bool = False):
    """VQModelTorch.decode — reference autoencoder.py:33-40."""
    idx = None
    if not force_not_quantize:
        h, idx = quantize(h, sd)
    out = decoder(_conv(h, sd, "post_quant_conv"), sd, cfg)
    return (out, idx) if return_indices else out


def bicubic_upsample(y, sf: int):
    """encode_first_stage's pre-upsample — reference models/gaussian_diffusion.py:503-504."""
    return F.interpolate(y, scale_factor=sf, mode="bicubic")
