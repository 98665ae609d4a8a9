"""ORACLE (test infrastructure, never shipped / never measured as the product).

A CPU, fp32, functional restatement of ``UNetModelSwin.forward`` — the denoiser on
ResShift's hot path — working directly on a reference-named ``state_dict``.
Every function cites the reference file:line it follows.  Pinned against outputs of the
imported reference itself (``oracle/make_golden.py`` -> ``tests/golden/*.npz``), see
``tests/test_oracle_golden.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from resshift_b200.arch import unet_block_plan, swin_geometry, shifted_window_mask
from resshift_b200.config import UNetConfig

SD = Dict[str, torch.Tensor]


def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    """reference models/basic_ops.py:99-117 — [cos | sin] halves, fp32."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def group_norm(x: torch.Tensor, sd: SD, name: str) -> torch.Tensor:
    """reference models/basic_ops.py:15-17,89-96 — 32 groups, eps 1e-5, fp32 math."""
    return F.group_norm(x.float(), 32, sd[f"{name}.weight"], sd[f"{name}.bias"], eps=1e-5)


def conv(x, sd: SD, name: str, stride: int = 1):
    w = sd[f"{name}.weight"]
    return F.conv2d(x, w, sd[f"{name}.bias"], stride=stride, padding=w.shape[-1] // 2)


def res_block(x, emb, sd: SD, p: str):
    """reference models/unet.py:186-206 with use_scale_shift_norm=True, no up/down."""
    h = conv(F.silu(group_norm(x, sd, f"{p}.in_layers.0")), sd, f"{p}.in_layers.2")
    e = F.linear(F.silu(emb), sd[f"{p}.emb_layers.1.weight"], sd[f"{p}.emb_layers.1.bias"])
    scale, shift = torch.chunk(e[:, :, None, None], 2, dim=1)
    h = group_norm(h, sd, f"{p}.out_layers.0") * (1 + scale) + shift
    h = conv(F.silu(h), sd, f"{p}.out_layers.3")
    if f"{p}.skip_connection.weight" in sd:
        x = conv(x, sd, f"{p}.skip_connection")
    return x + h


def window_attention(xw, sd: SD, p: str, heads: int, mask: Optional[torch.Tensor]):
    """reference models/swin_transformer.py:114-145.  xw: [nWin_total, N, C]."""
    bw, n, c = xw.shape
    hd = c // heads
    qkv = F.linear(xw, sd[f"{p}.qkv.weight"], sd[f"{p}.qkv.bias"]).view(bw, n, 3, heads, hd)
    q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))          # [bw, heads, n, hd]
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    table = sd[f"{p}.relative_position_bias_table"]                      # [(2w-1)^2, heads]
    index = sd[f"{p}.relative_position_index"].reshape(-1)
    attn = attn + table[index].view(n, n, heads).permute(2, 0, 1)[None]
    if mask is not None:
        nw = mask.shape[0]
        attn = (attn.view(bw // nw, nw, heads, n, n) + mask[None, :, None]).view(bw, heads, n, n)
    attn = attn.softmax(dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(bw, n, c)
    return F.linear(out, sd[f"{p}.proj.weight"], sd[f"{p}.proj.bias"])


def swin_block(x, sd: SD, p: str, heads: int, win: int, shift: int):
    """reference models/swin_transformer.py:238-281 (NCHW in, NCHW out)."""
    b, c, h, w = x.shape
    y = group_norm(x, sd, f"{p}.norm1")
    if shift:
        y = torch.roll(y, shifts=(-shift, -shift), dims=(2, 3))
    # window_partition (reference :35-47)
    yw = y.view(b, c, h // win, win, w // win, win).permute(0, 2, 4, 3, 5, 1).reshape(-1, win * win, c)
    # the reference always rebuilds the mask from the runtime size (:262-265); it is all-zero for shift 0
    mask = shifted_window_mask(h, w, win, shift).to(y.device) if shift else None
    aw = window_attention(yw, sd, f"{p}.attn", heads, mask)
    # window_reverse (reference :49-63)
    y = aw.view(b, h // win, w // win, win, win, c).permute(0, 5, 1, 3, 2, 4).reshape(b, c, h, w)
    if shift:
        y = torch.roll(y, shifts=(shift, shift), dims=(2, 3))
    x = x + y
    m = conv(group_norm(x, sd, f"{p}.norm2"), sd, f"{p}.mlp.fc1")
    m = conv(F.gelu(m), sd, f"{p}.mlp.fc2")                             # exact-erf GELU (:18)
    return x + m


def basic_layer(x, sd: SD, p: str, cfg: UNetConfig, ctor_res: int):
    """reference models/swin_transformer.py:427-442, patch_size 1, patch_norm False.
    Window / shift are fixed at construction from the level's nominal resolution (:191-194)."""
    win, shift = swin_geometry(cfg, ctor_res)
    x = conv(x, sd, f"{p}.patch_embed.proj")
    for i in range(cfg.swin_depth):
        x = swin_block(x, sd, f"{p}.blocks.{i}", cfg.swin_heads, win, shift if i % 2 else 0)
    return conv(x, sd, f"{p}.patch_unembed.proj")


def feature_extractor(lq, sd: SD, cfg: UNetConfig):
    """reference models/unet.py:689-702."""
    for st in range(cfg.fe_stages):
        lq = F.silu(conv(lq, sd, f"feature_extractor.{3 * st}"))
        lq = conv(lq, sd, f"feature_extractor.{3 * st + 2}.op", stride=2)
    return lq


def _run_block(h, emb, sd: SD, prefix: str, layers, cfg: UNetConfig):
    for j, layer in enumerate(layers):
        kind = layer[0]
        if kind == "conv":
            h = conv(h, sd, f"{prefix}.{j}")
        elif kind == "res":
            h = res_block(h, emb, sd, f"{prefix}.{j}")
        elif kind == "swin":
            h = basic_layer(h, sd, f"{prefix}.{j}", cfg, layer[2])
        elif kind == "down":
            h = conv(h, sd, f"{prefix}.{j}.op", stride=2)                # reference unet.py:99-108
        elif kind == "up":
            h = F.interpolate(h, scale_factor=2, mode="nearest")        # reference unet.py:71-81
            h = conv(h, sd, f"{prefix}.{j}.conv")
    return h


@torch.no_grad()
def unet_forward(sd: SD, cfg: UNetConfig, x, timesteps, lq=None, mask=None, probes: Optional[dict] = None):
    """reference models/unet.py:865-895.  All tensors NCHW fp32, on one device (CPU for the parity tests; bench.py's
    informational `gpu_library_baseline` runs the same functions on CUDA under fp16 autocast)."""
    emb = timestep_embedding(timesteps, cfg.model_channels)
    emb = F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    if lq is not None:
        if mask is not None:
            lq = torch.cat([lq, mask], dim=1)
        lq = feature_extractor(lq.float(), sd, cfg)
        x = torch.cat([x, lq], dim=1)
    input_blocks, middle, output_blocks = unet_block_plan(cfg)
    h = x.float()
    hs = []
    for i, layers in enumerate(input_blocks):
        h = _run_block(h, emb, sd, f"input_blocks.{i}", layers, cfg)
        hs.append(h)
        if probes is not None:
            probes[f"input_blocks.{i}"] = h
    h = _run_block(h, emb, sd, "middle_block", middle, cfg)
    if probes is not None:
        probes["middle_block"] = h
    for i, layers in enumerate(output_blocks):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_block(h, emb, sd, f"output_blocks.{i}", layers, cfg)
        if probes is not None:
            probes[f"output_blocks.{i}"] = h
    return conv(F.silu(group_norm(h, sd, "out.0")), sd, "out.2")
