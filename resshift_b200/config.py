"""Model / diffusion configuration mirroring the reference's yaml ``params`` trees.

``UNetConfig`` carries exactly the keyword arguments of ``UNetModelSwin.__init__``
(reference models/unet.py:632-657) and ``DiffusionConfig`` those of
``create_gaussian_diffusion`` (reference models/script_util.py:7-21).  The presets
restate the ``model.params`` / ``diffusion.params`` blocks of the shipped yaml files
(reference configs/*.yaml) so that benchmarks and tests do not need the reference tree.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, asdict
from typing import Optional, Sequence, Tuple


@dataclass
class UNetConfig:
    image_size: int = 64
    in_channels: int = 3
    model_channels: int = 160
    out_channels: int = 3
    num_res_blocks: Sequence[int] = (2, 2, 2, 2)
    attention_resolutions: Sequence[int] = (64, 32, 16, 8)
    dropout: float = 0.0
    channel_mult: Sequence[int] = (1, 2, 2, 4)
    conv_resample: bool = True
    dims: int = 2
    use_fp16: bool = False
    num_heads: int = 1
    num_head_channels: int = 32
    use_scale_shift_norm: bool = True
    resblock_updown: bool = False
    swin_depth: int = 2
    swin_embed_dim: int = 192
    window_size: int = 8
    mlp_ratio: float = 4.0
    patch_norm: bool = False
    cond_lq: bool = True
    cond_mask: bool = False
    lq_size: int = 64

    def __post_init__(self):
        if isinstance(self.num_res_blocks, int):
            self.num_res_blocks = (self.num_res_blocks,) * len(self.channel_mult)
        self.num_res_blocks = tuple(int(v) for v in self.num_res_blocks)
        self.channel_mult = tuple(int(v) for v in self.channel_mult)
        self.attention_resolutions = tuple(int(v) for v in self.attention_resolutions)
        # What this implementation covers (every shipped yaml satisfies these).
        assert self.dims == 2 and self.conv_resample and not self.resblock_updown
        assert self.use_scale_shift_norm and not self.patch_norm and self.dropout == 0
        assert len(self.num_res_blocks) == len(self.channel_mult)
        assert self.cond_lq, "the ResShift denoiser is always conditioned on the LQ image"

    # -- derived quantities (reference models/unet.py:689-709) -----------------
    @property
    def swin_heads(self) -> int:
        if self.num_head_channels == -1:
            return self.num_heads
        return self.swin_embed_dim // self.num_head_channels

    @property
    def time_embed_dim(self) -> int:
        return self.model_channels * 4

    @property
    def fe_stages(self) -> int:
        """Number of (conv3x3, SiLU, conv3x3-stride-2) stages in ``feature_extractor``."""
        if self.lq_size == self.image_size:
            return 0
        return int(math.log(self.lq_size / self.image_size) / math.log(2))

    @property
    def lq_in_channels(self) -> int:
        return 4 if self.cond_mask else 3

    @property
    def lq_feat_channels(self) -> int:
        """Channels that get concatenated to x (``base_chn`` in the reference)."""
        if self.lq_size == self.image_size:
            return self.lq_in_channels
        return 16 * (2 ** self.fe_stages)

    def to_kwargs(self) -> dict:
        d = asdict(self)
        d["num_res_blocks"] = list(self.num_res_blocks)
        d["channel_mult"] = list(self.channel_mult)
        d["attention_resolutions"] = list(self.attention_resolutions)
        return d


@dataclass
class DiffusionConfig:
    normalize_input: bool = True
    schedule_name: str = "exponential"
    sf: int = 4
    min_noise_level: float = 0.04
    steps: int = 15
    kappa: float = 2.0
    etas_end: float = 0.99
    schedule_kwargs: dict = field(default_factory=lambda: {"power": 0.3})
    weighted_mse: bool = False
    predict_type: str = "xstart"
    timestep_respacing: Optional[int] = None
    scale_factor: float = 1.0
    latent_flag: bool = True

    def to_kwargs(self) -> dict:
        return asdict(self)


# ---------------------------------------------------------------------------------
# Presets: configs/<name>.yaml  ->  (UNetConfig, DiffusionConfig, latent channels)
# ---------------------------------------------------------------------------------

def preset(name: str, steps: Optional[int] = None) -> Tuple[UNetConfig, DiffusionConfig]:
    if name in ("realsr", "realsr_swinunet_realesrgan256"):
        # configs/realsr_swinunet_realesrgan256.yaml: T=15, min_noise_level 0.04
        u = UNetConfig()
        d = DiffusionConfig()
    elif name in ("realsr_journal", "realsr_swinunet_realesrgan256_journal", "v3", "realsr_v3"):
        # configs/realsr_swinunet_realesrgan256_journal.yaml:38-73 (T=4 native)
        u = UNetConfig()
        d = DiffusionConfig(min_noise_level=0.2, steps=4)
    elif name in ("bicsr", "bicx4_swinunet_lpips"):
        u = UNetConfig()
        d = DiffusionConfig(min_noise_level=0.2, steps=4)
    elif name in ("faceir", "faceir_gfpgan512_lpips"):
        # configs/faceir_gfpgan512_lpips.yaml: f8 VQ (8 latent channels), LQ at 512
        u = UNetConfig(in_channels=8, out_channels=8, lq_size=512)
        d = DiffusionConfig(sf=1, min_noise_level=0.2, steps=4)
    elif name in ("inpaint", "inpaint_imagenet", "inpaint_lama256_imagenet", "inpaint_face", "inpaint_lama256_face"):
        # configs/inpaint_lama256_imagenet.yaml and inpaint_lama256_face.yaml: identical denoiser / schedule; they differ in
        # the VQ-GAN checkpoint only (autoencoder_vq_f4.pth vs celeba256_vq_f4_dim3_face.pth, same f4 architecture)
        u = UNetConfig(cond_mask=True, lq_size=256)
        d = DiffusionConfig(sf=1, min_noise_level=0.2, steps=4)
    elif name in ("realsr_x2", "realsr_realesrgan256_x2"):
        # configs/realsr_realesrgan256_x2.yaml: x2 SR — the LQ image enters at 128x128 through a one-stage feature extractor
        u = UNetConfig(lq_size=128)
        d = DiffusionConfig(sf=2, min_noise_level=0.2, steps=4)
    elif name == "tiny":
        # not a shipped config: a narrow model with the same topology, for fast tests
        u = UNetConfig(model_channels=32, swin_embed_dim=64)
        d = DiffusionConfig(steps=4, min_noise_level=0.2)
    elif name == "tiny_faceir":
        # narrow model with the face-restoration topology: 8 latent channels, three-stage LQ feature extractor at 512
        u = UNetConfig(model_channels=32, swin_embed_dim=64, in_channels=8, out_channels=8, lq_size=512)
        d = DiffusionConfig(sf=1, steps=4, min_noise_level=0.2)
    elif name == "tiny_inpaint":
        u = UNetConfig(model_channels=32, swin_embed_dim=64, cond_mask=True, lq_size=256)
        d = DiffusionConfig(sf=1, steps=4, min_noise_level=0.2)
    else:
        raise KeyError(f"unknown preset {name!r}")
    if steps is not None:
        d.steps = steps
    return u, d


def preset_vq_name(name: str) -> str:
    """VQ-GAN architecture (resshift_b200.vq_arch.vq_preset) each task's yaml names in its `autoencoder:` block."""
    return "f8_face" if name in ("faceir", "faceir_gfpgan512_lpips", "tiny_faceir") else "f4"


# inference_resshift.py --task / --version -> preset (reference inference_resshift.py:15-35,77-126)
TASKS = {
    ("realsr", "v1"): "realsr", ("realsr", "v2"): "realsr", ("realsr", "v3"): "realsr_journal",
    ("bicsr", None): "bicsr", ("inpaint_imagenet", None): "inpaint_imagenet", ("inpaint_face", None): "inpaint_face",
    ("faceir", None): "faceir",
}
