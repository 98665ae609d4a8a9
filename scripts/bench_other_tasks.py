"""Loop-only throughput of BASELINE configs 4 and 5 (face restoration: batch 8, 512x512 LQ, 8 latent channels, 15 steps;
inpainting: 16 images per GPU, 256x256 LQ + mask, 4 steps) on one GPU: synthetic inputs, random-init weights, CUDA-graph
replay through the same `sample_latent` path bench.py times for config 2.  Prints one JSON line per task."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from resshift_b200.config import preset  # noqa: E402
from resshift_b200.models.script_util import create_gaussian_diffusion  # noqa: E402
from resshift_b200.models.unet import UNetModelSwin  # noqa: E402
from resshift_b200.weights import random_state_dict  # noqa: E402


def run(name, batch, steps, iters=5):
    ucfg, dcfg = preset(name, steps)
    m = UNetModelSwin(**ucfg.to_kwargs())
    m.load_state_dict(random_state_dict(ucfg, 0))
    m = m.cuda().eval()
    diff = create_gaussian_diffusion(**dcfg.to_kwargs())
    g = torch.Generator(device="cuda").manual_seed(12345)
    zy = torch.randn(batch, ucfg.out_channels, 64, 64, device="cuda", generator=g) * 0.5
    lq = torch.rand(batch, 3, ucfg.lq_size, ucfg.lq_size, device="cuda", generator=g) * 2 - 1
    kw = {"lq": lq}
    if ucfg.cond_mask:
        kw["mask"] = (torch.rand(batch, 1, ucfg.lq_size, ucfg.lq_size, device="cuda", generator=g) > 0.5).float() * 2 - 1
    noises = torch.randn(steps + 1, batch, ucfg.out_channels, 64, 64, device="cuda", generator=g)
    for _ in range(3):
        out = diff.sample_latent(zy, m, kw, noises=noises, use_graph=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = diff.sample_latent(zy, m, kw, noises=noises, use_graph=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(json.dumps({"task": name, "batch": batch, "steps": steps, "ms_per_loop": ms, "ms_per_denoise_step": ms / steps,
                      "images_per_s": batch / (ms * 1e-3), "finite": bool(torch.isfinite(out).all()),
                      "launches_per_forward": m.num_launches(batch, 64, 64),
                      "note": "loop only (includes the LQ feature extractor once per loop and the host->graph input copies)"}), flush=True)


if __name__ == "__main__":
    run("faceir", 8, 15)
    run("inpaint", 16, 4)
