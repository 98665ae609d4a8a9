"""Profiling workload: N plain (un-graphed) denoiser forwards at the benchmark shape (batch 16, 64x64 latent,
full-width realsr model), for `ncu` launch lists / `--set full` captures.  Not a benchmark."""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from resshift_b200.config import preset  # noqa: E402
from resshift_b200.models.unet import UNetModelSwin  # noqa: E402
from resshift_b200.weights import random_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--iters", type=int, default=2)
args = ap.parse_args()
ucfg, _ = preset("realsr")
m = UNetModelSwin(**ucfg.to_kwargs())
m.load_state_dict(random_state_dict(ucfg, 0))
m = m.cuda().eval()
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(args.batch, 3, 64, 64, device="cuda", generator=g)
lq = torch.rand(args.batch, 3, 64, 64, device="cuda", generator=g) * 2 - 1
t = torch.full((args.batch,), 7, device="cuda")
print("launches per forward:", m.num_launches(args.batch, 64, 64))
out = m(x, t, lq=lq)                       # warm-up (weight packing, plan, first touch) outside the profiled range
torch.cuda.synchronize()
torch.cuda.profiler.start()                # ncu --profile-from-start off
for _ in range(args.iters):
    out = m(x, t, lq=lq)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(out.abs().mean()))
