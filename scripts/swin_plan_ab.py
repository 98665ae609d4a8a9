"""Fused vs unfused Swin attention at plan level WITH workspace reuse: final outputs for several batch sizes / switches."""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from resshift_b200.config import preset
from resshift_b200.weights import random_state_dict


def run(N, env):
    for k in ("RS_SWIN_FUSE", "RS_SWIN_IMPL", "RS_PDL", "RS_NO_REUSE", "RS_MLP_FUSE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    from resshift_b200.models.unet import UNetModelSwin
    ucfg, _ = preset("realsr")
    m = UNetModelSwin(**ucfg.to_kwargs())
    m.load_state_dict(random_state_dict(ucfg, 0), strict=True)
    m = m.cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(16, 3, 64, 64, device="cuda", generator=g)[5:5 + N].contiguous()
    lq = (torch.rand(16, 3, 64, 64, device="cuda", generator=g) * 2 - 1)[5:5 + N].contiguous()
    t = torch.full((N,), 9, device="cuda")
    out = m(x, t, lq=lq).clone()
    out2 = m(x, t, lq=lq).clone()
    del m
    return out, bool(torch.equal(out, out2))


for N in [int(a) for a in sys.argv[1:]] or [1, 2, 3]:
    ref, _ = run(N, {"RS_SWIN_FUSE": "0"})
    for name, env in (("tc", {}), ("tc nopdl", {"RS_PDL": "0"}), ("mma", {"RS_SWIN_IMPL": "mma"}), ("tc noreuse", {"RS_NO_REUSE": "1"}),
                      ("tc nomlpfuse", {"RS_MLP_FUSE": "0"})):
        out, rep = run(N, dict(env))
        d = (out - ref).abs()
        print(f"N={N} fused[{name:12s}] vs unfused: max|d|={d.max().item():.3e} mean={d.mean().item():.3e} nan={int(torch.isnan(out).sum())} reproducible={rep}")
