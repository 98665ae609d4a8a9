"""Plan-level A/B of the fused Swin attention path: the same forward with RS_SWIN_FUSE=0 / 1 (tc) / 1 (mma), block probes
compared in execution order to find the first block that diverges (batch sizes given on the command line)."""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
from resshift_b200.config import preset
from resshift_b200.weights import random_state_dict

os.environ["RS_NO_REUSE"] = "1"
g0 = np.load(ROOT / "tests" / "golden" / "unet_realsr.npz")
blocks = [k.split("/", 1)[1] for k in g0.files if k.startswith("probe_sub/")]


def run(N, env):
    for k in ("RS_SWIN_FUSE", "RS_SWIN_IMPL"):
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
    probes = {b: m.probe(N, 64, 64, b).clone() for b in blocks}
    nl = m.num_launches(N, 64, 64)
    del m
    return out, probes, nl


for N in [int(a) for a in sys.argv[1:]] or [1, 2]:
    ref, pref, nl0 = run(N, {"RS_SWIN_FUSE": "0"})
    for name, env in (("tc", {"RS_SWIN_FUSE": "1"}), ("mma", {"RS_SWIN_FUSE": "1", "RS_SWIN_IMPL": "mma"})):
        out, pr, nl = run(N, env)
        d = (out - ref).abs()
        print(f"N={N} fused[{name}] vs unfused: out max|d|={d.max().item():.3e} mean={d.mean().item():.3e}  launches {nl} vs {nl0}")
        for b in blocks:
            dd = (pr[b] - pref[b]).abs()
            flag = "  <<<<" if dd.max().item() > 5e-2 * max(1.0, pref[b].abs().max().item()) else ""
            print(f"    {b:40s} shape {tuple(pr[b].shape)} max|d|={dd.max().item():.3e} ref absmax {pref[b].abs().max().item():.2f}{flag}")
