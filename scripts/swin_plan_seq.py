"""The failing sequence of test_batch_independence_at_bench_size: one model, batch-16 forwards first, then a batch-1
forward of image 5 — fused Swin attention (tc / mma) against the unfused plan, optionally with poisoned free memory."""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from resshift_b200.config import preset
from resshift_b200.weights import random_state_dict


def run(env, poison=False, first16=True):
    for k in ("RS_SWIN_FUSE", "RS_SWIN_IMPL", "RS_PDL", "RS_NO_REUSE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    from resshift_b200.models.unet import UNetModelSwin
    ucfg, _ = preset("realsr")
    m = UNetModelSwin(**ucfg.to_kwargs())
    m.load_state_dict(random_state_dict(ucfg, 0), strict=True)
    m = m.cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(16, 3, 64, 64, device="cuda", generator=g)
    lq = torch.rand(16, 3, 64, 64, device="cuda", generator=g) * 2 - 1
    t = torch.full((16,), 9, device="cuda")
    full = None
    if first16:
        full = m(x, t, lq=lq).clone()
    if poison:
        junk = torch.full((1 << 28,), 1e4, dtype=torch.float16, device="cuda")   # 512 MB of large values, then freed
        del junk
    one = m(x[5:6], t[5:6], lq=lq[5:6]).clone()
    one2 = m(x[5:6], t[5:6], lq=lq[5:6]).clone()
    del m
    torch.cuda.empty_cache()
    return full, one, bool(torch.equal(one, one2))


ref16, ref1, _ = run({"RS_SWIN_FUSE": "0"})
print(f"unfused: batch-16[5] vs batch-1: max|d|={(ref16[5:6] - ref1).abs().max().item():.3e}")
for name, env, poison, f16 in (("tc", {}, False, True), ("tc poison", {}, True, True), ("tc poison, no batch-16 first", {}, True, False),
                               ("mma", {"RS_SWIN_IMPL": "mma"}, False, True), ("mma poison", {"RS_SWIN_IMPL": "mma"}, True, True),
                               ("tc noreuse", {"RS_NO_REUSE": "1"}, True, True), ("tc nopdl", {"RS_PDL": "0"}, True, True),
                               ("unfused poison", {"RS_SWIN_FUSE": "0"}, True, True)):
    full, one, rep = run(dict(env), poison, f16)
    d1 = (one - ref1).abs()
    msg = f"fused[{name:30s}]: batch-1 vs unfused batch-1 max|d|={d1.max().item():.3e} mean={d1.mean().item():.3e} reproducible={rep}"
    if full is not None:
        msg += f" | batch-16[5] vs unfused batch-16[5] max|d|={(full[5:6] - ref16[5:6]).abs().max().item():.3e}"
    print(msg)
