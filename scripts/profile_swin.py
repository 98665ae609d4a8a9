"""Profiling workload for the fused Swin attention kernel alone (ncu --set full --import-source on): batch 16, 64x64 and 8x8."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from resshift_b200 import _lib
from tests import gpu_util as G

def run(N, H, W, E=192, shift=4, iters=3):
    heads = E // 32
    g = torch.Generator(device="cuda").manual_seed(1)
    x = (torch.randn(N, H, W, E, device="cuda", generator=g)).half()
    gamma = torch.ones(E, device="cuda"); beta = torch.zeros(E, device="cuda")
    wqkv = torch.randn(3 * E, E, device="cuda", generator=g) / E ** 0.5
    bqkv = torch.zeros(3 * E, device="cuda")
    wproj = torch.randn(E, E, device="cuda", generator=g) / E ** 0.5
    bproj = torch.zeros(E, device="cuda")
    dense = torch.zeros(heads * 64 * 64, dtype=torch.float32, device="cuda")
    rows = 128 if H * W >= 128 else 64
    slots = H * W // rows
    xs = x.float().reshape(N, slots, rows, E)
    mean_s = xs.mean(dim=2)
    part = torch.stack([mean_s, ((xs - mean_s[:, :, None]) ** 2).sum(dim=2)], dim=-1).contiguous()
    wq_p, _ = G.pack_weight(wqkv); wp_p, _ = G.pack_weight(wproj)
    y = torch.empty_like(x)
    pout = torch.empty(N, (H // 8) * (W // 8), E, 2, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(iters):
        if i == iters - 1:
            e0.record()
        _lib.check(G.L.rs_op_swin_attn(x.data_ptr(), N, H, W, E, heads, shift, part.data_ptr(), slots, gamma.data_ptr(), beta.data_ptr(),
                                       wq_p.data_ptr(), bqkv.data_ptr(), dense.data_ptr(), wp_p.data_ptr(), bproj.data_ptr(),
                                       y.data_ptr(), pout.data_ptr(), None, None, G.stream()))
    e1.record()
    torch.cuda.synchronize()
    print(f"swin_attn N={N} {H}x{W} E={E}: {e0.elapsed_time(e1) * 1e3:.1f} us")

run(16, 64, 64)
run(16, 8, 8, shift=0)
run(16, 32, 32)
