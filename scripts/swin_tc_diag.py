"""Diagnostics for the tcgen05 Swin attention kernel (swin_attn_tc.cuh): the fused op against plain torch for weight
choices that isolate one stage each (projection = identity, attention = identity / uniform), with the error broken down
by head, window parity inside the CTA's pair and token-row quadrant; plus timings of both implementations and the
in-kernel timeline of CTA 0's first tile."""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import torch.nn.functional as F
from resshift_b200 import _lib
from resshift_b200.arch import relative_position_index, shifted_window_mask
from tests import gpu_util as G


def reference(x, gamma, beta, wqkv, bqkv, table, wproj, bproj, shift):
    N, H, W, E = x.shape
    heads = E // 32
    xc = x.float()
    xn = F.group_norm(xc.permute(0, 3, 1, 2), 32, gamma, beta, eps=1e-5).half().float()
    qkv = F.conv2d(xn, wqkv.half().float()[:, :, None, None], bqkv).half().float()
    if shift:
        qkv = torch.roll(qkv, (-shift, -shift), (2, 3))
    yw = qkv.reshape(N, 3 * E, H // 8, 8, W // 8, 8).permute(0, 2, 4, 3, 5, 1).reshape(-1, 64, 3, heads, 32)
    qq, kk, vv = (yw[:, :, i].transpose(1, 2) for i in range(3))
    attn = (qq * 32 ** -0.5) @ kk.transpose(-2, -1)
    idx = relative_position_index(8).reshape(-1).to(x.device)
    attn = attn + table[idx].view(64, 64, heads).permute(2, 0, 1)[None]
    if shift:
        m = shifted_window_mask(H, W, 8, shift).to(x.device)
        attn = (attn.view(-1, m.shape[0], heads, 64, 64) + m[None, :, None]).view(-1, heads, 64, 64)
    o = (attn.softmax(-1) @ vv).transpose(1, 2).reshape(-1, 64, E)
    o = o.view(N, H // 8, W // 8, 8, 8, E).permute(0, 5, 1, 3, 2, 4).reshape(N, E, H, W)
    if shift:
        o = torch.roll(o, (shift, shift), (2, 3))
    o = o.half().float()
    return (F.conv2d(o, wproj.half().float()[:, :, None, None], bproj) + xc.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)


def run_case(name, N, H, W, E, shift, mode, impl="tc"):
    os.environ["RS_SWIN_IMPL"] = impl
    heads = E // 32
    g = torch.Generator(device="cuda").manual_seed(E + H * 3 + shift + N)
    x = (torch.randn(N, H, W, E, device="cuda", generator=g) * 1.5 + 0.3).half()
    gamma = 1 + 0.2 * torch.randn(E, device="cuda", generator=g)
    beta = 0.2 * torch.randn(E, device="cuda", generator=g)
    wqkv = torch.randn(3 * E, E, device="cuda", generator=g) / E ** 0.5
    bqkv = torch.randn(3 * E, device="cuda", generator=g) * 0.1
    wproj = torch.randn(E, E, device="cuda", generator=g) / E ** 0.5 * 0.5
    bproj = torch.randn(E, device="cuda", generator=g) * 0.1
    table = torch.randn(225, heads, device="cuda", generator=g) * 0.5
    if mode in ("proj_id", "attn_id", "attn_uniform"):
        wproj = torch.eye(E, device="cuda"); bproj = torch.zeros(E, device="cuda")
    if mode == "attn_id":            # P = one-hot on the query itself: O = V
        table = torch.zeros(225, heads, device="cuda"); table[112] = 60.0     # relative offset (0, 0)
        wqkv[:2 * E] = 0; bqkv[:2 * E] = 0
    if mode == "attn_uniform":       # S = 0: O = mean of V over the window (no shift mask in this mode)
        table = torch.zeros(225, heads, device="cuda")
        wqkv[:2 * E] = 0; bqkv[:2 * E] = 0
    dense = torch.empty(heads * 64 * 64, dtype=torch.float32, device="cuda")
    _lib.check(G.L.rs_op_expand_relpos(table.data_ptr(), dense.data_ptr(), heads, G.stream()))
    rows = 128 if H * W >= 128 else 64
    slots = H * W // rows
    xs = x.float().reshape(N, slots, rows, E)
    mean_s = xs.mean(dim=2)
    part = torch.stack([mean_s, ((xs - mean_s[:, :, None]) ** 2).sum(dim=2)], dim=-1).contiguous()
    wq_p, _ = G.pack_weight(wqkv); wp_p, _ = G.pack_weight(wproj)
    nW = (H // 8) * (W // 8)
    y = torch.full_like(x, float("nan"))
    pout = torch.full((N, nW, E, 2), float("nan"), dtype=torch.float32, device="cuda")
    _lib.check(G.L.rs_op_swin_attn(x.data_ptr(), N, H, W, E, heads, shift, part.data_ptr(), slots, gamma.data_ptr(), beta.data_ptr(),
                                   wq_p.data_ptr(), bqkv.data_ptr(), dense.data_ptr(), wp_p.data_ptr(), bproj.data_ptr(),
                                   y.data_ptr(), pout.data_ptr(), None, None, G.stream()))
    torch.cuda.synchronize()
    ref = reference(x, gamma, beta, wqkv, bqkv, table, wproj, bproj, shift)
    d = (y.float() - ref).abs()
    nan = int(torch.isnan(y).sum())
    d = torch.nan_to_num(d, nan=1e3)
    print(f"[{name} impl={impl} mode={mode}] N={N} {H}x{W} E={E} shift={shift}: max={d.max():.3e} mean={d.mean():.3e} nan={nan} ref_max={ref.abs().max():.2f}")
    # windows of the shifted partition
    dd = d.permute(0, 3, 1, 2)
    if shift:
        dd = torch.roll(dd, (-shift, -shift), (2, 3))
    dw = dd.reshape(N, E, H // 8, 8, W // 8, 8).permute(0, 2, 4, 1, 3, 5).reshape(N * nW, E, 64)   # [window, ch, token]
    print("   by head      :", " ".join(f"{dw[:, 32 * h:32 * h + 32].max():.2e}" for h in range(heads)))
    print("   by win parity:", " ".join(f"{dw[k::2].max():.2e}" for k in range(2)))
    print("   by token quad:", " ".join(f"{dw[:, :, 16 * q:16 * q + 16].max():.2e}" for q in range(4)))
    print("   by window    :", " ".join(f"{dw[w].max():.1e}" for w in range(min(N * nW, 16))))
    return float(d.max())


def timing(N, H, W, E=192, shift=4, impl="tc", iters=5, timeline=False):
    os.environ["RS_SWIN_IMPL"] = impl
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
    tl = torch.zeros(128, dtype=torch.int64, device="cuda")
    if timeline:
        _lib.check(G.L.rs_debug_swin_timeline(tl.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(iters):
        if i == iters - 1:
            e0.record()
        _lib.check(G.L.rs_op_swin_attn(x.data_ptr(), N, H, W, E, heads, shift, part.data_ptr(), slots, gamma.data_ptr(), beta.data_ptr(),
                                       wq_p.data_ptr(), bqkv.data_ptr(), dense.data_ptr(), wp_p.data_ptr(), bproj.data_ptr(),
                                       y.data_ptr(), pout.data_ptr(), None, None, G.stream()))
    e1.record()
    torch.cuda.synchronize()
    _lib.check(G.L.rs_debug_swin_timeline(None))
    print(f"time impl={impl} N={N} {H}x{W} E={E} shift={shift}: {e0.elapsed_time(e1) * 1e3:.1f} us")
    if timeline:
        t = tl.cpu().tolist()
        for k in range(2):
            print(f"   tile {k} workers [0..26]:", t[64 * k:64 * k + 27])
            print(f"   tile {k} mma    [32..47]:", t[64 * k + 32:64 * k + 48])


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "diag"):
        for mode in ("attn_uniform", "attn_id", "proj_id", "full"):
            run_case("a", 2, 16, 16, 192, 0, mode)
        run_case("b", 2, 16, 32, 192, 4, "full")
        run_case("c", 3, 8, 8, 192, 0, "full")
        run_case("d", 2, 16, 16, 64, 0, "full")
        run_case("e", 1, 64, 64, 192, 4, "full")
    if what == "ncu":
        timing(16, 64, 64, shift=4, impl="tc", iters=2)
    if what in ("all", "time"):
        for (H, sh) in ((64, 4), (32, 4), (16, 4), (8, 0)):
            timing(16, H, H, shift=sh, impl="tc", timeline=(H == 64))
            timing(16, H, H, shift=sh, impl="mma")
