"""Diagnostics for the tcgen05 conv/GEMM kernel: structured inputs that reveal descriptor / swizzle /
layout mistakes.  Each experiment runs in its own process (a trapped kernel kills the CUDA context)."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def experiment(name):
    import torch
    from tests import gpu_util as G
    torch.manual_seed(0)
    dev = "cuda"
    if name == "identity_1x1":
        # out = x @ I : any permutation of channels / rows shows up directly
        N, H, W, Cc = 1, 8, 16, 64
        x = torch.arange(N * H * W * Cc, device=dev, dtype=torch.float32).reshape(N, H, W, Cc) % 251 / 16.0
        x = x.half()
        w = torch.eye(Cc, device=dev)[:, :, None, None]
        out = G.conv2d(x, w, None)
        d = (out.float() - x.float()).abs()
        print(name, "max|d|", d.max().item(), "nan", int(torch.isnan(out.float()).sum()))
        if d.max().item() > 0:
            bad = (d > 0).nonzero()
            print("first mismatches (n,h,w,c):", bad[:8].tolist())
            print("row 0 got :", out[0, 0, 0, :16].tolist())
            print("row 0 want:", x[0, 0, 0, :16].tolist())
            print("row 9 got :", out[0, 0, 9, :16].tolist())
            print("row 9 want:", x[0, 0, 9, :16].tolist())
            rows_ok = (d.reshape(-1, Cc).max(dim=1).values == 0).float().mean().item()
            cols_ok = (d.reshape(-1, Cc).max(dim=0).values == 0).float().mean().item()
            print("fraction of exact rows", rows_ok, "exact cols", cols_ok)
    elif name == "k128_1x1":
        N, H, W, Cc, Co = 1, 8, 16, 128, 64
        x = torch.randn(N, H, W, Cc, device=dev).half()
        w = torch.randn(Co, Cc, 1, 1, device=dev) / Cc ** 0.5
        out = G.nchw32(G.conv2d(x, w, None))
        ref = G.ref_conv(x, w, None)
        print(name, G.err_stats(out, ref))
    elif name == "shift_3x3":
        # weight = delta at tap (ky,kx): output must be the shifted input
        N, H, W, Cc = 1, 16, 16, 64
        x = torch.randn(N, H, W, Cc, device=dev).half()
        for ky in range(3):
            for kx in range(3):
                w = torch.zeros(Cc, Cc, 3, 3, device=dev)
                w[:, :, ky, kx] = torch.eye(Cc, device=dev)
                out = G.nchw32(G.conv2d(x, w, None))
                ref = G.ref_conv(x, w, None)
                print(name, (ky, kx), "max|d|", (out - ref).abs().max().item())
    elif name == "big":
        N, H, W, Cc, Co = 16, 64, 64, 160, 160
        x = torch.randn(N, H, W, Cc, device=dev).half()
        w = torch.randn(Co, Cc, 3, 3, device=dev) / (9 * Cc) ** 0.5
        b = torch.randn(Co, device=dev)
        import time
        for bn in (0,):
            out = G.conv2d(x, w, b, bn=bn)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                G.conv2d(x, w, b, bn=bn)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 5
            fl = 2.0 * N * H * W * Co * Cc * 9
            print(name, f"bn={bn} {dt*1e3:.3f} ms (incl. weight pack + alloc) {fl/dt/1e12:.1f} TFLOP/s")
        ref = G.ref_conv(x[:1], w, b)
        print(name, G.err_stats(G.nchw32(out[:1]), ref))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        experiment(sys.argv[1])
    else:
        for name in ("identity_1x1", "k128_1x1", "shift_3x3", "big"):
            r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=300)
            print(f"==== {name} (exit {r.returncode})")
            print(r.stdout[-3000:])
            if r.returncode != 0:
                print(r.stderr[-2000:])
