"""SUPERSEDED by scripts/conv_sweep.py (this one times from the host, one launch at a time, and is launch-bound for
the small layers it was meant to study; kept because profiles/r1_s12_smallm_sweep.log was produced by it).

Sweep tile configurations for the low-resolution (few-tile) conv layers: time per launch via CUDA events."""
import ctypes as C
import itertools
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from resshift_b200 import _lib
from tests import gpu_util as G
L = _lib.lib


def run(shape, env, iters=30):
    N, H, W, Ci, Co, k = shape
    for kk in ("RS_CONV_BN", "RS_CONV_CG", "RS_CONV_SPLITK", "RS_CONV_STAGES", "RS_CONV_OCC"):
        os.environ.pop(kk, None)
    os.environ.update({k2: str(v) for k2, v in env.items()})
    x = torch.randn(N, H, W, Ci, device="cuda").half()
    w = torch.randn(Co, Ci, k, k, device="cuda") / (Ci * k * k) ** 0.5
    b = torch.randn(Co, device="cuda")
    res = torch.randn(N, H, W, Co, device="cuda").half()
    wp, ipad = G.pack_weight(w)
    out = torch.empty(N, H, W, Co, dtype=torch.float16, device="cuda")
    part = torch.empty(N * 64 * Co * 2, dtype=torch.float32, device="cuda")
    scratch = torch.empty(8 * N * H * W * Co, dtype=torch.float32, device="cuda")
    S = C.c_int32()
    def call():
        _lib.check(L.rs_op_conv2d_splitk(x.data_ptr(), N, H, W, Ci, Ci, wp.data_ptr(), ipad, b.data_ptr(), Co, k, 1,
                                         res.data_ptr(), Co, out.data_ptr(), Co, 0, part.data_ptr(), Co, 0,
                                         scratch.data_ptr(), C.byref(S), _lib.current_stream()))
    # thrash L2 between launches with a big copy so that weights come from HBM like in the model
    junk_a = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    junk_b = torch.empty_like(junk_a)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    tot = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(iters):
        junk_b.copy_(junk_a)
        e0.record(); call(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    o = (C.c_int32 * 8)()
    print(f"  {str(env):70s} S={S.value}  {tot/iters*1e3:7.1f} us (cold L2)")


if __name__ == "__main__":
    shapes = [(16, 8, 8, 640, 640, 3), (16, 16, 16, 320, 320, 3), (16, 8, 8, 1280, 640, 3)]
    cfgs = [
        {},
        {"RS_CONV_BN": 64, "RS_CONV_CG": 1, "RS_CONV_SPLITK": 1},
        {"RS_CONV_BN": 64, "RS_CONV_CG": 2, "RS_CONV_SPLITK": 1},
        {"RS_CONV_BN": 32, "RS_CONV_CG": 2, "RS_CONV_SPLITK": 1},
        {"RS_CONV_BN": 160, "RS_CONV_CG": 2, "RS_CONV_SPLITK": 1},
        {"RS_CONV_BN": 160, "RS_CONV_CG": 2, "RS_CONV_SPLITK": 2},
        {"RS_CONV_BN": 160, "RS_CONV_CG": 2, "RS_CONV_SPLITK": 4},
        {"RS_CONV_BN": 160, "RS_CONV_CG": 1, "RS_CONV_SPLITK": 4},
        {"RS_CONV_BN": 80, "RS_CONV_CG": 2, "RS_CONV_SPLITK": 2},
        {"RS_CONV_BN": 80, "RS_CONV_CG": 2, "RS_CONV_SPLITK": 4},
        {"RS_CONV_BN": 64, "RS_CONV_CG": 2, "RS_CONV_SPLITK": 2},
        {"RS_CONV_BN": 128, "RS_CONV_CG": 2, "RS_CONV_SPLITK": 3},
    ]
    for sh in shapes:
        print("shape", sh)
        for c in cfgs:
            try:
                run(sh, c)
            except Exception as e:
                print("  ", c, "ERR", str(e)[:100])
