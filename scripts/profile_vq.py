"""Per-operator time table of the VQ-GAN bookends (f4: 256x256 <-> 64x64x3) at the benchmark batch (CUDA events around every launch)."""
import ctypes as C
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from resshift_b200 import _lib
from resshift_b200.models.autoencoder import VQModelTorch
from resshift_b200.vq_arch import vq_preset

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg = vq_preset("f4")
m = VQModelTorch(cfg.ddconfig(), cfg.n_embed, cfg.embed_dim).cuda().eval()
img = torch.rand(B, 3, 256, 256, device="cuda") * 2 - 1


def table(plan, what):
    cap, stride = 1024, 160
    ms = (C.c_double * cap)()
    desc = C.create_string_buffer(cap * stride)
    n = C.c_int32()
    for _ in range(2):
        _lib.check(_lib.lib.rs_vq_profile_ops(plan.handle, ms, desc, stride, cap, C.byref(n), _lib.current_stream()))
    rows = [(ms[i] * 1e3, desc.raw[i * stride:(i + 1) * stride].split(b"\0")[0].decode()) for i in range(n.value)]
    tot = sum(r[0] for r in rows)
    print(f"== {what}: ops {n.value}  total {tot/1e3:.3f} ms (batch {B})")
    agg = {}
    for us, d in rows:
        key = " ".join(d.split()[:8]) if d.startswith("conv") else " ".join(d.split()[:3])
        a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += us
    for k, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"{us:9.1f} us {us/tot*100:5.1f}%  n={cnt:3d}  avg {us/cnt:7.1f}  {k}")


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


z = m.encode(img)
print(f"encode {B}x3x256x256: {timed(lambda: m.encode(img)):.3f} ms")
table(m.plan(0, B, 256, 256), "encode")
zz = torch.randn(B, 3, 64, 64, device="cuda")
m.decode(zz)
print(f"decode {B}x3x64x64: {timed(lambda: m.decode(zz)):.3f} ms")
table(m.plan(1, B, 256, 256), "decode")
