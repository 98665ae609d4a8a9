"""Per-operator time table of one denoiser forward at the benchmark shape (CUDA events around every launch)."""
import ctypes as C
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from resshift_b200 import _lib
from resshift_b200.config import preset
from resshift_b200.models.unet import UNetModelSwin
from resshift_b200.weights import random_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ucfg, _ = preset("realsr")
m = UNetModelSwin(**ucfg.to_kwargs())
m.load_state_dict(random_state_dict(ucfg, 0))
m = m.cuda().eval()
x = torch.randn(B, 3, 64, 64, device="cuda")
lq = torch.rand(B, 3, 64, 64, device="cuda") * 2 - 1
t = torch.full((B,), 7.0, device="cuda")
plan = m.plan(B, 64, 64)
m(x, t, lq=lq)
cap, stride = 1024, 160
ms = (C.c_double * cap)()
desc = C.create_string_buffer(cap * stride)
n = C.c_int32()
for _ in range(2):
    _lib.check(_lib.lib.rs_plan_profile_ops(plan.handle, x.data_ptr(), t.data_ptr(), lq.data_ptr(), None, ms, desc, stride,
                                            cap, C.byref(n), _lib.current_stream()))
rows = [(ms[i] * 1e3, desc.raw[i * stride:(i + 1) * stride].split(b"\0")[0].decode()) for i in range(n.value)]
tot = sum(r[0] for r in rows)
print(f"ops {n.value}  total {tot/1e3:.3f} ms")
agg = {}
for us, d in rows:
    key = " ".join(d.split()[:6]) if d.startswith("conv") else " ".join(d.split()[:3])
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += us
for k, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{us:9.1f} us {us/tot*100:5.1f}%  n={cnt:3d}  avg {us/cnt:7.1f}  {k}")
