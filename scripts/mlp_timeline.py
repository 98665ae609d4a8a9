"""Clock timeline of CTA 0 of the fused MLP kernel (profiling aid)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from resshift_b200 import _lib
from tests import gpu_util as G

N, H, W, E, Hd = 16, 64, 64, 192, 768
if len(sys.argv) > 1:
    N, H, W = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x = torch.randn(N, H, W, E, device="cuda").half()
res = torch.randn(N, H, W, E, device="cuda").half()
w1 = torch.randn(Hd, E, device="cuda") / E ** 0.5
w2 = torch.randn(E, Hd, device="cuda") / Hd ** 0.5
b1 = torch.randn(Hd, device="cuda"); b2 = torch.randn(E, device="cuda")
w1p, _ = G.pack_weight(w1); w2p, _ = G.pack_weight(w2)
out = torch.empty_like(x)
dbg = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
for it in range(3):
    _lib.check(G.L.rs_op_mlp(x.data_ptr(), N, H, W, E, Hd, w1p.data_ptr(), b1.data_ptr(), w2p.data_ptr(), b2.data_ptr(),
                             res.data_ptr(), out.data_ptr(), dbg.data_ptr() if it == 2 else None, G.stream()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    _lib.check(G.L.rs_op_mlp(x.data_ptr(), N, H, W, E, Hd, w1p.data_ptr(), b1.data_ptr(), w2p.data_ptr(), b2.data_ptr(),
                             res.data_ptr(), out.data_ptr(), None, G.stream()))
e1.record(); torch.cuda.synchronize()
print(f"N={N} {H}x{W}: {e0.elapsed_time(e1)/10*1e3:.1f} us per launch")
d = dbg.cpu().view(64, 8)
print("chunk | MMA: GEMM1 start (acc1_empty ok), GEMM1 issued, h_full ok, GEMM2 issued | EPI: acc1_full, h_empty ok, gelu done, signalled   (cycles)")
for j in range(62):
    if int(d[j].abs().sum()):
        print(j, d[j].tolist())
print("final: acc2_full seen", int(d[63, 0]), " stores done", int(d[63, 1]))
