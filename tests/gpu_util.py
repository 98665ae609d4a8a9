"""Helpers shared by the GPU parity tests (thin wrappers over the C ABI single-operator entry points)."""
import ctypes as C

import torch
import torch.nn.functional as F

from resshift_b200 import _lib

L = _lib.lib


def stream():
    return _lib.current_stream()


def nhwc16(x_nchw: torch.Tensor) -> torch.Tensor:
    return x_nchw.permute(0, 2, 3, 1).contiguous().half()


def nchw32(x_nhwc: torch.Tensor) -> torch.Tensor:
    return x_nhwc.permute(0, 3, 1, 2).float().contiguous()


def pack_weight(w: torch.Tensor) -> tuple:
    """fp32 [O, I, kh, kw] (or [O, I]) on GPU -> packed fp16 [O][kh*kw][Ipad]."""
    if w.dim() == 2:
        w = w[:, :, None, None]
    O, I, KH, KW = w.shape
    ipad = (I + 7) // 8 * 8
    dst = torch.empty(O * KH * KW * ipad, dtype=torch.float16, device=w.device)
    _lib.check(L.rs_op_pack_conv_weight(w.contiguous().data_ptr(), dst.data_ptr(), O, I, KH, KW, ipad, stream()))
    return dst, ipad


def conv2d(x_nhwc, w, bias, stride=1, residual=None, act=0, out_f32=False, bn=0, in_view=None, out_view=None):
    """x_nhwc fp16 [N,H,W,C] (or a channel-slice view of a wider buffer given as (buffer, c0, C))."""
    if in_view is not None:
        buf, c0, Cc = in_view
        N, H, W, ld = buf.shape
        xptr = buf.data_ptr() + 2 * c0
    else:
        N, H, W, Cc = x_nhwc.shape
        ld = Cc
        xptr = x_nhwc.data_ptr()
    wp, ipad = pack_weight(w)
    O = w.shape[0]
    k = w.shape[-1] if w.dim() == 4 else 1
    Ho, Wo = H // stride, W // stride
    dev = w.device
    out = None
    out_ptr, out_ld = None, 0
    if out_view is not None:
        obuf, oc0 = out_view
        out_ptr, out_ld = obuf.data_ptr() + 2 * oc0, obuf.shape[-1]
    elif not out_f32:
        out = torch.full((N, Ho, Wo, O), float("nan"), dtype=torch.float16, device=dev)
        out_ptr, out_ld = out.data_ptr(), O
    o32 = torch.full((N, O, Ho, Wo), float("nan"), dtype=torch.float32, device=dev) if out_f32 else None
    _lib.check(L.rs_op_conv2d(xptr, N, H, W, Cc, ld, wp.data_ptr(), ipad, _lib.ptr(bias), O, k, stride,
                              _lib.ptr(residual), 0 if residual is None else residual.shape[-1], out_ptr, out_ld,
                              _lib.ptr(o32), act, bn, stream()))
    torch.cuda.synchronize()
    return o32 if out_f32 else out


def ref_conv(x_nhwc16, w, bias, stride=1, residual=None, act=0):
    """fp32 torch reference on the fp16-rounded operands."""
    x = nchw32(x_nhwc16)
    wq = w.half().float()
    if wq.dim() == 2:
        wq = wq[:, :, None, None]
    y = F.conv2d(x, wq, bias, stride=stride, padding=wq.shape[-1] // 2)
    if act == 1:
        y = F.gelu(y)
    elif act == 2:
        y = F.silu(y)
    if residual is not None:
        y = y + nchw32(residual)
    return y


def err_stats(got: torch.Tensor, ref: torch.Tensor) -> dict:
    d = (got.float() - ref.float()).abs()
    return {"max_abs": d.max().item(), "mean_abs": d.mean().item(), "ref_absmax": ref.abs().max().item(),
            "ref_std": ref.float().std().item(), "nan": int(torch.isnan(got.float()).sum().item())}
