"""Python face of the native sm_100a kernels (csrc/*.cu -> distributed_vgg_f_b200/_C*.so).

The extension is built in-tree by ``build_native.py`` (``__graft_entry__.build()``).  On a machine
with a CUDA device the ops FAIL LOUDLY when the extension is missing -- there is no silent torch
fallback on the GPU path; ``ops.ref`` holds the torch reference implementations used as the
numerical oracle in the tests and as the CPU path.

Layout conventions: activations NHWC bf16; conv weights [Cout][3][3][Cin] bf16 (OHWI);
FC weights [out][in] bf16; gradients fp32.
"""
from __future__ import annotations

import importlib
from typing import Optional, Tuple

import torch

from ..config import DATA

_C = None
_IMPORT_ERROR: Optional[BaseException] = None
try:
    _C = importlib.import_module("distributed_vgg_f_b200._C")
except BaseException as e:           # noqa: BLE001 - keep the reason for the loud failure below
    _IMPORT_ERROR = e


def available() -> bool:
    return _C is not None


def require() -> "module":
    """Return the extension module or raise with build instructions."""
    if _C is None:
        raise RuntimeError(
            "distributed_vgg_f_b200._C (the sm_100a kernel library) is not built/importable: %r.\n"
            "Run `python build_native.py` (or __graft_entry__.build()) in the repo root. "
            "The GPU path has no torch fallback by design." % (_IMPORT_ERROR,))
    return _C


def launch_count() -> int:
    return int(require().launch_count())


# ----------------------------------------------------------------------------------- GEMM / conv
def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, M: int, N: int, K: int,
         a_mn: bool = False, b_mn: bool = False, epi: str = "f32_store", ldo: Optional[int] = None,
         bias: Optional[torch.Tensor] = None, alpha: float = 1.0, ksplit: int = 1, bn: int = 0) -> None:
    """out = alpha * A @ B^T on tcgen05.  a_mn/b_mn: the operand is stored [K][M|N]."""
    C = require()
    code = {"f32_store": C.EPI_F32_STORE, "f32_atomic": C.EPI_F32_ATOMIC,
            "f32_atomic_t": C.EPI_F32_ATOMIC_T, "bf16_bias_relu": C.EPI_BF16_BIAS_RELU,
            "f32_store_t": C.EPI_F32_STORE_T, "bf16_store": C.EPI_BF16_STORE}[epi]
    if ldo is None:
        ldo = out.stride(0)
    C.gemm(a, a_mn, b, b_mn, M, N, K, out, ldo, code, bias, alpha, ksplit, bn)


def conv3x3_fprop(x, w, bias, relu: bool = True, bn: int = 0, out=None):
    N, H, W, _ = x.shape
    y = out if out is not None else torch.empty(N, H, W, w.shape[0], dtype=torch.bfloat16, device=x.device)
    require().conv_fprop(x, w, bias, y, relu, bn)
    return y


def pool_mask_like(n: int, h: int, w: int, c: int, device) -> torch.Tensor:
    """Mask buffer of conv3x3_fprop_pool for an un-pooled [n,h,w,c] activation."""
    return torch.zeros(n, h // 2, w // 2, c // 32, 4, dtype=torch.int32, device=device)


def conv3x3_fprop_pool(x, w, bias, bn: int = 0, out=None, mask=None):
    """EXPERIMENTAL (B200_FUSE_POOL=1): maxpool2x2(relu(conv(x) + bias)) without materialising the
    un-pooled tensor.  Returns (pooled bf16 [N,H/2,W/2,Cout], mask int32 [N,H/2,W/2,Cout/32,4])."""
    N, H, W, _ = x.shape
    cout = w.shape[0]
    p = out if out is not None else torch.empty(N, H // 2, W // 2, cout, dtype=torch.bfloat16, device=x.device)
    m = mask if mask is not None else pool_mask_like(N, H, W, cout, x.device)
    require().conv_fprop_pool(x, w, bias, p, m, bn)
    return p, m


def unpool2x2(dp, mask, out=None, colsum=None):
    """Backward of the fused pool: route dp to the recorded argmax where the maximum was > 0."""
    N, OH, OW, Cc = dp.shape
    dz = out if out is not None else torch.empty(N, OH * 2, OW * 2, Cc, dtype=dp.dtype, device=dp.device)
    require().unpool2x2(dp, mask, dz, colsum)
    return dz


def conv3x3_dgrad(dz, w, cin: int, mask_src=None, bn: int = 0, out=None, colsum=None):
    """dx = conv_transpose(dz, w) [* (mask_src > 0)]; colsum (fp32 [cin]) += dx.sum(pixels)."""
    N, H, W, _ = dz.shape
    dx = out if out is not None else torch.empty(N, H, W, cin, dtype=torch.bfloat16, device=dz.device)
    require().conv_dgrad(dz, w, mask_src, dx, colsum, bn)
    return dx


def conv3x3_wgrad(dz, x, dw, scale: float = 1.0, ksplit: int = 0, bn: int = 0):
    """dw (fp32 [Cout][3][3][Cin]) += scale * dz^T (*) x -- accumulates, caller zeroes."""
    require().conv_wgrad(dz, x, dw, scale, ksplit, bn)
    return dw


# ---------------------------------------------------------------------------------- element-wise
def maxpool2x2(x, out=None):
    N, H, W, Cc = x.shape
    y = out if out is not None else torch.empty(N, H // 2, W // 2, Cc, dtype=x.dtype, device=x.device)
    require().maxpool_fwd(x, y)
    return y


def maxpool2x2_relu_bwd(y, dp, out=None, colsum=None):
    dz = out if out is not None else torch.empty_like(y)
    require().maxpool_relu_bwd(y, dp, dz, colsum)
    return dz


def adaptive_avgpool(x, oh: int, ow: int, out=None):
    N, _, _, Cc = x.shape
    y = out if out is not None else torch.empty(N, oh, ow, Cc, dtype=x.dtype, device=x.device)
    require().avgpool_fwd(x, y)
    return y


def adaptive_avgpool_bwd(dy, h: int, w: int, out=None):
    N, _, _, Cc = dy.shape
    dx = out if out is not None else torch.empty(N, h, w, Cc, dtype=dy.dtype, device=dy.device)
    require().avgpool_bwd(dy, dx)
    return dx


def bias_grad(dz2d, db, rows: int, C: int, ld: Optional[int] = None, scale: float = 1.0):
    require().bias_grad(dz2d, db, rows, C, ld if ld is not None else C, scale)
    return db


def fc_bias_act(acc, bias, y=None, y_f32=None, *, B: int, N: int, relu: bool, drop_p: float = 0.0,
                seed: int = 0, offset: int = 0, clear: bool = True):
    require().fc_bias_act(acc, bias, y, y_f32, B, N, relu, drop_p, seed, offset, clear)


def fc_grad_act(acc, act, dz, *, B: int, N: int, relu: bool, drop_p: float = 0.0, clear: bool = True):
    require().fc_grad_act(acc, act, dz, B, N, relu, drop_p, clear)


def cross_entropy(logits, target, dlogits=None, ldd: int = 0, meter=None, loss_out=None,
                  grad_scale: Optional[float] = None, class_weights=None):
    if grad_scale is None:
        grad_scale = 1.0 / logits.shape[0]
    require().cross_entropy(logits, target, dlogits, ldd or logits.shape[1], meter, loss_out,
                            grad_scale, class_weights)


def adam_step(p, m, v, *, g32=None, g16=None, shadow=None, lr, beta1=0.9, beta2=0.999, eps=1e-8,
              weight_decay=0.0, step: int, grad_scale: float = 1.0, zero=None):
    require().adam(p, m, v, g32, g16, shadow, lr, beta1, beta2, eps, weight_decay, step, grad_scale, zero)


def sgd_step(p, mom, *, g32=None, g16=None, shadow=None, lr, momentum=0.9, weight_decay=0.0,
             first: bool, grad_scale: float = 1.0, zero=None):
    require().sgd(p, mom, g32, g16, shadow, lr, momentum, weight_decay, first, grad_scale, zero)


def augment(src_u8, params, out, resized_hw: Tuple[int, int], out_hw: int = DATA.crop,
            mode: str = "im2col", pad: int = 64):
    """uint8 [N,H,W,3] + params [N,8] -> normalised bf16 (NHWC padded, or layer-0 im2col rows)."""
    require().augment(src_u8, params, out, resized_hw[0], resized_hw[1], out_hw, out_hw,
                      1 if mode == "im2col" else 0, pad, list(DATA.mean), list(DATA.std))
    return out


def native_decode_pngs(paths, threads: int = 8):
    """Multi-threaded native PNG decode (csrc/png_decode.cpp) -> uint8 tensor [N,H,W,3], or None when
    a file is not a plain 8-bit PNG / sizes differ / the extension is absent (caller uses PIL)."""
    if _C is None or not hasattr(_C, "decode_pngs"):
        return None
    paths = list(paths)
    if not paths or not all(p.lower().endswith(".png") for p in paths):
        return None
    try:
        return _C.decode_pngs(paths, threads)
    except Exception:
        return None
