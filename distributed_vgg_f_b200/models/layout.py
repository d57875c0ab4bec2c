"""Native parameter layout and the torch <-> native conversions.

Checkpoints and the oracle use torchvision's layout (conv OIHW, ``classifier.0`` columns in
NCHW-flatten order).  The kernels want:
  * conv weights as ``[Cout][kh][kw][Cin]`` (the K-major B operand of the implicit GEMM, and --
    read MN-major with mirrored taps -- the dgrad operand);
  * the first conv as an im2col GEMM weight ``[Cout][K0]`` with ``k = (kh*3+kw)*3 + c`` zero-padded
    to ``K0 = 64`` (one 128-byte swizzle row);
  * ``classifier.0`` columns permuted to NHWC-flatten order, because activations are NHWC.
All parameters live in ONE flat fp32 arena laid out in gradient-ready order (parallel.buckets);
gradient, Adam moments and the bf16 shadow share the same offsets.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

from .vggf import VGGSpec

CONV0_K = 64


def native_shape(spec: VGGSpec, name: str) -> Tuple[int, ...]:
    base, kind = name.rsplit(".", 1)
    if kind == "bias":
        return spec.param_shape(name)
    for i, c in enumerate(spec.convs):
        if c.name == base:
            return (c.cout, CONV0_K) if i == 0 else (c.cout, 3, 3, c.cin)
    return spec.param_shape(name)


def native_numel(spec: VGGSpec, name: str) -> int:
    return math.prod(native_shape(spec, name))


def to_native(spec: VGGSpec, name: str, t: torch.Tensor) -> torch.Tensor:
    base, kind = name.rsplit(".", 1)
    if kind == "bias":
        return t
    for i, c in enumerate(spec.convs):
        if c.name == base:
            ohwi = t.permute(0, 2, 3, 1).contiguous()
            if i == 0:
                out = torch.zeros(c.cout, CONV0_K, dtype=t.dtype, device=t.device)
                out[:, :27] = ohwi.reshape(c.cout, 27)
                return out
            return ohwi
    if base == spec.fcs[0].name:
        f = spec.fcs[0]
        ch = f.fin // (spec.pooled_hw ** 2)
        return (t.view(f.fout, ch, spec.pooled_hw, spec.pooled_hw).permute(0, 2, 3, 1)
                .reshape(f.fout, f.fin).contiguous())
    return t


def to_torch(spec: VGGSpec, name: str, t: torch.Tensor) -> torch.Tensor:
    base, kind = name.rsplit(".", 1)
    if kind == "bias":
        return t
    for i, c in enumerate(spec.convs):
        if c.name == base:
            if i == 0:
                t = t[:, :27].reshape(c.cout, 3, 3, c.cin)
            return t.permute(0, 3, 1, 2).contiguous()
    if base == spec.fcs[0].name:
        f = spec.fcs[0]
        ch = f.fin // (spec.pooled_hw ** 2)
        return (t.view(f.fout, spec.pooled_hw, spec.pooled_hw, ch).permute(0, 3, 1, 2)
                .reshape(f.fout, f.fin).contiguous())
    return t


def ready_order(spec: VGGSpec):
    """(name, native numel) in the order backward produces the gradients."""
    return [(n, native_numel(spec, n)) for n in reversed(spec.param_names)]


def state_to_native(spec: VGGSpec, state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {n: to_native(spec, n, state[n]) for n in spec.param_names}
