"""Input transforms as *parameters + one fused sampling op*.

Reference pipelines (distributedVggf.py:88-95 train, :103-108 val), all on PIL images in the
training process:
  train: RandomResizedCrop(256, scale=(.8,1)) -> RandomRotation(10) -> RandomHorizontalFlip ->
         CenterCrop(224) -> ToTensor -> Normalize(ImageNet mean/std)
  val:   Resize(256) -> CenterCrop(224) -> ToTensor -> Normalize
Every stage is a coordinate map, so the whole chain collapses into "for each of the 224x224 output
pixels: undo centre-crop, undo flip, undo rotation (nearest, zero fill), bilinear-sample the crop
box of the source".  We draw the random parameters on the host with torchvision's exact
distributions (``sample_train_params``) and evaluate the chain in one pass:
``augment_reference`` (torch ops; CPU path and oracle) and ``ops.augment`` (sm_100a kernel that
reads uint8 HWC and writes normalised bf16 NHWC / the layer-0 im2col matrix directly).
``reference_transforms`` returns the literal torchvision pipeline for bit-level parity runs.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from ..config import DATA

# params row: [top, left, crop_h, crop_w, cos(theta), sin(theta), flip, _]
PARAM_DIM = 8


def sample_train_params(n: int, src_h: int, src_w: int,
                        generator: Optional[torch.Generator] = None,
                        scale=DATA.crop_scale, ratio=DATA.crop_ratio,
                        degrees: float = DATA.rotation_deg) -> torch.Tensor:
    """Random crop box / angle / flip per sample, distributed as torchvision's get_params.

    Vectorised over the batch (the loader thread must keep up with ~8k images/s per GPU):
    RandomResizedCrop's "up to 10 attempts, else centre crop" becomes a first-valid-attempt select.
    """
    out = torch.zeros(n, PARAM_DIM, dtype=torch.float32)
    area = float(src_h * src_w)
    log_lo, log_hi = math.log(ratio[0]), math.log(ratio[1])
    u = torch.rand(n, 10, 2, generator=generator, dtype=torch.float64)
    pos = torch.rand(n, 2, generator=generator, dtype=torch.float64)
    misc = torch.rand(n, 2, generator=generator, dtype=torch.float64)
    target_area = area * (scale[0] + (scale[1] - scale[0]) * u[..., 0])
    aspect = torch.exp(log_lo + (log_hi - log_lo) * u[..., 1])
    w = torch.round(torch.sqrt(target_area * aspect))                     # [n, 10]
    h = torch.round(torch.sqrt(target_area / aspect))
    ok = (w > 0) & (w <= src_w) & (h > 0) & (h <= src_h)
    first = torch.argmax(ok.to(torch.int8), dim=1)                        # first valid attempt
    any_ok = ok.any(dim=1)
    rows = torch.arange(n)
    cw, ch = w[rows, first], h[rows, first]
    # torchvision's centre-crop fallback when all 10 attempts fail
    in_ratio = src_w / src_h
    if in_ratio < min(ratio):
        fw, fh = float(src_w), float(round(src_w / min(ratio)))
    elif in_ratio > max(ratio):
        fh, fw = float(src_h), float(round(src_h * max(ratio)))
    else:
        fw, fh = float(src_w), float(src_h)
    cw = torch.where(any_ok, cw, torch.full_like(cw, fw))
    ch = torch.where(any_ok, ch, torch.full_like(ch, fh))
    top = torch.minimum(torch.floor(pos[:, 0] * (src_h - ch + 1)), src_h - ch)
    left = torch.minimum(torch.floor(pos[:, 1] * (src_w - cw + 1)), src_w - cw)
    top = torch.where(any_ok, top, torch.floor((src_h - ch) / 2))
    left = torch.where(any_ok, left, torch.floor((src_w - cw) / 2))
    theta = torch.deg2rad(-degrees + 2.0 * degrees * misc[:, 0])
    out[:, 0], out[:, 1], out[:, 2], out[:, 3] = top.float(), left.float(), ch.float(), cw.float()
    out[:, 4], out[:, 5] = torch.cos(theta).float(), torch.sin(theta).float()
    out[:, 6] = (misc[:, 1] < 0.5).float()
    return out


def val_params(n: int, src_h: int, src_w: int) -> torch.Tensor:
    out = torch.zeros(n, PARAM_DIM, dtype=torch.float32)
    out[:, 2], out[:, 3], out[:, 4] = src_h, src_w, 1.0
    return out


def resized_dims(train: bool, src_h: int, src_w: int, resize: int = DATA.resize) -> Tuple[int, int]:
    """Size of the intermediate 'resized' image: 256x256 for train, shorter-side-256 for val."""
    if train:
        return resize, resize
    if src_h <= src_w:
        return resize, max(int(resize * src_w / src_h), 1)
    return max(int(resize * src_h / src_w), 1), resize


def augment_reference(src: torch.Tensor, params: torch.Tensor, resized_hw: Tuple[int, int],
                      out_hw: int = DATA.crop, mean=DATA.mean, std=DATA.std) -> torch.Tensor:
    """Torch evaluation of the fused transform.  src: uint8 [n,H,W,3] -> float32 [n,3,out,out]."""
    n, H, W, _ = src.shape
    RH, RW = resized_hw
    dev = src.device
    p = params.to(dev, torch.float32)
    ys, xs = torch.meshgrid(torch.arange(out_hw, device=dev, dtype=torch.float32),
                            torch.arange(out_hw, device=dev, dtype=torch.float32), indexing="ij")
    # undo CenterCrop(out_hw) on the RH x RW image (torchvision rounds the offset)
    off_y = int(round((RH - out_hw) / 2.0))
    off_x = int(round((RW - out_hw) / 2.0))
    ay = (ys + off_y).expand(n, -1, -1)
    ax = (xs + off_x).expand(n, -1, -1)
    flip = p[:, 6].view(n, 1, 1) > 0.5
    bx = torch.where(flip, (RW - 1) - ax, ax)
    by = ay
    # undo rotation: PIL affine, output pixel centre -> input coordinate, nearest = floor
    c, s = p[:, 4].view(n, 1, 1), p[:, 5].view(n, 1, 1)
    cx, cy = RW / 2.0, RH / 2.0
    dx, dy = bx + 0.5 - cx, by + 0.5 - cy
    rx = torch.floor(c * dx - s * dy + cx)
    ry = torch.floor(s * dx + c * dy + cy)
    inside = (rx >= 0) & (rx < RW) & (ry >= 0) & (ry < RH)
    # bilinear sample of the crop box (clamped to the box == resize of the cropped image)
    top, left = p[:, 0].view(n, 1, 1), p[:, 1].view(n, 1, 1)
    ch, cw = p[:, 2].view(n, 1, 1), p[:, 3].view(n, 1, 1)
    sy = (ry + 0.5) * (ch / RH) - 0.5
    sx = (rx + 0.5) * (cw / RW) - 0.5
    sy = torch.minimum(torch.clamp(sy, min=0.0), ch - 1)
    sx = torch.minimum(torch.clamp(sx, min=0.0), cw - 1)
    y0, x0 = torch.floor(sy), torch.floor(sx)
    wy, wx = (sy - y0), (sx - x0)
    y1 = torch.minimum(y0 + 1, ch - 1)
    x1 = torch.minimum(x0 + 1, cw - 1)
    srcf = src.to(torch.float32)
    bidx = torch.arange(n, device=dev).view(n, 1, 1).expand(-1, out_hw, out_hw)

    def g(yy, xx):
        return srcf[bidx, (yy + top).long().clamp_(0, H - 1), (xx + left).long().clamp_(0, W - 1)]

    wy, wx = wy.unsqueeze(-1), wx.unsqueeze(-1)
    val = (g(y0, x0) * (1 - wy) * (1 - wx) + g(y0, x1) * (1 - wy) * wx +
           g(y1, x0) * wy * (1 - wx) + g(y1, x1) * wy * wx)
    val = torch.where(inside.unsqueeze(-1), val, torch.zeros_like(val)) / 255.0
    m = torch.tensor(mean, device=dev).view(1, 1, 1, 3)
    sd = torch.tensor(std, device=dev).view(1, 1, 1, 3)
    return ((val - m) / sd).permute(0, 3, 1, 2).contiguous()


def reference_transforms(train: bool):
    """The literal torchvision pipeline of the reference (distributedVggf.py:88-95 / :103-108)."""
    from torchvision import transforms as T

    norm = T.Normalize(list(DATA.mean), list(DATA.std))
    if train:
        return T.Compose([
            T.RandomResizedCrop(size=DATA.resize, scale=DATA.crop_scale),
            T.RandomRotation(degrees=DATA.rotation_deg),
            T.RandomHorizontalFlip(),
            T.CenterCrop(size=DATA.crop),
            T.ToTensor(),
            norm,
        ])
    return T.Compose([T.Resize(size=DATA.resize), T.CenterCrop(size=DATA.crop), T.ToTensor(), norm])


def torchvision_chain_fixed(img_hwc_u8, params_row, resize: int = DATA.resize, crop: int = DATA.crop) -> torch.Tensor:
    """The reference's TRAIN pipeline (distributedVggf.py:88-95) with its random draws pinned to one row
    of ``sample_train_params``: torchvision's own functional ops on a PIL image --
    resized_crop(bilinear) -> rotate(nearest, zero fill) -> hflip -> center_crop -> to_tensor -> normalize.
    Used by the parity tests of ``augment_reference`` and of the sm_100a augment kernel."""
    from PIL import Image
    from torchvision.transforms import InterpolationMode
    from torchvision.transforms import functional as TF

    top, left, ch, cw, c, s, flip = [float(v) for v in params_row[:7]]
    img = Image.fromarray(img_hwc_u8.numpy() if isinstance(img_hwc_u8, torch.Tensor) else img_hwc_u8)
    img = TF.resized_crop(img, int(top), int(left), int(ch), int(cw), [resize, resize], InterpolationMode.BILINEAR)
    img = TF.rotate(img, math.degrees(math.atan2(s, c)), InterpolationMode.NEAREST, fill=0)
    if flip > 0.5:
        img = TF.hflip(img)
    img = TF.center_crop(img, [crop, crop])
    return TF.normalize(TF.to_tensor(img), list(DATA.mean), list(DATA.std))
