"""Plain-PyTorch fp32 references of the engine's ops, chained exactly like the engine chains its
kernels -- including *where* tensors are rounded to bf16 (activations, activation gradients) --
so that an engine step can be checked layer by layer with tight tolerances.  Also the building
blocks used as oracles in tests/test_kernels_gpu.py.

All functions take/return NCHW fp32 tensors holding bf16-representable values where the engine
stores bf16.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F

from ..models.vggf import VGGSpec


def rb(t: torch.Tensor) -> torch.Tensor:
    """Round to bf16 and back (what storing a tensor as bf16 does)."""
    return t.to(torch.bfloat16).to(torch.float32)


def pool_relu_bwd(y: torch.Tensor, dp: torch.Tensor) -> torch.Tensor:
    """Backward of [ReLU -> maxpool2x2] given the post-ReLU map y: first maximal element wins."""
    with torch.enable_grad():
        yr = y.detach().clone().requires_grad_(True)
        F.max_pool2d(yr, 2, 2).backward(dp)
    return yr.grad * (y > 0)


@torch.no_grad()
def emulated_step(spec: VGGSpec, state: Dict[str, torch.Tensor], x: torch.Tensor, target: torch.Tensor,
                  dropout_masks: bool = False, return_intermediates: bool = False,
                  override: dict = None):
    """Forward + backward of the network on fp32 torch ops with the engine's rounding points.

    state: torch-layout fp32 tensors (weights are rounded to bf16 here, biases stay fp32).
    Returns (logits, mean loss, grads in torch layout).  Dropout is treated as identity.

    ``override`` (optional) pins the forward state to tensors observed elsewhere -- keys ``acts``
    (post-ReLU conv outputs, NCHW), ``feat`` (pooled feature map), ``fc_y`` (hidden FC outputs),
    ``logits``.  ReLU / max-pool gating is discontinuous: a 1-ulp difference in a near-zero
    activation flips a mask and moves a whole gradient row, so a *backward* check is only tight
    when both sides gate on the same forward values.
    """
    override = override or {}
    w = {k: (rb(v) if v.dim() > 1 else v.float()) for k, v in state.items()}
    acts: List[torch.Tensor] = []
    pools: List[torch.Tensor] = []
    a = rb(x)
    conv_in: List[torch.Tensor] = []
    for c in spec.convs:
        conv_in.append(a)
        y = rb(torch.relu(F.conv2d(a, w[c.name + ".weight"], w[c.name + ".bias"], padding=1)))
        if "acts" in override:
            y = override["acts"][len(acts)].float()
        acts.append(y)
        a = y
        if c.pool_after:
            a = F.max_pool2d(y, 2, 2)
        pools.append(a)
    feat_hw = a.shape[-1]
    feat = a
    if feat_hw != spec.pooled_hw:
        feat = rb(F.adaptive_avg_pool2d(a, (spec.pooled_hw, spec.pooled_hw)))
    if "feat" in override:
        feat = override["feat"].float()
    # the engine flattens NHWC; torch-layout FC-1 weights expect NCHW flatten -- same values
    h = torch.flatten(feat, 1)
    fc_in: List[torch.Tensor] = []
    last = len(spec.fcs) - 1
    for i, f in enumerate(spec.fcs):
        fc_in.append(h)
        z = h @ w[f.name + ".weight"].t() + w[f.name + ".bias"]
        if i == last:
            logits = override["logits"].float() if "logits" in override else z
        else:
            h = rb(torch.relu(z) if f.relu else z)
            if "fc_y" in override:
                h = override["fc_y"][i].float()
    B = x.shape[0]
    logp = torch.log_softmax(logits, dim=1)
    loss = -logp.gather(1, target.view(-1, 1)).mean()
    dz = rb((torch.softmax(logits, 1) - F.one_hot(target, logits.shape[1]).float()) / B)

    grads: Dict[str, torch.Tensor] = {}
    inter = {"fc_dz": {}, "fc_y": {i: fc_in[i + 1] for i in range(last)}, "conv_dz": {}}
    for i in range(last, -1, -1):
        f = spec.fcs[i]
        inter["fc_dz"][i] = dz
        grads[f.name + ".bias"] = dz.sum(0)
        grads[f.name + ".weight"] = dz.t() @ fc_in[i]
        dacc = dz @ w[f.name + ".weight"]
        if i > 0:
            prev = spec.fcs[i - 1]
            dz = rb(dacc * (fc_in[i] > 0)) if prev.relu else rb(dacc)
        else:
            g = rb(dacc).view_as(feat)
            inter["dfeat"] = g
    if feat_hw != spec.pooled_hw:
        src = pools[-1].detach().clone().requires_grad_(True)
        with torch.enable_grad():
            F.adaptive_avg_pool2d(src, (spec.pooled_hw, spec.pooled_hw)).backward(g)
        g = rb(src.grad)
    for i in range(len(spec.convs) - 1, -1, -1):
        c = spec.convs[i]
        y = acts[i]
        if c.pool_after:
            with torch.enable_grad():
                dzc = rb(pool_relu_bwd(y, g))
        else:
            dzc = g
        inter["conv_dz"][i] = dzc
        grads[c.name + ".bias"] = dzc.sum((0, 2, 3))
        xin = conv_in[i]
        grads[c.name + ".weight"] = torch.nn.grad.conv2d_weight(xin, w[c.name + ".weight"].shape, dzc, padding=1)
        if i == 0:
            break
        dx = F.conv_transpose2d(dzc, w[c.name + ".weight"], padding=1)
        if spec.convs[i - 1].pool_after:
            g = rb(dx)
        else:
            g = rb(dx * (xin > 0))
    if return_intermediates:
        return logits, loss, grads, inter
    return logits, loss, grads
