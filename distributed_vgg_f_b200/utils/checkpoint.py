"""Checkpoint save / resume.

The reference has no checkpointing at all (no torch.save / load / state_dict anywhere; SURVEY D1,
5.4).  The only layout that is "the same" as anything the reference could produce is the
``state_dict()`` of its wrapped model: both ``DistributedDataParallel`` (ws > 1,
distributedVggf.py:225) and ``DataParallel`` (ws == 1, :227) prefix every key with ``module.``, and
the tensors carry torchvision's VGG names and shapes plus the funnel
(``module.features.0.weight`` ... ``module.classifier.6.3.bias``, 34 tensors).  We therefore write

    {"model": {module.<torchvision name>: fp32 tensor in torch layout},
     "optimizer": {"name", "step", "exp_avg": {...}, "exp_avg_sq": {...} | "momentum_buffer": {...}},
     "epoch": int, "args": dict, "format": "distributed-vgg-f_b200/1"}

with ``torch.save`` on rank 0 only.  ``strip_module_prefix(ckpt["model"])`` loads directly into the
reference's *unwrapped* model object (``vgg_funnel_model(C).load_state_dict``).  The native engine
stores conv weights as OHWI and FC-1 columns in NHWC-flatten order; conversion to / from the torch
layout happens here (``engine.export_state`` / ``import_state``).
"""
from __future__ import annotations

import os
from typing import Any, Dict, Optional

import torch

FORMAT = "distributed-vgg-f_b200/1"


def add_module_prefix(state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {(k if k.startswith("module.") else "module." + k): v for k, v in state.items()}


def strip_module_prefix(state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {(k[len("module."):] if k.startswith("module.") else k): v for k, v in state.items()}


def model_state(model) -> Dict[str, torch.Tensor]:
    if hasattr(model, "export_state"):              # NativeEngine
        state = model.export_state()
    else:
        state = model.state_dict()
    return add_module_prefix({k: v.detach().to("cpu", torch.float32) for k, v in state.items()})


def optimizer_state(model, optimizer) -> Dict[str, Any]:
    if hasattr(model, "export_optimizer_state"):
        return model.export_optimizer_state()
    return optimizer.state_dict() if optimizer is not None else {}


def save_checkpoint(path: str, model, optimizer, epoch: int, args: Optional[dict] = None,
                    is_rank0: bool = True) -> None:
    if hasattr(model, "prepare_export"):     # collective: engines with sharded optimizer state gather it
        model.prepare_export()
    if not is_rank0:
        return
    payload = {"format": FORMAT, "model": model_state(model),
               "optimizer": optimizer_state(model, optimizer), "epoch": int(epoch),
               "args": dict(args or {})}
    tmp = path + ".tmp"
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(payload, tmp)
    os.replace(tmp, path)


def load_checkpoint(path: str, model, optimizer=None) -> int:
    """Restore model (+ optimizer) and return the epoch to resume *after*."""
    # our files hold tensors and plain containers only: no need to unpickle arbitrary objects
    payload = torch.load(path, map_location="cpu", weights_only=True)
    state = payload["model"] if "model" in payload else payload
    if hasattr(model, "import_state"):
        model.import_state(strip_module_prefix(state))
        if "optimizer" in payload and payload["optimizer"]:
            model.import_optimizer_state(payload["optimizer"])
    else:
        target = model.module if hasattr(model, "module") and not hasattr(model, "features") else model
        target.load_state_dict(strip_module_prefix(state))
        if optimizer is not None and payload.get("optimizer"):
            optimizer.load_state_dict(payload["optimizer"])
    return int(payload.get("epoch", 0))
