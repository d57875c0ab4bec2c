"""Classify images with a checkpoint written by ``--save``.

The reference stops at the per-epoch validation line; this is the missing last step for a user who
trained a model: load the checkpoint (its ``args`` record the architecture), apply the reference's
validation transform (Resize 256 -> CenterCrop 224 -> Normalize, distributedVggf.py:103-108) and
print one line per image::

    python -m distributed_vgg_f_b200.tools.predict ck.pt --classes edible,other,toy img1.png img2.png
    python -m distributed_vgg_f_b200.tools.predict ck.pt --data-root /data/coil3 /data/coil3/ValidationData/toy

``--classes`` (or ``--data-root``, whose TrainData sub-folders define them) only names the outputs;
directories are expanded to the image files inside.  Runs on the GPU through the native engine when
one is present (``--engine auto``), else on the CPU oracle.
"""
from __future__ import annotations

import argparse
import os
from typing import List, Optional, Sequence

import torch

from ..config import DATA
from ..data.folder import IMG_EXTENSIONS, find_classes
from ..data.transforms import reference_transforms
from ..models.vggf import build_oracle, get_spec
from ..utils.checkpoint import strip_module_prefix


def expand(paths: Sequence[str]) -> List[str]:
    out: List[str] = []
    for p in paths:
        if os.path.isdir(p):
            for dp, _, files in sorted(os.walk(p)):
                out += [os.path.join(dp, f) for f in sorted(files) if f.lower().endswith(IMG_EXTENSIONS)]
        else:
            out.append(p)
    return out


def load_model(checkpoint: str, model: Optional[str] = None, num_classes: Optional[int] = None):
    payload = torch.load(checkpoint, map_location="cpu", weights_only=True)
    state = strip_module_prefix(payload["model"] if "model" in payload else payload)
    saved = payload.get("args", {}) if isinstance(payload, dict) else {}
    last = [k for k in state if k.endswith(".weight")][-1]
    classes = num_classes or int(state[last].shape[0])
    spec = get_spec(model or saved.get("model") or "vggf", classes)
    return spec, state


@torch.no_grad()
def predict(checkpoint: str, images: Sequence[str], model: Optional[str] = None, engine: str = "auto",
            batch: int = 32):
    """Returns (logits [n, C] fp32 on the CPU, files)."""
    from PIL import Image

    spec, state = load_model(checkpoint, model)
    files = expand(images)
    tf = reference_transforms(train=False)
    use_native = engine == "native" or (engine == "auto" and torch.cuda.is_available()
                                        and torch.cuda.get_device_capability()[0] >= 10)
    if use_native:
        from ..engine.native_engine import NativeEngine
        net = NativeEngine(spec, device=torch.device("cuda", 0), batch=batch, init_state=state, distributed=False)
    else:
        net = build_oracle(spec, seed=0).eval()
        net.load_state_dict(state)
    out = []
    for i in range(0, len(files), batch):
        x = torch.stack([tf(Image.open(f).convert("RGB")) for f in files[i:i + batch]])
        if use_native:
            out.append(net.forward_logits((x, torch.zeros(x.shape[0], dtype=torch.int64))).float().cpu())
        else:
            out.append(net(x).float())
    return (torch.cat(out) if out else torch.zeros(0, spec.num_classes)), files


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("checkpoint")
    ap.add_argument("images", nargs="+", help="image files and/or directories")
    ap.add_argument("--classes", default=None, help="comma-separated class names in label order")
    ap.add_argument("--data-root", default=None, help="take the class names from <root>/%s" % DATA.train_dir)
    ap.add_argument("--model", default=None, help="architecture (default: the one recorded in the checkpoint)")
    ap.add_argument("--engine", default="auto", choices=["auto", "native", "oracle"])
    ap.add_argument("--batch", type=int, default=32)
    a = ap.parse_args(argv)
    logits, files = predict(a.checkpoint, a.images, a.model, a.engine, a.batch)
    names = a.classes.split(",") if a.classes else None
    if names is None and a.data_root:
        names = find_classes(os.path.join(a.data_root, DATA.train_dir))[0]
    prob = torch.softmax(logits, dim=1)
    for f, p in zip(files, prob):
        k = int(p.argmax())
        print("%s\t%s\t%.4f" % (f, names[k] if names and k < len(names) else k, float(p[k])))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
