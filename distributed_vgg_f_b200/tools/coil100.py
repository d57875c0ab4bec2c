"""Regroup a COIL-100 download into the reference's evaluation dataset.

The reference was evaluated on COIL-100 (100 objects x 72 views of 128x128 pixels, files named
``obj<N>__<angle>.png``) split BY HAND into three classes -- edible / toy / other -- with 80 % of the
images for training and 20 % for validation (Readme.md:81-107; 1296 / 1440 / 4464 images).  This tool
does that split reproducibly and writes the ImageFolder layout the trainer expects
(Readme.md:65-79)::

    python -m distributed_vgg_f_b200.tools.coil100 /data/coil-100 /data/coil3 [--val-fraction 0.2] [--seed 0] [--copy]

It also prints inverse-frequency class weights for ``--class-weights`` (the reference carries a
commented-out ``CLASS_OPTIM_WEIGHTS = [0.41, 0.19, 0.4]`` for exactly this imbalance,
distributedUtil.py:27-28).
"""
from __future__ import annotations

import argparse
import os
import random
import re
import shutil
from typing import Dict, List, Sequence

from ..config import DATA

# object ids of the two small classes (Readme.md:98-101); every other object is "other"
EDIBLE = frozenset({2, 4, 7, 47, 49, 53, 62, 63, 67, 72, 73, 75, 82, 83, 84, 93, 94, 98})
TOY = frozenset({6, 8, 14, 15, 17, 19, 20, 23, 27, 28, 34, 37, 48, 51, 52, 69, 74, 76, 91, 100})
_NAME = re.compile(r"^obj(\d+)__(\d+)\.(png|jpg|jpeg|ppm|bmp)$", re.IGNORECASE)


def class_of(obj_id: int) -> str:
    return "edible" if obj_id in EDIBLE else ("toy" if obj_id in TOY else "other")


def group_files(names: Sequence[str]) -> Dict[str, List[str]]:
    """file names -> {class: sorted files}; names that are not COIL-100 views are ignored."""
    groups: Dict[str, List[str]] = {"edible": [], "other": [], "toy": []}
    for n in names:
        m = _NAME.match(os.path.basename(n))
        if m:
            groups[class_of(int(m.group(1)))].append(n)
    return {k: sorted(v) for k, v in groups.items()}


def split_train_val(files: Sequence[str], val_fraction: float, seed: int):
    order = list(files)
    random.Random(seed).shuffle(order)
    n_val = int(round(len(order) * val_fraction))
    return sorted(order[n_val:]), sorted(order[:n_val])


def inverse_frequency_weights(counts: Dict[str, int]) -> List[float]:
    """Weights proportional to 1/count, normalised to sum 1, in sorted class order (= label order)."""
    inv = {k: (1.0 / c if c else 0.0) for k, c in counts.items()}
    tot = sum(inv.values()) or 1.0
    return [round(inv[k] / tot, 4) for k in sorted(counts)]


def prepare(src: str, dst: str, val_fraction: float = 0.2, seed: int = 0, copy: bool = False) -> Dict[str, Dict[str, int]]:
    names = [os.path.join(src, f) for f in sorted(os.listdir(src))]
    groups = group_files(names)
    if not any(groups.values()):
        raise FileNotFoundError("no obj<N>__<angle>.png files under %r" % src)
    report: Dict[str, Dict[str, int]] = {}
    for cls, files in groups.items():
        train, val = split_train_val(files, val_fraction, seed)
        report[cls] = {"train": len(train), "val": len(val)}
        for split, subset in ((DATA.train_dir, train), (DATA.val_dir, val)):
            out = os.path.join(dst, split, cls)
            os.makedirs(out, exist_ok=True)
            for f in subset:
                target = os.path.join(out, os.path.basename(f))
                if os.path.lexists(target):
                    os.remove(target)
                if copy:
                    shutil.copyfile(f, target)
                else:
                    os.symlink(os.path.abspath(f), target)
    return report


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("src", help="directory holding obj<N>__<angle>.png (the unpacked COIL-100 archive)")
    ap.add_argument("dst", help="root directory to create (TrainData/ and ValidationData/ inside)")
    ap.add_argument("--val-fraction", type=float, default=0.2)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--copy", action="store_true", help="copy files instead of symlinking them")
    a = ap.parse_args(argv)
    rep = prepare(a.src, a.dst, a.val_fraction, a.seed, a.copy)
    for cls in sorted(rep):
        print("[Info] %-7s train %5d  val %5d" % (cls, rep[cls]["train"], rep[cls]["val"]))
    w = inverse_frequency_weights({k: v["train"] for k, v in rep.items()})
    print("[Info] inverse-frequency class weights (edible, other, toy): --class-weights %s" % ",".join(map(str, w)))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
