"""Synthetic ImageFolder generator (BASELINE.json: "synthetic 128x128x3 ImageFolder").

Writes ``<root>/TrainData/<class>/*.png`` and ``<root>/ValidationData/<class>/*.png`` in the layout
the reference expects (Readme.md:65-79).  Each class gets a distinct low-frequency colour pattern
plus noise so that a few epochs of training separate the classes (used by the convergence test).
"""
from __future__ import annotations

import os
from typing import Sequence

import numpy as np

from ..config import DATA


def _class_image(rng: np.random.Generator, cls: int, size: int) -> np.ndarray:
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32) / size
    phase = rng.uniform(0, 2 * np.pi)
    freq = 1.5 + cls
    base = np.stack([
        0.5 + 0.5 * np.sin(2 * np.pi * freq * xx + phase + 0.9 * cls),
        0.5 + 0.5 * np.sin(2 * np.pi * freq * yy + phase + 2.1 * cls),
        0.5 + 0.5 * np.sin(2 * np.pi * freq * (xx + yy) + phase + 0.3 * cls),
    ], axis=-1)
    img = 0.75 * base + 0.25 * rng.random((size, size, 3), dtype=np.float32)
    return (np.clip(img, 0, 1) * 255).astype(np.uint8)


def make_synthetic_imagefolder(root: str, classes: Sequence[str] = ("edible", "other", "toy"),
                               train_per_class: int = 8, val_per_class: int = 4,
                               size: int = 128, seed: int = 0) -> str:
    from PIL import Image

    rng = np.random.default_rng(seed)
    for split, count in ((DATA.train_dir, train_per_class), (DATA.val_dir, val_per_class)):
        for ci, cname in enumerate(classes):
            d = os.path.join(root, split, cname)
            os.makedirs(d, exist_ok=True)
            for i in range(count):
                Image.fromarray(_class_image(rng, ci, size)).save(os.path.join(d, "img_%05d.png" % i))
    return root


def synthetic_uint8_batch(n: int, size: int = 128, num_classes: int = 3, seed: int = 0):
    """In-memory equivalent of decoding ``n`` synthetic PNGs: (uint8 [n,size,size,3], int64 [n])."""
    rng = np.random.default_rng(seed)
    labels = rng.integers(0, num_classes, size=n)
    imgs = np.stack([_class_image(rng, int(c), size) for c in labels])
    return imgs, labels.astype(np.int64)
