"""ImageFolder directory contract.

The reference reads ``<root>/TrainData/<class>/*`` and ``<root>/ValidationData/<class>/*`` through
``torchvision.datasets.ImageFolder`` (distributedVggf.py:96, :109; Readme.md:65-79): classes are
the sorted sub-directory names, the label is the index of the class, files are visited in sorted
order and filtered by extension.  ``scan_image_folder`` re-implements that contract with no
torchvision dependency so the native loader (C++ decode threads) can share the index.
"""
from __future__ import annotations

import os
from typing import List, Tuple

IMG_EXTENSIONS = (".jpg", ".jpeg", ".png", ".ppm", ".bmp", ".pgm", ".tif", ".tiff", ".webp")


def find_classes(directory: str) -> Tuple[List[str], dict]:
    classes = sorted(e.name for e in os.scandir(directory) if e.is_dir())
    if not classes:
        raise FileNotFoundError("Couldn't find any class folder in %s." % directory)
    return classes, {c: i for i, c in enumerate(classes)}


def scan_image_folder(directory: str) -> Tuple[List[str], List[Tuple[str, int]]]:
    """Return (class_names, [(path, label), ...]) in torchvision's order."""
    directory = os.path.expanduser(directory)
    classes, class_to_idx = find_classes(directory)
    samples: List[Tuple[str, int]] = []
    for cls in classes:
        target_dir = os.path.join(directory, cls)
        for root, _, fnames in sorted(os.walk(target_dir, followlinks=True)):
            for fname in sorted(fnames):
                if fname.lower().endswith(IMG_EXTENSIONS):
                    samples.append((os.path.join(root, fname), class_to_idx[cls]))
    if not samples:
        raise FileNotFoundError("Found no valid file in %s" % directory)
    return classes, samples
