"""DataManager: ImageFolder -> mini-batches, in two flavours.

Reference (distributedVggf.py:63-123): ``DataManager(root_folder, mini_batch, train)`` exposes
``class_names / number_classes / data_size / loader / get_loader()``; the loader is a synchronous
``DataLoader(num_workers=0, pin_memory=False)`` over PIL transforms, sharded by a
``DistributedSampler`` only for the training split (validation runs in full on every rank).

Same surface here, two pipelines:
  * ``pipeline="reference"`` -- PIL decode + the literal torchvision transforms per sample.  Yields
    ``(float32 [mb,3,224,224], int64 [mb])``.  Bit-level semantics of the reference; ~400 img/s.
  * ``pipeline="fused"`` (default) -- every image is decoded ONCE into a uint8 cache (thread pool),
    a background thread gathers each batch into a pinned staging ring together with the sampled
    transform parameters, and the consumer runs the fused augment op on the device.  Yields
    ``FusedBatch``.  Requires a uniform source size (true for COIL-100 / the synthetic set) and a
    split that fits in host memory; ``pipeline="auto"`` (the CLI default) switches to the
    reference pipeline when either does not hold.
Ragged final batches are preserved (``drop_last=False`` in the reference).
"""
from __future__ import annotations

import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Iterator, List, NamedTuple, Optional, Tuple

import numpy as np
import torch

from ..config import DATA
from . import transforms as T
from .folder import scan_image_folder
from .sampler import ShardedSampler


class FusedBatch(NamedTuple):
    images_u8: torch.Tensor     # uint8 [mb, H, W, 3] (pinned when CUDA is present)
    params: torch.Tensor        # float32 [mb, 8]     (same)
    labels: torch.Tensor        # int64 [mb]          (same)
    resized_hw: Tuple[int, int]
    state: Optional[dict] = None    # consumer may set state["event"]: slot is reusable after it fires

    def to_float(self, device="cpu") -> Tuple[torch.Tensor, torch.Tensor]:
        """Evaluate the fused transform with torch ops (CPU / oracle path)."""
        x = T.augment_reference(self.images_u8.to(device), self.params, self.resized_hw)
        return x, self.labels.to(device)


def _decode(path: str) -> np.ndarray:
    from PIL import Image

    with open(path, "rb") as f:
        return np.asarray(Image.open(f).convert("RGB"), dtype=np.uint8)


class CacheUnavailable(ValueError):
    """The split cannot be held as one uint8 tensor (mixed image sizes, or too big for host RAM)."""


def _cache_budget_bytes() -> int:
    """B200_MAX_CACHE_GB, else half of the memory the OS reports as available (at most 64 GB)."""
    env = os.environ.get("B200_MAX_CACHE_GB")
    if env:
        return int(float(env) * (1 << 30))
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return min(int(line.split()[1]) * 1024 // 2, 64 << 30)
    except OSError:
        pass
    return 16 << 30


class DecodedCache:
    """All images of a split, decoded once to uint8 HWC (optionally by the native PNG decoder)."""

    def __init__(self, samples: List[Tuple[str, int]], threads: int = 8, max_bytes: Optional[int] = None) -> None:
        paths = [p for p, _ in samples]
        if paths:          # size the cache from the first image before decoding everything
            h, w, _ = _decode(paths[0]).shape
            need, budget = len(paths) * h * w * 3, (_cache_budget_bytes() if max_bytes is None else max_bytes)
            if need > budget:
                raise CacheUnavailable("decoded cache needs %.1f GB for %d images of %dx%d, budget %.1f GB "
                                       "(B200_MAX_CACHE_GB)" % (need / 2**30, len(paths), h, w, budget / 2**30))
        images = None
        try:   # native multi-threaded PNG decode (csrc/png_decode.cpp), falls back to PIL
            from ..ops import native_decode_pngs
            images = native_decode_pngs(paths, threads)
        except Exception:
            images = None
        self.native_decode = images is not None
        if images is None:
            with ThreadPoolExecutor(max_workers=threads) as ex:
                arrays = list(ex.map(_decode, paths))
            shapes = {a.shape for a in arrays}
            if len(shapes) != 1:
                raise CacheUnavailable("fused pipeline needs a uniform image size, found %s; "
                                       "use pipeline='reference'" % sorted(shapes)[:4])
            images = torch.from_numpy(np.stack(arrays))
        self.images = images                                          # [N,H,W,3] uint8
        self.labels = torch.tensor([l for _, l in samples], dtype=torch.int64)
        self.src_hw = (int(self.images.shape[1]), int(self.images.shape[2]))


class _FusedLoader:
    """Batches for the fused GPU pipeline.  With the native extension present, a C++ worker thread
    (csrc/prefetch.cpp) gathers images from the decoded cache into a pinned ring and draws the
    transform parameters without ever taking the GIL; otherwise a Python thread does the same."""

    def __init__(self, cache: DecodedCache, mb: int, train: bool, sampler: Optional[ShardedSampler],
                 shuffle: bool, seed: int, prefetch: int = 3, pin: Optional[bool] = None,
                 native: Optional[bool] = None) -> None:
        self.cache, self.mb, self.train = cache, mb, train
        self.sampler, self.shuffle, self.seed = sampler, shuffle, seed
        self.prefetch = prefetch
        self.pin = torch.cuda.is_available() if pin is None else pin
        self.epoch = 0
        H, W = cache.src_hw
        self.resized_hw = T.resized_dims(train, H, W)
        slots = prefetch + 1
        self._img = self._buf((slots, mb, H, W, 3), torch.uint8)
        self._par = self._buf((slots, mb, T.PARAM_DIM), torch.float32)
        self._lab = self._buf((slots, mb), torch.int64)
        self._ring = [(self._img[s], self._par[s], self._lab[s]) for s in range(slots)]
        self._state = [{} for _ in range(slots)]
        self._pf = None
        if native is None or native:
            try:
                from .. import ops
                if ops.available() and hasattr(ops.require(), "Prefetcher"):
                    self._pf = ops.require().Prefetcher(
                        cache.images, cache.labels, self._img, self._par, self._lab, train,
                        DATA.crop_scale[0], DATA.crop_scale[1], DATA.crop_ratio[0], DATA.crop_ratio[1],
                        DATA.rotation_deg)
            except Exception:
                if native:
                    raise
                self._pf = None

    def _buf(self, shape, dtype):
        t = torch.empty(shape, dtype=dtype)
        return t.pin_memory() if self.pin else t

    def __len__(self) -> int:
        n = len(self.sampler) if self.sampler is not None else len(self.cache.labels)
        return (n + self.mb - 1) // self.mb

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch
        if self.sampler is not None:
            self.sampler.set_epoch(epoch)

    def _indices(self) -> List[int]:
        if self.sampler is not None:
            return list(iter(self.sampler))
        n = len(self.cache.labels)
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            return torch.randperm(n, generator=g).tolist()
        return list(range(n))

    def _epoch_seed(self) -> int:
        return 1_000_003 * (self.seed + 1) + self.epoch * 7919 + (self.sampler.rank if self.sampler else 0)

    def __iter__(self) -> Iterator[FusedBatch]:
        if self._pf is not None:
            return self._iter_native()
        return self._iter_python()

    # -- native worker thread ---------------------------------------------------------------------
    def _iter_native(self) -> Iterator[FusedBatch]:
        self._pf.start(self._indices(), self._epoch_seed())
        prev = None
        try:
            while True:
                item = self._pf.next()           # blocks with the GIL released
                if prev is not None:             # consumer asked for more: previous slot may be refilled
                    ev = self._state[prev].pop("event", None)
                    self._state[prev]["held"] = ev          # keep the event alive until the worker used it
                    self._pf.release(prev, int(ev.cuda_event) if ev is not None else 0)
                    prev = None
                if item is None:
                    break
                slot, k = item
                img, par, lab = self._ring[slot]
                prev = slot
                yield FusedBatch(img[:k], par[:k], lab[:k], self.resized_hw, self._state[slot])
        finally:
            self._pf.stop()

    # -- Python fallback ---------------------------------------------------------------------------
    def _iter_python(self) -> Iterator[FusedBatch]:
        idx = self._indices()
        gen = torch.Generator().manual_seed(self._epoch_seed())
        H, W = self.cache.src_hw
        q: "queue.Queue" = queue.Queue(maxsize=self.prefetch)
        free: "queue.Queue" = queue.Queue()
        for slot in range(len(self._ring)):
            free.put(slot)

        def produce():
            for b0 in range(0, len(idx), self.mb):
                sel = torch.tensor(idx[b0:b0 + self.mb], dtype=torch.int64)
                k = len(sel)
                slot = free.get()
                ev = self._state[slot].pop("event", None)
                if ev is not None:
                    ev.synchronize()        # async H2D of the previous occupant has finished
                img, par, lab = self._ring[slot]
                torch.index_select(self.cache.images, 0, sel, out=img[:k])
                torch.index_select(self.cache.labels, 0, sel, out=lab[:k])
                par[:k] = (T.sample_train_params(k, H, W, gen) if self.train
                           else T.val_params(k, H, W))
                q.put((slot, k))
            q.put(None)

        th = threading.Thread(target=produce, daemon=True)
        th.start()
        prev = None
        while True:
            item = q.get()
            if prev is not None:
                free.put(prev)      # consumer is done with the previous slot once it asks for more
            if item is None:
                break
            slot, k = item
            img, par, lab = self._ring[slot]
            prev = slot
            yield FusedBatch(img[:k], par[:k], lab[:k], self.resized_hw, self._state[slot])
        th.join()


class _ReferenceLoader:
    """Synchronous PIL + torchvision loader with the reference's exact semantics."""

    def __init__(self, samples, mb: int, train: bool, sampler: Optional[ShardedSampler],
                 shuffle: bool, seed: int) -> None:
        self.samples, self.mb, self.sampler = samples, mb, sampler
        self.shuffle, self.seed, self.epoch = shuffle, seed, 0
        self.tf = T.reference_transforms(train)

    def __len__(self) -> int:
        n = len(self.sampler) if self.sampler is not None else len(self.samples)
        return (n + self.mb - 1) // self.mb

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch
        if self.sampler is not None:
            self.sampler.set_epoch(epoch)

    def __iter__(self):
        from PIL import Image

        if self.sampler is not None:
            idx = list(iter(self.sampler))
        elif self.shuffle:
            idx = torch.randperm(len(self.samples)).tolist()
        else:
            idx = list(range(len(self.samples)))
        for b0 in range(0, len(idx), self.mb):
            xs, ys = [], []
            for i in idx[b0:b0 + self.mb]:
                path, label = self.samples[i]
                with open(path, "rb") as f:
                    xs.append(self.tf(Image.open(f).convert("RGB")))
                ys.append(label)
            yield torch.stack(xs), torch.tensor(ys, dtype=torch.int64)


class DataManager:
    """Same constructor / attributes as the reference's class (distributedVggf.py:63-123)."""

    def __init__(self, root_folder: str, mini_batch: int, train: bool = True, *,
                 world_size: int = 1, rank: int = 0, pipeline: str = "fused", seed: int = 0,
                 reference_order: bool = False, decode_threads: int = 8, shard_eval: bool = False) -> None:
        if root_folder is None:
            raise ValueError("-rd/--root_dir is required")
        self.root_folder, self.mb_size, self.train = root_folder, mini_batch, train
        self.path = os.path.join(root_folder, DATA.train_dir if train else DATA.val_dir)
        self.class_names, self.samples = scan_image_folder(self.path)
        self.number_classes = len(self.class_names)
        self.data_size = len(self.samples)
        sampler = None
        if train and world_size > 1:      # validation is NOT sharded (distributedVggf.py:115)
            sampler = ShardedSampler(self.data_size, world_size, rank, seed=seed,
                                     reference_order=reference_order)
        elif shard_eval and world_size > 1:   # extension: exact partition, metrics all-reduced
            sampler = ShardedSampler(self.data_size, world_size, rank, shuffle=False, pad=False)
        if pipeline in ("fused", "auto"):
            try:
                self.cache = DecodedCache(self.samples, decode_threads)
            except CacheUnavailable as e:
                if pipeline == "fused":
                    raise
                # mixed sizes / larger than host memory: the reference's per-sample PIL pipeline
                # handles anything an ImageFolder can hold (every rank sees the same files, so
                # every rank takes the same branch)
                print("[Info] %s -> using the per-sample (reference) input pipeline" % e, flush=True)
                pipeline = "reference"
        self.pipeline = pipeline = "fused" if pipeline == "auto" else pipeline
        if pipeline == "fused":
            self.loader = _FusedLoader(self.cache, mini_batch, train, sampler, True, seed)
        elif pipeline == "reference":
            self.loader = _ReferenceLoader(self.samples, mini_batch, train, sampler, True, seed)
        else:
            raise ValueError("pipeline must be 'fused', 'reference' or 'auto'")

    def get_loader(self):
        return self.loader
