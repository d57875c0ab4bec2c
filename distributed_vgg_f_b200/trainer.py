"""Training / evaluation loops.

Reference (distributedVggf.py:126-197): ``Trainer(model, optimizer, train_loader, test_loader,
device).fit(epochs)`` -- per batch: H2D, forward, cross-entropy, zero_grad, backward, step, two
``.item()`` syncs; per epoch: a full (unsharded) validation pass on every rank and one report line
``[Info] Epoch: e/E, train loss: .., train acc: ..%, test loss: .., test acc: ..%.``.

``Trainer`` keeps the constructor, ``fit`` and the report line, for two interchangeable back ends:
  * a torch ``nn.Module`` (optionally wrapped in ``parallel.ddp.FlatDDP``) -- the oracle / CPU path;
  * ``engine.NativeEngine`` -- the sm_100a path, where forward, loss, backward, gradient
    all-reduce and optimizer are one ``train_step`` call and metrics stay on the device.
Differences kept on purpose: metrics accumulate on the device and are read once per epoch
(``DeviceMeter``), ``set_epoch`` is called so shuffling changes between epochs, and throughput is
reported on an extra line.
"""
from __future__ import annotations

import time
from typing import Optional

import torch
import torch.nn.functional as F

from .data.loader import FusedBatch
from .utils.metrics import DeviceMeter


def _bf16_peak() -> float:
    """Sustained bf16 TFLOP/s of this machine as measured by the driver, else the B200 fallback."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        for key in ("bf16_tflops_sustained", "cublas_bf16_tflops_sustained", "bf16_tflops"):
            if key in d:
                return float(d[key])
    except (OSError, ValueError):
        pass
    return 1405.9


def sys_stdout_flush() -> None:
    import sys
    sys.stdout.flush()


def _is_native(model) -> bool:
    return hasattr(model, "train_step") and hasattr(model, "eval_step")


class Trainer:
    def __init__(self, model, optimizer, train_loader, test_loader, device,
                 class_weights: Optional[torch.Tensor] = None, verbose_throughput: bool = True,
                 on_epoch_end=None, shard_eval: bool = False) -> None:
        self.model = model
        self.optimizer = optimizer
        self.train_loader = train_loader
        self.test_loader = test_loader
        self.device = device
        # The reference carries a commented-out class-weighted loss (distributedVggf.py:164-166).
        self.class_weights = class_weights.to(device) if class_weights is not None else None
        self.verbose_throughput = verbose_throughput
        self.on_epoch_end = on_epoch_end
        self.shard_eval = shard_eval
        if self.class_weights is not None and _is_native(model):
            model.set_class_weights(self.class_weights)
        self.history = []

    # ------------------------------------------------------------------------------------------
    def fit(self, epochs: int, start_epoch: int = 1) -> None:
        for epoch in range(start_epoch, epochs + 1):
            for ld in (self.train_loader, self.test_loader):
                if hasattr(ld, "set_epoch"):
                    ld.set_epoch(epoch - 1)
            t0 = time.perf_counter()
            train_loss, train_acc = self._train()
            t1 = time.perf_counter()
            test_loss, test_acc = self._evaluate()
            print(
                "[Info] Epoch: {}/{},".format(epoch, epochs),
                "train loss: {}, train acc: {},".format(train_loss, train_acc),
                "test loss: {}, test acc: {}.".format(test_loss, test_acc),
                flush=True,
            )
            if self.verbose_throughput and train_loss.count:
                print("[Perf] Epoch: {}/{}, train images/sec (this rank): {:.1f}{}{}".format(
                    epoch, epochs, train_loss.count / max(t1 - t0, 1e-9),
                    self._utilisation(train_loss.count, t1 - t0), self._comm_line()), flush=True)
                ht = getattr(self, "epoch_host_times", None)
                if ht:
                    print("[Perf] Epoch: {}/{}, host: first batch after {:.1f} ms, steps enqueued in {:.1f} ms, "
                          "GPU drained {:.1f} ms later; validation pass {:.1f} ms".format(
                              epoch, epochs, ht["first_batch_ms"], ht["enqueue_ms"], ht["drain_ms"],
                              1e3 * (time.perf_counter() - t1)), flush=True)
            self.history.append(dict(epoch=epoch, train_loss=train_loss.average,
                                     train_acc=train_acc.accuracy, test_loss=test_loss.average,
                                     test_acc=test_acc.accuracy))
            if self.on_epoch_end is not None:
                self.on_epoch_end(epoch, self)

    @staticmethod
    def _print_timeline(rows) -> None:
        """--profile timeline: one line per mark, compute stream first, then the side streams."""
        if not rows:
            return
        print("[Timeline] last training step of the epoch, ms since its start (CUDA events)", flush=True)
        prev = 0.0
        for name, lane, ms in rows:
            if lane == "compute":
                print("[Timeline]   compute %-26s %8.3f  +%.3f" % (name, ms, ms - prev))
                prev = ms
        for name, lane, ms in rows:
            if lane != "compute":
                print("[Timeline]   %-7s %-26s %8.3f" % (lane, name, ms))
        sys_stdout_flush()

    def _comm_line(self) -> str:
        """Gradient all-reduce of the epoch as the step saw it: bus GB/s and fraction of the NVLink rate
        (SURVEY 5.5).  Needs ``engine.comm_timing(True)`` (the CLI's --profile comm)."""
        rep = getattr(self.model, "comm_report", None)
        steps = len(self.train_loader) if hasattr(self.train_loader, "__len__") else 0
        r = rep(max(steps, 1)) if rep is not None else None
        self.epoch_allreduce = r
        if not r:
            return ""
        return ", grad all-reduce {:.0f} MB/step in {:.2f} ms = {:.0f} GB/s bus ({:.0f}% of 770 measured, {:.0f}% of 900 nominal)".format(
            r["wire_MB_per_step"], r["ms_per_step"], r["bus_GBs"], 100 * r["frac_of_770_measured"],
            100 * r["frac_of_900_nominal"])

    def _utilisation(self, images: int, seconds: float) -> str:
        """Native engine only: achieved training TFLOP/s (3 x forward FLOPs of the layer table) and
        its fraction of the measured bf16 peak (MEASURED_PEAKS.json next to the package, if any)."""
        spec, hw = getattr(self.model, "spec", None), getattr(self.model, "HW", None)
        if not _is_native(self.model) or spec is None or not hw or seconds <= 0:
            return ""
        tflops = 3.0 * spec.flops_per_image(hw) * images / seconds / 1e12
        return ", {:.0f} TFLOP/s ({:.0f}% of the measured bf16 peak)".format(tflops, 100.0 * tflops / _bf16_peak())

    # ------------------------------------------------------------------------------------------
    def _to_device(self, batch):
        if isinstance(batch, FusedBatch):
            return batch.to_float(self.device)
        inputs, targets = batch
        return inputs.to(self.device, non_blocking=True), targets.to(self.device, non_blocking=True)

    def _train(self):
        meter = DeviceMeter(self.device)
        if _is_native(self.model):
            self.model.set_meter(meter)
            t0 = time.perf_counter()
            first = None
            want_tl = getattr(self, "profile_timeline", False) and hasattr(self.model, "timeline")
            n_steps = len(self.train_loader) if hasattr(self.train_loader, "__len__") else 0
            for k, batch in enumerate(self.train_loader):
                if first is None:
                    first = time.perf_counter() - t0       # loader start-up: time to the first batch
                if want_tl and k == n_steps - 1:           # trace the epoch's last step
                    self.model.sync()
                    self.model.timeline(True)
                self.model.train_step(batch)
            t1 = time.perf_counter()
            self.model.sync()
            if want_tl:
                self._print_timeline(self.model.timeline_report())
                self.model.timeline(False)
            self.epoch_host_times = {"first_batch_ms": 1e3 * (first or 0.0), "enqueue_ms": 1e3 * (t1 - t0),
                                     "drain_ms": 1e3 * (time.perf_counter() - t1)}
            return meter.snapshot()

        self.model.train()
        wrapper = self.model if hasattr(self.model, "finish_backward") else None
        for batch in self.train_loader:
            inputs, targets = self._to_device(batch)
            outputs = self.model(inputs)
            loss = F.cross_entropy(outputs, targets, weight=self.class_weights)
            if wrapper is not None:
                wrapper.zero_grad()
            else:
                self.optimizer.zero_grad()
            loss.backward()
            if wrapper is not None:
                wrapper.finish_backward()
            self.optimizer.step()
            meter.add_reference(outputs.detach(), targets, self.class_weights)
        return meter.snapshot()

    def _reduce_eval(self, meter: DeviceMeter):
        """--shard-eval: every rank saw a disjoint slice of the validation set; sum the counters."""
        import torch.distributed as dist

        if self.shard_eval and dist.is_available() and dist.is_initialized():
            dist.all_reduce(meter.buf)
        return meter.snapshot()

    def _evaluate(self):
        meter = DeviceMeter(self.device)
        if _is_native(self.model):
            self.model.set_meter(meter)
            for batch in self.test_loader:
                self.model.eval_step(batch)
            self.model.sync()
            return self._reduce_eval(meter)

        self.model.eval()
        with torch.no_grad():
            for batch in self.test_loader:
                inputs, targets = self._to_device(batch)
                outputs = self.model(inputs)
                meter.add_reference(outputs, targets)
        return self._reduce_eval(meter)
