"""Running metrics.

``Average`` / ``Accuracy2`` keep the reference's public surface and string formats
(distributedUtil.py:31-99: ``'{:.6f}'`` for the loss mean, ``'{:.2f}%'`` for accuracy) so that the
per-epoch report line is byte-compatible.  The reference feeds them with two ``.item()`` host
syncs per step (distributedVggf.py:174-175); ``DeviceMeter`` is the B200-native replacement: the
fused cross-entropy kernel adds (loss_sum, correct, count) into a 3-float device buffer and the
host reads it once per epoch.
"""
from __future__ import annotations

import torch


class Average:
    """Sample-weighted running mean (reference: distributedUtil.py:31-58)."""

    def __init__(self) -> None:
        self.sum = 0.0
        self.count = 0

    def __str__(self) -> str:
        return "{:.6f}".format(self.average)

    @property
    def average(self) -> float:
        return self.sum / self.count if self.count else float("nan")

    def update(self, value, number) -> None:
        self.sum += float(value) * number
        self.count += number


class Accuracy2:
    """Running top-1 accuracy (reference: distributedUtil.py:61-99)."""

    def __init__(self) -> None:
        self.correct = 0
        self.count = 0

    def __str__(self) -> str:
        return "{:.2f}%".format(self.accuracy * 100)

    @property
    def accuracy(self) -> float:
        return self.correct / self.count if self.count else float("nan")

    def update(self, output: torch.Tensor, target: torch.Tensor) -> None:
        with torch.no_grad():
            self.correct += int(output.argmax(dim=1).eq(target).sum().item())
        self.count += output.size(0)

    def update_counts(self, correct: int, number: int) -> None:
        self.correct += int(correct)
        self.count += int(number)


class DeviceMeter:
    """Three fp32 device accumulators: [sum of per-sample loss, #correct, #samples].

    The native engine's cross-entropy kernel atomically adds into ``buf``; nothing touches the
    host until ``snapshot`` is called (once per epoch).
    """

    def __init__(self, device) -> None:
        self.buf = torch.zeros(4, dtype=torch.float32, device=device)

    def reset(self) -> None:
        self.buf.zero_()

    def add_reference(self, logits: torch.Tensor, target: torch.Tensor, weight=None) -> None:
        """Torch-op fallback used by the oracle path (same arithmetic as the kernel): accumulates
        ``loss * batch`` exactly like ``Average.update(loss.item(), batch)`` in the reference's loop,
        where ``loss`` is the (optionally class-weighted) mean torch reports."""
        with torch.no_grad():
            logp = torch.log_softmax(logits.float(), dim=1)
            nll = -logp.gather(1, target.view(-1, 1)).view(-1)
            if weight is None:
                self.buf[0] += nll.sum()
            else:
                w = weight.to(nll.device, nll.dtype)[target]
                self.buf[0] += (w * nll).sum() / w.sum() * logits.shape[0]
            self.buf[1] += logits.argmax(dim=1).eq(target).sum()
            self.buf[2] += logits.shape[0]

    def snapshot(self):
        loss_sum, correct, count, _ = self.buf.tolist()
        avg, acc = Average(), Accuracy2()
        avg.sum, avg.count = loss_sum, int(count)
        acc.correct, acc.count = int(correct), int(count)
        return avg, acc
