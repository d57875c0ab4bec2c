"""Print the gradient-exchange plan of a model without touching a GPU: arena layout in gradient-ready order,
buckets, wire bytes, and which path each bucket takes at a given world size.

    python -m distributed_vgg_f_b200.tools.plan --model vggf --num-classes 3 --world 8

What DDP's Reducer decides at run time from autograd hooks (distributedVggf.py:225; rebuilt buckets of 1 / 25 MiB,
SURVEY N3), this framework fixes before the first step because it owns backward: the table below is exactly what
``NativeEngine`` executes (``parallel/buckets.py::engine_bucket_plan``, ``parallel/symm.py::pick_algo``)."""
from __future__ import annotations

import argparse

from ..models import layout as L
from ..models.vggf import get_spec
from ..parallel.buckets import engine_bucket_plan
from ..parallel.symm import ONESHOT_MAX_BYTES, ONESHOT_MAX_BYTES_NVLS


def describe(model: str = "vggf", num_classes: int = 3, world: int = 8, multicast: bool = True, zero1: str = "auto",
             bucket_mb: float = 32.0, conv_bucket_mb: float = 9.5, tail_bucket_kb: float = 2400.0):
    spec = get_spec(model, num_classes)
    plan = engine_bucket_plan(L.ready_order(spec), bucket_mb, conv_bucket_mb, tail_bucket_kb)
    fc_w = {f.name + ".weight" for f in spec.fcs}
    prepacked = set()
    for name in fc_w:
        ids = plan.bucket_of(name)
        if ids and all(plan.buckets[i].tensors == (name,) for i in ids):
            prepacked.update(ids)
    z1 = world >= 4 if zero1 == "auto" else zero1 == "on"
    rows = []
    for b in plan.buckets:
        wire = b.numel * 2
        if world == 1:
            path = "optimizer only"
        else:
            limit = ONESHOT_MAX_BYTES_NVLS if multicast else ONESHOT_MAX_BYTES
            algo = "oneshot" if wire <= limit else ("nvls" if multicast else "twoshot")
            if b.index in prepacked and algo == "oneshot":
                algo = "twoshot"
            fused = z1 and b.index in prepacked and algo != "oneshot"
            entry = "wire written by the wgrad GEMM" if b.index in prepacked else "pack fp32 -> bf16 * 1/ws"
            path = ("zero1: reduce-scatter + Adam + all-gather, " if fused else "all-reduce + Adam, ") + algo + ", " + entry
        rows.append((b.index, b.numel, wire, len(b.tensors), b.tensors[0], b.tensors[-1], path))
    return plan, rows


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--model", default="vggf")
    ap.add_argument("--num-classes", type=int, default=3)
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--no-multicast", action="store_true", help="plan for a fabric without NVLS")
    ap.add_argument("--zero1", default="auto", choices=["auto", "on", "off"])
    ap.add_argument("--bucket-mb", type=float, default=32.0)
    a = ap.parse_args(argv)
    plan, rows = describe(a.model, a.num_classes, a.world, not a.no_multicast, a.zero1, a.bucket_mb)
    total = sum(r[2] for r in rows)
    print("%s, %d classes: %d elements (native layouts) in %d tensors, arena %d elements (tensors aligned to %d), %d buckets, %.1f MB of bf16 wire per step"
          % (a.model, a.num_classes, sum(plan.numels.values()), len(plan.order), plan.total, plan.align, len(rows), total / 1e6))
    print("%3s %11s %9s %3s  %-24s %-24s %s" % ("#", "elements", "wire MB", "n", "first tensor", "last tensor (completes it)", "path at world=%d" % a.world))
    for i, n, wire, cnt, first, last, path in rows:
        print("%3d %11d %9.2f %3d  %-24s %-24s %s" % (i, n, wire / 1e6, cnt, first, last, path))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
