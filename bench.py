#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): VGG-F training images/sec, bs=64 per GPU, weak scaling over
1/2/4/8 B200, device-timed, max over ranks.  Synthetic 128x128 RGB source images, network input
224x224 (what the reference's transforms produce from them), random-init weights.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5
    python bench.py --impl reference ...      # the unmodified reference from baseline/_ref (NCCL DDP)

Two numbers per run:
  value  kernel-only: uint8 inputs already on the device; augment + forward + loss + backward +
         gradient all-reduce + optimizer, K steps between CUDA events (barrier + synchronize on
         both sides), max over ranks.
  e2e    the public API a user calls (Trainer-style ``engine.train_step(batch)`` fed by the
         pinned-memory batch ring): every step copies its uint8 inputs host->device and reads the
         step's loss back device->host.
One JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BASELINE_IMG_S = {1: 30.7, 2: 49.0}     # README table, mb=64, K80 (BASELINE.md "derived")


# ------------------------------------------------------------------------------------- utilities
class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.index, self.proc, self.lines = index, None, []

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self) -> None:
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


def timed_region(torch, dist, world, device, body, steps):
    """barrier + synchronize, K steps between CUDA events, synchronize + barrier; max over ranks."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    e0.record()
    for k in range(steps):
        body(k)
    e1.record()
    torch.cuda.synchronize(device)
    wall = time.perf_counter() - wall0
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]), float(t[1])


def make_host_batches(torch, n_batches, batch, num_classes, seed, pin=True):
    from distributed_vgg_f_b200.data import transforms as T
    from distributed_vgg_f_b200.data.loader import FusedBatch
    from distributed_vgg_f_b200.data.synthetic import synthetic_uint8_batch

    out = []
    g = torch.Generator().manual_seed(seed)
    for i in range(n_batches):
        imgs, labels = synthetic_uint8_batch(batch, 128, num_classes, seed=seed * 131 + i)
        img_t, lab_t = torch.from_numpy(imgs), torch.from_numpy(labels)
        par = T.sample_train_params(batch, 128, 128, g)
        if pin:
            img_t, lab_t, par = img_t.pin_memory(), lab_t.pin_memory(), par.pin_memory()
        out.append(FusedBatch(img_t, par, lab_t, (256, 256), None))
    return out


# --------------------------------------------------------------------------------- native arm
def run_native(args):
    import torch
    import torch.distributed as dist

    from distributed_vgg_f_b200 import ops
    from distributed_vgg_f_b200.data.loader import FusedBatch
    from distributed_vgg_f_b200.engine.native_engine import NativeEngine
    from distributed_vgg_f_b200.models.vggf import get_spec

    rank, world, local = dist_env()
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    spec = get_spec(args.model, args.num_classes)
    eng = NativeEngine(spec, device=device, batch=args.batch, lr=1e-5, optimizer=args.optimizer,
                       allreduce=args.allreduce, wire_dtype=args.wire_dtype, bucket_mb=args.bucket_mb,
                       seed=0, input_hw=args.hw, comm_ctas=args.comm_ctas,
                       zero1={"auto": "auto", "on": True, "off": False}[args.zero1] if world > 1 else False)
    host = make_host_batches(torch, 4, args.batch, args.num_classes, seed=rank + 1)
    dev_batches = [FusedBatch(b.images_u8.to(device), b.params.to(device), b.labels.to(device), b.resized_hw, None)
                   for b in host]

    for k in range(args.warmup):
        eng.train_step(dev_batches[k % len(dev_batches)])
    torch.cuda.synchronize(device)

    clocks = ClockSampler(local)
    clocks.start()
    l0 = ops.launch_count()
    ms, _ = timed_region(torch, dist, world, device,
                         lambda k: eng.train_step(dev_batches[k % len(dev_batches)]), args.steps)
    launches = ops.launch_count() - l0
    # gradient all-reduce as the step sees it (events on the comm stream; outside the timed region)
    ar_steps = min(args.steps, 10)
    eng.comm_timing(True)
    for k in range(ar_steps):
        eng.train_step(dev_batches[k % len(dev_batches)])
    ar_report = eng.comm_report(ar_steps)
    eng.comm_timing(False)

    # end-to-end through the public API: DataManager (decoded uint8 cache -> pinned batch ring filled
    # by the prefetch thread) -> engine.train_step(batch): every step copies its inputs host->device
    # and reads the step's loss back device->host.
    from distributed_vgg_f_b200.data.loader import DataManager
    from distributed_vgg_f_b200.data.synthetic import make_synthetic_imagefolder

    e2e_epoch_steps = min(args.steps, 16)
    root = os.path.join("/tmp", "b200_bench_synth_ws%d_b%d_s%d" % (world, args.batch, e2e_epoch_steps))
    if local == 0 and not os.path.isdir(os.path.join(root, "TrainData")):
        ncls = min(args.num_classes, 16)
        per_class = (e2e_epoch_steps * args.batch * world + ncls - 1) // ncls
        make_synthetic_imagefolder(root + ".tmp", classes=["c%02d" % i for i in range(ncls)],
                                   train_per_class=per_class, val_per_class=1, size=128, seed=7)
        os.replace(root + ".tmp", root)
    if world > 1:
        dist.barrier()
    dm = DataManager(root, args.batch, train=True, world_size=world, rank=rank, seed=0)
    loader = dm.get_loader()
    loss_host = torch.zeros(args.steps, dtype=torch.float32).pin_memory()

    def batches():
        ep = 0
        while True:
            loader.set_epoch(ep)
            for b in loader:
                if b.labels.shape[0] == args.batch:
                    yield b
            ep += 1

    gen = batches()
    for k in range(3):
        eng.train_step(next(gen))
    torch.cuda.synchronize(device)

    def e2e_step(k):
        loss = eng.train_step(next(gen))
        loss_host[k:k + 1].copy_(loss, non_blocking=True)

    ms_e2e, wall_e2e = timed_region(torch, dist, world, device, e2e_step, args.steps)
    clk = clocks.stop()
    b0 = host[0]
    h2d = b0.images_u8.numel() + b0.params.numel() * 4 + b0.labels.numel() * 8

    gb = args.batch * world
    value = gb * args.steps / (ms / 1e3)
    e2e = gb * args.steps / (max(ms_e2e, wall_e2e) / 1e3)
    if rank == 0:
        out = {
            "metric": "VGG-F training images/sec (device-timed, max over ranks), bs=%d/GPU" % args.batch,
            "value": round(value, 2), "unit": "images/sec", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": round(value / BASELINE_IMG_S[world], 2) if world in BASELINE_IMG_S else None,
            "dtype": "bf16", "impl": "native",
            "data": "synthetic 128x128 RGB uint8 images -> fused GPU augment -> 224x224; random-init weights",
            "config": {"model": args.model, "num_classes": args.num_classes, "global_batch": gb,
                       "per_gpu_batch": args.batch, "input": "%dx%d" % (args.hw, args.hw), "seq_len": None,
                       "parallelism": ("dp%d+zero1" if getattr(eng, "zero1", False) else "dp%d") % world,
                       "optimizer": args.optimizer,
                       "allreduce": args.allreduce, "wire_dtype": args.wire_dtype,
                       "l2_policy": "per-step working set (activations+weights, >2 GB) exceeds the 126 MB L2; "
                                    "4 distinct input batches rotated"},
            "e2e": {"value": round(e2e, 2), "unit": "images/sec", "ms_per_step": round(max(ms_e2e, wall_e2e) / args.steps, 4),
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                    "api": "DataManager(ImageFolder on disk, decoded uint8 cache, pinned ring) -> "
                           "NativeEngine.train_step(batch) + async loss read-back"},
            "gpu_launches": int(launches), "clocks": clk,
            "allreduce": ar_report,
            "final_loss": float(loss_host[-1]),
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------ reference arm
def run_reference(args, bf16: bool = False):
    """Unmodified reference (baseline/_ref/distributedVggf.py): its model factory, its Trainer loop,
    its DataManager; torch DDP over NCCL as in SURVEY D4 (BACKEND constant -> "nccl").

    ``bf16=True`` (--impl reference-bf16) is a CONTEXT arm, not the reference arm: the same model
    object and loop under ``torch.autocast(bfloat16)`` with channels_last tensors and DDP's
    ``bf16_compress_hook`` -- the best the library stack (cuDNN / cuBLAS / NCCL) does for this model at
    the precision the native engine computes in (SURVEY 5.8 "baseline to beat")."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "distributedVggf.py")):
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref missing: run python baseline/install_ref.py"}))
        return
    try:
        import torch
        import torch.distributed as dist
        import torchvision
        from torch import nn
    except Exception as e:      # noqa: BLE001
        print(json.dumps({"impl": "reference", "unavailable": "torch/torchvision import failed: %r" % (e,)}))
        return
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    sys.path.insert(0, ref_dir)
    # Harness shims (the reference files themselves stay byte-identical):
    #  * pretrained=True needs the network -> build the same torchvision VGG-16 with
    #    random init; BASELINE.json asks for random-init weights anyway.
    _orig_vgg16 = torchvision.models.vgg16
    torchvision.models.vgg16 = lambda pretrained=False, **kw: _orig_vgg16(weights=None, **kw)
    import distributedUtil as dstUt
    import distributedVggf as ref

    dstUt.BACKEND = "nccl"                       # the one-constant change
    if world > 1:
        dist.init_process_group(backend=dstUt.BACKEND, device_id=device)
    torch.manual_seed(0)
    if args.model == "vgg16":
        # BASELINE config #3 (VGG-16 / 1000 classes): the reference has no such factory -- its model IS
        # torchvision's VGG-16 with the last layer swapped (distributedVggf.py:46-57), so the arm is
        # torchvision's VGG-16 as is, driven by the reference's Trainer loop
        model = _orig_vgg16(weights=None, num_classes=args.num_classes)
    else:
        model = ref.vgg_funnel_model(args.num_classes)
    if bf16:
        model = model.to(memory_format=torch.channels_last)
    if ref.distributed_is_initialized():
        model.to(device)
        model = nn.parallel.DistributedDataParallel(model)     # distributedVggf.py:224-225
        if bf16:
            from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
            model.register_comm_hook(None, default_hooks.bf16_compress_hook)
    else:
        model = nn.DataParallel(model, device_ids=[local])     # distributedVggf.py:227 (1 visible GPU/process)
        model.to(device)
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-5)  # distributedVggf.py:230, README lr

    B = args.batch
    g = torch.Generator().manual_seed(rank)
    dev_batches = [(torch.randn(B, 3, args.hw, args.hw, generator=g).to(device),
                    torch.randint(0, args.num_classes, (B,), generator=g).to(device)) for _ in range(4)]
    if bf16:
        dev_batches = [(x.contiguous(memory_format=torch.channels_last), y) for x, y in dev_batches]
    import contextlib
    amp = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if bf16 else contextlib.nullcontext

    class Loader:
        def __init__(self, n):
            self.n = n

        def __iter__(self):
            for k in range(self.n):
                yield dev_batches[k % len(dev_batches)]

    trainer = ref.Trainer(model, optimizer, Loader(args.warmup), Loader(0), device)
    with amp():
        trainer._Trainer__train()                # warm-up through the reference's own loop
    torch.cuda.synchronize(device)
    clocks = ClockSampler(local)
    clocks.start()
    trainer.train_loader = Loader(args.steps)
    def whole_loop(k):
        if k == 0:
            with amp():
                trainer._Trainer__train()

    ms, _ = timed_region(torch, dist, world, device, whole_loop, args.steps)

    # end-to-end: the reference's DataManager (PIL + transforms, num_workers=0) on a synthetic ImageFolder
    e2e = None
    try:
        from distributed_vgg_f_b200.data.synthetic import make_synthetic_imagefolder
        root = os.path.join("/tmp", "b200_ref_synth_%d" % os.getpid())
        per_class = (args.e2e_steps * B * world + args.num_classes - 1) // args.num_classes
        make_synthetic_imagefolder(root, classes=["c%d" % i for i in range(args.num_classes)],
                                   train_per_class=per_class, val_per_class=1, size=128, seed=rank)
        dm = ref.DataManager(root_folder=root, mini_batch=B, train=True)
        n_steps = len(dm.get_loader())
        trainer.train_loader = dm.get_loader()
        ms_e, wall_e = timed_region(torch, dist, world, device, whole_loop, 1)
        e2e = {"value": round(B * world * n_steps / (max(ms_e, wall_e) / 1e3), 2), "unit": "images/sec",
               "steps": n_steps, "h2d_bytes_per_step": B * 3 * args.hw * args.hw * 4 + B * 8,
               "d2h_bytes_per_step": 8,
               "api": "reference DataManager + Trainer.__train (PIL transforms, pageable H2D, 2 .item() per step)"}
    except Exception as e:      # noqa: BLE001
        e2e = {"unavailable": repr(e)}
    clk = clocks.stop()
    gb = B * world
    value = gb * args.steps / (ms / 1e3)
    if rank == 0:
        print(json.dumps({
            "metric": "VGG-F training images/sec (device-timed, max over ranks), bs=%d/GPU" % B,
            "value": round(value, 2), "unit": "images/sec", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": round(value / BASELINE_IMG_S[world], 2) if world in BASELINE_IMG_S else None,
            "dtype": ("bf16 autocast + channels_last + bf16_compress_hook (context arm)" if bf16
                      else "fp32 (reference default: TF32 conv via cuDNN, fp32 matmul)"),
            "impl": "reference-bf16" if bf16 else "reference",
            "data": "synthetic: random fp32 224x224 device tensors (value) / synthetic 128x128 ImageFolder (e2e)",
            "config": {"model": args.model, "num_classes": args.num_classes, "global_batch": gb, "per_gpu_batch": B,
                       "input": "%dx%d" % (args.hw, args.hw), "parallelism": "dp%d" % world,
                       "backend": "nccl DDP (reference BACKEND constant set to nccl)", "optimizer": "adam"},
            "e2e": e2e, "gpu_launches": 0, "clocks": clk}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference", "reference-bf16"])
    ap.add_argument("--model", default="vggf")
    ap.add_argument("--num-classes", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--hw", type=int, default=224)
    ap.add_argument("--optimizer", default="adam")
    ap.add_argument("--allreduce", default="auto")
    ap.add_argument("--wire-dtype", default="bf16")
    ap.add_argument("--bucket-mb", type=float, default=32.0)
    ap.add_argument("--comm-ctas", type=int, default=48)
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--zero1", default="auto", choices=["auto", "on", "off"],
                    help="fused reduce-scatter + Adam + all-gather for the FC-weight buckets (auto: from 4 ranks up)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch ourselves under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533"),
               os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl in ("reference", "reference-bf16"):
        run_reference(args, bf16=args.impl == "reference-bf16")
    else:
        run_native(args)


if __name__ == "__main__":
    main()
