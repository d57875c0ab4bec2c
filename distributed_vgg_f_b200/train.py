"""Orchestration: data -> model -> parallel wrapper -> optimizer -> Trainer.fit.

Reference: ``manage_training(args)`` (distributedVggf.py:200-236) -- pick the device, build the
train / validation DataManagers, derive ``num_classes`` from the training folder, print three
``[Info]`` banner lines (+ one when distributed), build the model, wrap it (DDP when the process
group exists, DataParallel otherwise), build Adam(lr) and run ``Trainer.fit``.

Same sequence and same banner here.  What changes: rank -> GPU binding, the single-GPU path is a
plain engine (no DataParallel), the engine is either ``NativeEngine`` (sm_100a kernels, fused
gradient all-reduce over peer memory) or the torch-op oracle, and checkpoint / resume / LR step /
JSONL logging are available as additive options.
"""
from __future__ import annotations

import json
import os
import time

import torch

from .data.loader import DataManager
from .models.vggf import build_oracle, get_spec
from .parallel.ddp import FlatDDP
from .parallel.process_group import (device_name, distributed_is_initialized, init_distributed,
                                     pick_device)
from .trainer import Trainer
from .utils import checkpoint as ckpt


def _want_native(args, device: torch.device) -> bool:
    if args.engine == "oracle":
        return False
    native_ok = device.type == "cuda" and torch.cuda.get_device_capability(device)[0] >= 10
    if args.engine == "native" and not native_ok:
        raise RuntimeError("--engine native needs an sm_100 GPU (got %s)" % device)
    if native_ok and getattr(args, "dtype", "bf16") == "fp32":
        # The sm_100a kernels compute in bf16 (fp32 master weights, moments and accumulators).  The reference's
        # fp32 arithmetic (distributedVggf.py:162-172) is served by the torch-op engine -- same model, same
        # bucketed gradient averaging (FlatDDP) -- until a tcgen05 kind::tf32 path exists.
        if args.engine == "native":
            raise RuntimeError("--engine native computes in bf16; --dtype fp32 runs on --engine oracle (torch ops)")
        print("[Info] --dtype fp32: using the torch-op engine (the native kernels compute in bf16)", flush=True)
        return False
    return native_ok


def build_optimizer(params, args):
    if args.optimizer == "sgd":
        return torch.optim.SGD(params, lr=args.learning_rate, momentum=args.momentum)
    return torch.optim.Adam(params, lr=args.learning_rate)     # distributedVggf.py:230


def manage_training(args) -> Trainer:
    device = pick_device(args.rank, args.no_cuda)
    if args.world_size > 1 and not distributed_is_initialized():
        init_distributed(args.init_url, args.rank, args.world_size, device,
                         getattr(args, "backend", None))

    if getattr(args, "synthetic", 0):
        # Rank 0 alone decides whether the set has to be generated, writes it into a temporary
        # directory and renames it into place; EVERY rank then passes the same barrier.  (A per-rank
        # isdir() test races with rank 0 creating the directory: a late rank would skip the barrier,
        # scan a half-written folder and pair rank 0's barrier with some other collective.)
        if args.rank == 0 and not os.path.isdir(os.path.join(args.root_dir, "TrainData")):
            import shutil

            from .data.synthetic import make_synthetic_imagefolder
            tmp = args.root_dir.rstrip("/") + ".tmp%d" % os.getpid()
            make_synthetic_imagefolder(tmp, train_per_class=args.synthetic,
                                       val_per_class=max(args.synthetic // 4, 1), seed=args.seed)
            os.makedirs(args.root_dir, exist_ok=True)
            for sub in sorted(os.listdir(tmp), reverse=True):     # TrainData last
                os.replace(os.path.join(tmp, sub), os.path.join(args.root_dir, sub))
            shutil.rmtree(tmp, ignore_errors=True)
        if distributed_is_initialized():
            torch.distributed.barrier()

    dm_kw = dict(world_size=args.world_size, rank=args.rank, pipeline=args.pipeline,
                 seed=args.seed, reference_order=args.reference_order)
    train_data_mngr = DataManager(args.root_dir, args.mini_batch, train=True, **dm_kw)
    valid_data_mngr = DataManager(args.root_dir, args.mini_batch, train=False,
                                  shard_eval=getattr(args, "shard_eval", False), **dm_kw)
    train_loader, valid_loader = train_data_mngr.get_loader(), valid_data_mngr.get_loader()

    num_classes = args.num_classes or train_data_mngr.number_classes
    print("[Info] number of classes: {}".format(num_classes))
    print("[Info] class labels: {}".format(train_data_mngr.class_names))
    print("[Info] Running instance {} using {}".format(args.rank, device_name(device)))

    spec = get_spec(args.model, num_classes)
    pretrained = torch.load(args.pretrained, map_location="cpu") if args.pretrained else None

    if distributed_is_initialized():
        print("[Info] distributed training has been initialized")

    if _want_native(args, device):
        from .engine.native_engine import NativeEngine

        model = NativeEngine(spec, device=device, batch=args.mini_batch, lr=args.learning_rate,
                             optimizer=args.optimizer, momentum=args.momentum,
                             compute_dtype=args.dtype, allreduce=args.allreduce,
                             wire_dtype=args.wire_dtype, bucket_mb=args.bucket_mb, seed=args.seed,
                             pretrained_state=pretrained, profile=args.profile,
                             zero1={"auto": "auto", "on": True, "off": False}[getattr(args, "zero1", "auto")])
        optimizer = None            # fused into the engine
        if model.world > 1:
            model.comm_timing(True)   # [Perf] line: all-reduce GB/s and fraction of the NVLink rate per epoch
    else:
        model = build_oracle(spec, seed=args.seed, pretrained_state=pretrained).to(device)
        if distributed_is_initialized():
            model = FlatDDP(model, bucket_cap_mb=args.bucket_mb)
        optimizer = build_optimizer(model.parameters(), args)

    start_epoch = 1
    if args.resume:
        start_epoch = ckpt.load_checkpoint(args.resume, model, optimizer) + 1
        print("[Info] resumed from {} at epoch {}".format(args.resume, start_epoch))

    base_lr = args.learning_rate
    if args.resume and args.lr_step and start_epoch > 1:
        # a run resumed behind an LR-step boundary continues at the decayed rate on BOTH back ends
        resumed_lr = base_lr * (args.lr_gamma ** ((start_epoch - 1) // args.lr_step))
        if optimizer is not None:
            for g in optimizer.param_groups:
                g["lr"] = resumed_lr
        else:
            model.set_lr(resumed_lr)

    def on_epoch_end(epoch: int, trainer: Trainer) -> None:
        if args.lr_step and epoch % args.lr_step == 0:
            new_lr = base_lr * (args.lr_gamma ** (epoch // args.lr_step))
            if optimizer is not None:
                for g in optimizer.param_groups:
                    g["lr"] = new_lr
            else:
                model.set_lr(new_lr)
        if args.save:
            ckpt.save_checkpoint(args.save, model, optimizer, epoch, vars(args),
                                 is_rank0=(args.rank == 0))
            # rank 0 is busy with a ~1.6 GB D2H copy + torch.save: hold the other ranks HERE, on the
            # host, instead of letting them spin inside the next epoch's first all-reduce kernel
            if distributed_is_initialized():
                torch.distributed.barrier()
        if args.log_jsonl and args.rank == 0:
            rec = dict(trainer.history[-1], time=time.time(), world_size=args.world_size)
            if hasattr(model, "phase_times"):
                rec["phase_ms"] = model.phase_times()
            for key in ("epoch_host_times", "epoch_allreduce"):       # set by Trainer for native engines
                if getattr(trainer, key, None):
                    rec[key[len("epoch_"):]] = getattr(trainer, key)
            with open(args.log_jsonl, "a") as f:
                f.write(json.dumps(rec) + "\n")

    cw = None
    if getattr(args, "class_weights", None):
        cw = torch.tensor([float(v) for v in args.class_weights.split(",")], dtype=torch.float32)
        if cw.numel() != num_classes:
            raise ValueError("--class-weights needs %d comma-separated values" % num_classes)
    trainer = Trainer(model, optimizer, train_loader, valid_loader, device, class_weights=cw,
                      on_epoch_end=on_epoch_end, shard_eval=getattr(args, "shard_eval", False))
    trainer.profile_timeline = getattr(args, "profile", None) == "timeline"
    if getattr(args, "eval_only", False):        # score a checkpoint: one validation pass, no training
        test_loss, test_acc = trainer._evaluate()
        print("[Info] Evaluation: test loss: {}, test acc: {}.".format(test_loss, test_acc), flush=True)
        trainer.history.append(dict(epoch=start_epoch - 1, test_loss=test_loss.average, test_acc=test_acc.accuracy))
        return trainer
    trainer.fit(args.epochs, start_epoch=start_epoch)
    return trainer
