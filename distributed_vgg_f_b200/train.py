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

    if getattr(args, "synthetic", 0) and not os.path.isdir(os.path.join(args.root_dir, "TrainData")):
        if args.rank == 0:
            from .data.synthetic import make_synthetic_imagefolder
            make_synthetic_imagefolder(args.root_dir, train_per_class=args.synthetic,
                                       val_per_class=max(args.synthetic // 4, 1), seed=args.seed)
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
                             zero1=getattr(args, "zero1", False))
        optimizer = None            # fused into the engine
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
        if args.log_jsonl and args.rank == 0:
            rec = dict(trainer.history[-1], time=time.time(), world_size=args.world_size)
            if hasattr(model, "phase_times"):
                rec["phase_ms"] = model.phase_times()
            with open(args.log_jsonl, "a") as f:
                f.write(json.dumps(rec) + "\n")

    cw = None
    if getattr(args, "class_weights", None):
        cw = torch.tensor([float(v) for v in args.class_weights.split(",")], dtype=torch.float32)
        if cw.numel() != num_classes:
            raise ValueError("--class-weights needs %d comma-separated values" % num_classes)
    trainer = Trainer(model, optimizer, train_loader, valid_loader, device, class_weights=cw,
                      on_epoch_end=on_epoch_end, shard_eval=getattr(args, "shard_eval", False))
    if getattr(args, "eval_only", False):        # score a checkpoint: one validation pass, no training
        test_loss, test_acc = trainer._evaluate()
        print("[Info] Evaluation: test loss: {}, test acc: {}.".format(test_loss, test_acc), flush=True)
        trainer.history.append(dict(epoch=start_epoch - 1, test_loss=test_loss.average, test_acc=test_acc.accuracy))
        return trainer
    trainer.fit(args.epochs, start_epoch=start_epoch)
    return trainer
