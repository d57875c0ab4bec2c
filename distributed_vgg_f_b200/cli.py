"""Command line -- flag-for-flag compatible with the reference.

Reference (distributedVggf.py:239-302): ``-iu/--init_url`` (required), ``-rn/--rank`` (required),
``-ws/--world_size`` (required), ``-rd/--root_dir``, ``-ep/--epochs`` (20), ``-nc/--no_cuda``,
``-lr/--learning_rate`` (1e-3), ``-mb/--mini_batch`` (16); the parser prints the namespace and, as
a side effect, joins the process group when ``-ws > 1``; ``__main__`` then calls
``manage_training``.  Usage is one invocation per rank, e.g. (Readme.md:43-47)

    python -m distributed_vgg_f_b200 -iu tcp://10.0.0.1:23456 -rn 0 -ws 2 -rd /data -ep 5 \
        -lr 0.00001 -mb 64

All eight flags, short and long names and defaults are preserved; ``-rd`` is actually required
(the reference crashes with a TypeError without it).  Everything else is additive.
Under ``torchrun`` the three required flags may be omitted: they default from
``RANK / WORLD_SIZE / MASTER_ADDR:MASTER_PORT``.
"""
from __future__ import annotations

import argparse
import os
import sys
from typing import Optional, Sequence

from .config import TRAIN


def build_parser() -> argparse.ArgumentParser:
    desc = ("Trains a VGG-F convolutional neural network using a distributed cluster of machines. "
            "Train and validation data must be stored in folders TrainData and ValidationData under "
            "some root directory. Train and validation folders must have N subdirectories, one per "
            "class. The name of each of these N subdirectories is the class label.")
    env_ws = os.environ.get("WORLD_SIZE")
    env_rank = os.environ.get("RANK")
    env_url = None
    if "MASTER_ADDR" in os.environ and "MASTER_PORT" in os.environ:
        env_url = "tcp://%s:%s" % (os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"])
    parser = argparse.ArgumentParser(prog="distributed_vgg_f_b200", description=desc)
    required = parser.add_argument_group("required arguments")
    required.add_argument("-iu", "--init_url", type=str, required=env_url is None, default=env_url,
                          help="Initialization URL specifying the protocol and connection with a "
                               "master node like tcp://192.168.0.10:23456")
    required.add_argument("-rn", "--rank", type=int, required=env_rank is None,
                          default=int(env_rank) if env_rank is not None else 0,
                          help="Rank of the current instance (0..K-1).")
    required.add_argument("-ws", "--world_size", type=int, required=env_ws is None,
                          default=int(env_ws) if env_ws is not None else None,
                          help="Number of instances participating in the job")
    required.add_argument("-rd", "--root_dir", type=str, default=None,
                          help="Root directory for the training and validation sets")
    parser.add_argument("-ep", "--epochs", type=int, default=TRAIN.epochs,
                        help="Number of epochs. The default value is {}".format(TRAIN.epochs))
    parser.add_argument("-nc", "--no_cuda", action="store_true",
                        help="Flag to use the CPU in this instance. By default the program tries "
                             "to run on a GPU")
    parser.add_argument("-lr", "--learning_rate", type=float, default=TRAIN.learning_rate,
                        help="Selected learning rate. The default value is {}".format(
                            TRAIN.learning_rate))
    parser.add_argument("-mb", "--mini_batch", type=int, default=TRAIN.mini_batch,
                        help="Mini-batch size. The default value is {}".format(TRAIN.mini_batch))

    ext = parser.add_argument_group("extensions (not in the reference)")
    ext.add_argument("--model", default="vggf",
                     choices=["vggf", "vgg16", "vgg11", "vgg13", "vgg19", "vggf11", "vggf13", "vggf19",
                              "vggf-tiny", "vggf-mini"],
                     help="vggf = the reference's network (VGG-16 + funnel head); vggNN = torchvision's plain "
                          "VGG-NN; vggfNN = VGG-NN + funnel head; -tiny / -mini are test-sized")
    ext.add_argument("--num-classes", type=int, default=None,
                     help="override the class count (default: number of class folders)")
    ext.add_argument("--engine", default="auto", choices=["auto", "native", "oracle"],
                     help="native = sm_100a kernels; oracle = torch ops (CPU / reference numerics)")
    ext.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"],
                     help="bf16: the sm_100a engine (bf16 activations / weights / wire, fp32 master weights, moments and "
                          "accumulators).  fp32: the reference's arithmetic on the torch-op engine (the native kernels "
                          "have no tf32 path yet)")
    ext.add_argument("--optimizer", default="adam", choices=["adam", "sgd"])
    ext.add_argument("--momentum", type=float, default=TRAIN.momentum)
    ext.add_argument("--lr-step", type=int, default=0,
                     help="StepLR period in epochs (0 = constant lr, as the reference)")
    ext.add_argument("--lr-gamma", type=float, default=TRAIN.decay_gamma)
    ext.add_argument("--backend", default=None, choices=[None, "gloo", "nccl"],
                     help="rendezvous backend (default: nccl on GPU, gloo on CPU)")
    ext.add_argument("--allreduce", default="auto",
                     choices=["auto", "oneshot", "twoshot", "nvls", "nccl"],
                     help="gradient all-reduce algorithm (nccl = library baseline)")
    ext.add_argument("--wire-dtype", default="bf16", choices=["bf16", "fp32"])
    ext.add_argument("--bucket-mb", type=float, default=32.0)
    ext.add_argument("--pipeline", default="auto", choices=["auto", "fused", "reference"],
                     help="input pipeline: fused GPU augment over a decoded uint8 cache, the reference's "
                          "per-sample PIL transforms, or auto (fused when the split has one image size "
                          "and fits in host memory)")
    ext.add_argument("--pretrained", default=None,
                     help="path to a torchvision vgg16 state dict (offline pretrained=True)")
    ext.add_argument("--save", default=None, help="checkpoint path written by rank 0 every epoch")
    ext.add_argument("--resume", default=None, help="checkpoint path to resume from")
    ext.add_argument("--seed", type=int, default=0)
    ext.add_argument("--synthetic", type=int, default=0,
                     help="generate a synthetic ImageFolder with this many train images per class "
                          "under --root_dir if it does not exist")
    ext.add_argument("--socket-ifname", default=None,
                     help="network interface for the rendezvous / gloo / NCCL bootstrap sockets (the reference "
                          "hard-codes GLOO_SOCKET_IFNAME=eth0, distributedVggf.py:296; default: auto-detect)")
    ext.add_argument("--eval-only", action="store_true",
                     help="one validation pass (typically with --resume <checkpoint>) and exit")
    ext.add_argument("--zero1", default="auto", choices=["auto", "on", "off"],
                     help="shard the Adam state of the FC weights (88 %% of the parameters) across the ranks of one "
                          "NVLink domain: reduce-scatter + optimizer + all-gather of the new bf16 weights in ONE "
                          "kernel per bucket (auto: from 4 ranks up)")
    ext.add_argument("--reference-order", action="store_true",
                     help="replay the reference's identical-shuffle-every-epoch behaviour")
    ext.add_argument("--shard-eval", action="store_true",
                     help="shard validation across ranks and all-reduce the metrics")
    ext.add_argument("--class-weights", default=None,
                     help="comma-separated per-class loss weights (the reference's commented-out "
                          "CLASS_OPTIM_WEIGHTS, distributedUtil.py:27-28)")
    ext.add_argument("--profile", default=None, choices=[None, "events", "nvtx", "timeline"],
                     help="events: per-phase CUDA-event timings each epoch; nvtx: emit NVTX ranges; timeline: print, for "
                          "the last training step of every epoch, when each layer's kernels and each bucket's reduction / "
                          "update finished (CUDA events, ms since the start of the step; native engine)")
    ext.add_argument("--log-jsonl", default=None, help="append one JSON record per epoch")
    return parser


def parse_command_line(argv: Optional[Sequence[str]] = None, init: bool = True):
    """Parse, print the namespace (as the reference does) and join the process group."""
    args = build_parser().parse_args(argv)
    if args.root_dir is None:
        build_parser().error("the following arguments are required: -rd/--root_dir")
    print(args)
    if args.socket_ifname:
        os.environ["GLOO_SOCKET_IFNAME"] = args.socket_ifname
        os.environ.setdefault("NCCL_SOCKET_IFNAME", args.socket_ifname)
    if init and args.world_size > 1:
        from .parallel.process_group import init_distributed, pick_device

        device = pick_device(args.rank, args.no_cuda)
        init_distributed(args.init_url, args.rank, args.world_size, device, args.backend)
    return args


def main(argv: Optional[Sequence[str]] = None) -> int:
    from .train import manage_training

    args = parse_command_line(argv)
    from .parallel.process_group import shutdown
    from .parallel.watchdog import start_abort_watch

    watch = start_abort_watch(args.rank)          # None for single-process runs
    try:
        manage_training(args)
    except BaseException as e:                    # tell the peers before going down (they may be
        if watch is not None:                     # blocked in a collective this rank will never join)
            watch.signal("%s: %s" % (type(e).__name__, e))
            watch.stop()
        shutdown(graceful=False)
        raise
    if watch is not None:
        watch.stop()
    shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
