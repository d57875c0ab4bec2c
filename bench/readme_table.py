"""BASELINE config #5: the reference README's scaling table (Readme.md:133-138) on synthetic data.

The README reports whole-run wall-clock for 5 epochs of COIL-100 (5760 train / 1440 validation images
of 128x128, 3 classes), mini-batch 16/32/64/96 per rank, on 1 vs 2 workers -- data loading and a full
validation pass on every rank each epoch included.  This script reproduces that protocol through the
real entry points:

  native     ``distributed_vgg_f_b200.train.manage_training`` (what ``python -m distributed_vgg_f_b200``
             runs): 5 epochs, fused input pipeline, native engine.
  reference  the UNMODIFIED ``baseline/_ref/distributedVggf.py::manage_training`` (its DataManager with
             PIL transforms, its Trainer, DDP over NCCL).  Its loader does ~60-400 img/s, so ONE epoch
             is timed and the 5-epoch figure is 5x that (flagged in the output).

    python bench/readme_table.py --impl native            # 1 process
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 bench/readme_table.py --impl native
Each process loops over the mini-batch sizes; rank 0 appends JSON lines to gpurun_out/readme_table_<impl>.jsonl.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--mbs", default="16,32,64,96")
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--train-per-class", type=int, default=1920)     # 3 x 1920 = 5760 train, 3 x 480 = 1440 val
    ap.add_argument("--root", default="/tmp/b200_coil_like")
    ap.add_argument("--coil-train", type=int, default=5760, help="train images per epoch of the README's set")
    a = ap.parse_args()
    a.root = "%s_%d" % (a.root, a.train_per_class)
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0 and not os.path.isdir(os.path.join(a.root, "TrainData")):
        from distributed_vgg_f_b200.data.synthetic import make_synthetic_imagefolder
        make_synthetic_imagefolder(a.root + ".tmp", train_per_class=a.train_per_class,
                                   val_per_class=a.train_per_class // 4, size=128, seed=3)
        os.replace(a.root + ".tmp", a.root)
    if world > 1:
        dist.barrier()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = os.path.join(ROOT, "gpurun_out", "readme_table_%s.jsonl" % a.impl)

    if a.impl == "reference":
        import torchvision
        ref_dir = os.path.join(ROOT, "baseline", "_ref")
        sys.path.insert(0, ref_dir)
        orig = torchvision.models.vgg16
        torchvision.models.vgg16 = lambda pretrained=False, **kw: orig(weights=None, **kw)
        import distributedUtil as dstUt
        import distributedVggf as ref
        dstUt.BACKEND = "nccl"
        # the reference says torch.device("cuda"): with torch.cuda.set_device(local) above that
        # is this process's GPU.  At world size 1 it wraps the model in nn.DataParallel over ALL visible
        # GPUs (distributedVggf.py:227): launch with CUDA_VISIBLE_DEVICES=0 for the 1-GPU column.

    for mb in [int(v) for v in a.mbs.split(",")]:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        build_s = 0.0
        if a.impl == "reference":            # model construction is paid once per run, not per image
            tb = time.perf_counter()
            ref.vgg_funnel_model(3).to(dev)
            torch.cuda.synchronize(dev)
            build_s = time.perf_counter() - tb
        t0 = time.perf_counter()
        if a.impl == "native":
            from distributed_vgg_f_b200.cli import build_parser
            from distributed_vgg_f_b200.train import manage_training
            args = build_parser().parse_args(["-iu", "tcp://127.0.0.1:1", "-rn", str(rank), "-ws", str(world), "-rd", a.root,
                                              "-ep", str(a.epochs), "-lr", "0.00001", "-mb", str(mb)])
            trainer = manage_training(args)
            epochs_timed, hist = a.epochs, trainer.history
            acc = hist[-1]["test_acc"]
        else:
            ns = argparse.Namespace(init_url="", rank=rank, world_size=world, root_dir=a.root, epochs=1, no_cuda=False,
                                    learning_rate=1e-5, mini_batch=mb)
            ref.manage_training(ns)                               # distributedVggf.py:200-236, unmodified
            epochs_timed, acc = 1, None
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - t0
        t = torch.tensor([wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            n_train = 3 * a.train_per_class
            scale = (5.0 / epochs_timed) * (a.coil_train / n_train)
            rec = {"impl": a.impl, "mini_batch": mb, "world": world, "epochs_timed": epochs_timed,
                   "wall_s_timed": round(float(t), 3),
                   "wall_s_5_epochs": round(build_s + max(float(t) - build_s, 0.0) * scale, 3),
                   "extrapolated": scale != 1.0, "scale": scale, "model_build_s": round(build_s, 3),
                   "train_images_per_epoch": n_train,
                   "val_images_per_epoch": 3 * (a.train_per_class // 4), "final_val_acc": acc,
                   "includes": "model build, data loading, training, full validation pass on every rank per epoch"}
            with open(out, "a") as f:
                f.write(json.dumps(rec) + "\n")
            print(json.dumps(rec), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
