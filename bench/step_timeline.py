"""Step timeline without nsys: CUDA events after every layer (compute stream) and around every
gradient all-reduce / optimizer launch (comm stream), as milliseconds since the start of the step.

    python bench/step_timeline.py                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 bench/step_timeline.py

Shows which bucket's reduction runs under which layer's backward and how much of the gradient
exchange is exposed after the last weight gradient.  Median over --steps steps; rank 0 writes
gpurun_out/timeline_n<N>.json and prints a table.
"""
import argparse
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=7)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--comm-ctas", type=int, default=48)
    ap.add_argument("--model", default="vggf")
    ap.add_argument("--num-classes", type=int, default=3)
    ap.add_argument("--tag", default="")
    ap.add_argument("--zero1", default="auto", choices=["auto", "on", "off"])
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from bench import make_host_batches
    from distributed_vgg_f_b200.data.loader import FusedBatch
    from distributed_vgg_f_b200.engine.native_engine import NativeEngine
    from distributed_vgg_f_b200.models.vggf import get_spec

    eng = NativeEngine(get_spec(a.model, a.num_classes), device=dev, batch=a.batch, lr=1e-5, seed=0,
                       comm_ctas=a.comm_ctas, zero1={"auto": "auto", "on": True, "off": False}[a.zero1] if world > 1 else False)
    host = make_host_batches(torch, 4, a.batch, a.num_classes, seed=rank + 1, pin=False)
    batches = [FusedBatch(b.images_u8.to(dev), b.params.to(dev), b.labels.to(dev), b.resized_hw, None) for b in host]
    for k in range(a.warmup):
        eng.train_step(batches[k % 4])
    torch.cuda.synchronize(dev)
    runs = []
    for k in range(a.steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        eng.timeline(True)
        eng.train_step(batches[k % 4])
        runs.append(eng.timeline_report())
    eng.timeline(False)
    names = [(n, lane) for n, lane, _ in runs[0]]
    med = [statistics.median(r[i][2] for r in runs) for i in range(len(names))]
    rows = [{"name": n, "lane": lane, "ms": round(m, 4)} for (n, lane), m in zip(names, med)]
    # every rank's step end -> max over ranks; and the per-rank skew: when did each rank's LAST weight gradient
    # finish (its own speed) and when did its step end (after the last bucket's reduction, which needs everybody)
    t_end = torch.tensor([rows[-1]["ms"]], device=dev)
    last_w = max(r["ms"] for r in rows if r["lane"] == "compute" and r["name"].startswith("bwd "))
    per_rank = torch.zeros(world, 2, device=dev)
    per_rank[rank, 0], per_rank[rank, 1] = last_w, rows[-1]["ms"]
    if world > 1:
        dist.all_reduce(t_end, op=dist.ReduceOp.MAX)
        dist.all_reduce(per_rank)
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        out = "gpurun_out/timeline_n%d%s.json" % (world, a.tag)
        json.dump({"world": world, "comm_ctas": a.comm_ctas, "rows": rows, "step_end_max_over_ranks_ms": float(t_end),
                   "per_rank_last_wgrad_ms": [round(float(v), 4) for v in per_rank[:, 0]],
                   "per_rank_step_end_ms": [round(float(v), 4) for v in per_rank[:, 1]]},
                  open(out, "w"), indent=1)
        comp = [r for r in rows if r["lane"] == "compute"]
        comm = [r for r in rows if r["lane"] != "compute"]
        print("compute stream (ms since step start; delta):")
        prev = 0.0
        for r in comp:
            print("  %-28s %8.3f  +%.3f" % (r["name"], r["ms"], r["ms"] - prev))
            prev = r["ms"]
        print("side streams (comm = reductions, opt = optimizer):")
        for r in comm:
            print("  %-5s %-28s %8.3f" % (r["lane"], r["name"], r["ms"]))
        if world > 1:
            lw = [float(v) for v in per_rank[:, 0]]
            print("per-rank end of backward (ms): " + " ".join("%.3f" % v for v in lw)
                  + "  -> spread %.3f ms: the last bucket's reduction cannot finish before the slowest rank's" % (max(lw) - min(lw)))
            print("per-rank step end (ms):        " + " ".join("%.3f" % float(v) for v in per_rank[:, 1]))
        last_w = max(r["ms"] for r in comp if r["name"].startswith("bwd "))
        print("last wgrad enqueued-done at %.3f ms, step end (joined comm) at %.3f ms -> exposed tail %.3f ms"
              % (last_w, comp[-1]["ms"], comp[-1]["ms"] - last_w))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
