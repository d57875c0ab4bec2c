"""BASELINE config #4: gradient all-reduce bandwidth sweep, 1 KB - 1 GB, fused P2P / NVLS kernel vs
NCCL (bf16 on the wire in both cases; the NCCL arm is DDP's bf16_compress_hook recipe: a cast+scale
kernel, ncclAllReduce, a cast-back kernel).  Launch with torchrun --nproc-per-node N.

For every size: time per all-reduce (CUDA events on the stream, max over ranks), algorithm bandwidth
(fp32 gradient bytes / time), bus bandwidth on the bf16 wire (2(ws-1)/ws * bf16 bytes / time) and its
fraction of the 770 GB/s measured peer-copy rate / the 900 GB/s nominal NVLink rate per direction.
Writes gpurun_out/allreduce_sweep_wsN.json on rank 0.
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from distributed_vgg_f_b200.parallel.symm import SymmetricArena
    max_elems = int(os.environ.get("AR_MAX_ELEMS", str(256 * 1024 * 1024)))       # 1 GiB of fp32 gradient
    arena = SymmetricArena(max_elems, dev)
    grad = torch.randn(max_elems, device=dev)
    out = torch.empty_like(grad)
    tmp16 = torch.empty(max_elems, dtype=torch.bfloat16, device=dev)
    sizes = [256 * 4 ** i for i in range(0, 11)]          # fp32 elements: 1 KB ... 1 GB
    sizes = [s for s in sizes if s <= max_elems]
    algos = ["oneshot", "twoshot"] + (["nvls"] if arena.has_multicast else [])
    results = []

    def timed(fn, iters):
        for _ in range(3):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def nccl_arm(n):
        torch.mul(grad[:n], 1.0 / world, out=out[:n])          # DDP divides first
        tmp16[:n].copy_(out[:n])                                # compress
        dist.all_reduce(tmp16[:n])
        out[:n].copy_(tmp16[:n])                                # decompress

    cta_list = [int(c) for c in os.environ.get("AR_CTAS", "8,16,32,48,64,96,128").split(",")]
    for n in sizes:
        iters = 50 if n <= (1 << 22) else (20 if n <= (1 << 26) else 8)
        row = {"fp32_bytes": n * 4, "wire_bytes": n * 2}
        for algo in algos:
            if algo == "oneshot" and n > (1 << 22):
                continue
            for ctas in (cta_list if n >= (1 << 18) else (cta_list[0], cta_list[-1])):
                # (a) what a conv bucket does: pack fp32 -> wire, reduce, Adam reads the wire (no unpack)
                fn = (lambda: arena.allreduce(grad, out, 0, n, algo=algo, slot=1, max_ctas=ctas)) if algo == "oneshot" \
                    else (lambda: arena.allreduce(grad, None, 0, n, algo=algo, slot=1, max_ctas=ctas))
                ms = timed(fn, iters)
                key = "%s_c%d" % (algo, ctas)
                row[key + "_us"] = ms * 1e3
                row[key + "_busGBs"] = 2 * (world - 1) / world * n * 2 / (ms * 1e-3) / 1e9
                if algo != "oneshot":
                    # (b) what an FC-weight bucket does: the wgrad GEMM already wrote the wire (no pack, no unpack)
                    ms = timed(lambda: arena.allreduce(None, None, 0, n, algo=algo, slot=2, max_ctas=ctas), iters)
                    row[key + "_wireonly_us"] = ms * 1e3
                    row[key + "_wireonly_busGBs"] = 2 * (world - 1) / world * n * 2 / (ms * 1e-3) / 1e9
        ms = timed(lambda: nccl_arm(n), iters)
        row["nccl_bf16_hook_us"] = ms * 1e3
        row["nccl_busGBs"] = 2 * (world - 1) / world * n * 2 / (ms * 1e-3) / 1e9
        ms = timed(lambda: dist.all_reduce(tmp16[:n]), iters)
        row["nccl_allreduce_only_us"] = ms * 1e3
        row["nccl_only_busGBs"] = 2 * (world - 1) / world * n * 2 / (ms * 1e-3) / 1e9
        results.append(row)
        if rank == 0:
            best = max((v for k, v in row.items() if k.endswith("_busGBs") and not k.startswith("nccl")), default=0)
            print("n=%11d B  fused best bus %.1f GB/s (%.0f%% of 770, %.0f%% of 900) | nccl hook %.1f GB/s, nccl only %.1f GB/s | %s"
                  % (n * 4, best, 100 * best / 770, 100 * best / 900, row["nccl_busGBs"], row["nccl_only_busGBs"],
                     {k: round(v, 1) for k, v in row.items() if k.endswith("_us")}), flush=True)
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump({"world": world, "multicast": arena.has_multicast, "rows": results},
                  open("gpurun_out/allreduce_sweep_ws%d.json" % world, "w"), indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
