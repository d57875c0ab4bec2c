"""Debug aid: run three engines side by side on 2 ranks (replicated Adam x2, fused ZeRO-1) and print, per
step and per bucket, how many bf16 weights differ.  torchrun --nproc-per-node 2 bench/debug_zero1.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
from distributed_vgg_f_b200.engine.native_engine import NativeEngine
from distributed_vgg_f_b200.models.vggf import build_oracle, vggf_mini_spec
spec = vggf_mini_spec(3)
init = build_oracle(spec, seed=0).state_dict()
algo = os.environ.get("ALGO", "twoshot")
kw = dict(device=dev, batch=4, lr=1e-3, seed=0, input_hw=64, init_state=init, allreduce=algo, bucket_mb=0.25)
engs = {"ref": NativeEngine(spec, **kw), "ref2": NativeEngine(spec, **kw), "z": NativeEngine(spec, zero1=True, **kw)}
for e in engs.values():
    e.train_dropout = False
g = torch.Generator().manual_seed(7 + rank)
serial = os.environ.get("SERIAL", "0") == "1"
for step in range(4):
    x = torch.randn(4, 3, 64, 64, generator=g).to(torch.bfloat16).float()
    y = torch.randint(0, 3, (4,), generator=g)
    for e in engs.values():
        e.train_step((x, y))
        if serial:
            e.sync(); dist.barrier()
    for e in engs.values():
        e.sync()
    if rank == 0:
        for a, b in (("ref", "ref2"), ("ref", "z")):
            d = (engs[a].w16.float() - engs[b].w16.float()).abs()
            per = ["%d%s:%.3f" % (bi, "z" if bi in engs["z"]._zero1_buckets else "", float((d[bk.start:bk.end] > 0).float().mean()))
                   for bi, bk in enumerate(engs["z"].plan.buckets)]
            print("step %d %s-%s frac %.4f max %.5f | %s" % (step, a, b, float((d > 0).float().mean()), float(d.max()), " ".join(per)), flush=True)
    # replicas of z identical?
    w = engs["z"].w16.clone(); dist.broadcast(w, src=0)
    if not torch.equal(w, engs["z"].w16):
        print("rank %d: z replicas differ at step %d: %d elements" % (rank, step, int((w != engs["z"].w16).sum())), flush=True)
dist.barrier(); dist.destroy_process_group()
