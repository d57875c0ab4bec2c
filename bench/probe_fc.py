"""Probe: FC-dgrad-shaped GEMMs with tiny K / BN=32 / padded ld, engine-vs-emulated intermediates."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_vgg_f_b200 import ops

dev = "cuda"
torch.manual_seed(0)

def rel(a, b):
    return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12))

print("== gemm a_mn (W stored [K][M]) x B [N][K], store_t")
for (M, N, K, ldb_pad) in [(64, 4, 3, 8), (64, 4, 8, 8), (64, 4, 64, 64), (64, 64, 3, 8), (64, 32, 3, 8),
                           (256, 4, 64, 64), (256, 4, 256, 256), (3136, 4, 256, 256), (64, 8, 3, 8), (128, 4, 3, 8)]:
    A = torch.randn(K, M, device=dev).bfloat16()
    Bfull = torch.zeros(N, ldb_pad, device=dev).bfloat16()
    Bfull[:, :K] = torch.randn(N, K, device=dev).bfloat16()
    B = Bfull[:, :K] if ldb_pad != K else Bfull
    out = torch.zeros(N, M, device=dev)
    ops.gemm(A, Bfull, out, M=M, N=N, K=K, a_mn=True, epi="f32_store_t", ldo=M)
    ref = Bfull[:, :K].float() @ A.float()
    print("M=%d N=%d K=%d ldb=%d  rel=%.2e" % (M, N, K, ldb_pad, rel(out, ref)))

print("== gemm kk (fwd form) small N")
for (M, N, K) in [(64, 4, 3136), (256, 4, 256), (3, 4, 64), (64, 4, 256)]:
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    out = torch.zeros(N, M, device=dev)
    ops.gemm(A, B, out, M=M, N=N, K=K, epi="f32_atomic_t", ksplit=2, ldo=M)
    print("M=%d N=%d K=%d rel=%.2e" % (M, N, K, rel(out, B.float() @ A.float().t())))

print("== engine vs emulated intermediates")
from distributed_vgg_f_b200.engine.native_engine import NativeEngine
from distributed_vgg_f_b200.models.vggf import build_oracle, vggf_mini_spec
from distributed_vgg_f_b200.ops import ref as R
spec = vggf_mini_spec(3)
oracle = build_oracle(spec, seed=0)
with torch.no_grad():
    for p in oracle.parameters():
        if p.dim() > 1:
            p.copy_(p.to(torch.bfloat16).float())
eng = NativeEngine(spec, device=torch.device(dev), batch=4, lr=1e-3, seed=0, input_hw=64, init_state=oracle.state_dict())
eng.train_dropout = False
eng.apply_updates = False
x = torch.randn(4, 3, 64, 64, device=dev).bfloat16().float()
y = torch.randint(0, 3, (4,), device=dev)
state = {k: v.detach().to(dev) for k, v in oracle.state_dict().items()}
logits, loss, grads, inter = R.emulated_step(spec, state, x, y, return_intermediates=True)
eng.train_step((x, y)); eng.sync()
print("logits rel", rel(eng.logits[:4], logits))
for i in range(len(spec.fcs) - 1, -1, -1):
    f = spec.fcs[i]
    print("fc_dz[%d] (%s) rel=%.2e" % (i, f.name, rel(eng.fc_dz[i][:4, :f.fout], inter["fc_dz"][i])))
for i, f in enumerate(spec.fcs[:-1]):
    print("fc_y[%d] rel=%.2e" % (i, rel(eng.fc_y[i][:4], inter["fc_y"][i])))
print("dfeat rel=%.2e" % rel(eng.dfeat[:4].permute(0, 3, 1, 2), inter["dfeat"]))
