"""Probe unaligned / strided SWIZZLE_128B operand views on the tensor core (see csrc/umma_probe.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_vgg_f_b200 import ops
C = ops.require()
dev = "cuda"
torch.manual_seed(0)

def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))

print("== mode 0: K-major A view, row(r) = A[(r//8)*pitch + r%8 + shift]")
A = torch.randn(256, 64, device=dev).bfloat16()
B = torch.randn(64, 64, device=dev).bfloat16()
r = torch.arange(128, device=dev)
for pitch in (8, 10, 12, 16):
    for shift in (0, 1, 2, 3, 5, 7, 8, 9):
        idx = (r // 8) * pitch + r % 8 + shift
        if int(idx.max()) >= 256:
            continue
        ref = A[idx].float() @ B.float().t()
        res = []
        for ubo in (True, False):
            out = torch.zeros(128, 64, device=dev)
            try:
                C.shift_probe(A, B, out, shift, pitch, ubo, 0)
                torch.cuda.synchronize()
                res.append("%.1e" % rel(out, ref))
            except Exception as e:
                res.append("ERR " + str(e)[:60])
        print("pitch=%2d shift=%d  base_offset:on=%s off=%s" % (pitch, shift, res[0], res[1]))

print("== mode 1: MN-major B view, K-row k -> KN[k + shift]")
KN = torch.randn(200, 64, device=dev).bfloat16()        # [K rows][N=64]
Ad = torch.randn(64, 64, device=dev).bfloat16()         # dense A, 64 valid rows (rows 64..127 of the tile are zero-filled)
for shift in (0, 1, 2, 3, 7, 8, 9, 33, 100):
    ref = Ad.float() @ KN[shift:shift + 64].float()      # [64][64]
    res = []
    for ubo in (True, False):
        out = torch.zeros(128, 64, device=dev)
        C.shift_probe(KN, Ad, out, shift, 8, ubo, 1)
        torch.cuda.synchronize()
        res.append("%.1e" % rel(out[:64], ref))
    print("shift=%3d  base_offset:on=%s off=%s" % (shift, res[0], res[1]))
