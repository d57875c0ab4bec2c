"""Launch a handful of representative conv kernels once each (for ncu capture)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_vgg_f_b200 import ops
C = ops.require()
dev = "cuda"
B = 64
def run(h, cin, cout, what):
    x = torch.randn(B, h, h, cin, device=dev).bfloat16()
    w = (torch.randn(cout, 3, 3, cin, device=dev) * 0.05).bfloat16()
    bias = torch.randn(cout, device=dev)
    y = torch.empty(B, h, h, cout, dtype=torch.bfloat16, device=dev)
    dz = torch.randn(B, h, h, cout, device=dev).bfloat16()
    dx = torch.empty_like(x)
    dw = torch.zeros(cout, 3, 3, cin, device=dev)
    for _ in range(2):
        if "f" in what: C.conv_fprop(x, w, bias, y, True, 0)
        if "d" in what: C.conv_dgrad(dz, w, x, dx, None, 0)
        if "w" in what: C.conv_wgrad(dz, x, dw, 1.0, 0, 0)
    torch.cuda.synchronize()
run(224, 64, 64, "fdw")
run(56, 256, 256, "fdw")
run(28, 512, 512, "f")
run(14, 512, 512, "f")
# first conv (fused, no im2col)
x4 = torch.randn(B, 224, 224, 4, device=dev).bfloat16()
w0 = torch.randn(64, 64, device=dev).bfloat16()
y0 = torch.empty(B, 224, 224, 64, dtype=torch.bfloat16, device=dev)
dw0 = torch.zeros(64, 64, device=dev)
for _ in range(2):
    C.conv0_fprop(x4, w0, torch.zeros(64, device=dev), y0)
    C.conv0_wgrad(y0, x4, dw0)
torch.cuda.synchronize()
