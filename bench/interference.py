"""What slows the convolution kernels of backward while the comm stream works?  (1 GPU.)

A conv layer's dgrad + wgrad loop runs on stream A; stream B (high priority, like the engine's comm
stream) runs one of:
  none            nothing                                            -> baseline
  adam{1,4,8}     the fused Adam kernel, N CTAs per SM               -> what the engine does
  d2d_engine      cudaMemcpyAsync device->device                     -> memory traffic, NO SM footprint
  d2d_kernel      a torch elementwise copy kernel                    -> memory traffic + SM footprint
  spin            thin CTAs that only spin (no memory traffic)       -> SM footprint only
Reported: conv time per iteration (ms) and its ratio to the baseline.  Writes gpurun_out/interference.json.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from distributed_vgg_f_b200 import ops

dev = "cuda"
B = 64
C = ops.require()
LAYERS = {"features.21 (512->512, 28x28)": (28, 512, 512), "features.7 (128->128, 112x112)": (112, 128, 128),
          "features.12 (256->256, 56x56)": (56, 256, 256)}
n = 64 * 1024 * 1024                      # 64 M parameters of optimizer state: ~1.8 GB of traffic per pass
p = torch.zeros(n, device=dev); m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
g16 = torch.zeros(n, dtype=torch.bfloat16, device=dev); w16 = torch.zeros(n, dtype=torch.bfloat16, device=dev)
src = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev); dst = torch.empty_like(src)
side = torch.cuda.Stream(priority=-1)
results = {}


PROBE = {"fma": 0, "ld": 1, "st": 2, "ldst": 3, "mufu": 4, "sleep": 5}
KINDS = os.environ.get("KINDS", "none,adam,d2d_engine,d2d_kernel").split(",")
pbuf = torch.zeros(64 * 1024 * 1024, device=dev)          # 256 MB
sink = torch.zeros(1, device=dev)
NSM = torch.cuda.get_device_properties(0).multi_processor_count


def background(kind):
    if kind == "none":
        return
    if kind.startswith("probe_"):          # probe_<what>_<ctas per SM> e.g. probe_ld_1
        _, what, per = kind.split("_")
        mode = PROBE[what]
        reps = {0: 6, 4: 6, 5: 12}.get(mode, 2)
        C.probe_background(pbuf, mode, NSM * int(per), reps, sink)
        return
    if kind.startswith("adam"):
        ops.adam_step(p, m, v, g16=g16, shadow=w16, lr=1e-5, step=3)
    elif kind == "d2d_engine":
        dst.copy_(src, non_blocking=True)          # same-device copy of a contiguous tensor: cudaMemcpyAsync D2D
    elif kind == "d2d_kernel":
        torch.add(src[: n * 2].view(torch.bfloat16), 0, out=dst[: n * 2].view(torch.bfloat16))
    elif kind == "spin":
        torch.cuda._sleep(int(4e6))                # 1 CTA; SM-footprint control is weak here, kept for reference


for lname, (h, cin, cout) in LAYERS.items():
    x = torch.randn(B, h, h, cin, device=dev).bfloat16()
    w = (torch.randn(cout, 3, 3, cin, device=dev) * 0.05).bfloat16()
    dz = torch.randn(B, h, h, cout, device=dev).bfloat16()
    dx = torch.empty_like(x); dw = torch.zeros(cout, 3, 3, cin, device=dev)

    def conv_iter():
        C.conv_wgrad(dz, x, dw, 1.0, 0, 0)
        C.conv_dgrad(dz, w, x, dx, None, 0)

    base = None
    # the optimizer's CTAs-per-SM cap is read once per process (B200_ADAM_CTAS_PER_SM): run this
    # script once per value
    for kind in KINDS:
        iters = 12
        for _ in range(3):
            conv_iter()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(side):
            for _ in range(6):
                background(kind)
        e0.record()
        for _ in range(iters):
            conv_iter()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        base = base or ms
        results.setdefault(lname, {})[kind] = {"ms": round(ms, 4), "ratio": round(ms / base, 3)}
        print("%-34s %-11s %.3f ms  x%.3f" % (lname, kind + (os.environ.get("B200_ADAM_CTAS_PER_SM", "") if kind == "adam" else ""), ms, ms / base), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(results, open("gpurun_out/interference%s.json" % os.environ.get("TAG", ""), "w"), indent=1)
