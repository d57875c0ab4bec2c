"""Full-scale numerics: the native bf16 engine against the UNMODIFIED reference's own training loop
in fp32 (baseline/_ref/distributedVggf.py:158-175 ``Trainer.__train``) -- full VGG-F (136 M parameters),
224 x 224 inputs, batch 64, the same initial weights and the same 50 batches.

Dropout is switched off on both sides (``p = 0`` on the reference model's nn.Dropout modules, a
configuration of the model object, not of the reference's code): the two RNG streams cannot be made
equal, and the comparison is about the arithmetic.

What is asserted (numbers measured on a B200 are recorded in profiles/r2_fullscale_parity.json):
  * step-1 gradients, tensor by tensor, against fp32 autograd of the reference model (TF32 off):
    cosine similarity and relative L2 error (>= 0.99 for the classifier and the last conv block, >= 0.97 down
    to the first convolution: see the comment at the assertions for why it decays with depth);
  * the per-step training loss of 50 optimisation steps against the reference's loop from the same
    weights (Adam, lr 1e-5 as in the reference's README);
  * both runs learn the synthetic classes.
"""
import json
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
DEV = "cuda"
B, HW, STEPS, C = 64, 224, 50, 3


def _batches(n):
    """Learnable synthetic batches: class-dependent low-frequency patterns + noise (data/synthetic.py),
    bilinearly resized to the network input and normalised like the reference's transforms."""
    import torch.nn.functional as F

    from distributed_vgg_f_b200.config import DATA
    from distributed_vgg_f_b200.data.synthetic import synthetic_uint8_batch

    mean = torch.tensor(DATA.mean).view(1, 3, 1, 1)
    std = torch.tensor(DATA.std).view(1, 3, 1, 1)
    out = []
    for i in range(n):
        imgs, labels = synthetic_uint8_batch(B, 128, C, seed=1000 + i)
        x = torch.from_numpy(imgs).permute(0, 3, 1, 2).float() / 255.0
        x = F.interpolate(x, size=(HW, HW), mode="bilinear", align_corners=False)
        x = ((x - mean) / std).to(torch.bfloat16).float()        # values both sides represent exactly
        out.append((x, torch.from_numpy(labels)))
    return out


@pytest.fixture(scope="module")
def ref_module():
    if not os.path.exists(os.path.join(REF_DIR, "distributedVggf.py")):
        pytest.skip("baseline/_ref not installed (python baseline/install_ref.py)")
    import torchvision

    orig = torchvision.models.vgg16
    torchvision.models.vgg16 = lambda pretrained=False, **kw: orig(weights=None, **kw)
    sys.path.insert(0, REF_DIR)
    try:
        import distributedVggf as mod
        yield mod
    finally:
        torchvision.models.vgg16 = orig
        sys.path.remove(REF_DIR)


def test_fullscale_loss_curve_and_gradients_match_reference_trainer(ref_module):
    from distributed_vgg_f_b200.engine.native_engine import NativeEngine
    from distributed_vgg_f_b200.models import layout as L
    from distributed_vgg_f_b200.models.vggf import vggf_spec

    old_tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False          # a true fp32 reference
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(0)
        model = ref_module.vgg_funnel_model(C)                   # distributedVggf.py:35-59
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        init = {k: v.detach().clone() for k, v in model.state_dict().items()}
        model.to(DEV)
        spec = vggf_spec(C)
        eng = NativeEngine(spec, device=torch.device(DEV), batch=B, lr=1e-5, seed=0, input_hw=HW,
                           init_state=init, distributed=False)
        eng.train_dropout = False
        data = _batches(STEPS)

        # ---- step-1 gradients -----------------------------------------------------------------
        x0, y0 = data[0]
        model.train()
        model.zero_grad()
        loss0 = torch.nn.functional.cross_entropy(model(x0.to(DEV)), y0.to(DEV))
        loss0.backward()
        eng.apply_updates = False
        eng.train_step((x0, y0))
        eng.sync()
        grads = {}
        for name, p in model.named_parameters():
            g_ref = p.grad.detach().float().flatten()
            g_nat = L.to_torch(spec, name, eng._view(eng.g32, name)).float().flatten()
            cos = float(torch.dot(g_ref, g_nat) / (g_ref.norm() * g_nat.norm() + 1e-30))
            rel = float((g_ref - g_nat).norm() / (g_ref.norm() + 1e-30))
            grads[name] = {"cos": round(cos, 5), "rel_l2": round(rel, 5)}
        native_loss0 = float(eng.loss_buf)
        eng.g32.zero_()
        eng.apply_updates = True
        eng.step_count = 0
        model.zero_grad(set_to_none=True)

        # ---- 50 steps: the reference's own loop vs the engine ----------------------------------
        opt = torch.optim.Adam(model.parameters(), lr=1e-5)      # distributedVggf.py:230, README lr
        ref_losses = []

        class Loader:                                            # one batch per "epoch": loss per step
            def __init__(self, b):
                self.b = b

            def __iter__(self):
                yield self.b

        trainer = ref_module.Trainer(model, opt, None, None, torch.device(DEV))
        for xb, yb in data:
            trainer.train_loader = Loader((xb, yb))
            avg, _ = trainer._Trainer__train()                   # forward, CE, zero_grad, backward, step
            ref_losses.append(avg.average)
        nat_losses = []
        for xb, yb in data:
            eng.train_step((xb, yb))
            nat_losses.append(float(eng.loss_buf))
        eng.sync()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old_tf32

    diffs = [abs(a - b) for a, b in zip(nat_losses, ref_losses)]
    report = {"batch": B, "input": HW, "steps": STEPS, "loss_step1": {"reference_fp32": float(loss0), "native_bf16": native_loss0},
              "ref_losses": [round(v, 5) for v in ref_losses], "native_losses": [round(v, 5) for v in nat_losses],
              "max_abs_loss_diff": max(diffs), "mean_abs_loss_diff": sum(diffs) / len(diffs),
              "grad_step1": grads, "min_cos": min(v["cos"] for v in grads.values()),
              "max_rel_l2": max(v["rel_l2"] for v in grads.values())}
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "fullscale_parity.json"), "w") as f:
            json.dump(report, f, indent=1)
    except OSError:
        pass
    print(json.dumps({k: report[k] for k in ("loss_step1", "max_abs_loss_diff", "mean_abs_loss_diff", "min_cos", "max_rel_l2")}))
    worst = sorted(grads.items(), key=lambda kv: kv[1]["cos"])[:4]
    # Gradients.  Measured on a B200 (profiles/r2_fullscale_parity.json): cosine 1.0000 at the head, 0.995 at
    # classifier.0 / features.28, falling smoothly to 0.978 at features.0 (relative L2 error 0.001 -> 0.21).
    # The growth with depth is ReLU / max-pool GATE FLIPS, not arithmetic error: a pre-activation within bf16
    # rounding distance of zero (~0.3 % of them per layer) is on one side of the gate here and on the other in
    # fp32, which moves a whole gradient entry, and every layer below inherits it (relative error ~ sqrt(flipped
    # fraction) per layer, accumulating in quadrature over the 13 gated layers).  With the gates pinned to the same
    # forward values the two backward passes agree to < 2 % everywhere (test_backward_matches_reference_on_same_
    # forward), and the loss curves below coincide -- the noise is unbiased.
    assert report["min_cos"] >= 0.97, "step-1 gradient cosine below 0.97: %s" % worst
    assert report["max_rel_l2"] <= 0.25, "step-1 gradient relative L2 error above 25%%: %s" % worst
    for name, v in grads.items():
        if name.startswith("classifier.") or name.startswith("features.28"):
            assert v["cos"] >= 0.99, (name, v)
    # loss curve: same starting loss (measured: 6e-6 apart), same trajectory (measured: mean 7e-4, max 3.6e-3)
    assert abs(native_loss0 - float(loss0)) <= 1e-3 * max(1.0, float(loss0))
    assert report["mean_abs_loss_diff"] <= 0.005 and report["max_abs_loss_diff"] <= 0.02, (ref_losses, nat_losses)
    # and both learn
    assert sum(ref_losses[-5:]) / 5 < 0.8 * ref_losses[0] and sum(nat_losses[-5:]) / 5 < 0.8 * nat_losses[0]
