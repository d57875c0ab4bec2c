"""Parity with the UNMODIFIED reference (baseline/_ref/distributedVggf.py, vendored byte-for-byte by
baseline/install_ref.py) on CPU: same parameter names/shapes, same forward, same training step, and
checkpoints written by this framework load into the reference's model object (SURVEY D1, section 4
"DDP equivalence" tier at world size 1).  Skipped when baseline/_ref is not installed."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF_DIR, "distributedVggf.py")),
                                reason="baseline/_ref not installed (python baseline/install_ref.py)")


@pytest.fixture(scope="module")
def ref():
    import torchvision

    orig = torchvision.models.vgg16
    # pretrained=True would download ImageNet weights (no network): same architecture, random init
    torchvision.models.vgg16 = lambda pretrained=False, **kw: orig(weights=None, **kw)
    sys.path.insert(0, REF_DIR)
    try:
        import distributedVggf as mod
        yield mod
    finally:
        torchvision.models.vgg16 = orig
        sys.path.remove(REF_DIR)


@pytest.fixture(scope="module")
def pair(ref):
    """(reference model, our oracle module) holding the same weights."""
    from distributed_vgg_f_b200 import vgg_funnel_model

    torch.manual_seed(0)
    theirs = ref.vgg_funnel_model(3)                      # distributedVggf.py:35-59
    ours = vgg_funnel_model(3)
    ours.load_state_dict(theirs.state_dict())             # identical key set or this raises
    return theirs, ours


def test_state_dict_matches_reference_model(pair):
    theirs, ours = pair
    a, b = theirs.state_dict(), ours.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape for k in a)
    assert sum(p.numel() for p in ours.parameters()) == 136_359_747
    assert all(p.requires_grad for p in ours.parameters())     # nothing frozen, like the reference


def test_forward_matches_reference_model(pair):
    theirs, ours = pair
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ya, yb = theirs.eval()(x), ours.eval()(x)
    assert torch.allclose(ya, yb, rtol=1e-5, atol=1e-6)
    # the reference never resizes to a fixed network input: other resolutions go through the
    # adaptive average pool (128x128 sources -> 4x4 feature map -> 7x7)
    x2 = torch.randn(1, 3, 128, 128, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        assert torch.allclose(theirs(x2), ours(x2), rtol=1e-5, atol=1e-6)


def test_training_step_matches_reference_trainer(pair, ref):
    """One pass of the reference's Trainer.__train and of ours over the same two batches with the
    same RNG state: same loss / accuracy read-outs and the same updated weights."""
    import copy

    from distributed_vgg_f_b200.trainer import Trainer

    theirs, ours = copy.deepcopy(pair[0]), copy.deepcopy(pair[1])
    g = torch.Generator().manual_seed(3)
    batches = [(torch.randn(2, 3, 64, 64, generator=g), torch.randint(0, 3, (2,), generator=g)) for _ in range(2)]
    dev = torch.device("cpu")
    opt_a = torch.optim.Adam(theirs.parameters(), lr=1e-4)
    opt_b = torch.optim.Adam(ours.parameters(), lr=1e-4)
    ta = ref.Trainer(theirs, opt_a, batches, [], dev)
    tb = Trainer(ours, opt_b, batches, [], dev, verbose_throughput=False)
    torch.manual_seed(11)                                  # dropout masks
    loss_a, acc_a = ta._Trainer__train()
    torch.manual_seed(11)
    loss_b, acc_b = tb._train()
    assert str(loss_a) == str(loss_b) and str(acc_a) == str(acc_b)
    for (n, p), (_, q) in zip(theirs.named_parameters(), ours.named_parameters()):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-7), n


def test_checkpoint_loads_into_reference_model(pair, ref, tmp_path):
    from distributed_vgg_f_b200.utils import checkpoint as ck

    theirs, ours = pair
    path = str(tmp_path / "ck.pt")
    ck.save_checkpoint(path, ours, None, epoch=1, args={})
    payload = torch.load(path, weights_only=False)
    assert all(k.startswith("module.") for k in payload["model"])       # what a DDP-wrapped model saves
    fresh = ref.vgg_funnel_model(3)
    fresh.load_state_dict(ck.strip_module_prefix(payload["model"]))
    for (n, p), (_, q) in zip(fresh.named_parameters(), ours.named_parameters()):
        assert torch.equal(p, q), n
    # and the DDP-style keys load into a DataParallel-wrapped reference model as they are
    wrapped = torch.nn.DataParallel(ref.vgg_funnel_model(3))
    wrapped.load_state_dict(payload["model"])
