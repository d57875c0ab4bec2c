"""CPU-tier unit tests: CLI parity, metric arithmetic and formats, sampler
arithmetic vs torch, ImageFolder discovery, model topology / parameter count / state-dict keys,
bucket plan, transforms, checkpoint layout."""
import math
import os

import pytest
import torch

import distributed_vgg_f_b200 as pkg
from distributed_vgg_f_b200 import cli
from distributed_vgg_f_b200.data import transforms as T
from distributed_vgg_f_b200.data.folder import scan_image_folder
from distributed_vgg_f_b200.data.loader import DataManager, FusedBatch
from distributed_vgg_f_b200.data.sampler import ShardedSampler, shard_indices
from distributed_vgg_f_b200.models.vggf import build_oracle, vgg16_spec, vggf_spec
from distributed_vgg_f_b200.parallel.buckets import make_bucket_plan
from distributed_vgg_f_b200.utils.metrics import Accuracy2, Average, DeviceMeter


# ------------------------------------------------------------------------------------------- CLI
def test_cli_flags_and_defaults(capsys):
    args = cli.parse_command_line(["-iu", "tcp://127.0.0.1:1", "-rn", "0", "-ws", "1", "-rd", "/x"],
                                  init=False)
    assert (args.epochs, args.learning_rate, args.mini_batch, args.no_cuda) == (20, 0.001, 16, False)
    assert args.init_url == "tcp://127.0.0.1:1" and args.rank == 0 and args.world_size == 1
    assert "Namespace(" in capsys.readouterr().out          # the reference prints the namespace
    args = cli.parse_command_line(["--init_url", "tcp://h:2", "--rank", "1", "--world_size", "2",
                                   "--root_dir", "/d", "--epochs", "5", "--no_cuda",
                                   "--learning_rate", "0.00001", "--mini_batch", "64"], init=False)
    assert (args.epochs, args.learning_rate, args.mini_batch, args.no_cuda) == (5, 1e-5, 64, True)


def test_cli_requires_root_dir():
    with pytest.raises(SystemExit):
        cli.parse_command_line(["-iu", "tcp://127.0.0.1:1", "-rn", "0", "-ws", "1"], init=False)
    with pytest.raises(SystemExit):
        cli.parse_command_line(["-rn", "0", "-ws", "1", "-rd", "/x"], init=False)


# --------------------------------------------------------------------------------------- metrics
def test_average_and_accuracy_formats():
    a = Average()
    a.update(2.0, 4)
    a.update(1.0, 4)
    assert a.average == 1.5 and str(a) == "1.500000"
    acc = Accuracy2()
    out = torch.tensor([[0.1, 0.9], [0.8, 0.2], [0.3, 0.7]])
    acc.update(out, torch.tensor([1, 0, 0]))
    assert acc.correct == 2 and acc.count == 3 and str(acc) == "66.67%"


def test_device_meter_matches_host_metrics():
    torch.manual_seed(0)
    logits, tgt = torch.randn(10, 3), torch.randint(0, 3, (10,))
    m = DeviceMeter("cpu")
    m.add_reference(logits, tgt)
    avg, acc = m.snapshot()
    assert abs(avg.average - float(torch.nn.functional.cross_entropy(logits, tgt))) < 1e-6
    assert acc.correct == int((logits.argmax(1) == tgt).sum())


# --------------------------------------------------------------------------------------- sampler
@pytest.mark.parametrize("n,ws", [(24, 2), (25, 2), (7, 4), (5760, 8), (3, 4)])
def test_sampler_matches_torch_distributed_sampler(n, ws):
    from torch.utils.data import DistributedSampler

    ds = list(range(n))
    for epoch in (0, 3):
        for r in range(ws):
            ref = DistributedSampler(ds, num_replicas=ws, rank=r, shuffle=True, seed=0)
            ref.set_epoch(epoch)
            assert shard_indices(n, ws, r, epoch=epoch) == list(iter(ref))
    s = ShardedSampler(n, ws, 0, reference_order=True)
    s.set_epoch(5)
    assert list(iter(s)) == shard_indices(n, ws, 0, epoch=0)     # the reference never reshuffles


# ------------------------------------------------------------------------------------------ data
def test_imagefolder_contract(synth_root):
    classes, samples = scan_image_folder(os.path.join(synth_root, "TrainData"))
    assert classes == ["edible", "other", "toy"] and len(samples) == 24
    from torchvision.datasets import ImageFolder

    tv = ImageFolder(os.path.join(synth_root, "TrainData"))
    assert tv.classes == classes and tv.samples == samples


def test_datamanager_surface_and_batches(synth_root):
    dm = DataManager(synth_root, 5, train=True)
    assert dm.number_classes == 3 and dm.data_size == 24 and dm.class_names == ["edible", "other", "toy"]
    batches = list(dm.get_loader())
    assert [len(b.labels) for b in batches] == [5, 5, 5, 5, 4]      # ragged tail kept
    assert isinstance(batches[0], FusedBatch) and batches[0].images_u8.shape == (5, 128, 128, 3)
    x, y = batches[0].to_float()
    assert x.shape == (5, 3, 224, 224) and x.dtype == torch.float32 and y.dtype == torch.int64
    # sharded train / unsharded val (distributedVggf.py:115)
    d0 = DataManager(synth_root, 4, train=True, world_size=2, rank=0)
    d1 = DataManager(synth_root, 4, train=True, world_size=2, rank=1)
    assert len(d0.get_loader()) == 3 and len(d1.get_loader()) == 3
    v0 = DataManager(synth_root, 4, train=False, world_size=2, rank=0)
    assert sum(len(b.labels) for b in v0.get_loader()) == 12


def test_reference_pipeline_shapes(synth_root):
    dm = DataManager(synth_root, 4, train=True, pipeline="reference")
    x, y = next(iter(dm.get_loader()))
    assert x.shape == (4, 3, 224, 224) and y.shape == (4,)


def test_val_transform_matches_torchvision(synth_root):
    """Resize(256)+CenterCrop(224)+Normalize: fused evaluation vs PIL, up to PIL's uint8 rounding."""
    from PIL import Image

    _, samples = scan_image_folder(os.path.join(synth_root, "ValidationData"))
    img = Image.open(samples[0][0]).convert("RGB")
    ref = T.reference_transforms(train=False)(img)
    import numpy as np
    src = torch.from_numpy(np.asarray(img)).unsqueeze(0)
    got = T.augment_reference(src, T.val_params(1, 128, 128), (256, 256))[0]
    assert float((got - ref).abs().max()) < 0.06     # 1-2 uint8 steps / std
    assert float((got - ref).abs().mean()) < 0.01


def test_train_params_distribution():
    g = torch.Generator().manual_seed(0)
    p = T.sample_train_params(500, 128, 128, g)
    area = p[:, 2] * p[:, 3] / (128 * 128)
    assert float(area.min()) > 0.75 and float(area.max()) <= 1.0
    ang = torch.atan2(p[:, 5], p[:, 4]).abs() * 180 / math.pi
    assert float(ang.max()) <= 10.0 + 1e-4
    assert 0.4 < float(p[:, 6].mean()) < 0.6
    assert bool(((p[:, 0] + p[:, 2]) <= 128).all() and ((p[:, 1] + p[:, 3]) <= 128).all())


# ----------------------------------------------------------------------------------------- model
def test_vggf_topology_and_param_count():
    spec = vggf_spec(3)
    assert spec.num_params == 136_359_747
    assert vgg16_spec(1000).num_params == 138_357_544
    assert len(spec.param_names) == 34
    assert abs(spec.flops_per_image(224) / 1e9 - 30.94) < 0.05
    assert abs(spec.flops_per_image(128) / 1e9 - 10.27) < 0.35   # avgpool upsample not counted


def test_oracle_state_dict_keys_match_torchvision():
    from torch import nn
    from torchvision import models

    tv = models.vgg16(weights=None)
    tv.classifier[6] = nn.Sequential(nn.Linear(4096, 512), nn.ReLU(inplace=True), nn.Dropout(0.6),
                                     nn.Linear(512, 3))        # distributedVggf.py:52-57
    with torch.device("meta"):
        ours = pkg.models.vggf.VGGOracle.__new__(pkg.models.vggf.VGGOracle)
    ours = build_oracle(vggf_spec(3), seed=0)
    ref_sd, our_sd = tv.state_dict(), ours.state_dict()
    assert list(ref_sd.keys()) == list(our_sd.keys())
    assert all(ref_sd[k].shape == our_sd[k].shape for k in ref_sd)
    tv.load_state_dict(our_sd)                                   # interoperable
    x = torch.randn(1, 3, 64, 64)
    tv.eval(), ours.eval()
    assert torch.allclose(tv(x), ours(x), atol=1e-5)


def test_vgg_funnel_model_factory_is_dropin():
    m = pkg.vgg_funnel_model(5, seed=0)
    assert m.classifier[6][3].out_features == 5 and m.classifier[6][2].p == 0.6


# --------------------------------------------------------------------------------------- buckets
def test_bucket_plan_covers_arena_and_splits_big_tensors():
    spec = vggf_spec(3)
    ready = [(n, math.prod(spec.param_shape(n))) for n in reversed(spec.param_names)]
    plan = make_bucket_plan(ready, cap_elems=8 * 1024 * 1024)
    assert plan.buckets[0].start == 0 and plan.buckets[-1].end == plan.total
    for a, b in zip(plan.buckets, plan.buckets[1:]):
        assert a.end == b.start
    assert all(b.start % plan.align == 0 and b.end % plan.align == 0 for b in plan.buckets)
    assert max(b.numel for b in plan.buckets) <= 8 * 1024 * 1024
    big = plan.bucket_of("classifier.0.weight")
    assert len(big) == math.ceil(102_760_448 / (8 * 1024 * 1024))     # split, unlike DDP
    assert plan.order[0] == "classifier.6.3.bias" and plan.order[-1] == "features.0.weight"
    # every tensor lies inside the union of its buckets
    for n in plan.order:
        ids = plan.bucket_of(n)
        assert plan.buckets[ids[0]].start <= plan.offsets[n]
        assert plan.buckets[ids[-1]].end >= plan.offsets[n] + plan.numels[n]


def test_bucket_plan_late_caps_and_tail_bucket():
    """The engine's plan: big messages for the FC weights that are ready first, one bucket per big conv
    layer afterwards, and a small final bucket (its reduction + update is the only exposed part of the
    gradient exchange)."""
    from distributed_vgg_f_b200.models import layout as L

    spec = vggf_spec(3)
    order = L.ready_order(spec)
    first_conv = next(n for n, _ in order if n.startswith("features."))
    cap = 8 * 1024 * 1024
    plan = make_bucket_plan(order, cap_elems=cap, late_cap_elems=int(9.5 * 1024 * 1024 / 4), late_from=first_conv,
                            tail_elems=int(2400 * 1024 / 4))
    assert plan.buckets[0].start == 0 and plan.buckets[-1].end == plan.total
    assert all(a.end == b.start for a, b in zip(plan.buckets, plan.buckets[1:]))
    assert all(b.start % plan.align == 0 and b.end % plan.align == 0 for b in plan.buckets)
    conv_b = [b for b in plan.buckets if all(t.startswith("features.") for t in b.tensors)]
    fc_b = [b for b in plan.buckets if b not in conv_b]
    assert max(b.numel for b in fc_b) == cap                                   # 16.8 MB of bf16 wire
    assert max(b.numel for b in conv_b) <= int(9.5 * 1024 * 1024 / 4) + plan.align
    tail = plan.buckets[-1]
    assert tail.tensors[-1] == "features.0.weight" and tail.numel * 2 < 1.3e6   # ~1.1 MB of bf16 wire
    assert "features.10.weight" in tail.tensors and "features.12.weight" not in tail.tensors
    # the five 512-channel layers' weights each sit in a bucket of their own size class (one layer per bucket)
    for name in ("features.28.weight", "features.26.weight", "features.24.weight", "features.21.weight", "features.19.weight"):
        (bi,) = plan.bucket_of(name)
        assert sum(t.endswith(".weight") for t in plan.buckets[bi].tensors) == 1
    # default arguments reproduce the plain capped plan
    plain = make_bucket_plan(order, cap_elems=cap)
    assert [(b.start, b.end) for b in plain.buckets] == \
        [(b.start, b.end) for b in make_bucket_plan(order, cap_elems=cap, late_cap_elems=0, tail_elems=0).buckets]


# ------------------------------------------------------------------------------------ checkpoint
def test_checkpoint_layout_roundtrip(tmp_path):
    from distributed_vgg_f_b200.utils import checkpoint as ck

    m = build_oracle(vggf_spec(3), seed=1)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    path = str(tmp_path / "c.pt")
    ck.save_checkpoint(path, m, opt, epoch=2, args={"mini_batch": 4})
    payload = torch.load(path, weights_only=False)
    assert payload["format"] == ck.FORMAT and payload["epoch"] == 2
    keys = list(payload["model"].keys())
    assert len(keys) == 34 and keys[0] == "module.features.0.weight" and keys[-1] == "module.classifier.6.3.bias"
    m2 = build_oracle(vggf_spec(3), seed=2)
    assert ck.load_checkpoint(path, m2, None) == 2
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))


def test_native_png_decoder_matches_pil(synth_root, tmp_path):
    """csrc/png_decode.cpp (zlib inflate + PNG filters) == PIL, incl. RGBA / grey / palette inputs."""
    import numpy as np
    from PIL import Image

    from distributed_vgg_f_b200 import ops

    if not ops.available():
        pytest.skip("native extension not built")
    _, samples = scan_image_folder(os.path.join(synth_root, "TrainData"))
    paths = [p for p, _ in samples][:6]
    got = ops.native_decode_pngs(paths, 4)
    assert got is not None and got.shape == (6, 128, 128, 3)
    for i, p in enumerate(paths):
        assert np.array_equal(got[i].numpy(), np.asarray(Image.open(p).convert("RGB")))
    rng = np.random.default_rng(0)
    base = (rng.random((40, 52, 4)) * 255).astype(np.uint8)
    variants = {"rgba": Image.fromarray(base, "RGBA"), "grey": Image.fromarray(base[..., 0], "L"),
                "pal": Image.fromarray(base[..., :3], "RGB").quantize(64)}
    for name, im in variants.items():
        f = str(tmp_path / (name + ".png"))
        im.save(f)
        dec = ops.native_decode_pngs([f], 1)
        assert dec is not None, name
        assert np.array_equal(dec[0].numpy(), np.asarray(Image.open(f).convert("RGB"))), name
    assert ops.native_decode_pngs([str(tmp_path / "missing.png")], 1) is None
    f4 = str(tmp_path / "pal4.png")
    Image.fromarray(base[..., :3], "RGB").quantize(16).save(f4)      # 4-bit palette: not handled natively
    assert ops.native_decode_pngs([f4], 1) is None                   # -> caller falls back to PIL


def test_native_prefetcher_matches_python_loader(synth_root):
    """csrc/prefetch.cpp worker: same batches (images, labels, sizes) as the Python producer thread;
    transform parameters follow torchvision's ranges."""
    from distributed_vgg_f_b200 import ops
    from distributed_vgg_f_b200.data.loader import DataManager

    if not ops.available():
        pytest.skip("native extension not built")
    dm = DataManager(synth_root, 5, train=True, seed=3)
    ld = dm.get_loader()
    assert ld._pf is not None
    ld.set_epoch(1)
    nat = [(b.images_u8.clone(), b.labels.clone(), b.params.clone()) for b in ld]
    ld._pf, keep = None, ld._pf
    ld.set_epoch(1)
    py = [(b.images_u8.clone(), b.labels.clone(), b.params.clone()) for b in ld]
    ld._pf = keep
    assert [len(a[1]) for a in nat] == [5, 5, 5, 5, 4] == [len(a[1]) for a in py]
    for a, b in zip(nat, py):
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    p = torch.cat([a[2] for a in nat])
    area = p[:, 2] * p[:, 3] / (128 * 128)
    assert float(area.min()) > 0.75 and float(area.max()) <= 1.0
    assert bool(((p[:, 0] + p[:, 2]) <= 128).all() and ((p[:, 1] + p[:, 3]) <= 128).all())
    assert float(torch.atan2(p[:, 5], p[:, 4]).abs().max()) <= math.radians(10.0) + 1e-5
    # two epochs in a row and an abandoned iteration must not dead-lock the worker
    it = iter(ld)
    next(it)
    del it
    assert sum(len(b.labels) for b in ld) == 24


def test_native_layout_roundtrip_and_semantics():
    """models.layout: torch <-> native conversions are inverse, and the native layouts mean what the
    kernels assume (OHWI conv weights, im2col-ordered first conv, NHWC-ordered classifier.0 columns)."""
    from distributed_vgg_f_b200.models import layout as L
    from distributed_vgg_f_b200.models.vggf import vggf_mini_spec

    spec = vggf_mini_spec(3)
    torch.manual_seed(0)
    for name in spec.param_names:
        t = torch.randn(spec.param_shape(name))
        n = L.to_native(spec, name, t)
        assert tuple(n.shape) == L.native_shape(spec, name), name
        assert torch.equal(L.to_torch(spec, name, n), t), name
    # first conv: y = im2col(x) @ W0^T with k = (kh*3+kw)*3 + c
    w = torch.randn(64, 3, 3, 3)
    x = torch.randn(1, 3, 6, 6)
    w0 = L.to_native(spec, "features.0.weight", w)
    cols = torch.nn.functional.unfold(x, 3, padding=1).view(1, 3, 9, 36).permute(0, 3, 2, 1).reshape(36, 27)
    got = (cols @ w0[:, :27].t()).t().reshape(1, 64, 6, 6)
    assert torch.allclose(got, torch.nn.functional.conv2d(x, w, padding=1), atol=1e-4)
    assert torch.count_nonzero(w0[:, 27:]) == 0
    # classifier.0: NHWC flatten of the feature map times the permuted weight == NCHW flatten times the original
    f = spec.fcs[0]
    wfc = torch.randn(f.fout, f.fin)
    feat = torch.randn(2, f.fin // 49, 7, 7)
    a = torch.flatten(feat, 1) @ wfc.t()
    b = feat.permute(0, 2, 3, 1).reshape(2, -1) @ L.to_native(spec, f.name + ".weight", wfc).t()
    assert torch.allclose(a, b, atol=1e-3)


def test_emulated_reference_tracks_autograd():
    """ops.ref.emulated_step (the engine's oracle: torch fp32 ops with the engine's bf16 rounding
    points) stays close to plain autograd on a small network."""
    import torch.nn.functional as F

    from distributed_vgg_f_b200.models.vggf import vggf_tiny_spec
    from distributed_vgg_f_b200.ops import ref as R

    spec = vggf_tiny_spec(3)
    model = build_oracle(spec, seed=0).eval()
    torch.manual_seed(1)
    x = torch.randn(2, 3, 32, 32)
    y = torch.tensor([0, 2])
    loss = F.cross_entropy(model(x), y)
    loss.backward()
    state = {k: v.detach() for k, v in model.state_dict().items()}
    logits, eloss, grads = R.emulated_step(spec, state, x, y)
    assert abs(float(eloss) - float(loss)) < 5e-2
    for name, p in model.named_parameters():
        a, b = grads[name].flatten(), p.grad.flatten()
        cos = float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30))
        assert cos > 0.9, (name, cos)       # bf16 rounding drift grows towards the input layers


def test_coil100_regrouping(tmp_path):
    """tools/coil100: the reference's hand-made 3-class split of COIL-100 (Readme.md:81-107)."""
    from PIL import Image

    from distributed_vgg_f_b200.data.folder import scan_image_folder
    from distributed_vgg_f_b200.tools import coil100

    src = tmp_path / "coil-100"
    src.mkdir()
    for obj in (2, 6, 1, 100, 47, 99):                  # edible, toy, other, toy, edible, other
        for angle in range(0, 50, 5):                   # 10 views each
            Image.new("RGB", (8, 8), (obj, angle, 0)).save(src / ("obj%d__%d.png" % (obj, angle)))
    (src / "readme.txt").write_text("not an image")
    assert len(coil100.EDIBLE) == 18 and len(coil100.TOY) == 20           # 1296 / 1440 images of 72 views
    assert (len(coil100.EDIBLE) * 72, len(coil100.TOY) * 72, (100 - 38) * 72) == (1296, 1440, 4464)
    rep = coil100.prepare(str(src), str(tmp_path / "coil3"), val_fraction=0.2, seed=0)
    assert rep == {c: {"train": 16, "val": 4} for c in ("edible", "other", "toy")}
    classes, samples = scan_image_folder(str(tmp_path / "coil3" / "TrainData"))
    assert classes == ["edible", "other", "toy"] and len(samples) == 48
    _, val = scan_image_folder(str(tmp_path / "coil3" / "ValidationData"))
    train_names = {os.path.basename(p) for p, _ in samples}
    assert not train_names & {os.path.basename(p) for p, _ in val}      # disjoint split
    assert all(os.path.basename(p).startswith(("obj2__", "obj47__")) for p, y in samples if y == 0)
    # idempotent, and the weights follow 1/count
    assert coil100.prepare(str(src), str(tmp_path / "coil3"), val_fraction=0.2, seed=0) == rep
    w = coil100.inverse_frequency_weights({"edible": 1037, "other": 3571, "toy": 1152})
    assert abs(sum(w) - 1.0) < 1e-3 and w[1] < w[2] < w[0]
    assert coil100.main([str(src), str(tmp_path / "c2"), "--copy"]) == 0
    assert not os.path.islink(str(tmp_path / "c2" / "TrainData" / "toy" / os.listdir(str(tmp_path / "c2" / "TrainData" / "toy"))[0]))


def test_auto_pipeline_falls_back_for_mixed_sizes_and_big_splits(synth_root, tmp_path, monkeypatch, capsys):
    """pipeline='auto': mixed image sizes or a split larger than the cache budget use the per-sample
    pipeline (which is what the reference always does); pipeline='fused' reports the reason."""
    import shutil

    from PIL import Image

    from distributed_vgg_f_b200.data.loader import CacheUnavailable, DataManager, FusedBatch

    dm = DataManager(synth_root, 4, train=True, pipeline="auto")
    assert dm.pipeline == "fused" and isinstance(next(iter(dm.get_loader())), FusedBatch)
    monkeypatch.setenv("B200_MAX_CACHE_GB", "0.00001")
    dm = DataManager(synth_root, 4, train=True, pipeline="auto")
    assert dm.pipeline == "reference" and "B200_MAX_CACHE_GB" in capsys.readouterr().out
    x, y = next(iter(dm.get_loader()))
    assert tuple(x.shape) == (4, 3, 224, 224) and x.dtype == torch.float32 and y.dtype == torch.int64
    with pytest.raises(CacheUnavailable):
        DataManager(synth_root, 4, train=True, pipeline="fused")
    monkeypatch.delenv("B200_MAX_CACHE_GB")
    mixed = str(tmp_path / "mixed")
    shutil.copytree(synth_root, mixed)
    cls_dir = os.path.join(mixed, "TrainData", "toy")
    Image.new("RGB", (96, 160), (10, 200, 30)).save(os.path.join(cls_dir, "odd_size.png"))
    dm = DataManager(mixed, 4, train=True, pipeline="auto")
    assert dm.pipeline == "reference" and dm.data_size == len(dm.samples)
    assert sum(int(yb.numel()) for _, yb in dm.get_loader()) == dm.data_size


def test_weighted_loss_readout_matches_torch():
    """With --class-weights the epoch loss is the weighted mean torch reports (what the reference's
    commented-out variant would feed to Average.update), evaluation stays unweighted."""
    import torch.nn.functional as F

    from distributed_vgg_f_b200.utils.metrics import DeviceMeter

    g = torch.Generator().manual_seed(0)
    w = torch.tensor([0.41, 0.19, 0.4])                   # distributedUtil.py:28
    m, expect, n = DeviceMeter("cpu"), 0.0, 0
    for b in (5, 3):
        logits, y = torch.randn(b, 3, generator=g), torch.randint(0, 3, (b,), generator=g)
        m.add_reference(logits, y, w)
        expect += float(F.cross_entropy(logits, y, weight=w)) * b
        n += b
    avg, acc = m.snapshot()
    assert abs(avg.sum - expect) < 1e-5 and avg.count == n and acc.count == n


def test_native_png_decoder_rejects_corrupt_files(tmp_path):
    """Truncated / corrupt / non-PNG files must come back as "not handled" (None), never crash."""
    from PIL import Image

    from distributed_vgg_f_b200 import ops

    if not ops.available():
        pytest.skip("native extension not built")
    good = tmp_path / "good.png"
    Image.new("RGB", (16, 16), (1, 2, 3)).save(good)
    raw = good.read_bytes()
    cases = {"truncated.png": raw[:40], "header_only.png": raw[:8], "garbage.png": b"not a png at all" * 4,
             "short_ihdr.png": raw[:8] + b"\x00\x00\x00\x00IHDR" + b"\x00" * 4,
             "huge_dims.png": raw[:16] + b"\x7f\xff\xff\xff\x7f\xff\xff\xff" + raw[24:],
             "bad_zlib.png": raw[:60] + bytes(len(raw) - 60)}
    paths = []
    for name, data in cases.items():
        (tmp_path / name).write_bytes(data)
        paths.append(str(tmp_path / name))
    paths.append(str(tmp_path / "missing.png"))
    assert ops.native_decode_pngs(paths, 2) is None                 # nothing usable -> caller falls back
    ok = ops.native_decode_pngs([str(good)], 1)
    assert ok is not None and tuple(ok.shape) == (1, 16, 16, 3) and int(ok[0, 0, 0, 2]) == 3


def test_perf_line_reports_utilisation_for_native_engines():
    from distributed_vgg_f_b200.models.vggf import vggf_spec
    from distributed_vgg_f_b200.trainer import Trainer, _bf16_peak

    spec = vggf_spec(3)
    assert abs(spec.flops_per_image(224) / 1e9 - 30.94) < 0.05

    class FakeEngine:                       # quacks like NativeEngine for Trainer
        HW = 224

        def __init__(self):
            self.spec = spec

        def train_step(self, batch):
            pass

        def eval_step(self, batch):
            pass

    tr = Trainer(FakeEngine(), None, [], [], torch.device("cpu"), verbose_throughput=True)
    text = tr._utilisation(images=9340, seconds=1.0)                    # the round-1 single-GPU rate
    assert "TFLOP/s" in text and "% of the measured bf16 peak" in text
    tflops = float(text.split(",")[1].split()[0])
    assert abs(tflops - 3 * 30.94e9 * 9340 / 1e12) < 1.0 and 1000 < _bf16_peak() < 2500
    assert Trainer(torch.nn.Linear(2, 2), None, [], [], torch.device("cpu"))._utilisation(10, 1.0) == ""


@pytest.mark.parametrize("depth,params", [(11, 132_863_336), (13, 133_047_848), (19, 143_667_240)])
def test_other_vgg_depths_match_torchvision(depth, params):
    """VGG-11/13/19 from the same layer table: torchvision's parameter count, state-dict keys and
    forward (the native engine executes the table; the oracle is checked here)."""
    import torchvision

    from distributed_vgg_f_b200.models.vggf import build_oracle, get_spec, vgg_spec

    spec = vgg_spec(depth, 1000)
    assert spec.num_params == params
    assert get_spec("vgg%d" % depth, 1000).num_params == params
    funnel = get_spec("vggf%d" % depth, 3)
    assert [f.name for f in funnel.fcs][-2:] == ["classifier.6.0", "classifier.6.3"] and funnel.fcs[-1].fout == 3
    if depth != 11:
        return                                             # one full forward is enough for CI time
    tv = getattr(torchvision.models, "vgg%d" % depth)(weights=None).eval()
    ours = build_oracle(spec, seed=0).eval()
    assert list(tv.state_dict().keys()) == list(ours.state_dict().keys())
    ours.load_state_dict(tv.state_dict())
    x = torch.randn(1, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        assert torch.allclose(tv(x), ours(x), rtol=1e-5, atol=1e-6)


def test_eval_only_scores_a_checkpoint(synth_root, tmp_path, capsys):
    """--save then --resume --eval-only reproduces the validation line of the training run."""
    import re

    from distributed_vgg_f_b200.cli import parse_command_line
    from distributed_vgg_f_b200.train import manage_training

    ck = str(tmp_path / "ck.pt")
    base = ["-iu", "tcp://127.0.0.1:1", "-rn", "0", "-ws", "1", "-rd", synth_root, "-nc", "-mb", "4",
            "--model", "vggf-tiny", "--engine", "oracle"]
    manage_training(parse_command_line(base + ["-ep", "1", "--save", ck]))
    trained = re.search(r"test loss: ([\d.]+), test acc: ([\d.]+)%", capsys.readouterr().out).groups()
    manage_training(parse_command_line(base + ["--resume", ck, "--eval-only"]))
    out = capsys.readouterr().out
    assert "[Info] Evaluation:" in out and "Epoch:" not in out.split("[Info] Evaluation:")[1]
    assert re.search(r"Evaluation: test loss: ([\d.]+), test acc: ([\d.]+)%", out).groups() == trained


def test_predict_tool_reproduces_validation_accuracy(synth_root, tmp_path, capsys):
    """tools/predict on the validation folder gives the accuracy the training run printed."""
    import re

    from distributed_vgg_f_b200.cli import parse_command_line
    from distributed_vgg_f_b200.tools import predict as P
    from distributed_vgg_f_b200.train import manage_training

    ck = str(tmp_path / "ck.pt")
    manage_training(parse_command_line(["-iu", "tcp://127.0.0.1:1", "-rn", "0", "-ws", "1", "-rd", synth_root, "-nc",
                                        "-mb", "4", "-ep", "1", "--model", "vggf-tiny", "--engine", "oracle",
                                        "--pipeline", "reference", "--save", ck]))
    acc = float(re.search(r"test acc: ([\d.]+)%", capsys.readouterr().out).group(1))
    val = os.path.join(synth_root, "ValidationData")
    logits, files = P.predict(ck, [val], engine="oracle")
    classes = sorted(os.listdir(val))
    truth = torch.tensor([classes.index(os.path.basename(os.path.dirname(f))) for f in files])
    assert logits.shape == (len(files), 3)
    assert abs(100.0 * float((logits.argmax(1) == truth).float().mean()) - acc) < 0.01      # printed with 2 decimals
    assert P.main([ck, files[0], "--classes", "edible,other,toy", "--engine", "oracle"]) == 0
    line = capsys.readouterr().out.strip().split("\t")
    assert line[0] == files[0] and line[1] in ("edible", "other", "toy") and 0.0 < float(line[2]) <= 1.0


# ------------------------------------------------------------------ parity with torchvision itself
def test_train_augment_matches_torchvision_chain_on_fixed_draws():
    """The fused train transform (one coordinate chain per output pixel) against torchvision's OWN ops
    applied one after the other on a PIL image with the same drawn parameters (distributedVggf.py:88-95):
    RandomResizedCrop(256, scale=(.8, 1)) -> RandomRotation(10) -> RandomHorizontalFlip -> CenterCrop(224)
    -> ToTensor -> Normalize.  PIL rounds the resized image to uint8 before rotating (<= 0.5/255 per pixel,
    0.009 after Normalize) and nearest-neighbour rotation can pick the neighbouring pixel when a coordinate
    lands within float rounding of a pixel edge, hence "almost all pixels within 0.03"."""
    import numpy as np

    from distributed_vgg_f_b200.data import transforms as T
    from distributed_vgg_f_b200.data.synthetic import synthetic_uint8_batch

    imgs, _ = synthetic_uint8_batch(6, 128, 3, seed=11)
    rng = np.random.default_rng(0)
    imgs = np.clip(imgs.astype(np.int16) + rng.integers(-20, 20, imgs.shape), 0, 255).astype(np.uint8)   # texture
    params = T.sample_train_params(6, 128, 128, torch.Generator().manual_seed(5))
    # make sure both flip states and both rotation signs are exercised
    params[0, 6], params[1, 6] = 0.0, 1.0
    params[2, 4], params[2, 5] = math.cos(math.radians(9.5)), math.sin(math.radians(9.5))
    params[3, 4], params[3, 5] = math.cos(math.radians(-9.5)), math.sin(math.radians(-9.5))
    ours = T.augment_reference(torch.from_numpy(imgs), params, (256, 256))
    for i in range(6):
        theirs = T.torchvision_chain_fixed(imgs[i], params[i])
        d = (ours[i] - theirs).abs()
        frac_close = float((d < 0.03).float().mean())
        assert frac_close > 0.995, (i, frac_close, float(d.max()))
        assert float(d.mean()) < 0.008, (i, float(d.mean()))   # 0.25/255/std from PIL's uint8 rounding alone is 0.0043


def test_pretrained_vgg16_state_is_loaded_like_the_reference(tmp_path):
    """--pretrained (offline stand-in for models.vgg16(pretrained=True), distributedVggf.py:46): features.*
    and classifier.0 / .3 come from the torchvision VGG-16 file, the funnel (classifier.6.0 / 6.3,
    distributedVggf.py:52-57) keeps its fresh init -- exactly what the reference's factory produces."""
    import torchvision

    from distributed_vgg_f_b200 import vgg_funnel_model

    tv = torchvision.models.vgg16(weights=None)
    path = str(tmp_path / "vgg16.pth")
    torch.save(tv.state_dict(), path)
    fresh = vgg_funnel_model(3, seed=1).state_dict()
    got = vgg_funnel_model(3, pretrained_path=path, seed=1).state_dict()
    tvs = tv.state_dict()
    for k, v in got.items():
        if k.startswith("features.") or k.startswith("classifier.0.") or k.startswith("classifier.3."):
            assert torch.equal(v, tvs[k]), k
        else:                                    # the funnel head is not in the file
            assert k.startswith("classifier.6.") and torch.equal(v, fresh[k]), k
    assert "classifier.6.weight" in tvs and "classifier.6.weight" not in got


def test_resume_behind_an_lr_step_boundary_keeps_the_decayed_rate(synth_root, tmp_path):
    """--lr-step decays the rate at epochs that are multiples of the step.  A run resumed BEHIND such a boundary
    must continue at the decayed rate on both back ends (round-1 ADVICE: the native engine restarted at the base
    rate; manage_training now derives it from the resumed epoch, and the engine restores its own ``lr`` too)."""
    from distributed_vgg_f_b200.cli import parse_command_line
    from distributed_vgg_f_b200.train import manage_training

    ck = str(tmp_path / "ck.pt")
    base = ["-iu", "tcp://127.0.0.1:1", "-rn", "0", "-ws", "1", "-rd", synth_root, "-nc", "-mb", "4", "-lr", "0.01",
            "--model", "vggf-tiny", "--engine", "oracle", "--lr-step", "1", "--lr-gamma", "0.5"]
    first = manage_training(parse_command_line(base + ["-ep", "2", "--save", ck]))
    assert abs(first.optimizer.param_groups[0]["lr"] - 0.01 * 0.5 ** 2) < 1e-12
    # resume at epoch 3 of 2: nothing left to train, the optimizer shows the rate epoch 3 WOULD run at
    resumed = manage_training(parse_command_line(base + ["-ep", "2", "--resume", ck]))
    assert abs(resumed.optimizer.param_groups[0]["lr"] - 0.01 * 0.5 ** 2) < 1e-12
    # a checkpoint whose stored rate was lost (older file) still resumes at the right rate
    payload = torch.load(ck, map_location="cpu", weights_only=True)
    for g in payload["optimizer"]["param_groups"]:
        g["lr"] = 0.01
    torch.save(payload, ck)
    resumed = manage_training(parse_command_line(base + ["-ep", "2", "--resume", ck]))
    assert abs(resumed.optimizer.param_groups[0]["lr"] - 0.01 * 0.5 ** 2) < 1e-12


def test_trainer_perf_lines_with_a_native_style_engine(capsys):
    """Trainer's observability for engines that own the whole step (NativeEngine's surface, stubbed on the CPU):
    [Perf] throughput + gradient all-reduce GB/s and NVLink fraction, host-side epoch timing, and
    --profile timeline (marks of the epoch's LAST training step only)."""
    from distributed_vgg_f_b200.trainer import Trainer
    from distributed_vgg_f_b200.utils.metrics import DeviceMeter

    class Engine:
        def __init__(self):
            self.meter, self.steps, self.tl_on, self.tl_steps = None, 0, False, []

        def set_meter(self, m):
            self.meter = m

        def train_step(self, batch):
            self.steps += 1
            if self.tl_on:
                self.tl_steps.append(self.steps)
            self.meter.buf += torch.tensor([0.5 * 4, 3.0, 4.0, 0.0])

        def eval_step(self, batch):
            self.meter.buf += torch.tensor([0.25 * 4, 4.0, 4.0, 0.0])

        def sync(self):
            pass

        def timeline(self, on):
            self.tl_on = on

        def timeline_report(self):
            return [("step start", "compute", 0.0), ("bwd features.0", "compute", 6.5), ("ar23 end", "comm", 6.6),
                    ("opt23 end", "opt", 6.7)]

        def comm_report(self, steps):
            return {"wire_MB_per_step": 272.8, "ms_per_step": 2.5, "bus_GBs": 190.0, "frac_of_770_measured": 0.247,
                    "frac_of_900_nominal": 0.211}

    eng = Engine()
    batches = [object()] * 5
    tr = Trainer(eng, None, batches, batches[:2], torch.device("cpu"))
    tr.profile_timeline = True
    tr.fit(2)
    out = capsys.readouterr().out
    assert out.count("[Info] Epoch:") == 2 and "train loss: 0.500000, train acc: 75.00%" in out
    assert "grad all-reduce 273 MB/step in 2.50 ms = 190 GB/s bus (25% of 770 measured, 21% of 900 nominal)" in out
    assert out.count("host: first batch after") == 2 and "validation pass" in out
    assert eng.tl_steps == [5, 10], eng.tl_steps                  # only the last step of each epoch was traced
    assert out.count("[Timeline] last training step") == 2 and "compute bwd features.0" in out and "opt     opt23 end" in out
    assert tr.epoch_allreduce["bus_GBs"] == 190.0 and tr.epoch_host_times["enqueue_ms"] >= 0.0


def test_plan_tool_and_engine_bucket_plan(capsys):
    """tools/plan prints what NativeEngine executes: the same plan object (engine_bucket_plan), the ZeRO-1 / all-reduce
    split by world size, NVLS vs the no-multicast fallback."""
    from distributed_vgg_f_b200.models import layout as L
    from distributed_vgg_f_b200.parallel.buckets import engine_bucket_plan
    from distributed_vgg_f_b200.tools import plan as P

    spec = vggf_spec(3)
    order = L.ready_order(spec)
    first_conv = next(n for n, _ in order if n.startswith("features."))
    a = engine_bucket_plan(order)
    b = make_bucket_plan(order, cap_elems=8 * 1024 * 1024, late_cap_elems=int(9.5 * 1024 * 1024 / 4),
                         late_from=first_conv, tail_elems=int(2400 * 1024 / 4))
    assert [(x.start, x.end, x.tensors) for x in a.buckets] == [(x.start, x.end, x.tensors) for x in b.buckets]
    _, rows8 = P.describe("vggf", 3, world=8)
    _, rows2 = P.describe("vggf", 3, world=2)
    _, rows8_nomc = P.describe("vggf", 3, world=8, multicast=False)
    z8 = [r for r in rows8 if r[6].startswith("zero1")]
    assert len(z8) == 15 and all(r[4] in ("classifier.0.weight", "classifier.3.weight") for r in z8)
    assert sum(r[1] for r in z8) / sum(r[1] for r in rows8) > 0.85           # 88 % of the elements
    assert not any(r[6].startswith("zero1") for r in rows2)                  # auto: from 4 ranks
    assert all("nvls" in r[6] for r in rows8) and not any("nvls" in r[6] for r in rows8_nomc)
    assert "oneshot" in [r for r in rows8_nomc if r[4] == "classifier.0.bias"][0][6]
    assert P.main(["--world", "4"]) == 0
    out = capsys.readouterr().out
    assert "24 buckets" in out and "272.8 MB of bf16 wire per step" in out and "features.0.weight" in out
